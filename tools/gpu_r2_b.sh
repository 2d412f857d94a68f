#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_events.py -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
python tools/cc_probe.py > gpurun_out/r2b_cc_probe.log 2>&1
tail -5 gpurun_out/r2b_pytest.log; tail -20 gpurun_out/r2b_cc_probe.log
