#!/bin/bash
# quick iteration: bitwise/parity tests of the fused loop + device timing at 256 and 2048 trajectories
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/quick_pytest.log
tail -3 gpurun_out/quick_pytest.log
python tools/quick_bench.py mt40_ensemble 256 1000 | tail -3
python tools/quick_bench.py mt40_ensemble 2048 1000 | tail -3
