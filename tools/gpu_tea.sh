#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "tea or wide" > gpurun_out/tea_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/tea_pytest.log
tail -4 gpurun_out/tea_pytest.log
MADDY_GPU_PROFILE=1 python tools/quick_bench.py cylinder_tea 64 200 2>&1 | tail -12
MADDY_GPU_PROFILE=1 python tools/quick_bench.py cylinder_tea_large 1 200 2>&1 | tail -12
MADDY_GPU_PROFILE=1 python tools/quick_bench.py cylinder_tea 1 200 2>&1 | tail -6
