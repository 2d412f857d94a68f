#!/bin/bash
# quick iteration: parity tests + headline shapes (+ optional ncu capture with PROFILE=1)
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "--- 520x256"; python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -1
echo "--- 520x2048"; python tools/quick_bench.py mt40_ensemble 2048 400 2>&1 | grep "run 400" | tail -1
echo "--- 520x1"; python tools/quick_bench.py mt40_single 1 1000 2>&1 | grep "run 1000" | tail -1
echo "--- 1560x256"; python tools/quick_bench.py mt120_disassembly 256 400 hydrolysis=no 2>&1 | grep "run 400" | tail -1
if [ -n "$PROFILE" ]; then
  ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -o gpurun_out/prof_iter python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/ncu_iter.log 2>&1
  tail -2 gpurun_out/ncu_iter.log
fi
