#!/bin/bash
# one development iteration on the GPU box: parity suite + timings (+ optional ncu capture with PROFILE=1)
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | tail -5
python tools/quick_bench.py mt40_single 1 1000 2>&1 | tail -3
python tools/quick_bench.py mt120_disassembly 256 400 hydrolysis=no 2>&1 | tail -3
if [ -n "$PROFILE" ]; then
  ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -o gpurun_out/prof_iter python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/ncu_iter.log 2>&1
  tail -2 gpurun_out/ncu_iter.log
fi
