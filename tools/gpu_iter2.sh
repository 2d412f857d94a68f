#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
echo "--- shape 1 (2 CTAs/SM)"; python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -1
echo "--- shape 0 (1 CTA/SM)"; MADDY_CTAS_PER_SM=1 python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -1
echo "--- 2048 traj shape1"; python tools/quick_bench.py mt40_ensemble 2048 400 2>&1 | grep "run 400" | tail -1
echo "--- 2048 traj shape0"; MADDY_CTAS_PER_SM=1 python tools/quick_bench.py mt40_ensemble 2048 400 2>&1 | grep "run 400" | tail -1
echo "--- 1 traj"; python tools/quick_bench.py mt40_single 1 1000 2>&1 | grep "run 1000" | tail -1
echo "--- mt120 x256"; python tools/quick_bench.py mt120_disassembly 256 400 hydrolysis=no 2>&1 | grep "run 400" | tail -1
if [ -n "$PROFILE" ]; then
  ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 1 -c 1 -o gpurun_out/prof_iter python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/ncu_iter.log 2>&1
  tail -2 gpurun_out/ncu_iter.log
fi
