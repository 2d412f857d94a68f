#!/bin/bash
# ncu captures (one kernel each, source imported) of the three kernels VERDICT names
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:run_kernel -s 2 -c 1 -f -o gpurun_out/r2_run_kernel python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/r2e_run.log 2>&1
$NCU -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/r2_tea_pair python tools/quick_bench.py cylinder_tea 64 40 > gpurun_out/r2e_tea.log 2>&1
$NCU -k regex:wide_step_kernel -s 30 -c 1 -f -o gpurun_out/r2_wide_step python tools/quick_bench.py mt400_single 1 40 > gpurun_out/r2e_wide.log 2>&1
ls -la gpurun_out/*.ncu-rep
python tools/quick_bench.py mt40_ensemble 256 1000 | tail -4
python tools/quick_bench.py mt40_ensemble 2048 1000 | tail -4
