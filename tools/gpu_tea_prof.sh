#!/bin/bash
# ncu captures quoted in profiles/: TEA pair kernel (both shapes), wide step kernel, launch lists of a TEA step and a wide window
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/prof_tea_pair python tools/wide_bench.py tea 2600 1 60 > gpurun_out/ncu_tea_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/prof_tea_pair64 python tools/wide_bench.py tea 247 64 60 > gpurun_out/ncu_tea64_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:wide_step_kernel -s 30 -c 1 -f -o gpurun_out/prof_wide_step python tools/wide_bench.py lattice 400 1 60 > gpurun_out/ncu_wide_full.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 130 --csv --log-file gpurun_out/tea_launches.csv python tools/wide_bench.py tea 2600 1 100 > gpurun_out/tea_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 130 --csv --log-file gpurun_out/tea64_launches.csv python tools/wide_bench.py tea 247 64 100 > gpurun_out/tea64_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 130 --csv --log-file gpurun_out/wide_launches.csv python tools/wide_bench.py lattice 400 1 100 > gpurun_out/wide_ncu.log 2>&1
./tools/micro/_bin/ffma2_bench > gpurun_out/ffma2_bench.txt 2>&1; cat gpurun_out/ffma2_bench.txt
