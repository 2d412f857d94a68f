#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/prof_tea_pair python tools/wide_bench.py tea 2600 1 60 > gpurun_out/ncu_tea_full.log 2>&1; tail -2 gpurun_out/ncu_tea_full.log
