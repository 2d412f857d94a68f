#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu -k "tea" -x 2>&1 | tail -8
python tools/wide_bench.py tea 2600 1 500 --window 2>&1 | tail -2
python tools/wide_bench.py tea 247 64 1000 --window 2>&1 | tail -2
python tools/wide_bench.py tea 247 1 2000 --window 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/tea_launches.csv python tools/wide_bench.py tea 2600 1 100 > gpurun_out/tea_ncu.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/tea64_launches.csv python tools/wide_bench.py tea 247 64 100 > gpurun_out/tea64_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/prof_tea_pair python tools/wide_bench.py tea 2600 1 60 > gpurun_out/ncu_tea_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/prof_tea_pair64 python tools/wide_bench.py tea 247 64 60 > gpurun_out/ncu_tea64_full.log 2>&1
