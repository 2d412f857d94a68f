#!/bin/bash
# ncu --set full captures of the other kernels VERDICT names (TEA pair kernel, wide path) + launch lists of the hydrolysis plan
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:tea_pair_kernel -s 30 -c 1 -f -o $O/r2_tea_pair python tools/quick_bench.py cylinder_tea 64 40 > /dev/null 2>&1
$NCU -k regex:tea_pair_kernel -s 30 -c 1 -f -o $O/r2_tea_pair_large python tools/quick_bench.py cylinder_tea_large 1 40 > /dev/null 2>&1
$NCU -k regex:wide_run_kernel -s 2 -c 1 -f -o $O/r2_wide_run python tools/quick_bench.py mt400_single 1 100 > /dev/null 2>&1
NO_REF=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 200 --csv --log-file $O/r2_stride_launches.csv python tools/config_bench.py mt40_ensemble 256 6000 > /dev/null 2>&1
python tools/ncu_launch_summary.py $O/r2_stride_launches.csv
ls -la $O/*.ncu-rep
python bench.py --steps 20 --warmup 5 > $O/r2_bench_own.json 2> $O/r2_bench_own.err
NO_REF=1 python tools/config_bench.py mt120_constconc 128 6000 | tail -1
NO_REF=1 python tools/config_bench.py mt120_disassembly 256 10000 | tail -1
