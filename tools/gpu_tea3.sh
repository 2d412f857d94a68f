#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/r2_tea_pair_v2 python tools/quick_bench.py cylinder_tea 64 40 > /dev/null 2>&1
$NCU -k regex:tea_pair_kernel -s 30 -c 1 -f -o gpurun_out/r2_tea_pair_v2_large python tools/quick_bench.py cylinder_tea_large 1 40 > /dev/null 2>&1
ls -la gpurun_out/r2_tea_pair_v2*
