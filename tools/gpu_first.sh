#!/bin/bash
# first GPU contact: parity walk against the reference kernels + a first timing
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
export OMP_NUM_THREADS=8
timeout 600 python tools/parity_report.py mt40_single 2 60 > gpurun_out/parity_mt40.log 2>&1
tail -60 gpurun_out/parity_mt40.log
timeout 600 python tools/parity_report.py mt120_disassembly 1 40 probe_gdp_every=3 probe_ontub=1 > gpurun_out/parity_mt120.log 2>&1
tail -40 gpurun_out/parity_mt120.log
timeout 300 python tools/quick_bench.py mt40_ensemble 256 1000 > gpurun_out/qb_256.log 2>&1
cat gpurun_out/qb_256.log
timeout 300 python tools/quick_bench.py mt40_single 1 1000 > gpurun_out/qb_1.log 2>&1
cat gpurun_out/qb_1.log
