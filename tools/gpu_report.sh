#!/bin/bash
# refresh the numbers quoted in profiles/: parity report on two cases, every BASELINE config end to end next to the reference
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python tools/parity_report.py mt40_single 2 60 > gpurun_out/parity_mt40.log 2>&1; tail -12 gpurun_out/parity_mt40.log
python tools/parity_report.py mt120_disassembly 2 60 > gpurun_out/parity_mt120.log 2>&1; tail -12 gpurun_out/parity_mt120.log
bash tools/gpu_configs.sh 2>&1 | tee gpurun_out/configs.log
