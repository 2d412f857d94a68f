#!/bin/bash
# the stride-step launch (rebuild + energies) and the tests that pin its results
export OMP_NUM_THREADS=8
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/stride_launches.csv python bench.py --steps 3000 --warmup 100 --no-cpu-baseline > gpurun_out/ncu_stride.log 2>&1
python tools/ncu_launch_summary.py gpurun_out/stride_launches.csv | head -4
