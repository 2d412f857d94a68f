#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
NO_REF=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 200 --csv --log-file gpurun_out/r2_stride_launches.csv python tools/config_bench.py mt40_ensemble 256 6000 > /dev/null 2>&1
python tools/ncu_launch_summary.py gpurun_out/r2_stride_launches.csv
MADDY_HOST_PROFILE=1 NO_REF=1 python tools/config_bench.py mt40_ensemble 256 20000 2>&1 | tail -14
