#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "mt400_single 1" "mt400_single 16"; do
  set -- $cfg
  timeout 120 python tools/quick_bench.py $1 $2 1000 2>&1 | grep "us/step" | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 60 --csv --log-file gpurun_out/r2_wide_launches_$2.csv python tools/quick_bench.py $1 $2 200 > /dev/null 2>&1
  python tools/ncu_launch_summary.py gpurun_out/r2_wide_launches_$2.csv 2>/dev/null
done
