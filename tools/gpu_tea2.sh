#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
for cfg in "cylinder_tea 64" "cylinder_tea_large 1"; do
  set -- $cfg
  python tools/quick_bench.py $1 $2 300 2>&1 | grep "us/step" | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 150 --csv --log-file gpurun_out/r2_tea_launches_$1_$2.csv python tools/quick_bench.py $1 $2 100 > /dev/null 2>&1
  python tools/ncu_launch_summary.py gpurun_out/r2_tea_launches_$1_$2.csv
done
