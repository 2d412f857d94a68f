#!/bin/bash
set -x
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
timeout 900 python tests/golden/make_ref_golden.py > gpurun_out/golden.log 2>&1; tail -20 gpurun_out/golden.log
timeout 600 python __graft_entry__.py --smoke 2>&1 | tail -5
timeout 900 python bench.py --steps 2000 --warmup 200 > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err; cat gpurun_out/bench_own.json; tail -5 gpurun_out/bench_own.err
timeout 900 python bench.py --impl reference --steps 2000 --warmup 200 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
# launch list + one full capture of the fused kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1.csv python tools/quick_bench.py mt40_ensemble 256 200 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:traj_kernel -s 1 -c 1 -o gpurun_out/prof_r1_traj python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
