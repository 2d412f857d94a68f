#!/bin/bash
# full round check on a GPU box: test-suite, smoke, both bench arms, launch list + full ncu capture
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu 2>&1 | tail -15
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err; cat gpurun_out/bench_own.json; tail -3 gpurun_out/bench_own.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 400 --warmup 100 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 2 -c 1 -o gpurun_out/prof_full python tools/quick_bench.py mt40_ensemble 256 100 > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
