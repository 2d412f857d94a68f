#!/bin/bash
# round-2 first GPU trip: GPU tests, then both bench arms at the driver's invocation
set -x
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest.log
cat gpurun_out/r2_pytest.log
python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
cat gpurun_out/r2_bench_ref.json; tail -5 gpurun_out/r2_bench_ref.err
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_own.json 2> gpurun_out/r2_bench_own.err
cat gpurun_out/r2_bench_own.json; tail -5 gpurun_out/r2_bench_own.err
