#!/bin/bash
# wide path: parity tests + timings, then the full GPU suite and both bench arms
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu -k "wide" -x 2>&1 | tail -15
python tools/wide_bench.py lattice 400 1 2000 2>&1 | tail -5
python tools/wide_bench.py lattice 400 16 1000 2>&1 | tail -5
python tools/wide_bench.py tea 2600 1 500 2>&1 | tail -5
python -m pytest tests -q -m gpu 2>&1 | tail -8
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py > gpurun_out/bench_own.json 2> gpurun_out/bench_own.err; cat gpurun_out/bench_own.json; tail -3 gpurun_out/bench_own.err
