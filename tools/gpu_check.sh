#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu 2>&1 | tail -6
python tools/wide_bench.py tea 247 64 1000 --window 2>&1 | tail -2
python tools/wide_bench.py tea 247 1 2000 --window 2>&1 | tail -2
python tools/wide_bench.py tea 2600 1 500 --window 2>&1 | tail -2
python tools/config_bench.py cylinder_tea 64 4000 64
