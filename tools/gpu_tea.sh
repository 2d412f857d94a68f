#!/bin/bash
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu -k "tea" 2>&1 | tail -4
python tools/config_bench.py cylinder_tea 64 1000 64
python tools/config_bench.py cylinder_tea 1 4000 1
