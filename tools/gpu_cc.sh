#!/bin/bash
export OMP_NUM_THREADS=8
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python tools/cc_probe.py 2>&1 | grep -A1 "all_pairs_fallback': [1-9]\|near_overflow': [1-9]" | head -12
python tools/config_bench.py mt120_constconc 128 4000 x 2>&1 | tail -2
python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -1
python tools/quick_bench.py mt120_disassembly 256 400 hydrolysis=no 2>&1 | grep "run 400" | tail -1
