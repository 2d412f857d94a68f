#!/bin/bash
export OMP_NUM_THREADS=8
python -m pytest tests -q -m gpu -x -k "fused_equals_step_granular or lazy_verlet or crowded or wide_path_equals or list_hierarchy" 2>&1 | tail -3
python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -2
python tools/quick_bench.py mt40_ensemble 2048 1000 2>&1 | grep "run 1000" | tail -1
python tools/quick_bench.py mt120_disassembly 256 1000 2>&1 | grep "run 1000" | tail -1
