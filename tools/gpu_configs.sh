#!/bin/bash
export OMP_NUM_THREADS=8
python tools/config_bench.py mt40_single 1 20000 1
python tools/config_bench.py mt120_disassembly 256 4000 100
python tools/config_bench.py mt120_constconc 128 4000 100
python tools/config_bench.py cylinder_tea 64 600 64
python tools/config_bench.py cylinder_tea 1 2000 1
