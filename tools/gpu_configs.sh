#!/bin/bash
# every BASELINE config end to end (drop-in compute(), no file output) next to the reference's CUDA build
export OMP_NUM_THREADS=8
python tools/config_bench.py mt40_single 1 20000 1
python tools/config_bench.py mt40_ensemble 256 10000 100
python tools/config_bench.py mt120_constconc 128 4000 100
python tools/config_bench.py mt120_disassembly 256 4000 100
python tools/config_bench.py cylinder_tea 64 1000 64
python tools/config_bench.py cylinder_tea_large 1 400 1
