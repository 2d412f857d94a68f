#!/bin/bash
# config 5 (TEA) and the large-N systems end to end next to the reference's CUDA build
export OMP_NUM_THREADS=8
python tools/config_bench.py cylinder_tea 64 1000 64
python tools/config_bench.py cylinder_tea 1 2000 1
python tools/config_bench.py cylinder_tea_large 1 400 1
python tools/config_bench.py mt400_single 1 2000 1 hydrolysis=no
python tools/config_bench.py mt400_single 16 1000 16 hydrolysis=no
