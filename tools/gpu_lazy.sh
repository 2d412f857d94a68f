#!/bin/bash
export OMP_NUM_THREADS=8
echo "--- baseline"; python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000" | tail -1
echo "--- lazy"; MADDY_LAZY=1 python tools/quick_bench.py mt40_ensemble 256 1000 2>&1 | grep "run 1000\|finite" | tail -2
echo "--- lazy 2048"; MADDY_LAZY=1 python tools/quick_bench.py mt40_ensemble 2048 400 2>&1 | grep "run 400" | tail -1
echo "--- lazy 1560"; MADDY_LAZY=1 python tools/quick_bench.py mt120_disassembly 256 400 hydrolysis=no 2>&1 | grep "run 400" | tail -1
