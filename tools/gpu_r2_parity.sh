#!/bin/bash
# round-2 parity trip: the new parity tests (live reference binaries), goldens written to gpurun_out/golden
python -m pytest tests/test_gpu_parity_r2.py -m gpu -q -s -k "const_conc or stub" 2>&1 | tail -150 > gpurun_out/r2_parity.log
cat gpurun_out/r2_parity.log
