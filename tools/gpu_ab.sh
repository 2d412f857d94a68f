#!/bin/bash
cd "$GRAFT_REPO_ROOT"
for v in old nolj full old full; do
  echo "== $v"; MADDY_B200_LIB=$PWD/tools/micro/_bin/libmaddy_$v.so python tools/quick_bench.py mt40_ensemble 256 1000 | grep "us/step" | tail -2
done
for v in old full; do
  echo "== $v 2048"; MADDY_B200_LIB=$PWD/tools/micro/_bin/libmaddy_$v.so python tools/quick_bench.py mt40_ensemble 2048 1000 | grep "us/step" | tail -1
done
python -m pytest tests/test_gpu_parity.py tests/test_gpu_events.py -m gpu -x -q 2>&1 | tail -3
