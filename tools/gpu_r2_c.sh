#!/bin/bash
# 2 GPUs: whole GPU suite (incl. the multi-GPU tests), bench own arm at N=2 under torchrun
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2c_bench_n2.json 2> gpurun_out/r2c_bench_n2.err
tail -3 gpurun_out/r2c_pytest.log; tail -c 1500 gpurun_out/r2c_bench_n2.json; tail -5 gpurun_out/r2c_bench_n2.err
