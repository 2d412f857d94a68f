#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
N=${NGPU:-2}
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -q -m gpu 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20000 --warmup 500 > gpurun_out/bench_own_n$N.json 2> gpurun_out/bench_own_n$N.err; cat gpurun_out/bench_own_n$N.json | cut -c1-700; tail -3 gpurun_out/bench_own_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 5000 --warmup 200 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cat gpurun_out/bench_ref_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_ref_n$N.err
python bench.py --steps 100000 --warmup 1000 > gpurun_out/bench_own_n1.json 2> gpurun_out/bench_own_n1.err; cat gpurun_out/bench_own_n1.json | cut -c1-1200
