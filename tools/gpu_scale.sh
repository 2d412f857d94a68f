#!/bin/bash
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
N=${NGPU:-8}
nvidia-smi -L | head -8
for n in $N; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $n --steps 30000 --warmup 1000 > gpurun_out/scale_own_n$n.json 2> gpurun_out/scale_own_n$n.err
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/scale_own_n$n.json') if l.startswith('{')][-1]; print('own', d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['wall_s_runs'])"
tail -2 gpurun_out/scale_own_n$n.err
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus $N --steps 5000 --warmup 500 > gpurun_out/scale_ref_n$N.json 2> gpurun_out/scale_ref_n$N.err
python -c "
import json; d=[json.loads(l) for l in open('gpurun_out/scale_ref_n$N.json') if l.startswith('{')][-1]; print('ref', d['n_gpus'], d['value'])"
