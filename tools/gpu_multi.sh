#!/bin/bash
# N GPUs: the whole GPU suite (multi-GPU tests included), the bench at N under torchrun, the one-host probe
N=${1:-2}
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/multi_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/multi_pytest.log
tail -4 gpurun_out/multi_pytest.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi_bench_n$N.json 2> gpurun_out/multi_bench_n$N.err
tail -3 gpurun_out/multi_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/multi_bench_n$N.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['wall_s_runs'], 'shard_parity', d.get('shard_parity'), 'one_host', d.get('e2e_one_host'))
PY
python tools/one_host_probe.py $N 256 20000 4 2>&1 | tail -16
