#!/bin/bash
# N GPUs of one box: both bench arms under torchrun (as the driver launches them) and the one-host probe.
# usage (GPU box): bash tools/gpu_scale.sh <N> <tag>
N=${1:-8}; TAG=${2:-r2}
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
$RUN bench.py --gpus $N --steps 20 --warmup 5 > $O/${TAG}_scale_own_n$N.json 2> $O/${TAG}_scale_own_n$N.err
[ -z "$SKIP_REF" ] && $RUN bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/${TAG}_scale_ref_n$N.json 2> $O/${TAG}_scale_ref_n$N.err
python tools/one_host_probe.py $N 256 20000 4 > $O/${TAG}_one_host_n$N.log 2>&1
python - <<PY
import json
for f in ('$O/${TAG}_scale_own_n$N.json', '$O/${TAG}_scale_ref_n$N.json'):
    try:
        d = json.load(open(f))
        print(f, 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d.get('shard_parity'), (d.get('e2e_one_host') or {}).get('value'), (d.get('e2e_one_host') or {}).get('wall_s_runs'))
    except Exception as e:
        print(f, 'ERR', e)
PY
tail -14 $O/${TAG}_one_host_n$N.log
