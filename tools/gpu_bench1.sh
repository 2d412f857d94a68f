#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2f_bench_own.json 2> gpurun_out/r2f_bench_own.err
tail -3 gpurun_out/r2f_bench_own.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench_own.json'))
print('value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['wall_s_runs'], 'roofline', d['roofline']['frac'], 'launches', d['gpu_launches'], 'reps', d['config']['repetitions'], d['config']['window_pattern'])
PY
NO_REF=1 python tools/config_bench.py mt120_disassembly 256 10000 | tail -1
NO_REF=1 python tools/config_bench.py mt40_ensemble 256 100000 | tail -1
