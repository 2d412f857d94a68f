#!/bin/bash
# development check: GPU suite, the headline bench line, the constant-concentration probe
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/check_bench.json 2> gpurun_out/check_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/check_bench.json'))
print('value %.4g' % d['value'], 'ms/step', round(d['ms_per_step'], 4), 'e2e %.4g' % d['e2e']['value'], d['e2e']['wall_s_runs'], 'roofline', round(d['roofline']['frac'], 3))
PY
python tools/cc_probe.py 2>&1 | tail -4
NO_REF=1 python tools/config_bench.py mt120_constconc 128 6000 | tail -1
