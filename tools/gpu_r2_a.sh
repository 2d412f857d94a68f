#!/bin/bash
# round 2, call A: full GPU suite, both bench arms, device-side vs host-side stride events on configs 3 and 4
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_own.json 2> gpurun_out/r2a_bench_own.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
{
for cfg in "mt120_disassembly 256 10000" "mt120_constconc 128 10000"; do
  NO_REF=1 python tools/config_bench.py $cfg
  NO_REF=1 MADDY_HOST_EVENTS=1 MADDY_NO_OVERLAP=1 python tools/config_bench.py $cfg
  NO_REF=1 MADDY_HOST_PROFILE=1 python tools/config_bench.py $cfg 2>&1 | tail -25
done
} > gpurun_out/r2a_events.log 2>&1
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_events.log | grep "own e2e"
