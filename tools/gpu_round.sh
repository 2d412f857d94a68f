#!/bin/bash
# The round's standard single-GPU record: GPU test suite, both bench arms, TEA bench line, every BASELINE config end to end
# next to the reference's CUDA build, ncu launch list of the bench command and one full capture of the dominant kernel.
# usage (GPU box): bash tools/gpu_round.sh <tag>      -> gpurun_out/<tag>_*
TAG=${1:-r2}
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2> $O/${TAG}_bench_ref.err
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench_own.json 2> $O/${TAG}_bench_own.err
python bench.py --workload cylinder_tea --ntr 64 --steps 4 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_tea64.json 2> $O/${TAG}_bench_tea64.err
{
python tools/config_bench.py mt40_single 1 20000
python tools/config_bench.py mt40_ensemble 256 20000
python tools/config_bench.py mt120_constconc 128 6000
python tools/config_bench.py mt120_disassembly 256 10000
python tools/config_bench.py cylinder_tea 64 2000
python tools/config_bench.py cylinder_tea 1 2000
python tools/config_bench.py cylinder_tea_large 1 1000
python tools/config_bench.py mt400_single 1 4000
python tools/config_bench.py mt400_single 16 2000 16
} > $O/${TAG}_configs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --min-timed-s 0.01 > $O/${TAG}_bench_under_ncu.log 2>&1
python tools/ncu_launch_summary.py $O/${TAG}_launches.csv > $O/${TAG}_launch_summary.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:run_kernel -s 3 -c 1 -f -o $O/${TAG}_run_kernel python tools/quick_bench.py mt40_ensemble 256 100 > /dev/null 2>&1
tail -2 $O/${TAG}_pytest.log; cat $O/${TAG}_configs.log; cat $O/${TAG}_launch_summary.txt | head -12
