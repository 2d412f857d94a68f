mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 130 --csv --log-file gpurun_out/wide_launches.csv python tools/wide_bench.py lattice 400 1 100 > gpurun_out/wide_ncu.log 2>&1
MADDY_GPU_PROFILE=1 python tools/wide_bench.py lattice 400 1 300 2>&1 | tail -12
