#!/bin/bash
cd "$GRAFT_REPO_ROOT"
mkdir -p gpurun_out
{
python tools/one_host_probe.py 1 256
python tools/one_host_probe.py 2 256
python tools/one_host_probe.py 1 2048 10000 3
} > gpurun_out/r2d_one_host.log 2>&1
cat gpurun_out/r2d_one_host.log
