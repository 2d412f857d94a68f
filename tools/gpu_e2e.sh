#!/bin/bash
# where does the end-to-end wall time of the drop-in loop go?  (host profile of mt_system_compute on the bench workload)
mkdir -p gpurun_out
export OMP_NUM_THREADS=8
python - <<'P' 2>&1 | tail -40
import os, sys, time, tempfile
from pathlib import Path
sys.path.insert(0, os.getcwd())
import bench
from mt_b200 import workspace
os.environ["MADDY_HOST_PROFILE"] = "1"; os.environ["MADDY_GPU_PROFILE"] = "1"
for rep in range(2):
    tmp = Path(tempfile.mkdtemp(prefix="e2e_"))
    s = bench.make_system("mt40_ensemble", 256, tmp, write_files=True, steps=100000)
    s.srand(s.par.rseed)
    with workspace.chdir(tmp):
        t0 = time.perf_counter()
        st = s.compute(fused=True)
        dt = time.perf_counter() - t0
    print(f"rep {rep}: compute() wall {dt:.3f} s -> {520*256*100000/dt/1e9:.2f} G monomer-steps/s, launches {st.get('launches')}")
    s.close()
P
