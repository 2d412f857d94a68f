"""One host thread driving G handles (mt_system_compute, n_gpus=G): wall time per run and the host profile of the last run.
usage: python tools/one_host_probe.py <n_gpus> <ntr_per_gpu> [steps] [runs]"""
import os, sys, tempfile, time, shutil
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import HostSystem, workspace
G, per = int(sys.argv[1]), int(sys.argv[2])
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
runs = int(sys.argv[4]) if len(sys.argv) > 4 else 4
base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
root = Path(tempfile.mkdtemp(prefix="onehost_", dir=base))
for r in range(runs):
    d = root / f"r{r}"
    workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=G * per, steps=steps)
    if r == runs - 1:
        os.environ["MADDY_HOST_PROFILE"] = "1"
    with workspace.chdir(d):
        s = HostSystem("config.conf", ["device=0"], write_files=True)
        s.srand(s.par.rseed)
        t0 = time.perf_counter()
        s.compute(n_gpus=G)
        dt = time.perf_counter() - t0
        s.close()
    print(f"run {r}: G={G} ntr={G * per} steps={steps} wall {dt:.3f} s  {520 * G * per * steps / dt / 1e9:.2f} G monomer-steps/s", flush=True)
    shutil.rmtree(d, ignore_errors=True)
shutil.rmtree(root, ignore_errors=True)
