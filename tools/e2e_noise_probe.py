"""Repeats the bench's e2e leg (drop-in compute(), 20 strides, files on tmpfs) with the host profile on, to see which phase
carries the run-to-run spread (development aid)."""
import os, sys, tempfile, time, shutil, io, contextlib
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import HostSystem, workspace
os.environ["MADDY_HOST_PROFILE"] = "1"
base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
root = Path(tempfile.mkdtemp(prefix="e2e_noise_", dir=base))
for r in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    d = root / f"r{r}"
    workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=256, steps=20000)
    with workspace.chdir(d):
        t0 = time.perf_counter()
        s = HostSystem("config.conf", ["device=0"], write_files=True)
        t1 = time.perf_counter()
        s.srand(s.par.rseed)
        sys.stderr.write(f"--- run {r}\n"); sys.stderr.flush()
        s.compute()
        t2 = time.perf_counter()
        s.close()
        t3 = time.perf_counter()
    print(f"run {r}: load {t1 - t0:.3f}  compute {t2 - t1:.3f}  close {t3 - t2:.3f}", flush=True)
    shutil.rmtree(d, ignore_errors=True)
shutil.rmtree(root, ignore_errors=True)
