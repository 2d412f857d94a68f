"""Time one BASELINE config end to end (drop-in compute(), no file output) and the reference CUDA binary beside it.

usage: python tools/config_bench.py <config> <ntr> <steps> [ref_ntr] [overrides...]
Prints monomer-steps/s for both (the reference by the difference of two run lengths)."""
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import HostSystem, workspace  # noqa: E402
from oracle import refprobe  # noqa: E402


def main():
    name, ntr, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    ref_ntr = int(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4].isdigit() else min(ntr, 100)
    over = [a for a in sys.argv[4:] if "=" in a]
    d = Path(tempfile.mkdtemp(prefix="cb_"))
    workspace.make_baseline_rundir(d / "own", name, runnum=ntr, steps=steps)
    with workspace.chdir(d / "own"):
        s = HostSystem("config.conf", over)
        s.srand(s.par.rseed)
        s.compute(steps=min(steps, 100))
    with workspace.chdir(d / "own"):
        s = HostSystem("config.conf", over)
        s.srand(s.par.rseed)
        t0 = time.perf_counter()
        st = s.compute()
        dt = time.perf_counter() - t0
    N = s.Ntot
    print(f"{name}: N={N} Ntr={ntr} steps={steps}  own e2e {N * ntr * steps / dt / 1e9:.3f} G monomer-steps/s  ({dt / steps * 1e6:.1f} us/step, launches {st['launches']})")
    if refprobe.REF_MT.exists() and not os.environ.get('NO_REF'):
        def run(k):
            r = d / f"ref{k}"
            workspace.make_baseline_rundir(r, name, runnum=ref_ntr, steps=k)
            t0 = time.perf_counter()
            subprocess.run([str(refprobe.REF_MT), "config.conf", *over], cwd=str(r), stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
            return time.perf_counter() - t0
        k1 = max(20, steps // 20)
        run(k1)
        ta, tb = run(k1), run(k1 + steps // 4)
        per = (tb - ta) / (steps // 4)
        print(f"   reference CUDA build: Ntr={ref_ntr}  {N * ref_ntr / per / 1e9:.4f} G monomer-steps/s  ({per * 1e6:.1f} us/step)")
    shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    main()
