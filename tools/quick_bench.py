"""Quick device timing of maddy_run on a GPU box (development aid; bench.py is the contract)."""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import Engine, HostSystem, workspace  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "mt40_ensemble"
    ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
    d = Path(tempfile.mkdtemp(prefix="qb_"))
    workspace.make_baseline_rundir(d, name, runnum=ntr)
    with workspace.chdir(d):
        s = HostSystem("config.conf", sys.argv[4:])
    t0 = time.perf_counter()
    e = Engine(s)
    print(f"create: {time.perf_counter() - t0:.2f}s  N={s.Ntot} Ntr={s.Ntr}")
    e.run(0, 100)
    e.sync()
    for rep in range(3):
        t0 = time.perf_counter()
        e.run(100 + rep * steps, steps)
        e.sync()
        dt = time.perf_counter() - t0
        print(f"run {steps} steps: {dt * 1e3:.1f} ms  -> {dt / steps * 1e6:.2f} us/step, {s.Ntot * s.Ntr * steps / dt / 1e9:.3f} G monomer-steps/s")
    import numpy as np
    c = e.coords()
    print("finite:", bool(np.isfinite(c).all()), "energies[0]:", e.energies()[0])


if __name__ == "__main__":
    main()
