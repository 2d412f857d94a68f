"""Device timing of the wide path (N > MADDY_MAX_NTOT_CTA) on a GPU box (development aid; bench.py is the contract).

  python tools/wide_bench.py lattice <mt_len> <ntr> <steps>      fused windows (maddy_run: one launch per step)
  python tools/wide_bench.py tea <n_dimers> <ntr> <steps>        step-granular TEA loop (force -> tea_update -> tea_integrate)
"""
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import Engine, HostSystem, workspace  # noqa: E402


def main():
    kind, size, ntr, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    d = Path(tempfile.mkdtemp(prefix="wb_"))
    if kind == "lattice":
        workspace.make_rundir(d, ("lattice", size, 3), dict(workspace.BASELINE_CONFIGS["mt40_single"]["config"], runnum=ntr))
    else:
        spec = workspace.BASELINE_CONFIGS["cylinder_tea"]
        side = (size / 247.0) ** (1.0 / 3.0)
        workspace.make_rundir(d, ("free", size, 30.0 * side, 160.0 * side, 1), dict(spec["config"], runnum=ntr), None, spec["conditions"])
    with workspace.chdir(d):
        s = HostSystem("config.conf", ["hydrolysis=no"] + [x for x in sys.argv[5:] if not x.startswith("--")])
    e = Engine(s)
    print(f"N={s.Ntot} Ntr={s.Ntr} tea={bool(s.par.tea_on)}")
    freq = s.par.ljpairsupdatefreq

    def window(first, n):
        if not s.par.tea_on or "--window" in sys.argv:
            e.run(first, n)
            return
        for step in range(first, first + n):
            if step % freq == 0:
                if s.par.lj_on:
                    e.rebuild_lj()
                if s.par.is_assembly:
                    e.rebuild_bonds()
            e.force()
            e.tea_update(step)
            e.tea_integrate()

    window(0, 40)
    e.sync()
    for rep in range(3):
        l0 = e.launches
        t0 = time.perf_counter()
        window(40 + rep * steps, steps)
        e.sync()
        dt = time.perf_counter() - t0
        print(f"{steps} steps: {dt * 1e3:.1f} ms -> {dt / steps * 1e6:.2f} us/step, {s.Ntot * s.Ntr * steps / dt / 1e9:.3f} G monomer-steps/s, "
              f"{(e.launches - l0) / steps:.2f} launches/step")
    import numpy as np
    print("finite:", bool(np.isfinite(e.coords()).all()))


if __name__ == "__main__":
    main()
