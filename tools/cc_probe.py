"""Per-trajectory window time after constant-concentration insertion (development aid)."""
import os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from mt_b200 import Engine, HostSystem, workspace
d = Path(tempfile.mkdtemp())
workspace.make_baseline_rundir(d, "mt120_constconc", runnum=128, steps=1001)
with workspace.chdir(d):
    s = HostSystem("config.conf")
    s.srand(s.par.rseed)
    s.compute()
c = np.array(s.coords)
ex = np.array(s.extra).astype(bool)
print("finite", np.isfinite(c).all(), "max |xyz|", np.abs(c[..., :3]).max(), "inserted per traj", (~ex).sum(axis=1)[:4] - 1560)
rows = []
for t in range(128):
    live = ~ex[t]
    p = c[t, live, :3]
    free = p[1560:]
    if len(free) > 1:
        dd = np.linalg.norm(free[:, None] - free[None], axis=-1) + np.eye(len(free)) * 1e9
        mind = dd.min()
    else:
        mind = -1
    e = Engine(s, traj_first=t, n_tr_local=1)
    e.run(1000, 20); e.sync()
    t0 = time.perf_counter(); e.run(1020, 200); e.sync(); dt = time.perf_counter() - t0
    cc = e.coords(); stats = e.list_stats()
    rows.append((dt, t, stats, mind, bool(np.isfinite(cc).all())))
    if 0: print(f"traj {t:2d}: {dt / 200 * 1e6:7.1f} us/step   min free-free distance at insertion {mind:6.2f} nm   finite after {np.isfinite(cc).all()}  max|xyz| {np.abs(cc[..., :3]).max():.1f}")
    e.close()
rows.sort(reverse=True)
for dt, t, stats, mind, fin in rows[:12]:
    print(f"traj {t:3d}: {dt / 200 * 1e6:7.1f} us/step  finite {fin}  min free-free {mind:5.2f}  {stats}")
print("median", sorted(r[0] for r in rows)[64] / 200 * 1e6)
# whole ensemble: before and after the insertions
def timed(e, first, nsteps):
    e.run(first, 20); e.sync(); e.list_stats(reset=True)
    t0 = time.perf_counter(); e.run(first + 20, nsteps); e.sync(); dt = time.perf_counter() - t0
    return dt / nsteps * 1e6, e.list_stats()
e = Engine(s)
print("after insertion, 128 traj:", timed(e, 1000, 980))
e.close()
workspace.make_baseline_rundir(d / "b", "mt120_constconc", runnum=128, steps=1001)
with workspace.chdir(d / "b"):
    s0 = HostSystem("config.conf")
e = Engine(s0)
print("before insertion, 128 traj:", timed(e, 0, 980))
e.close()
