import sys, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import Engine, HostSystem, capi, workspace
ntr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
d = Path(tempfile.mkdtemp())
workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=ntr)
with workspace.chdir(d):
    s = HostSystem("config.conf", ["hydrolysis=no"])
for nsteps in (1, 2, 20, 21, 41):
    a, b = Engine(s), Engine(s)
    a.run(0, nsteps)
    for step in range(nsteps):
        if step % 20 == 0:
            b.rebuild_lj(); b.rebuild_bonds()
        b.force(); b.integrate()
    ca, cb = a.coords(), b.coords()
    diff = np.argwhere((ca != cb).any(axis=2))
    print("steps", nsteps, "differing (traj,monomer):", len(diff), diff[:6].tolist(), "max|d|", np.abs(ca - cb).max())
    for kind in (capi.LIST_LJ, capi.LIST_LONGITUDINAL, capi.LIST_LATERAL):
        ac, ae = a.download_list(kind); bc, be = b.download_list(kind)
        bad = np.argwhere(ac != bc)
        print("   list", kind, "count mismatches", len(bad), bad[:5].tolist(), "fixed?", [bool(s.fixed[i]) for _, i in bad[:5]])
    a.close(); b.close()
