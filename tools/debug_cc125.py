import os, sys, shutil, subprocess, tempfile
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import mt_b200
from mt_b200 import workspace
G = ROOT / "tests" / "golden" / "inputs"
spec = workspace.BASELINE_CONFIGS["mt120_constconc"]
base = Path(tempfile.mkdtemp(prefix="cc125_"))
outs = {}
for mode in ("device", "host", "nohyd"):
    d = base / mode
    cfg = dict(spec["config"]); cfg.update(runnum=2, steps=650, stride=200)
    cond = dict(spec["conditions"]); cond.update(conc=200)
    workspace.make_rundir(d, ("files", str(G / "constconc125_xyz.pdb"), str(G / "constconc125_ang.pdb")), cfg, dict(spec["forcefield"]), cond)
    env = dict(os.environ)
    if mode == "host":
        env["MADDY_HOST_EVENTS"] = "1"; env["MADDY_NO_OVERLAP"] = "1"
    if mode == "nohyd":
        env["MADDY_HOST_HYDROLYSIS"] = "1"
    r = subprocess.run([str(ROOT / "mt_b200" / "mt"), "config.conf"], cwd=str(d), capture_output=True, text=True, env=env)
    print(mode, "rc", r.returncode, r.stderr[-300:])
    outs[mode] = (r.stdout, [mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd") for t in range(2)])
for m in ("device", "nohyd"):
    for t in range(2):
        a, b = outs[m][1][t], outs["host"][1][t]
        print(m, "traj", t, "shape", a.shape, b.shape)
        for f in range(min(a.shape[0], b.shape[0])):
            df = np.abs(a[f] - b[f])
            df = np.where(np.isnan(df), 0, df)
            i = np.unravel_index(np.argmax(df), df.shape)
            print("   frame", f, "max diff", df.max(), "at", i, a[f][i[0]], b[f][i[0]], "nan:", np.isnan(a[f]).sum(), np.isnan(b[f]).sum())
    so, sh = outs[m][0].splitlines(), outs["host"][0].splitlines()
    diff = [(i, x, y) for i, (x, y) in enumerate(zip(so, sh)) if x != y and "Computation time" not in x and "Estimated" not in x]
    print(m, "stdout lines", len(so), len(sh), "first diffs", diff[:4])
from oracle import refprobe
d = base / "ref"
cfg = dict(spec["config"]); cfg.update(runnum=2, steps=650, stride=200)
cond = dict(spec["conditions"]); cond.update(conc=200)
workspace.make_rundir(d, ("files", str(G / "constconc125_xyz.pdb"), str(G / "constconc125_ang.pdb")), cfg, dict(spec["forcefield"]), cond)
_, out_ref = refprobe.run_reference_mt(d)
for t in range(2):
    a, b = outs["host"][1][t], mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd")
    for f in range(4):
        df = np.abs(a[f] - b[f]); df = np.where(np.isnan(df), 0, df)
        bad = np.argwhere(df.max(axis=1) > 1e-2)[:, 0]
        print("ref vs own traj", t, "frame", f, "max", df.max(), "n bad", len(bad), bad[:10], "nan own/ref", np.isnan(a[f]).any(axis=1).sum(), np.isnan(b[f]).any(axis=1).sum())
        for i in bad[:3]:
            print("     monomer", i, "own", a[f][i], "ref", b[f][i])
import re
ins = re.compile(r"New x,y coordinates for extra particle: .*")
io, ir = ins.findall(outs["host"][0]), ins.findall(out_ref)
print("insertions own/ref", len(io), len(ir), io == ir, io[:2], ir[:2])
print("---- criteria")
for t in range(2):
    for suffix in (".dcd", ".dcd_ang"):
        a = mt_b200.read_dcd(base / "device" / "dcd" / f"run_{t}{suffix}")
        b = mt_b200.read_dcd(d / "dcd" / f"run_{t}{suffix}")
        bx = mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd")
        dist = np.linalg.norm(bx, axis=-1, keepdims=True)
        flung = np.broadcast_to(dist > 2000.0, a.shape)
        err = np.abs(a - b)
        sane_err = np.where(flung, 0, err)
        i = np.unravel_index(np.argmax(sane_err), err.shape)
        rel = (err / np.maximum(dist, 1.0))
        relf = np.where(flung, rel, 0)
        j = np.unravel_index(np.argmax(relf), err.shape)
        print(t, suffix, "flung frac", flung.mean(), "worst sane", sane_err.max(), "at", i, a[i[0], i[1]], b[i[0], i[1]], "dist", dist[i[0], i[1]],
              "| worst flung rel", relf.max(), "at", j, a[j[0], j[1]], b[j[0], j[1]])
