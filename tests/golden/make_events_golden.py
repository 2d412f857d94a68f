"""Generates tests/golden/ref_events.npz: what the REFERENCE's own host callbacks - mt_length() (updater.cpp:154-227),
hydrolyse() (:229-257), change_conc() (:97-152), compiled from /root/reference/src by mt_b200/build.py into
oracle/_ref/ref_events_probe - do on adversarial inputs (bending angles on and around every crossing of
cos(theta) = cos(1), radii on and around both bounds, mixed GTP / reserve / previous-stride flags, a concentration that
forces insertions until some trajectories run out of reserve dimers).  Host code only: runs without a GPU.

    python tests/golden/make_events_golden.py
"""
import struct
import subprocess
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import HostSystem, workspace  # noqa: E402

PROBE = ROOT / "oracle" / "_ref" / "ref_events_probe"
STRUCTURE, NTR, SEED, NHYD, CONC = ("reserve", 40, 20), 5, 20260117, 4, 300.0


def adversarial(c, rng):
    n = c.shape[0] * c.shape[1]
    flat = c.reshape(n, 7)
    cross = np.array([k * 2 * np.pi + sgn for k in range(-3, 4) for sgn in (-1.0, 1.0)])
    base = rng.choice(cross, size=n).astype(np.float32)
    ulps = rng.integers(-40, 41, size=n).astype(np.int32)
    theta = (base.view(np.int32) + ulps).view(np.float32)
    flat[:, 4] = np.where(rng.random(n) < 0.3, rng.uniform(-20.0, 20.0, n).astype(np.float32), theta)
    which = rng.integers(0, 4, size=n)
    rad = np.where(which == 0, 24.12, np.where(which == 1, 1.0, rng.uniform(0.0, 40.0, n))).astype(np.float32)
    rad = (rad.view(np.int32) + rng.integers(-6, 7, size=n).astype(np.int32)).view(np.float32)
    phi = rng.uniform(0, 2 * np.pi, n)
    flat[:, 0] = (rad * np.cos(phi)).astype(np.float32)
    flat[:, 1] = (rad * np.sin(phi)).astype(np.float32)
    on_axis = rng.random(n) < 0.2
    flat[on_axis, 0] = rad[on_axis]
    flat[on_axis, 1] = 0.0


def main():
    rng = np.random.default_rng(SEED)
    tmp = Path(tempfile.mkdtemp(prefix="events_golden_"))
    spec = workspace.BASELINE_CONFIGS["mt120_constconc"]
    cond = dict(spec["conditions"], conc=CONC)
    workspace.make_rundir(tmp, STRUCTURE, dict(spec["config"], runnum=NTR), dict(spec["forcefield"]), cond)
    with workspace.chdir(tmp):
        s = HostSystem("config.conf")
    N = s.Ntot
    c = np.array(s.coords, dtype=np.float32).copy()
    adversarial(c, rng)
    extra = np.array(s.extra).reshape(NTR, N).astype(np.int32)
    # trajectory 1 keeps two reserve dimers only (it runs out: "No more extra particles"), trajectory 3 none at all
    res = np.flatnonzero(extra[1])
    extra[1, res[4:]] = 0
    extra[3, :] = 0
    gtp = (rng.random((NTR, N // 2)) < 0.7).astype(np.int32).repeat(2, axis=1)
    on_cur = (rng.random((NTR, N)) < 0.5).astype(np.int32)   # overwritten by mt_length()
    on_prev = (rng.random((NTR, N // 2)) < 0.6).astype(np.int32).repeat(2, axis=1)
    mon_type = np.array(s.mon_type, dtype=np.int32)
    len_prev = rng.integers(0, N, size=NTR).astype(np.int32)
    case = tmp / "case.bin"
    with open(case, "wb") as f:
        f.write(struct.pack("<4i4f", N, NTR, SEED % 2**31, NHYD, s.par.rep_r, s.par.rep_h, s.par.rep_leftborder, s.host.conc))
        for a in (c, gtp, on_cur, on_prev, extra, mon_type, len_prev):
            f.write(np.ascontiguousarray(a).tobytes())
    subprocess.run([str(PROBE), str(case), str(tmp / "out.bin")], cwd=str(tmp), check=True)
    raw = (tmp / "out.bin").read_bytes()
    n, o = NTR * N, 0

    def take(count, dtype):
        nonlocal o
        a = np.frombuffer(raw, dtype=dtype, count=count, offset=o).copy()
        o += a.nbytes
        return a

    out = {"on_tubule": take(n, np.int32).reshape(NTR, N), "mt_len": take(NTR, np.int32),
           "gtp_after": np.stack([take(n, np.int32).reshape(NTR, N) for _ in range(NHYD)]), "flag": take(1, np.int32),
           "extra_after": take(n, np.int32).reshape(NTR, N), "coords_after": take(n * 7, np.float32).reshape(NTR, N, 7),
           "next_rand": take(4, np.int32)}
    assert o == len(raw)
    np.savez_compressed(ROOT / "tests" / "golden" / "ref_events.npz", structure=np.array(STRUCTURE[1:]), ntr=NTR, seed=SEED % 2**31, conc=CONC,
                        coords=c, gtp=gtp, on_cur=on_cur, on_prev=on_prev, extra=extra, len_prev=len_prev, **out)
    print("on-tubule fraction", out["on_tubule"].mean(), "mt_len", out["mt_len"], "insertions", int(out["flag"][0]),
          "hydrolysed per event", [(int((out["gtp_after"][k] == 0).sum())) for k in range(NHYD)])


if __name__ == "__main__":
    main()
