"""Shared helpers of the test-suite."""
import ctypes as C

import numpy as np


class RefParameters(C.Structure):
    """Memory layout of the reference's `Parameters` (src/parameters.h:246-318, x86-64) — used to read the
    `params_raw` record of the ref_probe golden files and compare the host's derived constants."""
    _fields_ = [
        ("hdi_on", C.c_bool), ("rseed", C.c_int), ("Temp", C.c_float), ("varR", C.c_float), ("gammaR", C.c_float),
        ("varTheta", C.c_float), ("gammaTheta", C.c_float), ("viscosity", C.c_float), ("freeze_temp", C.c_float),
        ("is_assembly", C.c_bool), ("is_const_conc", C.c_bool), ("out_energy", C.c_bool), ("out_force", C.c_bool),
        ("tub_length", C.c_bool), ("conc", C.c_float), ("alpha", C.c_float), ("dt", C.c_float), ("device", C.c_int),
        ("steps", C.c_longlong), ("firststep", C.c_longlong), ("stride", C.c_longlong),
        ("Ntot", C.c_int), ("Ntr", C.c_int), ("firstrun", C.c_int),
        ("C", C.c_float), ("B_psi", C.c_float), ("B_fi", C.c_float), ("B_theta", C.c_float), ("psi_0", C.c_float),
        ("fi_0", C.c_float), ("theta0_gtp", C.c_float), ("theta0_gdp", C.c_float), ("A_lat", C.c_float), ("A_long", C.c_float),
        ("D_lat", C.c_float), ("D_long", C.c_float), ("seam_coeff", C.c_float),
        ("hydrolysis", C.c_bool), ("khydro", C.c_float), ("hydrostep", C.c_long),
        ("is_wall", C.c_bool), ("rep_h", C.c_float), ("rep_r", C.c_float), ("rep_eps", C.c_float), ("zs", C.c_float * 100),
        ("rep_leftborder", C.c_float),
        ("barrier", C.c_bool), ("a_barr_long", C.c_float), ("r_barr_long", C.c_float), ("w_barr_long", C.c_float),
        ("a_barr_lat", C.c_float), ("r_barr_lat", C.c_float), ("w_barr_lat", C.c_float),
        ("ljpairscutoff", C.c_float), ("ljpairsupdatefreq", C.c_int), ("lj_on", C.c_bool), ("ljscale", C.c_float),
        ("ljsigma6", C.c_float),
    ]


def ref_parameters(raw: np.ndarray) -> RefParameters:
    p = RefParameters()
    C.memmove(C.byref(p), raw.tobytes(), C.sizeof(RefParameters))
    return p


def lists_equal(cnt_a, ent_a, cnt_b, ent_b):
    """compare neighbour lists entry by entry up to their counts"""
    if not np.array_equal(cnt_a, cnt_b):
        return False
    width = min(ent_a.shape[-1], ent_b.shape[-1])
    assert cnt_a.max(initial=0) <= width
    mask = np.arange(width)[None, None, :] < cnt_a[..., None]
    return bool(np.array_equal(np.where(mask, ent_a[..., :width], 0), np.where(mask, ent_b[..., :width], 0)))


def system_from_golden(g, rundir, load_system, ntr=None, extra_overrides=()):
    """rebuild the run directory a golden file was generated from and load it with this repo's host"""
    over = [o for o in str(g["overrides"]).split() if not o.startswith("probe_")] + list(extra_overrides)
    d = rundir(str(g["case"]), runnum=int(g["ntr"]) if ntr is None else ntr, steps=int(g["window"]), stride=100000)
    return load_system(d, over)


# ---------------------------------------------------------------- analysis tools of the reference
def write_dcd(path, frames, stride=1000):
    """float32 [F, N, 3] -> a DCD file in the layout of the reference's dcdio.cpp:98-203 (what its tools read)"""
    import struct
    frames = np.asarray(frames, dtype=np.float32)
    f_count, n = frames.shape[0], frames.shape[1]
    with open(path, "wb") as f:
        f.write(struct.pack("<i4s9if10i", 84, b"CORD", f_count, stride, stride, 0, 0, 0, 0, 0, 0, 1.0, *([0] * 9), 24))
        f.write(struct.pack("<ii", 84, 164))
        f.write(struct.pack("<i", 2) + b"REMARKS CREATED BY dcdio.c".ljust(80, b"\0") + b"REMARKS DATE: test".ljust(80, b"\0"))
        f.write(struct.pack("<iiii", 164, 4, n, 4))
        for fr in frames:
            for k in range(3):
                f.write(struct.pack("<i", 4 * n) + np.ascontiguousarray(fr[:, k]).tobytes() + struct.pack("<i", 4 * n))


def run_reference_analysis(ref_dir, workdir, pdb, xyz_frames, ang_frames, stride):
    """Runs the reference's own tools (oracle/_ref/{temp_calc,p3d22d,disc}) over DCD files written from the frames.
    -> dict(temp=[F-1, 8] printed columns, proj=[F, N, 3], timeline=[F] values, summary=(mean, frames, mean_curled))"""
    import subprocess
    from pathlib import Path
    from mt_b200 import read_dcd
    ref_dir, workdir = Path(ref_dir), Path(workdir)
    workdir.mkdir(parents=True, exist_ok=True)
    fx, fa, fp = workdir / "xyz.dcd", workdir / "ang.dcd", workdir / "proj.dcd"
    write_dcd(fx, xyz_frames, stride)
    write_dcd(fa, ang_frames, stride)
    n_frames = len(xyz_frames)
    out = subprocess.run([str(ref_dir / "temp_calc"), str(pdb), str(fx), str(fa), str(stride)], capture_output=True, text=True, check=True).stdout
    rows = [l.split() for l in out.splitlines() if l and l[0].isdigit()]
    temp = np.array([[float(v) for v in r[1:9]] for r in rows if 2 <= int(r[0]) <= n_frames])
    subprocess.run([str(ref_dir / "p3d22d"), str(pdb), str(fx), str(fa), str(fp)], capture_output=True, text=True, check=True)
    proj = read_dcd(fp)[:n_frames]
    subprocess.run([str(ref_dir / "disc"), str(pdb), str(fp), "timeline"], capture_output=True, text=True, check=True)
    timeline = np.array([float(l.split()[1]) for l in open(str(fp) + ".dat").read().splitlines()][:n_frames])
    summ = subprocess.run([str(ref_dir / "disc"), str(pdb), str(fp)], capture_output=True, text=True, check=True).stdout.split("\n")[-1].split()
    return {"temp": temp, "proj": proj, "timeline": timeline, "summary": (float(summ[0]), int(summ[1]), float(summ[2]))}


def synthetic_disassembly_frames(xyz0, ang0, chain, resid, n_frames=5, seed=3):
    """A lattice whose protofilaments progressively peel outwards: per frame, the top dimers of some protofilaments get a
    larger radius (breaks > 5 nm), theta curls, and everything jitters.  float32 [F, N, 3] x 2 (ang = fi, psi, theta)."""
    rng = np.random.default_rng(seed)
    xyz, ang = [np.asarray(xyz0, np.float32).copy()], [np.asarray(ang0, np.float32).copy()]
    top = resid.max()
    for f in range(1, n_frames):
        x, a = xyz[-1].copy(), ang[-1].copy()
        x += rng.normal(0, 0.03, x.shape).astype(np.float32)
        a += rng.normal(0, 0.01, a.shape).astype(np.float32)
        for c in rng.choice(13, size=5, replace=False):
            cut = top - rng.integers(1, 3 * f + 1)
            sel = (chain == c) & (resid > cut)
            scale = (1.0 + 0.25 * f * (resid[sel] - cut) / 3.0).astype(np.float32)
            x[sel, 0] *= scale
            x[sel, 1] *= scale
            a[sel, 2] += np.float32(0.15 * f)
        xyz.append(x)
        ang.append(a)
    return np.stack(xyz), np.stack(ang)
