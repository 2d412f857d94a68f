"""Phase-by-phase comparison of this repo's kernels with the reference's own kernels (ref_probe) on a GPU box.

usage: python tools/parity_report.py [config_name] [ntr] [window] [overrides...]
Prints statistics only (the asserting version lives in tests/test_gpu_parity.py).
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import Engine, HostSystem, capi, workspace  # noqa: E402
from oracle import refprobe  # noqa: E402
from oracle.pyoracle import OracleState  # noqa: E402


def stats(name, a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.abs(a - b)
    scale = np.abs(b).max() if b.size else 0
    print(f"  {name:28s} max|d|={d.max():.3e}  mean|d|={d.mean():.3e}  max|ref|={scale:.3e}  rel={d.max() / (scale + 1e-30):.3e}")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "mt40_single"
    ntr = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    window = int(sys.argv[3]) if len(sys.argv) > 3 else 60
    overrides = sys.argv[4:]
    probe_over = [o for o in overrides]
    host_over = [o for o in overrides if not o.startswith("probe_")]
    d = Path(tempfile.mkdtemp(prefix="parity_"))
    workspace.make_baseline_rundir(d, name, runnum=ntr, steps=window, stride=100000)
    print(f"== {name} ntr={ntr} window={window} overrides={overrides}")
    dump = refprobe.run_probe(d, d / "probe.bin", window, window, probe_over + ["hydrolysis=no"])
    with workspace.chdir(d):
        sysm = HostSystem("config.conf", host_over + ["hydrolysis=no"])
    N = sysm.Ntot
    # host topology vs reference host
    harm_ref = np.frombuffer(dump._raw("harm", -1), dtype=np.int32).reshape(N, dump.maxH)
    print("host: harmonic equal", np.array_equal(harm_ref, sysm.harmonic), " mon_type equal",
          np.array_equal(np.frombuffer(dump._raw("montype", -1), dtype=np.int32), sysm.mon_type),
          " fixed equal", np.array_equal(np.frombuffer(dump._raw("fixed", -1), dtype=np.uint8), sysm.fixed),
          " extra equal", np.array_equal(np.frombuffer(dump._raw("extra", -1), dtype=np.uint8).reshape(ntr, N), sysm.extra))
    stats("host coords r", sysm.coords, dump.coords(-1, "r_host"))
    gtp = np.frombuffer(dump._raw("gtp", -1), dtype=np.int32).reshape(ntr, N)
    ontub = np.frombuffer(dump._raw("ontub", -1), dtype=np.int32).reshape(ntr, N)
    seeds_ref = dump.seeds(-1)
    eng = Engine(sysm)
    eng.upload_gtp(gtp)
    eng.upload_on_tubule(ontub)
    print("seed table equal:", np.array_equal(seeds_ref, eng.rng_state()))
    print("initial coords equal:", np.array_equal(dump.coords(0), eng.coords()))

    # ---- step-granular walk, restarting each phase from the reference's state
    freq = sysm.par.ljpairsupdatefreq
    worst = {}
    for step in range(window):
        ref_c = dump.coords(step)
        eng.upload_coords(ref_c)
        if step % freq == 0:
            if sysm.par.lj_on:
                eng.rebuild_lj()
                cnt, ent = eng.download_list(capi.LIST_LJ)
                rc, re = dump.lj(step) if step in dump.steps("lj") else (dump.ints("ljcnt", step), None)
                ok = np.array_equal(cnt, rc) and (re is None or all(
                    np.array_equal(ent[t, i, :cnt[t, i]], re[t, i, :rc[t, i]]) for t in range(ntr) for i in range(N)))
                print(f"step {step}: LJ list exact = {ok}  (mean count {cnt.mean():.1f}, max {cnt.max()})")
            if sysm.par.is_assembly:
                eng.rebuild_bonds()
                lc, le = eng.download_list(capi.LIST_LONGITUDINAL)
                tc, te = eng.download_list(capi.LIST_LATERAL)
                rlc, rle, rtc, rte = dump.bonds(step)
                okl = np.array_equal(lc, rlc) and all(np.array_equal(le[t, i, :lc[t, i]], rle[t, i, :rlc[t, i]]) for t in range(ntr) for i in range(N))
                okt = np.array_equal(tc, rtc) and all(np.array_equal(te[t, i, :tc[t, i]], rte[t, i, :rtc[t, i]]) for t in range(ntr) for i in range(N))
                print(f"step {step}: longitudinal exact = {okl}, lateral exact = {okt}  (counts {np.bincount(lc.ravel())} / {np.bincount(tc.ravel())})")
                if not okl:
                    bad = np.argwhere(lc != rlc)[:5]
                    print("   first long mismatches", bad.tolist())
                if not okt:
                    bad = np.argwhere(tc != rtc)[:5]
                    print("   first lat mismatches", bad.tolist(), [(te[t, i, :4].tolist(), rte[t, i, :4].tolist()) for t, i in bad[:3]])
        eng.force()
        F = eng.forces()
        RF = dump.forces(step)
        e_tr, e_mono = eng.energies(per_monomer=True)
        RE = dump.energy(step)
        eng.integrate()
        C1 = eng.coords()
        R1 = dump.coords(step + 1)
        for key, a, b in (("F xyz", F[..., :3], RF[..., :3]), ("F ang", F[..., 3:6], RF[..., 3:6]), ("E mono", e_mono, RE),
                          ("step xyz", C1[..., :3], R1[..., :3]), ("step ang", C1[..., 3:6], R1[..., 3:6])):
            dd = np.abs(a.astype(np.float64) - b.astype(np.float64)).max()
            worst[key] = max(worst.get(key, 0), dd)
        if step in (0, 1, window - 1):
            print(f"step {step}:")
            stats("forces xyz", F[..., :3], RF[..., :3])
            stats("forces ang", F[..., 3:6], RF[..., 3:6])
            stats("energy per monomer", e_mono, RE)
            stats("energy per traj", e_tr, RE.sum(axis=1))
            stats("one step xyz", C1[..., :3], R1[..., :3])
            stats("one step ang", C1[..., 3:6], R1[..., 3:6])
            print("   one-step coords bit-exact fraction:", float((C1[..., :6] == R1[..., :6]).mean()))
    print("worst single-phase deviations over the window:", {k: float(f"{v:.3e}") for k, v in worst.items()})
    print("rng state after window equal:", np.array_equal(dump.seeds(window), eng.rng_state()))

    # ---- free-running fused trajectory vs reference trajectory
    eng2 = Engine(sysm)
    eng2.upload_gtp(gtp)
    eng2.upload_on_tubule(ontub)
    eng2.run(0, window)
    stats("fused run xyz (end)", eng2.coords()[..., :3], dump.coords(window)[..., :3])
    stats("fused run ang (end)", eng2.coords()[..., 3:6], dump.coords(window)[..., 3:6])
    # fused == step-granular (bitwise)
    eng3 = Engine(sysm)
    eng3.upload_gtp(gtp)
    eng3.upload_on_tubule(ontub)
    for step in range(window):
        if step % freq == 0:
            eng3.rebuild_lj()
            if sysm.par.is_assembly:
                eng3.rebuild_bonds()
        eng3.force()
        eng3.integrate()
    print("fused == step-granular bitwise:", np.array_equal(eng2.coords(), eng3.coords()), np.array_equal(eng2.rng_state(), eng3.rng_state()))
    # ---- CPU oracle against the reference (pins the oracle)
    o = OracleState(sysm)
    o.top.gtp = capi.as_ptr(np.ascontiguousarray(gtp), __import__("ctypes").c_int)
    o.top.on_tubule_cur = capi.as_ptr(np.ascontiguousarray(ontub), __import__("ctypes").c_int)
    o.coords[:] = dump.coords(0)
    o.rebuild_lj()
    if sysm.par.is_assembly:
        o.rebuild_bonds()
    stats("oracle forces xyz", o.force()[..., :3], dump.forces(0)[..., :3])
    stats("oracle forces ang", o.forces[..., 3:6], dump.forces(0)[..., 3:6])
    stats("oracle energy", o.energies(), dump.energy(0))
    np.savez_compressed(d / "summary.npz", worst=np.array(list(worst.values())))
    print("rundir:", d)


if __name__ == "__main__":
    main()
