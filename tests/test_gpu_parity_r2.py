"""GPU parity, second set: the BASELINE configs on the inputs they NAME and through the executables.

  config 5  initial/cylinder_xyz.pdb + cylinder_ang.pdb with TEA hydrodynamics: 250-step window against the reference's own
            kernels (ref_probe) crossing the epsilon/beta updates at steps 100 and 200 and 12 list rebuilds, 1 and 4
            trajectories; drop-in `mt` (tea_on yes) against the reference's `mt`.
  config 4  disassembly: non-zero Morse barrier amplitudes (a_barr_long 3.4, a_barr_lat 1.9) + hydrolysis, drop-in `mt`
            against the reference's `mt` over several strides (the serial stride path, on_tubule uploads).
  config 3  constant concentration on the shipped initial/constconc/125 structure, drop-in `mt` against the reference's `mt`.

The reference inputs are committed fixtures (tests/golden/inputs/).  The kernel-level comparison uses the golden file
tests/golden/ref_tea_cylinder_*.npz when it exists and otherwise the live ref_probe binary (and then writes the golden
file into gpurun_out/golden/, from where it is committed): tests/golden/make_ref_golden.py documents the same recipe.

Bars as in test_gpu_parity.py: integer streams and lists bit-exact; |dxyz| <= 1e-3 nm, |dangle| <= 1e-4 rad over the window;
TEA epsilon sums / beta to float rounding (summation order differs from the reference's sequential j loop)."""
import re
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from mt_b200 import Engine, capi

pytestmark = pytest.mark.gpu

INPUTS = GOLDEN / "inputs"
CYL = ("files", str(INPUTS / "cylinder_xyz.pdb"), str(INPUTS / "cylinder_ang.pdb"))
CC125 = ("files", str(INPUTS / "constconc125_xyz.pdb"), str(INPUTS / "constconc125_ang.pdb"))
TEA_KEYS = dict(tea_on="yes", tea_a=1.5, tea_epsilon_freq=100, tea_capricious="yes")  # config/config.conf:37-41
TEA_WINDOW = 250
TEA_MARKS = (100, 200, TEA_WINDOW)


def tea_cylinder_golden(ntr, rundir):
    """coords at the marks, TEA state at the epsilon updates, final seeds, LJ counts at the last rebuild: from the golden
    file, or from the reference's kernels run here"""
    name = f"ref_tea_cylinder_{ntr}.npz"
    p = GOLDEN / name
    if p.exists():
        return dict(np.load(p))
    from oracle import refprobe
    if not refprobe.REF_PROBE.exists():
        pytest.skip(f"{name} not generated and oracle/_ref/ref_probe did not travel with the tree")
    d = rundir("cylinder_tea", structure=CYL, runnum=ntr, steps=TEA_WINDOW, stride=100000, **TEA_KEYS)
    dump = refprobe.run_probe(d, d / "probe.bin", TEA_WINDOW, TEA_WINDOW, ["hydrolysis=no"])
    N = dump.N
    g = {"ntr": np.array(ntr), "window": np.array(TEA_WINDOW), "seeds_end": dump.seeds(TEA_WINDOW),
         "ljcnt_last": dump.ints("ljcnt", 240), "coords0": dump.coords(0), "forces0": dump.forces(0)}
    for m in TEA_MARKS:
        g[f"coords_{m}"] = dump.coords(m)
    for m in (0, 100, 200):
        g[f"tea_ci_{m}"] = dump.floats("tea_ci", m, (ntr * N, 4))
        g[f"tea_eps_{m}"] = dump.floats("tea_eps", m, (ntr * N,))
        g[f"tea_beta_{m}"] = dump.floats("tea_beta", m, (ntr,))
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(out / name, **g)
    return g


@pytest.mark.parametrize("ntr", [1, 4])
def test_tea_cylinder_250_steps_against_reference_kernels(ntr, rundir, load_system):
    g = tea_cylinder_golden(ntr, rundir)
    d = rundir("cylinder_tea", structure=CYL, runnum=ntr, steps=TEA_WINDOW, stride=100000, **TEA_KEYS)
    s = load_system(d, ["hydrolysis=no"])
    assert s.Ntot == 520 and s.par.tea_on
    e = Engine(s)
    assert np.array_equal(e.coords()[..., :6], g["coords0"][..., :6])
    worst = {}
    step = 0
    for mark in TEA_MARKS:
        # the epsilon/beta update of step `step` happens inside the first step of this window: run one step, compare, go on
        e.run(step, 1)
        if step in (0, 100, 200):
            ci, eps, beta = e.tea_state()
            ref_eps, ref_beta, ref_ci = g[f"tea_eps_{step}"], g[f"tea_beta_{step}"], g[f"tea_ci_{step}"]
            # per-bead epsilon sums are O(N) sums of O(1e-2) terms in a different order: relative 1e-4 of the largest entry
            assert np.abs(eps - ref_eps).max() <= 1e-4 * np.abs(ref_eps).max() + 1e-6, (step, np.abs(eps - ref_eps).max())
            assert np.allclose(beta, ref_beta, rtol=2e-5, atol=0), (step, beta, ref_beta)
            assert np.allclose(ci[:, :3], ref_ci[:, :3], rtol=1e-4, atol=1e-6), (step, np.abs(ci[:, :3] - ref_ci[:, :3]).max())
        e.run(step + 1, mark - step - 1)
        step = mark
        c, r = e.coords(), g[f"coords_{mark}"]
        dx, da = np.abs(c[..., :3] - r[..., :3]).max(), np.abs(c[..., 3:6] - r[..., 3:6]).max()
        worst[mark] = (float(dx), float(da))
        assert dx < 1e-3 and da < 1e-4, (mark, dx, da)
    print(f"TEA cylinder x{ntr}: |dxyz|, |dang| at steps {worst}")
    assert np.array_equal(e.rng_state(), g["seeds_end"])
    cnt, _ = e.download_list(capi.LIST_LJ)  # the list of the rebuild at step 240
    assert np.array_equal(cnt, g["ljcnt_last"])


def _both_executables(d_ref, timeout=900):
    """run the reference's mt and this repo's mt on copies of one run directory -> (stdout_ref, stdout_own, d_own)"""
    from oracle import refprobe
    if not refprobe.REF_MT.exists():
        pytest.skip("oracle/_ref/mt did not travel with the tree")
    d_own = d_ref.parent / (d_ref.name + "_own")
    shutil.copytree(d_ref, d_own)
    _, out_ref = refprobe.run_reference_mt(d_ref)
    r = subprocess.run([str(ROOT / "mt_b200" / "mt"), "config.conf"], cwd=str(d_own), capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stderr[-2000:]
    return out_ref, r.stdout, d_own


def _compare_dcd(d_ref, d_own, ntr, frames, n, tol_xyz=1e-3, tol_ang=1e-4):
    import mt_b200
    worst = [0.0, 0.0]
    for t in range(ntr):
        for k, (suffix, tol) in enumerate(((".dcd", tol_xyz), (".dcd_ang", tol_ang))):
            a = mt_b200.read_dcd(d_own / "dcd" / f"run_{t}{suffix}")
            b = mt_b200.read_dcd(d_ref / "dcd" / f"run_{t}{suffix}")
            assert a.shape == b.shape == (frames, n, 3), (a.shape, b.shape)
            # absolute tolerance, plus float rounding relative to the distance from the origin: a dimer the reference's
            # change_conc() drops ON another one (same 2-nm grid point, updater.cpp:118-119) is shot out to ~1e8 nm by the
            # r^-6 repulsion in both executables, where one ulp is 8 nm
            # (such a dimer and what it hits on its way out leave the 160 x 80 nm cylinder by orders of magnitude; their
            # trajectories amplify rounding like any collision at r -> 0 and are compared to 1e-3 relative, everything
            # that stays in the box to the absolute tolerance;
            # their angles, driven by forces of 1e7 and more, are not compared at all)
            dist = np.linalg.norm(mt_b200.read_dcd(d_ref / "dcd" / f"run_{t}.dcd"), axis=-1, keepdims=True)
            flung = np.broadcast_to(dist > 2000.0, a.shape)
            err = np.where(flung, 0.0, np.abs(a - b))
            worst[k] = max(worst[k], float(err.max()))
            assert err.max() < tol, (t, suffix, err.max())
            assert flung.mean() < 0.05
            if suffix == ".dcd":
                assert (np.abs(a - b) <= 1e-3 * np.maximum(dist, 1.0))[flung].all()
    return worst


def _energies(out):
    pat = re.compile(r"Energies\[(\d+)\]:\s+(.*)")
    return np.array([[float(x) for x in m.group(2).split()] for m in map(pat.match, out.splitlines()) if m])


def test_drop_in_tea_executable_vs_reference_executable(rundir):
    """BASELINE config 5 through the executables: `mt` with tea_on yes on initial/cylinder_{xyz,ang}.pdb, 2 trajectories,
    frames at steps 0/150/300/450 (epsilon/beta updates at 100, 200, 300, 400 in between)."""
    d_ref = rundir("cylinder_tea", structure=CYL, runnum=2, steps=451, stride=150, **TEA_KEYS)
    out_ref, out_own, d_own = _both_executables(d_ref)
    worst = _compare_dcd(d_ref, d_own, 2, 4, 520)
    e_ref, e_own = _energies(out_ref), _energies(out_own)
    assert e_ref.shape == e_own.shape == (8, 7)
    assert np.allclose(e_own, e_ref, rtol=1e-4, atol=2e-2)
    assert re.findall(r"tubule\[\d\]: \d+", out_own) == re.findall(r"tubule\[\d\]: \d+", out_ref)
    print(f"drop-in TEA: worst |dxyz| {worst[0]:.2e} nm, |dang| {worst[1]:.2e} rad over 450 steps")


def test_drop_in_disassembly_barrier_vs_reference_executable(rundir):
    """BASELINE config 4 through the executables: initial/xyz_120 + ang_120 (regenerated byte-identically, 1560 monomers, curled
    tails), barrier yes with a_barr_long 3.4 / a_barr_lat 1.9, hydrolysis every 100 steps, 2 trajectories, 4 strides: with
    non-zero amplitudes the on-tubule flags classified at every stride feed back into the forces (compute_cuda.cu:1189-1193,
    :259-262, :454-457), so the loop takes its serial stride path and uploads them."""
    d_ref = rundir("mt120_disassembly", runnum=2, steps=850, stride=200)
    out_ref, out_own, d_own = _both_executables(d_ref)
    worst = _compare_dcd(d_ref, d_own, 2, 5, 1560)
    ev = re.compile(r"\*\*\* .* \*\*\*")
    assert "Hydrolysis occured" in out_ref and ev.findall(out_own) == ev.findall(out_ref)
    assert re.findall(r"tubule\[\d\]: \d+", out_own) == re.findall(r"tubule\[\d\]: \d+", out_ref)
    e_ref, e_own = _energies(out_ref), _energies(out_own)
    assert e_ref.shape == e_own.shape == (10, 7)
    assert np.allclose(e_own, e_ref, rtol=1e-4, atol=2e-2)
    assert (d_own / "mt_len.dat").read_text() == (d_ref / "mt_len.dat").read_text()
    print(f"drop-in disassembly: worst |dxyz| {worst[0]:.2e} nm, |dang| {worst[1]:.2e} rad over 800 steps")


def test_drop_in_const_conc_on_shipped_structure_vs_reference_executable(rundir):
    """BASELINE config 3 on the reference's shipped fallback structure initial/constconc/125 (520 lattice + 520 reserve
    monomers): walls, constant concentration (insertions consume the shared rand() stream), LJ list; 2 trajectories."""
    # conc 200: in this (equilibrated) structure few monomers pass the reference's on-tubule test, so the template's conc 30 is
    # already exceeded by the 260 free dimers and nothing would be inserted; 200 muM needs >= 386 free dimers
    d_ref = rundir("mt120_constconc", structure=CC125, runnum=2, steps=650, stride=200, conditions={"conc": 200})
    out_ref, out_own, d_own = _both_executables(d_ref)
    ins = re.compile(r"New x,y coordinates for extra particle: .*")
    assert ins.findall(out_ref) and ins.findall(out_own) == ins.findall(out_ref)
    conc = re.compile(r"Concentration for tajectory\[\d+\]: .*")
    assert conc.findall(out_own) == conc.findall(out_ref)
    assert re.findall(r"tubule\[\d\]: \d+", out_own) == re.findall(r"tubule\[\d\]: \d+", out_ref)
    worst = _compare_dcd(d_ref, d_own, 2, 4, 1040)
    print(f"drop-in const-conc (constconc/125): worst |dxyz| {worst[0]:.2e} nm, |dang| {worst[1]:.2e} rad over 600 steps")


@pytest.mark.parametrize("case", ["mt40", "tea_cylinder", "constconc125"])
def test_reference_host_with_stub_compute_equals_own_executable(case, rundir):
    """The boundary proven from the reference's side: oracle/_ref/mt_stub is the reference's OWN main/preparator/updater/... with
    compute() replaced by oracle/compute_b200_stub.cpp on top of libmaddy_b200.so.  Same kernels, same
    rand() stream, so its frames equal this repo's `mt` bit for bit (whatever windows each host loop chooses), and both agree
    with the reference's `mt` to the window tolerance."""
    import mt_b200
    stub = ROOT / "oracle" / "_ref" / "mt_stub"
    if not stub.exists():
        pytest.skip("oracle/_ref/mt_stub did not travel with the tree")
    if case == "mt40":
        d_ref, ntr, frames, n = rundir("mt40_single", runnum=2, steps=800, stride=350), 2, 3, 520
    elif case == "tea_cylinder":
        d_ref, ntr, frames, n = rundir("cylinder_tea", structure=CYL, runnum=2, steps=301, stride=150, **TEA_KEYS), 2, 3, 520
    else:
        d_ref, ntr, frames, n = rundir("mt120_constconc", structure=CC125, runnum=2, steps=450, stride=200), 2, 3, 1040
    out_ref, out_own, d_own = _both_executables(d_ref)
    d_stub = d_ref.parent / (d_ref.name + "_stub")
    shutil.copytree(d_own, d_stub, ignore=shutil.ignore_patterns("*.dcd", "*.dcd_ang", "mt_len.dat", "result_*.pdb", "hydrolysis.pdb", "ensemble.dat"))
    r = subprocess.run([str(stub), "config.conf"], cwd=str(d_stub), capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, (r.stdout[-1000:], r.stderr[-1000:])
    for t in range(ntr):
        for suffix in (".dcd", ".dcd_ang"):
            a = mt_b200.read_dcd(d_stub / "dcd" / f"run_{t}{suffix}")
            b = mt_b200.read_dcd(d_own / "dcd" / f"run_{t}{suffix}")
            assert a.shape == b.shape == (frames, n, 3)
            assert np.array_equal(a, b), (case, t, suffix, np.abs(a - b).max())
    _compare_dcd(d_ref, d_stub, ntr, frames, n)
    for pat in (r"\*\*\* .* \*\*\*", r"tubule\[\d\]: \d+", r"New x,y coordinates for extra particle: .*", r"Energies\[\d+\]:.*"):
        assert re.findall(pat, r.stdout) == re.findall(pat, out_own), pat
    assert (d_stub / "mt_len.dat").read_text() == (d_own / "mt_len.dat").read_text() == (d_ref / "mt_len.dat").read_text()
