"""In-situ analysis: the numpy oracle against the reference's own tools.

CPU: oracle/analysis.py vs the committed outputs of scripts/temp_calc + scripts/disas_speed (tests/golden/
analysis_golden.json) and, where the tools were built (oracle/_ref), vs the tools run live on fresh frames.
GPU (`-m gpu`): the device reductions through the C-ABI vs the oracle (integers and projection bit-exact, displacement
sums to double rounding) and vs the tools run over DCD files of the frames of a real disassembly run."""
import hashlib
import json

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from helpers import run_reference_analysis, synthetic_disassembly_frames
from oracle import analysis as oa

REF = ROOT / "oracle" / "_ref"
HAVE_TOOLS = all((REF / t).exists() for t in ("temp_calc", "p3d22d", "disc"))


def lattice_case(tmp_path, mt_len=40, tail=3):
    from mt_b200 import pdb_labels, structures
    xyz, ang = structures.lattice(mt_len, tail)
    structures.write_pair(xyz, ang, tmp_path / "xyz.pdb", tmp_path / "ang.pdb")
    chain, resid, name1 = pdb_labels(tmp_path / "xyz.pdb")
    x0 = np.array([[a.x, a.y, a.z] for a in xyz], dtype=np.float32)
    a0 = np.array([[a.x, a.y, a.z] for a in ang], dtype=np.float32)
    return chain, resid, name1, x0, a0


def oracle_outputs(fx, fa, chain, resid, name1, stride):
    proj = np.stack([oa.project(x, a) for x, a in zip(fx, fa)])
    pf = np.stack([oa.protofilaments(p, chain, resid, name1) for p in proj])
    temp = np.stack([oa.temperature_scale(oa.temperature_sums(fx[k], fa[k], fx[k - 1], fa[k - 1]), stride) for k in range(1, len(fx))])
    timeline = np.array([oa.timeline_value(p) for p in pf])
    half = np.array([np.sum(chain == c) // 2 for c in range(13)])
    summary = (float(np.float32(2 * int(np.sum(half - pf[-1][:, 2]))) / np.float32(13)),
               float(np.float32(2 * int(np.sum(pf[-1][:, 0] - pf[-1][:, 1]))) / np.float32(13)))
    return proj, pf, temp, timeline, summary


def check_against_tools(out, proj, temp, timeline, summary):
    assert np.array_equal(out["proj"], proj)                    # 3d22d: bit-exact
    assert np.allclose(out["timeline"], timeline, rtol=0, atol=2e-6)  # printed with %f
    assert np.allclose(out["temp"], temp, rtol=1e-9, atol=1.5e-6)     # printed with %f (six decimals)
    assert abs(out["summary"][0] - summary[0]) < 2e-6 and abs(out["summary"][2] - summary[1]) < 2e-6


def test_oracle_matches_committed_tool_outputs(tmp_path):
    g = json.loads((GOLDEN / "analysis_golden.json").read_text())
    chain, resid, name1, x0, a0 = lattice_case(tmp_path, *g["structure"][1:])
    fx, fa = synthetic_disassembly_frames(x0, a0, chain, resid, n_frames=g["n_frames"], seed=g["seed"])
    proj, pf, temp, timeline, summary = oracle_outputs(fx, fa, chain, resid, name1, g["stride"])
    assert hashlib.sha256(np.ascontiguousarray(proj).tobytes()).hexdigest() == g["proj_sha256"]
    assert np.allclose(g["timeline"], timeline, rtol=0, atol=2e-6)
    assert np.allclose(g["temp"], temp, rtol=1e-9, atol=1.5e-6)
    assert abs(g["summary"][0] - summary[0]) < 2e-6 and abs(g["summary"][2] - summary[1]) < 2e-6
    assert (pf[0][:, 0] == 20).all() and (pf[-1][:, 0] < 20).any()  # the frames do break protofilaments


@pytest.mark.skipif(not HAVE_TOOLS, reason="reference analysis tools not built (oracle/_ref)")
@pytest.mark.parametrize("seed,mt_len", [(11, 40), (12, 60)])
def test_oracle_matches_live_tools(tmp_path, seed, mt_len):
    chain, resid, name1, x0, a0 = lattice_case(tmp_path, mt_len, 2)
    fx, fa = synthetic_disassembly_frames(x0, a0, chain, resid, n_frames=4, seed=seed)
    out = run_reference_analysis(REF, tmp_path / "run", tmp_path / "xyz.pdb", fx, fa, 500)
    proj, pf, temp, timeline, summary = oracle_outputs(fx, fa, chain, resid, name1, 500)
    check_against_tools(out, proj, temp, timeline, summary)


# ------------------------------------------------------------------ device reductions
@pytest.mark.gpu
def test_device_analysis_matches_oracle_and_tools(rundir, load_system, tmp_path):
    """Disassembly configuration (curled tails, barrier), 3 trajectories: every 100 steps the in-situ reductions are
    compared with the oracle on the downloaded state, and at the end with the reference's tools run over DCD files
    written from those very frames."""
    from mt_b200 import Engine, pdb_labels
    d = rundir("mt120_disassembly", structure=("lattice", 40, 3), runnum=3)
    s = load_system(d, ["hydrolysis=no", "theta0_gdp=0.35"])
    chain, resid, name1 = pdb_labels(d / "dcd" / "xyz.pdb")
    e = Engine(s)
    e.upload_gtp(np.zeros((3, s.Ntot), dtype=np.int32))  # all GDP: the tails curl and peel
    e.analysis_setup(chain, resid, name1)
    e.analysis_reference()
    frames = [e.coords()]
    for k in range(4):
        e.run(100 * k, 100)
        sums, proj, pf = e.analysis_temperature(), e.analysis_project(), e.analysis_protofilaments()
        c = e.coords()
        frames.append(c)
        for t in range(3):
            xn, an = c[t, :, :3], c[t][:, [3, 5, 4]]
            xo, ao = frames[-2][t, :, :3], frames[-2][t][:, [3, 5, 4]]
            assert np.allclose(sums[t], oa.temperature_sums(xn, an, xo, ao), rtol=1e-12, atol=0)
            p = oa.project(xn, an)
            assert np.array_equal(proj[t], p)
            assert np.array_equal(pf[t], oa.protofilaments(p, chain, resid, name1))
    if HAVE_TOOLS:
        for t in range(3):
            fx = np.stack([f[t, :, :3] for f in frames])
            fa = np.stack([f[t][:, [3, 5, 4]] for f in frames])
            out = run_reference_analysis(REF, tmp_path / f"t{t}", d / "dcd" / "xyz.pdb", fx, fa, 100)
            proj, pf, temp, timeline, summary = oracle_outputs(fx, fa, chain, resid, name1, 100)
            check_against_tools(out, proj, temp, timeline, summary)


@pytest.mark.gpu
def test_device_analysis_on_broken_protofilaments(rundir, load_system):
    """Uploaded synthetic frames with breaks and curls (the simulated window above stays intact): integers bit-exact."""
    from mt_b200 import Engine, pdb_labels
    d = rundir("mt40_ensemble", structure=("lattice", 40, 3), runnum=4)
    s = load_system(d, ["hydrolysis=no"])
    chain, resid, name1 = pdb_labels(d / "dcd" / "xyz.pdb")
    e = Engine(s)
    e.analysis_setup(chain, resid, name1)
    c0 = e.coords()
    fx, fa = synthetic_disassembly_frames(c0[0, :, :3], c0[0][:, [3, 5, 4]], chain, resid, n_frames=5, seed=9)
    c = c0.copy()
    for t in range(4):
        c[t, :, :3] = fx[t + 1]
        c[t][:, [3, 5, 4]] = fa[t + 1]
    e.analysis_reference()
    e.upload_coords(c)
    c = e.coords()
    sums, proj, pf = e.analysis_temperature(), e.analysis_project(), e.analysis_protofilaments()
    broke = 0
    for t in range(4):
        xn, an = c[t, :, :3], c[t][:, [3, 5, 4]]
        p = oa.project(xn, an)
        want = oa.protofilaments(p, chain, resid, name1)
        assert np.array_equal(proj[t], p) and np.array_equal(pf[t], want)
        assert np.allclose(sums[t], oa.temperature_sums(xn, an, c0[t, :, :3], c0[t][:, [3, 5, 4]]), rtol=1e-12, atol=0)
        broke += int((want[:, 0] < 20).sum())
    assert broke > 0
    with pytest.raises(Exception):
        Engine(s).analysis_temperature()  # no setup: fails loudly


@pytest.mark.gpu
@pytest.mark.skipif(not HAVE_TOOLS, reason="reference analysis tools not built (oracle/_ref)")
def test_drop_in_insitu_analysis_vs_tools_on_the_dcd_output(rundir):
    """`mt config.conf` with the extension key insitu_analysis: the per-stride lines it writes beside the DCD files
    against the reference's tools run afterwards over those DCD files (what a user of the reference does today)."""
    import subprocess
    from mt_b200 import HostSystem, read_dcd, workspace
    d = rundir("mt120_disassembly", structure=("lattice", 40, 3), runnum=2, steps=500, stride=100, insitu_analysis="yes")
    r = subprocess.run([str(ROOT / "mt_b200" / "mt"), "config.conf"], cwd=str(d), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    with workspace.chdir(d):
        s = HostSystem("config.conf")
    par = s.par
    for t in range(2):
        fx, fa = read_dcd(d / "dcd" / f"run_{t}.dcd"), read_dcd(d / "dcd" / f"run_{t}.dcd_ang")
        assert fx.shape == (5, 520, 3)  # frames at steps 0 .. 400
        out = run_reference_analysis(REF, d / f"tools_{t}", d / "dcd" / "xyz.pdb", fx, fa, 100)
        mine = np.loadtxt(d / "dcd" / f"run_{t}.dcd.temp.dat")
        assert mine.shape == (4, 9) and list(mine[:, 0]) == [2, 3, 4, 5]
        # both print scaled sums with six decimals; undo the two scalings (the tool's constants are hard-coded, main.cpp:94-97)
        six = 6.0 * 100 * par.dt * 520 * 0.0019872041
        own_scale = np.array([par.gammaR / six, par.gammaTheta / six] + [par.gammaR / (six / 3)] * 3 + [par.gammaTheta / (six / 3)] * 3)
        tool_scale = oa.temperature_scale(np.ones(8), 100)
        assert np.allclose(mine[:, 1:] / own_scale, out["temp"] / tool_scale, rtol=2e-5, atol=1e-7)
        disc = [l.split() for l in (d / "dcd" / f"run_{t}.dcd.disc.dat").read_text().splitlines()]
        assert len(disc) == 5
        assert np.allclose([float(l[1]) for l in disc], out["timeline"], rtol=0, atol=2e-6)
    s.close()


@pytest.mark.gpu
def test_device_analysis_skips_atoms_outside_the_protofilaments(rundir, load_system):
    """Reserve dimers (chain 'X', parked above the lattice) belong to no protofilament: disc.cpp would index out of its
    13-entry arrays; here they are skipped, and the numbers equal those of the lattice alone."""
    from mt_b200 import Engine, pdb_labels
    d = rundir("mt120_constconc", structure=("reserve", 40, 10), runnum=2)
    s = load_system(d, ["hydrolysis=no"])
    chain, resid, name1 = pdb_labels(d / "dcd" / "xyz.pdb")
    assert (chain == -1).sum() > 0
    e = Engine(s)
    e.analysis_setup(chain, resid, name1)
    e.analysis_reference()
    e.run(0, 60)
    pf, proj, c = e.analysis_protofilaments(), e.analysis_project(), e.coords()
    for t in range(2):
        p = oa.project(c[t, :, :3], c[t][:, [3, 5, 4]])
        assert np.array_equal(proj[t], p)
        assert np.array_equal(pf[t], oa.protofilaments(p, chain, resid, name1))
        keep = chain >= 0
        assert np.array_equal(pf[t], oa.protofilaments(p[keep], chain[keep], resid[keep], bytes(np.frombuffer(name1, dtype=np.uint8)[keep])))
