"""The CPU oracle pinned against the REFERENCE'S OWN KERNELS (tests/golden/ref_*.npz, dumped on a B200 by
oracle/_ref/ref_probe = unmodified reference translation units + oracle/ref_probe.cu).

Exact: seed table, RNG state after the window, LJ lists, bond lists.  Toleranced (libm here vs MUFU
approximations under -use_fast_math there): forces, energies, one integrator step, the step window."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_npz
from helpers import lists_equal, system_from_golden
from mt_b200.capi import as_ptr
from oracle.pyoracle import OracleState

CASES = ["mt40", "mt120_gdp_barrier", "reserve_walls", "mt40_static"]

# tolerances vs the reference's CUDA kernels (observed maxima in comments, see profiles/r1_parity.md)
F_ATOL, F_RTOL = 5e-3, 2e-5      # observed |dF| <= 2.1e-3 on |F| up to 5e2
E_ATOL = 5e-5                    # per-monomer energy, observed <= 7e-6
STEP_ATOL_XYZ, STEP_ATOL_ANG = 5e-5, 2e-6  # one step: a few float ulps at |x| ~ 5e2


def oracle_from_golden(g, rundir, load_system):
    s = system_from_golden(g, rundir, load_system)
    o = OracleState(s)
    o._keep = (np.ascontiguousarray(g["gtp"]), np.ascontiguousarray(g["ontub"]))
    o.top.gtp = as_ptr(o._keep[0], C.c_int)
    o.top.on_tubule_cur = as_ptr(o._keep[1], C.c_int)
    return s, o


@pytest.mark.parametrize("name", CASES)
def test_seed_table_and_initial_state(name, rundir, load_system):
    g = golden_npz(name)
    s, o = oracle_from_golden(g, rundir, load_system)
    assert np.array_equal(o.rng, g["seeds0"])
    assert np.array_equal(o.coords[..., :6], g["coords0"][..., :6])


@pytest.mark.parametrize("name", CASES)
def test_lists_bit_exact(name, rundir, load_system):
    g = golden_npz(name)
    s, o = oracle_from_golden(g, rundir, load_system)
    o.rebuild_lj()
    assert lists_equal(o.lj_count, o.lj, g["ljcnt0"], g["lj0"])
    if s.par.is_assembly:
        o.rebuild_bonds()
    assert lists_equal(o.long_count, o.long, g["longcnt0"], g["long0"])
    assert lists_equal(o.lat_count, o.lat, g["latcnt0"], g["lat0"])


@pytest.mark.parametrize("name", CASES)
def test_forces_and_energies(name, rundir, load_system):
    g = golden_npz(name)
    s, o = oracle_from_golden(g, rundir, load_system)
    o.rebuild_lj()
    if s.par.is_assembly:
        o.rebuild_bonds()
    F = o.force().copy()
    assert np.allclose(F[..., :6], g["forces0"][..., :6], rtol=F_RTOL, atol=F_ATOL), np.abs(F[..., :6] - g["forces0"][..., :6]).max()
    E = o.energies()
    assert np.allclose(E, g["energy0"], rtol=1e-6, atol=E_ATOL), np.abs(E - g["energy0"]).max()
    # second step of the reference trajectory (thermalised coordinates, same lists)
    o.coords[:] = g["coords1"]
    F1 = o.force().copy()
    assert np.allclose(F1[..., :6], g["forces1"][..., :6], rtol=F_RTOL, atol=F_ATOL), np.abs(F1[..., :6] - g["forces1"][..., :6]).max()


@pytest.mark.parametrize("name", CASES)
def test_one_step_and_window(name, rundir, load_system):
    g = golden_npz(name)
    s, o = oracle_from_golden(g, rundir, load_system)
    o.run(0, 1)
    assert np.abs(o.coords[..., :3] - g["coords1"][..., :3]).max() < STEP_ATOL_XYZ
    assert np.abs(o.coords[..., 3:6] - g["coords1"][..., 3:6]).max() < STEP_ATOL_ANG
    window = int(g["window"])
    o.run(1, window - 1)
    # integer stream: identical consumption (only free, non-extra monomers draw; 8 draws per step)
    assert np.array_equal(o.rng, g["seeds_end"])
    assert np.abs(o.coords[..., :3] - g["coords_end"][..., :3]).max() < 1e-3   # stated fp32 window tolerance (nm)
    assert np.abs(o.coords[..., 3:6] - g["coords_end"][..., 3:6]).max() < 1e-4  # rad


def test_force_is_minus_energy_gradient(rundir, load_system):
    """compute_kernel vs energy_kernel are independent code in the reference; for the distance-dependent terms
    (harmonic, Morse, LJ) the force must be the negative gradient of the summed per-monomer energies."""
    s = load_system(rundir(runnum=1), ["B_psi=0", "B_fi=0", "B_theta=0", "hydrolysis=no"])
    o = OracleState(s)
    rng = np.random.default_rng(5)
    o.coords[0, :, :6] += (rng.normal(0, 0.02, (520, 6)) * np.array([1, 1, 1, .2, .2, .2])).astype(np.float32)
    o.rebuild_lj()
    o.rebuild_bonds()
    F = o.force().copy()

    def etot():
        e = o.energies()[0]
        return e[:, 0].sum() + e[:, 1].sum() + e[:, 2].sum() + e[:, 6].sum()

    for i, k in ((100, 0), (100, 2), (101, 3), (300, 5), (57, 4), (57, 1), (400, 3), (402, 5), (27, 4)):
        h = 2e-3
        c0 = o.coords[0, i, k]
        o.coords[0, i, k] = c0 + h
        ep = etot()
        o.coords[0, i, k] = c0 - h
        em = etot()
        o.coords[0, i, k] = c0
        assert abs(F[0, i, k] + (ep - em) / (2 * h)) < 0.02 + 2e-3 * abs(F[0, i, k]), (i, k)


def test_newton_third_law_and_list_symmetry(rundir, load_system):
    s = load_system(rundir(runnum=2), ["hydrolysis=no"])
    o = OracleState(s)
    o.rebuild_lj()
    o.rebuild_bonds()
    o.run(0, 25)
    o.rebuild_lj()
    o.rebuild_bonds()
    F = o.force()
    # walls off, pair potentials only: the xyz forces of a trajectory sum to ~0 (updater.cpp:45-57 prints this norm)
    assert np.abs(F[..., :3].sum(axis=1)).max() < 0.05
    for t in range(2):
        A = np.zeros((520, 520), dtype=bool)
        for i in range(520):
            A[i, o.lj[t, i, :o.lj_count[t, i]]] = True
        assert np.array_equal(A, A.T) and not A.diagonal().any()
        assert all(np.all(np.diff(o.lj[t, i, :o.lj_count[t, i]]) > 0) for i in range(520))  # ascending j


def test_extras_and_fixed_are_inert(rundir, load_system):
    d = rundir("mt120_constconc", structure=("reserve", 20, 6), runnum=1)
    s = load_system(d, ["hydrolysis=no", "is_const_conc=no"])
    o = OracleState(s)
    before = o.coords.copy()
    rng0 = o.rng.copy()
    o.run(0, 3)
    extra = s.extra[0].astype(bool)
    fixed = s.fixed.astype(bool)
    assert extra.sum() == 78 and fixed.sum() == 26
    assert np.array_equal(o.coords[0, extra | fixed], before[0, extra | fixed])
    n = s.Ntot
    still = np.where(extra | fixed)[0]
    assert np.array_equal(o.rng[0, still], rng0[0, still]) and np.array_equal(o.rng[1, still], rng0[1, still])
    assert (o.lj_count[0, extra] == 0).all() and (o.lat_count[0, extra] == 0).all()
    assert not np.array_equal(o.rng[0, ~(extra | fixed)], rng0[0, ~(extra | fixed)])


def test_tea_against_reference(rundir, load_system):
    g = golden_npz("tea")
    s, o = oracle_from_golden(g, rundir, load_system)
    o.rebuild_lj()
    o.rebuild_bonds()
    F = o.force().copy()
    assert np.allclose(F[..., :6], g["forces0"][..., :6], rtol=F_RTOL, atol=F_ATOL)
    assert o.tea_update() == 0
    assert np.allclose(o.tea_ci, g["tea_ci0"], rtol=2e-5, atol=1e-6)
    assert np.allclose(o.tea_eps, g["tea_eps0"], rtol=2e-5, atol=1e-5)
    assert np.allclose(o.tea_beta, g["tea_beta0"], rtol=1e-6)
    o.tea_integrate()
    assert np.abs(o.coords[..., :3] - g["coords1"][..., :3]).max() < 1e-4
    assert np.abs(o.coords[..., 3:6] - g["coords1"][..., 3:6]).max() < 1e-5
    # the TEA path advances BOTH streams of EVERY bead, fixed ones included (bdhitea_kernel.cu:22,:194)
    for step in range(1, int(g["window"])):
        o.force()
        if step % s.par.tea_epsilon_freq == 0:
            o.tea_update()
        o.tea_integrate()
    assert np.array_equal(o.rng, g["seeds_end"])
    assert np.abs(o.coords[..., :3] - g["coords_end"][..., :3]).max() < 1e-3
