"""SURVEY 8 f1 — the host events of the stride block on the device side of the boundary (all through the C-ABI):

  * mt_length()'s on-tubule classification (updater.cpp:154-227) evaluated by the snapshot (MADDY_SNAP_ONTUBULE), exact;
  * change_conc()'s insertions (updater.cpp:97-152) handed over as sparse records (maddy_insert_dimers);
  * the drop-in loop with both == the reference's serial stride block (MADDY_HOST_EVENTS=1), bit for bit.

The checker for the classification is the host's own mt_length() (libm cosf / sqrt of this process) — the function the
reference's host executes; the drop-in executables are compared with the reference's `mt` in test_gpu_parity*.py.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mt_b200 import Engine, capi  # noqa: E402


def _adversarial_state(s, seed):
    """coordinates whose radius and theta sit on, next to and far from every threshold of updater.cpp:161"""
    rng = np.random.default_rng(seed)
    c = np.array(s.coords, dtype=np.float32).reshape(s.Ntr, s.Ntot, 7).copy()
    n = c.shape[0] * c.shape[1]
    flat = c.reshape(n, 7)
    # theta: every crossing of cos(theta) = cos(1) within +-3 turns, +-40 ulps around it, both signs; plus bulk values
    cross = np.array([k * 2 * np.pi + sgn * 1.0 for k in range(-3, 4) for sgn in (-1.0, 1.0)], dtype=np.float64)
    base = rng.choice(cross, size=n).astype(np.float32)
    ulps = rng.integers(-40, 41, size=n).astype(np.int32)
    theta = (base.view(np.int32) + np.where(base >= 0, ulps, -ulps)).view(np.float32)
    bulk = rng.random(n) < 0.3
    theta = np.where(bulk, rng.uniform(-20.0, 20.0, n).astype(np.float32), theta)
    flat[:, 4] = theta
    # radius: around R_MT + R_THRES = 24.12 and around 1.0, a few ulps either side, random direction; plus bulk values
    which = rng.integers(0, 4, size=n)
    rad = np.where(which == 0, 24.12, np.where(which == 1, 1.0, rng.uniform(0.0, 40.0, n))).astype(np.float32)
    rad = (rad.view(np.int32) + rng.integers(-6, 7, size=n).astype(np.int32)).view(np.float32)
    phi = rng.uniform(0, 2 * np.pi, n)
    flat[:, 0] = (rad * np.cos(phi)).astype(np.float32)
    flat[:, 1] = (rad * np.sin(phi)).astype(np.float32)
    exact = rng.random(n) < 0.2  # exact representable radii on an axis: the compare itself is on the boundary
    flat[exact, 0] = rad[exact]
    flat[exact, 1] = 0.0
    return c


def test_device_on_tubule_classification_is_exact(rundir, load_system):
    """MADDY_SNAP_ONTUBULE on adversarial states == the host's mt_length() on the same coordinates, flag for flag; APPLY makes
    them the live flags (the forces that follow read them)."""
    s = load_system(rundir("mt120_disassembly", runnum=24))
    e = Engine(s)
    assert capi.lib.maddy_has_exact_on_tubule(e._h) == 1
    for seed in range(3):
        c = _adversarial_state(s, seed)
        e.upload_coords(c)
        e.snapshot_begin(coords=True, energies=False, on_tubule=True)
        ln, und = e.snapshot_tubule_lengths()  # counts first
        snap = e.snapshot_end()
        assert und == 0
        s.coords[...] = snap["coords"]
        want_len = s.mt_length(1000)
        want = np.array(s.on_tubule_cur).reshape(s.Ntr, s.Ntot)
        assert np.array_equal(snap["coords"], c)
        assert np.array_equal(snap["on_tubule"], want), int((snap["on_tubule"] != want).sum())
        assert np.array_equal(snap["mt_len"], want_len) and np.array_equal(ln, want_len)
        assert 0 < want.sum() < want.size
    # APPLY == upload of the same flags: forces with a non-zero barrier are bitwise equal
    ref = Engine(s)
    ref.upload_coords(c)
    ref.upload_on_tubule(want)
    e.snapshot_begin(coords=False, energies=False, apply_on_tubule=True)
    e.snapshot_end()
    for eng in (e, ref):
        eng.rebuild_lj()
        eng.rebuild_bonds()
        eng.force()
    assert np.array_equal(e.forces(), ref.forces())
    # a bending angle several turns away is reported, not guessed
    c[0, 5, 4] = 40.0
    c[0, 5, 0], c[0, 5, 1] = 8.0, 0.0
    e.upload_coords(c)
    e.snapshot_begin(coords=False, energies=False, on_tubule=True)
    _, und = e.snapshot_tubule_lengths()
    assert und != 0
    from mt_b200 import MaddyError
    with pytest.raises(MaddyError):
        e.snapshot_end()


def test_insert_dimers_equals_full_upload(rundir, load_system):
    """maddy_insert_dimers == maddy_upload_extra + maddy_upload_coords of the arrays change_conc() modified: state, lists and
    the following 60 steps bitwise."""
    s = load_system(rundir("mt120_constconc", runnum=6, steps=200), ["hydrolysis=no"])
    a, b = Engine(s), Engine(s)
    for eng in (a, b):
        eng.run(0, 40)
        eng.rebuild_lj()
        eng.rebuild_bonds()
    c = a.coords()
    extra = np.array(s.extra).reshape(s.Ntr, s.Ntot).copy()
    assert extra.any()
    rng = np.random.default_rng(1)
    idx, rec = [], []
    for t in (0, 2, 2, 5):
        free = np.flatnonzero(extra[t])
        i = int(free[0])
        assert i % 2 == 0 and extra[t, i + 1]
        x, y = np.float32(rng.integers(-30, 30) * 2.0), np.float32(rng.integers(-30, 30) * 2.0)
        z = np.float32(160.0 + 12.0)
        extra[t, i] = extra[t, i + 1] = 0
        c[t, i, :3] = (x, y, z)
        c[t, i + 1, :3] = (x, y, z + np.float32(4.0))
        idx.append(t * s.Ntot + i)
        rec.append((x, y, z, z + np.float32(4.0)))
    a.upload_extra(extra)
    a.upload_coords(c)
    b.insert_dimers(idx, rec)
    assert np.array_equal(a.coords(), b.coords())
    for eng in (a, b):
        eng.run(40, 60, skip_first_rebuild=True)
    assert np.array_equal(a.coords(), b.coords()) and np.array_equal(a.rng_state(), b.rng_state())
    for kind in (capi.LIST_LJ, capi.LIST_LONGITUDINAL, capi.LIST_LATERAL):
        (ca, ea), (cb, eb) = a.download_list(kind), b.download_list(kind)
        assert np.array_equal(ca, cb) and np.array_equal(ea, eb)
    assert np.array_equal(a.energies(), b.energies())
    from mt_b200 import MaddyError
    with pytest.raises(MaddyError):
        b.insert_dimers([s.Ntot - 1], [(0, 0, 0, 0)])  # not the first monomer of a dimer of one trajectory


@pytest.mark.parametrize("case,ntr,over", [
    ("mt120_disassembly", 5, dict(steps=850, stride=200)),                                    # flags feed back into the forces
    ("mt120_constconc", 4, dict(steps=850, stride=200, conditions={"conc": 60})),             # insertions at every stride
    ("mt40_ensemble", 3, dict(steps=700, stride=200)),                                        # nothing feeds back: counts not awaited
])
def test_device_events_loop_equals_host_events_loop(case, ntr, over, rundir, monkeypatch):
    """compute() with the stride events on the device side (default) == the reference's serial stride block with mt_length()
    and change_conc() uploads on the host (MADDY_HOST_EVENTS=1 / MADDY_NO_OVERLAP=1): final state, flags, frames, mt_len.dat."""
    import mt_b200
    from mt_b200 import HostSystem, workspace
    out = {}
    over = dict(over)
    conditions = over.pop("conditions", None)
    for mode in ("device", "host"):
        d = rundir(case, runnum=ntr, **over) if conditions is None else _rundir_with(rundir, case, ntr, over, conditions)
        if mode == "host":
            monkeypatch.setenv("MADDY_HOST_EVENTS", "1")
            monkeypatch.setenv("MADDY_NO_OVERLAP", "1")
        with workspace.chdir(d):
            s = HostSystem("config.conf", [], write_files=True)
            s.srand(s.par.rseed)
            reserve0 = int(np.array(s.extra).sum())
            st = s.compute()
            out[mode] = dict(reserve0=reserve0, coords=np.array(s.coords).copy(), energies=np.array(s.energies).copy(), gtp=np.array(s.gtp).copy(),
                             on=np.array(s.on_tubule_cur).copy(), prev=np.array(s.on_tubule_prev).copy(), extra=np.array(s.extra).copy(),
                             dcd=[mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd") for t in range(ntr)],
                             ang=[mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd_ang") for t in range(ntr)],
                             mt_len=(d / "mt_len.dat").read_text(), h2d=st["h2d_bytes"])
            s.close()
    monkeypatch.delenv("MADDY_HOST_EVENTS")
    monkeypatch.delenv("MADDY_NO_OVERLAP")
    a, b = out["device"], out["host"]
    for key in ("coords", "energies", "gtp", "on", "prev", "extra"):
        assert np.array_equal(a[key], b[key]), key
    frames = over["steps"] // over["stride"] + 1
    assert all(np.array_equal(x, y) and x.shape[0] == frames for x, y in zip(a["dcd"], b["dcd"]))
    assert all(np.array_equal(x, y) for x, y in zip(a["ang"], b["ang"]))
    assert a["mt_len"] == b["mt_len"] and len(a["mt_len"].splitlines()) == frames
    if case == "mt120_constconc":
        assert int(a["extra"].sum()) < a["reserve0"]  # some reserve dimers were inserted
        assert a["h2d"] < b["h2d"]  # sparse records instead of whole-ensemble uploads


def _rundir_with(rundir, case, ntr, over, conditions):
    from mt_b200 import workspace
    spec = workspace.BASELINE_CONFIGS[case]
    return rundir(case, structure=spec["structure"], runnum=ntr, conditions=conditions, **over)


def test_device_hydrolysis_plan_equals_host_hydrolyse(rundir, load_system):
    """maddy_hydrolysis_plan (all events of a stride on the device, rand() stream produced there by jump-ahead) == hydrolyse()
    called event by event on the host with the same generator: GTP state after every event, number of draws, next draw."""
    ntr, n_events = 37, 6
    s = load_system(rundir("mt120_disassembly", runnum=ntr))
    N = s.Ntot
    e = Engine(s)
    rng = np.random.default_rng(5)
    c = np.array(s.coords).copy()

    def classify(frac_off):
        """a classification with ~frac_off of the monomers off the tubule (radius beyond R_MT + R_THRES), device and host"""
        cc = c.copy()
        off = rng.random((ntr, N // 2)) < frac_off
        cc[..., 0] = np.where(off.repeat(2, axis=1), 100.0, cc[..., 0])
        e.upload_coords(cc)
        e.snapshot_begin(coords=False, energies=False, on_tubule=True)
        e.snapshot_end()
        s.coords[...] = cc
        s.on_tubule_prev[...] = s.on_tubule_cur
        s.mt_length(1000)

    s.on_tubule_prev[...] = s.on_tubule_cur  # as the stride block of step 0 does
    classify(0.2)
    classify(0.3)
    gtp0 = (rng.random((ntr, N // 2)) < 0.7).astype(np.int32).repeat(2, axis=1)  # some dimers start as GDP
    s.gtp[...] = gtp0
    e.upload_gtp(gtp0)
    s.srand(4242)
    w = s.rand_window()
    e.hydrolysis_plan(w, 1100, 100, n_events, keep_slots=True)
    total, first, slots = e.hydrolysis_result()
    draws = 0
    for k in range(n_events):
        before = np.array(s.gtp).copy()
        elig = int(((before[:, ::2] == 1) & (np.array(s.on_tubule_cur)[:, ::2] * np.array(s.on_tubule_prev)[:, ::2] == 1)).sum())
        assert int(first[k]) == draws
        s.hydrolyse()
        draws += elig
        assert np.array_equal(slots[k], np.array(s.gtp)), k
    assert total == draws and 0 < (slots[-1] == 0).sum() and (slots[0] != gtp0).any()
    # the host generator (which made the draws) and a copy that jumps over them agree on what comes next
    nxt = [s.rand_next() for _ in range(5)]
    s.srand(4242)
    s.rand_discard(total)
    assert [s.rand_next() for _ in range(5)] == nxt
    # the schedule is what the fused loop sees: a window over the events ends in the last slot's state
    e.run(1000, 100 * n_events + 50)
    e.snapshot_begin(coords=False, energies=False, gtp=True)
    assert np.array_equal(e.snapshot_end()["gtp"], slots[-1])
    # ... and a slot can be made current ahead of its window (stride block)
    f = Engine(s)
    f.upload_gtp(gtp0)
    f.snapshot_begin(coords=False, energies=False, on_tubule=True)
    f.snapshot_end()
    from mt_b200 import MaddyError
    g = Engine(s, traj_first=0, n_tr_local=ntr - 1)
    with pytest.raises(MaddyError):
        g.hydrolysis_plan(w, 1100, 100, 2)  # a shard: draw positions are global


def test_guarded_classification_poisons_what_is_queued_behind_it(rundir, load_system):
    """MADDY_SNAP_ONTUBULE_GUARD: an undecided classification turns the plan and the window queued behind it into no-ops
    (state, RNG streams, lists untouched) until maddy_clear_guard; a decided one changes nothing."""
    s = load_system(rundir("mt120_disassembly", runnum=3))
    e, ref = Engine(s), Engine(s)
    for eng in (e, ref):
        eng.run(0, 40)
    c = e.coords()
    bad = c.copy()
    bad[1, 7, 4] = 40.0  # several turns away: the device does not guess
    e.upload_coords(bad)
    ref.upload_coords(bad)
    e.snapshot_begin(coords=False, energies=False, apply_on_tubule=True, guard=True)
    e.hydrolysis_plan(s.rand_window(), 100, 100, 2)
    e.run(40, 60)            # no-op
    _, und = e.snapshot_tubule_lengths()
    assert und != 0
    from mt_b200 import MaddyError
    with pytest.raises(MaddyError):
        e.snapshot_end()
    total, _, _ = e.hydrolysis_result()
    assert total == 0
    assert np.array_equal(e.coords(), bad) and np.array_equal(e.rng_state(), ref.rng_state())
    e.clear_guard()
    e.schedule_gtp(0, 1, [])
    # the host's verdict, then the same window on both engines
    s.coords[...] = bad
    s.mt_length(40)
    for eng in (e, ref):
        eng.upload_on_tubule(np.array(s.on_tubule_cur))
        eng.run(40, 60)
    assert np.array_equal(e.coords(), ref.coords()) and not np.array_equal(e.coords(), bad)


@pytest.mark.parametrize("case", ["mt120_disassembly", "mt40_ensemble"])
def test_loop_takes_over_on_the_host_when_the_device_cannot_decide(case, rundir, monkeypatch):
    """compute() with a crippled device rule (MADDY_ONTUB_AMAX, test hook: |theta| >= 0.05 counts as undecided) == the
    host-events loop: the guarded windows are redone, hydrolysis returns to the host, nothing is guessed."""
    import mt_b200
    from mt_b200 import HostSystem, workspace
    out = {}
    for mode in ("crippled", "host"):
        d = rundir(case, runnum=3, steps=650, stride=200)
        if mode == "crippled":
            monkeypatch.setenv("MADDY_ONTUB_AMAX", "0.05")
        else:
            monkeypatch.delenv("MADDY_ONTUB_AMAX")
            monkeypatch.setenv("MADDY_HOST_EVENTS", "1")
            monkeypatch.setenv("MADDY_NO_OVERLAP", "1")
        with workspace.chdir(d):
            s = HostSystem("config.conf", [], write_files=True)
            s.srand(s.par.rseed)
            s.compute()
            out[mode] = (np.array(s.coords).copy(), np.array(s.gtp).copy(), np.array(s.on_tubule_cur).copy(), np.array(s.energies).copy(),
                         [mt_b200.read_dcd(d / "dcd" / f"run_{t}.dcd") for t in range(3)], (d / "mt_len.dat").read_text(), s.rand_next())
            s.close()
    monkeypatch.delenv("MADDY_HOST_EVENTS")
    monkeypatch.delenv("MADDY_NO_OVERLAP")
    a, b = out["crippled"], out["host"]
    for x, y in zip(a[:4], b[:4]):
        assert np.array_equal(x, y)
    assert all(np.array_equal(x, y) and x.shape[0] == 4 for x, y in zip(a[4], b[4])) and a[5] == b[5] and a[6] == b[6]


def test_guard_escalation_equals_step_granular_path(rundir, load_system):
    """Dimers dropped next to one another (the reference's insertion rule does that) fly apart and trip the near-list
    displacement guard step after step: the fused loop escalates them and their listed partners to their exact Verlet rows
    instead of refreshing the whole trajectory - bitwise the step-granular path (full list walk), lists included."""
    s = load_system(rundir("mt120_constconc", runnum=5, steps=400), ["hydrolysis=no"])
    N = s.Ntot
    a, b = Engine(s), Engine(s)
    extra = np.array(s.extra).reshape(s.Ntr, N)
    idx, rec = [], []
    for t in (0, 2, 3):
        free = np.flatnonzero(extra[t])[::2][:3]
        for q, (x, y, z) in zip(free, ((10.0, 12.0, 100.0), (10.5, 12.3, 100.2), (30.0, -6.0, 60.0))):
            idx.append(t * N + int(q))
            rec.append((x, y, z, z + 4.0))
    for eng in (a, b):
        eng.run(0, 20)
        eng.rebuild_lj()
        eng.rebuild_bonds()
        eng.insert_dimers(idx, rec)
        eng.list_stats(reset=True)
    a.run(20, 180, skip_first_rebuild=True)
    for step in range(20, 200):
        if step % 20 == 0 and step != 20:
            b.rebuild_lj()
            b.rebuild_bonds()
        b.force()
        b.integrate()
    assert a.list_stats()["near_refresh"] > 0  # the guard did trip
    assert np.isfinite(a.coords()).all()
    assert np.array_equal(a.coords(), b.coords()) and np.array_equal(a.rng_state(), b.rng_state())
    for kind in (capi.LIST_LJ, capi.LIST_LONGITUDINAL, capi.LIST_LATERAL):
        (ca, ea), (cb, eb) = a.download_list(kind), b.download_list(kind)
        assert np.array_equal(ca, cb) and np.array_equal(ea, eb)
