"""Ensemble sizes beyond what the other event tests use (this file sorts last on purpose: the cases are bigger).

The device hydrolysis plan cuts a dimer row into segments of 256 trajectories and reaches the far end of the rand() stream
through a third level of jump matrices (draws beyond 1024 x 248); neither is touched by ensembles of a few dozen
trajectories.  Checker: the host's hydrolyse() (pinned to the reference's updater.cpp:229-257 in test_events_golden.py and
to libc rand() in test_events.py) called event by event with the same generator.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from mt_b200 import Engine  # noqa: E402


def test_device_hydrolysis_plan_on_300_trajectories(rundir, load_system):
    ntr, n_events = 300, 6   # two row segments (256 + 44); > 254 k draws in the plan
    s = load_system(rundir("mt40_ensemble", runnum=ntr))
    N = s.Ntot
    e = Engine(s)
    rng = np.random.default_rng(17)
    c = np.array(s.coords).copy()

    def classify(frac_off):
        """a classification with ~frac_off of the dimers off the tubule (radius beyond R_MT + R_THRES), device and host"""
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
    classify(0.05)
    classify(0.05)
    gtp0 = (rng.random((ntr, N // 2)) < 0.95).astype(np.int32).repeat(2, axis=1)  # a few dimers start as GDP
    s.gtp[...] = gtp0
    e.upload_gtp(gtp0)
    s.srand(987654)
    e.hydrolysis_plan(s.rand_window(), 1100, 100, n_events, keep_slots=True)
    total, first, slots = e.hydrolysis_result()
    draws = 0
    for k in range(n_events):
        before = np.array(s.gtp).copy()
        elig = int(((before[:, ::2] == 1) & (np.array(s.on_tubule_cur)[:, ::2] * np.array(s.on_tubule_prev)[:, ::2] == 1)).sum())
        assert int(first[k]) == draws, k
        s.hydrolyse()
        draws += elig
        assert np.array_equal(slots[k], np.array(s.gtp)), k
    assert total == draws and total > 1024 * 248  # the third level of the stream's jump matrices was needed
    assert (slots[-1] == 0).sum() > (gtp0 == 0).sum() and (slots[0] != gtp0).any()
    # the host generator (which made the draws) and a copy that jumps over them agree on what comes next
    nxt = [s.rand_next() for _ in range(5)]
    s.srand(987654)
    s.rand_discard(total)
    assert [s.rand_next() for _ in range(5)] == nxt
    # the schedule is what the fused loop sees: a window over the events ends in the last slot's state
    e.run(1000, 100 * n_events + 50)
    e.snapshot_begin(coords=False, energies=False, gtp=True)
    assert np.array_equal(e.snapshot_end()["gtp"], slots[-1])
