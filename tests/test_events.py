"""Host events of the step loop (updater.cpp): tubule length / on-tubule flags, hydrolysis, constant concentration."""
import ctypes

import numpy as np

libc = ctypes.CDLL("libc.so.6")


def test_mt_length_classification(rundir, load_system):
    s = load_system(rundir(runnum=2))
    c = s.coords
    n = s.mt_length(1000)
    assert n.tolist() == [520, 520] and (s.on_tubule_cur == 1).all()  # intact lattice: radius 8.12, theta 0
    c[0, 10, 0] += 30.0          # radius beyond R_MT + 8 r_mon = 24.12
    c[0, 11, 4] = 1.2            # theta beyond ANG_THRES (cos(theta) <= cos(1.0)); Coord order x,y,z,fi,theta,psi
    c[1, 5, 0], c[1, 5, 1] = 0.5, 0.1  # radius below 1.0
    n = s.mt_length(2000)
    assert n.tolist() == [518, 519]
    assert s.on_tubule_cur[0, 10] == 0 and s.on_tubule_cur[0, 11] == 0 and s.on_tubule_cur[1, 5] == 0


def test_hydrolysis_uses_libc_rand_in_reference_order(rundir, load_system):
    """2 % per on-tubule GTP dimer, loop i-outer / trajectory-inner, one rand() per candidate (updater.cpp:229-257)."""
    s = load_system(rundir(runnum=3))
    s.on_tubule_cur[:] = 1
    s.on_tubule_prev[:] = 1
    s.on_tubule_cur[2, 100:104] = 0   # not on tubule now -> not a candidate (and consumes no rand())
    s.srand(1234567)
    s.hydrolyse()
    got = s.gtp.copy()
    # replay the stream
    libc.srand(1234567)
    exp = np.ones((3, 520), dtype=np.int32)
    for i in range(0, 520, 2):
        for tr in range(3):
            if tr == 2 and 100 <= i < 104:
                continue
            if libc.rand() / 2147483647.0 < 0.02:
                exp[tr, i] = exp[tr, i + 1] = 0
    assert np.array_equal(got, exp) and 0 < (got == 0).sum() < 200
    # GDP dimers that are off the tubule now and at the previous stride turn back to GTP, deterministically
    idx = np.argwhere(got == 0)[0]
    s.on_tubule_cur[idx[0], idx[1] - idx[1] % 2] = 0
    s.on_tubule_prev[idx[0], idx[1] - idx[1] % 2] = 0
    before = (s.gtp == 0).sum()
    s.srand(1)
    s.hydrolyse()
    assert s.gtp[idx[0], idx[1]] == 1


def test_change_conc_inserts_reserve_dimers(rundir, load_system):
    d = rundir("mt120_constconc", structure=("reserve", 20, 6), runnum=2)
    s = load_system(d, ["is_const_conc=yes", "conc=30", "rep_r=20", "rep_h=60", "repulsive_walls=yes"])
    assert s.extra.sum() == 2 * 78
    mt_len = np.array([260, 260], dtype=np.int32)
    s.srand(7)
    changed = s.change_conc(np.zeros(2, dtype=np.int32), mt_len)
    # V = 3.14 r^2 h ; dimers wanted: conc * V * 6e-7
    want = int(np.ceil(30 * 3.14 * 20 * 20 * 60 * 6e-7))
    assert changed == 2 * want
    freed = np.argwhere(s.extra[0] == 0)
    freed = [i for i in freed.ravel() if i >= 260]
    assert len(freed) == 2 * want
    c = s.coords
    for i in freed[::2]:
        assert s.mon_type[i] == 0 and c[0, i, 0] ** 2 + c[0, i, 1] ** 2 <= 400.0
        assert c[0, i, 2] == 60 + 0.0 + 12 and c[0, i + 1, 2] == c[0, i, 2] + 4 and c[0, i + 1, 0] == c[0, i, 0]


def test_consecutive_hydrolysis_events_continue_the_libc_stream(rundir, load_system):
    """The draws of an event are pre-drawn in bulk (HostRand::fill) and tested in parallel: two consecutive events on an
    ensemble large enough for the parallel path consume exactly the libc rand() sequence, in the reference's order."""
    ntr = 130
    s = load_system(rundir(runnum=ntr))
    s.on_tubule_cur[:] = 1
    s.on_tubule_prev[:] = 1
    rng = np.random.default_rng(0)
    off = rng.random((ntr, 260)) < 0.1   # some dimers are off the tubule: no draw for them
    s.on_tubule_cur[:] = np.where(off.repeat(2, axis=1), 0, 1)
    s.srand(42)
    s.hydrolyse()
    first = s.gtp.copy()
    s.hydrolyse()
    second = s.gtp.copy()
    libc.srand(42)
    exp = np.ones((ntr, 520), dtype=np.int32)
    snaps = []
    for event in range(2):
        for i in range(0, 520, 2):
            for tr in range(ntr):
                if exp[tr, i] != 1 or off[tr, i // 2]:
                    continue
                if libc.rand() / 2147483647.0 < 0.02:
                    exp[tr, i] = exp[tr, i + 1] = 0
        snaps.append(exp.copy())
    assert np.array_equal(first, snaps[0]) and np.array_equal(second, snaps[1])
    assert (second == 0).sum() > (first == 0).sum() > 0


def test_rand_jump_ahead_matches_libc(rundir, load_system):
    """maddy_rand_discard / HostRand::discard (polynomial jump-ahead of glibc's TYPE_3 generator, maddy_lfib.h): the window
    after n draws is the one libc reaches by drawing, for small, block-sized and large n; the documented next-draw formula."""
    import ctypes as C
    from mt_b200 import capi
    s = load_system(rundir(runnum=1))
    for seed, n in ((1, 1), (1234567, 2), (7, 30), (7, 31), (7, 32), (99, 124), (99, 248 * 32 + 5), (4242, 1_000_003)):
        s.srand(seed)
        libc.srand(seed)
        w = s.rand_window()
        assert ((int(w[0]) + int(w[28])) & 0xFFFFFFFF) >> 1 == libc.rand()  # the next draw is (w[0] + w[28]) >> 1
        libc.srand(seed)
        s.rand_discard(n)
        for _ in range(n):
            libc.rand()
        assert [s.rand_next() for _ in range(64)] == [libc.rand() for _ in range(64)], (seed, n)
        # the C-ABI entry on a raw window
        s.srand(seed)
        w = s.rand_window().copy()
        capi.lib.maddy_rand_discard(w.ctypes.data_as(C.POINTER(C.c_uint)), n)
        s.rand_discard(n)
        assert np.array_equal(w, s.rand_window())


def test_device_on_tubule_rule_equals_the_host_cosf_test(rundir, load_system):
    """The classification kernel never calls cosf: it is given interval edges of |theta| bisected from this process's cosf
    (maddy_on_tubule_rule, host-only).  The rule, restated in numpy exactly as ontubule_kernel evaluates it, must agree with
    the host's mt_length() - the reference's own predicate (updater.cpp:164-170, pinned in test_events_golden.py) - on every
    float within 300 ulps of each edge, on both signs, on random angles up to the rule's range, and at the radius limits."""
    import ctypes as C
    from mt_b200 import capi
    rad_hi, a_max = C.c_float(), C.c_float()
    edges = np.zeros(7, dtype=np.float32)
    assert capi.lib.maddy_on_tubule_rule(C.byref(rad_hi), C.byref(a_max), edges.ctypes.data_as(C.POINTER(C.c_float))) == 0
    rad_hi, a_max = np.float32(rad_hi.value), np.float32(a_max.value)
    assert rad_hi == np.float32(8.12) + np.float32(16.0) and np.all(np.diff(edges) > 0) and a_max > edges[-1]
    # cos(theta) > cos(1): the crossings are at 1, 2 pi - 1, 2 pi + 1, 4 pi - 1, ...
    exact = np.array([1, 2 * np.pi - 1, 2 * np.pi + 1, 4 * np.pi - 1, 4 * np.pi + 1, 6 * np.pi - 1, 6 * np.pi + 1])
    assert np.allclose(edges, exact, rtol=1e-6)

    rng = np.random.default_rng(11)
    near = np.concatenate([(e.view(np.int32) + np.arange(-300, 301, dtype=np.int32)).view(np.float32) for e in edges.reshape(-1, 1)])
    theta = np.concatenate([near, -near, rng.uniform(-float(a_max), float(a_max), 4000).astype(np.float32),
                            np.array([0.0, -0.0, np.nextafter(a_max, np.float32(0))], dtype=np.float32)])
    x = np.full(theta.shape, 8.12, dtype=np.float32)
    y = np.zeros_like(x)
    # radius limits: rad < rad_hi and rad > 1.0, a few ulps either side (theta = 0 there)
    for lim in (np.float32(1.0), rad_hi):
        r = (lim.view(np.int32) + np.arange(-40, 41, dtype=np.int32)).view(np.float32)
        ang = rng.uniform(0, 2 * np.pi, r.size)
        x = np.concatenate([x, (r * np.cos(ang)).astype(np.float32)])
        y = np.concatenate([y, (r * np.sin(ang)).astype(np.float32)])
        theta = np.concatenate([theta, np.zeros(r.size, dtype=np.float32)])

    def rule(x, y, theta):  # ontubule_kernel (mt_b200/csrc/maddy_analysis.cu), float arithmetic in its order
        rad = np.sqrt(x * x + y * y, dtype=np.float32)
        a = np.abs(theta)
        assert (a < a_max).all()
        inside = a < edges[0]
        for k in (1, 3, 5):
            inside |= (a > edges[k]) & (a < edges[k + 1])
        return (rad < rad_hi) & (rad.astype(np.float64) > 1.0) & inside

    s = load_system(rundir(runnum=1))
    N = s.Ntot
    c = s.coords
    want = rule(x, y, theta)
    assert 0.2 < want.mean() < 0.8
    checked = 0
    for first in range(0, theta.size, N):
        n = min(N, theta.size - first)
        c[0, :, 0], c[0, :, 1], c[0, :, 4] = 8.12, 0.0, 0.0
        c[0, :n, 0], c[0, :n, 1], c[0, :n, 4] = x[first:first + n], y[first:first + n], theta[first:first + n]
        s.mt_length(1000)
        got = s.on_tubule_cur[0, :n] == 1
        bad = np.flatnonzero(got != want[first:first + n])
        assert bad.size == 0, (first + bad[:5], theta[first + bad[:5]], x[first + bad[:5]], y[first + bad[:5]])
        checked += n
    assert checked == theta.size > 12000
