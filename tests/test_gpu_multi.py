"""Multi-GPU paths (skipped on a single-GPU box): the C++ host sharding trajectories over n_gpus handles and the
NCCL ensemble reduction of the C-ABI."""
import ctypes as C

import numpy as np
import pytest

from mt_b200 import Engine, capi

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


def test_host_compute_sharded_equals_single_gpu(rundir, load_system):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    d = rundir("mt40_single", runnum=5, steps=250, stride=100)
    a = load_system(d)
    b = load_system(d)
    a.srand(1234567)
    a.compute(n_gpus=1)
    b.srand(1234567)
    b.compute(n_gpus=2)
    assert np.array_equal(a.coords, b.coords) and np.array_equal(a.gtp, b.gtp)
    assert np.array_equal(a.energies, b.energies)


def test_ensemble_allreduce_nccl(rundir, load_system):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    s = load_system(rundir("mt40_single", runnum=4), ["hydrolysis=no"])
    e0 = Engine(s, traj_first=0, n_tr_local=2, device=0)
    e1 = Engine(s, traj_first=2, n_tr_local=2, device=1)
    v0 = e0.energies().sum(axis=0)
    v1 = e1.energies().sum(axis=0)
    want = v0 + v1
    hs = (C.c_void_p * 2)(e0._h, e1._h)
    bufs = (C.POINTER(C.c_double) * 2)(capi.as_ptr(v0, C.c_double), capi.as_ptr(v1, C.c_double))
    rc = capi.lib.maddy_ensemble_allreduce(hs, 2, bufs, 7)
    assert rc == 0, capi.lib.maddy_last_error(e0._h)
    assert np.allclose(v0, want, rtol=1e-14) and np.allclose(v1, want, rtol=1e-14)


def test_ensemble_stats_two_gpus_equal_host_statistics(rundir, load_system):
    """maddy_ensemble_stats in the stride block of the one-host loop (NCCL all-reduce over the shards) against numpy
    statistics of the per-trajectory energies the same run returns."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    d = rundir("mt40_single", runnum=6, steps=201, stride=100)
    s = load_system(d)
    s.srand(1234567)
    s.compute(n_gpus=2)
    st, e = s.ensemble_stats, s.energies
    assert st is not None and st[14] == 6
    assert np.allclose(st[:7], e.sum(axis=0), rtol=1e-13, atol=1e-9)
    assert np.allclose(st[7:14], (e * e).sum(axis=0), rtol=1e-13, atol=1e-9)


def test_sharded_device_hydrolysis_equals_single_gpu(rundir, load_system):
    """Equal shards: every GPU evaluates the ensemble's hydrolysis plan from the all-gathered inputs (maddy_hydrolysis_plan_all)
    - same draws in the same global order as the single-GPU plan: state, GTP flags, energies and the host generator agree."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    d = rundir("mt40_ensemble", runnum=6, steps=450, stride=200)
    out = []
    for g in (1, 2):
        s = load_system(d)
        s.srand(1234567)
        s.compute(n_gpus=g)
        out.append((np.array(s.coords).copy(), np.array(s.gtp).copy(), np.array(s.energies).copy(), np.array(s.on_tubule_cur).copy(), s.rand_next()))
    a, b = out
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3]) and a[4] == b[4]
    assert (a[1] == 0).any()  # some dimers were hydrolysed
