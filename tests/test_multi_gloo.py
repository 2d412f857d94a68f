"""N > 1 host logic on CPU: two `gloo` processes shard the ensemble the way bench.py / the host do.

Checks (no GPU): contiguous trajectory blocks cover the ensemble, every shard slices the GLOBAL seed table,
the oracle run of a shard equals the same trajectories of an unsharded oracle run, and the periodic ensemble
statistic (sum of per-trajectory energies) all-reduces to the unsharded value; and the one exchange step of the path,
the per-stride all-gather of the hydrolysis plan inputs, after which every rank evaluates the ensemble-wide draw order on
its own (oracle/hyd_plan.py restates what mt_b200/csrc/maddy_events.cu does on the GPUs)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, rundir, ntr, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mt_b200 import HostSystem, workspace
    from oracle.pyoracle import OracleState
    with workspace.chdir(rundir):
        s = HostSystem("config.conf", ["hydrolysis=no"])
    first = ntr * rank // world
    count = ntr * (rank + 1) // world - first
    o = OracleState(s, traj_first=first, n_tr_local=count)
    o.run(0, 25)
    e = torch.tensor(o.energies().sum(axis=(0, 1)), dtype=torch.float64)
    dist.all_reduce(e)
    cover = torch.zeros(ntr, dtype=torch.int64)
    cover[first:first + count] = 1
    dist.all_reduce(cover)
    np.savez(Path(out) / f"rank{rank}.npz", coords=o.coords, rng=o.rng, first=first, count=count, esum=e.numpy(), cover=cover.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded(rundir, tmp_path, load_system):
    ntr = 3
    d = rundir(runnum=ntr)
    mp.spawn(_worker, args=(2, _free_port(), str(d), ntr, str(tmp_path)), nprocs=2, join=True)
    from oracle.pyoracle import OracleState
    s = load_system(d, ["hydrolysis=no"])
    full = OracleState(s)
    full.run(0, 25)
    esum = full.energies().sum(axis=(0, 1))
    N = s.Ntot
    for rank in range(2):
        r = np.load(tmp_path / f"rank{rank}.npz")
        first, count = int(r["first"]), int(r["count"])
        assert (r["cover"] == 1).all()
        assert np.array_equal(r["coords"], full.coords[first:first + count])  # bit-identical to the unsharded run
        assert np.array_equal(r["rng"][0], full.rng[0, first * N:(first + count) * N])
        assert np.array_equal(r["rng"][1], full.rng[1, first * N:(first + count) * N])
        assert np.allclose(r["esum"], esum, rtol=1e-12)


def _flags(ntr, N):
    """seeded on-tubule history and GTP state of the whole ensemble (every process derives the same)"""
    rng = np.random.default_rng(5)
    cur = (rng.random((ntr, N // 2)) < 0.85).repeat(2, axis=1).astype(np.int32)
    prev = (rng.random((ntr, N // 2)) < 0.85).repeat(2, axis=1).astype(np.int32)
    gtp = (rng.random((ntr, N // 2)) < 0.9).repeat(2, axis=1).astype(np.int32)
    return cur, prev, gtp


def _plan_worker(rank, world, port, rundir, ntr, n_events, seed, out):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), OMP_NUM_THREADS="1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mt_b200 import HostSystem, workspace
    from oracle import hyd_plan
    with workspace.chdir(rundir):
        s = HostSystem("config.conf", [])
    N = s.Ntot
    cur, prev, gtp = _flags(ntr, N)
    per = ntr // world
    sl = slice(rank * per, (rank + 1) * per)
    extra = np.asarray(s.extra).reshape(ntr, N)
    gt, st = hyd_plan.shard_cells(gtp[sl], extra[sl], cur[sl], prev[sl])
    own = torch.from_numpy(np.concatenate([gt.ravel(), st.ravel()]))
    parts = [torch.empty_like(own) for _ in range(world)]
    dist.all_gather(parts, own)  # the path's only exchange: once per stride
    cells = gt.size
    gts = [p.numpy()[:cells].reshape(gt.shape) for p in parts]
    sts = [p.numpy()[cells:].reshape(st.shape) for p in parts]
    s.srand(seed)
    slots, used = hyd_plan.plan(gts, sts, rank, s.rand_window(), n_events)
    s.rand_discard(used)
    np.savez(Path(out) / f"plan{rank}.npz", slots=np.stack(slots), used=used, nxt=np.array([s.rand_next() for _ in range(8)]))
    dist.destroy_process_group()


def test_two_rank_hydrolysis_plan_matches_the_host_events(rundir, tmp_path, load_system):
    """Each rank evaluates the plan of a whole stride from the all-gathered cells and keeps its own trajectories; together
    they are the GTP states the host's hydrolyse() (= the reference's, tests/test_events_golden.py) leaves after each event,
    and every rank's generator ends where the host's does."""
    ntr, n_events, seed = 6, 3, 20260117
    d = rundir(runnum=ntr)
    mp.spawn(_plan_worker, args=(2, _free_port(), str(d), ntr, n_events, seed, str(tmp_path)), nprocs=2, join=True)
    s = load_system(d)
    N = s.Ntot
    cur, prev, gtp = _flags(ntr, N)
    s.on_tubule_cur[:], s.on_tubule_prev[:], s.gtp[:] = cur, prev, gtp
    s.srand(seed)
    after = []
    for _ in range(n_events):
        s.hydrolyse()
        after.append(s.gtp.copy())
    nxt = [s.rand_next() for _ in range(8)]
    changed = 0
    for rank in range(2):
        r = np.load(tmp_path / f"plan{rank}.npz")
        sl = slice(rank * 3, rank * 3 + 3)
        for k in range(n_events):
            assert np.array_equal(r["slots"][k], after[k][sl, 0::2]) and np.array_equal(after[k][sl, 0::2], after[k][sl, 1::2])
        changed += int((r["slots"][-1] != gtp[sl, 0::2]).sum())
        assert r["nxt"].tolist() == nxt  # same rand() position as the host after the stride's events
    assert changed > 10  # hydrolysis and returns to GTP both happened
