"""N > 1 host logic on CPU: two `gloo` processes shard the ensemble the way bench.py / the host do.

Checks (no GPU): contiguous trajectory blocks cover the ensemble, every shard slices the GLOBAL seed table,
the oracle run of a shard equals the same trajectories of an unsharded oracle run, and the periodic ensemble
statistic (sum of per-trajectory energies) all-reduces to the unsharded value."""
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
