"""CPU restatement (numpy) of the hydrolysis PLAN of one stride — TEST INFRASTRUCTURE ONLY: imported by tests/, never
by the product path (the plan runs on the GPU, mt_b200/csrc/maddy_events.cu; the host's own events are
mt_b200/host/events.cpp).

What it restates: the reference's hydrolyse() (updater.cpp:229-257) applied `n_events` times with the on-tubule flags
of the last stride block held fixed, written the way the device evaluates it — from per-shard INPUT CELLS rather than
from a loop over rand():

  * a shard (a contiguous block of trajectories) reduces its state to two byte tables, rows = dimers, columns = its
    trajectories:  gt[d][tr] = GTP state of the dimer's first monomer;  st[d][tr] bit 0 = can hydrolyse (not a reserve
    particle, on the tubule now AND at the previous stride), bit 1 = returns to GTP (not reserve, off the tubule both
    times).  These are the bytes maddy_hydrolysis_inputs() exposes and the ranks all-gather once per stride;
  * the reference draws one rand() per ELIGIBLE dimer in the order dimer-outer / trajectory-inner over the WHOLE
    ensemble (updater.cpp:233-236), so the position of a cell's draw in the stream is
        draws consumed by earlier events + eligible cells in earlier rows + eligible cells left of it in its row,
    a pure function of the gathered tables: every shard can evaluate it without talking to the others again;
  * the draw itself: glibc TYPE_3 rand(), x[n] = x[n-31] + x[n-3] (mod 2^32), output x >> 1, from the 31-word window of
    the host generator (mt_b200/csrc/maddy_lfib.h states the same recurrence); hydrolysed when
    (int)draw / (double)RAND_MAX < 0.02 (updater.cpp:235-236);
  * then GDP cells with bit 1 return to GTP, without a draw (updater.cpp:246-254).

Pinned by tests/test_multi_gloo.py against mt_b200's host hydrolyse() (itself pinned bit for bit against the
reference's updater.cpp through oracle/ref_events_probe.cu, tests/test_events_golden.py) and libc rand().
"""
import numpy as np

RAND_MAX = 2147483647


def shard_cells(gtp, extra, cur, prev):
    """[ntr_l, N] arrays of one shard -> (gt, st), each uint8 [N/2, ntr_l]"""
    g = np.asarray(gtp)[:, 0::2]
    ex = np.asarray(extra)[:, 0::2] != 0
    c = np.asarray(cur)[:, 0::2] != 0
    p = np.asarray(prev)[:, 0::2] != 0
    st = (~ex & c & p).astype(np.uint8) | ((~ex & ~c & ~p).astype(np.uint8) << 1)
    return np.ascontiguousarray(g.T.astype(np.uint8)), np.ascontiguousarray(st.T)


def rand_stream(window31, n):
    """the next n rand() values of the generator whose last 31 state words are window31 (oldest first)"""
    x = [int(v) for v in window31]
    out = np.empty(n, dtype=np.int64)
    for i in range(n):
        v = (x[i] + x[i + 28]) & 0xFFFFFFFF
        x.append(v)
        out[i] = v >> 1
    return out


def plan(gt_shards, st_shards, shard, window31, n_events):
    """gt_shards / st_shards: the gathered tables of EVERY shard, in rank order.  Returns (slots, draws_used):
    slots[k] = GTP state [ntr_l, N/2] of shard `shard`'s dimers after event k; draws_used = rand() calls of the whole
    plan (the same on every shard: the host generator is advanced by it)."""
    gt = np.concatenate(gt_shards, axis=1).astype(np.uint8)  # [nd, ntr] over the whole ensemble
    st = np.concatenate(st_shards, axis=1)
    first = sum(g.shape[1] for g in gt_shards[:shard])
    width = gt_shards[shard].shape[1]
    can, back = (st & 1) != 0, (st & 2) != 0
    # an upper bound of the draws: every cell in every event
    stream = rand_stream(window31, int(can.sum()) * n_events)
    used = 0
    slots = []
    for _ in range(n_events):
        elig = (gt == 1) & can
        pos = used + np.cumsum(elig.ravel()).reshape(elig.shape) - 1  # row-major = dimer-outer / trajectory-inner
        hit = np.zeros_like(elig)
        hit[elig] = stream[pos[elig]] / float(RAND_MAX) < 0.02
        used += int(elig.sum())
        gt = np.where(hit, 0, gt).astype(np.uint8)
        gt = np.where((gt == 0) & back, 1, gt).astype(np.uint8)
        slots.append(gt[:, first:first + width].T.copy())
    return slots, used
