#!/usr/bin/env python
"""bench.py — monomer-steps/s of the MADDY Langevin/BD step loop on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--ntr T]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (N > 1)

One bench "step" = ONE OUTPUT STRIDE of the workload (`stride` = 1000 MD steps in every BASELINE config; stated as
config.md_steps_per_bench_step): the stride block of the reference loop (list rebuild, energies, coordinate read-back;
compute_cuda.cu:1163-1226) followed by 1000 MD steps of all Ntot x Ntr monomers (list rebuild every LJPairsUpdateFreq =
20 steps, GTP-flag update every hydrostep = 100 steps, force evaluation, Langevin integration).  `--steps 20 --warmup 5`
therefore times 20 000 MD steps after 5 000 warm-up steps, in BOTH arms.  Workload at every N: BASELINE.json configs[1],
the 13-protofilament seed (520 monomers) x 256 trajectories PER GPU (weak scaling; the global ensemble of 256 N
trajectories is sharded in contiguous blocks, every shard uses the GLOBAL RNG stream ids).

value   device-timed whole job, state resident in HBM: K strides issued exactly as the drop-in loop issues them
        (snapshot_begin[rebuild + energies + on-tubule classification + coordinates -> pinned host], the hydrolysis events of
        the stride planned on the device, ONE fused maddy_run window per stride, snapshot_end), ONE CUDA event pair on the launching stream around the K
        strides, a 256 MiB L2 flush write between bench steps (inside the pair), max over ranks; repeated R times, median.
e2e     the drop-in compute() call of the C++ host (mt_system_compute) over K strides with HOST buffers: device
        allocation + upload of coordinates / topology / seeds, hydrolysis draws + uploads, energies + coordinate download
        every stride, DCD / PDB / mt_len output, all inside the timed region (wall clock); median of 9 runs after 2 untimed ones.
roofline  algorithmic bytes of the step-granular contract (352 B per monomer-step on the intact lattice)
        per fused-window launch / average launch duration (event pairs around every maddy_run inside the timed region),
        against the measured HBM copy bandwidth.  The fused kernel keeps the state on-chip, so `traffic` (ncu dram
        bytes) is far BELOW the algorithmic bytes.
cpu_baseline  the CPU oracle port (oracle/maddy_oracle.c, OpenMP) on a bounded sample, rank 0, N = 1.
--impl reference  the reference has NO CPU path: its own CUDA build (oracle/_ref/mt, sm_100 recompile, unmodified
        sources) runs the same conf files on the same GPU(s).  Its `zs[100]` array limits a process to 100 trajectories
        (SURVEY.md 8d), so the 256 trajectories per GPU run as three processes of 86/85/85, one after the other and
        all at once; the faster way is reported.  Timing: the parent time-stamps the "Saving coordinates at step S"
        line the reference prints in every stride block; the timed region is from the line of stride W to the line of
        stride W + K (exactly K strides, process start-up and initialisation excluded, nothing differenced).
N > 1   every rank runs the two legs above on its shard; in addition rank 0 reports `shard_parity` (each rank's shard
        bitwise equal to the same trajectories of a single-GPU run of the global ensemble, first 40 steps) and
        `e2e_one_host`: the product's own multi-GPU path, ONE host thread driving N handles (mt_system_compute with
        n_gpus = N, runnum = 256 N; global rand() order; NCCL ensemble statistics in the stride block).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "monomer-steps/s (ensemble, device-timed)"
UNIT = "monomer-steps/s"
# algorithmic bytes per monomer-step of the step-granular contract (BASELINE.md 3 / SURVEY.md 8d): intact lattice
# (46 LJ partners, 4 bonded) and free dimers in a cylinder (13 LJ partners)
B_ALG = {"lattice": 352.0, "free": 220.0}
B_ALG_TERMS = "64 state r/w + 64 RNG r/w + 4*(n_LJ+1) LJ list + 4*(n_bonded+3) bond lists + 8 flags"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (a 100-step fused window at 520 x 256) from the
# `ncu --set full` capture summarised in profiles/r2_run_kernel_ncu_full.txt (32.8 MB read + 0.2 MB written)
NCU_TRAFFIC_PER_LAUNCH = {("mt40_ensemble", 256): (33.0e6, 100)}
REF_NTR_LIMIT = 100  # Parameters::zs[100], parameters.h:12,293


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return json.loads(p.read_text()), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.1)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        sm = sorted(int(s[0]) for s in self.samples if s and s[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples if len(s) >= 6 for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        mx = max((int(s[1]) for s in self.samples if len(s) > 1 and s[1].isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def scratch_dir(prefix: str) -> Path:
    """Run directories of BOTH arms (config files in, DCD / PDB trajectories out).  At 8 GPUs the ensemble writes
    ~5 GB/s of trajectory frames; on a box whose /tmp sits on a virtual disk that rate is throttled by dirty-page
    write-back (measured: 5 ms per stride instead of 0.5 ms), which times the disk and not the path.  So the output
    goes to tmpfs when there is one with room; MADDY_BENCH_SCRATCH overrides."""
    root = os.environ.get("MADDY_BENCH_SCRATCH")
    if not root:
        shm = Path("/dev/shm")
        try:
            if shm.is_dir() and os.access(shm, os.W_OK) and shutil.disk_usage(shm).free > (8 << 30):
                root = str(shm)
        except OSError:
            root = None
    return Path(tempfile.mkdtemp(prefix=prefix, dir=root))


def make_system(workload: str, ntr_global: int, tmp: Path, overrides=(), write_files=False, **config):
    from mt_b200 import HostSystem, workspace
    workspace.make_baseline_rundir(tmp, workload, runnum=ntr_global, **config)
    with workspace.chdir(tmp):
        return HostSystem("config.conf", list(overrides), write_files=write_files)


def workload_facts(workload: str):
    """(Ntot, stride, hydrostep or 0, B_alg key) of a workload, from the run directory the host itself parses (no GPU)."""
    from mt_b200 import workspace
    tmp = Path(tempfile.mkdtemp(prefix="bench_facts_"))
    try:
        s = make_system(workload, 1, tmp)
        kind = workspace.BASELINE_CONFIGS[workload]["structure"][0]
        hyd = int(s.host.hydrostep) if s.host.hydrolysis and s.host.hydrostep > 0 else 0
        out = (int(s.Ntot), int(s.host.stride), hyd, "free" if kind == "free" else "lattice", bool(s.par.tea_on))
        s.close()
        return out
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def workload_text(workload: str, N: int, ntr_local: int, ntr_global: int) -> str:
    return f"{workload}: {N} monomers x {ntr_local} trajectories per GPU ({ntr_global} total), BASELINE.json configs[1] when mt40_ensemble " \
           f"(13-PF MT seed, Morse + LJ, dynamic bond lists, hydrolysis, dt 200)"


def common_config(workload: str, N: int, ntr_local: int, world: int, stride: int) -> dict:
    """identical in both arms, so the driver can compare them key by key"""
    return {"workload": workload_text(workload, N, ntr_local, ntr_local * world), "ntot": N, "ntr_per_gpu": ntr_local,
            "ntr_total": ntr_local * world, "md_steps_per_bench_step": stride,
            "bench_step": "one output stride: stride block (rebuild + energies + coordinate read-back) + `stride` MD steps",
            "l2": "per-GPU state + neighbour lists (~190 MB at 520 x 256) exceed the 126 MB L2; this repo's arm additionally "
                  "writes a 256 MiB flush buffer between bench steps, inside the timed region"}


def cpu_baseline(workload: str, budget_s: float = 10.0):
    """CPU oracle (port) timed on a bounded sample of the same workload: one trajectory per core."""
    from oracle.pyoracle import OracleState
    cores = os.cpu_count() or 1
    ntr = max(1, min(cores, 16))
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    tmp = Path(tempfile.mkdtemp(prefix="bench_cpu_"))
    try:
        s = make_system(workload, ntr, tmp)
        o = OracleState(s)
        o.run(0, 20)  # warm-up incl. first rebuild
        t0 = time.perf_counter()
        o.run(20, 40)
        per_step = (time.perf_counter() - t0) / 40
        steps = int(max(40, min(20000, budget_s / max(per_step, 1e-6))))
        steps -= steps % 20
        t0 = time.perf_counter()
        o.run(60, steps)
        dt = time.perf_counter() - t0
        return {"value": s.Ntot * ntr * steps / dt, "unit": UNIT, "cores": min(cores, ntr), "kind": "port",
                "sample": f"oracle/maddy_oracle.c (OpenMP over trajectories), {ntr} trajectories x {s.Ntot} monomers x {steps} MD steps of {workload} "
                          f"({dt:.1f} s; the reference itself has no CPU path)"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


class _DevicePtr:
    """a raw device buffer of the C-ABI as something torch can wrap (no copy)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


class StrideIssuer:
    """Issues strides on one Engine the way mt_b200/host/events.cpp::compute does in its overlapped mode: at a stride step
    the read-back (list rebuild + energies + coordinates) is queued, the first window launched, the snapshot collected;
    inside the stride the windows grow by one hydrolysis period (100, 100, 200, 300, 300 steps at the template's periods),
    each preceded by the GTP upload / schedule of its hydrolysis events (here: the resident flags, re-sent)."""

    def __init__(self, eng, system, torch_stream, single_event=False):
        import numpy as np
        self.eng, self.stream = eng, torch_stream
        self.stride = int(system.host.stride)
        hyd = int(system.host.hydrostep) if system.host.hydrolysis and system.host.hydrostep > 0 else 0
        self.period = hyd if 0 < hyd < self.stride else 0
        self.tea = bool(system.par.tea_on)
        first = eng.par.traj_first
        self.gtp = np.ascontiguousarray(system.gtp[first:first + eng.ntr], dtype=np.int32)
        # hydrolysis on the device (the drop-in loop's default when one handle holds the ensemble): the events of a stride
        # are planned right after its stride block and applied inside ONE window that spans the stride
        self.system = system
        self.shards = system.Ntr // eng.ntr if system.Ntr % eng.ntr == 0 else 0  # equal contiguous blocks, one per rank
        self.dev_hyd = bool(self.period) and not self.tea and system.Ntot % 2 == 0 and bool(system.host.tub_length) \
            and self.shards >= 1 and not os.environ.get("MADDY_HOST_HYDROLYSIS") and not os.environ.get("MADDY_HOST_EVENTS")
        self.gathered = None  # sharded ensemble: every rank's plan inputs, all-gathered once per stride (NCCL over NVLink)
        self.apply_flags = bool(system.par.barrier)
        self.plan_pending = False
        self.windows = [(0, self.period, 0), (self.period, self.stride - self.period, self.stride // self.period)] if self.dev_hyd \
            else self._pattern(single_event)
        self.window_events = []  # (start, end) torch events around every maddy_run since the last reset
        self.md_in_windows = 0

    def _pattern(self, single_event):
        if not self.period:
            return [(0, self.stride, 0)]
        out, s, last = [], 0, 0
        h = self.period
        while s < self.stride:
            if s == 0 or single_event:
                n = h
            elif last == h and s == h:   # the window after the post-stride one stays short
                n = h
            else:
                n = min((last // h + 1) * h, self.stride - s)
            n = min(n, self.stride - s)
            out.append((s, n, 0 if s == 0 else (n + h - 1) // h))  # events at s, s+h, ... inside the window
            s += n
            last = n
        return out

    def issue(self, s0: int, timed: bool):
        import numpy as np
        import torch
        eng = self.eng
        if self.dev_hyd:
            # exactly the calls of mt::compute() in its device-events mode (mt_b200/host/events.cpp)
            h = self.period
            if s0 != 0 and s0 % h == 0:
                eng.apply_scheduled_gtp(s0)  # the event AT a stride step precedes that stride's energies
            if self.plan_pending:
                total, _, _ = eng.hydrolysis_result()
                self.system.rand_discard(total)
                self.plan_pending = False
            eng.snapshot_begin(coords=True, energies=True, rebuild=True, on_tubule=True, apply_on_tubule=(s0 != 0 and self.apply_flags), gtp=True,
                               guard=True)
            if self.shards == 1:
                eng.hydrolysis_plan(self.system.rand_window(), s0 + h, h, self.stride // h)
            else:
                # the draw positions are global: each rank evaluates the ensemble's plan from the gathered inputs
                import torch.distributed as dist
                ptr, nb = eng.hydrolysis_inputs()
                own = torch.as_tensor(_DevicePtr(ptr, nb), device="cuda")
                if self.gathered is None:
                    self.gathered = torch.empty(self.shards * nb, dtype=torch.uint8, device="cuda")
                dist.all_gather_into_tensor(self.gathered, own)
                eng.hydrolysis_plan_sharded(self.gathered.data_ptr(), self.shards, self.system.rand_window(), s0 + h, h, self.stride // h)
            self.plan_pending = True
            # the plan is evaluated beside the window up to its first event; the window to the next stride step waits for it
            for first, n in ((s0, h), (s0 + h, self.stride - h)):
                if timed:
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(self.stream)
                eng.run(first, n, skip_first_rebuild=(first == s0))
                if timed:
                    b.record(self.stream)
                    self.window_events.append((a, b))
                    self.md_in_windows += n
            eng.snapshot_end()  # the host collects while the windows run
            if eng.snapshot_tubule_lengths()[1]:
                raise RuntimeError("bench: the on-tubule classification was undecided (the drop-in loop would redo the stride on the host)")
            return
        for off, n, n_ev in self.windows:
            if off == 0:
                if self.period and s0 != 0:
                    eng.upload_gtp(self.gtp)  # the event AT a stride step precedes that stride's energies
                eng.snapshot_begin(coords=True, energies=True, rebuild=True)
            elif n_ev == 1:
                eng.upload_gtp(self.gtp)
            elif n_ev > 1:
                eng.schedule_gtp(s0 + off, self.period, np.broadcast_to(self.gtp, (n_ev,) + self.gtp.shape))
            if timed:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(self.stream)
            eng.run(s0 + off, n, skip_first_rebuild=(off == 0))
            if timed:
                b.record(self.stream)
                self.window_events.append((a, b))
                self.md_in_windows += n
            if off == 0:
                eng.snapshot_end()  # the host collects while the first window runs


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mt_b200 import Engine, workspace

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path")
    torch.cuda.set_device(local)
    cpu_group = None
    if world > 1:
        import datetime
        # a collective that a rank never joins fails after 5 minutes instead of holding the box for the default 10
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=300))
        cpu_group = dist.new_group(backend="gloo")  # host-side waits that must not put a spinning kernel on a GPU
    pk, pk_src = peaks()
    ntr_local = args.ntr
    ntr_global = ntr_local * world
    K, warm = args.steps, max(args.warmup, 3)
    tmp = scratch_dir(f"bench_r{rank}_")
    try:
        system = make_system(args.workload, ntr_global, tmp)
        N = system.Ntot
        stride = int(system.host.stride)
        kind = "free" if workspace.BASELINE_CONFIGS[args.workload]["structure"][0] == "free" else "lattice"
        b_alg = B_ALG[kind]
        stream = torch.cuda.Stream()

        # ---- shard parity (N > 1): this rank's shard against the same trajectories of the global ensemble on ONE GPU
        shard_parity = None
        if world > 1:
            a = Engine(system, traj_first=rank * ntr_local, n_tr_local=ntr_local, device=local)
            b = Engine(system, traj_first=0, n_tr_local=ntr_global, device=local)
            a.run(0, 40)
            b.run(0, 40)
            sl = slice(rank * ntr_local, (rank + 1) * ntr_local)
            n0, n1 = rank * ntr_local * N, (rank + 1) * ntr_local * N
            ok = np.array_equal(a.coords(), b.coords()[sl]) and np.array_equal(a.rng_state(), b.rng_state()[:, n0:n1]) \
                and np.array_equal(a.energies(), b.energies()[sl])
            a.close()
            b.close()
            t = torch.tensor([1 if ok else 0], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            shard_parity = bool(t.item())

        eng = Engine(system, traj_first=rank * ntr_local, n_tr_local=ntr_local, device=local, stream=stream.cuda_stream)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
        issuer = StrideIssuer(eng, system, stream, single_event=bool(os.environ.get("MADDY_SINGLE_EVENT_WINDOWS")))

        def region(first_stride: int, count: int, timed: bool):
            """`count` bench steps; returns device ms between one event pair on the launching stream"""
            with torch.cuda.stream(stream):
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ev0.record(stream)
                for k in range(count):
                    flush.zero_()  # L2 flush between bench steps (inside the timed region)
                    issuer.issue((first_stride + k) * stride, timed)
                ev1.record(stream)
            stream.synchronize()
            return ev0.elapsed_time(ev1)

        t_warm = region(0, warm, False)
        eng.sync()
        est = max(t_warm / warm * K * 1e-3, 1e-6)  # seconds per repetition
        reps = int(min(61, max(1, math.ceil(args.min_timed_s / est))))
        reps += 1 - reps % 2
        if world > 1:
            # every rank must issue the SAME number of strides: the sharded hydrolysis plan all-gathers inside a stride, and a
            # rank that stopped early would leave the others waiting (each rank's estimate comes from its own warm-up time)
            rt = torch.tensor([reps], dtype=torch.int64, device="cuda")
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
            reps = int(rt.item())
        launches0 = eng.launches
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        rep_ms = []
        issuer.window_events, issuer.md_in_windows = [], 0
        for r in range(reps):
            rep_ms.append(region(warm + r * K, K, True))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.summary()
        launches = (eng.launches - launches0) / reps
        win_ms = [a.elapsed_time(b) for a, b in issuer.window_events]
        win_md = issuer.md_in_windows
        eng.sync()
        # periodic ensemble statistics: per-trajectory energies reduced over ranks with NCCL (north_star e)
        en = eng.energies()
        esum = torch.tensor(en.sum(axis=0), dtype=torch.float64, device="cuda")
        t = torch.tensor(rep_ms, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(esum)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)  # per repetition, the slowest rank
        rep_max = sorted(t.tolist())
        ms_med = rep_max[len(rep_max) // 2]
        finite = bool(np.isfinite(eng.coords()).all())
        value = N * ntr_global * stride * K / (ms_med * 1e-3)
        eng.close()
        del flush

        # ---- e2e through the drop-in compute() with host buffers
        # DCD frames, mt_len.dat and hydrolysis.pdb are written like the reference executable does (background writer).
        # Two untimed runs of the SAME length first: on a fresh box the first full-length runs pay for first-touch of pinned
        # and tmpfs pages (single runs of 0.4 s and 2.6 s were seen before the 0.20 s steady state), like the W warm-up steps
        # of the device-timed leg.
        walls = []
        for rep in range(-2, 9):
            e2e_sys = make_system(args.workload, ntr_local, tmp / f"e2e_{rep}", [f"device={local}"], write_files=True, steps=K * stride)
            e2e_sys.srand(e2e_sys.par.rseed)
            if world > 1:
                dist.barrier()
            with workspace.chdir(tmp / f"e2e_{rep}"):
                t0 = time.perf_counter()
                st = e2e_sys.compute()
                if rep >= 0:
                    walls.append(time.perf_counter() - t0)
            e2e_sys.close()
            shutil.rmtree(tmp / f"e2e_{rep}", ignore_errors=True)
        tw = torch.tensor(walls, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        walls_max = sorted(tw.tolist())
        wall = walls_max[len(walls_max) // 2]
        e2e = {"value": N * ntr_global * stride * K / wall, "unit": UNIT,
               "h2d_bytes_per_step": st["h2d_bytes"] / K, "d2h_bytes_per_step": st["d2h_bytes"] / K,
               "call": "mt_system_compute (drop-in compute(): create + upload, fused windows, hydrolysis draws + uploads, asynchronous stride read-back, DCD output)",
               "md_steps": K * stride, "wall_s": wall, "wall_s_runs": [round(w, 4) for w in tw.tolist()], "statistic": "median of 9 runs (each the max over ranks) after 2 untimed runs of the same length",
               "scratch": str(tmp.parent)}

        # ---- N > 1: the product's own multi-GPU path, one host thread driving N handles
        one_host = None
        if world > 1:
            if cpu_group is not None:
                dist.barrier(group=cpu_group)
            if rank == 0:
                try:
                    ws = []
                    for rep in range(-2, 3):  # two untimed runs of the same length first (NCCL communicator, first-touch of pinned pages)
                        d = tmp / f"one_host_{rep}"
                        s1 = make_system(args.workload, ntr_global, d, ["device=0"], write_files=True, steps=K * stride)
                        s1.srand(s1.par.rseed)
                        with workspace.chdir(d):
                            t0 = time.perf_counter()
                            st1 = s1.compute(n_gpus=world)
                            if rep >= 0:
                                ws.append(time.perf_counter() - t0)
                        s1.close()
                        shutil.rmtree(d, ignore_errors=True)
                    w1 = sorted(ws)[1]
                    one_host = {"value": N * ntr_global * stride * K / w1, "unit": UNIT, "wall_s": w1, "wall_s_runs": [round(w, 4) for w in ws],
                                "h2d_bytes_per_step": st1["h2d_bytes"] / K, "d2h_bytes_per_step": st1["d2h_bytes"] / K,
                                "call": f"mt_system_compute(n_gpus={world}, runnum={ntr_global}): one host thread, {world} handles, global rand() order, "
                                        "NCCL ensemble statistics (maddy_ensemble_stats) in every stride block"}
                except Exception as ex:  # reported, not hidden
                    one_host = {"value": None, "reason": f"{type(ex).__name__}: {ex}"}
            if cpu_group is not None:
                dist.barrier(group=cpu_group)

        if rank == 0:
            n_win = max(len(win_ms), 1)
            avg_launch_ms = sum(win_ms) / n_win
            md_per_launch = win_md / n_win
            alg_per_launch = b_alg * N * ntr_local * md_per_launch
            achieved = alg_per_launch / (avg_launch_ms * 1e-3) / 1e9 if win_ms else None
            traffic = NCU_TRAFFIC_PER_LAUNCH.get((args.workload, ntr_local))
            cfg = common_config(args.workload, N, ntr_local, world, stride)
            cfg.update({"parallelism": f"trajectory-sharded x{world}",
                        "window_pattern": [n for _, n, _ in issuer.windows], "repetitions": reps})
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warm,
                "ms_per_step": ms_med / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": cfg,
                "timed_region_ms": {"median": ms_med, "repetitions": [round(x, 3) for x in t.tolist()],
                                    "what": f"{K} bench steps = {K * stride} MD steps per repetition, one event pair each, max over ranks"},
                "e2e": e2e, "gpu_launches": int(round(launches)), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / pk["hbm_gbs"] if achieved else None,
                             "traffic": traffic[0] * md_per_launch / traffic[1] if traffic else None,
                             "traffic_unit": "bytes per average launch, scaled from the ncu dram read+write of a 100-step window (profiles/r2_run_kernel_ncu_full.txt)",
                             "algorithmic_bytes_per_launch": alg_per_launch, "avg_md_steps_per_launch": md_per_launch,
                             "avg_launch_ms": avg_launch_ms, "launches_timed": len(win_ms),
                             "window_share_of_timed_region": sum(win_ms) / max(sum(rep_ms), 1e-9),
                             "algorithmic_bytes_per_bench_step": b_alg * N * ntr_local * stride,
                             "peak_source": pk_src,
                             "kernel": "maddy::run_kernel (one fused window of MD steps per launch; the rest of the region is the stride-step "
                                       "rebuild+energies launch, the snapshot kernel and the L2 flush fill)",
                             "algorithmic_bytes_per_monomer_step": b_alg, "terms": B_ALG_TERMS,
                             "note": "state stays on-chip across the fused steps, so DRAM traffic is far below the algorithmic bytes; "
                                     "the binding limit is SM issue/latency (see profiles/)"},
                "ensemble_energy_sum": [float(x) for x in esum.tolist()], "finite": finite,
            }
            if system.par.tea_on and win_ms:
                # TEA: the O(N^2) Rotne-Prager pair work is FP32-pipe bound (~45 flop per ordered pair); the
                # windows timed above hold force + prepare + pair kernel of every step, so this is a lower bound for the
                # pair kernel alone (profiles/ has its own duration)
                flop = 45.0 * (N - 1) * N * ntr_local * md_per_launch
                tf = flop / (avg_launch_ms * 1e-3) / 1e12
                fp32_peak = 148 * 128 * 2 * pk.get("sm_max_mhz", 1965.0) * 1e6 / 1e12
                line["roofline_fp32"] = {"bound": "fp32", "achieved": tf, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tf / fp32_peak,
                                         "flop_per_monomer_step": 45.0 * (N - 1), "peak_source": "148 SMs x 128 lanes x 2 flop x SM clock (nominal; "
                                         "the packed FFMA2 ceiling measured on this GPU is 66 TFLOP/s, profiles/r1_ffma2_microbench.txt)",
                                         "kernel": "maddy::tea_pair_kernel inside whole TEA steps (force + prepare launch included in the time)"}
            if world > 1:
                line["shard_parity"] = shard_parity
                line["e2e_one_host"] = one_host
            if world == 1 and not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(args.workload)
            emit(line)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


# --------------------------------------------------------------------------------------------- reference arm
_SAVING = re.compile(rb"Saving coordinates at step (\d+)")


class StrideClock(threading.Thread):
    """Reads the stdout of one reference process and time-stamps the line its update() prints in every stride block
    (updater.cpp:78).  With >= 8 KB of output per stride (Energies[..] / tubule[..] lines of >= 80 trajectories) a plain
    pipe delivers each stride's line inside that stride block; smaller runs get a pseudo-terminal so that the reference's
    stdout is line-buffered."""

    def __init__(self, fd: int):
        super().__init__(daemon=True)
        self.fd, self.stamps, self.tail = fd, {}, b""

    def run(self):
        while True:
            try:
                chunk = os.read(self.fd, 1 << 16)
            except OSError:
                break
            now = time.perf_counter()
            if not chunk:
                break
            data = self.tail + chunk
            for m in _SAVING.finditer(data):
                if m.end() < len(data):  # the number is complete
                    self.stamps.setdefault(int(m.group(1)), now)
            self.tail = data[-64:]


def _reference_batch(ref, jobs, timeout):
    """start one reference process per job (rundir, ntr) at once; -> (list of StrideClock, list of error strings)"""
    procs, errs = [], []
    for d, ntr in jobs:
        use_pty = ntr < 80
        if use_pty:
            import pty
            master, slave = pty.openpty()
            p = subprocess.Popen([str(ref), "config.conf"], cwd=str(d), stdout=slave, stderr=subprocess.PIPE)
            os.close(slave)
            fd = master
        else:
            p = subprocess.Popen([str(ref), "config.conf"], cwd=str(d), stdout=subprocess.PIPE, stderr=subprocess.PIPE)
            fd = p.stdout.fileno()
        clock = StrideClock(fd)
        clock.start()
        procs.append((p, clock))
    for p, clock in procs:
        try:
            p.wait(timeout=timeout)
        except subprocess.TimeoutExpired:
            p.kill()
            errs.append(f"timeout after {timeout} s")
        clock.join(timeout=10)
        if p.returncode:
            errs.append(f"exit code {p.returncode}: {p.stderr.read().decode(errors='replace')[-300:]}")
    return [c for _, c in procs], errs


def run_reference(args):
    """The unmodified reference CUDA binary on the same conf files; rank 0 drives every process.

    The reference's zs[100] array limits a process to 100 trajectories, so the args.ntr trajectories of a GPU are covered
    by the fewest equal-sized processes (256 -> 86/85/85).  Two ways to run them on the one GPU are measured and the
    FASTER is reported (the other is kept in `timing`): one process after the other per GPU (the GPUs in parallel), and
    all of a GPU's processes at once (time-sliced contexts)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref = ROOT / "oracle" / "_ref" / "mt"
    K, warm = args.steps, max(args.warmup, 3)
    if not ref.exists():
        emit({"impl": "reference", "unavailable": "oracle/_ref/mt not built (needs /root/reference at build time)"})
        return
    from mt_b200 import workspace
    N, stride, _, _, _ = workload_facts(args.workload)
    n_proc = max(1, math.ceil(args.ntr / REF_NTR_LIMIT))
    split = [args.ntr // n_proc + (1 if i < args.ntr % n_proc else 0) for i in range(n_proc)]
    steps_total = (warm + K) * stride + 1  # the stride block of step (warm + K) * stride prints the closing line
    s_a, s_b = warm * stride, (warm + K) * stride
    units = N * args.ntr * world * stride * K
    tmp = scratch_dir("bench_ref_")
    cfg = common_config(args.workload, N, args.ntr, world, stride)
    cfg.update({"parallelism": f"{n_proc} reference processes per GPU ({'/'.join(map(str, split))} trajectories: zs[100] limit of the reference) x {world} GPU(s)"})
    base = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": warm, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg}

    def rundirs(tag, i):
        out = []
        for g in range(world):
            d = tmp / f"{tag}_g{g}_p{i}"
            workspace.make_baseline_rundir(d, args.workload, runnum=split[i], steps=steps_total, device=g)
            out.append((d, split[i]))
        return out

    def spans(clocks):
        missing = [i for i, c in enumerate(clocks) if s_a not in c.stamps or s_b not in c.stamps]
        if missing:
            raise RuntimeError(f"stride lines {s_a}/{s_b} not seen from process(es) {missing}")
        return [c.stamps[s_a] for c in clocks], [c.stamps[s_b] for c in clocks]

    try:
        modes, singles = {}, None
        try:
            # (a) one process after the other on each GPU (all GPUs at the same time): time = sum over the rounds of the slowest GPU
            total, per_round = 0.0, []
            for i in range(n_proc):
                clocks, errs = _reference_batch(ref, rundirs("seq", i), args.ref_timeout)
                if errs:
                    raise RuntimeError("; ".join(errs))
                st, en = spans(clocks)
                per_round.append(max(en) - min(st))
                total += per_round[-1]
                if singles is None:
                    c0 = clocks[0].stamps
                    singles = sorted(c0[(warm + k + 1) * stride] - c0[(warm + k) * stride] for k in range(K)
                                     if (warm + k + 1) * stride in c0 and (warm + k) * stride in c0)
            modes["sequential"] = {"seconds": total, "value": units / total, "per_round_s": [round(x, 4) for x in per_round]}
        except Exception as ex:
            modes["sequential"] = {"value": None, "reason": f"{type(ex).__name__}: {ex}"}
        if n_proc > 1 and not args.ref_sequential_only:
            try:
                # (b) all processes of a GPU at once
                jobs = [j for i in range(n_proc) for j in rundirs("con", i)]
                clocks, errs = _reference_batch(ref, jobs, args.ref_timeout)
                if errs:
                    raise RuntimeError("; ".join(errs))
                st, en = spans(clocks)
                span = max(en) - min(st)
                modes["concurrent"] = {"seconds": span, "value": units / span, "concurrency_overlap": round((min(en) - max(st)) / span, 4)}
            except Exception as ex:
                modes["concurrent"] = {"value": None, "reason": f"{type(ex).__name__}: {ex}"}
        good = {k: v for k, v in modes.items() if v.get("value")}
        if not good:
            base.update({"value": None, "ms_per_step": None, "reason": "; ".join(f"{k}: {v.get('reason')}" for k, v in modes.items()), "timing": modes})
            emit(base)
            return
        best = max(good, key=lambda k: good[k]["value"])
        value, seconds = good[best]["value"], good[best]["seconds"]
        note = f"reference's own CUDA build (oracle/_ref/mt, unmodified sources, -arch=sm_100 -use_fast_math) on {world} x B200; {n_proc} processes " \
               f"per GPU run {best}ly; timed between its own 'Saving coordinates at step {s_a}' and '... {s_b}' lines ({K} strides), host cores {os.cpu_count()}; " \
               f"run directories under {tmp.parent}"
        base.update({"value": value, "ms_per_step": seconds / K * 1e3,
                     "timing": {"reported_mode": best, "modes": modes,
                                "single_stride_s_min_med_max_first_process": [round(singles[0], 4), round(singles[len(singles) // 2], 4), round(singles[-1], 4)] if singles else None,
                                "stdout": "pty (line-buffered)" if split[0] < 80 else "pipe"},
                     "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                      "sample": "the reference has no CPU implementation of the step loop; " + note},
                     "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        emit(base)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


_JSON_FD = None


def emit(line: dict):
    """Writes the one JSON line to the process's ORIGINAL stdout (see main())."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    # stdout carries exactly one line, the JSON result: anything a library prints on fd 1 while the run is in flight
    # (NCCL's version banner, for one) is sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20, help="bench steps to time; one bench step = one output stride = 1000 MD steps")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="mt40_ensemble")
    ap.add_argument("--ntr", type=int, default=256, help="trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-timed-s", type=float, default=8.0, help="repeat the K-step timed region until this much device time is covered (median reported)")
    ap.add_argument("--ref-timeout", type=float, default=1500.0)
    ap.add_argument("--ref-sequential-only", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
