#!/usr/bin/env python
"""bench.py — monomer-steps/s of the MADDY Langevin/BD step loop on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--ntr T]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W      (N > 1)

One bench "step" = one MD step of the whole ensemble (Ntot * Ntr monomer-steps): list rebuild every
LJPairsUpdateFreq steps, force evaluation, Langevin integration.  Workload at every N: the
configuration the metric is quoted on — BASELINE.json configs[1], the 13-protofilament seed
(520 monomers) x 256 trajectories PER GPU (weak scaling; the global ensemble of 256*N trajectories is
sharded in contiguous blocks and every shard uses the GLOBAL RNG stream ids).

value   device-timed: state resident in HBM, K steps issued as fused maddy_run() windows of
        `hydrostep` (=100) steps — the launch granularity the host events allow; CUDA events on the
        launching stream, max over ranks; L2 flushed between windows (outside the event pairs).
e2e     the drop-in compute() call of the C++ host (mt_system_compute) over K steps with HOST buffers:
        device allocation + upload of coordinates/topology/seeds, hydrolysis uploads every 100 steps,
        energies + coordinate download every `stride` steps, all inside the timed region (wall clock).
roofline  algorithmic bytes of the step-granular contract (352 B per monomer-step on the
        intact lattice) / device time, against the measured HBM copy bandwidth.  The fused kernel keeps
        the state on-chip, so `traffic` (ncu dram bytes) is far BELOW the algorithmic bytes.
cpu_baseline  the CPU oracle port (oracle/maddy_oracle.c, OpenMP) on a bounded sample, rank 0, N = 1.
--impl reference  the reference has NO CPU path: its own CUDA build (oracle/_ref/mt, sm_100 recompile,
        unmodified sources) is run on the GPU(s) with the same conf files; runnum is capped at 100 per
        process by the reference's zs[100] array and the value is per monomer-step.
"""
from __future__ import annotations

import argparse
import json
import os
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
B_ALG = 352.0  # algorithmic bytes per monomer-step, intact lattice (BASELINE.md 3 / SURVEY.md 8d)
B_ALG_TERMS = "64 state r/w + 64 RNG r/w + 4*(46+1) LJ list + 4*(4+3) bond lists + 8 flags"
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (a 100-step fused window at 520 x 256) from the
# `ncu --set full` capture summarised in profiles/r1_run_kernel_ncu_full.txt (32.8 MB read + 0.13 MB written)
NCU_TRAFFIC_PER_LAUNCH = {("mt40_ensemble", 256, 100): 32.9e6}
REF_NTR_LIMIT = 100


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


def cpu_baseline(workload: str, budget_s: float = 12.0):
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
        steps = int(max(40, min(4000, budget_s / max(per_step, 1e-6))))
        steps -= steps % 20
        t0 = time.perf_counter()
        o.run(60, steps)
        dt = time.perf_counter() - t0
        return {"value": s.Ntot * ntr * steps / dt, "unit": UNIT, "cores": min(cores, ntr), "kind": "port",
                "sample": f"oracle/maddy_oracle.c (OpenMP over trajectories), {ntr} trajectories x {s.Ntot} monomers x {steps} steps of {workload}"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def run_own(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from mt_b200 import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk, pk_src = peaks()
    ntr_local = args.ntr
    ntr_global = ntr_local * world
    tmp = scratch_dir(f"bench_r{rank}_")
    try:
        system = make_system(args.workload, ntr_global, tmp)
        N = system.Ntot
        window = int(system.host.hydrostep) if system.host.hydrolysis and system.host.hydrostep > 0 else int(system.host.stride)
        window = max(1, min(window, int(system.host.stride)))
        stream = torch.cuda.Stream()
        eng = Engine(system, traj_first=rank * ntr_local, n_tr_local=ntr_local, device=local, stream=stream.cuda_stream)
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

        stride = int(system.host.stride)
        multi = bool(system.host.hydrolysis) and window < stride and not os.environ.get("MADDY_SINGLE_EVENT_WINDOWS")

        def run_steps(first, count, timed):
            """count steps as fused windows; returns summed event time (ms) when timed.  The window lengths are those of
            the drop-in loop (mt_b200/host/events.cpp): one hydrolysis period after a stride step, then one period longer
            per window up to the next stride (100, 100, 200, 300, 300 steps at the template's periods)."""
            evs = []
            s = first
            last = window
            while s < first + count:
                if multi and s % stride != 0:
                    room = stride - s % stride
                    grow = 0 if (s - last) % stride == 0 else window  # the window after the post-stride one stays short
                    n = min((last // window) * window + grow, room)
                else:
                    n = window - (s % window)
                n = min(n, first + count - s)
                last = n
                with torch.cuda.stream(stream):
                    flush.zero_()  # L2 flush, outside the event pair
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(stream)
                    eng.run(s, n)
                    b.record(stream)
                evs.append((a, b))
                s += n
            stream.synchronize()
            return sum(a.elapsed_time(b) for a, b in evs) if timed else 0.0

        warm = max(args.warmup, 3)
        run_steps(0, warm, False)
        eng.sync()
        launches0 = eng.launches
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        sampler = ClockSampler(local)
        sampler.start()
        ms = run_steps(warm, args.steps, True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        clocks = sampler.summary()
        launches = eng.launches - launches0
        eng.sync()
        # periodic ensemble statistics: per-trajectory energies reduced over ranks with NCCL (north_star e)
        en = eng.energies()
        esum = torch.tensor(en.sum(axis=0), dtype=torch.float64, device="cuda")
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(esum)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        finite = bool(np.isfinite(eng.coords()).all())
        value = N * ntr_global * args.steps / (ms_max * 1e-3)

        # ---- e2e through the drop-in compute() with host buffers
        from mt_b200 import workspace
        e2e_sys = make_system(args.workload, ntr_local, tmp / "e2e", [f"device={local}"])
        e2e_sys.compute(steps=min(args.steps, 200))  # untimed: module load / context warm-up
        e2e_sys.close()
        # DCD frames, mt_len.dat and hydrolysis.pdb are written like the reference executable does (background writer).
        # The host side (file system, scheduler) makes single runs noisy (0.95 .. 1.6 s on a 16-core box): five runs, the
        # median is reported.
        walls = []
        for rep in range(5):
            e2e_sys = make_system(args.workload, ntr_local, tmp / f"e2e_{rep}", [f"device={local}"], write_files=True, steps=args.steps)
            e2e_sys.srand(e2e_sys.par.rseed)
            if world > 1:
                dist.barrier()
            with workspace.chdir(tmp / f"e2e_{rep}"):
                t0 = time.perf_counter()
                st = e2e_sys.compute()
                walls.append(time.perf_counter() - t0)
            e2e_sys.close()
            shutil.rmtree(tmp / f"e2e_{rep}", ignore_errors=True)
        wall = sorted(walls)[len(walls) // 2]
        tw = torch.tensor([wall], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        e2e = {"value": N * ntr_global * args.steps / float(tw.item()), "unit": UNIT,
               "h2d_bytes_per_step": st["h2d_bytes"] / args.steps, "d2h_bytes_per_step": st["d2h_bytes"] / args.steps,
               "call": "mt_system_compute (drop-in compute(): create + upload, fused windows, hydrolysis uploads, asynchronous stride read-back, DCD output)",
               "wall_s": float(tw.item()), "wall_s_runs": [round(w, 4) for w in walls], "statistic": "median of 5 runs",
               "scratch": str(tmp.parent)}

        if rank == 0:
            achieved = value / world * B_ALG / 1e9  # per-GPU algorithmic GB/s of the dominant (only) kernel
            line = {
                "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"{args.workload}: 13-PF MT seed, {N} monomers x {ntr_local} trajectories per GPU "
                                       f"({ntr_global} total), Morse + LJ, dynamic bond lists, dt 200",
                           "ntot": N, "ntr_per_gpu": ntr_local, "ntr_total": ntr_global, "window_steps": window,
                           "window_pattern": "per stride: one hydrolysis period, then one period longer per window (the drop-in loop's windows)" if multi else "one hydrolysis period",
                           "parallelism": f"trajectory-sharded x{world}",
                           "l2": "256 MiB flush write between fused windows (outside the timed event pairs); "
                                 "within a window the state is register/SMEM resident by design"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                             "frac": achieved / pk["hbm_gbs"], "traffic": NCU_TRAFFIC_PER_LAUNCH.get((args.workload, ntr_local, window)),
                             "traffic_unit": "bytes per launch of a 100-step window (ncu dram read+write, profiles/r1_run_kernel_ncu_full.txt)",
                             "algorithmic_bytes_per_launch": B_ALG * N * ntr_local * args.steps / max(launches, 1), "avg_steps_per_launch": args.steps / max(launches, 1), "peak_source": pk_src,
                             "kernel": "maddy::run_kernel<1,2> (one fused window of `window_steps` MD steps per launch; 92 % of the kernel time in the ncu launch list profiles/r1_launches.csv; the rest is the stride-step rebuild+energies launch and the L2 flush fill)",
                             "algorithmic_bytes_per_monomer_step": B_ALG, "terms": B_ALG_TERMS,
                             "note": "state stays on-chip across the fused steps, so DRAM traffic is far below the algorithmic bytes; "
                                     "the binding limit is SM issue/latency (see profiles/)"},
                "ensemble_energy_sum": [float(x) for x in esum.tolist()], "finite": finite,
            }
            if world == 1 and not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(args.workload)
            emit(line)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def run_reference(args):
    """The unmodified reference CUDA binary on the same conf files; rank 0 drives one process per GPU."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    ref = ROOT / "oracle" / "_ref" / "mt"
    if not ref.exists():
        emit({"impl": "reference", "unavailable": "oracle/_ref/mt not built (needs /root/reference at build time)"})
        return
    from mt_b200 import workspace
    ntr = min(args.ntr, REF_NTR_LIMIT)
    warm = max(args.warmup, 3)
    tmp = scratch_dir("bench_ref_")

    def timed(steps: int) -> float:
        procs, t0 = [], time.perf_counter()
        for g in range(world):
            d = tmp / f"g{g}_{steps}"
            workspace.make_baseline_rundir(d, args.workload, runnum=ntr, steps=steps, device=g)
            procs.append(subprocess.Popen([str(ref), "config.conf"], cwd=str(d), stdout=subprocess.DEVNULL, stderr=subprocess.PIPE))
        for p in procs:
            _, err = p.communicate(timeout=3000)
            if p.returncode != 0:
                raise RuntimeError(f"reference mt failed: {err.decode()[-500:]}")
        return time.perf_counter() - t0

    try:
        timed(warm)                      # cold start (driver/module load)
        t_a = min(timed(warm), timed(warm))
        t_b = timed(warm + args.steps)
        dt = max(t_b - t_a, 1e-9)        # difference of two run lengths removes initialisation and file set-up
        N = 520 if "40" in args.workload else None
        if N is None:
            from mt_b200 import structures
            N = len(structures.lattice(40, 0)[0])
        value = N * ntr * world * args.steps / dt
        cores_note = f"reference's own CUDA build (oracle/_ref/mt, unmodified sources, -arch=sm_100 -use_fast_math), {world} x B200, " \
                     f"runnum {ntr} per process (zs[100] limit of the reference), wall clock of {warm + args.steps} minus {warm} steps, run directories under {tmp.parent}"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": f"{args.workload}: 13-PF MT seed, {N} monomers x {ntr} trajectories per GPU (reference limit 100)",
                           "ntot": N, "ntr_per_gpu": ntr, "ntr_total": ntr * world, "parallelism": f"independent processes x{world}"},
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": 0, "kind": "reference",
                                 "sample": "the reference has no CPU implementation of the step loop; " + cores_note},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        emit(line)
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
    ap.add_argument("--steps", type=int, default=100000, help="MD steps to time (default: the 1e5-step run length of the BASELINE configs)")
    ap.add_argument("--warmup", type=int, default=1000)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--workload", default="mt40_ensemble")
    ap.add_argument("--ntr", type=int, default=256, help="trajectories per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
