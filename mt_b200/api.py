"""Python mirror of the reference's host interface over the C-ABI (thin: ctypes + numpy views).

  HostSystem  = what initParameters()/AssemblyInit() leave in the reference's globals
                (par, top, r): loaded by the C++ host from config.conf / forcefield / conditions.
  Engine      = one maddy_handle: the device side of compute() for a block of trajectories.

The names follow the reference (Ntot, Ntr, gtp, extra, on_tubule, harmonic/longitudinal/lateral/LJ).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import capi
from .capi import MaddyError, MaddyParams, MaddyTopology, HostParams, as_ptr


class HostSystem:
    """Config + topology + coordinates on the host (reference globals `par`, `top`, `r`)."""

    def __init__(self, config_path, overrides: Sequence[str] = (), quiet: bool = True, write_files: bool = False):
        self._h = C.c_void_p()
        arr = (C.c_char_p * max(1, len(overrides)))(*[o.encode() for o in overrides])
        flags = (capi.LOAD_QUIET if quiet else 0) | (0 if write_files else capi.LOAD_NO_FILES)
        if capi.hostlib.mt_system_load(str(config_path).encode(), len(overrides), arr, flags, C.byref(self._h)):
            raise MaddyError(1, capi.hostlib.mt_host_last_error().decode())
        self.par = MaddyParams()
        self.host = HostParams()
        self.refresh()

    def refresh(self):
        capi.hostlib.mt_system_params(self._h, C.byref(self.par), C.byref(self.host))
        self._top = MaddyTopology()
        capi.hostlib.mt_system_topology(self._h, C.byref(self._top))

    def close(self):
        if self._h:
            capi.hostlib.mt_system_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- sizes
    @property
    def Ntot(self) -> int:
        return self.par.n_tot

    @property
    def Ntr(self) -> int:
        return self.par.n_tr

    # ---- live views of the host arrays
    def _view(self, ptr, shape, dtype):
        n = int(np.prod(shape))
        if n == 0:
            return np.zeros(shape, dtype=dtype)
        ct = {np.float32: C.c_float, np.int32: C.c_int, np.uint8: C.c_ubyte, np.float64: C.c_double}[dtype]
        return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).reshape(shape)

    @property
    def coords(self) -> np.ndarray:
        """[Ntr, Ntot, 7] float32 in the reference's Coord order x,y,z,fi,theta,psi,w."""
        return self._view(capi.hostlib.mt_system_coords(self._h), (self.Ntr, self.Ntot, 7), np.float32)

    @property
    def gtp(self):
        return self._view(capi.hostlib.mt_system_gtp(self._h), (self.Ntr, self.Ntot), np.int32)

    @property
    def on_tubule_cur(self):
        return self._view(capi.hostlib.mt_system_on_tubule(self._h, 0), (self.Ntr, self.Ntot), np.int32)

    @property
    def on_tubule_prev(self):
        return self._view(capi.hostlib.mt_system_on_tubule(self._h, 1), (self.Ntr, self.Ntot), np.int32)

    @property
    def extra(self):
        return self._view(capi.hostlib.mt_system_extra(self._h), (self.Ntr, self.Ntot), np.uint8)

    @property
    def fixed(self):
        return self._view(self._top.fixed, (self.Ntot,), np.uint8)

    @property
    def mon_type(self):
        return self._view(self._top.mon_type, (self.Ntot,), np.int32)

    @property
    def harmonic(self):
        return self._view(self._top.harmonic, (self.Ntot, self.par.max_harmonic), np.int32)

    @property
    def harmonic_count(self):
        return self._view(self._top.harmonic_count, (self.Ntot,), np.int32)

    @property
    def longitudinal(self):
        return self._view(self._top.longitudinal, (self.Ntr, self.Ntot, self.par.max_longitudinal), np.int32)

    @property
    def longitudinal_count(self):
        return self._view(self._top.longitudinal_count, (self.Ntr, self.Ntot), np.int32)

    @property
    def lateral(self):
        return self._view(self._top.lateral, (self.Ntr, self.Ntot, self.par.max_lateral), np.int32)

    @property
    def lateral_count(self):
        return self._view(self._top.lateral_count, (self.Ntr, self.Ntot), np.int32)

    @property
    def energies(self):
        return self._view(capi.hostlib.mt_system_energies(self._h), (self.Ntr, 7), np.float64)

    @property
    def ensemble_stats(self):
        """[16] all-reduced energy statistics of the last stride (sum(7), sum of squares(7), count, 0), or None"""
        p = capi.hostlib.mt_system_ensemble_stats(self._h)
        return np.ctypeslib.as_array(p, shape=(16,)).copy() if p else None

    def topology(self, traj_first: int = 0) -> MaddyTopology:
        """maddy_topology for the trajectories starting at traj_first (pointer arithmetic on the live arrays)."""
        t = MaddyTopology()
        C.memmove(C.byref(t), C.byref(self._top), C.sizeof(MaddyTopology))
        o = traj_first * self.Ntot

        def adv(ptr, ctype, off):
            if not ptr:
                return ptr
            return C.cast(C.addressof(ptr.contents) + off * C.sizeof(ctype), C.POINTER(ctype))

        t.longitudinal_count = adv(t.longitudinal_count, C.c_int, o)
        t.longitudinal = adv(t.longitudinal, C.c_int, o * self.par.max_longitudinal)
        t.lateral_count = adv(t.lateral_count, C.c_int, o)
        t.lateral = adv(t.lateral, C.c_int, o * self.par.max_lateral)
        t.extra = adv(t.extra, C.c_ubyte, o)
        t.gtp = adv(t.gtp, C.c_int, o)
        t.on_tubule_cur = adv(t.on_tubule_cur, C.c_int, o)
        return t

    # ---- the drop-in compute() and host events
    def compute(self, fused: bool = True, n_gpus: Optional[int] = None, steps: Optional[int] = None):
        if n_gpus is not None:
            capi.hostlib.mt_system_set_ngpus(self._h, int(n_gpus))
        if steps is not None:
            capi.hostlib.mt_system_set_steps(self._h, int(steps))
            self.refresh()
        st = (C.c_double * 4)()
        if capi.hostlib.mt_system_compute(self._h, int(fused), st):
            raise MaddyError(1, capi.hostlib.mt_host_last_error().decode())
        return {"steps": int(st[0]), "launches": int(st[1]), "h2d_bytes": st[2], "d2h_bytes": st[3]}

    def srand(self, seed: int):
        """srand(seed) for this system's host events (hydrolysis, insertion): same sequence as libc rand()"""
        capi.hostlib.mt_system_srand(self._h, int(seed) & 0xffffffff)

    def rand_window(self) -> np.ndarray:
        """31 words of the host rand() stream, oldest first (what Engine.hydrolysis_plan takes)"""
        w = np.zeros(31, dtype=np.uint32)
        capi.hostlib.mt_system_rand_window(self._h, as_ptr(w, C.c_uint))
        return w

    def rand_discard(self, n: int):
        capi.hostlib.mt_system_rand_discard(self._h, int(n))

    def rand_next(self) -> int:
        return int(capi.hostlib.mt_system_rand_next(self._h))

    def mt_length(self, step: int) -> np.ndarray:
        out = np.zeros(self.Ntr, dtype=np.int32)
        if capi.hostlib.mt_system_mt_length(self._h, int(step), as_ptr(out, C.c_int)):
            raise MaddyError(1, capi.hostlib.mt_host_last_error().decode())
        return out

    def hydrolyse(self):
        if capi.hostlib.mt_system_hydrolyse(self._h):
            raise MaddyError(1, capi.hostlib.mt_host_last_error().decode())

    def change_conc(self, delta, mt_len) -> int:
        d = np.ascontiguousarray(delta, dtype=np.int32)
        m = np.ascontiguousarray(mt_len, dtype=np.int32)
        ch = C.c_int()
        if capi.hostlib.mt_system_change_conc(self._h, as_ptr(d, C.c_int), as_ptr(m, C.c_int), C.byref(ch)):
            raise MaddyError(1, capi.hostlib.mt_host_last_error().decode())
        return ch.value


class Engine:
    """One maddy_handle: device state + kernels for trajectories [traj_first, traj_first + n_tr_local)."""

    def __init__(self, system: HostSystem, traj_first: int = 0, n_tr_local: Optional[int] = None, device: Optional[int] = None,
                 stream: Optional[int] = None, coords: Optional[np.ndarray] = None, par: Optional[MaddyParams] = None):
        self.system = system
        p = (par or system.par).copy()
        p.traj_first = traj_first
        p.n_tr_local = system.Ntr - traj_first if n_tr_local is None else n_tr_local
        if device is not None:
            p.device = device
        self.par = p
        self.N, self.ntr = p.n_tot, p.n_tr_local
        top = system.topology(traj_first)
        if coords is None:
            coords = system.coords[traj_first:traj_first + self.ntr]
        c = np.ascontiguousarray(coords, dtype=np.float32).reshape(self.ntr, self.N, 7)
        self._h = C.c_void_p()
        rc = capi.lib.maddy_create(C.byref(p), C.byref(top), as_ptr(c, C.c_float), C.c_void_p(stream or 0), C.byref(self._h))
        if rc:
            raise MaddyError(rc, capi.lib.maddy_last_error(None).decode())

    def _ck(self, rc):
        if rc:
            raise MaddyError(rc, capi.lib.maddy_last_error(self._h).decode())

    def close(self):
        if self._h:
            capi.lib.maddy_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def stream(self) -> int:
        return capi.lib.maddy_stream(self._h) or 0

    @property
    def launches(self) -> int:
        return capi.lib.maddy_launch_count(self._h)

    def sync(self):
        self._ck(capi.lib.maddy_sync(self._h))

    # step-granular (one per reference launch)
    def rebuild_lj(self):
        self._ck(capi.lib.maddy_rebuild_lj(self._h))

    def rebuild_bonds(self):
        self._ck(capi.lib.maddy_rebuild_bonds(self._h))

    def force(self):
        self._ck(capi.lib.maddy_force(self._h))

    def integrate(self):
        self._ck(capi.lib.maddy_integrate(self._h))

    def tea_update(self, step: int):
        self._ck(capi.lib.maddy_tea_update(self._h, int(step)))

    def tea_integrate(self):
        self._ck(capi.lib.maddy_tea_integrate(self._h))

    def tea_state(self):
        """(C_i [ntr*N, 4], epsilon [ntr*N], beta [ntr]) of the TEA integrator (d_ci, d_epsilon, d_beta_ij of the reference)"""
        ci = np.empty((self.ntr * self.N, 4), dtype=np.float32)
        eps = np.empty(self.ntr * self.N, dtype=np.float32)
        beta = np.empty(self.ntr, dtype=np.float32)
        self._ck(capi.lib.maddy_download_tea(self._h, as_ptr(ci, C.c_float), as_ptr(eps, C.c_float), as_ptr(beta, C.c_float)))
        return ci, eps, beta

    def run(self, first_step: int, n_steps: int, skip_first_rebuild: bool = False):
        self._ck(capi.lib.maddy_run(self._h, int(first_step), int(n_steps),
                                    capi.RUN_SKIP_FIRST_REBUILD if skip_first_rebuild else 0))

    def energies(self, per_monomer: bool = False):
        out = np.empty((self.ntr, 7), dtype=np.float64)
        mono = np.empty((self.ntr, self.N, 7), dtype=np.float64) if per_monomer else None
        self._ck(capi.lib.maddy_energies(self._h, as_ptr(out, C.c_double), as_ptr(mono, C.c_double) if per_monomer else None))
        return (out, mono) if per_monomer else out

    def rebuild_and_energies(self):
        """list rebuild + energies of the reference's stride block in one launch"""
        out = np.empty((self.ntr, 7), dtype=np.float64)
        self._ck(capi.lib.maddy_rebuild_and_energies(self._h, as_ptr(out, C.c_double), None))
        return out

    def list_stats(self, reset: bool = False) -> dict:
        out = (C.c_ulonglong * 4)()
        self._ck(capi.lib.maddy_list_stats(self._h, out, int(reset)))
        return {"near_refresh": out[0], "candidate_rescan": out[1], "all_pairs_fallback": out[2], "near_overflow": out[3]}

    def snapshot_begin(self, coords=True, forces=False, energies=True, rebuild=False, on_tubule=False, apply_on_tubule=False, gtp=False, guard=False):
        """queue the stride read-back (maddy_snapshot_begin); work queued afterwards overlaps with snapshot_end()"""
        what = (capi.SNAP_COORDS if coords else 0) | (capi.SNAP_FORCES if forces else 0) | (capi.SNAP_ENERGIES if energies else 0) \
            | (capi.SNAP_REBUILD if rebuild else 0) | (capi.SNAP_ONTUBULE if on_tubule or apply_on_tubule else 0) \
            | (capi.SNAP_ONTUBULE_APPLY if apply_on_tubule else 0) | (capi.SNAP_GTP if gtp else 0) \
            | (capi.SNAP_ONTUBULE_GUARD if guard else 0)
        self._snap = what
        self._ck(capi.lib.maddy_snapshot_begin(self._h, what))

    def snapshot_end(self):
        """-> dict with the arrays requested in snapshot_begin (state as of the point where it was queued)"""
        what = self._snap
        c = np.empty((self.ntr, self.N, 7), dtype=np.float32) if what & capi.SNAP_COORDS else None
        f = np.empty((self.ntr, self.N, 7), dtype=np.float32) if what & capi.SNAP_FORCES else None
        e = np.empty((self.ntr, 7), dtype=np.float64) if what & capi.SNAP_ENERGIES else None
        self._ck(capi.lib.maddy_snapshot_end(self._h, as_ptr(c, C.c_float) if c is not None else None,
                                             as_ptr(f, C.c_float) if f is not None else None,
                                             as_ptr(e, C.c_double) if e is not None else None))
        out = {"coords": c, "forces": f, "energies": e}
        if what & capi.SNAP_ONTUBULE:
            on = np.empty((self.ntr, self.N), dtype=np.int32)
            ln = np.empty(self.ntr, dtype=np.int32)
            self._ck(capi.lib.maddy_snapshot_on_tubule(self._h, as_ptr(on, C.c_int), as_ptr(ln, C.c_int)))
            out["on_tubule"], out["mt_len"] = on, ln
        if what & capi.SNAP_GTP:
            g = np.empty((self.ntr, self.N), dtype=np.int32)
            self._ck(capi.lib.maddy_snapshot_gtp(self._h, as_ptr(g, C.c_int)))
            out["gtp"] = g
        return out

    # ---- hydrolysis events of a stride on the device (maddy_hydrolysis_plan)
    def hydrolysis_plan(self, rand_window, first_event: int, period: int, n_events: int, keep_slots: bool = False):
        w = np.ascontiguousarray(rand_window, dtype=np.uint32)
        assert w.size == 31
        self._hyd = (int(n_events), bool(keep_slots))
        self._ck(capi.lib.maddy_hydrolysis_plan(self._h, as_ptr(w, C.c_uint), int(first_event), int(period), int(n_events),
                                                capi.HYD_KEEP_SLOTS if keep_slots else 0))

    def hydrolysis_inputs(self):
        """(device pointer, bytes) of this shard's transposed plan inputs, prepared in stream order (maddy_hydrolysis_inputs)"""
        ptr, nb = C.c_void_p(), C.c_ulong()
        self._ck(capi.lib.maddy_hydrolysis_inputs(self._h, C.byref(ptr), C.byref(nb)))
        return int(ptr.value), int(nb.value)

    def hydrolysis_plan_sharded(self, gathered_ptr: int, n_shards: int, rand_window, first_event: int, period: int, n_events: int, keep_slots=False):
        w = np.ascontiguousarray(rand_window, dtype=np.uint32)
        self._hyd = (int(n_events), bool(keep_slots))
        self._ck(capi.lib.maddy_hydrolysis_plan_sharded(self._h, C.c_void_p(gathered_ptr), int(n_shards), as_ptr(w, C.c_uint), int(first_event),
                                                        int(period), int(n_events), capi.HYD_KEEP_SLOTS if keep_slots else 0))

    def hydrolysis_result(self):
        """-> (draws_total, first draw of every event, slots [n_events, ntr, N] or None)"""
        ne, keep = self._hyd
        total = C.c_ulonglong()
        first = np.zeros(ne, dtype=np.uint64)
        slots = np.empty((ne, self.ntr, self.N), dtype=np.int32) if keep else None
        self._ck(capi.lib.maddy_hydrolysis_result(self._h, C.byref(total), as_ptr(first, C.c_ulonglong), as_ptr(slots, C.c_int) if keep else None))
        return int(total.value), first, slots

    def clear_guard(self):
        self._ck(capi.lib.maddy_clear_guard(self._h))

    def apply_scheduled_gtp(self, step: int):
        self._ck(capi.lib.maddy_apply_scheduled_gtp(self._h, int(step)))

    def snapshot_tubule_lengths(self):
        """(mt_len [ntr], undecided) of the snapshot in flight: waits for the counts only (maddy_snapshot_tubule_lengths)"""
        ln = np.empty(self.ntr, dtype=np.int32)
        und = C.c_int()
        self._ck(capi.lib.maddy_snapshot_tubule_lengths(self._h, as_ptr(ln, C.c_int), C.byref(und)))
        return ln, und.value

    def insert_dimers(self, index, xyzz):
        """sparse constant-concentration insertion (maddy_insert_dimers): index [k] local first-monomer ids, xyzz [k, 4]"""
        idx = np.ascontiguousarray(index, dtype=np.int32)
        v = np.ascontiguousarray(xyzz, dtype=np.float32).reshape(-1, 4)
        self._ck(capi.lib.maddy_insert_dimers(self._h, int(idx.size), as_ptr(idx, C.c_int), as_ptr(v, C.c_float)))

    # ---- in-situ analysis (scripts/temp_calc, scripts/disas_speed of the reference)
    def analysis_setup(self, chain, resid, name1, n_pf: int = 13):
        """PDB labels per monomer: chain index (chain - 'A', -1 outside the protofilaments), residue number, second
        character of the atom name (see pdb_labels)."""
        chain = np.ascontiguousarray(chain, dtype=np.int32)
        resid = np.ascontiguousarray(resid, dtype=np.int32)
        name1 = bytes(name1)
        assert chain.size == self.N and resid.size == self.N and len(name1) == self.N
        self._npf = int(n_pf)
        self._ck(capi.lib.maddy_analysis_setup(self._h, as_ptr(chain, C.c_int), as_ptr(resid, C.c_int), name1, self._npf))

    def analysis_reference(self):
        self._ck(capi.lib.maddy_analysis_reference(self._h))

    def analysis_temperature(self) -> np.ndarray:
        """raw displacement sums [ntr, 8] against the previous frame, which then becomes the current state"""
        out = np.empty((self.ntr, 8), dtype=np.float64)
        self._ck(capi.lib.maddy_analysis_temperature(self._h, as_ptr(out, C.c_double)))
        return out

    def analysis_project(self) -> np.ndarray:
        out = np.empty((self.ntr, self.N, 3), dtype=np.float32)
        self._ck(capi.lib.maddy_analysis_project(self._h, as_ptr(out, C.c_float)))
        return out

    def analysis_protofilaments(self) -> np.ndarray:
        """[ntr, n_pf, 3] = pf_end_number, curled_start, mt_end_number per protofilament"""
        out = np.empty((self.ntr, self._npf, 3), dtype=np.int32)
        self._ck(capi.lib.maddy_analysis_protofilaments(self._h, as_ptr(out, C.c_int)))
        return out

    @property
    def energies_device_ptr(self) -> int:
        return capi.lib.maddy_energies_device(self._h) or 0

    def coords(self) -> np.ndarray:
        out = np.empty((self.ntr, self.N, 7), dtype=np.float32)
        self._ck(capi.lib.maddy_download_coords(self._h, as_ptr(out, C.c_float)))
        return out

    def forces(self) -> np.ndarray:
        out = np.empty((self.ntr, self.N, 7), dtype=np.float32)
        self._ck(capi.lib.maddy_download_forces(self._h, as_ptr(out, C.c_float)))
        return out

    def upload_coords(self, c):
        c = np.ascontiguousarray(c, dtype=np.float32)
        self._ck(capi.lib.maddy_upload_coords(self._h, as_ptr(c, C.c_float)))

    def upload_gtp(self, g):
        g = np.ascontiguousarray(g, dtype=np.int32)
        self._ck(capi.lib.maddy_upload_gtp(self._h, as_ptr(g, C.c_int)))

    def schedule_gtp(self, first_event: int, period: int, slots):
        """slots[k] ([ntr, N] ints) becomes the GTP state at the start of step first_event + k*period (see maddy_schedule_gtp)"""
        g = np.ascontiguousarray(slots, dtype=np.int32)
        n = 0 if g.size == 0 else g.reshape(-1, self.ntr * self.N).shape[0]
        self._ck(capi.lib.maddy_schedule_gtp(self._h, int(first_event), int(period), n, as_ptr(g, C.c_int) if n else None))

    def upload_on_tubule(self, g):
        g = np.ascontiguousarray(g, dtype=np.int32)
        self._ck(capi.lib.maddy_upload_on_tubule(self._h, as_ptr(g, C.c_int)))

    def upload_extra(self, e):
        e = np.ascontiguousarray(e, dtype=np.uint8)
        self._ck(capi.lib.maddy_upload_extra(self._h, as_ptr(e, C.c_ubyte)))

    def _cap(self, kind):
        return {capi.LIST_LONGITUDINAL: self.par.max_longitudinal, capi.LIST_LATERAL: self.par.max_lateral,
                capi.LIST_LJ: capi.LJ_CAPACITY}[kind]

    def download_list(self, kind: int):
        """(counts [ntr, N], entries [ntr, N, capacity]) in the reference encoding."""
        cnt = np.zeros((self.ntr, self.N), dtype=np.int32)
        ent = np.zeros((self.ntr, self.N, self._cap(kind)), dtype=np.int32)
        self._ck(capi.lib.maddy_download_list(self._h, kind, as_ptr(cnt, C.c_int), as_ptr(ent, C.c_int)))
        return cnt, ent

    def upload_list(self, kind: int, counts, entries):
        cnt = np.ascontiguousarray(counts, dtype=np.int32)
        ent = np.ascontiguousarray(entries, dtype=np.int32)
        self._ck(capi.lib.maddy_upload_list(self._h, kind, as_ptr(cnt, C.c_int), as_ptr(ent, C.c_int)))

    def rng_state(self) -> np.ndarray:
        out = np.empty((2, self.ntr * self.N, 4), dtype=np.uint32)
        self._ck(capi.lib.maddy_download_rng(self._h, as_ptr(out, C.c_uint)))
        return out

    def upload_rng(self, st):
        st = np.ascontiguousarray(st, dtype=np.uint32)
        self._ck(capi.lib.maddy_upload_rng(self._h, as_ptr(st, C.c_uint)))


def pdb_labels(path, n_pf: int = 13):
    """(chain index, residue number, second character of the atom name) per ATOM record of a PDB file, in the form
    maddy_analysis_setup takes them (fixed PDB columns: name 13-16, chain 22, resSeq 23-26)."""
    chain, resid, name1 = [], [], bytearray()
    with open(path) as f:
        for line in f:
            if not line.startswith(("ATOM", "HETATM")):
                continue
            c = ord(line[21]) - ord("A")
            chain.append(c if 0 <= c < n_pf else -1)
            resid.append(int(line[22:26]))
            name = line[12:16].strip()
            name1.append(ord(name[1]) if len(name) > 1 else 32)
    return np.asarray(chain, dtype=np.int32), np.asarray(resid, dtype=np.int32), bytes(name1)
