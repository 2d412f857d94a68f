"""ctypes wrapper of the CPU oracle (oracle/maddy_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference
legs) — never by mt_b200/.  Works on the reference's own layouts (AoS7 coordinates, reference
list encodings) so results compare 1:1 with mt_b200.Engine downloads and with ref_probe dumps.
"""
from __future__ import annotations

import ctypes as C
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200.capi import MaddyParams, MaddyTopology, LJ_CAPACITY, as_ptr  # noqa: E402  (struct definitions only)

LIB_PATH = ROOT / "oracle" / "_build" / "libmaddy_oracle.so"


class OracleLists(C.Structure):
    _fields_ = [("longitudinal_count", C.POINTER(C.c_int)), ("longitudinal", C.POINTER(C.c_int)),
                ("lateral_count", C.POINTER(C.c_int)), ("lateral", C.POINTER(C.c_int)),
                ("lj_count", C.POINTER(C.c_int)), ("lj", C.POINTER(C.c_int))]


def _lib():
    if not LIB_PATH.exists():
        from mt_b200 import build
        build.build_oracle()
    lib = C.CDLL(str(LIB_PATH))
    pp, pt, pl = C.POINTER(MaddyParams), C.POINTER(MaddyTopology), C.POINTER(OracleLists)
    pf, pd, pu, pi = C.POINTER(C.c_float), C.POINTER(C.c_double), C.POINTER(C.c_uint), C.POINTER(C.c_int)
    lib.oracle_generate_seeds.argtypes = [pu, C.c_int, C.c_longlong]
    lib.oracle_hybrid_taus.argtypes = [pu]
    lib.oracle_hybrid_taus.restype = C.c_uint
    lib.oracle_rforce.argtypes = [pu, pf]
    lib.oracle_lj_lists.argtypes = [pp, pt, pf, pl]
    lib.oracle_pair_lists.argtypes = [pp, pt, pf, pl]
    lib.oracle_forces.argtypes = [pp, pt, pl, pf, pf]
    lib.oracle_energies.argtypes = [pp, pt, pl, pf, pd]
    lib.oracle_integrate.argtypes = [pp, pt, pf, pf, pu]
    lib.oracle_run.argtypes = [pp, pt, pl, pf, pf, pu, C.c_longlong, C.c_longlong, C.c_int]
    lib.oracle_tea_beta.argtypes = [C.c_double, C.c_int, C.c_int, C.c_float, C.c_float, pf, pd]
    lib.oracle_tea_update.argtypes = [pp, pt, pf, pf, pf, pf]
    lib.oracle_tea_integrate.argtypes = [pp, pt, pf, pf, pu, pf, pf]
    for n in ("oracle_generate_seeds", "oracle_rforce", "oracle_lj_lists", "oracle_pair_lists", "oracle_forces", "oracle_energies",
              "oracle_integrate", "oracle_run", "oracle_tea_integrate"):
        getattr(lib, n).restype = None
    lib.oracle_tea_beta.restype = C.c_int
    lib.oracle_tea_update.restype = C.c_int
    return lib


lib = _lib()


def generate_seeds(rseed: int, n: int) -> np.ndarray:
    out = np.empty((n, 4), dtype=np.uint32)
    lib.oracle_generate_seeds(as_ptr(out, C.c_uint), int(rseed), int(n))
    return out


def hybrid_taus_stream(state, count: int):
    st = np.array(state, dtype=np.uint32).copy()
    out = np.empty(count, dtype=np.uint32)
    for k in range(count):
        out[k] = lib.oracle_hybrid_taus(as_ptr(st, C.c_uint))
    return out, st


def rforce(state):
    st = np.array(state, dtype=np.uint32).copy()
    out = np.empty(4, dtype=np.float32)
    lib.oracle_rforce(as_ptr(st, C.c_uint), as_ptr(out, C.c_float))
    return out, st


class OracleState:
    """CPU mirror of one Engine: coordinates, forces, lists, RNG of a block of trajectories."""

    def __init__(self, system, traj_first: int = 0, n_tr_local=None, wrap_angles: bool = True, par=None):
        p = (par or system.par).copy()
        p.traj_first = traj_first
        p.n_tr_local = system.Ntr - traj_first if n_tr_local is None else n_tr_local
        self.par, self.system = p, system
        self.N, self.ntr = p.n_tot, p.n_tr_local
        n = self.N * self.ntr
        self.top = system.topology(traj_first)
        self.coords = np.array(system.coords[traj_first:traj_first + self.ntr], dtype=np.float32).copy()
        if wrap_angles:  # initIntegration, compute_cuda.cu:1004-1010
            for k in (3, 4, 5):
                a = self.coords[..., k].astype(np.float64)
                self.coords[..., k] = (a - (2 * np.pi) * np.trunc(a / (2 * np.pi))).astype(np.float32)
        self.forces = np.zeros((self.ntr, self.N, 7), dtype=np.float32)
        self.long_count = np.array(system.longitudinal_count[traj_first:traj_first + self.ntr], dtype=np.int32).copy()
        self.long = np.array(system.longitudinal[traj_first:traj_first + self.ntr], dtype=np.int32).copy()
        self.lat_count = np.array(system.lateral_count[traj_first:traj_first + self.ntr], dtype=np.int32).copy()
        self.lat = np.array(system.lateral[traj_first:traj_first + self.ntr], dtype=np.int32).copy()
        self.lj_count = np.zeros((self.ntr, self.N), dtype=np.int32)
        self.lj = np.zeros((self.ntr, self.N, LJ_CAPACITY), dtype=np.int32)
        seeds = generate_seeds(p.rseed, 2 * p.n_tot * p.n_tr)
        o = traj_first * self.N
        self.rng = np.stack([seeds[o:o + n], seeds[p.n_tot * p.n_tr + o:p.n_tot * p.n_tr + o + n]]).copy()
        self.tea_ci = np.zeros((n, 4), dtype=np.float32)
        self.tea_eps = np.zeros(n, dtype=np.float32)
        self.tea_beta = np.zeros(self.ntr, dtype=np.float32)

    def _lists(self) -> OracleLists:
        l = OracleLists()
        l.longitudinal_count = as_ptr(self.long_count, C.c_int)
        l.longitudinal = as_ptr(self.long, C.c_int)
        l.lateral_count = as_ptr(self.lat_count, C.c_int)
        l.lateral = as_ptr(self.lat, C.c_int)
        l.lj_count = as_ptr(self.lj_count, C.c_int)
        l.lj = as_ptr(self.lj, C.c_int)
        return l

    def _pc(self):
        return C.byref(self.par), C.byref(self.top)

    def rebuild_lj(self):
        l = self._lists()
        lib.oracle_lj_lists(*self._pc(), as_ptr(self.coords, C.c_float), C.byref(l))

    def rebuild_bonds(self):
        l = self._lists()
        lib.oracle_pair_lists(*self._pc(), as_ptr(self.coords, C.c_float), C.byref(l))

    def force(self):
        l = self._lists()
        lib.oracle_forces(*self._pc(), C.byref(l), as_ptr(self.coords, C.c_float), as_ptr(self.forces, C.c_float))
        return self.forces

    def energies(self):
        l = self._lists()
        e = np.zeros((self.ntr, self.N, 7), dtype=np.float64)
        lib.oracle_energies(*self._pc(), C.byref(l), as_ptr(self.coords, C.c_float), as_ptr(e, C.c_double))
        return e

    def integrate(self):
        lib.oracle_integrate(*self._pc(), as_ptr(self.coords, C.c_float), as_ptr(self.forces, C.c_float), as_ptr(self.rng, C.c_uint))

    def run(self, first_step: int, n_steps: int, skip_first_rebuild: bool = False):
        l = self._lists()
        lib.oracle_run(*self._pc(), C.byref(l), as_ptr(self.coords, C.c_float), as_ptr(self.forces, C.c_float),
                       as_ptr(self.rng, C.c_uint), int(first_step), int(n_steps), int(skip_first_rebuild))

    def tea_update(self):
        return lib.oracle_tea_update(*self._pc(), as_ptr(self.coords, C.c_float), as_ptr(self.tea_ci, C.c_float),
                                     as_ptr(self.tea_eps, C.c_float), as_ptr(self.tea_beta, C.c_float))

    def tea_integrate(self):
        lib.oracle_tea_integrate(*self._pc(), as_ptr(self.coords, C.c_float), as_ptr(self.forces, C.c_float),
                                 as_ptr(self.rng, C.c_uint), as_ptr(self.tea_ci, C.c_float), as_ptr(self.tea_beta, C.c_float))
