"""ctypes bindings of the two C-ABI libraries (include/maddy_b200.h, include/maddy_host.h).

Nothing here computes: every call goes into libmaddy_b200.so (sm_100a kernels) or
libmaddy_host.so (drop-in C++ host).  Import fails loudly if the libraries are not built.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent

COORD_STRIDE = 7
ENERGY_TERMS = 7
ZERO_SENTINEL = 999999
LJ_CAPACITY = 256
RUN_SKIP_FIRST_REBUILD = 1
SNAP_COORDS, SNAP_FORCES, SNAP_ENERGIES, SNAP_REBUILD = 1, 2, 4, 8
SNAP_ONTUBULE, SNAP_ONTUBULE_APPLY, SNAP_GTP, SNAP_ONTUBULE_GUARD = 16, 32, 64, 128
HYD_KEEP_SLOTS = 1
LIST_LONGITUDINAL, LIST_LATERAL, LIST_LJ = 0, 1, 2
LOAD_QUIET, LOAD_NO_FILES = 1, 2


class MaddyParams(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("n_tot", C.c_int), ("n_tr", C.c_int), ("traj_first", C.c_int),
        ("n_tr_local", C.c_int), ("device", C.c_int), ("rseed", C.c_int),
        ("dt", C.c_float), ("Temp", C.c_float),
        ("gammaR", C.c_float), ("gammaTheta", C.c_float), ("varR", C.c_float), ("varTheta", C.c_float),
        ("alpha", C.c_float), ("freeze_temp", C.c_float),
        ("C", C.c_float), ("B_psi", C.c_float), ("B_fi", C.c_float), ("B_theta", C.c_float),
        ("psi_0", C.c_float), ("fi_0", C.c_float), ("theta0_gtp", C.c_float), ("theta0_gdp", C.c_float),
        ("A_long", C.c_float), ("D_long", C.c_float), ("A_lat", C.c_float), ("D_lat", C.c_float), ("seam_coeff", C.c_float),
        ("barrier", C.c_int),
        ("a_barr_long", C.c_float), ("r_barr_long", C.c_float), ("w_barr_long", C.c_float),
        ("a_barr_lat", C.c_float), ("r_barr_lat", C.c_float), ("w_barr_lat", C.c_float),
        ("lj_on", C.c_int), ("ljscale", C.c_float), ("ljsigma6", C.c_float), ("ljpairscutoff", C.c_float),
        ("ljpairsupdatefreq", C.c_int),
        ("is_wall", C.c_int), ("rep_leftborder", C.c_float), ("rep_r", C.c_float), ("rep_eps", C.c_float), ("rep_h", C.c_float),
        ("is_assembly", C.c_int),
        ("tea_on", C.c_int), ("tea_a", C.c_float), ("tea_epsilon_freq", C.c_int), ("tea_capricious", C.c_int),
        ("tea_epsmax", C.c_float),
        ("max_harmonic", C.c_int), ("max_longitudinal", C.c_int), ("max_lateral", C.c_int),
    ]

    def copy(self) -> "MaddyParams":
        p = MaddyParams()
        C.memmove(C.byref(p), C.byref(self), C.sizeof(MaddyParams))
        return p


class MaddyTopology(C.Structure):
    _fields_ = [
        ("harmonic_count", C.POINTER(C.c_int)), ("harmonic", C.POINTER(C.c_int)),
        ("longitudinal_count", C.POINTER(C.c_int)), ("longitudinal", C.POINTER(C.c_int)),
        ("lateral_count", C.POINTER(C.c_int)), ("lateral", C.POINTER(C.c_int)),
        ("fixed", C.POINTER(C.c_ubyte)), ("extra", C.POINTER(C.c_ubyte)),
        ("mon_type", C.POINTER(C.c_int)), ("gtp", C.POINTER(C.c_int)), ("on_tubule_cur", C.POINTER(C.c_int)),
    ]


class HostParams(C.Structure):
    _fields_ = [
        ("steps", C.c_longlong), ("firststep", C.c_longlong), ("stride", C.c_longlong), ("hydrostep", C.c_longlong),
        ("fix", C.c_int), ("tub_length", C.c_int), ("out_energy", C.c_int), ("out_force", C.c_int),
        ("is_restart", C.c_int), ("is_const_conc", C.c_int), ("hydrolysis", C.c_int), ("n_gpus", C.c_int),
        ("conc", C.c_float), ("khydro", C.c_float), ("viscosity", C.c_float),
    ]


def _load(name: str) -> C.CDLL:
    path = PKG / name
    if name == "libmaddy_b200.so" and os.environ.get("MADDY_B200_LIB"):  # development: A/B builds of the kernels
        path = Path(os.environ["MADDY_B200_LIB"])
    if not path.exists():
        raise ImportError(f"{path} is not built; run `python -m mt_b200.build` (there is no Python/CPU fallback)")
    return C.CDLL(str(path), mode=C.RTLD_GLOBAL)


lib = _load("libmaddy_b200.so")
hostlib = _load("libmaddy_host.so")

_vp, _i, _ll, _u = C.c_void_p, C.c_int, C.c_longlong, C.c_uint
_pf, _pi, _pd, _pu8 = C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_ubyte)


def _sig(fn, res, args):
    fn.restype = res
    fn.argtypes = args


# every symbol of include/maddy_b200.h
KERNEL_SYMBOLS = [
    "maddy_create", "maddy_destroy", "maddy_last_error", "maddy_stream", "maddy_sync", "maddy_rebuild_lj",
    "maddy_rebuild_bonds", "maddy_force", "maddy_integrate", "maddy_tea_update", "maddy_tea_integrate", "maddy_run",
    "maddy_energies", "maddy_energies_device", "maddy_download_coords", "maddy_download_forces", "maddy_upload_coords",
    "maddy_upload_gtp", "maddy_upload_on_tubule", "maddy_upload_extra", "maddy_download_list", "maddy_upload_list",
    "maddy_download_rng", "maddy_upload_rng", "maddy_generate_seeds", "maddy_tea_beta", "maddy_ensemble_allreduce",
    "maddy_launch_count", "maddy_schedule_gtp", "maddy_rebuild_and_energies", "maddy_snapshot_begin", "maddy_snapshot_end",
    "maddy_list_stats", "maddy_analysis_setup", "maddy_analysis_reference", "maddy_analysis_temperature", "maddy_analysis_project",
    "maddy_analysis_protofilaments", "maddy_ensemble_stats_begin", "maddy_ensemble_stats_end", "maddy_download_tea",
    "maddy_snapshot_tubule_lengths", "maddy_snapshot_on_tubule", "maddy_insert_dimers", "maddy_has_exact_on_tubule",
    "maddy_hydrolysis_plan", "maddy_hydrolysis_result", "maddy_apply_scheduled_gtp", "maddy_rand_discard", "maddy_snapshot_gtp", "maddy_clear_guard",
    "maddy_hydrolysis_inputs", "maddy_hydrolysis_plan_sharded", "maddy_hydrolysis_plan_all", "maddy_on_tubule_rule",
]
HOST_SYMBOLS = [
    "mt_host_last_error", "mt_system_load", "mt_system_free", "mt_system_params", "mt_system_topology", "mt_system_coords",
    "mt_system_gtp", "mt_system_on_tubule", "mt_system_extra", "mt_system_energies", "mt_system_ensemble_stats", "mt_system_set_ngpus", "mt_system_srand",
    "mt_system_rand_window", "mt_system_rand_discard", "mt_system_rand_next",
    "mt_system_set_steps", "mt_system_compute", "mt_system_mt_length", "mt_system_hydrolyse", "mt_system_change_conc",
    "mt_system_save_pdb", "mt_dcd_read", "mt_pdb_count",
]

_sig(lib.maddy_create, _i, [C.POINTER(MaddyParams), C.POINTER(MaddyTopology), _pf, _vp, C.POINTER(_vp)])
_sig(lib.maddy_destroy, _i, [_vp])
_sig(lib.maddy_last_error, C.c_char_p, [_vp])
_sig(lib.maddy_stream, _vp, [_vp])
_sig(lib.maddy_sync, _i, [_vp])
for _n in ("maddy_rebuild_lj", "maddy_rebuild_bonds", "maddy_force", "maddy_integrate", "maddy_tea_integrate"):
    _sig(getattr(lib, _n), _i, [_vp])
_sig(lib.maddy_tea_update, _i, [_vp, _ll])
_sig(lib.maddy_run, _i, [_vp, _ll, _ll, _u])
_sig(lib.maddy_energies, _i, [_vp, _pd, _pd])
_sig(lib.maddy_energies_device, _vp, [_vp])
_sig(lib.maddy_rebuild_and_energies, _i, [_vp, _pd, _pd])
_sig(lib.maddy_download_coords, _i, [_vp, _pf])
_sig(lib.maddy_download_forces, _i, [_vp, _pf])
_sig(lib.maddy_upload_coords, _i, [_vp, _pf])
_sig(lib.maddy_upload_gtp, _i, [_vp, _pi])
_sig(lib.maddy_upload_on_tubule, _i, [_vp, _pi])
_sig(lib.maddy_upload_extra, _i, [_vp, _pu8])
_sig(lib.maddy_download_list, _i, [_vp, _i, _pi, _pi])
_sig(lib.maddy_upload_list, _i, [_vp, _i, _pi, _pi])
_sig(lib.maddy_download_rng, _i, [_vp, C.POINTER(C.c_uint)])
_sig(lib.maddy_upload_rng, _i, [_vp, C.POINTER(C.c_uint)])
_sig(lib.maddy_generate_seeds, None, [C.POINTER(C.c_uint), _i, _ll])
_sig(lib.maddy_tea_beta, _i, [C.c_double, _i, _i, C.c_float, C.c_float, _pf, _pd])
_sig(lib.maddy_ensemble_allreduce, _i, [C.POINTER(_vp), _i, C.POINTER(_pd), _i])
_sig(lib.maddy_launch_count, _ll, [_vp])
_sig(lib.maddy_download_tea, _i, [_vp, _pf, _pf, _pf])
_sig(lib.maddy_ensemble_stats_begin, _i, [C.POINTER(_vp), _i])
_sig(lib.maddy_ensemble_stats_end, _i, [C.POINTER(_vp), _i, _pd])
_sig(lib.maddy_schedule_gtp, _i, [_vp, _ll, _ll, _i, _pi])
_sig(lib.maddy_list_stats, _i, [_vp, C.POINTER(C.c_ulonglong), _i])
_sig(lib.maddy_analysis_setup, _i, [_vp, _pi, _pi, C.c_char_p, _i])
_sig(lib.maddy_analysis_reference, _i, [_vp])
_sig(lib.maddy_analysis_temperature, _i, [_vp, _pd])
_sig(lib.maddy_analysis_project, _i, [_vp, _pf])
_sig(lib.maddy_analysis_protofilaments, _i, [_vp, _pi])
_sig(lib.maddy_snapshot_begin, _i, [_vp, _u])
_sig(lib.maddy_snapshot_end, _i, [_vp, _pf, _pf, _pd])
_sig(lib.maddy_snapshot_tubule_lengths, _i, [_vp, _pi, _pi])
_sig(lib.maddy_snapshot_on_tubule, _i, [_vp, _pi, _pi])
_sig(lib.maddy_insert_dimers, _i, [_vp, _i, _pi, _pf])
_sig(lib.maddy_has_exact_on_tubule, _i, [_vp])
_sig(lib.maddy_on_tubule_rule, _i, [_pf, _pf, _pf])
_sig(lib.maddy_hydrolysis_plan, _i, [_vp, C.POINTER(C.c_uint), _ll, _ll, _i, _u])
_sig(lib.maddy_hydrolysis_result, _i, [_vp, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), _pi])
_sig(lib.maddy_apply_scheduled_gtp, _i, [_vp, _ll])
_sig(lib.maddy_rand_discard, None, [C.POINTER(C.c_uint), C.c_ulonglong])
_sig(lib.maddy_snapshot_gtp, _i, [_vp, _pi])
_sig(lib.maddy_clear_guard, _i, [_vp])
_sig(lib.maddy_hydrolysis_inputs, _i, [_vp, C.POINTER(_vp), C.POINTER(C.c_ulong)])
_sig(lib.maddy_hydrolysis_plan_sharded, _i, [_vp, _vp, _i, C.POINTER(C.c_uint), _ll, _ll, _i, _u])
_sig(lib.maddy_hydrolysis_plan_all, _i, [C.POINTER(_vp), _i, C.POINTER(C.c_uint), _ll, _ll, _i, _u])

_sig(hostlib.mt_host_last_error, C.c_char_p, [])
_sig(hostlib.mt_system_load, _i, [C.c_char_p, _i, C.POINTER(C.c_char_p), _u, C.POINTER(_vp)])
_sig(hostlib.mt_system_free, None, [_vp])
_sig(hostlib.mt_system_params, _i, [_vp, C.POINTER(MaddyParams), C.POINTER(HostParams)])
_sig(hostlib.mt_system_topology, _i, [_vp, C.POINTER(MaddyTopology)])
_sig(hostlib.mt_system_coords, _pf, [_vp])
_sig(hostlib.mt_system_gtp, _pi, [_vp])
_sig(hostlib.mt_system_on_tubule, _pi, [_vp, _i])
_sig(hostlib.mt_system_extra, _pu8, [_vp])
_sig(hostlib.mt_system_energies, _pd, [_vp])
_sig(hostlib.mt_system_ensemble_stats, _pd, [_vp])
_sig(hostlib.mt_system_set_ngpus, _i, [_vp, _i])
_sig(hostlib.mt_system_srand, _i, [_vp, _u])
_sig(hostlib.mt_system_rand_window, _i, [_vp, C.POINTER(C.c_uint)])
_sig(hostlib.mt_system_rand_discard, _i, [_vp, C.c_ulonglong])
_sig(hostlib.mt_system_rand_next, _i, [_vp])
_sig(hostlib.mt_system_set_steps, _i, [_vp, _ll])
_sig(hostlib.mt_system_compute, _i, [_vp, _i, _pd])
_sig(hostlib.mt_system_mt_length, _i, [_vp, _ll, _pi])
_sig(hostlib.mt_system_hydrolyse, _i, [_vp])
_sig(hostlib.mt_system_change_conc, _i, [_vp, _pi, _pi, _pi])
_sig(hostlib.mt_system_save_pdb, _i, [_vp, C.c_char_p, C.c_char_p])
_sig(hostlib.mt_dcd_read, _i, [C.c_char_p, _pi, _pi, _pf, _ll])
_sig(hostlib.mt_pdb_count, _i, [C.c_char_p])


class MaddyError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"[{code}] {msg}")
        self.code = code


def as_ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


def generate_seeds(rseed: int, np_count: int) -> np.ndarray:
    """HybridTaus seed table, uint32 [np_count, 4] (reference generateSeeds on a fresh ran2 state)."""
    out = np.empty((np_count, 4), dtype=np.uint32)
    lib.maddy_generate_seeds(as_ptr(out, C.c_uint), int(rseed), int(np_count))
    return out


def tea_beta(epsilon_sum: float, n_noextra: int, capricious: bool, tea_a: float, epsmax: float):
    beta, eps = C.c_float(), C.c_double()
    rc = lib.maddy_tea_beta(float(epsilon_sum), int(n_noextra), int(capricious), float(tea_a), float(epsmax),
                            C.byref(beta), C.byref(eps))
    return rc, beta.value, eps.value


def read_dcd(path) -> np.ndarray:
    """All frames of a DCD file as float32 [frames, atoms, 3]."""
    n, fr = C.c_int(), C.c_int()
    if hostlib.mt_dcd_read(str(path).encode(), C.byref(n), C.byref(fr), None, 0):
        raise MaddyError(1, hostlib.mt_host_last_error().decode())
    out = np.empty((fr.value, n.value, 3), dtype=np.float32)
    if fr.value:
        hostlib.mt_dcd_read(str(path).encode(), C.byref(n), C.byref(fr), as_ptr(out, C.c_float), out.size)
    return out
