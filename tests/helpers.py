"""Shared helpers of the test-suite."""
import ctypes as C

import numpy as np


class RefParameters(C.Structure):
    """Memory layout of the reference's `Parameters` (src/parameters.h:246-318, x86-64) — used to read the
    `params_raw` record of the ref_probe golden files and compare the host's derived constants."""
    _fields_ = [
        ("hdi_on", C.c_bool), ("rseed", C.c_int), ("Temp", C.c_float), ("varR", C.c_float), ("gammaR", C.c_float),
        ("varTheta", C.c_float), ("gammaTheta", C.c_float), ("viscosity", C.c_float), ("freeze_temp", C.c_float),
        ("is_assembly", C.c_bool), ("is_const_conc", C.c_bool), ("out_energy", C.c_bool), ("out_force", C.c_bool),
        ("tub_length", C.c_bool), ("conc", C.c_float), ("alpha", C.c_float), ("dt", C.c_float), ("device", C.c_int),
        ("steps", C.c_longlong), ("firststep", C.c_longlong), ("stride", C.c_longlong),
        ("Ntot", C.c_int), ("Ntr", C.c_int), ("firstrun", C.c_int),
        ("C", C.c_float), ("B_psi", C.c_float), ("B_fi", C.c_float), ("B_theta", C.c_float), ("psi_0", C.c_float),
        ("fi_0", C.c_float), ("theta0_gtp", C.c_float), ("theta0_gdp", C.c_float), ("A_lat", C.c_float), ("A_long", C.c_float),
        ("D_lat", C.c_float), ("D_long", C.c_float), ("seam_coeff", C.c_float),
        ("hydrolysis", C.c_bool), ("khydro", C.c_float), ("hydrostep", C.c_long),
        ("is_wall", C.c_bool), ("rep_h", C.c_float), ("rep_r", C.c_float), ("rep_eps", C.c_float), ("zs", C.c_float * 100),
        ("rep_leftborder", C.c_float),
        ("barrier", C.c_bool), ("a_barr_long", C.c_float), ("r_barr_long", C.c_float), ("w_barr_long", C.c_float),
        ("a_barr_lat", C.c_float), ("r_barr_lat", C.c_float), ("w_barr_lat", C.c_float),
        ("ljpairscutoff", C.c_float), ("ljpairsupdatefreq", C.c_int), ("lj_on", C.c_bool), ("ljscale", C.c_float),
        ("ljsigma6", C.c_float),
    ]


def ref_parameters(raw: np.ndarray) -> RefParameters:
    p = RefParameters()
    C.memmove(C.byref(p), raw.tobytes(), C.sizeof(RefParameters))
    return p


def lists_equal(cnt_a, ent_a, cnt_b, ent_b):
    """compare neighbour lists entry by entry up to their counts"""
    if not np.array_equal(cnt_a, cnt_b):
        return False
    width = min(ent_a.shape[-1], ent_b.shape[-1])
    assert cnt_a.max(initial=0) <= width
    mask = np.arange(width)[None, None, :] < cnt_a[..., None]
    return bool(np.array_equal(np.where(mask, ent_a[..., :width], 0), np.where(mask, ent_b[..., :width], 0)))


def system_from_golden(g, rundir, load_system, ntr=None, extra_overrides=()):
    """rebuild the run directory a golden file was generated from and load it with this repo's host"""
    over = [o for o in str(g["overrides"]).split() if not o.startswith("probe_")] + list(extra_overrides)
    d = rundir(str(g["case"]), runnum=int(g["ntr"]) if ntr is None else ntr, steps=int(g["window"]), stride=100000)
    return load_system(d, over)
