"""mt_b200 — B200-native Langevin/BD step loop of MADDY (klyshko/MT) behind a C-ABI.

Only what the hot path needs lives here:
  csrc/     hand-written sm_100a CUDA kernels + the C-ABI (include/maddy_b200.h)
  host/     drop-in C++ host (config/PDB/DCD/updater/main) + its C-ABI (include/maddy_host.h)
  capi.py   ctypes bindings          api.py        HostSystem / Engine (reference vocabulary)
  structures.py  lattice generators  workspace.py  run directories of the BASELINE configs
  build.py  in-tree build of all native artefacts

The bindings load lazily so that `python -m mt_b200.build` works before the libraries exist; touching
Engine / HostSystem / capi without built libraries raises ImportError (there is no Python or CPU fallback).
"""
import importlib

_LAZY = {
    "MaddyError": "capi", "generate_seeds": "capi", "read_dcd": "capi",
    "Engine": "api", "HostSystem": "api", "pdb_labels": "api",
}


def __getattr__(name):
    if name in _LAZY:
        return getattr(importlib.import_module(f".{_LAZY[name]}", __name__), name)
    if name in ("capi", "api", "build", "structures", "workspace"):
        return importlib.import_module(f".{name}", __name__)
    raise AttributeError(name)
