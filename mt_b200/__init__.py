"""mt_b200 — B200-native Langevin/BD step loop of MADDY (klyshko/MT) behind a C-ABI.

Only what the hot path needs lives here:
  csrc/     hand-written sm_100a CUDA kernels + the C-ABI (include/maddy_b200.h)
  host/     drop-in C++ host (config/PDB/DCD/updater/main) + its C-ABI (include/maddy_host.h)
  capi.py   ctypes bindings          api.py        HostSystem / Engine (reference vocabulary)
  structures.py  lattice generators  workspace.py  run directories of the BASELINE configs
  build.py  in-tree build of all native artefacts
"""
from .capi import MaddyError, generate_seeds, read_dcd  # noqa: F401
from .api import Engine, HostSystem  # noqa: F401
