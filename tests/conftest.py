"""pytest configuration: the `gpu` marker, in-tree build of the native libraries, shared fixtures."""
import hashlib
import os
import sys
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"
REFERENCE = Path("/root/reference")  # exists only in the build container; never required


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 box with `-m gpu`)")
    # build (or reuse) the shared libraries before any test imports them
    from mt_b200 import build
    build.build_kernels()
    build.build_host()
    build.build_oracle()


@pytest.fixture(scope="session")
def tmp_root():
    d = Path(tempfile.mkdtemp(prefix="maddy_tests_"))
    yield d


@pytest.fixture()
def rundir(tmp_root, request):
    """factory: rundir(name, **config overrides) -> Path with config.conf, morse.conf, cond.conf, dcd/*.pdb"""
    from mt_b200 import workspace
    counter = {"n": 0}

    def make(name="mt40_single", structure=None, forcefield=None, conditions=None, **config):
        counter["n"] += 1
        tag = hashlib.sha1(request.node.nodeid.encode()).hexdigest()[:6]  # parametrized ids share their first 40 characters
        d = tmp_root / f"{request.node.name[:40]}_{tag}_{counter['n']}"
        if structure is None:
            return workspace.make_baseline_rundir(d, name, **config)
        spec = workspace.BASELINE_CONFIGS[name]
        cfg = dict(spec.get("config", {}))
        cfg.update(config)
        ff = dict(spec.get("forcefield") or {})
        ff.update(forcefield or {})
        cond = dict(spec.get("conditions") or {})
        cond.update(conditions or {})
        return workspace.make_rundir(d, structure, cfg, ff, cond)

    return make


@pytest.fixture()
def load_system():
    """factory: load_system(rundir, overrides) -> HostSystem (quiet, no output files)"""
    from mt_b200 import HostSystem, workspace
    made = []

    def load(d, overrides=(), **kw):
        with workspace.chdir(d):
            s = HostSystem("config.conf", list(overrides), **kw)
        made.append(s)
        return s

    yield load
    for s in made:
        s.close()


def golden_npz(name):
    import numpy as np
    p = GOLDEN / f"ref_{name}.npz"
    if not p.exists():
        pytest.skip(f"{p.name} not generated yet")
    return np.load(p)
