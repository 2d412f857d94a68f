"""The C-ABI libraries load and export every symbol include/*.h declares (no compute calls: CPU-only)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

from mt_b200 import capi

ROOT = Path(__file__).resolve().parent.parent


def declared_functions(header: Path):
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b((?:maddy|mt)_[a-z0-9_]+)\s*\(", text)))


def test_kernel_library_exports_every_declared_symbol():
    names = declared_functions(ROOT / "include" / "maddy_b200.h")
    assert len(names) >= 25
    for n in names:
        assert hasattr(capi.lib, n), f"libmaddy_b200.so does not export {n}"
    assert sorted(capi.KERNEL_SYMBOLS) == names, "capi.KERNEL_SYMBOLS is out of sync with include/maddy_b200.h"


def test_host_library_exports_every_declared_symbol():
    names = [n for n in declared_functions(ROOT / "include" / "maddy_host.h")]
    for n in names:
        assert hasattr(capi.hostlib, n), f"libmaddy_host.so does not export {n}"
    assert sorted(capi.HOST_SYMBOLS) == names


def test_params_struct_layout_matches_header_order():
    # field order of the ctypes mirror == field order in the header
    text = (ROOT / "include" / "maddy_b200.h").read_text()
    body = text[text.index("typedef struct maddy_params {"):text.index("} maddy_params;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(?:typedef struct maddy_params \{\s*)?(int|float)\s+(.*)", decl, flags=re.S)
        if m:
            fields += [(m.group(1), f.strip()) for f in m.group(2).split(",")]
    mirror = [("int" if t is C.c_int else "float", n) for n, t in capi.MaddyParams._fields_]
    assert fields == mirror


def test_no_cpu_fallback_without_device(rundir, load_system):
    """Without a GPU every compute entry point must fail loudly (and creation must not succeed)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from mt_b200 import Engine, MaddyError
    s = load_system(rundir(runnum=1))
    with pytest.raises(MaddyError) as e:
        Engine(s)
    assert "no CPU path" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """mt_b200/ must never import, link or execute anything under oracle/."""
    for p in (ROOT / "mt_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h") and p.name != "build.py":
            txt = p.read_text()
            assert "pyoracle" not in txt and "maddy_oracle" not in txt and "libmaddy_oracle" not in txt, p
    import subprocess
    out = subprocess.run(["ldd", str(ROOT / "mt_b200" / "libmaddy_b200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out
