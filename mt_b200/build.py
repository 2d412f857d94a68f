"""In-tree build of every native artefact (no JIT cache: the .so files travel with the tree).

  mt_b200/libmaddy_b200.so   sm_100a kernels + the C-ABI of include/maddy_b200.h   (nvcc)
  mt_b200/libmaddy_host.so   drop-in C++ host + the C-ABI of include/maddy_host.h   (g++)
  mt_b200/mt                 drop-in `mt <config.conf>` executable                  (g++)
  oracle/_build/libmaddy_oracle.so   CPU restatement used by tests / smoke / cpu_baseline (gcc)
  oracle/_ref/{mt,ref_probe} the reference's own CUDA build, compiled IN PLACE from
                             /root/reference/src when that tree is present (nvcc)
  oracle/_ref/mt_stub        the reference's own HOST (main/preparator/updater/...) with compute() replaced by
                             oracle/compute_b200_stub.cpp, linked against libmaddy_b200.so
  oracle/_ref/ref_events_probe  the reference's host callbacks (updater.cpp: mt_length, hydrolyse, change_conc) behind a
                             dump driver (oracle/ref_events_probe.cu); host code, makes tests/golden/ref_events.npz
  oracle/_ref/{disc,p3d22d,temp_calc}  the reference's analysis tools (scripts/), same rule (g++)

Run:  python -m mt_b200.build [--force] [--no-ref]
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "mt_b200"
CSRC = PKG / "csrc"
HOST = PKG / "host"
ORACLE = ROOT / "oracle"
REFERENCE = Path("/root/reference")

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]

LIB_KERNELS = PKG / "libmaddy_b200.so"
LIB_HOST = PKG / "libmaddy_host.so"
MT_BIN = PKG / "mt"
LIB_ORACLE = ORACLE / "_build" / "libmaddy_oracle.so"
REF_MT = ORACLE / "_ref" / "mt"
REF_PROBE = ORACLE / "_ref" / "ref_probe"
REF_STUB = ORACLE / "_ref" / "mt_stub"


def _stale(target: Path, sources) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in sources)


def _run(cmd, **kw):
    cmd = [str(c) for c in cmd]
    print("+", " ".join(cmd), flush=True)
    subprocess.run(cmd, check=True, **kw)


def build_kernels(force=False, verbose_ptxas=False):
    srcs = [CSRC / "maddy_kernels.cu", CSRC / "maddy_tea.cu", CSRC / "maddy_analysis.cu", CSRC / "maddy_events.cu", CSRC / "maddy_capi.cu", CSRC / "maddy_seeds.cpp"]
    deps = srcs + list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "maddy_b200.h"]
    if not force and not _stale(LIB_KERNELS, deps):
        return LIB_KERNELS
    # -use_fast_math: the reference is built with it (CMakeLists.txt:62) and the MUFU lowering of
    # sinf/cosf/expf/logf/sqrtf and `/` is part of the arithmetic contract.
    # maddy_kernels.cu additionally gets -fmad=false: its device functions are shared by several kernels (fused
    # loop, step-granular phases) that must round identically, so every FMA there is an explicit fmaf().
    common = [*ARCH, "-lineinfo", "-O3", "-use_fast_math", "-std=c++17", "-Xcompiler", "-fPIC", f"-I{ROOT / 'include'}", f"-I{CSRC}"]
    if verbose_ptxas:
        common += ["-Xptxas", "-v"]
    objdir = PKG / "_build"
    objdir.mkdir(exist_ok=True)
    objs = []
    for src in srcs:
        obj = objdir / (src.stem + ".o")
        extra = ["-fmad=false"] if src.name == "maddy_kernels.cu" else []
        if src.name == "maddy_capi.cu":
            extra += ["-Xcompiler", "-fopenmp"]  # host-side flag conversions / read-back transposes
        if force or _stale(obj, deps):
            _run([NVCC, "-c", *common, *extra, "-o", obj, src])
        objs.append(obj)
    _run([NVCC, "-shared", *ARCH, "-o", LIB_KERNELS, *objs, "-ldl", "-lgomp"])
    return LIB_KERNELS


def build_host(force=False):
    srcs = [HOST / "io.cpp", HOST / "system.cpp", HOST / "events.cpp", HOST / "checkpoint.cpp", HOST / "host_capi.cpp"]
    deps = srcs + [HOST / "mt_host.hpp", HOST / "main.cpp", ROOT / "include" / "maddy_host.h", ROOT / "include" / "maddy_b200.h"]
    build_kernels(force)
    if force or _stale(LIB_HOST, deps + [LIB_KERNELS]):
        _run(["g++", "-O2", "-std=c++17", "-pthread", "-fopenmp", "-shared", "-fPIC", f"-I{ROOT / 'include'}", f"-I{HOST}", "-o", LIB_HOST, *srcs,
              f"-L{PKG}", "-lmaddy_b200", "-Wl,-rpath,$ORIGIN"])
    if force or _stale(MT_BIN, deps + [LIB_HOST]):
        _run(["g++", "-O2", "-std=c++17", "-pthread", "-fopenmp", f"-I{ROOT / 'include'}", f"-I{HOST}", "-o", MT_BIN, HOST / "main.cpp",
              f"-L{PKG}", "-lmaddy_host", "-lmaddy_b200", "-Wl,-rpath,$ORIGIN"])
    return LIB_HOST


def build_oracle(force=False):
    srcs = sorted(ORACLE.glob("*.c"))
    if not srcs:
        return None
    deps = srcs + sorted(ORACLE.glob("*.h"))
    if force or _stale(LIB_ORACLE, deps):
        LIB_ORACLE.parent.mkdir(parents=True, exist_ok=True)
        # -ffp-contract=off: the restatement states every rounding explicitly
        _run(["gcc", "-O2", "-std=gnu11", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", f"-I{ROOT / 'include'}", "-o", LIB_ORACLE,
              *srcs, "-lm"])
    return LIB_ORACLE


def build_reference(force=False):
    """Compile the reference's own `mt` and the probe harness from the sources where they lie.

    Recipe = the reference's CMake flags (CMakeLists.txt:11,62,77) as one direct nvcc line; its
    CMakeLists does not configure under CMake 4.  Nothing is copied: only binaries land in oracle/_ref.
    """
    src = REFERENCE / "src"
    if not src.is_dir():
        return None
    REF_MT.parent.mkdir(parents=True, exist_ok=True)
    common = ["dcdio.cpp", "xyzio.cpp", "globals.cpp", "pdbio.cpp", "configreader.cpp", "wrapper.cpp", "timer.cpp",
              "parameters.cpp", "compute_cuda.cu", "updater.cpp", "preparator.cpp", "bdhitea.cu", "bdhitea_kernel.cu",
              "HybridTaus.cu"]
    flags = ["-O2", "-arch=sm_100", "-rdc=true", "-use_fast_math", "-DCUDA", "-DMORSE", "-w", f"-I{src}"]
    if force or not REF_MT.exists():
        _run([NVCC, *flags, "-o", REF_MT, *[src / f for f in common], src / "main.cpp"])
    # the reference's offline analysis tools (plain g++, their Makefile's recipe): checkers for the in-situ analysis
    for out, main, d in (("disc", "disc.cpp", "disas_speed"), ("p3d22d", "3d22d.cpp", "disas_speed"), ("temp_calc", "main.cpp", "temp_calc")):
        sdir = REFERENCE / "scripts" / d
        if sdir.is_dir() and (force or not (REF_MT.parent / out).exists()):
            _run(["g++", "-O3", "-w", "-o", REF_MT.parent / out, sdir / main, sdir / "dcdio.cpp", sdir / "pdbio.cpp"])
    # the reference's own host with compute() swapped for the C-ABI binding: every reference
    # source except its four CUDA units, plus oracle/compute_b200_stub.cpp, linked against libmaddy_b200.so
    stub_src = ORACLE / "compute_b200_stub.cpp"
    if stub_src.exists() and (force or _stale(REF_STUB, [stub_src, LIB_KERNELS, ROOT / "include" / "maddy_b200.h"])):
        build_kernels()
        host_only = [f for f in common if not f.endswith(".cu")]
        _run([NVCC, "-O2", "-DCUDA", "-DMORSE", "-w", f"-I{src}", f"-I{ROOT / 'include'}", "-o", REF_STUB, *[src / f for f in host_only],
              src / "main.cpp", stub_src, f"-L{PKG}", "-lmaddy_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../mt_b200"])
    # the reference's host callbacks of the stride block (updater.cpp) behind a dump driver: runs without a GPU, makes
    # tests/golden/ref_events.npz (tests/golden/make_events_golden.py)
    events_src = ORACLE / "ref_events_probe.cu"
    events_bin = REF_MT.parent / "ref_events_probe"
    if events_src.exists() and (force or _stale(events_bin, [events_src])):
        _run([NVCC, "-O2", "-arch=sm_100", "-rdc=true", "-DCUDA", "-DMORSE", "-w", f"-I{src}", "-o", events_bin, src / "updater.cpp",
              src / "globals.cpp", events_src])
    probe_src = ORACLE / "ref_probe.cu"
    if probe_src.exists() and (force or _stale(REF_PROBE, [probe_src])):
        _run([NVCC, *flags, "-o", REF_PROBE, *[src / f for f in common], probe_src])
    return REF_MT


def build_all(force=False, with_reference=True):
    build_kernels(force)
    build_host(force)
    build_oracle(force)
    if with_reference:
        build_reference(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, with_reference="--no-ref" not in sys.argv)
