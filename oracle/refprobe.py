"""Reader / runner for oracle/_ref/ref_probe dumps (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import struct
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
REF_PROBE = ROOT / "oracle" / "_ref" / "ref_probe"
REF_MT = ROOT / "oracle" / "_ref" / "mt"

COORD = np.dtype([("x", "f4"), ("y", "f4"), ("z", "f4"), ("fi", "f4"), ("theta", "f4"), ("psi", "f4"), ("w", "f4")])


class ProbeDump:
    """records[tag] = list of (step, ndarray).  Shapes are resolved from the header."""

    def __init__(self, path):
        raw = Path(path).read_bytes()
        self.records = {}
        off = 0
        while off < len(raw):
            tag = raw[off:off + 8].split(b"\0")[0].decode()
            step, nbytes = struct.unpack_from("<qq", raw, off + 8)
            off += 24
            self.records.setdefault(tag, []).append((step, raw[off:off + nbytes]))
            off += nbytes
        hdr = np.frombuffer(self.records["header"][0][1], dtype=np.int64)
        self.N, self.Ntr, self.window, self.detail, self.maxH, self.capLong, self.capLat, self.sizeof_par = [int(v) for v in hdr]

    def steps(self, tag):
        return [s for s, _ in self.records.get(tag, [])]

    def _raw(self, tag, step):
        for s, b in self.records.get(tag, []):
            if s == step:
                return b
        raise KeyError((tag, step))

    def coords(self, step, tag="coords") -> np.ndarray:
        """[Ntr, N, 7] float32 (x,y,z,fi,theta,psi,w)"""
        return np.frombuffer(self._raw(tag, step), dtype=np.float32).reshape(self.Ntr, self.N, 7).copy()

    def forces(self, step):
        return self.coords(step, "forces")

    def energy(self, step):
        return np.frombuffer(self._raw("energy", step), dtype=np.float64).reshape(self.Ntr, self.N, 7).copy()

    def ints(self, tag, step, cap=None):
        a = np.frombuffer(self._raw(tag, step), dtype=np.int32)
        return a.reshape(self.Ntr, self.N, cap).copy() if cap else a.reshape(self.Ntr, self.N).copy()

    def lj(self, step):
        return self.ints("ljcnt", step), self.ints("lj", step, 256)

    def bonds(self, step):
        return (self.ints("longcnt", step), self.ints("long", step, self.capLong),
                self.ints("latcnt", step), self.ints("lat", step, self.capLat))

    def seeds(self, step):
        return np.frombuffer(self._raw("seeds", step), dtype=np.uint32).reshape(2, self.Ntr * self.N, 4).copy()

    def floats(self, tag, step, shape):
        return np.frombuffer(self._raw(tag, step), dtype=np.float32).reshape(shape).copy()


def run_probe(rundir, out, window, detail, overrides=(), timeout=600):
    cmd = [str(REF_PROBE), "config.conf", str(out), str(window), str(detail), *overrides]
    r = subprocess.run(cmd, cwd=str(rundir), capture_output=True, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_probe failed ({r.returncode}): {r.stdout[-2000:]} {r.stderr[-2000:]}")
    return ProbeDump(out)


def run_reference_mt(rundir, overrides=(), timeout=3600):
    """Run the unmodified reference binary in rundir; returns (wall seconds, stdout)."""
    import time
    t0 = time.perf_counter()
    r = subprocess.run([str(REF_MT), "config.conf", *overrides], cwd=str(rundir), capture_output=True, text=True, timeout=timeout)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"reference mt failed ({r.returncode}): {r.stdout[-2000:]} {r.stderr[-2000:]}")
    return dt, r.stdout
