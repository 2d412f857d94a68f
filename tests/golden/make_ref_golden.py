"""Generate the golden vectors of tests/golden/ref_*.npz from the REFERENCE'S OWN kernels.

Runs oracle/_ref/ref_probe (the unmodified reference translation units + the dump driver
oracle/ref_probe.cu, built in place from /root/reference/src) on a B200 and keeps a compact
subset of its dumps.  Must run on a GPU box:  gpurun -- python tests/golden/make_ref_golden.py
Outputs land in gpurun_out/golden/ and are copied to tests/golden/ by hand.
"""
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from mt_b200 import workspace  # noqa: E402
from oracle import refprobe  # noqa: E402

# name -> (baseline config, ntr, window, overrides)
CASES = {
    "mt40": ("mt40_single", 1, 40, []),
    "mt120_gdp_barrier": ("mt120_disassembly", 1, 20, ["probe_gdp_every=3", "probe_ontub=1"]),
    "reserve_walls": ("mt120_constconc", 1, 20, ["repulsive_walls=yes", "rep_r=7.0", "rep_h=100"]),
    "mt40_static": ("mt40_single", 1, 20, ["is_assembly=no"]),
    "tea": ("cylinder_tea", 1, 20, []),
}


def trim_lists(cnt, ent):
    width = max(1, int(cnt.max()))
    return ent[..., :width].astype(np.int32)


def main():
    out = ROOT / "gpurun_out" / "golden"
    out.mkdir(parents=True, exist_ok=True)
    for name, (cfg, ntr, window, over) in CASES.items():
        d = Path(tempfile.mkdtemp(prefix=f"golden_{name}_"))
        workspace.make_baseline_rundir(d, cfg, runnum=ntr, steps=window, stride=100000)
        dump = refprobe.run_probe(d, d / "probe.bin", window, 2, over + ["hydrolysis=no"])
        N = dump.N
        a = {"case": np.array(cfg), "overrides": np.array(" ".join(over + ["hydrolysis=no"])), "ntr": np.array(ntr), "window": np.array(window),
             "params_raw": np.frombuffer(dump._raw("params", -1), dtype=np.uint8),
             "harm": np.frombuffer(dump._raw("harm", -1), dtype=np.int32).reshape(N, dump.maxH),
             "harmcnt": np.frombuffer(dump._raw("harmcnt", -1), dtype=np.int32),
             "montype": np.frombuffer(dump._raw("montype", -1), dtype=np.int32),
             "fixed": np.frombuffer(dump._raw("fixed", -1), dtype=np.uint8),
             "extra": np.frombuffer(dump._raw("extra", -1), dtype=np.uint8).reshape(ntr, N),
             "gtp": np.frombuffer(dump._raw("gtp", -1), dtype=np.int32).reshape(ntr, N),
             "ontub": np.frombuffer(dump._raw("ontub", -1), dtype=np.int32).reshape(ntr, N),
             "seeds0": dump.seeds(-1), "seeds_end": dump.seeds(window),
             "coords0": dump.coords(0), "coords1": dump.coords(1), "coords_end": dump.coords(window),
             "forces0": dump.forces(0), "forces1": dump.forces(1)}
        if "energy" in dump.records:
            a["energy0"] = dump.energy(0)
        if "lj" in dump.records:
            c, e = dump.lj(0)
            a["ljcnt0"], a["lj0"] = c, trim_lists(c, e)
        lc, le, tc, te = dump.bonds(0)
        a["longcnt0"], a["long0"], a["latcnt0"], a["lat0"] = lc, trim_lists(lc, le), tc, trim_lists(tc, te)
        a["cap_long"], a["cap_lat"] = np.array(dump.capLong), np.array(dump.capLat)
        if "tea_ci" in dump.records:
            a["tea_ci0"] = dump.floats("tea_ci", 0, (ntr * N, 4))
            a["tea_eps0"] = dump.floats("tea_eps", 0, (ntr * N,))
            a["tea_beta0"] = dump.floats("tea_beta", 0, (ntr,))
        np.savez_compressed(out / f"ref_{name}.npz", **a)
        print(name, "N", N, "window", window, "->", (out / f"ref_{name}.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
