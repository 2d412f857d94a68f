"""Writes tests/golden/analysis_golden.json: outputs of the REFERENCE'S OWN analysis tools (scripts/temp_calc,
scripts/disas_speed, built in place into oracle/_ref by mt_b200/build.py) over synthetic disassembly frames of the
13 x 40 lattice.  Run where /root/reference exists:  python tests/golden/make_analysis_golden.py"""
import json
import sys
import tempfile
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
from helpers import run_reference_analysis, synthetic_disassembly_frames  # noqa: E402
from mt_b200 import build, pdb_labels, structures  # noqa: E402

build.build_reference()
d = Path(tempfile.mkdtemp(prefix="an_golden_"))
xyz, ang = structures.lattice(40, 3)
structures.write_pair(xyz, ang, d / "xyz.pdb", d / "ang.pdb")
chain, resid, name1 = pdb_labels(d / "xyz.pdb")
x0 = np.array([[a.x, a.y, a.z] for a in xyz], dtype=np.float32)
a0 = np.array([[a.x, a.y, a.z] for a in ang], dtype=np.float32)
fx, fa = synthetic_disassembly_frames(x0, a0, chain, resid, n_frames=5, seed=3)
out = run_reference_analysis(ROOT / "oracle" / "_ref", d, d / "xyz.pdb", fx, fa, 1000)
gold = {"structure": ["lattice", 40, 3], "n_frames": 5, "seed": 3, "stride": 1000,
        "temp": out["temp"].tolist(), "timeline": out["timeline"].tolist(), "summary": list(out["summary"]),
        "proj_sha256": __import__("hashlib").sha256(np.ascontiguousarray(out["proj"]).tobytes()).hexdigest()}
(ROOT / "tests" / "golden" / "analysis_golden.json").write_text(json.dumps(gold, indent=1))
print(json.dumps(gold)[:600])
