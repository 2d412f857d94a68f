"""tests/golden/seeds.json from the reference's own generateSeeds (see oracle/gen_golden_seeds.cu).
Runs in the build container only (needs /root/reference)."""
import json
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent.parent
exe = ROOT / "oracle" / "_ref" / "gen_golden_seeds"
exe.parent.mkdir(parents=True, exist_ok=True)
subprocess.run(["nvcc", "-O2", "-arch=sm_100", "-w", "-I/root/reference/src", "-o", str(exe), str(ROOT / "oracle" / "gen_golden_seeds.cu"),
                "/root/reference/src/HybridTaus.cu"], check=True)
cases = [(1234567, 1040), (1234567, 2 * 520 * 256), (42, 2 * 1560 * 4), (7, 64), (2147483000, 1000)]
out = [json.loads(subprocess.run([str(exe), str(s), str(n)], check=True, capture_output=True, text=True).stdout) for s, n in cases]
(ROOT / "tests" / "golden" / "seeds.json").write_text(json.dumps(out, indent=1))
print("wrote", len(out), "cases")
