#!/usr/bin/env python
"""Static SASS opcode histogram per kernel of the built library (no GPU needed):
    python tools/sass_histogram.py > profiles/r2_sass_histogram.txt
Counts instructions of `cuobjdump -sass mt_b200/libmaddy_b200.so` by opcode stem; the tensor-core / TMA columns are
there to show their absence."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "mt_b200" / "libmaddy_b200.so"
COLS = ["FFMA2", "FMUL2", "FADD2", "FFMA", "FMUL", "FADD", "MUFU", "DFMA", "DMUL", "DADD", "F2F", "LDS", "STS", "LDG", "STG", "LDGSTS",
        "UBLKCP", "UTMALDG", "UTCHMMA", "HMMA", "LDL", "STL", "BAR", "SHFL", "VOTE", "ATOM", "ATOMG", "RED", "IMAD", "LOP3", "SHF"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", out)))
    kernels, cur, k = collections.OrderedDict(), None, 0
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = demangle[k] if k < len(demangle) and demangle[k] else m.group(1)
            k += 1
            name = re.sub(r"\(.*", "", name.replace("(int)", ""))
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
            cur["__total"] += 1
    print(f"# SASS opcode histogram per kernel (static instruction counts, `cuobjdump -sass mt_b200/libmaddy_b200.so`); code objects: {', '.join(archs)}")
    print("# tensor-core / TMA opcodes (UTCHMMA, HMMA, UTMALDG, UBLKCP) are absent by design: the path is gather / scan / generated 3x3 mat-vec work (DESIGN.md 3)")
    print(f"{'kernel':<42}{'total':>8}" + "".join(f"{c:>7}" for c in COLS))
    for name, c in sorted(kernels.items(), key=lambda kv: -kv[1]["__total"]):
        print(f"{name[:41]:<42}{c['__total']:>8}" + "".join(f"{c[col]:>7}" for col in COLS))


if __name__ == "__main__":
    sys.exit(main())
