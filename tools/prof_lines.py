"""Attribute an ncu SASS profile to source lines: joins `ncu --page source --print-source sass --csv` with the line table
of the cubin (`nvdisasm -gi`).  usage: prof_lines.py <report.ncu-rep> <object.o> <mangled kernel name> <warps*steps> [min]"""
import collections, csv, io, re, subprocess, sys, tempfile, os
rep, obj, fun, per = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 8.0
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(f".text.{fun}:"))
ins = []  # ((innermost file, line), outermost line, chain)
cur = (("", 0), 0, ())
pat = re.compile(r'File "([^"]+)", line (\d+)')
fresh = True  # the next comment starts a new annotation block (innermost frame first, outermost last)
for l in dis[start + 1:]:
    if l.startswith("\t.section") or l.startswith(".text."):
        break
    if "//##" in l:
        m = pat.findall(l)
        if m:
            f0, l0 = os.path.basename(m[0][0]), int(m[0][1])
            last = int(m[-1][1])
            if fresh:
                cur = ((f0, l0), last, ((f0, l0),))
                fresh = False
            else:
                cur = (cur[0], last, cur[2] + ((f0, l0),))
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        ins.append(cur)
        fresh = True
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
hdr, data = rows[1], rows[2:]
ci, csamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(f"instructions in line table {len(ins)}, in profile {len(data)}")
n = min(len(ins), len(data))
inner, outer = collections.Counter(), collections.Counter()
inner_s, outer_s = collections.Counter(), collections.Counter()
mid = collections.Counter(); mid_s = collections.Counter()
for (a, b, chain), r in zip(ins[:n], data[:n]):
    # second-outermost frame = the line inside the function called from the kernel body
    key = chain[-2] if len(chain) >= 2 else (chain[-1] if chain else a)
    mid[key] += int(r[ci]); mid_s[key] += int(r[csamp])
    inner[a] += int(r[ci]); outer[b] += int(r[ci])
    inner_s[a] += int(r[csamp]); outer_s[b] += int(r[csamp])
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mt_b200", "csrc", "maddy_kernels.cu")).read().split("\n")
tot_s = sum(inner_s.values())
print("== by call site in the kernel body (outermost line), instructions per warp-step / % samples")
for line, c in sorted(outer.items()):
    if c / per >= thr:
        print(f"  {line:5d} {c / per:7.1f} {100 * outer_s[line] / tot_s:5.1f}%  {src[line - 1].strip()[:100] if 0 < line <= len(src) else ''}")
def text(f, line):
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "mt_b200", "csrc", f)
    try:
        return open(path).read().split("\n")[line - 1].strip()[:90]
    except Exception:
        return ""
print("== by line of the function called from the kernel body")
for (f, line), c in sorted(mid.items()):
    if c / per >= thr:
        print(f"  {f}:{line:5d} {c / per:7.1f} {100 * mid_s[(f, line)] / tot_s:5.1f}%  {text(f, line)}")
print("== by innermost line")
for (f, line), c in sorted(inner.items()):
    if c / per >= thr:
        print(f"  {f}:{line:5d} {c / per:7.1f} {100 * inner_s[(f, line)] / tot_s:5.1f}%  {text(f, line)}")
