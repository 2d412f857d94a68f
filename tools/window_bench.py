"""Per-launch overhead of the fused window: K steps as windows of W steps, with and without an L2 flush in between."""
import sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
from mt_b200 import Engine, HostSystem, workspace
d = Path(tempfile.mkdtemp())
workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=256)
with workspace.chdir(d):
    s = HostSystem("config.conf")
stream = torch.cuda.Stream()
e = Engine(s, stream=stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
e.run(0, 200); e.sync()
for W, do_flush in ((1000, False), (100, False), (100, True), (20, False), (500, False), (5, False), (1, False)):
    K = 4000 if W >= 20 else 400
    evs = []
    step = 200
    with torch.cuda.stream(stream):
        for w in range(K // W):
            if do_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); e.run(step, W); b.record(stream)
            evs.append((a, b)); step += W
    stream.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in evs)
    print(f"window {W:5d} flush={do_flush}: {ms / K * 1e3:.2f} us/step  ({520 * 256 * K / ms / 1e6:.2f} G/s)")
