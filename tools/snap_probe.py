"""Why is the snapshot D2H slow inside the stride loop?  Engine-level reproduction with MADDY_GPU_PROFILE segments."""
import os, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ["MADDY_GPU_PROFILE"] = "1"
import numpy as np
from mt_b200 import Engine, HostSystem, workspace
d = Path(tempfile.mkdtemp())
workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=256)
with workspace.chdir(d):
    s = HostSystem("config.conf")
g = np.ones((256, s.Ntot), dtype=np.int32)
for mode in sys.argv[1:] or ["ahead", "synced", "ahead-noupload"]:
    print("mode", mode, flush=True)
    e = Engine(s)
    step = 0
    for stride in range(6):
        for w in range(10):
            if w == 0:
                if mode == "synced":
                    e.sync()
                e.snapshot_begin(coords=True, energies=True, rebuild=True)
            if mode != "ahead-noupload":
                e.upload_gtp(g)
            e.run(step, 100, skip_first_rebuild=(w == 0))
            if w == 0:
                e.snapshot_end()
            step += 100
    e.sync()
    e.close()
