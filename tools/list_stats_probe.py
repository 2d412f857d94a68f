import sys, tempfile
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mt_b200 import Engine, HostSystem, workspace
d = Path(tempfile.mkdtemp()); workspace.make_baseline_rundir(d, "mt40_ensemble", runnum=256)
with workspace.chdir(d): s = HostSystem("config.conf")
e = Engine(s); e.run(0, 1000); e.sync(); e.list_stats(reset=True)
for w in (100, 100, 100, 1000):
    e.run(1000, w) if w == 100 else e.run(2000, w); e.sync()
    print(w, e.list_stats(reset=True))
