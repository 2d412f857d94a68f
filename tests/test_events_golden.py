"""The host callbacks of the stride block (mt_b200/host/events.cpp: mt_length, hydrolyse, change_conc) against what the
REFERENCE's own functions (updater.cpp:97-257, compiled from /root/reference/src into oracle/_ref/ref_events_probe) did on
the same adversarial inputs: tests/golden/ref_events.npz, made by tests/golden/make_events_golden.py.  Bit for bit: flags,
counts, GTP state after every event, inserted coordinates, reserve flags and the position of the rand() stream afterwards.
(The device-side versions of the same three are compared with these host functions in tests/test_gpu_events.py.)"""
from pathlib import Path

import numpy as np

from mt_b200 import workspace

GOLDEN = Path(__file__).parent / "golden" / "ref_events.npz"


def test_host_events_equal_the_reference_functions(tmp_path, load_system):
    g = np.load(GOLDEN)
    ntr = int(g["ntr"])
    spec = workspace.BASELINE_CONFIGS["mt120_constconc"]
    d = workspace.make_rundir(tmp_path / "run", ("reserve", int(g["structure"][0]), int(g["structure"][1])), dict(spec["config"], runnum=ntr),
                              dict(spec["forcefield"]), dict(spec["conditions"], conc=float(g["conc"])))
    s = load_system(d)
    N = s.Ntot
    assert g["coords"].shape == (ntr, N, 7)
    s.coords[...] = g["coords"]
    s.gtp[...] = g["gtp"]
    s.on_tubule_cur[...] = g["on_cur"]
    s.on_tubule_prev[...] = g["on_prev"]
    s.extra[...] = g["extra"].astype(np.uint8)
    s.srand(int(g["seed"]))
    # mt_length(): the classification with libm cosf / sqrt on every threshold
    mt_len = s.mt_length(1000)
    assert np.array_equal(np.array(s.on_tubule_cur), g["on_tubule"])
    assert np.array_equal(mt_len, g["mt_len"])
    assert 0.1 < g["on_tubule"].mean() < 0.9
    # hydrolyse(), event after event: same draws in the same order
    for k in range(g["gtp_after"].shape[0]):
        s.hydrolyse()
        assert np.array_equal(np.array(s.gtp), g["gtp_after"][k]), k
    # change_conc(): insertions until the concentration is reached or the reserve runs out
    delta = mt_len - g["len_prev"]
    flag = s.change_conc(delta, mt_len)
    assert flag == int(g["flag"][0]) and flag > 0
    assert np.array_equal(np.array(s.extra).astype(np.int32), g["extra_after"])
    assert np.array_equal(np.array(s.coords), g["coords_after"])
    # ... and the stream stands where the reference's libc rand() stands
    assert [s.rand_next() for _ in range(4)] == [int(v) for v in g["next_rand"]]
