"""CPU restatement (numpy) of the reference's offline DCD analysis tools — TEST INFRASTRUCTURE ONLY: imported by
tests/, never by the product path (the in-situ analysis runs on the GPU, mt_b200/csrc/maddy_analysis.cu).

Follows, frame by frame:
  scripts/temp_calc/main.cpp:70-92         displacement sums between consecutive frames
  scripts/disas_speed/3d22d.cpp:42-51      projection {sqrtf(x*x + y*y), z, theta}
  scripts/disas_speed/disc.cpp:62-117      per-protofilament break / curl / tip numbers, :163-166 timeline value

Pinned against the tools themselves (oracle/_ref/{temp_calc,p3d22d,disc}, built in place from the reference) and
the committed outputs of those tools in tests/golden/analysis_golden.json (tests/golden/make_analysis_golden.py).
"""
import numpy as np

THRES, HOR_THRES, THETA_THRES, PF_NUMBER = 5.0, 2.0, 0.2, 13  # disc.cpp:9-12


def temperature_sums(xyz_new, ang_new, xyz_old, ang_old):
    """float32 [N,3] frames (ang = fi, psi, theta: the angular DCD's X, Y, Z) -> the 8 raw sums of main.cpp:70-92,
    in the tool's order of evaluation: float differences, double squares, sequential double sums."""
    d = (xyz_new.astype(np.float32) - xyz_old.astype(np.float32)).astype(np.float64)
    a = (ang_new.astype(np.float32) - ang_old.astype(np.float32)).astype(np.float64)
    c2 = np.cos(ang_new[:, 1].astype(np.float64)) ** 2
    t_xyz = (d[:, 0] ** 2 + d[:, 1] ** 2) + d[:, 2] ** 2
    t_rot = ((a[:, 0] ** 2 + a[:, 1] ** 2) + a[:, 2] ** 2) - ((2 * a[:, 0]) * a[:, 2]) * c2
    cols = [t_xyz, t_rot, d[:, 0] ** 2, d[:, 1] ** 2, d[:, 2] ** 2, a[:, 0] ** 2, a[:, 1] ** 2, a[:, 2] ** 2]
    return np.array([np.cumsum(c)[-1] for c in cols])  # cumsum = the tool's sequential accumulation


def temperature_scale(sums, stride, n=1560, gamma_r=1.06e+06, gamma_t=5e+06, k=0.002, dt=200):
    """the tool's printed columns from the raw sums (main.cpp:94-106; its constants are hard-coded, float products)"""
    f32 = np.float32
    six = f32(6 * stride * dt * n) * f32(k)
    two = f32(2 * stride * dt * n) * f32(k)
    gr, gt = f32(gamma_r), f32(gamma_t)
    s = np.asarray(sums, dtype=np.float64)
    return np.array([s[0] * float(gr / six), s[1] * float(gt / six), s[2] * float(gr / two), s[3] * float(gr / two),
                     s[4] * float(gr / two), s[5] * float(gt / two), s[6] * float(gt / two), s[7] * float(gt / two)])


def project(xyz, ang):
    """3d22d.cpp:42-51 in float32, every operation rounded to float"""
    x, y = xyz[:, 0].astype(np.float32), xyz[:, 1].astype(np.float32)
    r = np.sqrt((x * x + y * y).astype(np.float32)).astype(np.float32)
    return np.stack([r, xyz[:, 2].astype(np.float32), ang[:, 2].astype(np.float32)], axis=1)


def protofilaments(proj, chain, resid, name1, n_pf=PF_NUMBER):
    """disc.cpp:62-124 for one frame of the projection {X = radius, Y = z, Z = theta} -> int [n_pf, 3] =
    pf_end_number, curled_start (clamped, :119-124), mt_end_number.  Plain loops: small systems only."""
    n = len(chain)
    X, Y, Z = proj[:, 0], proj[:, 1], proj[:, 2]
    chain_len = np.zeros(n_pf, dtype=np.int64)
    pf_end = np.full(n_pf, n, dtype=np.int64)
    curled = np.full(n_pf, n, dtype=np.int64)
    mt_end_y = np.zeros(n_pf, dtype=np.float32)
    mt_end_number = np.zeros(n_pf, dtype=np.int64)
    for i in range(n):
        if 0 <= chain[i] < n_pf:
            chain_len[chain[i]] += 1
    for i in range(n):  # the tool's double loop admits only |id_i - id_j| == 1: consecutive records
        for j in (i - 1, i + 1):
            if j < 0 or j >= n or chain[i] != chain[j] or not (0 <= chain[i] < n_pf):
                continue
            if abs(int(resid[i]) - int(resid[j])) != 1 or name1[i] == name1[j]:
                continue
            dx = np.float64(np.float32(X[i]) - np.float32(X[j]))
            dy = np.float64(np.float32(Y[i]) - np.float32(Y[j]))
            if np.sqrt(dx * dx + dy * dy) > THRES and pf_end[chain[i]] > min(resid[i], resid[j]):
                pf_end[chain[i]] = min(resid[i], resid[j])
    for c in range(n_pf):
        if pf_end[c] > chain_len[c] // 2:
            pf_end[c] = chain_len[c] // 2
    for i in range(n):
        c = chain[i]
        if not (0 <= c < n_pf):
            continue
        if resid[i] <= pf_end[c] and Y[i] > mt_end_y[c] and np.float64(Z[i]) < HOR_THRES:
            mt_end_y[c] = Y[i]
            mt_end_number[c] = resid[i]
        if np.float64(Z[i]) > THETA_THRES and resid[i] < pf_end[c] and resid[i] < curled[c]:
            curled[c] = resid[i]
    curled = np.minimum(curled, pf_end)
    return np.stack([pf_end, curled, mt_end_number], axis=1).astype(np.int32)


def timeline_value(pf):
    """disc.cpp:163-166: 2 * sum(float(mt_end_number) / 13.0) accumulated in float"""
    lt = np.float32(0)
    for v in pf[:, 2]:
        lt = np.float32(np.float64(lt) + np.float64(np.float32(v)) / 13.0)  # float += double quotient
    return float(np.float32(2) * lt)
