/*
 * maddy_kernels.cu — sm_100a kernels of the MADDY Langevin/BD step loop.
 *
 * Execution model (B200-first, not the reference's one-thread-per-(traj,monomer) 32-thread
 * blocks with every neighbour gathered from global memory):
 *
 *   one CTA = one trajectory.  Each thread owns MPT monomers whose state (6 coordinates +
 *   two HybridTaus streams) lives in REGISTERS for the whole launch.  At the top of every
 *   step each thread publishes, for its monomers, {position, angles, the three site-offset
 *   vectors r_mon*e3, R*p1, R*p2, flags} into a shared-memory stage (4 x float4 = 64 B per
 *   monomer, double-buffered => ONE __syncthreads per step).  Forces, the list rebuilds and
 *   the energies read neighbours from that stage only; HBM is touched for the neighbour
 *   lists (uint16, k-major => coalesced) and, at launch start/end, for the state.
 *
 *   OP_RUN iterates the steps between two host events inside the kernel
 *   (replaces >= 2 launches + 2 cudaDeviceSynchronize per step, compute_cuda.cu:1228-1238).
 *   The step-granular entry points run the same device functions for a single phase, so
 *   fused and unfused execution are bit-identical.
 *
 * Reference semantics restated here (file:line in /root/reference/src):
 *   forces      compute_cuda.cu:32-525     lists   :527-674 (bonds), :913-940 (LJ)
 *   energies    compute_cuda.cu:676-911    integrator :943-975, HybridTaus.cu:53-98
 */
#include "maddy_kernels.cuh"

#ifndef MADDY_OPT_LJ
#define MADDY_OPT_LJ 1
#endif

namespace maddy {

#define KB_BOLTZ 0.0019872041f // kcal/(mol*K), mt.h:39

// Dynamic shared memory of both trajectory kernels (declared once so that every access below is a plain shared-space
// load/store with a 32-bit address and folded immediate offsets - no generic pointers, no per-step pointer set-up).
extern __shared__ float4 g_smem[];

// One stage buffer: four float4 arrays of N entries (array-major: a warp gathering the positions of 32 different
// neighbours then spreads over all banks; a monomer-major layout was measured 4x worse in bank conflicts).
//   P: x, y, z, fi     E: r_mon*e3, psi     L1: R*p1, theta     L2: R*p2, flag bits (mon_type | gtp<<8 | ontub<<9 | extra<<10)
struct Stage {
    static constexpr bool kGlobal = false;
    int oP, oE, oL1, oL2; // float4 indices into g_smem
    __device__ __forceinline__ float4 &P(int j) const { return g_smem[oP + j]; }
    __device__ __forceinline__ float4 &E(int j) const { return g_smem[oE + j]; }
    __device__ __forceinline__ float4 &L1(int j) const { return g_smem[oL1 + j]; }
    __device__ __forceinline__ float4 &L2(int j) const { return g_smem[oL2 + j]; }
};

__device__ __forceinline__ Stage stage_at(int N, int buf)
{
    Stage s;
    s.oP = buf * 4 * N;
    s.oE = s.oP + N;
    s.oL1 = s.oE + N;
    s.oL2 = s.oL1 + N;
    return s;
}

struct Near { // shared-memory near list of the fused loop (see MD_NEAR_R2 in maddy_kernels.cuh)
    uint16_t *list; // [cap][N], k-major
    uint8_t *cnt;   // [N]
    float4 *tlo, *thi; // bounding boxes of tiles of MD_TILE consecutive monomers
    int cap, ntiles;
    int stride;     // row stride of `list` for stages in global memory (the wide path's per-CTA list); N otherwise
    bool ok;        // CTA-uniform: the near list is valid for this step
    bool stale_lj;  // CTA-uniform: the Verlet list in HBM predates the last list-update step (lazy fused loop): a monomer
                    // whose near list overflowed walks its near-candidates instead (same pairs inside the force cut-off,
                    // same order: every such pair is certainly listed, see KArgs::lazy)
    uint4 *topo;    // [N] packed topology words (see load_topo), or nullptr: read the lists from HBM
};

// Per-monomer topology words the force evaluation needs every step, packed so that one LDS.128 from the
// thread's own shared-memory slot replaces six dependent global loads (they change only at list rebuilds):
//   x = first harmonic entry (raw, signed)        y = harmonic count | longitudinal count << 8 | lateral count << 16
//   z = longitudinal codes 0,1 (16 bit each)      w = lateral codes 0,1
// Entries beyond these (rare) are read from the lists in HBM.
__device__ __forceinline__ uint4 load_topo(const DevSys &a, int traj, int i)
{
    const uint16_t *bl = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
    const uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
    const unsigned nh = (unsigned)a.harm_count[i], nlong = bc[0], nlat = bc[a.Npad];
    uint4 t;
    t.x = (unsigned)a.harm[a.maxH * i];
    t.y = (nh & 0xffu) | (nlong << 8) | (nlat << 16);
    t.z = (nlong > 0 ? (unsigned)bl[0] : 0u) | (nlong > 1 ? (unsigned)bl[a.Npad] << 16 : 0u);
    t.w = (nlat > 0 ? (unsigned)bl[(size_t)a.capLong * a.Npad] : 0u) | (nlat > 1 ? (unsigned)bl[(size_t)(a.capLong + 1) * a.Npad] << 16 : 0u);
    return t;
}
__device__ __forceinline__ int topo_harm(const DevSys &a, const uint4 &t, int i, int k) { return k == 0 ? (int)t.x : a.harm[a.maxH * i + k]; }
__device__ __forceinline__ unsigned topo_long(const DevSys &a, const uint4 &t, int traj, int i, int k)
{
    if (k < 2) return (t.z >> (16 * k)) & 0xffffu;
    return a.bl[((size_t)traj * (a.capLong + a.capLat) + k) * a.Npad + i];
}
__device__ __forceinline__ unsigned topo_lat(const DevSys &a, const uint4 &t, int traj, int i, int k)
{
    if (k < 2) return (t.w >> (16 * k)) & 0xffffu;
    return a.bl[((size_t)traj * (a.capLong + a.capLat) + a.capLong + k) * a.Npad + i];
}

struct Mono { // register-resident state of one monomer
    float x, y, z, fi, psi, theta;
    uint4 rx, ra;
    int flags; // bit0 fixed, bits1-7 type, bit8 gtp, bit9 ontub, bit10 extra
};
#define MF_FIXED 1
#define MF_TYPE(f) (((f) >> 1) & 0x7f)
#define MF_GTP 0x100
#define MF_ONTUB 0x200
#define MF_EXTRA 0x400

__device__ __forceinline__ Frame publish(const Stage &s, int i, const Mono &m, const LatSite &ls)
{
    F3 a, l1, l2;
    const Frame fr = make_frame(m.fi, m.psi, m.theta, ls, a, l1, l2);
    s.P(i) = make_float4(m.x, m.y, m.z, m.fi);
    s.E(i) = make_float4(a.x, a.y, a.z, m.psi);
    s.L1(i) = make_float4(l1.x, l1.y, l1.z, m.theta);
    int jf = MF_TYPE(m.flags) | (m.flags & (MF_GTP | MF_ONTUB | MF_EXTRA));
    s.L2(i) = make_float4(l2.x, l2.y, l2.z, __int_as_float(jf));
    return fr;
}

// ------------------------------------------------------------------ forces
// Generalized force on monomer i (non-extra) from the staged trajectory.
// kBatchWalk: walk the full Verlet list in batches of 16 (index loads first, then positions): for callers whose list
// walk is a chain of L2 round trips (step-granular phase kernel, wide path).  Same arithmetic in the same order.
template <bool kBatchWalk = false, class S>
__device__ __forceinline__ G6 monomer_force(const KArgs &k, const S &s, const Near &near, int traj, int i, const Mono &m,
                                            const Frame &fr)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    G6 f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const float4 Ei = s.E(i);
    const float xi = m.x, yi = m.y, zi = m.z;
    const bool gtp_i = (m.flags & MF_GTP) != 0;
    const bool ontub_i = (m.flags & MF_ONTUB) != 0;

    const uint4 tw = near.topo ? near.topo[i] : load_topo(a, traj, i);

    // ---- harmonic intra-dimer bond + bending (compute_cuda.cu:70-182)
    const int nh = tw.y & 0xff;
    for (int kk = 0; kk < nh; kk++) {
        const int raw = topo_harm(a, tw, i, kk);
        const float sg = raw < 0 ? 1.0f : -1.0f; // R_MON / r_mon
        const int j = raw < 0 ? -raw : raw;
        const float4 Pj = s.P(j), Ej = s.E(j);
        F3 d;
        d.x = fmaf(-sg, Ej.x, fmaf(-sg, Ei.x, Pj.x - xi));
        d.y = fmaf(-sg, Ej.y, fmaf(-sg, Ei.y, Pj.y - yi));
        d.z = fmaf(-sg, Ej.z, fmaf(-sg, Ei.z, Pj.z - zi));
        const float dr = site_distance(d);
        bond_accumulate(f, p.C, d, mk3(sg * Ei.x, sg * Ei.y, sg * Ei.z), fr);
        if (dr < MD_ANGLE_CUTOFF) {
            const float psiji = Ej.w - m.psi;
            const float fiji = Pj.w - m.fi;
            const float thetaji = s.L1(j).w - m.theta;
            const float th0 = gtp_i ? p.theta0_gtp : p.theta0_gdp;
            // +B sin(q_j - q_i - q0) on the R_MON > 0 side, -B sin(q_i - q_j - q0) on the other: sg (q_j - q_i) - q0, scaled by sg B
            f.psi = fmaf(sg * p.B_psi, sinf(fmaf(sg, psiji, -p.psi_0)), f.psi);
            f.fi = fmaf(sg * p.B_fi, sinf(fmaf(sg, fiji, -p.fi_0)), f.fi);
            f.theta = fmaf(sg * p.B_theta, sinf(fmaf(sg, thetaji, -th0)), f.theta);
        }
    }

    // ---- longitudinal Morse (+ barrier) + bending (compute_cuda.cu:189-299)
    const int nlong = (tw.y >> 8) & 0xff;
    for (int kk = 0; kk < nlong; kk++) {
        const unsigned code = topo_long(a, tw, traj, i, kk);
        const int j = code >> 1;
        const float sg = (code & 1u) ? 1.0f : -1.0f;
        const float4 Pj = s.P(j), Ej = s.E(j);
        const float4 L2j = s.L2(j);
        const int jf = __float_as_int(L2j.w);
        F3 d;
        d.x = fmaf(-sg, Ej.x, fmaf(-sg, Ei.x, Pj.x - xi));
        d.y = fmaf(-sg, Ej.y, fmaf(-sg, Ei.y, Pj.y - yi));
        d.z = fmaf(-sg, Ej.z, fmaf(-sg, Ei.z, Pj.z - zi));
        const float dr = site_distance(d);
        float dUdr;
        if (dr == 0) dUdr = 0.0f;
        else dUdr = dmorse(p.D_long, p.A_long, dr) / dr;
        if (k.barr_long_on && !ontub_i && !(jf & MF_ONTUB)) {
            if (dr != 0.0f) dUdr += dbarr(p.a_barr_long, p.r_barr_long, p.w_barr_long, dr) / dr;
        }
        bond_accumulate(f, dUdr, d, mk3(sg * Ei.x, sg * Ei.y, sg * Ei.z), fr);
        if (dr < MD_ANGLE_CUTOFF) {
            const float psiji = Ej.w - m.psi;
            const float fiji = Pj.w - m.fi;
            const float thetaji = s.L1(j).w - m.theta;
            // the dimer closer to the plus end rules theta0 (compute_cuda.cu:282-283)
            const bool gtp_last = (zi > Pj.z) ? gtp_i : ((jf & MF_GTP) != 0);
            const float th0 = gtp_last ? p.theta0_gtp : p.theta0_gdp;
            // +B sin(q_j - q_i - q0) on the R_MON > 0 side, -B sin(q_i - q_j - q0) on the other: sg (q_j - q_i) - q0, scaled by sg B
            f.psi = fmaf(sg * p.B_psi, sinf(fmaf(sg, psiji, -p.psi_0)), f.psi);
            f.fi = fmaf(sg * p.B_fi, sinf(fmaf(sg, fiji, -p.fi_0)), f.fi);
            f.theta = fmaf(sg * p.B_theta, sinf(fmaf(sg, thetaji, -th0)), f.theta);
        }
    }

    // ---- lateral Morse (+ barrier), seam scaling (compute_cuda.cu:304-466)
    const int nlat = (tw.y >> 16) & 0xff;
    if (nlat > 0) {
        const float4 L1i = s.L1(i), L2i = s.L2(i);
        const int type_i = MF_TYPE(m.flags);
        for (int kk = 0; kk < nlat; kk++) {
            const unsigned code = topo_lat(a, tw, traj, i, kk);
            const int j = code >> 1;
            const bool neg = (code & 1u) != 0;
            // negative entry: i interacts through p1, j through p2; positive: i through p2, j through p1
            const float4 Pj = s.P(j);
            const float4 L1j = s.L1(j), L2j = s.L2(j);
            const int jf = __float_as_int(L2j.w);
            const F3 oi = neg ? mk3(L1i.x, L1i.y, L1i.z) : mk3(L2i.x, L2i.y, L2i.z);
            const F3 oj = neg ? mk3(L2j.x, L2j.y, L2j.z) : mk3(L1j.x, L1j.y, L1j.z);
            F3 d;
            d.x = (Pj.x - xi) - oi.x + oj.x;
            d.y = (Pj.y - yi) - oi.y + oj.y;
            d.z = (Pj.z - zi) - oi.z + oj.z;
            const float dr = site_distance(d);
            float dUdr;
            if (dr == 0) dUdr = 0.0f;
            else if (type_i != (jf & 0x7f)) dUdr = dmorse(k.c.D_lat_seam, p.A_lat, dr) / dr;
            else dUdr = dmorse(p.D_lat, p.A_lat, dr) / dr;
            if (k.barr_lat_on && !ontub_i && !(jf & MF_ONTUB)) {
                if (dr != 0.0f) dUdr += dbarr(p.a_barr_lat, p.r_barr_lat, p.w_barr_lat, dr) / dr;
            }
            bond_accumulate(f, dUdr, d, oi, fr);
        }
    }

    // ---- LJ r^-6 repulsion from the Verlet list (compute_cuda.cu:470-495)
    if (p.lj_on) {
        const float amp = p.ljscale * p.ljsigma6;
        float fx = 0.f, fy = 0.f, fz = 0.f;
        const int ncnt = near.ok ? (int)near.cnt[i] : MD_NEAR_FULL;
        if (ncnt < MD_NEAR_FULL_EXACT) {
            // shared-memory near list: the listed pairs that can be inside the 6-nm cut-off (see MD_NEAR_R2)
            // An entry outside the force cut-off (or not LJ-listed) gets the coefficient 0, which leaves the accumulators
            // bit-for-bit unchanged, so the loop body has no branch.  A pair inside the +-1e-6 band around the exact
            // threshold (practically never) is only noted; the sum is then redone with the fp64 tie-break.
            const int n = ncnt;
            const int nstride = S::kGlobal ? near.stride : a.N; // compile-time choice: the shared-memory stage keeps rows of N
            const int16_t *nl = reinterpret_cast<const int16_t *>(near.list) + i; // bit 15 (LJ-listed) read as the sign
            const float lo = k.cut_force.lo, hi = k.cut_force.hi;
            bool band = false;
            if (MADDY_OPT_LJ && k.near_all_listed) {
                // every entry is LJ-listed (pairs cut-off beyond the near radius): no flag test, no index mask, and the
                // band check is one subtract + compare (a superset of [lo, hi]; a hit only triggers the exact redo below)
                const uint16_t *nu = near.list + i;
                const float mid = k.band_mid, hw = k.band_hw;
#pragma unroll 4
                for (int kk = 0; kk < n; kk++) {
                    const float4 Pj = s.P(nu[kk * nstride] & 0x7fffu);
                    const float dx = xi - Pj.x, dy = yi - Pj.y, dz = zi - Pj.z;
                    const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    band |= fabsf(sf - mid) <= hw;
                    const float inv = 1.0f / sf;
                    const float inv2 = inv * inv;
                    const float c = sf < lo ? amp * (6.0f * (inv2 * inv2)) : 0.0f; // 6 / dr^8
                    fx = fmaf(c, dx, fx);
                    fy = fmaf(c, dy, fy);
                    fz = fmaf(c, dz, fz);
                }
            } else
            for (int kk = 0; kk < n; kk++) {
                const int e = nl[kk * nstride];
                const float4 Pj = s.P(e & 0x7fff);
                const float dx = xi - Pj.x, dy = yi - Pj.y, dz = zi - Pj.z;
                const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const bool in = sf < lo && e < 0;
                band |= sf <= hi && !in; // also trips for an unlisted pair inside the cut-off (rare): the exact redo below gives the same sum
                const float inv = 1.0f / sf;
                const float inv2 = inv * inv;
                const float c = in ? amp * (6.0f * (inv2 * inv2)) : 0.0f; // 6 / dr^8
                fx = fmaf(c, dx, fx);
                fy = fmaf(c, dy, fy);
                fz = fmaf(c, dz, fz);
            }
            if (band) {
                fx = fy = fz = 0.f;
                for (int kk = 0; kk < n; kk++) {
                    const int e = nl[kk * nstride];
                    const float4 Pj = s.P(e & 0x7fff);
                    const float dx = xi - Pj.x, dy = yi - Pj.y, dz = zi - Pj.z;
                    const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const bool in = e < 0 && inside_cut(k.cut_force, dx, dy, dz, sf);
                    const float inv = 1.0f / sf;
                    const float inv2 = inv * inv;
                    const float c = in ? amp * (6.0f * (inv2 * inv2)) : 0.0f;
                    fx = fmaf(c, dx, fx);
                    fy = fmaf(c, dy, fy);
                    fz = fmaf(c, dz, fz);
                }
            }
        } else {
            const uint16_t *lj = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
            int n = a.ljcnt[(size_t)traj * a.Npad + i];
            if (near.stale_lj && ncnt != MD_NEAR_FULL_EXACT) { // (an escalated monomer's row was written out when it was escalated)
                const int nnc = a.ncandcnt[(size_t)traj * a.Npad + i];
                const bool all = nnc == MD_NEAR_FULL;
                lj = all ? a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i : a.ncand + (size_t)traj * MD_NCAND_CAPACITY * a.Npad + i;
                n = all ? (int)a.candcnt[(size_t)traj * a.Npad + i] : nnc;
            }
            if (kBatchWalk) {
                // every index load is an L2 round trip: fetch 16 indices, then their 16 positions, then accumulate in list
                // order (same arithmetic, same order: only the number of dependent trips changes)
                for (int k0 = 0; k0 < n; k0 += 16) {
                    int jv[16];
                    float4 Pv[16];
#pragma unroll
                    for (int q = 0; q < 16; q++) jv[q] = k0 + q < n ? (int)lj[(size_t)(k0 + q) * a.Npad] : -1;
#pragma unroll
                    for (int q = 0; q < 16; q++) Pv[q] = s.P(jv[q] < 0 ? i : jv[q]);
#pragma unroll
                    for (int q = 0; q < 16; q++) {
                        const float dx = xi - Pv[q].x, dy = yi - Pv[q].y, dz = zi - Pv[q].z;
                        const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (jv[q] >= 0 && inside_cut(k.cut_force, dx, dy, dz, sf)) {
                            const float inv = 1.0f / sf;
                            const float inv2 = inv * inv;
                            const float c = amp * (6.0f * (inv2 * inv2)); // 6 / dr^8
                            fx = fmaf(c, dx, fx);
                            fy = fmaf(c, dy, fy);
                            fz = fmaf(c, dz, fz);
                        }
                    }
                }
            } else {
#pragma unroll 4
            for (int kk = 0; kk < n; kk++) {
                const int j = lj[(size_t)kk * a.Npad];
                const float4 Pj = s.P(j);
                const float dx = xi - Pj.x, dy = yi - Pj.y, dz = zi - Pj.z;
                const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (inside_cut(k.cut_force, dx, dy, dz, sf)) {
                    const float inv = 1.0f / sf;
                    const float inv2 = inv * inv;
                    const float c = amp * (6.0f * (inv2 * inv2)); // 6 / dr^8
                    fx = fmaf(c, dx, fx);
                    fy = fmaf(c, dy, fy);
                    fz = fmaf(c, dz, fz);
                }
            }
            }
        }
        f.x += fx;
        f.y += fy;
        f.z += fz;
    }

    // ---- cylindrical repulsive walls (compute_cuda.cu:497-517); zs[traj] == rep_h
    if (p.is_wall) {
        if (zi < p.rep_leftborder) {
            f.z += p.rep_eps * fabsf(zi - p.rep_leftborder);
        } else if (zi > p.rep_h + p.rep_leftborder) {
            f.z += -p.rep_eps * fabsf(zi - (p.rep_h + p.rep_leftborder));
        }
        const float rad2 = xi * xi + yi * yi;
        if (rad2 > p.rep_r * p.rep_r) {
            const float rad = sqrtf(rad2);
            const float coeff = -p.rep_eps * (rad - p.rep_r);
            f.x += (xi / rad) * coeff;
            f.y += (yi / rad) * coeff;
        }
    }
    return f;
}

// ------------------------------------------------------------------ energies (compute_cuda.cu:676-911)
struct E7 { double harm, lng, lat, psi, fi, teta, lj; };

template <class S>
__device__ __forceinline__ E7 monomer_energy(const KArgs &k, const S &s, int traj, int i, const Mono &m)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    float U_lat = 0.f, U_long = 0.f, U_harm = 0.f, U_fi = 0.f, U_psi = 0.f, U_teta = 0.f, U_lj = 0.f;
    if (!(m.flags & MF_EXTRA)) {
        const float4 Ei = s.E(i);
        const float xi = m.x, yi = m.y, zi = m.z;
        const bool gtp_i = (m.flags & MF_GTP) != 0;
        const bool ontub_i = (m.flags & MF_ONTUB) != 0;
        const int nh = a.harm_count[i];
        for (int kk = 0; kk < nh; kk++) {
            int raw = a.harm[a.maxH * i + kk];
            const float sg = raw < 0 ? 1.0f : -1.0f;
            const int j = raw < 0 ? -raw : raw;
            const float4 Pj = s.P(j), Ej = s.E(j);
            F3 d;
            d.x = fmaf(-sg, Ej.x, fmaf(-sg, Ei.x, Pj.x - xi));
            d.y = fmaf(-sg, Ej.y, fmaf(-sg, Ei.y, Pj.y - yi));
            d.z = fmaf(-sg, Ej.z, fmaf(-sg, Ei.z, Pj.z - zi));
            const float dr = site_distance_d(d);
            U_harm = (float)((double)U_harm + (double)(p.C / 2) * ((double)dr * (double)dr));
            if (dr < MD_ANGLE_CUTOFF) {
                // quirk kept: psi and fi use (i - j), theta uses (j - i)  (compute_cuda.cu:751-757)
                const float psiij = -(Ej.w - m.psi);
                const float fiij = -(Pj.w - m.fi);
                const float thetaji = s.L1(j).w - m.theta;
                U_psi += p.B_psi * (1 - cosf(psiij - p.psi_0));
                U_fi += p.B_fi * (1 - cosf(fiij - p.fi_0));
                U_teta += p.B_theta * (1 - cosf(thetaji - (gtp_i ? p.theta0_gtp : p.theta0_gdp)));
            }
        }
        const uint16_t *bl = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
        const uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
        const int nlong = bc[0];
        for (int kk = 0; kk < nlong; kk++) {
            const unsigned code = bl[(size_t)kk * a.Npad];
            const int j = code >> 1;
            const float sg = (code & 1u) ? 1.0f : -1.0f;
            const float4 Pj = s.P(j), Ej = s.E(j);
            const int jf = __float_as_int(s.L2(j).w);
            F3 d;
            d.x = fmaf(-sg, Ej.x, fmaf(-sg, Ei.x, Pj.x - xi));
            d.y = fmaf(-sg, Ej.y, fmaf(-sg, Ei.y, Pj.y - yi));
            d.z = fmaf(-sg, Ej.z, fmaf(-sg, Ei.z, Pj.z - zi));
            // dr2 is rounded to float before the sqrt here (compute_cuda.cu:783-791)
            double s2 = (double)d.z * (double)d.z;
            s2 += (double)d.x * (double)d.x;
            s2 += (double)d.y * (double)d.y;
            const float dr = sqrtf((float)s2);
            U_long += morse_en(p.D_long, p.A_long, dr);
            if (p.barrier && !ontub_i && !(jf & MF_ONTUB)) U_long += barr(p.a_barr_long, p.r_barr_long, p.w_barr_long, dr);
            if (dr < MD_ANGLE_CUTOFF) {
                const float psiij = -(Ej.w - m.psi);
                const float fiij = -(Pj.w - m.fi);
                const float thetaij = -(s.L1(j).w - m.theta);
                const bool gtp_last = (zi > Pj.z) ? gtp_i : ((jf & MF_GTP) != 0);
                const float th0 = gtp_last ? p.theta0_gtp : p.theta0_gdp;
                U_psi += p.B_psi * (1 - cosf(psiij - p.psi_0));
                U_fi += p.B_fi * (1 - cosf(fiij - p.fi_0));
                U_teta += p.B_theta * (1 - cosf(thetaij - th0));
            }
        }
        const int nlat = bc[a.Npad];
        if (nlat > 0) {
            const float4 L1i = s.L1(i), L2i = s.L2(i);
            const int type_i = MF_TYPE(m.flags);
            for (int kk = 0; kk < nlat; kk++) {
                const unsigned code = bl[(size_t)(a.capLong + kk) * a.Npad];
                const int j = code >> 1;
                const bool neg = (code & 1u) != 0;
                const float4 Pj = s.P(j);
                const float4 L1j = s.L1(j), L2j = s.L2(j);
                const int jf = __float_as_int(L2j.w);
                const F3 oi = neg ? mk3(L1i.x, L1i.y, L1i.z) : mk3(L2i.x, L2i.y, L2i.z);
                const F3 oj = neg ? mk3(L2j.x, L2j.y, L2j.z) : mk3(L1j.x, L1j.y, L1j.z);
                F3 d;
                d.x = (Pj.x - xi) - oi.x + oj.x;
                d.y = (Pj.y - yi) - oi.y + oj.y;
                d.z = (Pj.z - zi) - oi.z + oj.z;
                const float dr = site_distance_d(d);
                if (type_i != (jf & 0x7f)) U_lat += morse_en(p.D_lat / p.seam_coeff, p.A_lat, dr);
                else U_lat += morse_en(p.D_lat, p.A_lat, dr);
                if (p.barrier && !ontub_i && !(jf & MF_ONTUB)) U_lat += barr(p.a_barr_lat, p.r_barr_lat, p.w_barr_lat, dr);
            }
        }
        if (p.lj_on) {
            const uint16_t *lj = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
            const int n = a.ljcnt[(size_t)traj * a.Npad + i];
            // 16 index loads per round trip (the list lives in HBM/L2), then the pairs in list order
            for (int k0 = 0; k0 < n; k0 += 16) {
                int jv[16];
#pragma unroll
                for (int q = 0; q < 16; q++) jv[q] = k0 + q < n ? (int)lj[(size_t)(k0 + q) * a.Npad] : -1;
#pragma unroll
                for (int q = 0; q < 16; q++) {
                    if (jv[q] < 0) continue;
                    const float4 Pj = s.P(jv[q]);
                    const float dx = xi - Pj.x, dy = yi - Pj.y, dz = zi - Pj.z;
                    const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (inside_cut(k.cut_force, dx, dy, dz, sf)) {
                        const float dr = (float)sqrt(dist2_exact(dx, dy, dz));
                        const double dr2 = (double)dr * (double)dr;
                        U_lj = (float)((double)U_lj + (double)(p.ljscale * p.ljsigma6) / (dr2 * dr2 * dr2));
                    }
                }
            }
        }
    }
    E7 e;
    e.harm = U_harm / 2;
    e.lng = U_long / 2;
    e.lat = U_lat / 2;
    e.lj = U_lj / 2;
    e.psi = U_psi / 2;
    e.fi = U_fi / 2;
    e.teta = U_teta / 2;
    return e;
}

// ------------------------------------------------------------------ list rebuild
// Dynamic bond candidates of monomer i against monomer j (pairs_kernel, compute_cuda.cu:564-669):
// longitudinal (other monomer type, end sites opposite to the dimer bond) then the two lateral pairings.
struct BondOut {
    uint16_t *col; // &bl[traj][0][i]
    int nlong, nlat, status;
};
// bit0: longitudinal hit, bit1: lateral (i:p1, j:p2) hit, bit2: lateral (i:p2, j:p1) hit
template <class S>
__device__ __forceinline__ unsigned bond_tests(const S &s, float xi, float yi, float zi, int type_i, int j, int hraw, const float4 &Pj,
                                               const float4 &Ei, const float4 &L1i, const float4 &L2i)
{
    const float4 Ej = s.E(j), L1j = s.L1(j), L2j = s.L2(j);
    const int jf = __float_as_int(L2j.w);
    unsigned hit = 0;
    if (type_i != (jf & 0x7f)) {
        const float sg = hraw < 0 ? -1.0f : 1.0f; // R_MON / r_mon (compute_cuda.cu:548-551)
        F3 d;
        d.x = fmaf(-sg, Ej.x, fmaf(-sg, Ei.x, Pj.x - xi));
        d.y = fmaf(-sg, Ej.y, fmaf(-sg, Ei.y, Pj.y - yi));
        d.z = fmaf(-sg, Ej.z, fmaf(-sg, Ei.z, Pj.z - zi));
        if (site_distance(d) < MD_PAIR_CUTOFF) hit |= 1u;
    }
    // lateral: (i:p1, j:p2) stored negative, then (i:p2, j:p1) stored positive (compute_cuda.cu:612-661)
    F3 d;
    d.x = (Pj.x - xi) - L1i.x + L2j.x;
    d.y = (Pj.y - yi) - L1i.y + L2j.y;
    d.z = (Pj.z - zi) - L1i.z + L2j.z;
    if (site_distance(d) < MD_PAIR_CUTOFF) hit |= 2u;
    d.x = (Pj.x - xi) - L2i.x + L1j.x;
    d.y = (Pj.y - yi) - L2i.y + L1j.y;
    d.z = (Pj.z - zi) - L2i.z + L1j.z;
    if (site_distance(d) < MD_PAIR_CUTOFF) hit |= 4u;
    return hit;
}
// longitudinal entries are stored as +j when harmonic < 0, -j otherwise; -0 == 0 loses its sign (compute_cuda.cu:588-592)
__device__ __forceinline__ uint16_t long_code(int j, int hraw) { return (uint16_t)((j << 1) | ((hraw < 0 || j == 0) ? 0u : 1u)); }

template <class S>
__device__ __forceinline__ void bond_candidates(const DevSys &a, const S &s, int i, int j, const Mono &m, int hraw, const float4 &Pj,
                                                const float4 &Ei, const float4 &L1i, const float4 &L2i, BondOut &o)
{
    const unsigned hit = bond_tests(s, m.x, m.y, m.z, MF_TYPE(m.flags), j, hraw, Pj, Ei, L1i, L2i);
    if (hit & 1u) {
        if (o.nlong < a.capLong) o.col[(size_t)o.nlong * a.Npad] = long_code(j, hraw);
        else o.status |= ST_LONG_OVERFLOW;
        o.nlong++;
    }
    if (hit & 2u) {
        if (o.nlat < a.capLat) o.col[(size_t)(a.capLong + o.nlat) * a.Npad] = (uint16_t)((j << 1) | 1u);
        else o.status |= ST_LAT_OVERFLOW;
        o.nlat++;
    }
    if (hit & 4u) {
        if (o.nlat < a.capLat) o.col[(size_t)(a.capLong + o.nlat) * a.Npad] = (uint16_t)(j << 1);
        else o.status |= ST_LAT_OVERFLOW;
        o.nlat++;
    }
}

// Warp-cooperative rebuild of ONE monomer's rows from its candidate list: 32 lanes test 32 candidates at a time and
// the hits are compacted in candidate (= ascending j) order with ballots.  Used for the fixed monomers of the fused
// loop, which own no thread: every warp looks after one or two of them, so no warp pays a whole extra pass.
__device__ __forceinline__ void rebuild_row_cooperative(const KArgs &k, const Stage &s, int traj, int i, unsigned ops)
{
    const DevSys &a = k.a;
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const bool do_lj = (ops & OP_REBUILD_LJ) != 0, do_b = (ops & OP_REBUILD_BONDS) != 0;
    const float4 Pi = s.P(i), Ei = s.E(i), L1i = s.L1(i), L2i = s.L2(i);
    const int fl = __float_as_int(L2i.w);
    const int hraw = a.harm[a.maxH * i];
    const int hp = hraw < 0 ? -hraw : hraw;
    const bool extra = (fl & MF_EXTRA) != 0;
    const int n = extra ? 0 : (int)a.candcnt[(size_t)traj * a.Npad + i];
    const uint16_t *cp = a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i;
    uint16_t *lp = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
    uint16_t *bp = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
    int nlj = 0, nlong = 0, nlat = 0, status = 0;
    for (int b0 = 0; b0 < n; b0 += 32) {
        const int kk = b0 + lane;
        const bool valid = kk < n;
        const int j = valid ? (int)cp[(size_t)kk * a.Npad] : 0;
        const float4 Pj = s.P(j);
        const float dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
        const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        if (do_lj) {
            const bool in = valid && inside_cut(k.cut_pairs, dx, dy, dz, sf);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (in) {
                const int pos = nlj + __popc(m & lt);
                if (pos < MADDY_LJ_CAPACITY) lp[(size_t)pos * a.Npad] = (uint16_t)j;
                else status |= ST_LJ_OVERFLOW;
            }
            nlj += __popc(m);
        }
        if (do_b) {
            unsigned hit = 0;
            if (valid && sf < MD_BOND_PREFILTER2 && hp != j) hit = bond_tests(s, Pi.x, Pi.y, Pi.z, fl & 0x7f, j, hraw, Pj, Ei, L1i, L2i);
            const unsigned ml = __ballot_sync(0xffffffffu, hit & 1u), m1 = __ballot_sync(0xffffffffu, hit & 2u),
                           m2 = __ballot_sync(0xffffffffu, hit & 4u);
            if (hit & 1u) {
                const int pos = nlong + __popc(ml & lt);
                if (pos < a.capLong) bp[(size_t)pos * a.Npad] = long_code(j, hraw);
                else status |= ST_LONG_OVERFLOW;
            }
            const int before = nlat + __popc(m1 & lt) + __popc(m2 & lt);
            if (hit & 2u) {
                if (before < a.capLat) bp[(size_t)(a.capLong + before) * a.Npad] = (uint16_t)((j << 1) | 1u);
                else status |= ST_LAT_OVERFLOW;
            }
            if (hit & 4u) {
                const int pos = before + ((hit & 2u) ? 1 : 0);
                if (pos < a.capLat) bp[(size_t)(a.capLong + pos) * a.Npad] = (uint16_t)(j << 1);
                else status |= ST_LAT_OVERFLOW;
            }
            nlong += __popc(ml);
            nlat += __popc(m1) + __popc(m2);
        }
    }
    if (lane == 0) {
        if (do_lj) a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
        if (do_b) {
            uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
            bc[0] = (uint8_t)min(nlong, a.capLong);
            bc[a.Npad] = (uint8_t)min(nlat, a.capLat);
        }
    }
    if (status) atomicOr(a.status, status);
}

// General path: one pass over ALL j for the MPT monomers of this thread (LJ Verlet list,
// compute_cuda.cu:913-940, and bond lists).  Used by the step-granular entry points when the near
// list is disabled and as the fallback when a near list overflows.
template <int MPT, class S>
__device__ __forceinline__ void rebuild_lists_all_pairs(const KArgs &k, const S &s, int traj, const Mono (&mo)[MPT],
                                                        const int (&idx)[MPT], unsigned ops)
{
    const DevSys &a = k.a;
    const int N = a.N;
    const bool do_lj = (ops & OP_REBUILD_LJ) != 0;
    const bool do_b = (ops & OP_REBUILD_BONDS) != 0;
    uint16_t *ljb = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad;
    int status = 0;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        int nlj = 0;
        BondOut bo;
        bo.col = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
        bo.nlong = bo.nlat = bo.status = 0;
        if (!(mo[t].flags & MF_EXTRA)) {
            const int hraw = a.harm[a.maxH * i]; // first entry, whatever harmonicCount says (compute_cuda.cu:548)
            const int hp = hraw < 0 ? -hraw : hraw;
            const float4 Ei = s.E(i), L1i = s.L1(i), L2i = s.L2(i);
            for (int j = 0; j < N; j++) {
                const float4 Pj = s.P(j);
                const float dx = mo[t].x - Pj.x, dy = mo[t].y - Pj.y, dz = mo[t].z - Pj.z;
                const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (i == j) continue;
                if (do_lj && inside_cut(k.cut_pairs, dx, dy, dz, sf)) {
                    if (nlj < MADDY_LJ_CAPACITY) ljb[(size_t)nlj * a.Npad + i] = (uint16_t)j;
                    else status |= ST_LJ_OVERFLOW;
                    nlj++;
                }
                if (do_b && sf < MD_BOND_PREFILTER2 && hp != j) bond_candidates(a, s, i, j, mo[t], hraw, Pj, Ei, L1i, L2i, bo);
            }
        }
        if (do_lj) a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
        if (do_b) {
            uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
            bc[0] = (uint8_t)min(bo.nlong, a.capLong);
            bc[a.Npad] = (uint8_t)min(bo.nlat, a.capLat);
        }
        status |= bo.status;
    }
    if (status) atomicOr(a.status, status);
}

// Bounding boxes of tiles of MD_TILE consecutive monomers (caller syncs afterwards).
__device__ __forceinline__ void compute_tiles(const Stage &s, const Near &near, int N)
{
    for (int t = threadIdx.x; t < near.ntiles; t += blockDim.x) {
        const int j0 = t * MD_TILE, j1 = min(N, j0 + MD_TILE);
        float4 p = s.P(j0);
        float lx = p.x, ly = p.y, lz = p.z, hx = p.x, hy = p.y, hz = p.z;
        for (int j = j0 + 1; j < j1; j++) {
            p = s.P(j);
            lx = fminf(lx, p.x); ly = fminf(ly, p.y); lz = fminf(lz, p.z);
            hx = fmaxf(hx, p.x); hy = fmaxf(hy, p.y); hz = fmaxf(hz, p.z);
        }
        near.tlo[t] = make_float4(lx, ly, lz, 0.f);
        near.thi[t] = make_float4(hx, hy, hz, 0.f);
    }
}

// Fast path, phase 1a (rare): tile-culled scan that builds the CANDIDATE list.  Every lane walks ITS OWN
// sequence of tiles whose bounding box is within the candidate radius (lanes of a warp sit far apart along
// the protofilament, so a warp-uniform tile loop would visit nearly every tile); tiles are visited in
// ascending order, so candidates come out in ascending j like the reference's all-pairs loop.
// Returns true if a candidate list overflowed.
template <int MPT>
__device__ __forceinline__ bool scan_candidates(const KArgs &k, const Stage &s, const Near &near, int traj, const Mono (&mo)[MPT],
                                                const int (&idx)[MPT])
{
    const DevSys &a = k.a;
    const int N = a.N;
    const float rc2 = k.rcand2;
    bool ovf = false;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        int nc = 0, nnc = 0;
        if (!(mo[t].flags & MF_EXTRA)) {
            uint16_t *cp = a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i;
            uint16_t *np = a.ncand + (size_t)traj * MD_NCAND_CAPACITY * a.Npad + i;
            const float x = mo[t].x, y = mo[t].y, z = mo[t].z;
            int tile = 0;
            for (;;) {
                while (tile < near.ntiles) {
                    const float4 lo = near.tlo[tile], hi = near.thi[tile];
                    const float ex = fmaxf(fmaxf(lo.x - x, x - hi.x), 0.f);
                    const float ey = fmaxf(fmaxf(lo.y - y, y - hi.y), 0.f);
                    const float ez = fmaxf(fmaxf(lo.z - z, z - hi.z), 0.f);
                    if (fmaf(ez, ez, fmaf(ey, ey, ex * ex)) < rc2) break;
                    tile++;
                }
                if (tile >= near.ntiles) break;
                const int j0 = tile * MD_TILE, j1 = min(N, j0 + MD_TILE);
                for (int j = j0; j < j1; j++) {
                    const float4 Pj = s.P(j);
                    const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (d2 < rc2 && j != i) {
                        if (nc < MD_CAND_CAPACITY) {
                            *cp = (uint16_t)j;
                            cp += a.Npad;
                        }
                        nc++;
                        if (d2 < MD_NCAND_R2) {
                            if (nnc < MD_NCAND_CAPACITY) {
                                *np = (uint16_t)j;
                                np += a.Npad;
                            }
                            nnc++;
                        }
                    }
                }
                tile++;
            }
        }
        a.candcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nc, MD_CAND_CAPACITY);
        a.ncandcnt[(size_t)traj * a.Npad + i] = (uint8_t)(nnc > MD_NCAND_CAPACITY ? MD_NEAR_FULL : nnc);
        ovf |= nc > MD_CAND_CAPACITY;
    }
    return ovf;
}

__device__ __forceinline__ float4 lds_f4(unsigned addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u16(unsigned addr, unsigned short v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v)); }
__device__ __forceinline__ void stg_u16(const uint16_t *p, unsigned short v) { asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"(v)); }

// Fast path, phase 1b (every list-update step): the reference's Verlet list = candidates that pass the exact
// cut-off test (HBM, k-major); near list (SMEM) = candidates within MD_NEAR_R.  Returns near-list overflow.
template <int MPT>
__device__ __forceinline__ bool filter_candidates(const KArgs &k, const Stage &s, const Near &near, int traj, const Mono (&mo)[MPT],
                                                  const int (&idx)[MPT], bool do_lj)
{
    const DevSys &a = k.a;
    const int N = a.N;
    int status = 0;
    bool ovf = false;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        int nlj = 0, nn = 0;
        if (!(mo[t].flags & MF_EXTRA)) {
            const uint16_t *cp = a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i;
            const uint16_t *lj_col = (const uint16_t *)__cvta_generic_to_global(a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i);
            asm volatile("" : "+l"(lj_col)); // keep the column base in registers (otherwise re-derived at every store)
            unsigned near_s = (unsigned)__cvta_generic_to_shared(near.list + i);
            const unsigned near_end = near_s + (unsigned)(near.cap * N) * 2u, near_step = (unsigned)N * 2u;
            const int n = a.candcnt[(size_t)traj * a.Npad + i];
            const float x = mo[t].x, y = mo[t].y, z = mo[t].z;
            const int row = a.Npad;
            const float lo = do_lj ? k.cut_pairs.lo : -1.f, hi = do_lj ? k.cut_pairs.hi : -1.f;
            int lj_off = 0; // element offset of the next free slot in this monomer's Verlet-list column
            const int lj_end = MADDY_LJ_CAPACITY * row;
            // candidates are fetched MD_FILTER_BATCH at a time so the L2 latency of the index loads is paid once per
            // batch instead of once per candidate; list writes are predicated stores
            for (int k0 = 0; k0 < n; k0 += MD_FILTER_BATCH, cp += MD_FILTER_BATCH * (size_t)row) {
                unsigned jj[MD_FILTER_BATCH];
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) jj[u] = k0 + u < n ? (unsigned)cp[u * (size_t)row] : 0u;
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) {
                    const float4 Pj = s.P((int)jj[u]);
                    const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const float sf = k0 + u < n ? d2 : 3.0e38f; // slots past the end of the list are outside every radius
                    bool inlj = sf < lo;
                    if (sf >= lo && sf <= hi) inlj = dist2_exact(dx, dy, dz) < k.cut_pairs.t; // +-1e-6 band: practically never
                    if (inlj && lj_off < lj_end) stg_u16(lj_col + lj_off, (unsigned short)jj[u]);
                    lj_off += inlj ? row : 0;
                    const bool innear = sf < MD_NEAR_R2;
                    if (innear && near_s < near_end) sts_u16(near_s, (unsigned short)(jj[u] | (inlj ? MD_NEAR_LJ_FLAG : 0u)));
                    near_s += innear ? near_step : 0u;
                    nn += innear;
                    nlj += inlj;
                }
            }
            if (nlj > MADDY_LJ_CAPACITY) status |= ST_LJ_OVERFLOW;
        }
        if (do_lj) a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
        // a crowded monomer (more near partners than the cache holds) simply keeps walking its full lists
        near.cnt[i] = (uint8_t)(nn > near.cap ? MD_NEAR_FULL : nn);
        ovf |= nn > near.cap;
    }
    if (status) atomicOr(a.status, status);
    return ovf;
}

// The exact Verlet row of monomer i as of the last list-update step of the fused loop: candidates (valid for that step)
// re-tested with the reference's cut-off test on the positions recorded then (a.rpos, gathered from HBM/L2).  Same
// order and same arithmetic as filter_candidates.  Runs on demand only (maddy_download_list, energies, a window that
// starts between two list-update steps, a tripped displacement guard).
__device__ __forceinline__ void materialise_row(const KArgs &k, int traj, int i)
{
    const DevSys &a = k.a;
    const size_t base = (size_t)traj * a.N;
    int nlj = 0, status = 0;
    if (!a.extra[base + i]) {
        const float4 Pi = a.rpos[base + i];
        const uint16_t *cp = a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i;
        uint16_t *lp = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
        const int n = a.candcnt[(size_t)traj * a.Npad + i];
        for (int kk = 0; kk < n; kk++) {
            const int j = cp[(size_t)kk * a.Npad];
            const float4 Pj = a.rpos[base + j];
            const float dx = Pi.x - Pj.x, dy = Pi.y - Pj.y, dz = Pi.z - Pj.z;
            const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (inside_cut(k.cut_pairs, dx, dy, dz, sf)) {
                if (nlj < MADDY_LJ_CAPACITY) lp[(size_t)nlj * a.Npad] = (uint16_t)j;
                else status |= ST_LJ_OVERFLOW;
                nlj++;
            }
        }
    }
    a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
    if (status) atomicOr(a.status, status);
}

// Lazy variant of phase 1b (fused loop): only the near list is refreshed, from the near-candidates.  Valid when the
// pairs cut-off exceeds MD_NEAR_R + MD_CAND_SKIN, so that every pair inside the near radius is certainly listed.
template <int MPT>
__device__ __forceinline__ bool near_from_candidates(const KArgs &k, const Stage &s, const Near &near, int traj, const Mono (&mo)[MPT],
                                                     const int (&idx)[MPT])
{
    const DevSys &a = k.a;
    const int N = a.N;
    bool ovf = false;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        int nn = 0;
        if (!(mo[t].flags & MF_EXTRA)) {
            const int nnc = a.ncandcnt[(size_t)traj * a.Npad + i];
            const bool full = nnc == MD_NEAR_FULL;
            const int n = full ? (int)a.candcnt[(size_t)traj * a.Npad + i] : nnc;
            const uint16_t *cp = full ? a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i : a.ncand + (size_t)traj * MD_NCAND_CAPACITY * a.Npad + i;
            const float x = mo[t].x, y = mo[t].y, z = mo[t].z;
            const size_t row = a.Npad;
            unsigned near_s = (unsigned)__cvta_generic_to_shared(near.list + i);
            const unsigned near_end = near_s + (unsigned)(near.cap * N) * 2u, near_step = (unsigned)N * 2u;
            for (int k0 = 0; k0 < n; k0 += MD_FILTER_BATCH, cp += MD_FILTER_BATCH * row) {
                unsigned jj[MD_FILTER_BATCH];
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) jj[u] = k0 + u < n ? (unsigned)cp[u * row] : 0u;
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) {
                    const float4 Pj = s.P((int)jj[u]);
                    const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                    const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    const bool innear = k0 + u < n && d2 < MD_NEAR_R2;
                    if (innear && near_s < near_end) sts_u16(near_s, (unsigned short)(jj[u] | MD_NEAR_LJ_FLAG));
                    near_s += innear ? near_step : 0u;
                    nn += innear;
                }
            }
        }
        near.cnt[i] = (uint8_t)(nn > near.cap ? MD_NEAR_FULL : nn);
        ovf |= nn > near.cap;
    }
    return ovf;
}

// Fast path, phase 2: bond lists from the near list (every candidate lies within 6.6 nm < MD_NEAR_R).
template <int MPT>
__device__ __forceinline__ void bonds_from_near(const KArgs &k, const Stage &s, const Near &near, int traj, const Mono (&mo)[MPT],
                                                const int (&idx)[MPT])
{
    const DevSys &a = k.a;
    const int N = a.N;
    int status = 0;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        BondOut bo;
        bo.col = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
        bo.nlong = bo.nlat = bo.status = 0;
        if (!(mo[t].flags & MF_EXTRA)) {
            const int hraw = a.harm[a.maxH * i];
            const int hp = hraw < 0 ? -hraw : hraw;
            const float4 Ei = s.E(i), L1i = s.L1(i), L2i = s.L2(i);
            const bool full = near.cnt[i] == MD_NEAR_FULL; // near list overflowed: take the candidates (same ascending order)
            const int n = full ? (int)a.candcnt[(size_t)traj * a.Npad + i] : (int)near.cnt[i];
            const uint16_t *cp = a.cand + (size_t)traj * MD_CAND_CAPACITY * a.Npad + i;
            for (int kk = 0; kk < n; kk++) {
                const int j = full ? (int)cp[(size_t)kk * a.Npad] : (int)(near.list[kk * N + i] & 0x7fffu);
                const float4 Pj = s.P(j);
                const float dx = mo[t].x - Pj.x, dy = mo[t].y - Pj.y, dz = mo[t].z - Pj.z;
                const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                if (sf < MD_BOND_PREFILTER2 && hp != j) bond_candidates(a, s, i, j, mo[t], hraw, Pj, Ei, L1i, L2i, bo);
            }
        }
        uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
        bc[0] = (uint8_t)min(bo.nlong, a.capLong);
        bc[a.Npad] = (uint8_t)min(bo.nlat, a.capLat);
        status |= bo.status;
    }
    if (status) atomicOr(a.status, status);
}

// Near list := entries of the full LJ list (HBM) whose CURRENT distance is below MD_NEAR_R.
// Used at the start of a fused window and when the displacement guard trips.  Returns overflow.
template <int MPT>
__device__ __forceinline__ bool refresh_near(const KArgs &k, const Stage &s, const Near &near, int traj, const Mono (&mo)[MPT],
                                             const int (&idx)[MPT])
{
    const DevSys &a = k.a;
    const int N = a.N;
    bool ovf = false;
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i >= N) continue;
        int nn = 0;
        if (k.p.lj_on && !(mo[t].flags & MF_EXTRA)) {
            const uint16_t *lj = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
            const int n = a.ljcnt[(size_t)traj * a.Npad + i];
            const size_t row = a.Npad;
            // indices fetched MD_FILTER_BATCH at a time (one L2 round trip per batch, as in filter_candidates)
            for (int k0 = 0; k0 < n; k0 += MD_FILTER_BATCH, lj += MD_FILTER_BATCH * row) {
                unsigned jj[MD_FILTER_BATCH];
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) jj[u] = k0 + u < n ? (unsigned)lj[u * row] : 0u;
#pragma unroll
                for (int u = 0; u < MD_FILTER_BATCH; u++) {
                    const float4 Pj = s.P(jj[u]);
                    const float dx = mo[t].x - Pj.x, dy = mo[t].y - Pj.y, dz = mo[t].z - Pj.z;
                    if (k0 + u < n && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < MD_NEAR_R2) {
                        if (nn < near.cap) near.list[nn * N + i] = (uint16_t)(jj[u] | MD_NEAR_LJ_FLAG);
                        nn++;
                    }
                }
            }
        }
        near.cnt[i] = (uint8_t)(nn > near.cap ? MD_NEAR_FULL : nn);
        ovf |= nn > near.cap;
    }
    return ovf;
}

// Candidate-list state of one trajectory, carried in registers through a launch (and in HBM between launches).
struct CandState {
    int valid; // CTA-uniform; the positions the list was built from live in HBM (a.cpos), read every list-update step
    bool dirty;
};

// Rebuild at a list-update step.  CTA-uniform control flow; returns whether the near list is valid.
// with_fixed (fused loop): the fixed monomers own no thread.  On the common path (valid candidate lists) every warp
// rebuilds one or two of their rows cooperatively; on the rare paths (candidate scan, overflow fallback) thread f
// handles fixed monomer f with the per-thread functions.
template <int MPT>
__device__ __forceinline__ bool rebuild_lists(const KArgs &k, const Stage &s, Near &near, CandState &cs, int traj,
                                              const Mono (&mo)[MPT], const int (&idx)[MPT], unsigned ops, bool with_fixed,
                                              bool lazy, bool &lj_exact)
{
    lj_exact = true; // set to false only on the lazy path below
    const DevSys &a = k.a;
    const int N = a.N;
    Mono fm[1];
    int fi[1] = {N};
    fm[0].flags = MF_EXTRA | MF_FIXED;
    if (with_fixed && (int)threadIdx.x < a.n_fixed) {
        fi[0] = (int)a.fmap[threadIdx.x];
        const float4 P = s.P(fi[0]);
        const int jf = __float_as_int(s.L2(fi[0]).w);
        fm[0].x = P.x; fm[0].y = P.y; fm[0].z = P.z;
        fm[0].flags = MF_FIXED | ((jf & 0x7f) << 1) | (jf & (MF_GTP | MF_ONTUB | MF_EXTRA));
    }
    auto count_event = [&](int which) {
        if (threadIdx.x == 0) atomicAdd(a.stats + which, 1ull);
    };
    if (near.cap == 0) {
        rebuild_lists_all_pairs<MPT>(k, s, traj, mo, idx, ops);
        if (with_fixed) rebuild_lists_all_pairs<1>(k, s, traj, fm, fi, ops);
        return false;
    }
    // has anything moved more than half the candidate skin since the candidate list was built?
    const size_t base = (size_t)traj * N;
    bool moved = !cs.valid;
    if (cs.valid) {
#pragma unroll
        for (int t = 0; t < MPT; t++) {
            if (idx[t] < N) {
                const float4 c = a.cpos[base + idx[t]];
                const float dx = mo[t].x - c.x, dy = mo[t].y - c.y, dz = mo[t].z - c.z;
                moved |= fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > MD_CAND_GUARD2;
            }
        }
    }
    if (__syncthreads_or(moved)) {
        count_event(1);
        compute_tiles(s, near, N);
        __syncthreads();
        bool covf = scan_candidates<MPT>(k, s, near, traj, mo, idx);
        if (with_fixed) covf |= scan_candidates<1>(k, s, near, traj, fm, fi);
        cs.dirty = true;
        if (__syncthreads_or(covf)) { // more candidates than MD_CAND_CAPACITY: general path, candidates stay invalid
            count_event(2);
            cs.valid = 0;
            rebuild_lists_all_pairs<MPT>(k, s, traj, mo, idx, ops);
            if (with_fixed) rebuild_lists_all_pairs<1>(k, s, traj, fm, fi, ops);
            return false;
        }
        cs.valid = 1;
#pragma unroll
        for (int t = 0; t < MPT; t++)
            if (idx[t] < N) a.cpos[base + idx[t]] = make_float4(mo[t].x, mo[t].y, mo[t].z, 0.f);
        if (fi[0] < N) a.cpos[base + fi[0]] = make_float4(fm[0].x, fm[0].y, fm[0].z, 0.f);
        __syncthreads(); // candidate rows of the fixed monomers are read by other threads below
    }
    lj_exact = !lazy;
    if (lazy ? near_from_candidates<MPT>(k, s, near, traj, mo, idx)
             : filter_candidates<MPT>(k, s, near, traj, mo, idx, (ops & OP_REBUILD_LJ) != 0))
        atomicAdd(a.stats + 3, 1ull);
    if (ops & OP_REBUILD_BONDS) bonds_from_near<MPT>(k, s, near, traj, mo, idx);
    if (with_fixed) {
        const int nwarp = blockDim.x >> 5;
        const unsigned fops = lazy ? (ops & ~(unsigned)OP_REBUILD_LJ) : ops;
        if (fops)
            for (int f = threadIdx.x >> 5; f < a.n_fixed; f += nwarp) rebuild_row_cooperative(k, s, traj, (int)a.fmap[f], fops);
    }
    return (ops & OP_REBUILD_LJ) != 0 || !k.p.lj_on;
}

// ------------------------------------------------------------------ integrator (compute_cuda.cu:943-975)
__device__ __forceinline__ void integrate_monomer(const KArgs &k, Mono &m, const G6 &f)
{
    const maddy_params &p = k.p;
    if (!(m.flags & MF_FIXED) && !(m.flags & MF_EXTRA)) {
        const float4 rf_xyz = rforce(m.rx);
        const float4 rf_ang = rforce(m.ra);
        // same shape as the reference's SASS: noise * var (FMUL), force * (dt/gamma) + that (FFMA), coordinate + that (FADD)
        const float aR = k.c.aR, aA = k.c.aA, aT = k.c.aT, vA = k.c.vA;
        m.x += fmaf(f.x, aR, p.varR * rf_xyz.x);
        m.y += fmaf(f.y, aR, p.varR * rf_xyz.y);
        m.z += fmaf(f.z, aR, p.varR * rf_xyz.z);
        m.fi += fmaf(f.fi, aA, vA * rf_ang.x);
        m.psi += fmaf(f.psi, aA, vA * rf_ang.y);
        m.theta += fmaf(f.theta, aT, p.varTheta * rf_ang.z);
    }
}

// ------------------------------------------------------------------ block reduction of 7 doubles
__device__ __forceinline__ void block_reduce_e7(E7 e, double *out7, double *scratch /* [32*7] shared */)
{
    double v[7] = {e.harm, e.lng, e.lat, e.psi, e.fi, e.teta, e.lj};
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int q = 0; q < 7; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_down_sync(0xffffffffu, v[q], o);
    }
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 7; q++) scratch[warp * 7 + q] = v[q];
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int q = 0; q < 7; q++) {
            double w = lane < nwarp ? scratch[lane * 7 + q] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
            if (lane == 0) out7[q] = w;
        }
    }
}

__device__ __forceinline__ void store_cand_state(const DevSys &a, const CandState &cs, int traj)
{
    if (cs.dirty && threadIdx.x == 0) a.cand_valid[traj] = cs.valid;
}

// ------------------------------------------------------------------ the trajectory kernels
// Dynamic shared memory (both kernels): [nbuf][4][N] float4 stage | 2*ntiles float4 tile boxes |
// near list u16[cap][N] | u8[N] counts | (run kernel) RNG streams uint4[2][N] | topology words uint4[N]
__device__ __forceinline__ Near carve_near(float4 *smem, const KArgs &k, int N)
{
    Near near;
    near.cap = k.near_cap;
    near.ntiles = (N + MD_TILE - 1) / MD_TILE;
    near.tlo = smem + (size_t)k.nbuf * 4 * N;
    near.thi = near.tlo + near.ntiles;
    near.list = reinterpret_cast<uint16_t *>(near.thi + near.ntiles);
    near.cnt = reinterpret_cast<uint8_t *>(near.list + (size_t)near.cap * N);
    near.ok = false;
    near.stale_lj = false;
    near.topo = nullptr;
    near.stride = N;
    return near;
}

__device__ __forceinline__ void load_mono(const DevSys &a, size_t base, int i, Mono &m)
{
    const float4 P = a.pos[base + i], A = a.ang[base + i];
    m.x = P.x; m.y = P.y; m.z = P.z;
    m.fi = A.x; m.psi = A.y; m.theta = A.z;
    m.flags = (int)a.sflags[i] | (a.gtp[base + i] == 1 ? MF_GTP : 0) | (a.ontub[base + i] ? MF_ONTUB : 0) | (a.extra[base + i] ? MF_EXTRA : 0);
}

/*
 * run_kernel<MPT, MINB> — the fused step loop (OP_RUN).
 *
 * Threads are handed out to the monomers that can MOVE: a.amap lists the non-fixed monomers (the bottom ring of a
 * seeded microtubule, `fix`, never moves: its force is computed and discarded by the reference, compute_cuda.cu:949).
 * Fixed monomers are published once into both stage buffers and get their list rows rebuilt by threads
 * 0..n_fixed-1 as one extra monomer per thread.  For the 520-monomer seed that leaves 494 movers = 16 warps exactly,
 * which is what lets TWO CTAs share an SM at 64 registers each (registers are allocated per 4-warp group).
 * MINB = 2: HybridTaus streams wait in shared memory between integrator calls; MINB = 1: up to 128 registers.
 */
template <int MPT, int MINB>
__global__ void __launch_bounds__(MD_RUN_THREADS, MINB) run_kernel(const __grid_constant__ KArgs k)
{
    float4 *const smem = g_smem;
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    if (*a.guard) return; // an on-tubule classification queued before this window could not be decided: the host redoes the stride
    const int N = a.N;
    const int traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    const LatSite ls = lateral_site();

    Near near = carve_near(smem, k, N);
    constexpr bool kRngShared = MINB == 2;
    uint4 *srng = reinterpret_cast<uint4 *>(smem) + (k.rng_smem_offset >> 4);
    if (k.topo_smem_offset >= 0) near.topo = reinterpret_cast<uint4 *>(smem) + (k.topo_smem_offset >> 4);

    Mono mo[MPT];
    int idx[MPT];
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int slot = threadIdx.x + t * blockDim.x;
        const int i = slot < a.n_active ? (int)a.amap[slot] : N;
        idx[t] = i;
        mo[t].flags = MF_EXTRA | MF_FIXED;
        if (i < N) {
            load_mono(a, base, i, mo[t]);
            if (kRngShared) {
                srng[i] = a.rng_xyz[base + i];
                srng[N + i] = a.rng_ang[base + i];
            } else {
                mo[t].rx = a.rng_xyz[base + i];
                mo[t].ra = a.rng_ang[base + i];
            }
            if (near.topo) near.topo[i] = load_topo(a, traj, i);
        }
    }
    if ((int)threadIdx.x < a.n_fixed) { // fixed monomers never move: published once, into both stage buffers
        const int i = (int)a.fmap[threadIdx.x];
        Mono fm;
        load_mono(a, base, i, fm);
        publish(stage_at(N, 0), i, fm, ls);
        if (k.nbuf == 2) publish(stage_at(N, 1), i, fm, ls);
    }

    CandState cs; // candidate-list state (persists in HBM between launches)
    cs.valid = near.cap > 0 ? a.cand_valid[traj] : 0;
    cs.dirty = false;

    int near_state = 0; // 0: not built, 1: valid, 2: overflowed (full list until the next rebuild)
    float gx[MPT], gy[MPT], gz[MPT]; // positions when the near list was formed (displacement guard)
    unsigned esc = 0;                 // bit t: monomer t tripped the guard and was escalated (no further watching until the next formation)
#pragma unroll
    for (int t = 0; t < MPT; t++) gx[t] = gy[t] = gz[t] = 0.f;
    const unsigned rops = (p.lj_on ? OP_REBUILD_LJ : 0u) | (p.is_assembly ? OP_REBUILD_BONDS : 0u);

    bool lazy = k.lazy != 0;              // list-update steps refresh only the near / bond lists (see KArgs::lazy)
    bool stale = a.lj_stale[traj] != 0;   // the Verlet list in HBM predates the last list-update step
    int fixed_flags_dirty = 0; // stage buffers whose copy of a fixed monomer's GTP bit is stale
    bool gtp_changed = false;
    // Step bookkeeping without 64-bit divisions in the loop: steps left until the next list-update step and until the next
    // scheduled hydrolysis event are counted down (the modulo of a long long costs ~30 instructions per warp and step).
    const int freq = p.ljpairsupdatefreq > 0 ? p.ljpairsupdatefreq : 1;
    int to_update = (int)((freq - k.first_step % freq) % freq); // 0: this step is a list-update step
    int sched_slot = 0;                                         // slot that becomes current at the next event
    int to_event = -1;                                          // steps until that event; < 0: none (left) in this launch
    if (k.sched_slots > 0) {
        const long long rel = k.first_step - k.sched_first;
        const long long slot0 = rel <= 0 ? 0 : (rel + k.sched_period - 1) / k.sched_period;
        const long long wait = k.sched_first + slot0 * k.sched_period - k.first_step;
        if (slot0 < k.sched_slots && wait < k.n_steps) {
            sched_slot = (int)slot0;
            to_event = (int)wait;
        }
    }
    const int period_m1 = k.sched_period - 1 > 0x7ffffff0LL ? 0x7ffffff0 : (int)(k.sched_period - 1);
    const int stage_toggle = k.nbuf == 2 ? 4 * N : 0;
    int boff = 0; // float4 offset of the current stage buffer (replaces buf * 4 * N)
    for (long long step = k.first_step; step < k.first_step + k.n_steps; step++) {
        Stage s;
        s.oP = boff;
        s.oE = boff + N;
        s.oL1 = boff + 2 * N;
        s.oL2 = boff + 3 * N;
        // scheduled hydrolysis event: the GTP flags uploaded for this step become current (maddy_schedule_gtp)
        const bool event_now = to_event == 0;
        if (to_event >= 0) to_event--;
        if (event_now) {
            const int slot = sched_slot++;
            to_event = sched_slot < k.sched_slots ? period_m1 : -1;
            {
                const uint8_t *g = a.gtp_sched + (size_t)slot * a.ntr * N + base;
#pragma unroll
                for (int t = 0; t < MPT; t++)
                    if (idx[t] < N) mo[t].flags = (mo[t].flags & ~MF_GTP) | (g[idx[t]] == 1 ? MF_GTP : 0);
                fixed_flags_dirty = k.nbuf; // fixed monomers are staged once: patch their flag word in each buffer in turn
                gtp_changed = true;
            }
        }
        if (fixed_flags_dirty > 0) {
            if ((int)threadIdx.x < a.n_fixed) {
                const int slot = sched_slot - 1; // latest slot at or before this step
                const uint8_t *g = a.gtp_sched + (size_t)slot * a.ntr * N + base;
                const int i = (int)a.fmap[threadIdx.x];
                int jf = __float_as_int(s.L2(i).w);
                jf = (jf & ~MF_GTP) | (g[i] == 1 ? MF_GTP : 0);
                s.L2(i).w = __int_as_float(jf);
            }
            fixed_flags_dirty--;
        }
        unsigned moved = 0; // bit t: monomer t of this thread has left its guard sphere (and has not been escalated yet)
        Frame fr[MPT];
#pragma unroll
        for (int t = 0; t < MPT; t++) {
            if (idx[t] < N) {
                fr[t] = publish(s, idx[t], mo[t], ls);
                const float dx = mo[t].x - gx[t], dy = mo[t].y - gy[t], dz = mo[t].z - gz[t];
                if (fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > MD_NEAR_GUARD2 && !(esc & (1u << t))) moved |= 1u << t;
            }
        }
        const bool any_moved = __syncthreads_or(moved != 0 && near_state == 1) != 0;
        const bool update_now = to_update == 0;
        to_update = update_now ? freq - 1 : to_update - 1;
        const bool do_rebuild = rops != 0 && update_now && !(step == k.first_step && (k.run_flags & MADDY_RUN_SKIP_FIRST_REBUILD));
        bool formed = false;
        if (do_rebuild) {
            bool lj_exact;
            near_state = rebuild_lists<MPT>(k, s, near, cs, traj, mo, idx, rops, true, lazy, lj_exact) ? 1 : 2;
            stale = !lj_exact;
            if (stale) { // the Verlet list of this step is (cand, these positions): written out only if someone asks
#pragma unroll
                for (int t = 0; t < MPT; t++)
                    if (idx[t] < N) a.rpos[base + idx[t]] = make_float4(mo[t].x, mo[t].y, mo[t].z, 0.f);
            }
            formed = true;
            if (near.topo) { // own rows were just rewritten by this thread
#pragma unroll
                for (int t = 0; t < MPT; t++)
                    if (idx[t] < N) near.topo[idx[t]] = load_topo(a, traj, idx[t]);
            }
        } else if (near.cap > 0 && p.lj_on && near_state == 1 && any_moved) {
            // Displacement guard tripped: only pairs of the monomers that left their guard sphere can be missing from a
            // near row.  Each of them - and every partner in its Verlet row - is ESCALATED to walking its exact Verlet row
            // (written out first where the list is lazy) until the next list-update step; everybody else keeps the near row,
            // whose pairs are still covered by their own guards.  A monomer that tripped needs no further watching (all its
            // listed pairs are exact now).  Same forces as a whole-trajectory refresh, at the cost of ~50 monomers instead
            // of the trajectory - free dimers that the reference's insertion rule drops onto one another trip at every step.
            if (stale) __syncthreads(); // rpos of the whole trajectory is in place
#pragma unroll
            for (int t = 0; t < MPT; t++) {
                if (!(moved & (1u << t))) continue;
                const int i = idx[t];
                esc |= 1u << t;
                if (stale) materialise_row(k, traj, i);
                const uint16_t *lj = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
                const int n = a.ljcnt[(size_t)traj * a.Npad + i];
                for (int kk = 0; kk < n; kk++) near.cnt[lj[(size_t)kk * a.Npad]] = MD_NEAR_PENDING;
                near.cnt[i] = MD_NEAR_FULL_EXACT;
            }
            __syncthreads();
#pragma unroll
            for (int t = 0; t < MPT; t++) {
                const int i = idx[t];
                if (i < N && near.cnt[i] == MD_NEAR_PENDING) {
                    if (stale) materialise_row(k, traj, i);
                    near.cnt[i] = MD_NEAR_FULL_EXACT;
                }
            }
            if (threadIdx.x == 0) atomicAdd(a.stats + 0, 1ull);
        } else if (near.cap > 0 && p.lj_on && near_state == 0) {
            if (stale) { // the refresh below reads the Verlet list: write it out now and stay exact for the rest of the launch
                __syncthreads(); // rpos of the whole trajectory is in place
#pragma unroll
                for (int t = 0; t < MPT; t++)
                    if (idx[t] < N) materialise_row(k, traj, idx[t]);
                if ((int)threadIdx.x < a.n_fixed) materialise_row(k, traj, (int)a.fmap[threadIdx.x]);
                stale = false;
                lazy = false;
            }
            if (refresh_near<MPT>(k, s, near, traj, mo, idx)) atomicAdd(a.stats + 3, 1ull);
            __syncthreads();
            near_state = 1;
            formed = true;
        }
        if (formed) {
            esc = 0;
#pragma unroll
            for (int t = 0; t < MPT; t++) {
                gx[t] = mo[t].x;
                gy[t] = mo[t].y;
                gz[t] = mo[t].z;
            }
        }
        near.ok = near.cap > 0 && near_state == 1;
        near.stale_lj = stale;
#pragma unroll
        for (int t = 0; t < MPT; t++) {
            if (idx[t] < N && !(mo[t].flags & (MF_EXTRA | MF_FIXED))) {
                const G6 f = monomer_force(k, s, near, traj, idx[t], mo[t], fr[t]);
                if (kRngShared) {
                    mo[t].rx = srng[idx[t]];
                    mo[t].ra = srng[N + idx[t]];
                }
                integrate_monomer(k, mo[t], f);
                if (kRngShared) {
                    srng[idx[t]] = mo[t].rx;
                    srng[N + idx[t]] = mo[t].ra;
                }
            }
        }
        boff ^= stage_toggle;
        if (k.nbuf != 2) __syncthreads();
    }
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = idx[t];
        if (i < N) {
            a.pos[base + i] = make_float4(mo[t].x, mo[t].y, mo[t].z, 0.f);
            a.ang[base + i] = make_float4(mo[t].fi, mo[t].psi, mo[t].theta, 0.f);
            a.rng_xyz[base + i] = kRngShared ? srng[i] : mo[t].rx;
            a.rng_ang[base + i] = kRngShared ? srng[N + i] : mo[t].ra;
            if (gtp_changed) a.gtp[base + i] = (mo[t].flags & MF_GTP) ? 1 : 0;
        }
    }
    if (gtp_changed && (int)threadIdx.x < a.n_fixed) { // fixed monomers: last slot applied in this launch
        const long long last = k.first_step + k.n_steps - 1;
        long long slot = (last - k.sched_first) / k.sched_period;
        if (slot >= k.sched_slots) slot = k.sched_slots - 1;
        const int i = (int)a.fmap[threadIdx.x];
        a.gtp[base + i] = a.gtp_sched[(size_t)slot * a.ntr * N + base + i] == 1 ? 1 : 0;
    }
    store_cand_state(a, cs, traj);
    if (threadIdx.x == 0) a.lj_stale[traj] = stale ? 1 : 0;
}

// integrateTea_prepare (bdhitea_kernel.cu:16-36) for one bead, fused behind its force evaluation: the same statements as
// tea_prepare_kernel (maddy_tea.cu) on the same values (products only: no contraction either way), so the TEA window of
// maddy_run and the step-granular calls stay bit-identical.
__device__ __forceinline__ void tea_prepare_bead(const KArgs &k, size_t q, const Mono &m, const G6 &f, bool extra)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const float var = sqrtf(2.0f * KB_BOLTZ * p.Temp * p.gammaR / p.dt);
    uint4 st = a.rng_xyz[q];
    float4 df = rforce(st); // every bead draws, fixed and reserve ones included (:22)
    a.rng_xyz[q] = st;
    df.x *= var;
    df.y *= var;
    df.z *= var;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    a.tea_rf[q] = extra ? zero : df;
    a.tea_mf[q] = extra ? zero : make_float4(f.x, f.y, f.z, 0.f);
    a.tea_co[q] = make_float4(m.x, m.y, m.z, extra ? 1.f : 0.f);
    a.fpos[q] = zero; // only xyz is zeroed (:31-33)
}

/*
 * phase_kernel<MPT> — one phase of the step for EVERY monomer (the step-granular entry points: one call per
 * reference launch).  Same device functions as the fused loop, lists read from HBM, so results are bit-identical.
 */
template <int MPT>
__global__ void __launch_bounds__(MD_MAX_THREADS, 1) phase_kernel(const __grid_constant__ KArgs k)
{
    float4 *const smem = g_smem;
    __shared__ double red_scratch[32 * 7];
    const DevSys &a = k.a;
    const int N = a.N;
    const int traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    const LatSite ls = lateral_site();
    Near near = carve_near(smem, k, N);

    Mono mo[MPT];
    int idx[MPT];
#pragma unroll
    for (int t = 0; t < MPT; t++) {
        const int i = threadIdx.x + t * blockDim.x;
        idx[t] = i;
        mo[t].flags = MF_EXTRA | MF_FIXED;
        if (i < N) load_mono(a, base, i, mo[t]);
    }
    const Stage s = stage_at(N, 0);
    Frame fr[MPT];
#pragma unroll
    for (int t = 0; t < MPT; t++)
        if (idx[t] < N) fr[t] = publish(s, idx[t], mo[t], ls);
    __syncthreads();

    if ((k.ops & OP_MATERIALISE) && a.lj_stale[traj]) { // CTA-uniform
#pragma unroll
        for (int t = 0; t < MPT; t++)
            if (idx[t] < N) materialise_row(k, traj, idx[t]);
        __syncthreads(); // every thread has read the flag
        if (threadIdx.x == 0) a.lj_stale[traj] = 0;
    }
    if (k.ops & (OP_REBUILD_LJ | OP_REBUILD_BONDS)) {
        CandState cs;
        cs.valid = near.cap > 0 ? a.cand_valid[traj] : 0;
        cs.dirty = false;
        bool lj_exact;
        rebuild_lists<MPT>(k, s, near, cs, traj, mo, idx, k.ops, false, false, lj_exact);
        store_cand_state(a, cs, traj);
        if ((k.ops & OP_REBUILD_LJ) && threadIdx.x == 0) a.lj_stale[traj] = 0; // the list written above is current
    }
    near.ok = false;

    if (k.ops & OP_FORCE) {
#pragma unroll
        for (int t = 0; t < MPT; t++) {
            const int i = idx[t];
            if (i >= N) continue;
            const bool extra = (mo[t].flags & MF_EXTRA) != 0;
            G6 f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (!extra) {
                // extras keep the zero written by the integrator (compute_cuda.cu:55, :966-972)
                f = monomer_force<true>(k, s, near, traj, i, mo[t], fr[t]);
                if (!(k.ops & OP_TEA_PREP)) a.fpos[base + i] = make_float4(f.x, f.y, f.z, 0.f);
                a.fang[base + i] = make_float4(f.fi, f.psi, f.theta, 0.f);
            }
            if (k.ops & OP_TEA_PREP) tea_prepare_bead(k, base + i, mo[t], f, extra);
        }
    }

    if (k.ops & OP_ENERGY) {
        E7 acc = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
        for (int t = 0; t < MPT; t++) {
            const int i = idx[t];
            if (i < N) {
                const E7 e = monomer_energy(k, s, traj, i, mo[t]);
                double *o = a.en_mono + (base + i) * 7;
                o[0] = e.harm; o[1] = e.lng; o[2] = e.lat; o[3] = e.psi; o[4] = e.fi; o[5] = e.teta; o[6] = e.lj;
                acc.harm += e.harm; acc.lng += e.lng; acc.lat += e.lat; acc.psi += e.psi;
                acc.fi += e.fi; acc.teta += e.teta; acc.lj += e.lj;
            }
        }
        // per-trajectory order for the host: harm,long,lat,psi,fi,teta,lj (updater.cpp:35-36)
        block_reduce_e7(acc, a.en_traj + (size_t)traj * 7, red_scratch);
    }
}

// Evaluates StepConsts with the device's fast-math arithmetic (the expressions of compute_cuda.cu:955-961, :448).
__global__ void consts_kernel(maddy_params p, StepConsts *out)
{
    StepConsts c;
    c.aR = p.dt / p.gammaR;
    c.aA = p.dt / (p.gammaTheta * p.alpha);
    c.aT = p.dt / p.gammaTheta;
    c.vA = p.varTheta * sqrtf(p.freeze_temp / p.alpha);
    c.D_lat_seam = p.D_lat / p.seam_coeff;
    *out = c;
}
cudaError_t launch_consts_kernel(const maddy_params &p, StepConsts *d_out, cudaStream_t st)
{
    consts_kernel<<<1, 1, 0, st>>>(p, d_out);
    return cudaGetLastError();
}

// Stand-alone integrator over forces stored in HBM (step-granular maddy_integrate).
__global__ void __launch_bounds__(256) integrate_kernel(const __grid_constant__ KArgs k)
{
    const DevSys &a = k.a;
    const size_t n = (size_t)a.ntr * a.N;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % a.N);
        Mono m;
        m.flags = (int)a.sflags[i] | (a.extra[q] ? MF_EXTRA : 0);
        if (!(m.flags & MF_FIXED) && !(m.flags & MF_EXTRA)) {
            const float4 P = a.pos[q], A = a.ang[q], FP = a.fpos[q], FA = a.fang[q];
            m.x = P.x; m.y = P.y; m.z = P.z; m.fi = A.x; m.psi = A.y; m.theta = A.z;
            m.rx = a.rng_xyz[q];
            m.ra = a.rng_ang[q];
            G6 f = {FP.x, FP.y, FP.z, FA.x, FA.y, FA.z};
            integrate_monomer(k, m, f);
            a.pos[q] = make_float4(m.x, m.y, m.z, 0.f);
            a.ang[q] = make_float4(m.fi, m.psi, m.theta, 0.f);
            a.rng_xyz[q] = m.rx;
            a.rng_ang[q] = m.ra;
        }
        a.fpos[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        a.fang[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// ------------------------------------------------------------------ launch helpers (called from the C-ABI)
template <int MPT, int MINB>
static cudaError_t launch_run(const KArgs &k, int threads, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(run_kernel<MPT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    run_kernel<MPT, MINB><<<k.a.ntr, threads, smem, st>>>(k);
    return cudaGetLastError();
}
template <int MPT>
static cudaError_t launch_phase(const KArgs &k, int threads, size_t smem, cudaStream_t st)
{
    cudaError_t e = cudaFuncSetAttribute(phase_kernel<MPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    phase_kernel<MPT><<<k.a.ntr, threads, smem, st>>>(k);
    return cudaGetLastError();
}

// fused loop; ctas_per_sm = 2 only with mpt == 1
cudaError_t launch_run_kernel(const KArgs &k, int mpt, int ctas_per_sm, int threads, size_t smem, cudaStream_t st)
{
    if (ctas_per_sm == 2) return mpt == 1 ? launch_run<1, 2>(k, threads, smem, st) : cudaErrorInvalidValue;
    switch (mpt) {
    case 1: return launch_run<1, 1>(k, threads, smem, st);
    case 2: return launch_run<2, 1>(k, threads, smem, st);
    case 3: return launch_run<3, 1>(k, threads, smem, st);
    case 4: return launch_run<4, 1>(k, threads, smem, st);
    case 5: return launch_run<5, 1>(k, threads, smem, st);
    case 6: return launch_run<6, 1>(k, threads, smem, st);
    case 7: return launch_run<7, 1>(k, threads, smem, st);
    default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_phase_kernel(const KArgs &k, int mpt, int threads, size_t smem, cudaStream_t st)
{
    switch (mpt) {
    case 1: return launch_phase<1>(k, threads, smem, st);
    case 2: return launch_phase<2>(k, threads, smem, st);
    case 3: return launch_phase<3>(k, threads, smem, st);
    case 4: return launch_phase<4>(k, threads, smem, st);
    case 5: return launch_phase<5>(k, threads, smem, st);
    case 6: return launch_phase<6>(k, threads, smem, st);
    default: return cudaErrorInvalidValue;
    }
}

} // namespace maddy
#include "maddy_wide.cuh"
namespace maddy {

// State read-back, device leg: float4 {x,y,z,-} + {fi,psi,theta,-} -> the boundary's AoS-7 record {x,y,z,fi,theta,psi,0}
// (mt.h:63-71) in a device staging buffer; the PCIe leg is a plain copy on a second stream (maddy_snapshot_begin), so the
// host receives the layout it hands to the DCD writer and does no transposition of its own.
__global__ void __launch_bounds__(256) snapshot_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ ang, float *__restrict__ out,
                                                       size_t n_monomers)
{
    // 256 monomers per tile: coalesced float4 loads -> AoS-7 in shared memory -> contiguous float4 stores
    // (a tile is 256 * 28 B = 7168 B, a multiple of 16, so every tile starts 16-byte aligned)
    __shared__ __align__(16) float tile[256 * 7];
    const size_t ntiles = (n_monomers + 255) / 256;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const size_t q = t * 256 + threadIdx.x;
        if (q < n_monomers) {
            const float4 p = __ldg(pos + q), a = __ldg(ang + q);
            float *r = tile + threadIdx.x * 7;
            r[0] = p.x; r[1] = p.y; r[2] = p.z;
            r[3] = a.x; r[4] = a.z; r[5] = a.y; // fi, theta, psi
            r[6] = 0.f;
        }
        __syncthreads();
        const size_t left = n_monomers - t * 256;
        const int nfl = (int)(left < 256 ? left : 256) * 7;
        float *dst = out + t * 256 * 7;
        const int n4 = nfl / 4;
        for (int g = threadIdx.x; g < n4; g += 256) reinterpret_cast<float4 *>(dst)[g] = reinterpret_cast<const float4 *>(tile)[g];
        if ((int)threadIdx.x < (nfl & 3)) dst[4 * n4 + threadIdx.x] = tile[4 * n4 + threadIdx.x];
        __syncthreads();
    }
}
cudaError_t launch_snapshot_kernel(const float4 *pos, const float4 *ang, float *out_mapped, size_t n_monomers, cudaStream_t st)
{
    int blocks = (int)((n_monomers + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    snapshot_kernel<<<blocks, 256, 0, st>>>(pos, ang, out_mapped, n_monomers);
    return cudaGetLastError();
}

cudaError_t launch_integrate_kernel(const KArgs &k, cudaStream_t st)
{
    const size_t n = (size_t)k.a.ntr * k.a.N;
    int blocks = (int)((n + 255) / 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    integrate_kernel<<<blocks, 256, 0, st>>>(k);
    return cudaGetLastError();
}

} // namespace maddy
