/*
 * maddy_tea.cu — TEA hydrodynamic-interaction integrator (Geyer & Winter 2009,
 * doi:10.1063/1.3089668) with the Rotne-Prager-Yamakawa tensor, all-pairs ("unlisted").
 *
 * What the reference does (src/bdhitea_kernel.cu:16-213, src/bdhitea.cu:37-118): one thread per bead walks
 * all N partners sequentially with every term gathered from global memory (3 float4 per pair), plus a D2H of
 * per-bead epsilon sums, a serial host loop for beta and an H2D every tea_epsilon_freq steps, and a host loop
 * over all particles every step.
 *
 * Here the O(N^2) work is spread over the whole GPU even for ONE trajectory.  The epsilon statistics (every
 * tea_epsilon_freq steps) use a warp per bead, lanes striding over the partners; the pair kernel of every step gives a
 * warp four beads as two packed pairs (FFMA2 / FMUL2 / FADD2), streams the partners through shared memory with cp.async
 * and keeps two rounds of 32 partners in flight (see tea_pair_kernel).  All reductions are shuffles and shared-memory
 * sums in a fixed order (deterministic, independent of the launch shape).  epsilon is reduced per trajectory on the
 * device and beta (eq. 26) is evaluated there too, so nothing crosses PCIe.  The pair work is a generated 3x3 mat-vec
 * per (i,j) whose right-hand side (f_j + C_i o r_j) depends on i: not a GEMM, FP32-pipe bound; tensor cores do not apply.
 *
 * Summation order differs from the reference's sequential j loop (lane-strided partial sums + butterfly), so
 * TEA displacements agree with the reference to float rounding (~1e-7 relative), not bit for bit.
 */
#include "maddy_kernels.cuh"

namespace maddy {

#define KB_BOLTZ 0.0019872041f // kcal/(mol*K), mt.h:39
#define ST_TEA_ABORT 0x100
#define TEA_WARPS 8 // beads (warps) per CTA

struct Sym6 { float xx, xy, xz, yy, yz, zz; };

// D_ij / D_ii for a unit vector (x,y,z) and distance w (bdhitea_kernel.cu:38-58, eq. 3-5)
__device__ __forceinline__ Sym6 rpy(float x, float y, float z, float w, float a)
{
    const float ra = w / a;
    float crr, cii;
    if (ra > 2.f) {
        crr = 0.75f / ra * (1.f - 2.f / ra / ra);
        cii = 0.75f / ra * (1.f + 2.f / 3.f / ra / ra);
    } else {
        crr = 3.f * ra / 32.f;
        cii = 1.f - 9.f * ra / 32.f;
    }
    Sym6 d;
    d.xx = x * x * crr + cii;
    d.xy = x * y * crr;
    d.xz = x * z * crr;
    d.yy = y * y * crr + cii;
    d.yz = y * z * crr;
    d.zz = z * z * crr + cii;
    return d;
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- snapshot of coordinates (+extra flag in .w) for the epsilon pass
__global__ void __launch_bounds__(256) tea_snapshot_kernel(const __grid_constant__ KArgs k)
{
    const DevSys &a = k.a;
    const size_t n = (size_t)a.ntr * a.N;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const float4 P = a.pos[q];
        a.tea_co[q] = make_float4(P.x, P.y, P.z, a.extra[q] ? 1.f : 0.f);
    }
}

// ---- per-bead epsilon / C_i statistics (integrateTea_epsilon_unlisted, bdhitea_kernel.cu:61-101): warp per bead
__global__ void __launch_bounds__(TEA_WARPS * 32) tea_epsilon_kernel(const __grid_constant__ KArgs k)
{
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.y;
    const int i = blockIdx.x * TEA_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const size_t base = (size_t)traj * N;
    const float4 *co = a.tea_co + base;
    const float4 ci = co[i];
    float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
    if (ci.w == 0.f) {
        for (int j = lane; j < N; j += 32) {
            const float4 cj = co[j];
            if (j == i || cj.w != 0.f) continue;
            float dx = cj.x - ci.x, dy = cj.y - ci.y, dz = cj.z - ci.z;
            const float w = sqrtf(dx * dx + dy * dy + dz * dz);
            dx /= w;
            dy /= w;
            dz /= w;
            const Sym6 d = rpy(dx, dy, dz, w, k.p.tea_a);
            sw += d.xx + 2 * d.xy + 2 * d.xz + d.yy + 2 * d.yz + d.zz;
            sx += d.xx * d.xx + d.xy * d.xy + d.xz * d.xz;
            sy += d.xy * d.xy + d.yy * d.yy + d.yz * d.yz;
            sz += d.xz * d.xz + d.yz * d.yz + d.zz * d.zz;
        }
    }
    sx = warp_sum(sx);
    sy = warp_sum(sy);
    sz = warp_sum(sz);
    sw = warp_sum(sw);
    if (lane == 0) {
        a.tea_ci[base + i] = make_float4(sx, sy, sz, 0.f);
        a.tea_eps[base + i] = sw;
    }
}

// ---- per-trajectory epsilon and beta (host part of updateTea, bdhitea.cu:57-118): one CTA per trajectory
__global__ void __launch_bounds__(256) tea_beta_kernel(const __grid_constant__ KArgs k)
{
    __shared__ double red[8];
    __shared__ int redn[8];
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    double e = 0.0;
    int n = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        e += (double)a.tea_eps[base + i];
        n += a.extra[base + i] ? 0 : 1;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        e += __shfl_down_sync(0xffffffffu, e, o);
        n += __shfl_down_sync(0xffffffffu, n, o);
    }
    if (lane == 0) {
        red[warp] = e;
        redn[warp] = n;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            e += red[w];
            n += redn[w];
        }
        const double n3 = 3. * n;
        double eps = e / (n3 * (n3 - 3.));
        bool bad = false;
        if (eps > 1.0) {
            if (k.p.tea_capricious) bad = true;
            eps = 1.0;
        }
        if (eps > (double)k.p.tea_epsmax) bad = true;
        const double aa = (n3 - 1.) * eps * eps - (n3 - 2.) * eps;
        float beta;
        if (fabs(aa) < 1e-7) {
            beta = .5f;
            if (k.p.tea_capricious && k.p.tea_a > 0.0f) bad = true;
        } else {
            beta = (float)((1. - sqrt(1. - aa)) / aa);
        }
        a.tea_beta[traj] = beta;
        if (bad) atomicOr(a.status, ST_TEA_ABORT);
    }
}

// ---- integrateTea_prepare (bdhitea_kernel.cu:16-36): every bead draws, fixed and extra included
__global__ void __launch_bounds__(256) tea_prepare_kernel(const __grid_constant__ KArgs k)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const size_t n = (size_t)a.ntr * a.N;
    const float var = sqrtf(2.0f * KB_BOLTZ * p.Temp * p.gammaR / p.dt);
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        uint4 st = a.rng_xyz[q];
        float4 df = rforce(st);
        a.rng_xyz[q] = st;
        df.x *= var;
        df.y *= var;
        df.z *= var;
        const float4 F = a.fpos[q], P = a.pos[q];
        // reserve beads act on nobody (bdhitea_kernel.cu:166): their records carry zero force, so the pair kernel needs no test
        const bool extra = a.extra[q] != 0;
        a.tea_rf[q] = extra ? make_float4(0.f, 0.f, 0.f, 0.f) : df;
        a.tea_mf[q] = extra ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(F.x, F.y, F.z, 0.f);
        a.tea_co[q] = make_float4(P.x, P.y, P.z, extra ? 1.f : 0.f);
        a.fpos[q] = make_float4(0.f, 0.f, 0.f, 0.f); // only xyz is zeroed (:31-33)
    }
}

// ---- integrateTea_kernel_unlisted (bdhitea_kernel.cu:148-213)
// Work split: a group of TEA_IB = 4 beads is shared by TEA_S = 4 warps ("parts"); part s takes every 4th round of 32
// partners, so even ONE trajectory of a few thousand beads puts ~8 warps on every SM sub-partition.  A CTA holds TEA_G
// groups and streams the partners through shared memory in tiles of 512 (two-stage cp.async ring).  The partial sums of the parts are combined in a fixed order ((p0+p1)+(p2+p3), then a lane butterfly), so
// the result does not depend on the launch shape.
// Arithmetic: the four beads of a group are two PACKED pairs: sm_100a's FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2) do the same
// operation for both beads of a pair in one instruction - measured 66 TFLOP/s against 42 TFLOP/s for three-register
// scalar FFMA on this GPU (tools/micro/ffma2_bench.cu).  A partner's component enters as a scalar register broadcast to
// both halves (the .F32 operand form), so the tile holds the plain records.  Per pair the tensor is applied in its dyadic form  D g = cii g + crr (u.g) u
// (eq. 3-5 of the paper; the reference multiplies the six entries out, bdhitea_kernel.cu:38-58) with one MUFU (rsqrt):
// 1/ra = a / w needs no division.  Reserve beads carry zero force records (exact zero contribution); the overlap
// branch ra <= 2, the self pair and the padding beyond N are handled in a warp-uniform slow path.
#define TEA_IB 4
#define TEA_THREADS 256
#define TEA_TILE 512
#define TEA_STAGES 2
#define TEA_SPLIT_NTOT 2048 // trajectories at least this long split the partners of a bead group over 4 warps

struct TeaBeads { // per warp: two packed bead pairs
    float2 npx[2], npy[2], npz[2]; // negated positions
    float2 cx[2], cy[2], cz[2];    // beta * C_i
    float2 ax[2], ay[2], az[2];    // accumulators
};

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

template <bool MASKED>
__device__ __forceinline__ void tea_round(TeaBeads &B, const float4 *sco, const float4 *smf, const float4 *srf, int jj, int j, int i0, int N, float ta,
                                          float inv_a, float near2)
{
    const float4 c = sco[jj], m = smf[jj], r = srf[jj];
#pragma unroll
    for (int pk = 0; pk < 2; pk++) {
        // (v, v) operands: FADD2 / FFMA2 take a scalar register broadcast to both halves, no move needed
        const float2 dx = __fadd2_rn(f2(c.x, c.x), B.npx[pk]);
        const float2 dy = __fadd2_rn(f2(c.y, c.y), B.npy[pk]);
        const float2 dz = __fadd2_rn(f2(c.z, c.z), B.npz[pk]);
        float2 w2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
        bool on0 = true, on1 = true;
        if (MASKED) { // the self pair and the padding beyond N
            on0 = j != i0 + 2 * pk && j < N;
            on1 = j != i0 + 2 * pk + 1 && j < N;
            w2.x = on0 ? w2.x : 1.f;
            w2.y = on1 ? w2.y : 1.f;
        }
        const float2 iw = f2(rsqrtf(w2.x), rsqrtf(w2.y));
        const float2 ux = __fmul2_rn(dx, iw), uy = __fmul2_rn(dy, iw), uz = __fmul2_rn(dz, iw);
        const float2 ira = __fmul2_rn(f2(ta, ta), iw); // 1 / ra
        const float2 ira2 = __fmul2_rn(ira, ira);
        const float2 far = __fmul2_rn(f2(0.75f, 0.75f), ira);
        float2 crr = __fmul2_rn(far, __ffma2_rn(f2(-2.f, -2.f), ira2, f2(1.f, 1.f)));
        float2 cii = __fmul2_rn(far, __ffma2_rn(f2(2.f / 3.f, 2.f / 3.f), ira2, f2(1.f, 1.f)));
        const bool near = w2.x <= near2 || w2.y <= near2;
        if (MASKED || __any_sync(0xffffffffu, near)) { // overlapping beads (ra <= 2) and excluded pairs: rare, scalar
            if (w2.x <= near2) {
                const float ra = (w2.x * iw.x) * inv_a;
                crr.x = (3.f / 32.f) * ra;
                cii.x = 1.f - (9.f / 32.f) * ra;
            }
            if (w2.y <= near2) {
                const float ra = (w2.y * iw.y) * inv_a;
                crr.y = (3.f / 32.f) * ra;
                cii.y = 1.f - (9.f / 32.f) * ra;
            }
            if (!on0) crr.x = cii.x = 0.f;
            if (!on1) crr.y = cii.y = 0.f;
        }
        const float2 gx = __ffma2_rn(f2(r.x, r.x), B.cx[pk], f2(m.x, m.x));
        const float2 gy = __ffma2_rn(f2(r.y, r.y), B.cy[pk], f2(m.y, m.y));
        const float2 gz = __ffma2_rn(f2(r.z, r.z), B.cz[pk], f2(m.z, m.z));
        const float2 pr = __fmul2_rn(crr, __ffma2_rn(uz, gz, __ffma2_rn(uy, gy, __fmul2_rn(ux, gx))));
        B.ax[pk] = __ffma2_rn(cii, gx, __ffma2_rn(pr, ux, B.ax[pk]));
        B.ay[pk] = __ffma2_rn(cii, gy, __ffma2_rn(pr, uy, B.ay[pk]));
        B.az[pk] = __ffma2_rn(cii, gz, __ffma2_rn(pr, uz, B.az[pk]));
    }
}

// Two rounds of 32 partners at once, common case only (no self pair, no padding): the four bead-pair streams are
// independent until the accumulators, which doubles the instruction-level parallelism a warp offers (the top stall of
// the one-round version was `wait`: dependent packed operations).  Accumulation order = round A, then round B, as when
// the rounds are taken one by one, so the result is bit-identical.
__device__ __forceinline__ void tea_round2(TeaBeads &B, const float4 *sco, const float4 *smf, const float4 *srf, int jjA, int jjB, float ta, float inv_a,
                                           float near2)
{
    const float4 c[2] = {sco[jjA], sco[jjB]}, m[2] = {smf[jjA], smf[jjB]}, r[2] = {srf[jjA], srf[jjB]};
    float2 ux[4], uy[4], uz[4], crr[4], cii[4], w2[4], iw[4];
    bool near = false;
#pragma unroll
    for (int q = 0; q < 4; q++) { // q = 2 * round + pair
        const int rd = q >> 1, pk = q & 1;
        const float2 dx = __fadd2_rn(f2(c[rd].x, c[rd].x), B.npx[pk]);
        const float2 dy = __fadd2_rn(f2(c[rd].y, c[rd].y), B.npy[pk]);
        const float2 dz = __fadd2_rn(f2(c[rd].z, c[rd].z), B.npz[pk]);
        w2[q] = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
        iw[q] = f2(rsqrtf(w2[q].x), rsqrtf(w2[q].y));
        ux[q] = __fmul2_rn(dx, iw[q]);
        uy[q] = __fmul2_rn(dy, iw[q]);
        uz[q] = __fmul2_rn(dz, iw[q]);
        const float2 ira = __fmul2_rn(f2(ta, ta), iw[q]);
        const float2 ira2 = __fmul2_rn(ira, ira);
        const float2 far = __fmul2_rn(f2(0.75f, 0.75f), ira);
        crr[q] = __fmul2_rn(far, __ffma2_rn(f2(-2.f, -2.f), ira2, f2(1.f, 1.f)));
        cii[q] = __fmul2_rn(far, __ffma2_rn(f2(2.f / 3.f, 2.f / 3.f), ira2, f2(1.f, 1.f)));
        near |= fminf(w2[q].x, w2[q].y) <= near2;
    }
    if (__any_sync(0xffffffffu, near)) { // overlapping beads (ra <= 2): rare, scalar
#pragma unroll
        for (int q = 0; q < 4; q++) {
            if (w2[q].x <= near2) {
                const float ra = (w2[q].x * iw[q].x) * inv_a;
                crr[q].x = (3.f / 32.f) * ra;
                cii[q].x = 1.f - (9.f / 32.f) * ra;
            }
            if (w2[q].y <= near2) {
                const float ra = (w2[q].y * iw[q].y) * inv_a;
                crr[q].y = (3.f / 32.f) * ra;
                cii[q].y = 1.f - (9.f / 32.f) * ra;
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int rd = q >> 1, pk = q & 1;
        const float2 gx = __ffma2_rn(f2(r[rd].x, r[rd].x), B.cx[pk], f2(m[rd].x, m[rd].x));
        const float2 gy = __ffma2_rn(f2(r[rd].y, r[rd].y), B.cy[pk], f2(m[rd].y, m[rd].y));
        const float2 gz = __ffma2_rn(f2(r[rd].z, r[rd].z), B.cz[pk], f2(m[rd].z, m[rd].z));
        const float2 pr = __fmul2_rn(crr[q], __ffma2_rn(uz[q], gz, __ffma2_rn(uy[q], gy, __fmul2_rn(ux[q], gx))));
        B.ax[pk] = __ffma2_rn(cii[q], gx, __ffma2_rn(pr, ux[q], B.ax[pk]));
        B.ay[pk] = __ffma2_rn(cii[q], gy, __ffma2_rn(pr, uy[q], B.ay[pk]));
        B.az[pk] = __ffma2_rn(cii[q], gz, __ffma2_rn(pr, uz[q], B.az[pk]));
    }
}

// 16-byte asynchronous copy global -> shared (LDGSTS); src_bytes = 0 fills zeros
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem, int src_bytes)
{
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

// TEA_S parts per bead group, TEA_G groups per CTA (TEA_S * TEA_G = 8 warps).  <4, 2> for long trajectories, <1, 8> for
// short ones (their ensembles have warps to spare, and a part would only see a handful of rounds).  The choice depends
// on N alone, so a trajectory's result does not depend on how many trajectories share the launch.
template <int TEA_S, int TEA_G>
__global__ void __launch_bounds__(TEA_THREADS, 2) tea_pair_kernel(const __grid_constant__ KArgs k)
{
    constexpr int TEA_ROUNDS = TEA_TILE / 32 / TEA_S;
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / TEA_S, part = warp % TEA_S;
    const int i0 = (blockIdx.x * TEA_G + grp) * TEA_IB; // may be >= N: the warp still loads tiles and meets the barriers
    const size_t base = (size_t)traj * N;
    const float4 *co = a.tea_co + base, *mf = a.tea_mf + base, *rf = a.tea_rf + base;
    // partner tiles {position, molecular force, random force}: TEA_STAGES-deep ring filled with cp.async
    __shared__ float4 tile[TEA_STAGES][3][TEA_TILE];

    const float beta = a.tea_beta[traj];
    const float b2 = beta * beta;
    TeaBeads B;
#pragma unroll
    for (int pk = 0; pk < 2; pk++) {
        float q[2][6];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int i = min(i0 + 2 * pk + h, N - 1);
            const float4 raw = a.tea_ci[base + i], ci = co[i];
            q[h][0] = -ci.x; q[h][1] = -ci.y; q[h][2] = -ci.z;
            q[h][3] = beta * (1.f / sqrtf(1.f + b2 * raw.x));
            q[h][4] = beta * (1.f / sqrtf(1.f + b2 * raw.y));
            q[h][5] = beta * (1.f / sqrtf(1.f + b2 * raw.z));
        }
        B.npx[pk] = f2(q[0][0], q[1][0]); B.npy[pk] = f2(q[0][1], q[1][1]); B.npz[pk] = f2(q[0][2], q[1][2]);
        B.cx[pk] = f2(q[0][3], q[1][3]); B.cy[pk] = f2(q[0][4], q[1][4]); B.cz[pk] = f2(q[0][5], q[1][5]);
        B.ax[pk] = B.ay[pk] = B.az[pk] = f2(0.f, 0.f);
    }
    const float ta = p.tea_a, inv_a = 1.f / p.tea_a, near2 = 4.f * p.tea_a * p.tea_a;

    const int ntiles = (N + TEA_TILE - 1) / TEA_TILE;
    auto fetch = [&](int t) { // tile t -> stage t % TEA_STAGES; TEA_TILE / TEA_THREADS partners per thread
        if (t < ntiles) {
            float4(*T)[TEA_TILE] = tile[t % TEA_STAGES];
#pragma unroll
            for (int u = 0; u < TEA_TILE / TEA_THREADS; u++) {
                const int jj = u * TEA_THREADS + tid, j = t * TEA_TILE + jj, jc = min(j, N - 1), nb = j < N ? 16 : 0;
                cp_async16(&T[0][jj], co + jc, nb);
                cp_async16(&T[1][jj], mf + jc, nb);
                cp_async16(&T[2][jj], rf + jc, nb);
            }
        }
        cp_async_commit();
    };
#pragma unroll
    for (int t = 0; t < TEA_STAGES - 1; t++) fetch(t);
    for (int t = 0; t < ntiles; t++) {
        cp_async_wait<TEA_STAGES - 2>(); // this thread's part of tile t has landed
        __syncthreads();                 // everybody's has; and everybody is done with the stage refilled below
        fetch(t + TEA_STAGES - 1);
        const float4(*T)[TEA_TILE] = tile[t % TEA_STAGES];
        auto special = [&](int j0) { return (j0 <= i0 + TEA_IB - 1 && i0 <= j0 + 31) || j0 + 31 >= N; }; // self pair or padding inside
        auto one = [&](int jj0, int j0) {
            if (special(j0)) tea_round<true>(B, T[0], T[1], T[2], jj0 + lane, j0 + lane, i0, N, ta, inv_a, near2);
            else tea_round<false>(B, T[0], T[1], T[2], jj0 + lane, j0 + lane, i0, N, ta, inv_a, near2);
        };
#pragma unroll 1
        for (int r = 0; r < TEA_ROUNDS; r += 2) {
            const int jjA = (part + TEA_S * r) * 32, jA = t * TEA_TILE + jjA;
            const int jjB = jjA + TEA_S * 32, jB = jA + TEA_S * 32;
            if (jA >= N) break;
            if (jB < N && !special(jA) && !special(jB)) tea_round2(B, T[0], T[1], T[2], jjA + lane, jjB + lane, ta, inv_a, near2);
            else {
                one(jjA, jA);
                if (jB < N) one(jjB, jB);
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
    // combine the parts of a group in a fixed order (the tile buffers are free now)
    static_assert(TEA_S == 1 || TEA_S == 4, "parts per bead group");
    float sum[6][2];
    if (TEA_S == 4) {
        float2 *ps = reinterpret_cast<float2 *>(&tile[0][0][0]);
        float2 *mine = ps + ((grp * TEA_S + part) * 6) * 32 + lane;
        mine[0 * 32] = B.ax[0]; mine[1 * 32] = B.ay[0]; mine[2 * 32] = B.az[0];
        mine[3 * 32] = B.ax[1]; mine[4 * 32] = B.ay[1]; mine[5 * 32] = B.az[1];
        __syncthreads();
        if (part != 0) return;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const float2 p0 = ps[((grp * TEA_S + 0) * 6 + c) * 32 + lane], p1 = ps[((grp * TEA_S + 1) * 6 + c) * 32 + lane];
            const float2 p2 = ps[((grp * TEA_S + 2) * 6 + c) * 32 + lane], p3 = ps[((grp * TEA_S + 3) * 6 + c) * 32 + lane];
            sum[c][0] = warp_sum((p0.x + p1.x) + (p2.x + p3.x));
            sum[c][1] = warp_sum((p0.y + p1.y) + (p2.y + p3.y));
        }
    } else {
        const float2 acc[6] = {B.ax[0], B.ay[0], B.az[0], B.ax[1], B.ay[1], B.az[1]};
#pragma unroll
        for (int c = 0; c < 6; c++) {
            sum[c][0] = warp_sum(acc[c].x);
            sum[c][1] = warp_sum(acc[c].y);
        }
    }
    const int i = i0 + lane;
    if (lane >= TEA_IB || i >= N) return;
    // lane b finishes bead i0 + b = pair b / 2, half b % 2
    float sx = 0.f, sy = 0.f, sz = 0.f;
#pragma unroll
    for (int b = 0; b < TEA_IB; b++)
        if (lane == b) {
            sx = sum[3 * (b / 2) + 0][b % 2];
            sy = sum[3 * (b / 2) + 1][b % 2];
            sz = sum[3 * (b / 2) + 2][b % 2];
        }
    const float4 ci = co[i], fm = mf[i], fr = rf[i], raw = a.tea_ci[base + i];
    sx += fm.x + fr.x * (1.f / sqrtf(1.f + b2 * raw.x));
    sy += fm.y + fr.y * (1.f / sqrtf(1.f + b2 * raw.y));
    sz += fm.z + fr.z * (1.f / sqrtf(1.f + b2 * raw.z));
    // The angular stream advances for every bead (:194), the update only for free ones (:196-204)
    uint4 st = a.rng_ang[base + i];
    const float4 rf_ang = rforce(st);
    a.rng_ang[base + i] = st;
    if (!(a.sflags[i] & 1) && ci.w == 0.f) {
        const float mult = p.dt / p.gammaR;
        const float4 A = a.ang[base + i], FA = a.fang[base + i];
        a.pos[base + i] = make_float4(ci.x + mult * sx, ci.y + mult * sy, ci.z + mult * sz, 0.f);
        float fi = A.x, psi = A.y, theta = A.z;
        fi += (p.dt / (p.gammaTheta * p.alpha)) * FA.x + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.x;
        psi += (p.dt / (p.gammaTheta * p.alpha)) * FA.y + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.y;
        theta += (p.dt / p.gammaTheta) * FA.z + p.varTheta * rf_ang.z;
        a.ang[base + i] = make_float4(fi, psi, theta, 0.f);
    }
}

// which = 0: epsilon update (snapshot, per-bead statistics, per-trajectory beta); which = 1: prepare + pair step;
// which = 2: pair step only (the force launch did the prepare part, OP_TEA_PREP)
cudaError_t launch_tea_kernels(const KArgs &k, int which, long long /*step*/, cudaStream_t st)
{
    const int N = k.a.N;
    const size_t n = (size_t)k.a.ntr * N;
    int eblocks = (int)((n + 255) / 256);
    if (eblocks > 148 * 8) eblocks = 148 * 8;
    const dim3 grid((N + TEA_WARPS - 1) / TEA_WARPS, k.a.ntr);
    if (which == 0) {
        tea_snapshot_kernel<<<eblocks, 256, 0, st>>>(k);
        tea_epsilon_kernel<<<grid, TEA_WARPS * 32, 0, st>>>(k);
        tea_beta_kernel<<<k.a.ntr, 256, 0, st>>>(k);
    } else {
        if (which == 1) tea_prepare_kernel<<<eblocks, 256, 0, st>>>(k);
        if (N >= TEA_SPLIT_NTOT) tea_pair_kernel<4, 2><<<dim3((N + 2 * TEA_IB - 1) / (2 * TEA_IB), k.a.ntr), TEA_THREADS, 0, st>>>(k);
        else tea_pair_kernel<1, 8><<<dim3((N + 8 * TEA_IB - 1) / (8 * TEA_IB), k.a.ntr), TEA_THREADS, 0, st>>>(k);
    }
    return cudaGetLastError();
}

} // namespace maddy
