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
// Mapping: LANES ARE BEADS.  A warp owns a block of 32 beads (one per lane) and walks partners four at a time; the partner
// records live in shared memory as structure-of-arrays, so one broadcast LDS.128 per component hands every lane the same
// four partners as two PACKED pairs: sm_100a's FFMA2 / FMUL2 / FADD2 (fma.rn.f32x2) then do the same operation for two
// partners of the lane's bead in one instruction - measured 66 TFLOP/s against 42 TFLOP/s for three-register scalar FFMA
// on this GPU (tools/micro/ffma2_bench.cu) - while the bead's own quantities enter as scalars broadcast to both halves.
// Nothing is reduced across lanes (the previous mapping - partners on lanes, four beads per warp - spent more instructions
// on its butterfly reductions, its per-group prologue / epilogue with 4 of 32 lanes active and its padding rounds than on
// the pairs: 2196 warp instructions per 2080 pairs, 43 % of them in the pair loop).
// Per pair the tensor is applied in its dyadic form  D g = cii g + crr (u.g) u  (eq. 3-5 of the paper; the reference
// multiplies the six entries out, bdhitea_kernel.cu:38-58) with one MUFU (rsqrt): 1/ra = a / w needs no division.
// Reserve beads carry zero force records and the padding up to a multiple of four partners is a far-away bead with zero
// forces (exact zero contributions, no test); the overlap branch ra <= 2 is a warp-voted slow path and the self pair is
// masked only in the eight steps that cover the warp's own block.
// Work split (functions of N alone, so a trajectory's result never depends on how many trajectories share the launch):
// TEA_S warps of a CTA share one bead block and take every TEA_S-th step; for long trajectories the partners are also cut
// into Z segments handled by different CTAs, whose partial sums meet in HBM and are added in segment order by the CTA that
// finishes last (ticket counter), so even ONE trajectory of a few thousand beads fills the GPU.
#define TEA_BLOCK 32
#define TEA_TILE 512 // partners per shared-memory tile (9 floats each: 18 KB)
#define TEA_SPLIT_NTOT 2048 // trajectories at least this long use 4 warps per bead block and partner segments

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

struct TeaAcc { float2 x[2], y[2], z[2]; }; // [h]: partners 4k+2h (.x) and 4k+2h+1 (.y)

// One step = partners j..j+3 of the tile against the lane's bead.  i: the lane's bead (global index in the trajectory).
template <bool MASKED>
__device__ __forceinline__ void tea_step(TeaAcc &A, const float (*T)[TEA_TILE], int jj, int j, int i, float npx, float npy, float npz, float Cx,
                                         float Cy, float Cz, float ta, float inv_a, float near2)
{
    const float4 X = *reinterpret_cast<const float4 *>(&T[0][jj]), Y = *reinterpret_cast<const float4 *>(&T[1][jj]),
                 Z = *reinterpret_cast<const float4 *>(&T[2][jj]);
    const float4 MX = *reinterpret_cast<const float4 *>(&T[3][jj]), MY = *reinterpret_cast<const float4 *>(&T[4][jj]),
                 MZ = *reinterpret_cast<const float4 *>(&T[5][jj]);
    const float4 RX = *reinterpret_cast<const float4 *>(&T[6][jj]), RY = *reinterpret_cast<const float4 *>(&T[7][jj]),
                 RZ = *reinterpret_cast<const float4 *>(&T[8][jj]);
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const float2 dx = __fadd2_rn(h ? f2(X.z, X.w) : f2(X.x, X.y), f2(npx, npx));
        const float2 dy = __fadd2_rn(h ? f2(Y.z, Y.w) : f2(Y.x, Y.y), f2(npy, npy));
        const float2 dz = __fadd2_rn(h ? f2(Z.z, Z.w) : f2(Z.x, Z.y), f2(npz, npz));
        float2 w2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
        bool on0 = true, on1 = true;
        if (MASKED) { // the self pair
            on0 = j + 2 * h != i;
            on1 = j + 2 * h + 1 != i;
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
        const bool near = fminf(w2.x, w2.y) <= near2;
        if (MASKED || __any_sync(0xffffffffu, near)) { // overlapping beads (ra <= 2) and the self pair: rare, scalar
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
        const float2 gx = __ffma2_rn(h ? f2(RX.z, RX.w) : f2(RX.x, RX.y), f2(Cx, Cx), h ? f2(MX.z, MX.w) : f2(MX.x, MX.y));
        const float2 gy = __ffma2_rn(h ? f2(RY.z, RY.w) : f2(RY.x, RY.y), f2(Cy, Cy), h ? f2(MY.z, MY.w) : f2(MY.x, MY.y));
        const float2 gz = __ffma2_rn(h ? f2(RZ.z, RZ.w) : f2(RZ.x, RZ.y), f2(Cz, Cz), h ? f2(MZ.z, MZ.w) : f2(MZ.x, MZ.y));
        const float2 pr = __fmul2_rn(crr, __ffma2_rn(uz, gz, __ffma2_rn(uy, gy, __fmul2_rn(ux, gx))));
        A.x[h] = __ffma2_rn(cii, gx, __ffma2_rn(pr, ux, A.x[h]));
        A.y[h] = __ffma2_rn(cii, gy, __ffma2_rn(pr, uy, A.y[h]));
        A.z[h] = __ffma2_rn(cii, gz, __ffma2_rn(pr, uz, A.z[h]));
    }
}

// TEA_NS steps (4 * TEA_NS partners) at once, common case only (no self pair): 2 * TEA_NS independent packed streams.
// A warp issues one dependent packed operation every ~10 cycles (measured: `wait` is the top stall with two streams) and
// registers cap the SM at ~4 warps per scheduler, so the parallelism has to come from inside the warp.  Only the squared
// distances are formed before the ONE overlap vote (8 registers per stream); behind it the streams share nothing but the
// accumulators, so the compiler interleaves them freely.  A hit (rare) redoes the steps one by one with the scalar
// fix-ups.  Accumulation order = step 0 (h = 0, 1), step 1, ... as when the steps are taken one by one: bit-identical.
#define TEA_NS 2
__device__ __forceinline__ void tea_steps(TeaAcc &A, const float (*T)[TEA_TILE], int jj0, int jstride, int j0, int i, float npx, float npy, float npz,
                                          float Cx, float Cy, float Cz, float ta, float inv_a, float near2)
{
    float2 dx[2 * TEA_NS], dy[2 * TEA_NS], dz[2 * TEA_NS], w2[2 * TEA_NS];
    float wmin = 3.0e38f;
#pragma unroll
    for (int q = 0; q < 2 * TEA_NS; q++) { // q = 2 * step + h
        const int jj = jj0 + (q >> 1) * jstride + 2 * (q & 1);
        dx[q] = __fadd2_rn(*reinterpret_cast<const float2 *>(&T[0][jj]), f2(npx, npx));
        dy[q] = __fadd2_rn(*reinterpret_cast<const float2 *>(&T[1][jj]), f2(npy, npy));
        dz[q] = __fadd2_rn(*reinterpret_cast<const float2 *>(&T[2][jj]), f2(npz, npz));
        w2[q] = __ffma2_rn(dz[q], dz[q], __ffma2_rn(dy[q], dy[q], __fmul2_rn(dx[q], dx[q])));
        wmin = fminf(wmin, fminf(w2[q].x, w2[q].y));
    }
    if (__any_sync(0xffffffffu, wmin <= near2)) { // overlapping beads (ra <= 2) somewhere in the warp: rare
#pragma unroll 1
        for (int st = 0; st < TEA_NS; st++)
            tea_step<false>(A, T, jj0 + st * jstride, j0 + st * jstride, i, npx, npy, npz, Cx, Cy, Cz, ta, inv_a, near2);
        return;
    }
#pragma unroll
    for (int q = 0; q < 2 * TEA_NS; q++) {
        const int jj = jj0 + (q >> 1) * jstride + 2 * (q & 1), h = q & 1;
        const float2 iw = f2(rsqrtf(w2[q].x), rsqrtf(w2[q].y));
        const float2 ux = __fmul2_rn(dx[q], iw), uy = __fmul2_rn(dy[q], iw), uz = __fmul2_rn(dz[q], iw);
        const float2 ira = __fmul2_rn(f2(ta, ta), iw); // 1 / ra
        const float2 ira2 = __fmul2_rn(ira, ira);
        const float2 far = __fmul2_rn(f2(0.75f, 0.75f), ira);
        const float2 crr = __fmul2_rn(far, __ffma2_rn(f2(-2.f, -2.f), ira2, f2(1.f, 1.f)));
        const float2 cii = __fmul2_rn(far, __ffma2_rn(f2(2.f / 3.f, 2.f / 3.f), ira2, f2(1.f, 1.f)));
        const float2 gx = __ffma2_rn(*reinterpret_cast<const float2 *>(&T[6][jj]), f2(Cx, Cx), *reinterpret_cast<const float2 *>(&T[3][jj]));
        const float2 gy = __ffma2_rn(*reinterpret_cast<const float2 *>(&T[7][jj]), f2(Cy, Cy), *reinterpret_cast<const float2 *>(&T[4][jj]));
        const float2 gz = __ffma2_rn(*reinterpret_cast<const float2 *>(&T[8][jj]), f2(Cz, Cz), *reinterpret_cast<const float2 *>(&T[5][jj]));
        const float2 pr = __fmul2_rn(crr, __ffma2_rn(uz, gz, __ffma2_rn(uy, gy, __fmul2_rn(ux, gx))));
        A.x[h] = __ffma2_rn(cii, gx, __ffma2_rn(pr, ux, A.x[h]));
        A.y[h] = __ffma2_rn(cii, gy, __ffma2_rn(pr, uy, A.y[h]));
        A.z[h] = __ffma2_rn(cii, gz, __ffma2_rn(pr, uz, A.z[h]));
    }
}

// number of partner segments for a trajectory of N beads (a function of N alone, see above)
__host__ __device__ inline int tea_segments(int N) { return N < TEA_SPLIT_NTOT ? 1 : ((N + 384) / 768 > 16 ? 16 : (N + 384) / 768); }

template <int TEA_S>
__global__ void __launch_bounds__(TEA_S * 32, 16 / TEA_S) tea_pair_kernel(const __grid_constant__ KArgs k)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.z, seg = blockIdx.y, nseg = gridDim.y;
    const int tid = threadIdx.x, lane = tid & 31, part = tid >> 5;
    const int i0 = blockIdx.x * TEA_BLOCK;
    const int i = min(i0 + lane, N - 1); // lanes beyond N shadow the last bead (their results are dropped)
    const size_t base = (size_t)traj * N;
    const float4 *co = a.tea_co + base, *mf = a.tea_mf + base, *rf = a.tea_rf + base;
    // partner tile, structure of arrays: x, y, z | molecular force x, y, z | random force x, y, z
    __shared__ __align__(16) float tile[9][TEA_TILE];
    __shared__ bool last_cta;

    const float beta = a.tea_beta[traj];
    const float b2 = beta * beta;
    const float4 ci = co[i], raw = a.tea_ci[base + i];
    const float npx = -ci.x, npy = -ci.y, npz = -ci.z;
    const float sxc = 1.f / sqrtf(1.f + b2 * raw.x), syc = 1.f / sqrtf(1.f + b2 * raw.y), szc = 1.f / sqrtf(1.f + b2 * raw.z);
    const float Cx = beta * sxc, Cy = beta * syc, Cz = beta * szc; // beta * C_i
    const float ta = p.tea_a, inv_a = 1.f / p.tea_a, near2 = 4.f * p.tea_a * p.tea_a;
    TeaAcc A;
#pragma unroll
    for (int h = 0; h < 2; h++) A.x[h] = A.y[h] = A.z[h] = f2(0.f, 0.f);

    // this CTA's partner segment, in steps of four partners
    const int steps_all = (N + 3) / 4;
    const int per_seg = (steps_all + nseg - 1) / nseg;
    const int st_lo = seg * per_seg, st_hi = min(st_lo + per_seg, steps_all);
    const int self_lo = i0 / 4; // steps self_lo .. self_lo + 7 hold the warp's own beads
    for (int t0 = st_lo; t0 < st_hi; t0 += TEA_TILE / 4) {
        const int nst = min(TEA_TILE / 4, st_hi - t0);
        if (t0 != st_lo) __syncthreads(); // everybody is done with the previous tile
        for (int jj = tid; jj < nst * 4; jj += TEA_S * 32) {
            const int j = t0 * 4 + jj;
            float4 c = make_float4(3.0e5f, 3.0e5f, 3.0e5f, 0.f), m = make_float4(0.f, 0.f, 0.f, 0.f), r = m; // padding: far away, no force
            if (j < N) {
                c = co[j];
                m = mf[j];
                r = rf[j];
            }
            tile[0][jj] = c.x; tile[1][jj] = c.y; tile[2][jj] = c.z;
            tile[3][jj] = m.x; tile[4][jj] = m.y; tile[5][jj] = m.z;
            tile[6][jj] = r.x; tile[7][jj] = r.y; tile[8][jj] = r.z;
        }
        __syncthreads();
        auto one = [&](int st) {
            const int gs = t0 + st;
            if ((unsigned)(gs - self_lo) < 8u) tea_step<true>(A, tile, st * 4, gs * 4, i, npx, npy, npz, Cx, Cy, Cz, ta, inv_a, near2);
            else tea_step<false>(A, tile, st * 4, gs * 4, i, npx, npy, npz, Cx, Cy, Cz, ta, inv_a, near2);
        };
        int st = part;
#pragma unroll 1
        for (; st + (TEA_NS - 1) * TEA_S < nst; st += TEA_NS * TEA_S) {
            const int g0 = t0 + st - self_lo; // the group covers steps g0, g0 + TEA_S, ... relative to the warp's own block
            if (g0 < 8 && g0 + (TEA_NS - 1) * TEA_S >= 0) {
#pragma unroll 1
                for (int q = 0; q < TEA_NS; q++) one(st + q * TEA_S);
            } else {
                tea_steps(A, tile, st * 4, TEA_S * 4, (t0 + st) * 4, i, npx, npy, npz, Cx, Cy, Cz, ta, inv_a, near2);
            }
        }
#pragma unroll 1
        for (; st + TEA_S < nst; st += TEA_S) one(st);
        if (st < nst) one(st);
    }
    // per lane: the four partner streams in a fixed order, then the parts of the CTA in part order
    float sx = (A.x[0].x + A.x[0].y) + (A.x[1].x + A.x[1].y);
    float sy = (A.y[0].x + A.y[0].y) + (A.y[1].x + A.y[1].y);
    float sz = (A.z[0].x + A.z[0].y) + (A.z[1].x + A.z[1].y);
    if (TEA_S > 1) {
        __syncthreads(); // the tile is free
        float *ps = &tile[0][0];
        ps[(part * 3 + 0) * 32 + lane] = sx;
        ps[(part * 3 + 1) * 32 + lane] = sy;
        ps[(part * 3 + 2) * 32 + lane] = sz;
        __syncthreads();
        if (part != 0) return;
        sx = ps[0 * 32 + lane];
        sy = ps[1 * 32 + lane];
        sz = ps[2 * 32 + lane];
#pragma unroll
        for (int q = 1; q < TEA_S; q++) {
            sx += ps[(q * 3 + 0) * 32 + lane];
            sy += ps[(q * 3 + 1) * 32 + lane];
            sz += ps[(q * 3 + 2) * 32 + lane];
        }
    }
    // (one warp from here on)
    if (nseg > 1) {
        // partial sums of the segments meet in HBM; the CTA whose ticket is the last one adds them in segment order
        float4 *mine = a.tea_part + ((size_t)traj * nseg + seg) * N;
        if (i0 + lane < N) mine[i0 + lane] = make_float4(sx, sy, sz, 0.f);
        __threadfence();
        __syncwarp();
        unsigned *cnt = a.tea_cnt + (size_t)traj * gridDim.x + blockIdx.x;
        if (lane == 0) {
            const unsigned ticket = atomicAdd(cnt, 1u);
            last_cta = ticket == (unsigned)nseg - 1;
            if (last_cta) *cnt = 0; // ready for the next step
        }
        __syncwarp();
        if (!last_cta) return;
        __threadfence();
        sx = sy = sz = 0.f;
        if (i0 + lane < N)
            for (int q = 0; q < nseg; q++) {
                const float4 v = __ldcg(a.tea_part + ((size_t)traj * nseg + q) * N + i0 + lane);
                sx += v.x;
                sy += v.y;
                sz += v.z;
            }
    }
    if (i0 + lane >= N) return;
    const float4 fm = mf[i], fr = rf[i];
    sx += fm.x + fr.x * sxc;
    sy += fm.y + fr.y * syc;
    sz += fm.z + fr.z * szc;
    // The angular stream advances for every bead (:194), the update only for free ones (:196-204)
    uint4 st = a.rng_ang[base + i];
    const float4 rf_ang = rforce(st);
    a.rng_ang[base + i] = st;
    if (!(a.sflags[i] & 1) && ci.w == 0.f) {
        const float mult = p.dt / p.gammaR;
        const float4 AN = a.ang[base + i], FA = a.fang[base + i];
        a.pos[base + i] = make_float4(ci.x + mult * sx, ci.y + mult * sy, ci.z + mult * sz, 0.f);
        float fi = AN.x, psi = AN.y, theta = AN.z;
        fi += (p.dt / (p.gammaTheta * p.alpha)) * FA.x + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.x;
        psi += (p.dt / (p.gammaTheta * p.alpha)) * FA.y + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.y;
        theta += (p.dt / p.gammaTheta) * FA.z + p.varTheta * rf_ang.z;
        a.ang[base + i] = make_float4(fi, psi, theta, 0.f);
    }
}

int tea_partner_segments(int N) { return tea_segments(N); }

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
        const dim3 pgrid((N + TEA_BLOCK - 1) / TEA_BLOCK, tea_segments(N), k.a.ntr);
        if (N >= TEA_SPLIT_NTOT) tea_pair_kernel<4><<<pgrid, 4 * 32, 0, st>>>(k);
        else tea_pair_kernel<2><<<pgrid, 2 * 32, 0, st>>>(k);
    }
    return cudaGetLastError();
}

} // namespace maddy
