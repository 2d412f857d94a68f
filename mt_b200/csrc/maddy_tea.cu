/*
 * maddy_tea.cu — TEA hydrodynamic-interaction integrator (Geyer & Winter 2009,
 * doi:10.1063/1.3089668) with the Rotne-Prager-Yamakawa tensor, all-pairs ("unlisted").
 *
 * What the reference does (src/bdhitea_kernel.cu:16-213, src/bdhitea.cu:37-118): one thread per bead walks
 * all N partners sequentially with every term gathered from global memory (3 float4 per pair), plus a D2H of
 * per-bead epsilon sums, a serial host loop for beta and an H2D every tea_epsilon_freq steps, and a host loop
 * over all particles every step.
 *
 * Here the O(N^2) work is spread over the whole GPU even for ONE trajectory: a WARP owns a bead, its 32 lanes
 * stride over the partners (coalesced float4 loads of the snapshot arrays, served by L1/L2: 25 KB per
 * trajectory), and the 3-vector is reduced with shuffles in a fixed order (deterministic).  520 beads = 520
 * warps = every SM busy at Ntr = 1.  epsilon is reduced per trajectory on the device and beta (eq. 26) is
 * evaluated there too, so nothing crosses PCIe.  The pair work is a generated 3x3 mat-vec per (i,j) whose
 * right-hand side (f_j + C_i o r_j) depends on i: not a GEMM, FP32-pipe bound; tensor cores do not apply.
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
        a.tea_rf[q] = df;
        a.tea_mf[q] = make_float4(F.x, F.y, F.z, 0.f);
        a.tea_co[q] = make_float4(P.x, P.y, P.z, a.extra[q] ? 1.f : 0.f);
        a.fpos[q] = make_float4(0.f, 0.f, 0.f, 0.f); // only xyz is zeroed (:31-33)
    }
}

// ---- integrateTea_kernel_unlisted (bdhitea_kernel.cu:148-213): warp per bead
__global__ void __launch_bounds__(TEA_WARPS * 32) tea_pair_kernel(const __grid_constant__ KArgs k)
{
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.y;
    const int i = blockIdx.x * TEA_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= N) return;
    const size_t base = (size_t)traj * N;
    const float4 *co = a.tea_co + base, *mf = a.tea_mf + base, *rf = a.tea_rf + base;

    const float beta = a.tea_beta[traj];
    const float4 raw = a.tea_ci[base + i];
    const float b2 = beta * beta;
    float cx = 1.f / sqrtf(1.f + b2 * raw.x);
    float cy = 1.f / sqrtf(1.f + b2 * raw.y);
    float cz = 1.f / sqrtf(1.f + b2 * raw.z);
    const float4 ci = co[i];
    const float4 fm = mf[i], fr = rf[i];
    const float f0x = fm.x + fr.x * cx, f0y = fm.y + fr.y * cy, f0z = fm.z + fr.z * cz;
    cx *= beta;
    cy *= beta;
    cz *= beta;
    float fx = 0.f, fy = 0.f, fz = 0.f;
    if (ci.w == 0.f) {
        for (int j = lane; j < N; j += 32) {
            const float4 cj = co[j];
            if (j == i || cj.w != 0.f) continue;
            float dx = cj.x - ci.x, dy = cj.y - ci.y, dz = cj.z - ci.z;
            const float w = sqrtf(dx * dx + dy * dy + dz * dz);
            dx /= w;
            dy /= w;
            dz /= w;
            const float4 mj = mf[j], rj = rf[j];
            const float gx = mj.x + rj.x * cx, gy = mj.y + rj.y * cy, gz = mj.z + rj.z * cz;
            const Sym6 d = rpy(dx, dy, dz, w, p.tea_a);
            fx += d.xx * gx + d.xy * gy + d.xz * gz;
            fy += d.xy * gx + d.yy * gy + d.yz * gz;
            fz += d.xz * gx + d.yz * gy + d.zz * gz;
        }
    }
    fx = f0x + warp_sum(fx);
    fy = f0y + warp_sum(fy);
    fz = f0z + warp_sum(fz);
    if (lane != 0) return;
    // angular stream advances for every bead (:194), the update only for free ones (:196-204)
    uint4 st = a.rng_ang[base + i];
    const float4 rf_ang = rforce(st);
    a.rng_ang[base + i] = st;
    if (!(a.sflags[i] & 1) && ci.w == 0.f) {
        const float mult = p.dt / p.gammaR;
        const float4 A = a.ang[base + i], FA = a.fang[base + i];
        a.pos[base + i] = make_float4(ci.x + mult * fx, ci.y + mult * fy, ci.z + mult * fz, 0.f);
        float fi = A.x, psi = A.y, theta = A.z;
        fi += (p.dt / (p.gammaTheta * p.alpha)) * FA.x + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.x;
        psi += (p.dt / (p.gammaTheta * p.alpha)) * FA.y + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.y;
        theta += (p.dt / p.gammaTheta) * FA.z + p.varTheta * rf_ang.z;
        a.ang[base + i] = make_float4(fi, psi, theta, 0.f);
    }
}

// which = 0: epsilon update (snapshot, per-bead statistics, per-trajectory beta); which = 1: prepare + pair step
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
        tea_prepare_kernel<<<eblocks, 256, 0, st>>>(k);
        tea_pair_kernel<<<grid, TEA_WARPS * 32, 0, st>>>(k);
    }
    return cudaGetLastError();
}

} // namespace maddy
