/*
 * maddy_tea.cu — TEA hydrodynamic-interaction integrator (Geyer & Winter 2009,
 * doi:10.1063/1.3089668) with the Rotne-Prager-Yamakawa tensor, all-pairs ("unlisted").
 *
 * What the reference does (src/bdhitea_kernel.cu:16-213, src/bdhitea.cu:37-118): three
 * launches per step with every pair term gathered from global memory, plus a D2H of
 * per-bead epsilon sums, a serial host loop for beta and an H2D every tea_epsilon_freq steps.
 *
 * Here: one CTA per trajectory; coordinates, molecular forces and pre-drawn random forces
 * are staged once in shared memory (3 x float4 per bead) and the O(N^2) pair loop reads
 * broadcast LDS.128 only; epsilon is reduced per trajectory with warp shuffles and beta
 * (eq. 26) is evaluated on the device, so nothing crosses PCIe.
 * The pair work is a generated 3x3 mat-vec per (i,j) with a per-i right-hand side
 * (f_j + C_i o r_j): not a GEMM, FP32-pipe bound; tensor cores do not apply.
 */
#include "maddy_kernels.cuh"

namespace maddy {

#define KB_BOLTZ 0.0019872041f // kcal/(mol*K), mt.h:39
#define ST_TEA_ABORT 0x100

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

// ---- epsilon / C_i statistics + beta  (integrateTea_epsilon_unlisted :84-101, updateTea bdhitea.cu:57-118)
__global__ void __launch_bounds__(MD_MAX_THREADS) tea_epsilon_kernel(const __grid_constant__ KArgs k)
{
    extern __shared__ float4 sC[]; // x,y,z,extra
    __shared__ double red[32];
    __shared__ int redn[32];
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float4 P = a.pos[base + i];
        sC[i] = make_float4(P.x, P.y, P.z, a.extra[base + i] ? 1.f : 0.f);
    }
    __syncthreads();
    double eps_acc = 0.0;
    int n_acc = 0;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float4 ci = sC[i];
        float sx = 0.f, sy = 0.f, sz = 0.f, sw = 0.f;
        if (ci.w == 0.f) {
            n_acc++;
            for (int j = 0; j < N; j++) {
                const float4 cj = sC[j];
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
        a.tea_ci[base + i] = make_float4(sx, sy, sz, 0.f);
        a.tea_eps[base + i] = sw;
        eps_acc += (double)sw;
    }
    // per-trajectory reduction
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    for (int o = 16; o > 0; o >>= 1) {
        eps_acc += __shfl_down_sync(0xffffffffu, eps_acc, o);
        n_acc += __shfl_down_sync(0xffffffffu, n_acc, o);
    }
    if (lane == 0) {
        red[warp] = eps_acc;
        redn[warp] = n_acc;
    }
    __syncthreads();
    if (warp == 0) {
        double e = lane < nwarp ? red[lane] : 0.0;
        int n = lane < nwarp ? redn[lane] : 0;
        for (int o = 16; o > 0; o >>= 1) {
            e += __shfl_down_sync(0xffffffffu, e, o);
            n += __shfl_down_sync(0xffffffffu, n, o);
        }
        if (lane == 0) {
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
}

// ---- integrateTea_prepare + integrateTea_kernel_unlisted (bdhitea_kernel.cu:16-36, :148-213)
__global__ void __launch_bounds__(MD_MAX_THREADS) tea_integrate_kernel(const __grid_constant__ KArgs k)
{
    extern __shared__ float4 sm[];
    const maddy_params &p = k.p;
    const DevSys &a = k.a;
    const int N = a.N, traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    float4 *sC = sm, *sM = sm + N, *sR = sm + 2 * N; // coords(+extra), molecular force, random force

    // prepare: every bead draws, fixed and extra included (bdhitea_kernel.cu:22)
    const float var = sqrtf(2.0f * KB_BOLTZ * p.Temp * p.gammaR / p.dt);
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        uint4 st = a.rng_xyz[base + i];
        float4 df = rforce(st);
        a.rng_xyz[base + i] = st;
        df.x *= var;
        df.y *= var;
        df.z *= var;
        const float4 F = a.fpos[base + i];
        const float4 P = a.pos[base + i];
        sR[i] = df;
        sM[i] = make_float4(F.x, F.y, F.z, 0.f);
        sC[i] = make_float4(P.x, P.y, P.z, a.extra[base + i] ? 1.f : 0.f);
        a.fpos[base + i] = make_float4(0.f, 0.f, 0.f, 0.f); // only xyz is zeroed (:31-33)
    }
    __syncthreads();

    const float beta = a.tea_beta[traj];
    const float mult = p.dt / p.gammaR;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float4 raw = a.tea_ci[base + i];
        const float b2 = beta * beta;
        float cx = 1.f / sqrtf(1.f + b2 * raw.x);
        float cy = 1.f / sqrtf(1.f + b2 * raw.y);
        float cz = 1.f / sqrtf(1.f + b2 * raw.z);
        const float4 co = sC[i];
        const float4 fm = sM[i], fr = sR[i];
        float fx = fm.x + fr.x * cx, fy = fm.y + fr.y * cy, fz = fm.z + fr.z * cz;
        cx *= beta;
        cy *= beta;
        cz *= beta;
        if (co.w == 0.f) {
            for (int j = 0; j < N; j++) {
                const float4 cj = sC[j];
                if (j == i || cj.w != 0.f) continue;
                float dx = cj.x - co.x, dy = cj.y - co.y, dz = cj.z - co.z;
                const float w = sqrtf(dx * dx + dy * dy + dz * dz);
                dx /= w;
                dy /= w;
                dz /= w;
                const float4 mj = sM[j], rj = sR[j];
                const float gx = mj.x + rj.x * cx, gy = mj.y + rj.y * cy, gz = mj.z + rj.z * cz;
                const Sym6 d = rpy(dx, dy, dz, w, p.tea_a);
                fx += d.xx * gx + d.xy * gy + d.xz * gz;
                fy += d.xy * gx + d.yy * gy + d.yz * gz;
                fz += d.xz * gx + d.yz * gy + d.zz * gz;
            }
        }
        // angular stream advances for every bead (:194), the update only for free ones (:196-204)
        uint4 st = a.rng_ang[base + i];
        const float4 rf_ang = rforce(st);
        a.rng_ang[base + i] = st;
        const int sf = a.sflags[i];
        if (!(sf & 1) && co.w == 0.f) {
            const float4 A = a.ang[base + i], FA = a.fang[base + i];
            a.pos[base + i] = make_float4(co.x + mult * fx, co.y + mult * fy, co.z + mult * fz, 0.f);
            float fi = A.x, psi = A.y, theta = A.z;
            fi += (p.dt / (p.gammaTheta * p.alpha)) * FA.x + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.x;
            psi += (p.dt / (p.gammaTheta * p.alpha)) * FA.y + (p.varTheta * sqrtf(p.freeze_temp / p.alpha)) * rf_ang.y;
            theta += (p.dt / p.gammaTheta) * FA.z + p.varTheta * rf_ang.z;
            a.ang[base + i] = make_float4(fi, psi, theta, 0.f);
        }
    }
}

cudaError_t launch_tea_kernels(const KArgs &k, int which, long long /*step*/, cudaStream_t st)
{
    const int N = k.a.N;
    int threads = ((N < MD_MAX_THREADS ? N : MD_MAX_THREADS) + 31) & ~31;
    if (which == 0) {
        const size_t smem = (size_t)N * sizeof(float4);
        cudaError_t e = cudaFuncSetAttribute(tea_epsilon_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tea_epsilon_kernel<<<k.a.ntr, threads, smem, st>>>(k);
    } else {
        const size_t smem = (size_t)3 * N * sizeof(float4);
        cudaError_t e = cudaFuncSetAttribute(tea_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        tea_integrate_kernel<<<k.a.ntr, threads, smem, st>>>(k);
    }
    return cudaGetLastError();
}

} // namespace maddy
