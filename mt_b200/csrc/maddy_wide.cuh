/*
 * maddy_wide.cuh — the WIDE path: trajectories too long for one CTA's shared memory (N > MADDY_MAX_NTOT_CTA; the
 * "large-N single system" of BASELINE.json config 5) are spread over the whole GPU, one thread per monomer, many CTAs
 * per trajectory.  Included at the end of maddy_kernels.cu (same translation unit, same -fmad=false arithmetic), so
 * the wide kernels call the very device functions of the one-CTA path — monomer_force / monomer_energy /
 * rebuild_lists_all_pairs, instantiated on a stage that lives in HBM/L2 instead of shared memory — and a system that
 * fits both paths gives bit-identical lists, forces, coordinates and RNG streams on either
 * (tests/test_gpu_parity.py::test_wide_path_equals_cta_path_bitwise).
 *
 *   stage in HBM: [2 buffers][P,E,L1,L2][ntr*N] float4 (64 B per monomer and buffer; 5200 monomers = 333 KB, L2-resident)
 *   wide_publish_kernel   state -> stage[buf]                                (before a step-granular phase, once per window)
 *   wide_phase_kernel     rebuild / force / energy from stage[buf]           (compute_cuda.cu:913-940, :527-674, :32-525, :676-911)
 *   wide_step_kernel      force from stage[buf] -> integrate -> state + stage[buf^1]: ONE launch per step
 *                         (replaces compute_kernel + integrate_kernel + 2 cudaDeviceSynchronize, compute_cuda.cu:1228-1238)
 *   wide_reduce_kernel    per-trajectory energy sums (OutputAllEnergies, updater.cpp:3-43)
 */
#pragma once

namespace maddy {

#define WIDE_THREADS 128

struct GStage { // read side: the buffer is read-only for the lifetime of the kernel that reads it (ld.global.nc)
    const float4 *p, *e, *l1, *l2;
    __device__ __forceinline__ float4 P(int j) const { return __ldg(p + j); }
    __device__ __forceinline__ float4 E(int j) const { return __ldg(e + j); }
    __device__ __forceinline__ float4 L1(int j) const { return __ldg(l1 + j); }
    __device__ __forceinline__ float4 L2(int j) const { return __ldg(l2 + j); }
};

__device__ __forceinline__ float4 *gstage_array(const DevSys &a, int buf, int r)
{
    return a.gstage + (size_t)(buf * 4 + r) * ((size_t)a.ntr * a.N);
}
__device__ __forceinline__ GStage gstage_at(const DevSys &a, int buf, int traj)
{
    const size_t base = (size_t)traj * a.N;
    GStage s;
    s.p = gstage_array(a, buf, 0) + base;
    s.e = gstage_array(a, buf, 1) + base;
    s.l1 = gstage_array(a, buf, 2) + base;
    s.l2 = gstage_array(a, buf, 3) + base;
    return s;
}

// same words as publish() writes into the shared-memory stage
__device__ __forceinline__ Frame publish_global(const DevSys &a, int buf, size_t q, const Mono &m, const LatSite &ls)
{
    F3 e, l1, l2;
    const Frame fr = make_frame(m.fi, m.psi, m.theta, ls, e, l1, l2);
    const int jf = MF_TYPE(m.flags) | (m.flags & (MF_GTP | MF_ONTUB | MF_EXTRA));
    gstage_array(a, buf, 0)[q] = make_float4(m.x, m.y, m.z, m.fi);
    gstage_array(a, buf, 1)[q] = make_float4(e.x, e.y, e.z, m.psi);
    gstage_array(a, buf, 2)[q] = make_float4(l1.x, l1.y, l1.z, m.theta);
    gstage_array(a, buf, 3)[q] = make_float4(l2.x, l2.y, l2.z, __int_as_float(jf));
    return fr;
}

__device__ __forceinline__ Near no_near()
{
    Near near;
    near.list = nullptr;
    near.cnt = nullptr;
    near.tlo = near.thi = nullptr;
    near.cap = 0;
    near.ntiles = 0;
    near.ok = false;
    near.stale_lj = false;
    near.topo = nullptr;
    return near;
}

__global__ void __launch_bounds__(256) wide_publish_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const LatSite ls = lateral_site();
    const size_t n = (size_t)a.ntr * a.N;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(q % a.N);
        Mono m;
        load_mono(a, q - i, i, m);
        publish_global(a, buf, q, m, ls);
    }
}

// One phase for every monomer, grid (ceil(N / WIDE_THREADS), ntr).  A monomer's list rows are written and read by its
// own thread only, so rebuild + energies may share a launch like in phase_kernel.
__global__ void __launch_bounds__(WIDE_THREADS) wide_phase_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const int N = a.N;
    const int traj = blockIdx.y;
    const size_t base = (size_t)traj * N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const LatSite ls = lateral_site();
    const GStage s = gstage_at(a, buf, traj);
    const Near near = no_near();

    Mono mo[1];
    int idx[1] = {i < N ? i : N};
    mo[0].flags = MF_EXTRA | MF_FIXED;
    if (i < N) load_mono(a, base, i, mo[0]);

    if (k.ops & (OP_REBUILD_LJ | OP_REBUILD_BONDS)) rebuild_lists_all_pairs<1>(k, s, traj, mo, idx, k.ops);
    if (i >= N) return;

    if ((k.ops & OP_FORCE) && !(mo[0].flags & MF_EXTRA)) {
        F3 e, l1, l2;
        const Frame fr = make_frame(mo[0].fi, mo[0].psi, mo[0].theta, ls, e, l1, l2);
        const G6 f = monomer_force(k, s, near, traj, i, mo[0], fr);
        a.fpos[base + i] = make_float4(f.x, f.y, f.z, 0.f);
        a.fang[base + i] = make_float4(f.fi, f.psi, f.theta, 0.f);
    }
    if (k.ops & OP_ENERGY) {
        const E7 en = monomer_energy(k, s, traj, i, mo[0]);
        double *o = a.en_mono + (base + i) * 7;
        o[0] = en.harm; o[1] = en.lng; o[2] = en.lat; o[3] = en.psi; o[4] = en.fi; o[5] = en.teta; o[6] = en.lj;
    }
}

// per-trajectory sums of the per-monomer energies, order harm,long,lat,psi,fi,teta,lj (updater.cpp:35-36)
__global__ void __launch_bounds__(256) wide_reduce_kernel(const __grid_constant__ KArgs k)
{
    __shared__ double red_scratch[32 * 7];
    const DevSys &a = k.a;
    const int traj = blockIdx.x;
    const double *e = a.en_mono + (size_t)traj * a.N * 7;
    E7 acc = {0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
        const double *o = e + (size_t)i * 7;
        acc.harm += o[0]; acc.lng += o[1]; acc.lat += o[2]; acc.psi += o[3];
        acc.fi += o[4]; acc.teta += o[5]; acc.lj += o[6];
    }
    block_reduce_e7(acc, a.en_traj + (size_t)traj * 7, red_scratch);
}

// One whole step: forces from stage[buf], Euler-Maruyama update in registers, new state to HBM and to stage[buf^1].
__global__ void __launch_bounds__(WIDE_THREADS) wide_step_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const int N = a.N;
    const int traj = blockIdx.y;
    const size_t base = (size_t)traj * N;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const LatSite ls = lateral_site();
    const GStage s = gstage_at(a, buf, traj);
    const Near near = no_near();
    Mono m;
    load_mono(a, base, i, m);
    if (!(m.flags & (MF_EXTRA | MF_FIXED))) {
        F3 e, l1, l2;
        const Frame fr = make_frame(m.fi, m.psi, m.theta, ls, e, l1, l2);
        const G6 f = monomer_force(k, s, near, traj, i, m, fr);
        m.rx = a.rng_xyz[base + i];
        m.ra = a.rng_ang[base + i];
        integrate_monomer(k, m, f);
        a.pos[base + i] = make_float4(m.x, m.y, m.z, 0.f);
        a.ang[base + i] = make_float4(m.fi, m.psi, m.theta, 0.f);
        a.rng_xyz[base + i] = m.rx;
        a.rng_ang[base + i] = m.ra;
    }
    publish_global(a, buf ^ 1, base + i, m, ls);
}

static int wide_blocks(size_t n)
{
    size_t b = (n + 255) / 256;
    return (int)(b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b));
}

cudaError_t launch_wide_publish(const KArgs &k, int buf, cudaStream_t st)
{
    wide_publish_kernel<<<wide_blocks((size_t)k.a.ntr * k.a.N), 256, 0, st>>>(k, buf);
    return cudaGetLastError();
}
// ops: any of OP_REBUILD_LJ | OP_REBUILD_BONDS | OP_FORCE | OP_ENERGY, from stage[buf] (publish first)
cudaError_t launch_wide_phase(const KArgs &k, int buf, cudaStream_t st)
{
    const dim3 grid((k.a.N + WIDE_THREADS - 1) / WIDE_THREADS, k.a.ntr);
    wide_phase_kernel<<<grid, WIDE_THREADS, 0, st>>>(k, buf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (k.ops & OP_ENERGY) {
        wide_reduce_kernel<<<k.a.ntr, 256, 0, st>>>(k);
        e = cudaGetLastError();
    }
    return e;
}
cudaError_t launch_wide_step(const KArgs &k, int buf, cudaStream_t st)
{
    const dim3 grid((k.a.N + WIDE_THREADS - 1) / WIDE_THREADS, k.a.ntr);
    wide_step_kernel<<<grid, WIDE_THREADS, 0, st>>>(k, buf);
    return cudaGetLastError();
}

} // namespace maddy
