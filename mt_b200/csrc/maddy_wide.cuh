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
 *   wide_grid/count/scan/fill/rebuild_kernel   cell grid + sorted gather: the list rebuild in O(N)       (compute_cuda.cu:913-940, :527-674)
 *   wide_phase_kernel     force / energy from stage[buf]                     (compute_cuda.cu:32-525, :676-911)
 *   wide_step_kernel      force from stage[buf] -> integrate -> state + stage[buf^1]: ONE launch per step
 *                         (replaces compute_kernel + integrate_kernel + 2 cudaDeviceSynchronize, compute_cuda.cu:1228-1238)
 *   wide_reduce_kernel    per-trajectory energy sums (OutputAllEnergies, updater.cpp:3-43)
 */
#pragma once

namespace maddy {

#define WIDE_THREADS 128

struct GStage { // read side: the buffer is read-only for the lifetime of the kernel that reads it (ld.global.nc)
    static constexpr bool kGlobal = true;
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
    near.stride = 0;
    return near;
}

// ---- cell grid for the list rebuild (per trajectory, rebuilt with the lists).  Cells are at least WIDE_CELL_MARGIN x
// the search radius wide, so every partner of a monomer sits in the 27 cells around its own; the members of those cells
// inside the radius are gathered, SORTED (the reference's lists are in ascending j, compute_cuda.cu:913-940) and then put
// through the very statements of rebuild_lists_all_pairs.  O(N) per rebuild instead of O(N^2), whatever the index order
// of the structure (a lattice is index-coherent, free dimers in a cylinder are not).
#define WIDE_MAX_CELLS 32768
#define WIDE_AXIS_CELLS 128
#define WIDE_CELL_MARGIN 1.001f
#define WIDE_GATHER_CAP 320 // partners inside the radius: the Verlet list holds at most 256 of them (more is MADDY_EOVERFLOW anyway)
struct WGrid {
    float ox, oy, oz, ihx, ihy, ihz;
    int nx, ny, nz, ncells;
};
// per trajectory: WGrid header (64 B), then unsigned count[WIDE_MAX_CELLS], start[WIDE_MAX_CELLS], cursor[WIDE_MAX_CELLS], then uint16 members[Npad]
__device__ __forceinline__ size_t wgrid_stride(const DevSys &a) { return 64 + (size_t)3 * WIDE_MAX_CELLS * 4 + (size_t)a.Npad * 2; }
struct WCells {
    WGrid *g;
    unsigned *count, *start, *cursor;
    uint16_t *members;
};
__device__ __forceinline__ WCells wcells_at(const DevSys &a, int traj)
{
    char *b = reinterpret_cast<char *>(a.wgrid) + (size_t)traj * wgrid_stride(a);
    WCells c;
    c.g = reinterpret_cast<WGrid *>(b);
    c.count = reinterpret_cast<unsigned *>(b + 64);
    c.start = c.count + WIDE_MAX_CELLS;
    c.cursor = c.start + WIDE_MAX_CELLS;
    c.members = reinterpret_cast<uint16_t *>(c.cursor + WIDE_MAX_CELLS);
    return c;
}
__device__ __forceinline__ float search_radius2(const KArgs &k, unsigned ops)
{
    return fmaxf((ops & OP_REBUILD_LJ) ? k.cut_pairs.hi : 0.f, (ops & OP_REBUILD_BONDS) ? MD_BOND_PREFILTER2 : 0.f) * 1.00001f;
}
__device__ __forceinline__ int cell_coord(float x, float o, float ih, int n)
{
    return min(n - 1, max(0, (int)((x - o) * ih)));
}

// (1) bounding box of the trajectory -> grid geometry; clears the counts.  One CTA per trajectory.
__global__ void __launch_bounds__(256) wide_grid_kernel(const __grid_constant__ KArgs k, int buf)
{
    __shared__ float red[6][8];
    __shared__ WGrid sg;
    const DevSys &a = k.a;
    const int traj = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float4 *P = gstage_array(a, buf, 0) + (size_t)traj * a.N;
    const float inf = __int_as_float(0x7f800000);
    float v[6] = {inf, inf, inf, inf, inf, inf}; // min x,y,z and min of the negated coordinates
    for (int j = threadIdx.x; j < a.N; j += blockDim.x) {
        const float4 p = P[j];
        v[0] = fminf(v[0], p.x); v[1] = fminf(v[1], p.y); v[2] = fminf(v[2], p.z);
        v[3] = fminf(v[3], -p.x); v[4] = fminf(v[4], -p.y); v[5] = fminf(v[5], -p.z);
    }
#pragma unroll
    for (int q = 0; q < 6; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] = fminf(v[q], __shfl_xor_sync(0xffffffffu, v[q], o));
        if (lane == 0) red[q][warp] = v[q];
    }
    __syncthreads();
    const WCells c = wcells_at(a, traj);
    if (threadIdx.x == 0) {
        for (int q = 0; q < 6; q++)
            for (int w = 1; w < 8; w++) red[q][0] = fminf(red[q][0], red[q][w]);
        const float h = sqrtf(search_radius2(k, k.ops)) * WIDE_CELL_MARGIN;
        const float ext[3] = {-red[3][0] - red[0][0], -red[4][0] - red[1][0], -red[5][0] - red[2][0]};
        int n[3];
        for (int q = 0; q < 3; q++) n[q] = min(WIDE_AXIS_CELLS, (int)(ext[q] / h) + 1);
        while ((long long)n[0] * n[1] * n[2] > WIDE_MAX_CELLS) { // coarsen the axis with the most cells
            const int q = n[0] >= n[1] && n[0] >= n[2] ? 0 : (n[1] >= n[2] ? 1 : 2);
            n[q] = (n[q] + 1) / 2;
        }
        WGrid g;
        g.ox = red[0][0]; g.oy = red[1][0]; g.oz = red[2][0];
        // cell width = max(h, extent / n): never below the margin-widened radius
        g.ihx = 1.0f / fmaxf(h, ext[0] / (float)n[0] * WIDE_CELL_MARGIN);
        g.ihy = 1.0f / fmaxf(h, ext[1] / (float)n[1] * WIDE_CELL_MARGIN);
        g.ihz = 1.0f / fmaxf(h, ext[2] / (float)n[2] * WIDE_CELL_MARGIN);
        g.nx = n[0]; g.ny = n[1]; g.nz = n[2];
        g.ncells = n[0] * n[1] * n[2];
        sg = g;
        *c.g = g;
    }
    __syncthreads();
    for (int q = threadIdx.x; q < sg.ncells; q += blockDim.x) c.count[q] = 0;
}
__device__ __forceinline__ int cell_of(const WGrid &g, const float4 &p)
{
    return (cell_coord(p.z, g.oz, g.ihz, g.nz) * g.ny + cell_coord(p.y, g.oy, g.ihy, g.ny)) * g.nx + cell_coord(p.x, g.ox, g.ihx, g.nx);
}
// (2) members per cell
__global__ void __launch_bounds__(256) wide_count_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const int traj = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.N) return;
    const WCells c = wcells_at(a, traj);
    atomicAdd(c.count + cell_of(*c.g, gstage_array(a, buf, 0)[(size_t)traj * a.N + j]), 1u);
}
// (3) exclusive scan of the counts.  One CTA of 1024 threads per trajectory, 32 cells per thread at most.
__global__ void __launch_bounds__(1024) wide_scan_kernel(const __grid_constant__ KArgs k)
{
    __shared__ unsigned wsum[32];
    const DevSys &a = k.a;
    const WCells c = wcells_at(a, blockIdx.x);
    const int ncells = c.g->ncells;
    const int per = (ncells + 1023) / 1024;
    const int q0 = threadIdx.x * per, q1 = min(ncells, q0 + per);
    unsigned s = 0;
    for (int q = q0; q < q1; q++) s += c.count[q];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
    }
    __syncthreads();
    unsigned run = wsum[warp] + inc - s;
    for (int q = q0; q < q1; q++) {
        c.start[q] = run;
        c.cursor[q] = run;
        run += c.count[q];
    }
}
// (4) member lists (order inside a cell is whatever the atomics give: the gather sorts)
__global__ void __launch_bounds__(256) wide_fill_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const int traj = blockIdx.y, j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.N) return;
    const WCells c = wcells_at(a, traj);
    const unsigned slot = atomicAdd(c.cursor + cell_of(*c.g, gstage_array(a, buf, 0)[(size_t)traj * a.N + j]), 1u);
    c.members[slot] = (uint16_t)j;
}

// (5b) the rebuild proper, one THREAD per monomer (large ensembles: throughput over latency, ~3x fewer instructions per
// monomer than the warp version below, same lists): gather from the 27 cells, sort ascending, then the statements of rebuild_lists_all_pairs
// (LJ_kernel compute_cuda.cu:913-940, pairs_kernel :527-674) on the survivors only.
__global__ void __launch_bounds__(WIDE_THREADS) wide_rebuild_thread_kernel(const __grid_constant__ KArgs k, int buf)
{
    const DevSys &a = k.a;
    const int N = a.N;
    const int traj = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const size_t base = (size_t)traj * N;
    const GStage s = gstage_at(a, buf, traj);
    const unsigned ops = k.ops;
    const bool do_lj = (ops & OP_REBUILD_LJ) != 0;
    const bool do_b = (ops & OP_REBUILD_BONDS) != 0;
    Mono m;
    load_mono(a, base, i, m);
    uint16_t *ljb = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad;
    int status = 0, nlj = 0;
    BondOut bo;
    bo.col = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
    bo.nlong = bo.nlat = bo.status = 0;
    if (!(m.flags & MF_EXTRA)) {
        const WCells c = wcells_at(a, traj);
        const WGrid g = *c.g;
        const float rc2 = search_radius2(k, ops);
        const float x = m.x, y = m.y, z = m.z;
        uint16_t found[WIDE_GATHER_CAP];
        int nf = 0;
        const int cx = cell_coord(x, g.ox, g.ihx, g.nx), cy = cell_coord(y, g.oy, g.ihy, g.ny), cz = cell_coord(z, g.oz, g.ihz, g.nz);
        for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.nz - 1); zz++)
            for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); yy++) {
                // the cells x-1..x+1 of a row are contiguous in memory, and so are their member ranges
                const int q0 = (zz * g.ny + yy) * g.nx + max(cx - 1, 0), q1 = (zz * g.ny + yy) * g.nx + min(cx + 1, g.nx - 1);
                const unsigned m0 = c.start[q0], m1 = c.start[q1] + c.count[q1];
                for (unsigned q = m0; q < m1; q++) {
                    const int j = c.members[q];
                    const float4 Pj = s.P(j);
                    const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                    const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (sf <= rc2 && j != i) {
                        if (nf < WIDE_GATHER_CAP) found[nf] = (uint16_t)j;
                        else status |= do_lj ? ST_LJ_OVERFLOW : ST_LAT_OVERFLOW;
                        nf = min(nf + 1, WIDE_GATHER_CAP);
                    }
                }
            }
        for (int q = 1; q < nf; q++) { // insertion sort (a few dozen entries, thread-local)
            const uint16_t v = found[q];
            int r = q - 1;
            while (r >= 0 && found[r] > v) {
                found[r + 1] = found[r];
                r--;
            }
            found[r + 1] = v;
        }
        const int hraw = a.harm[a.maxH * i]; // first entry, whatever harmonicCount says (compute_cuda.cu:548)
        const int hp = hraw < 0 ? -hraw : hraw;
        const float4 Ei = s.E(i), L1i = s.L1(i), L2i = s.L2(i);
        for (int q = 0; q < nf; q++) {
            const int j = found[q];
            const float4 Pj = s.P(j);
            const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
            const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            if (do_lj && inside_cut(k.cut_pairs, dx, dy, dz, sf)) {
                if (nlj < MADDY_LJ_CAPACITY) ljb[(size_t)nlj * a.Npad + i] = (uint16_t)j;
                else status |= ST_LJ_OVERFLOW;
                nlj++;
            }
            if (do_b && sf < MD_BOND_PREFILTER2 && hp != j) bond_candidates(a, s, i, j, m, hraw, Pj, Ei, L1i, L2i, bo);
        }
    }
    if (do_lj) a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
    if (do_b) {
        uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
        bc[0] = (uint8_t)min(bo.nlong, a.capLong);
        bc[a.Npad] = (uint8_t)min(bo.nlat, a.capLat);
    }
    status |= bo.status;
    if (status) atomicOr(a.status, status);
}

// (5) the rebuild proper, ONE WARP PER MONOMER: the lanes scan the members of the 27 cells side by side (the member
// ranges of a cell row are contiguous), mark the partners inside the radius in a per-warp BITMAP over the monomer indices
// (shared memory), and enumerate its set bits - which yields them in ascending j, the order of the reference's lists
// (compute_cuda.cu:913-940), without a sort.  The survivors then take the very statements of rebuild_lists_all_pairs
// (LJ_kernel, pairs_kernel :527-674): the exact list test on the lanes with a ballot-ordered append, the bond candidates
// (a handful per monomer, order-dependent) one after the other on lane 0.  (The one-thread-per-monomer version walked
// ~500 members through dependent L2 loads and insertion-sorted in local memory: 105 us per rebuild of 5200 monomers.)
#define WIDE_RB_WARPS 4
__global__ void __launch_bounds__(WIDE_RB_WARPS * 32) wide_rebuild_kernel(const __grid_constant__ KArgs k, int buf)
{
    extern __shared__ unsigned s_rb[]; // [warps][nwords] bitmaps, then uint16 [warps][WIDE_GATHER_CAP] survivors
    const DevSys &a = k.a;
    const int N = a.N;
    const int traj = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    const int nwords = (N + 31) / 32;
    unsigned *bits = s_rb + (size_t)warp * nwords;
    uint16_t *found = reinterpret_cast<uint16_t *>(s_rb + (size_t)WIDE_RB_WARPS * nwords) + warp * WIDE_GATHER_CAP;
    for (int w = lane; w < nwords; w += 32) bits[w] = 0;
    __syncwarp();
    const size_t base = (size_t)traj * N;
    const GStage s = gstage_at(a, buf, traj);
    const unsigned ops = k.ops;
    const bool do_lj = (ops & OP_REBUILD_LJ) != 0;
    const bool do_b = (ops & OP_REBUILD_BONDS) != 0;
    const WCells c = wcells_at(a, traj);
    const WGrid g = *c.g;
    const float rc2 = search_radius2(k, ops);
    uint16_t *ljb = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad;
    int status = 0;
    for (int i = blockIdx.x * WIDE_RB_WARPS + warp; i < N; i += gridDim.x * WIDE_RB_WARPS) {
        Mono m;
        load_mono(a, base, i, m); // every lane: the same addresses (broadcast)
        int nlj = 0;
        BondOut bo;
        bo.col = a.bl + (size_t)traj * (a.capLong + a.capLat) * a.Npad + i;
        bo.nlong = bo.nlat = bo.status = 0;
        if (!(m.flags & MF_EXTRA)) {
            const float x = m.x, y = m.y, z = m.z;
            const int cx = cell_coord(x, g.ox, g.ihx, g.nx), cy = cell_coord(y, g.oy, g.ihy, g.ny), cz = cell_coord(z, g.oz, g.ihz, g.nz);
            int wlo = nwords, whi = -1;
            for (int zz = max(cz - 1, 0); zz <= min(cz + 1, g.nz - 1); zz++)
                for (int yy = max(cy - 1, 0); yy <= min(cy + 1, g.ny - 1); yy++) {
                    // the cells x-1..x+1 of a row are contiguous in memory, and so are their member ranges
                    const int q0 = (zz * g.ny + yy) * g.nx + max(cx - 1, 0), q1 = (zz * g.ny + yy) * g.nx + min(cx + 1, g.nx - 1);
                    const unsigned m0 = c.start[q0], m1 = c.start[q1] + c.count[q1];
                    for (unsigned q = m0 + lane; q < m1; q += 32) {
                        const int j = c.members[q];
                        const float4 Pj = s.P(j);
                        const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                        const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                        if (sf <= rc2 && j != i) {
                            atomicOr(bits + (j >> 5), 1u << (j & 31));
                            wlo = min(wlo, j >> 5);
                            whi = max(whi, j >> 5);
                        }
                    }
                }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wlo = min(wlo, __shfl_xor_sync(0xffffffffu, wlo, o));
                whi = max(whi, __shfl_xor_sync(0xffffffffu, whi, o));
            }
            __syncwarp();
            // set bits in ascending order -> found[]
            int nf = 0;
            for (int w0 = wlo; w0 <= whi; w0 += 32) {
                const int w = w0 + lane;
                unsigned word = w <= whi ? bits[w] : 0u;
                if (w <= whi) bits[w] = 0u; // clean for the next monomer
                const int cnt = __popc(word);
                int inc = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                int pos = nf + inc - cnt;
                while (word) {
                    const int bit = __ffs(word) - 1;
                    word &= word - 1;
                    if (pos < WIDE_GATHER_CAP) found[pos] = (uint16_t)(w * 32 + bit);
                    pos++;
                }
                nf += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (nf > WIDE_GATHER_CAP) {
                status |= do_lj ? ST_LJ_OVERFLOW : ST_LAT_OVERFLOW;
                nf = WIDE_GATHER_CAP;
            }
            __syncwarp();
            const int hraw = a.harm[a.maxH * i]; // first entry, whatever harmonicCount says (compute_cuda.cu:548)
            const int hp = hraw < 0 ? -hraw : hraw;
            const float4 Ei = s.E(i), L1i = s.L1(i), L2i = s.L2(i);
            for (int q0 = 0; q0 < nf; q0 += 32) {
                const int q = q0 + lane;
                const bool valid = q < nf;
                const int j = valid ? (int)found[q] : i;
                const float4 Pj = s.P(j);
                const float dx = x - Pj.x, dy = y - Pj.y, dz = z - Pj.z;
                const float sf = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                const bool inlj = valid && do_lj && inside_cut(k.cut_pairs, dx, dy, dz, sf);
                const unsigned bl = __ballot_sync(0xffffffffu, inlj);
                if (inlj) {
                    const int pos = nlj + __popc(bl & lt);
                    if (pos < MADDY_LJ_CAPACITY) ljb[(size_t)pos * a.Npad + i] = (uint16_t)j;
                    else status |= ST_LJ_OVERFLOW;
                }
                nlj += __popc(bl);
                unsigned bb = __ballot_sync(0xffffffffu, valid && do_b && sf < MD_BOND_PREFILTER2 && hp != j);
                while (bb) { // order-dependent and rare: one candidate at a time, on lane 0
                    const int src = __ffs(bb) - 1;
                    bb &= bb - 1;
                    const int jb = __shfl_sync(0xffffffffu, j, src);
                    float4 Pb;
                    Pb.x = __shfl_sync(0xffffffffu, Pj.x, src);
                    Pb.y = __shfl_sync(0xffffffffu, Pj.y, src);
                    Pb.z = __shfl_sync(0xffffffffu, Pj.z, src);
                    Pb.w = __shfl_sync(0xffffffffu, Pj.w, src);
                    if (lane == 0) bond_candidates(a, s, i, jb, m, hraw, Pb, Ei, L1i, L2i, bo);
                }
            }
        }
        if (lane == 0) {
            if (do_lj) a.ljcnt[(size_t)traj * a.Npad + i] = (uint16_t)min(nlj, MADDY_LJ_CAPACITY);
            if (do_b) {
                uint8_t *bc = a.bcnt + (size_t)traj * 2 * a.Npad + i;
                bc[0] = (uint8_t)min(bo.nlong, a.capLong);
                bc[a.Npad] = (uint8_t)min(bo.nlat, a.capLat);
            }
            status |= bo.status;
        }
        __syncwarp();
    }
    if (status) atomicOr(a.status, status);
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

    if (i >= N) return;
    if ((k.ops & (OP_REBUILD_LJ | OP_REBUILD_BONDS)) && !a.wgrid)
        rebuild_lists_all_pairs<1>(k, s, traj, mo, idx, k.ops); // MADDY_WIDE_ALL_PAIRS=1 (test hook): O(N^2) scan

    if (k.ops & OP_FORCE) {
        const bool extra = (mo[0].flags & MF_EXTRA) != 0;
        G6 f = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (!extra) {
            F3 e, l1, l2;
            const Frame fr = make_frame(mo[0].fi, mo[0].psi, mo[0].theta, ls, e, l1, l2);
            f = monomer_force<true>(k, s, near, traj, i, mo[0], fr);
            if (!(k.ops & OP_TEA_PREP)) a.fpos[base + i] = make_float4(f.x, f.y, f.z, 0.f);
            a.fang[base + i] = make_float4(f.fi, f.psi, f.theta, 0.f);
        }
        if (k.ops & OP_TEA_PREP) tea_prepare_bead(k, base + i, mo[0], f, extra);
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
    if (*a.guard) return;
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
        const G6 f = monomer_force<true>(k, s, near, traj, i, m, fr);
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

// ---- persistent window: all steps between two list-update steps (or host / scheduled events) in ONE launch.
// The per-step launch above is latency-bound by construction (41 CTAs for 5200 monomers, launch gap + ramp per step, the
// full Verlet row gathered through L2 every step).  Here the CTAs of a trajectory stay resident (cooperative launch) and
// meet at a PER-TRAJECTORY barrier once per step (a monotone ticket counter in HBM: trajectories never wait for one
// another); state, RNG streams and the packed topology words stay in registers / shared memory for the whole window,
// and the force loop walks a per-monomer NEAR list (listed partners within MD_NEAR_R when it was formed, kept in the
// CTA's shared memory, guarded by the same displacement test as the one-CTA path and re-formed by the whole trajectory
// when the guard trips).  The stage stays in HBM/L2 (double-buffered), read with ld.global.cg: it is rewritten every
// step, so the read-only path of GStage does not apply.  Same device functions, same order of operations as
// wide_step_kernel and the one-CTA path: bit-identical results.
struct GStageCG {
    static constexpr bool kGlobal = true;
    const float4 *p, *e, *l1, *l2;
    __device__ __forceinline__ float4 P(int j) const { return __ldcg(p + j); }
    __device__ __forceinline__ float4 E(int j) const { return __ldcg(e + j); }
    __device__ __forceinline__ float4 L1(int j) const { return __ldcg(l1 + j); }
    __device__ __forceinline__ float4 L2(int j) const { return __ldcg(l2 + j); }
};
__device__ __forceinline__ GStageCG gstagecg_at(const DevSys &a, int buf, int traj)
{
    const size_t base = (size_t)traj * a.N;
    GStageCG s;
    s.p = gstage_array(a, buf, 0) + base;
    s.e = gstage_array(a, buf, 1) + base;
    s.l1 = gstage_array(a, buf, 2) + base;
    s.l2 = gstage_array(a, buf, 3) + base;
    return s;
}

#define WIDE_NEAR_CAP 16

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs of one trajectory: arrive + wait until `target` arrivals have been counted since the launch
__device__ __forceinline__ void wide_traj_barrier(unsigned *counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_u32(counter) < target) {
        }
        __threadfence();
    }
    __syncthreads();
}

// near list of monomer i := entries of its Verlet row whose CURRENT distance (stage s) is below MD_NEAR_R (as refresh_near)
template <class S>
__device__ __forceinline__ void wide_form_near(const KArgs &k, const S &s, int traj, int i, const Mono &m, uint16_t *row /* [cap][WIDE_THREADS] + tid */,
                                               uint8_t *cnt)
{
    const DevSys &a = k.a;
    int nn = 0;
    if (k.p.lj_on && !(m.flags & MF_EXTRA)) {
        const uint16_t *lj = a.lj + (size_t)traj * MADDY_LJ_CAPACITY * a.Npad + i;
        const int n = a.ljcnt[(size_t)traj * a.Npad + i];
        const size_t stride = a.Npad;
        for (int k0 = 0; k0 < n; k0 += 16, lj += 16 * stride) {
            unsigned jj[16];
            float4 Pv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) jj[u] = k0 + u < n ? (unsigned)lj[u * stride] : (unsigned)i;
#pragma unroll
            for (int u = 0; u < 16; u++) Pv[u] = s.P(jj[u]);
#pragma unroll
            for (int u = 0; u < 16; u++) {
                const float dx = m.x - Pv[u].x, dy = m.y - Pv[u].y, dz = m.z - Pv[u].z;
                if (k0 + u < n && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) < MD_NEAR_R2) {
                    if (nn < WIDE_NEAR_CAP) row[nn * WIDE_THREADS] = (uint16_t)(jj[u] | MD_NEAR_LJ_FLAG);
                    nn++;
                }
            }
        }
    }
    *cnt = (uint8_t)(nn > WIDE_NEAR_CAP ? MD_NEAR_FULL : nn);
}

__global__ void __launch_bounds__(WIDE_THREADS, 5) wide_run_kernel(const __grid_constant__ KArgs k, int buf, int n_steps, int publish_first)
{
    __shared__ uint16_t s_near[WIDE_NEAR_CAP * WIDE_THREADS];
    __shared__ uint8_t s_cnt[WIDE_THREADS];
    __shared__ uint4 s_topo[WIDE_THREADS];
    const DevSys &a = k.a;
    if (*a.guard) return; // see run_kernel (every CTA reads the same word: nobody is left waiting at a barrier)
    const int N = a.N;
    const int traj = blockIdx.y, tid = threadIdx.x;
    const size_t base = (size_t)traj * N;
    const int i0 = blockIdx.x * WIDE_THREADS, i = i0 + tid;
    const bool active = i < N;
    const LatSite ls = lateral_site();
    unsigned *counter = a.wbar + 4 * (size_t)traj; // [0] barrier tickets, [1..2] step parity words of the displacement guard
    unsigned nbar = 0;                             // barriers passed since the launch
    const unsigned nctas = gridDim.x;

    Near near = no_near();
    near.list = s_near - i0; // (near.list + i)[kk * stride] == s_near[kk * WIDE_THREADS + tid]
    near.cnt = s_cnt - i0;
    near.topo = s_topo - i0;
    near.cap = WIDE_NEAR_CAP;
    near.stride = WIDE_THREADS;

    Mono m;
    m.flags = MF_EXTRA | MF_FIXED;
    Frame fr = {};
    if (active) {
        load_mono(a, base, i, m);
        s_topo[tid] = load_topo(a, traj, i);
        s_cnt[tid] = MD_NEAR_FULL;
    }
    const bool mobile = active && !(m.flags & (MF_EXTRA | MF_FIXED));
    if (mobile) {
        m.rx = a.rng_xyz[base + i];
        m.ra = a.rng_ang[base + i];
    }
    if (publish_first) {
        if (active) publish_global(a, buf, base + i, m, ls);
        wide_traj_barrier(counter, ++nbar * nctas);
    }
    if (mobile) {
        F3 e, l1, l2;
        fr = make_frame(m.fi, m.psi, m.theta, ls, e, l1, l2);
    }
    const bool use_near = k.p.lj_on != 0;
    if (use_near && active) wide_form_near(k, gstagecg_at(a, buf, traj), traj, i, m, s_near + tid, s_cnt + tid);
    near.ok = use_near;
    float gx = m.x, gy = m.y, gz = m.z; // positions when the near list was formed (displacement guard)

    for (int step = 0; step < n_steps; step++) {
        const GStageCG s = gstagecg_at(a, buf, traj);
        if (mobile) {
            const G6 f = monomer_force<true>(k, s, near, traj, i, m, fr);
            integrate_monomer(k, m, f);
            const float dx = m.x - gx, dy = m.y - gy, dz = m.z - gz;
            if (use_near && fmaf(dz, dz, fmaf(dy, dy, dx * dx)) > MD_NEAR_GUARD2) atomicMax(counter + 1 + (nbar & 1u), nbar + 1);
        }
        if (step + 1 == n_steps) break;
        if (active) fr = publish_global(a, buf ^ 1, base + i, m, ls);
        const unsigned parity = nbar & 1u;
        wide_traj_barrier(counter, ++nbar * nctas);
        buf ^= 1;
        if (use_near && ld_acquire_u32(counter + 1 + parity) == nbar) { // some monomer of this trajectory moved beyond the guard
            if (active) wide_form_near(k, gstagecg_at(a, buf, traj), traj, i, m, s_near + tid, s_cnt + tid);
            gx = m.x;
            gy = m.y;
            gz = m.z;
            if (tid == 0 && blockIdx.x == 0) atomicAdd(a.stats + 0, 1ull);
        }
    }
    if (mobile) {
        a.pos[base + i] = make_float4(m.x, m.y, m.z, 0.f);
        a.ang[base + i] = make_float4(m.fi, m.psi, m.theta, 0.f);
        a.rng_xyz[base + i] = m.rx;
        a.rng_ang[base + i] = m.ra;
    }
    if (active) publish_global(a, buf ^ 1, base + i, m, ls); // the stage of the state just written (for a rebuild / the next window)
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
    if ((k.ops & (OP_REBUILD_LJ | OP_REBUILD_BONDS)) && k.a.wgrid) {
        const dim3 mgrid((k.a.N + 255) / 256, k.a.ntr);
        wide_grid_kernel<<<k.a.ntr, 256, 0, st>>>(k, buf);
        wide_count_kernel<<<mgrid, 256, 0, st>>>(k, buf);
        wide_scan_kernel<<<k.a.ntr, 1024, 0, st>>>(k);
        wide_fill_kernel<<<mgrid, 256, 0, st>>>(k, buf);
        if ((size_t)k.a.ntr * k.a.N <= 32768 && !getenv("MADDY_WIDE_THREAD_REBUILD")) {
            // few monomers: a warp per monomer (latency); measured 29 us against 105 us at 5200 x 1
            const int nwords = (k.a.N + 31) / 32;
            const size_t smem = (size_t)WIDE_RB_WARPS * nwords * 4 + (size_t)WIDE_RB_WARPS * WIDE_GATHER_CAP * 2;
            wide_rebuild_kernel<<<dim3((k.a.N + WIDE_RB_WARPS - 1) / WIDE_RB_WARPS, k.a.ntr), WIDE_RB_WARPS * 32, smem, st>>>(k, buf);
        } else {
            // the GPU is full either way: a thread per monomer (164 us against 316 us at 5200 x 16)
            wide_rebuild_thread_kernel<<<grid, WIDE_THREADS, 0, st>>>(k, buf);
        }
        if (!(k.ops & (OP_FORCE | OP_ENERGY))) return cudaGetLastError();
    }
    wide_phase_kernel<<<grid, WIDE_THREADS, 0, st>>>(k, buf);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (k.ops & OP_ENERGY) {
        wide_reduce_kernel<<<k.a.ntr, 256, 0, st>>>(k);
        e = cudaGetLastError();
    }
    return e;
}
// Persistent window of n_steps steps from stage[buf] (published first if asked).  Returns cudaErrorCooperativeLaunchTooLarge
// when the CTAs of the ensemble cannot all be resident (the caller then falls back to one launch per step).
cudaError_t launch_wide_run(const KArgs &k, int buf, int n_steps, int publish_first, cudaStream_t st)
{
    static int per_sm = -1, n_sm = 0;
    if (per_sm < 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wide_run_kernel, WIDE_THREADS, 0);
    }
    const dim3 grid((k.a.N + WIDE_THREADS - 1) / WIDE_THREADS, k.a.ntr);
    if ((long long)grid.x * grid.y > (long long)per_sm * n_sm) return cudaErrorCooperativeLaunchTooLarge;
    cudaError_t e = cudaMemsetAsync(k.a.wbar, 0, (size_t)k.a.ntr * 4 * sizeof(unsigned), st);
    if (e != cudaSuccess) return e;
    KArgs kk = k;
    void *args[] = {(void *)&kk, (void *)&buf, (void *)&n_steps, (void *)&publish_first};
    return cudaLaunchCooperativeKernel((const void *)wide_run_kernel, grid, dim3(WIDE_THREADS), args, 0, st);
}
cudaError_t launch_wide_step(const KArgs &k, int buf, cudaStream_t st)
{
    const dim3 grid((k.a.N + WIDE_THREADS - 1) / WIDE_THREADS, k.a.ntr);
    wide_step_kernel<<<grid, WIDE_THREADS, 0, st>>>(k, buf);
    return cudaGetLastError();
}

} // namespace maddy
