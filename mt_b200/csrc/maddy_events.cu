/*
 * maddy_events.cu — hydrolyse() (updater.cpp:229-257) on the device, for ALL events of one output stride at once.
 *
 * What the reference does at every hydrostep: a host loop, dimer-outer / trajectory-inner, that draws one libc rand()
 * per eligible dimer (GTP, not in the reserve, on the tubule now and at the previous stride), hydrolyses it with
 * probability 0.02, then a second loop that returns GDP dimers off the tubule (now and before) to GTP; followed by an
 * H2D copy of the whole gtp array (compute_cuda.cu:1153-1160).
 *
 * Here: the eligibility of a dimer changes inside a stride only through hydrolysis itself (the on-tubule flags and the
 * reserve flags change at stride steps), so right after the stride block the device evaluates every event up to the
 * next stride step in one go and leaves the results where the fused loop already looks for them - the GTP schedule
 * (KArgs::gtp_sched, one slot per event).  The position of a dimer's draw in the rand() stream is a prefix sum over the
 * eligibility mask in the reference's order (rows = dimers, columns = trajectories); the stream itself is produced on
 * the device from the host generator's 31-word window by polynomial jump-ahead (maddy_lfib.h).  The host is left with
 * advancing its generator by the number of draws the device reports - it issues no per-event work at all, and a fused
 * window can span the whole stride.  Bit-identical to hydrolyse(): same draws, same order, same threshold.
 */
#include "maddy_kernels.cuh"
#include "maddy_lfib.h"
#include <stdlib.h>

namespace maddy {

#define HYD_ROUNDS 8                        // rounds of 31 draws per thread of the stream kernel
#define HYD_PER_THREAD (31 * HYD_ROUNDS)    // 248
#define HYD_RADIX 32                        // thread t = (a * 32 + b) * 32 + c starts at draw t * HYD_PER_THREAD
#define HYD_TOP 256                         // values of a: streams of up to 256 * 1024 * 248 = 65 M draws per plan

// Jump matrices, made once per handle: level 0 holds the offsets c * B, level 1 b * 32 B, level 2 a * 1024 B (B =
// HYD_PER_THREAD); row i of a matrix = coefficients of z^(offset + i) mod p, so that (M w)[i] = x[offset + i] for the
// window w.  A thread of the stream kernel then reaches ITS window with three 31 x 31 matrix-vector products instead
// of ~20 polynomial products (the jump was 8/9 of the stream kernel's instructions).
__global__ void __launch_bounds__(32) hyd_matrices_kernel(const LfibPoly *__restrict__ table, uint32_t *__restrict__ mats)
{
    const int m = blockIdx.x, i = threadIdx.x; // matrix, row
    if (i >= LFIB_DEG) return;
    const int level = m < HYD_RADIX ? 0 : (m < 2 * HYD_RADIX ? 1 : 2);
    const unsigned long long j = level == 0 ? m : (level == 1 ? m - HYD_RADIX : m - 2 * HYD_RADIX);
    const unsigned long long unit = level == 0 ? HYD_PER_THREAD : (level == 1 ? (unsigned long long)HYD_RADIX * HYD_PER_THREAD
                                                                               : (unsigned long long)HYD_RADIX * HYD_RADIX * HYD_PER_THREAD);
    LfibPoly q;
    lfib_power(q, table, j * unit + i);
    for (int k = 0; k < LFIB_DEG; k++) mats[((size_t)m * LFIB_DEG + i) * LFIB_DEG + k] = q.c[k];
}

__device__ __forceinline__ void hyd_matvec(uint32_t (&w)[LFIB_DEG], const uint32_t *__restrict__ M)
{
    uint32_t out[LFIB_DEG];
#pragma unroll
    for (int i = 0; i < LFIB_DEG; i++) {
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < LFIB_DEG; k++) v += M[i * LFIB_DEG + k] * w[k];
        out[i] = v;
    }
#pragma unroll
    for (int i = 0; i < LFIB_DEG; i++) w[i] = out[i];
}

// out[v] = draw v of the plan = x[31 + v] >> 1, v < count; W[k] = x[k] is the generator's window (oldest word first)
__global__ void __launch_bounds__(64) hyd_stream_kernel(const uint32_t *__restrict__ W, const uint32_t *__restrict__ mats, unsigned long long count,
                                                        uint32_t *__restrict__ out, const int *__restrict__ guard)
{
    if (*guard) return;
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long first = t * HYD_PER_THREAD;
    if (first >= count) return;
    uint32_t r[LFIB_DEG];
#pragma unroll
    for (int k = 0; k < LFIB_DEG; k++) r[k] = W[k];
    const unsigned c = (unsigned)(t % HYD_RADIX), b = (unsigned)(t / HYD_RADIX % HYD_RADIX), a = (unsigned)(t / (HYD_RADIX * HYD_RADIX));
    if (a) hyd_matvec(r, mats + (size_t)(2 * HYD_RADIX + a) * LFIB_DEG * LFIB_DEG);
    if (b) hyd_matvec(r, mats + (size_t)(HYD_RADIX + b) * LFIB_DEG * LFIB_DEG);
    if (c) hyd_matvec(r, mats + (size_t)c * LFIB_DEG * LFIB_DEG);
    // r[j] = x[first + j]: the 31 words in front of draw `first`
    for (int round = 0; round < HYD_ROUNDS; round++) {
        const unsigned long long v0 = first + (unsigned long long)round * 31;
        if (v0 >= count) break;
#pragma unroll
        for (int j = 0; j < 31; j++) {
            r[j] = r[j] + r[(j + 28) % 31]; // x[n] = x[n-31] + x[n-3]; (j + 28) % 31 < j holds the value made three steps ago
            if (v0 + j < count) out[v0 + j] = r[j] >> 1;
        }
    }
}

// The plan's working set: rows = dimers, columns = trajectories in GLOBAL order (the reference's loop order).  A sharded
// ensemble gathers every shard's block (contiguous trajectories) and each shard evaluates the whole plan on its own copy:
// the draw positions are global, the slots it writes are those of its own trajectories.
__device__ __forceinline__ size_t hyd_cell(const HydArgs &h, int which, int d, int tr)
{
    const int g = tr / h.ntr_l, tl = tr - g * h.ntr_l;
    return (((size_t)g * 2 + which) * h.nd + d) * h.ntr_l + tl;
}

// transposed inputs of this shard: rows = dimers, columns = its trajectories
//   gt[d][tr]  GTP state of the dimer's first monomer as the events go by (1 / 0 / other)
//   st[d][tr]  bit 0: can hydrolyse (not reserve, on the tubule now and before)   bit 1: returns to GTP (not reserve, off both)
__global__ void __launch_bounds__(256) hyd_prepare_kernel(HydArgs h)
{
    if (*h.guard) return;
    const size_t cells = (size_t)h.nd * h.ntr_l;
    for (size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(c / h.ntr_l), tr = (int)(c % h.ntr_l);
        const size_t q = (size_t)tr * h.N + 2 * d;
        const bool ex = h.extra[q] != 0, cur = h.cur[q] != 0, prev = h.prev[q] != 0;
        h.own[c] = h.gtp[q];
        h.own[cells + c] = (uint8_t)((!ex && cur && prev ? 1 : 0) | (!ex && !cur && !prev ? 2 : 0));
    }
}

// draws of one event per dimer row
__global__ void __launch_bounds__(128) hyd_count_kernel(HydArgs h)
{
    __shared__ unsigned wsum[4];
    if (*h.guard) return;
    const int r = blockIdx.x, d = r / h.nseg, t0 = (r % h.nseg) * h.seg, t1 = min(h.ntr, t0 + h.seg);
    unsigned c = 0;
    for (int tr = t0 + threadIdx.x; tr < t1; tr += blockDim.x) c += h.all[hyd_cell(h, 0, d, tr)] == 1 && (h.all[hyd_cell(h, 1, d, tr)] & 1);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) h.rowcount[r] = wsum[0] + wsum[1] + wsum[2] + wsum[3];
}

// first draw of every row = draws consumed so far + exclusive prefix of the row counts; one CTA
__global__ void __launch_bounds__(1024) hyd_scan_kernel(HydArgs h, int event)
{
    __shared__ unsigned long long wsum[32];
    if (*h.guard) return;
    const int per = (h.nrows + 1023) / 1024;
    const int d0 = threadIdx.x * per, d1 = min(h.nrows, d0 + per);
    unsigned long long s = 0;
    for (int d = d0; d < d1; d++) s += h.rowcount[d];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long inc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = wsum[lane], winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        wsum[lane] = winc - w;
    }
    __syncthreads();
    const unsigned long long base = h.cursor[0];
    unsigned long long run = base + wsum[warp] + inc - s;
    for (int d = d0; d < d1; d++) {
        h.rowstart[d] = run;
        run += h.rowcount[d];
    }
    __syncthreads(); // every thread has read the cursor
    if (threadIdx.x == 1023) {
        h.event_start[event] = base;
        h.cursor[0] = run; // the last thread's running sum is the total
    }
}

// one event, one dimer row, one warp: the trajectories in order (ballot prefix = position in the stream)
__device__ __forceinline__ void hyd_apply_row(const HydArgs &h, int r, unsigned long long pos, uint8_t *__restrict__ slot, int lane)
{
    const unsigned lt = (1u << lane) - 1u;
    const int own_lo = h.shard * h.ntr_l;
    const int d = r / h.nseg, s0 = (r % h.nseg) * h.seg, s1 = min(h.ntr, s0 + h.seg);
    for (int t0 = s0; t0 < s1; t0 += 32) {
        const int tr = t0 + lane;
        const bool in = tr < s1;
        const size_t cg = in ? hyd_cell(h, 0, d, tr) : 0;
        uint8_t g = in ? h.all[cg] : (uint8_t)2;
        const uint8_t s = in ? h.all[hyd_cell(h, 1, d, tr)] : (uint8_t)0;
        const bool elig = g == 1 && (s & 1);
        const unsigned bl = __ballot_sync(0xffffffffu, elig);
        if (elig) {
            const unsigned long long v = pos + __popc(bl & lt);
            if (v < h.stream_count) {
                if (h.stream[v] <= h.threshold) g = 0; // rand() / (double)RAND_MAX < 0.02  (updater.cpp:236-237)
            } else {
                atomicOr(h.status, 1);
            }
        }
        pos += __popc(bl);
        if (g == 0 && (s & 2)) g = 1; // off the tubule now and before: back to GTP (updater.cpp:246-254), no draw
        if (in) {
            h.all[cg] = g;
            // this shard's trajectories: both monomers of the dimer (2 d is even and N is even: the pair is 2-byte aligned)
            const int tl = tr - own_lo;
            if (tl >= 0 && tl < h.ntr_l) *reinterpret_cast<uint16_t *>(slot + (size_t)tl * h.N + 2 * d) = (uint16_t)(g | (g << 8));
        }
    }
}
__global__ void __launch_bounds__(128) hyd_apply_kernel(HydArgs h, uint8_t *__restrict__ slot)
{
    if (*h.guard) return;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= h.nrows) return;
    hyd_apply_row(h, r, h.rowstart[r], slot, lane);
}

// The whole plan in ONE launch of ONE CTA (ensembles up to ~1 M dimer-trajectory cells: the per-event kernels above are
// three launches per event, and a stride has ten events): prepare, then per event count -> scan -> apply with CTA
// barriers in between; the row offsets live in shared memory.
__global__ void __launch_bounds__(1024) hyd_plan_fused_kernel(HydArgs h, int n_events, uint8_t *__restrict__ slots)
{
    extern __shared__ unsigned s_row[]; // [nd]: draws per row, then exclusive prefix
    __shared__ unsigned s_wsum[32];
    __shared__ unsigned s_total;
    if (*h.guard) return;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // (one shard only: the inputs were not gathered, this kernel prepares them itself)
    const size_t cells = (size_t)h.nd * h.ntr_l;
    for (size_t c = tid; c < cells; c += 1024) {
        const int d = (int)(c / h.ntr_l), tr = (int)(c % h.ntr_l);
        const size_t q = (size_t)tr * h.N + 2 * d;
        const bool ex = h.extra[q] != 0, cur = h.cur[q] != 0, prev = h.prev[q] != 0;
        h.own[c] = h.gtp[q];
        h.own[cells + c] = (uint8_t)((!ex && cur && prev ? 1 : 0) | (!ex && !cur && !prev ? 2 : 0));
    }
    __syncthreads();
    unsigned long long cursor = 0;
    const int per = (h.nrows + 1023) / 1024;
    for (int k = 0; k < n_events; k++) {
        for (int r = warp; r < h.nrows; r += 32) { // draws of this event per row segment
            const int d = r / h.nseg, t0 = (r % h.nseg) * h.seg, t1 = min(h.ntr, t0 + h.seg);
            unsigned c = 0;
            for (int tr = t0 + lane; tr < t1; tr += 32) c += h.all[hyd_cell(h, 0, d, tr)] == 1 && (h.all[hyd_cell(h, 1, d, tr)] & 1);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if (lane == 0) s_row[r] = c;
        }
        __syncthreads();
        { // exclusive prefix over the rows
            const int d0 = tid * per, d1 = min(h.nrows, d0 + per);
            unsigned sum = 0;
            for (int d = d0; d < d1; d++) sum += s_row[d];
            unsigned inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_wsum[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                const unsigned w = s_wsum[lane];
                unsigned winc = w;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const unsigned t = __shfl_up_sync(0xffffffffu, winc, o);
                    if (lane >= o) winc += t;
                }
                s_wsum[lane] = winc - w;
                if (lane == 31) s_total = winc;
            }
            __syncthreads();
            unsigned run = s_wsum[warp] + inc - sum;
            for (int d = d0; d < d1; d++) {
                const unsigned c = s_row[d];
                s_row[d] = run;
                run += c;
            }
        }
        __syncthreads();
        for (int r = warp; r < h.nrows; r += 32) hyd_apply_row(h, r, cursor + s_row[r], slots + (size_t)k * h.ntr_l * h.N, lane);
        if (tid == 0) h.event_start[k] = cursor;
        cursor += s_total;
        __syncthreads();
    }
    if (tid == 0) h.cursor[0] = cursor;
}

// jump matrices of the stream kernel: (2 * HYD_RADIX + HYD_TOP) x 31 x 31 words, computed on the device
size_t hyd_matrices_bytes() { return (size_t)(2 * HYD_RADIX + HYD_TOP) * LFIB_DEG * LFIB_DEG * sizeof(uint32_t); }
unsigned long long hyd_stream_capacity() { return (unsigned long long)HYD_TOP * HYD_RADIX * HYD_RADIX * HYD_PER_THREAD; }
cudaError_t launch_hyd_matrices(const void *table, void *mats, cudaStream_t st)
{
    hyd_matrices_kernel<<<2 * HYD_RADIX + HYD_TOP, 32, 0, st>>>(reinterpret_cast<const LfibPoly *>(table), reinterpret_cast<uint32_t *>(mats));
    return cudaGetLastError();
}
cudaError_t launch_hyd_stream(const uint32_t *W, const void *mats, unsigned long long count, uint32_t *out, const int *guard, cudaStream_t st)
{
    const unsigned long long threads = (count + HYD_PER_THREAD - 1) / HYD_PER_THREAD;
    if (threads == 0) return cudaSuccess;
    hyd_stream_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>(W, reinterpret_cast<const uint32_t *>(mats), count, out, guard);
    return cudaGetLastError();
}

// this shard's transposed inputs (h.own), for the gather of a sharded ensemble
cudaError_t launch_hyd_prepare(const HydArgs &h, cudaStream_t st)
{
    const size_t cells = (size_t)h.nd * h.ntr_l;
    int pb = (int)((cells + 255) / 256);
    if (pb > 148 * 8) pb = 148 * 8;
    hyd_prepare_kernel<<<pb, 256, 0, st>>>(h);
    return cudaGetLastError();
}

// all events of a plan on h.all (prepared / gathered by the caller when `prepared`), slot k of `slots` ([n_events][ntr_l * N]
// bytes) = GTP state of THIS shard's trajectories after event k
cudaError_t launch_hyd_plan(const HydArgs &h, int n_events, uint8_t *slots, bool prepared, cudaStream_t st)
{
    const size_t cells = (size_t)h.nd * h.ntr;
    // (one CTA is latency-bound on its dependent loads: 1.3 ms per plan at 260 x 256 against 0.18 ms for the per-event
    // launches below; kept for very small ensembles, where the launches dominate)
    if (!prepared && h.shards == 1 && cells <= 4096 && h.nrows <= 8192 && !getenv("MADDY_HYD_PER_EVENT_KERNELS")) {
        hyd_plan_fused_kernel<<<1, 1024, (size_t)h.nrows * sizeof(unsigned), st>>>(h, n_events, slots);
        return cudaGetLastError();
    }
    if (!prepared) {
        cudaError_t e = launch_hyd_prepare(h, st);
        if (e != cudaSuccess) return e;
    }
    for (int k = 0; k < n_events; k++) {
        hyd_count_kernel<<<h.nrows, 128, 0, st>>>(h);
        hyd_scan_kernel<<<1, 1024, 0, st>>>(h, k);
        hyd_apply_kernel<<<(h.nrows + 3) / 4, 128, 0, st>>>(h, slots + (size_t)k * h.ntr_l * h.N);
    }
    return cudaGetLastError();
}

} // namespace maddy
