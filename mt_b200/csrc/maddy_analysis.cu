/*
 * maddy_analysis.cu — in-situ analysis kernels: the reference's offline DCD post-processing
 * (scripts/temp_calc/main.cpp, scripts/disas_speed/3d22d.cpp, scripts/disas_speed/disc.cpp) as reductions over the
 * state in HBM.  One CTA per trajectory; the integer results and the projection are bit-identical to the tools run
 * over the DCD files of the same frames, the displacement sums agree to double rounding (parallel summation order).
 *
 * What the tools compute on the host from float DCD frames is restated with the same widths: differences of two
 * floats are float subtractions, squares and sums are double (C's usual arithmetic conversions, pow(double, 2)).
 */
#include "maddy_kernels.cuh"

namespace maddy {

#define AN_THREADS 256
#define AN_MAX_PF 32
// disc.cpp:9-11
#define AN_THRES 5.0
#define AN_HOR_THRES 2.0
#define AN_THETA_THRES 0.2

__device__ __forceinline__ double block_sum(double v, double *scratch /* [AN_THREADS / 32] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads(); // scratch may still be read from the previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (threadIdx.x == 0)
        for (int w = 0; w < AN_THREADS / 32; w++) r += scratch[w];
    return r; // valid in thread 0
}

// scripts/temp_calc/main.cpp:70-92; then previous frame := current state (:112-117)
__global__ void __launch_bounds__(AN_THREADS) analysis_temperature_kernel(AnalysisArgs a)
{
    __shared__ double scratch[AN_THREADS / 32];
    const int traj = blockIdx.x;
    const size_t base = (size_t)traj * a.N;
    double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < a.N; i += AN_THREADS) {
        const float4 p = a.pos[base + i], q = a.ppos[base + i], u = a.ang[base + i], v = a.pang[base + i];
        // ang = {fi, psi, theta}: the angular DCD's X, Y, Z columns
        const double dx = (double)__fsub_rn(p.x, q.x), dy = (double)__fsub_rn(p.y, q.y), dz = (double)__fsub_rn(p.z, q.z);
        const double da = (double)__fsub_rn(u.x, v.x), db = (double)__fsub_rn(u.y, v.y), dg = (double)__fsub_rn(u.z, v.z);
        const double c = cos((double)u.y);
        s[0] += (dx * dx + dy * dy) + dz * dz;
        s[1] += ((da * da + db * db) + dg * dg) - ((2 * da) * dg) * (c * c);
        s[2] += dx * dx; s[3] += dy * dy; s[4] += dz * dz;
        s[5] += da * da; s[6] += db * db; s[7] += dg * dg;
        a.ppos[base + i] = p;
        a.pang[base + i] = u;
    }
#pragma unroll
    for (int q = 0; q < 8; q++) {
        const double r = block_sum(s[q], scratch);
        if (threadIdx.x == 0) a.temp[(size_t)traj * 8 + q] = r;
    }
}

// radius in the xy plane as the host tool rounds it: two products, one sum, one square root, each rounded to float
__device__ __forceinline__ float planar_radius(float x, float y) { return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }

// scripts/disas_speed/3d22d.cpp:42-51
__global__ void __launch_bounds__(AN_THREADS) analysis_project_kernel(AnalysisArgs a)
{
    const size_t n = (size_t)a.ntr * a.N;
    for (size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (size_t)gridDim.x * blockDim.x) {
        const float4 p = a.pos[q], u = a.ang[q];
        a.proj[q * 3 + 0] = planar_radius(p.x, p.y);
        a.proj[q * 3 + 1] = p.z;
        a.proj[q * 3 + 2] = u.z;
    }
}

// scripts/disas_speed/disc.cpp:62-117, per frame and protofilament, on the projection {X = radius, Y = z, Z = theta}
__global__ void __launch_bounds__(AN_THREADS) analysis_protofilament_kernel(AnalysisArgs a)
{
    __shared__ int pf_end[AN_MAX_PF], curled[AN_MAX_PF], chain_len[AN_MAX_PF];
    __shared__ unsigned long long tip[AN_MAX_PF];
    const int traj = blockIdx.x, N = a.N;
    const size_t base = (size_t)traj * N;
    if (threadIdx.x < AN_MAX_PF) {
        pf_end[threadIdx.x] = N;   // fill_n(pf_end_number, pf_number, atomCount), :65
        curled[threadIdx.x] = N;   // :66
        chain_len[threadIdx.x] = 0;
        tip[threadIdx.x] = 0ull;
    }
    __syncthreads();
    // :67-87  chain lengths; first break of every protofilament: consecutive atoms (ids differ by one) of the same
    // chain, residues differing by one, different monomer kinds, further apart than thres in the (radius, z) plane
    for (int i = threadIdx.x; i < N; i += AN_THREADS) {
        const int c = a.chain[i];
        if (c < 0 || c >= a.n_pf) continue;
        atomicAdd(&chain_len[c], 1);
        const int j = i + 1;
        if (j >= N || a.chain[j] != c) continue;
        const int ri = a.resid[i], rj = a.resid[j];
        if (abs(ri - rj) != 1 || a.name1[i] == a.name1[j]) continue;
        const float4 pi = a.pos[base + i], pj = a.pos[base + j];
        const double dX = (double)__fsub_rn(planar_radius(pi.x, pi.y), planar_radius(pj.x, pj.y));
        const double dY = (double)__fsub_rn(pi.z, pj.z);
        if (sqrt(dX * dX + dY * dY) > AN_THRES) atomicMin(&pf_end[c], min(ri, rj));
    }
    __syncthreads();
    if (threadIdx.x < a.n_pf && pf_end[threadIdx.x] > chain_len[threadIdx.x] / 2) pf_end[threadIdx.x] = chain_len[threadIdx.x] / 2; // :88-94
    __syncthreads();
    // :97-107  straight tip = the highest atom (first one on ties) at or below the break with theta < hor_thres;
    // curl start = lowest residue below the break with theta > theta_thres
    for (int i = threadIdx.x; i < N; i += AN_THREADS) {
        const int c = a.chain[i];
        if (c < 0 || c >= a.n_pf) continue;
        const int r = a.resid[i];
        const float z = a.pos[base + i].z, theta = a.ang[base + i].z;
        if (r <= pf_end[c] && z > 0.0f && (double)theta < AN_HOR_THRES)
            atomicMax(&tip[c], ((unsigned long long)__float_as_uint(z) << 32) | (unsigned long long)(0xffffffffu - (unsigned)i));
        if ((double)theta > AN_THETA_THRES && r < pf_end[c]) atomicMin(&curled[c], r);
    }
    __syncthreads();
    if (threadIdx.x < a.n_pf) {
        const int c = threadIdx.x;
        int *o = a.pf + ((size_t)traj * a.n_pf + c) * 3;
        o[0] = pf_end[c];
        o[1] = min(curled[c], pf_end[c]); // :119-124
        o[2] = tip[c] ? (int)a.resid[0xffffffffu - (unsigned)(tip[c] & 0xffffffffull)] : 0;
    }
}

// ensemble statistics of the per-trajectory energies: out[0..6] = sum over the local trajectories of each
// term, out[7..13] = sum of squares, out[14] = trajectory count, out[15] = 0.  One CTA; the result is all-reduced over
// the handles with NCCL (maddy_ensemble_stats_begin).
__global__ void __launch_bounds__(AN_THREADS) ensemble_stats_kernel(const double *__restrict__ en_traj, int ntr, double *__restrict__ out)
{
    __shared__ double scratch[AN_THREADS / 32];
    double s[14];
#pragma unroll
    for (int q = 0; q < 14; q++) s[q] = 0.0;
    for (int t = threadIdx.x; t < ntr; t += AN_THREADS) {
#pragma unroll
        for (int q = 0; q < 7; q++) {
            const double e = en_traj[(size_t)t * 7 + q];
            s[q] += e;
            s[7 + q] += e * e;
        }
    }
#pragma unroll
    for (int q = 0; q < 14; q++) {
        const double r = block_sum(s[q], scratch);
        if (threadIdx.x == 0) out[q] = r;
    }
    if (threadIdx.x == 0) {
        out[14] = (double)ntr;
        out[15] = 0.0;
    }
}

// mt_length() of the reference (updater.cpp:154-227) on the device, EXACTLY: on the tubule iff
//   rad < R_MT + R_THRES  &&  rad > 1.0  &&  cosf(theta) > cosf(ANG_THRES),   rad = sqrtf(x*x + y*y)  (float products, float sum).
// The radius part is IEEE arithmetic (__fmul_rn/__fadd_rn/__fsqrt_rn = what the host compiler emits without contraction).
// The cosine part cannot be evaluated here (the host's libm cosf is the authority), so the caller hands over the float
// thresholds at which ITS cosf crosses cosf(ANG_THRES) on each monotone branch (maddy_ontub_rule, bisected on the host):
// with a = |theta|, on iff a in [0, e0) U (e1, e2) U (e3, e4) ...; a >= a_max (several turns away) is reported as
// undecided in the status word.  One CTA per trajectory; flags to out_flags (and to live_flags when apply != 0), the
// per-trajectory count to out_count.
__global__ void __launch_bounds__(AN_THREADS) ontubule_kernel(const float4 *__restrict__ pos, const float4 *__restrict__ ang, int N, OnTubRule rule,
                                                              uint8_t *__restrict__ out_flags, uint8_t *__restrict__ live_flags, int apply,
                                                              int *__restrict__ out_count, int *__restrict__ status, int *__restrict__ guard)
{
    __shared__ int wsum[AN_THREADS / 32];
    const int traj = blockIdx.x;
    const size_t base = (size_t)traj * N;
    int sum = 0, undecided = 0;
    for (int i = threadIdx.x; i < N; i += AN_THREADS) {
        const float4 p = pos[base + i];
        const float theta = ang[base + i].z;
        const float rad = __fsqrt_rn(__fadd_rn(__fmul_rn(p.x, p.x), __fmul_rn(p.y, p.y)));
        bool on = rad < rule.rad_hi && (double)rad > 1.0;
        if (on) {
            const float a = fabsf(theta);
            if (!(a < rule.a_max)) { // also NaN
                undecided = 1;
                on = false;
            } else {
                // number of edges below or at a decides the branch: inside [0,e0): 0 edges passed -> on; (e0..e1]: off; ...
                bool in = a < rule.edge[0];
#pragma unroll
                for (int k = 1; k + 1 < ONTUB_EDGES; k += 2) in = in || (a > rule.edge[k] && a < rule.edge[k + 1]);
                on = in;
            }
        }
        out_flags[base + i] = on ? 1 : 0;
        if (apply) live_flags[base + i] = on ? 1 : 0;
        sum += on;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int w = 0; w < AN_THREADS / 32; w++) t += wsum[w];
        out_count[traj] = t;
    }
    if (undecided) {
        atomicOr(status, 1); // the word behind the counts: some |theta| was beyond the rule's range
        if (guard) atomicOr(guard, 1); // ... and what the caller queued behind this classification must not run on a guess
    }
}

cudaError_t launch_ontubule(const float4 *pos, const float4 *ang, int ntr, int N, const OnTubRule &rule, uint8_t *out_flags, uint8_t *live_flags,
                            int apply, int *out_count, int *status, int *guard, cudaStream_t st)
{
    ontubule_kernel<<<ntr, AN_THREADS, 0, st>>>(pos, ang, N, rule, out_flags, live_flags, apply, out_count, status, guard);
    return cudaGetLastError();
}

// change_conc()'s insertions (updater.cpp:118-135) as a sparse update (maddy_insert_dimers): dimer (q, q+1) leaves the
// reserve at the drawn position; the candidate lists of its trajectory refer to the old positions from here on.
__global__ void insert_dimers_kernel(float4 *__restrict__ pos, float4 *__restrict__ rpos, uint8_t *__restrict__ extra, int *__restrict__ cand_valid,
                                     int N, int n_insert, const int *__restrict__ index, const float4 *__restrict__ xyzz)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_insert) return;
    const int q = index[k];
    const float4 v = xyzz[k];
    const float4 a = make_float4(v.x, v.y, v.z, 0.f), b = make_float4(v.x, v.y, v.w, 0.f);
    pos[q] = a;
    pos[q + 1] = b;
    rpos[q] = a;
    rpos[q + 1] = b;
    extra[q] = 0;
    extra[q + 1] = 0;
    if (cand_valid) cand_valid[q / N] = 0;
}

cudaError_t launch_insert_dimers(float4 *pos, float4 *rpos, uint8_t *extra, int *cand_valid, int N, int n_insert, const int *index,
                                 const float4 *xyzz, cudaStream_t st)
{
    insert_dimers_kernel<<<(n_insert + 127) / 128, 128, 0, st>>>(pos, rpos, extra, cand_valid, N, n_insert, index, xyzz);
    return cudaGetLastError();
}

cudaError_t launch_ensemble_stats(const double *en_traj, int ntr, double *out, cudaStream_t st)
{
    ensemble_stats_kernel<<<1, AN_THREADS, 0, st>>>(en_traj, ntr, out);
    return cudaGetLastError();
}

cudaError_t launch_analysis(int which, const AnalysisArgs &a, cudaStream_t st)
{
    if (which == 0) analysis_temperature_kernel<<<a.ntr, AN_THREADS, 0, st>>>(a);
    else if (which == 1) {
        size_t b = ((size_t)a.ntr * a.N + AN_THREADS - 1) / AN_THREADS;
        analysis_project_kernel<<<(int)(b > 148 * 8 ? 148 * 8 : b), AN_THREADS, 0, st>>>(a);
    } else analysis_protofilament_kernel<<<a.ntr, AN_THREADS, 0, st>>>(a);
    return cudaGetLastError();
}

} // namespace maddy
