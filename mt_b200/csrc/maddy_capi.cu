/*
 * maddy_capi.cu — implementation of the C-ABI declared in include/maddy_b200.h.
 *
 * Replaces, for the hot path only, what compute() owns in the reference
 * (src/compute_cuda.cu:977-1260): device allocation and upload, the launch sequence, the
 * host<->device transfers around host events, and teardown.  No CPU fallback exists: every
 * compute entry point launches the sm_100a kernels of maddy_kernels.cu or fails.
 */
#include <cuda_runtime.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <thread>
#include <mutex>
#include <future>
#include <memory>
#include <dlfcn.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "maddy_kernels.cuh"
#include "maddy_lfib.h"

namespace maddy {
cudaError_t launch_run_kernel(const KArgs &k, int mpt, int ctas_per_sm, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_phase_kernel(const KArgs &k, int mpt, int threads, size_t smem, cudaStream_t st);
cudaError_t launch_integrate_kernel(const KArgs &k, cudaStream_t st);
cudaError_t launch_snapshot_kernel(const float4 *pos, const float4 *ang, float *out_mapped, size_t n_monomers, cudaStream_t st);
cudaError_t launch_consts_kernel(const maddy_params &p, StepConsts *d_out, cudaStream_t st);
cudaError_t launch_tea_kernels(const KArgs &k, int which, long long step, cudaStream_t st);
int tea_partner_segments(int N);
cudaError_t launch_hyd_stream(const uint32_t *W, const void *mats, unsigned long long count, uint32_t *out, const int *guard, cudaStream_t st);
cudaError_t launch_hyd_matrices(const void *table, void *mats, cudaStream_t st);
size_t hyd_matrices_bytes();
unsigned long long hyd_stream_capacity();
cudaError_t launch_hyd_prepare(const HydArgs &h, cudaStream_t st);
cudaError_t launch_hyd_plan(const HydArgs &h, int n_events, uint8_t *slots, bool prepared, cudaStream_t st);
cudaError_t launch_wide_publish(const KArgs &k, int buf, cudaStream_t st);
cudaError_t launch_wide_phase(const KArgs &k, int buf, cudaStream_t st);
cudaError_t launch_wide_step(const KArgs &k, int buf, cudaStream_t st);
cudaError_t launch_wide_run(const KArgs &k, int buf, int n_steps, int publish_first, cudaStream_t st);
cudaError_t launch_analysis(int which, const AnalysisArgs &a, cudaStream_t st);
cudaError_t launch_ensemble_stats(const double *en_traj, int ntr, double *out, cudaStream_t st);
cudaError_t launch_ontubule(const float4 *pos, const float4 *ang, int ntr, int N, const OnTubRule &rule, uint8_t *out_flags, uint8_t *live_flags,
                            int apply, int *out_count, int *status, int *guard, cudaStream_t st);
cudaError_t launch_insert_dimers(float4 *pos, float4 *rpos, uint8_t *extra, int *cand_valid, int N, int n_insert, const int *index,
                                 const float4 *xyzz, cudaStream_t st);
} // namespace maddy

using namespace maddy;

// The flag conversions and the SoA -> AoS read-back sit on the host's critical path between two fused windows; split
// them over a few threads when the ensemble is large (no-op when built without OpenMP or with OMP_NUM_THREADS=1).
#define HOST_PRAGMA(x) _Pragma(#x)
#define HOST_PARALLEL_FOR(n) HOST_PRAGMA(omp parallel for schedule(static) num_threads(host_threads()) if ((n) > 65536))
static int host_threads()
{
#ifdef _OPENMP
    static const int k = [] {
        // OMP_NUM_THREADS if it asks for several threads; launchers that pin it to 1 per rank (torchrun) still get a
        // share of a many-core host: one sixteenth of the hardware threads, at most 8
        int m = omp_get_max_threads();
        const int share = (int)(std::thread::hardware_concurrency() / 16);
        if (m < share) m = share;
        return m > 8 ? 8 : (m < 1 ? 1 : m);
    }();
    return k;
#else
    return 1;
#endif
}

struct LaunchCfg {
    int mpt = 1, threads = 32, ctas = 1, nbuf = 1, near_cap = 0, rng_off = 0, topo_off = -1;
    size_t smem = 0;
};

struct maddy_handle {
    maddy_params p;
    DevSys a;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    LaunchCfg phase, run; // step-granular phase kernel / fused run kernel
    StepConsts consts{};
    // GTP schedule (maddy_schedule_gtp)
    uint8_t *d_sched = nullptr;   // the schedule the next maddy_run reads (one of d_sched_buf)
    size_t sched_capacity = 0;
    // double-buffered, asynchronous upload: a schedule is converted into pinned memory and copied on the copy stream into
    // the buffer the running window does NOT read, so the host can hand over the next window's events beside the current one
    uint8_t *d_sched_buf[2] = {nullptr, nullptr};
    uint8_t *h_sched_buf[2] = {nullptr, nullptr};
    cudaEvent_t sched_copied[2] = {nullptr, nullptr};
    cudaEvent_t sched_reader[2] = {nullptr, nullptr}; // recorded behind the last run that read buffer b
    int sched_cur = 0;
    long long sched_first = 0, sched_period = 1;
    int sched_slots = 0;
    std::vector<uint16_t> amap, fmap;
    CutTest cut_pairs, cut_force;
    std::string err;
    long long launches = 0;
    AnalysisArgs an{};          // in-situ analysis buffers (maddy_analysis_setup), an.temp == nullptr until then
    bool an_has_reference = false;
    bool wide = false;          // trajectories spread over many CTAs, stage in HBM (maddy_wide.cuh): N > MADDY_MAX_NTOT_CTA
    bool lazy = false;          // fused loop keeps the Verlet list lazily (see ensure_lj)
    bool lj_maybe_stale = false; // some trajectory's Verlet list may have to be materialised before it is read
    std::vector<void *> allocs;
    int *h_status = nullptr; // pinned
    // ring of pinned staging buffers for the flag uploads: the copies are asynchronous, so the host can queue the
    // next fused window while the GPU is still running the current one
    static const int kStage = 4;
    uint8_t *stage[kStage] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t stage_done[kStage] = {nullptr, nullptr, nullptr, nullptr};
    int stage_next = 0;
    // asynchronous stride snapshot (maddy_snapshot_begin/_end): pinned host mirrors of state, forces and energies
    float4 *snap_pos = nullptr, *snap_ang = nullptr; // synchronous download path
    float *snap_r = nullptr, *snap_f = nullptr;      // pinned AoS-7 mirrors of the snapshot
    float *d_snap_r = nullptr, *d_snap_f = nullptr;  // device staging written by snapshot_kernel
    double *d_snap_en = nullptr;
    int *d_snap_status = nullptr;
    cudaStream_t copy_stream = nullptr;              // the D2H leg runs beside the next window
    cudaEvent_t snap_staged = nullptr;
    double *snap_en = nullptr;
    cudaEvent_t snap_done = nullptr;
    unsigned snap_what = 0;
    // ensemble statistics (maddy_ensemble_stats_begin/_end): 16 doubles on the device, pinned mirror, events
    double *d_ens = nullptr, *h_ens = nullptr;
    cudaEvent_t ens_ready = nullptr, ens_done = nullptr;
    bool ens_pending = false;
    // on-tubule classification with the snapshot (MADDY_SNAP_ONTUBULE): rule, device results, pinned mirrors
    OnTubRule ontub_rule{};
    bool ontub_rule_ok = false;
    uint8_t *d_cls_pair[2] = {nullptr, nullptr}; // the last two classifications (ping-pong): [cls_cur] newest, [cls_cur ^ 1] the one before
    int cls_cur = 0;                             // (both start as the flags given to maddy_create: updater.cpp's on_tubule_prev = on_tubule_cur at step 0)
    uint8_t *h_cls_flags = nullptr;
    int *d_cls_count = nullptr, *h_cls_count = nullptr; // [ntr + 1]: per-trajectory counts, then the undecided word
    unsigned cls_what = 0;                              // classification bits of the snapshot collected last
    cudaEvent_t cls_staged = nullptr, cls_done = nullptr; // counts written on the main stream / landed in h_cls_count
    // GTP flags with the snapshot (MADDY_SNAP_GTP)
    uint8_t *d_snap_gtp = nullptr, *h_snap_gtp = nullptr;
    // hydrolysis events of a stride on the device (maddy_hydrolysis_plan)
    uint8_t *d_hyd_own = nullptr, *d_hyd_all = nullptr; // this shard's transposed inputs / the working copy of every shard's (== own for one shard)
    size_t hyd_all_cap = 0;
    unsigned *d_hyd_rowcount = nullptr;
    unsigned long long *d_hyd_rowstart = nullptr, *d_hyd_counters = nullptr, *h_hyd_counters = nullptr; // [0] draws consumed, [1 + k] first draw of event k
    int hyd_counters_cap = 0;
    uint32_t *d_hyd_stream = nullptr, *d_hyd_window = nullptr, *h_hyd_window = nullptr;
    unsigned long long hyd_stream_cap = 0;
    void *d_lfib_table = nullptr, *d_hyd_mats = nullptr;
    int hyd_rows_cap = 0;
    int *d_hyd_status = nullptr, *h_hyd_status = nullptr;
    uint8_t *h_hyd_slots = nullptr;
    size_t hyd_slots_cap = 0;
    cudaEvent_t hyd_staged = nullptr, hyd_done = nullptr, hyd_in = nullptr;
    cudaStream_t aux_stream = nullptr; // the plan's kernels run beside the window that precedes its first event
    bool plan_wait = false;            // the main stream has not yet been made to wait for the plan
    int hyd_events = 0;      // events of the plan whose result is pending / was collected last
    bool hyd_pending = false, hyd_keep = false;
    // sparse insertions (maddy_insert_dimers): pinned + device record buffers {index, then float4 xyzz}, reuse event
    char *h_ins = nullptr, *d_ins = nullptr;
    size_t ins_capacity = 0;
    cudaEvent_t ins_done = nullptr;
    // MADDY_GPU_PROFILE=1: event pair around every kernel launch, summarised by maddy_destroy (development aid)
    bool gpu_prof = getenv("MADDY_GPU_PROFILE") != nullptr;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events;
    std::vector<unsigned> prof_ops;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_copy;
};

static thread_local std::string g_create_error;

static int fail(maddy_handle *h, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf;
    else g_create_error = buf;
    return code;
}

#define CU(h, call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess) return fail(h, MADDY_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));         \
    } while (0)

// All device arrays of a handle are carved out of ONE cudaMalloc (create/destroy sit inside the timed region of
// the drop-in compute()): requests are recorded first, then committed.
struct PoolReq {
    void **slot;
    size_t bytes;
};
static thread_local std::vector<PoolReq> *g_pool_reqs;
template <typename T> static void pool_req(T **p, size_t n)
{
    g_pool_reqs->push_back({(void **)p, ((n ? n : 1) * sizeof(T) + 255) & ~(size_t)255});
}
static int pool_commit(maddy_handle *h, std::vector<PoolReq> &reqs)
{
    size_t total = 0;
    for (const PoolReq &r : reqs) total += r.bytes;
    void *base = nullptr;
    cudaError_t e = cudaMalloc(&base, total);
    if (e != cudaSuccess) return fail(h, MADDY_ENOMEM, "cudaMalloc(%zu bytes): %s", total, cudaGetErrorString(e));
    h->allocs.push_back(base);
    char *q = (char *)base;
    for (const PoolReq &r : reqs) {
        *r.slot = q;
        q += r.bytes;
    }
    return MADDY_OK;
}

// smallest double s with float(sqrt(s)) >= c  (sqrt, the double->float conversion and the
// comparison are all monotone, so bisection over the bit pattern is exact)
static CutTest make_cut(float c)
{
    CutTest t;
    if (!(c > 0.f)) {
        t.t = 0.0;
        t.lo = t.hi = 0.f;
        return t;
    }
    uint64_t lo = 0, hi;
    double big = 4.0 * (double)c * (double)c;
    memcpy(&hi, &big, 8);
    while (hi - lo > 1) { // invariant: pred(lo) true ("inside"), pred(hi) false
        uint64_t mid = lo + (hi - lo) / 2;
        double s;
        memcpy(&s, &mid, 8);
        if ((float)sqrt(s) < c) lo = mid;
        else hi = mid;
    }
    memcpy(&t.t, &hi, 8);
    t.lo = (float)(t.t * (1.0 - 1e-6));
    t.hi = (float)(t.t * (1.0 + 1e-6));
    return t;
}

static KArgs kargs(const maddy_handle *h, unsigned ops)
{
    KArgs k;
    k.p = h->p;
    k.a = h->a;
    k.c = h->consts;
    k.barr_long_on = h->p.barrier && h->p.a_barr_long != 0.0f;
    k.barr_lat_on = h->p.barrier && h->p.a_barr_lat != 0.0f;
    k.first_step = 0;
    k.n_steps = 0;
    k.sched_first = h->sched_first;
    k.sched_period = h->sched_period;
    k.sched_slots = (ops & OP_RUN) ? h->sched_slots : 0;
    k.a.gtp_sched = h->d_sched;
    k.ops = ops;
    k.run_flags = 0;
    const LaunchCfg &c = (ops & OP_RUN) ? h->run : h->phase;
    k.nbuf = c.nbuf;
    k.near_cap = c.near_cap;
    k.rng_smem_offset = c.rng_off;
    k.topo_smem_offset = c.topo_off;
    {
        const float rb = fmaxf(h->p.lj_on ? h->p.ljpairscutoff : 0.f, 7.0f) + MD_CAND_SKIN;
        k.rcand2 = rb * rb;
    }
    k.lazy = (ops & OP_RUN) && h->lazy;
    k.near_all_listed = h->p.lj_on && h->p.ljpairscutoff >= 7.5f; // near radius 7.0 (MD_NEAR_R2)
    k.band_mid = 0.5f * (h->cut_force.lo + h->cut_force.hi);
    k.band_hw = h->cut_force.hi - h->cut_force.lo; // twice the half-width: a superset of the band whatever the rounding
    k.cut_pairs = h->cut_pairs;
    k.cut_force = h->cut_force;
    return k;
}

static int check_status(maddy_handle *h)
{
    // caller has synchronised the stream
    int st = *h->h_status;
    if (st) {
        *h->h_status = 0;
        cudaMemsetAsync(h->a.status, 0, sizeof(int), h->stream);
        return fail(h, MADDY_EOVERFLOW, "neighbour list overflow:%s%s%s (capacities LJ %d, longitudinal %d, lateral %d)",
                    (st & ST_LJ_OVERFLOW) ? " LJ" : "", (st & ST_LONG_OVERFLOW) ? " longitudinal" : "",
                    (st & ST_LAT_OVERFLOW) ? " lateral" : "", MADDY_LJ_CAPACITY, h->a.capLong, h->a.capLat);
    }
    return MADDY_OK;
}

static int sync_and_check(maddy_handle *h)
{
    CU(h, cudaMemcpyAsync(h->h_status, h->a.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return check_status(h);
}

// Wide path: the same operations as launch sequences over the whole GPU (maddy_wide.cuh).  A step-granular phase is
// [publish -> phase]; a fused window is cut at list-update steps and scheduled hydrolysis events into segments of ONE
// persistent cooperative launch each (wide_run_kernel: per-trajectory barrier per step, near list, state in registers);
// ensembles whose CTAs cannot all be resident, or MADDY_WIDE_PER_STEP=1, take one launch per step (wide_step_kernel).
static cudaError_t wide_dispatch(maddy_handle *h, const KArgs &k)
{
    cudaStream_t st = h->stream;
    cudaError_t e;
    if (!(k.ops & OP_RUN)) {
        KArgs kk = k;
        kk.ops &= OP_REBUILD_LJ | OP_REBUILD_BONDS | OP_FORCE | OP_ENERGY | OP_TEA_PREP; // nothing is ever lazy here: OP_MATERIALISE is a no-op
        if (!(kk.ops & ~(unsigned)OP_TEA_PREP)) return cudaSuccess;
        if ((e = launch_wide_publish(kk, 0, st)) != cudaSuccess) return e;
        h->launches += 1 + ((kk.ops & OP_ENERGY) ? 1 : 0);
        return launch_wide_phase(kk, 0, st);
    }
    const maddy_params &p = h->p;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    const unsigned rops = (p.lj_on ? OP_REBUILD_LJ : 0u) | (p.is_assembly ? OP_REBUILD_BONDS : 0u);
    KArgs ks = k, kr = k;
    ks.ops = OP_FORCE;
    kr.ops = rops;
    const long long end = k.first_step + k.n_steps;
    const long long freq = p.ljpairsupdatefreq > 0 ? p.ljpairsupdatefreq : 1;
    auto event_at = [&](long long step) { // scheduled hydrolysis event (maddy_schedule_gtp) at this step: its slot, or -1
        if (k.sched_slots <= 0 || step < k.sched_first || (step - k.sched_first) % k.sched_period != 0) return -1LL;
        const long long slot = (step - k.sched_first) / k.sched_period;
        return slot < k.sched_slots ? slot : -1LL;
    };
    auto rebuild_at = [&](long long step) {
        return rops != 0 && step % freq == 0 && !(step == k.first_step && (k.run_flags & MADDY_RUN_SKIP_FIRST_REBUILD));
    };
    int buf = 0;
    bool staged = false; // stage[buf] holds the current state
    bool persistent = !getenv("MADDY_WIDE_PER_STEP");
    long long step = k.first_step;
    while (step < end) {
        // scheduled hydrolysis event: the flags of this slot become current (the stage carries them: re-publish)
        const long long slot = event_at(step);
        if (slot >= 0) {
            e = cudaMemcpyAsync(h->a.gtp, h->d_sched + (size_t)slot * n, n, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) return e;
            staged = false;
        }
        if (rebuild_at(step)) {
            if (!staged) {
                if ((e = launch_wide_publish(ks, buf, st)) != cudaSuccess) return e;
                h->launches++;
                staged = true;
            }
            if ((e = launch_wide_phase(kr, buf, st)) != cudaSuccess) return e;
            h->launches++;
        }
        // steps up to the next list-update step / scheduled event: one persistent launch
        long long seg_end = end;
        if (rops != 0) seg_end = std::min(seg_end, (step / freq + 1) * freq);
        if (k.sched_slots > 0) {
            for (long long q = step + 1; q < seg_end; q++) // (segments are at most one list-update period long)
                if (event_at(q) >= 0) {
                    seg_end = q;
                    break;
                }
        }
        if (rops == 0 && seg_end - step > 4096) seg_end = step + 4096;
        const int count = (int)(seg_end - step);
        if (persistent) {
            e = launch_wide_run(ks, buf, count, staged ? 0 : 1, st);
            if (e == cudaErrorCooperativeLaunchTooLarge) {
                persistent = false; // the ensemble does not fit the GPU at once: one launch per step
                (void)cudaGetLastError();
            } else if (e != cudaSuccess) {
                return e;
            } else {
                h->launches++;
                buf ^= count & 1;
                staged = true;
                step = seg_end;
                continue;
            }
        }
        if (!staged) {
            if ((e = launch_wide_publish(ks, buf, st)) != cudaSuccess) return e;
            h->launches++;
            staged = true;
        }
        for (; step < seg_end; step++) {
            if ((e = launch_wide_step(ks, buf, st)) != cudaSuccess) return e;
            h->launches++;
            buf ^= 1;
        }
    }
    return cudaSuccess;
}

static int launch(maddy_handle *h, const KArgs &k)
{
    CU(h, cudaSetDevice(h->p.device));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (h->gpu_prof) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, h->stream);
    }
    cudaError_t e = h->wide ? wide_dispatch(h, k)
                    : (k.ops & OP_RUN) ? launch_run_kernel(k, h->run.mpt, h->run.ctas, h->run.threads, h->run.smem, h->stream)
                                       : launch_phase_kernel(k, h->phase.mpt, h->phase.threads, h->phase.smem, h->stream);
    if (h->gpu_prof) {
        cudaEventRecord(e1, h->stream);
        h->prof_events.push_back({e0, e1});
        h->prof_ops.push_back(k.ops);
    }
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "trajectory kernel launch (ops=%u): %s", k.ops, cudaGetErrorString(e));
    h->launches++;
    return MADDY_OK;
}

// ---- AoS7 {x,y,z,fi,theta,psi,w}  <->  float4 {x,y,z,0} + {fi,psi,theta,0}
static void aos_to_soa(const float *aos, size_t n, std::vector<float4> &pos, std::vector<float4> &ang, bool wrap)
{
    pos.resize(n);
    ang.resize(n);
    for (size_t q = 0; q < n; q++) {
        const float *c = aos + q * MADDY_COORD_STRIDE;
        float fi = c[3], theta = c[4], psi = c[5];
        if (wrap) { // compute_cuda.cu:1004-1010: truncation toward zero, in double
            fi -= (2 * M_PI) * (int)(fi / (2 * M_PI));
            psi -= (2 * M_PI) * (int)(psi / (2 * M_PI));
            theta -= (2 * M_PI) * (int)(theta / (2 * M_PI));
        }
        pos[q] = make_float4(c[0], c[1], c[2], 0.f);
        ang[q] = make_float4(fi, psi, theta, 0.f);
    }
}
static void soa_to_aos(const std::vector<float4> &pos, const std::vector<float4> &ang, float *aos)
{
    for (size_t q = 0; q < pos.size(); q++) {
        float *c = aos + q * MADDY_COORD_STRIDE;
        c[0] = pos[q].x; c[1] = pos[q].y; c[2] = pos[q].z;
        c[3] = ang[q].x; c[4] = ang[q].z; c[5] = ang[q].y;
        c[6] = 0.f;
    }
}

extern "C" const char *maddy_last_error(const maddy_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }
extern "C" void *maddy_stream(const maddy_handle *h) { return h ? (void *)h->stream : nullptr; }
extern "C" long long maddy_launch_count(const maddy_handle *h) { return h ? h->launches : 0; }

extern "C" int maddy_list_stats(maddy_handle *h, unsigned long long out[4], int reset)
{
    if (!h || !out) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    CU(h, cudaMemcpyAsync(out, h->a.stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    if (reset) CU(h, cudaMemsetAsync(h->a.stats, 0, 4 * sizeof(unsigned long long), h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}

extern "C" int maddy_sync(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    return sync_and_check(h);
}

extern "C" int maddy_destroy(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    cudaSetDevice(h->p.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->gpu_prof && !h->prof_events.empty()) {
        double busy = 0;
        float ms = 0, span = 0;
        for (auto &pr : h->prof_events) {
            cudaEventElapsedTime(&ms, pr.first, pr.second);
            busy += ms;
        }
        cudaEventElapsedTime(&span, h->prof_events.front().first, h->prof_events.back().second);
        for (size_t q = 0; q < h->prof_copy.size() && q < 12; q++) {
            cudaEventElapsedTime(&ms, h->prof_copy[q].first, h->prof_copy[q].second);
            fprintf(stderr, "[maddy gpu profile] snapshot copy segment %zu: %.3f ms\n", q, ms);
        }
        double gsum[4] = {0, 0, 0, 0};
        int gcnt[4] = {0, 0, 0, 0};
        for (size_t q = 1; q < h->prof_events.size(); q++) {
            cudaEventElapsedTime(&ms, h->prof_events[q - 1].second, h->prof_events[q].first);
            const int b = ms < 0.02f ? 0 : ms < 0.1f ? 1 : ms < 0.5f ? 2 : 3;
            if (b == 3 && gcnt[3] < 6) {
                float k0 = 0, k1 = 0;
                cudaEventElapsedTime(&k0, h->prof_events[q - 1].first, h->prof_events[q - 1].second);
                cudaEventElapsedTime(&k1, h->prof_events[q].first, h->prof_events[q].second);
                fprintf(stderr, "[maddy gpu profile] gap %.3f ms between launch %zu (ops %u, %.3f ms) and %zu (ops %u, %.3f ms)\n", ms, q - 1,
                        h->prof_ops[q - 1], k0, q, h->prof_ops[q], k1);
            }
            gsum[b] += ms;
            gcnt[b]++;
        }
        fprintf(stderr, "[maddy gpu profile] gaps <20us: %d (%.2f ms)  20-100us: %d (%.2f ms)  0.1-0.5ms: %d (%.2f ms)  >0.5ms: %d (%.2f ms)\n",
                gcnt[0], gsum[0], gcnt[1], gsum[1], gcnt[2], gsum[2], gcnt[3], gsum[3]);
        fprintf(stderr, "[maddy gpu profile] %zu launches, kernels busy %.3f ms, first-to-last span %.3f ms (idle %.3f ms)\n",
                h->prof_events.size(), busy, (double)span, (double)span - busy);
        for (auto &pr : h->prof_events) {
            cudaEventDestroy(pr.first);
            cudaEventDestroy(pr.second);
        }
    }
    for (void *q : h->allocs) cudaFree(q);
    if (h->h_status) cudaFreeHost(h->h_status);
    for (int b = 0; b < 2; b++) {
        if (h->d_sched_buf[b]) cudaFree(h->d_sched_buf[b]);
        if (h->h_sched_buf[b]) cudaFreeHost(h->h_sched_buf[b]);
        if (h->sched_copied[b]) cudaEventDestroy(h->sched_copied[b]);
        if (h->sched_reader[b]) cudaEventDestroy(h->sched_reader[b]);
    }
    for (int k = 0; k < maddy_handle::kStage; k++) {
        if (h->stage[k]) cudaFreeHost(h->stage[k]);
        if (h->stage_done[k]) cudaEventDestroy(h->stage_done[k]);
    }
    for (void *q : {(void *)h->snap_pos, (void *)h->snap_ang, (void *)h->snap_r, (void *)h->snap_f, (void *)h->snap_en})
        if (q) cudaFreeHost(q);
    if (h->snap_done) cudaEventDestroy(h->snap_done);
    if (h->snap_staged) cudaEventDestroy(h->snap_staged);
    for (void *q : {(void *)h->d_cls_pair[0], (void *)h->d_cls_pair[1], (void *)h->d_snap_gtp, (void *)h->d_hyd_own, (void *)(h->d_hyd_all != h->d_hyd_own ? h->d_hyd_all : nullptr), (void *)h->d_hyd_rowcount,
                    (void *)h->d_hyd_rowstart, (void *)h->d_hyd_counters, (void *)h->d_hyd_stream, (void *)h->d_hyd_window, h->d_lfib_table, h->d_hyd_mats,
                    (void *)h->d_hyd_status})
        if (q) cudaFree(q);
    for (void *q : {(void *)h->h_snap_gtp, (void *)h->h_hyd_counters, (void *)h->h_hyd_window, (void *)h->h_hyd_status, (void *)h->h_hyd_slots})
        if (q) cudaFreeHost(q);
    if (h->hyd_staged) cudaEventDestroy(h->hyd_staged);
    if (h->hyd_in) cudaEventDestroy(h->hyd_in);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    if (h->hyd_done) cudaEventDestroy(h->hyd_done);
    if (h->cls_staged) cudaEventDestroy(h->cls_staged);
    if (h->cls_done) cudaEventDestroy(h->cls_done);
    if (h->ins_done) cudaEventDestroy(h->ins_done);
    if (h->d_ins) cudaFree(h->d_ins);
    if (h->h_ins) cudaFreeHost(h->h_ins);
    if (h->d_cls_count) cudaFree(h->d_cls_count);
    if (h->h_cls_flags) cudaFreeHost(h->h_cls_flags);
    if (h->h_cls_count) cudaFreeHost(h->h_cls_count);
    if (h->ens_ready) cudaEventDestroy(h->ens_ready);
    if (h->ens_done) cudaEventDestroy(h->ens_done);
    if (h->d_ens) cudaFree(h->d_ens);
    if (h->h_ens) cudaFreeHost(h->h_ens);
    if (h->copy_stream) {
        cudaStreamSynchronize(h->copy_stream);
        cudaStreamDestroy(h->copy_stream);
    }
    for (void *q : {(void *)h->d_snap_r, (void *)h->d_snap_f, (void *)h->d_snap_en, (void *)h->d_snap_status})
        if (q) cudaFree(q);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return MADDY_OK;
}

extern "C" int maddy_upload_list(maddy_handle *h, int kind, const int *counts, const int *entries);
static bool process_ontub_rule(OnTubRule &r);
static int cls_pair_ready(maddy_handle *h);

// The HybridTaus seed table of an ensemble (HybridTaus.cu:21-48): ran2 is sequential (9 ns per draw, 10 ms at 520 x 256), so
// the table is generated once per process and ensemble - not once per shard - and on a thread of its own while
// maddy_create allocates and uploads.
static std::shared_ptr<std::vector<unsigned>> seed_table(int rseed, long long np)
{
    static std::mutex seed_mutex;
    static std::shared_ptr<std::vector<unsigned>> seed_cache;
    static int seed_cache_rseed = 0;
    std::lock_guard<std::mutex> lock(seed_mutex);
    if (!seed_cache || seed_cache_rseed != rseed || (long long)seed_cache->size() != np * 4) {
        seed_cache = std::make_shared<std::vector<unsigned>>((size_t)np * 4);
        maddy_generate_seeds(seed_cache->data(), rseed, np);
        seed_cache_rseed = rseed;
    }
    return seed_cache;
}

extern "C" int maddy_create(const maddy_params *par, const maddy_topology *top, const float *coords, void *stream,
                            maddy_handle **out)
{
    if (!par || !top || !coords || !out) return fail(nullptr, MADDY_EINVAL, "maddy_create: null argument");
    if (par->abi_version != MADDY_ABI_VERSION)
        return fail(nullptr, MADDY_EINVAL, "maddy_create: abi_version %d, library has %d", par->abi_version, MADDY_ABI_VERSION);
    const int N = par->n_tot, ntr = par->n_tr_local;
    if (N <= 0 || N > MADDY_MAX_NTOT)
        return fail(nullptr, MADDY_EINVAL, "maddy_create: n_tot=%d outside [1,%d]", N, MADDY_MAX_NTOT);
    const bool wide = N > MADDY_MAX_NTOT_CTA || getenv("MADDY_FORCE_WIDE") != nullptr;
    if (wide && par->n_tr_local > 65535)
        return fail(nullptr, MADDY_EINVAL, "maddy_create: the wide path takes at most 65535 trajectories per handle (%d asked)", par->n_tr_local);
    if (ntr <= 0 || par->traj_first < 0 || par->traj_first + ntr > par->n_tr)
        return fail(nullptr, MADDY_EINVAL, "maddy_create: bad shard [%d,%d) of %d trajectories", par->traj_first,
                    par->traj_first + ntr, par->n_tr);
    if (par->max_harmonic < 1 || par->max_longitudinal < 0 || par->max_lateral < 0 || par->max_longitudinal > 255 ||
        par->max_lateral > 255)
        return fail(nullptr, MADDY_EINVAL, "maddy_create: bad list capacities (harmonic %d, longitudinal %d, lateral %d)",
                    par->max_harmonic, par->max_longitudinal, par->max_lateral);
    if (par->ljpairsupdatefreq <= 0) return fail(nullptr, MADDY_EINVAL, "maddy_create: ljpairsupdatefreq must be > 0");
    if (par->tea_on && par->tea_epsilon_freq <= 0) return fail(nullptr, MADDY_EINVAL, "maddy_create: tea_epsilon_freq must be > 0");

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MADDY_ECUDA, "maddy_create: no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (par->device < 0 || par->device >= ndev) return fail(nullptr, MADDY_EINVAL, "maddy_create: device %d of %d", par->device, ndev);

    std::future<std::shared_ptr<std::vector<unsigned>>> seed_future =
        std::async(std::launch::async, seed_table, par->rseed, 2LL * N * par->n_tr);
    maddy_handle *h = new maddy_handle;
    h->p = *par;
    int rc = MADDY_OK;
#define CK(x)                                   \
    do {                                        \
        rc = (x);                               \
        if (rc != MADDY_OK) goto bad;           \
    } while (0)
#define CUK(call)                                                                              \
    do {                                                                                       \
        cudaError_t e_ = (call);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            rc = fail(h, MADDY_ECUDA, "%s: %s", #call, cudaGetErrorString(e_));                \
            goto bad;                                                                          \
        }                                                                                      \
    } while (0)
    {
        CUK(cudaSetDevice(par->device));
        if (stream) h->stream = (cudaStream_t)stream;
        else {
            CUK(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
            h->own_stream = true;
        }
        CUK(cudaMallocHost(&h->h_status, 2 * sizeof(int))); // [0] synchronous checks, [1] snapshot
        h->h_status[0] = h->h_status[1] = 0;

        DevSys &a = h->a;
        memset(&a, 0, sizeof a);
        a.N = N;
        a.Npad = (N + 31) & ~31;
        a.ntr = ntr;
        a.maxH = par->max_harmonic;
        a.capLong = par->max_longitudinal;
        a.capLat = par->max_lateral;
        const size_t n = (size_t)ntr * N;

        // ---- launch geometry.  64 B of stage per monomer; near list 2 B per row and monomer.
        const bool no_near = getenv("MADDY_NO_NEAR") != nullptr;
        const size_t tiles = (size_t)2 * 16 * ((N + MD_TILE - 1) / MD_TILE);
        auto layout = [&](LaunchCfg &c, int nbuf, int cap, bool rng, bool topo) {
            c.nbuf = nbuf;
            c.near_cap = no_near ? 0 : cap;
            size_t sm = (size_t)nbuf * 64 * N + tiles + (size_t)c.near_cap * N * 2 + N + 64;
            sm = (sm + 15) & ~(size_t)15;
            c.rng_off = rng ? (int)sm : 0;
            if (rng) sm += (size_t)32 * N;
            c.topo_off = topo ? (int)sm : -1;
            if (topo) sm += (size_t)16 * N;
            c.smem = sm;
        };
        auto best_cap = [&](int nbuf, size_t extra, size_t budget) {
            const size_t fixed = (size_t)nbuf * 64 * N + tiles + N + 80 + extra;
            if (fixed + (size_t)12 * 2 * N > budget) return 0; // too little room: all-pairs path, lists from HBM only
            size_t cap = (budget - fixed) / ((size_t)2 * N);
            if (const char *e = getenv("MADDY_NEAR_CAP")) { // test hook: a tiny cache makes crowded-monomer overflows common
                const size_t want = (size_t)atoi(e);
                if (want >= 1 && want < cap) cap = want;
            }
            return (int)(cap > 32 ? 32 : cap);
        };
        h->wide = wide;
        if (wide) {
            // wide path (maddy_wide.cuh): one thread per monomer over many CTAs, stage in HBM, lists walked from HBM
            h->phase = LaunchCfg();
            h->run = LaunchCfg();
            h->amap.resize(N);
            for (int i = 0; i < N; i++) h->amap[i] = (uint16_t)i;
            a.n_active = N;
            a.n_fixed = 0;
        } else {
        // (1) phase kernel (step-granular entry points): every monomer owns a thread, one stage buffer
        h->phase.mpt = (N + MD_MAX_THREADS - 1) / MD_MAX_THREADS;
        if (h->phase.mpt > MD_MAX_MPT) {
            rc = fail(h, MADDY_EINVAL, "n_tot=%d needs %d monomers per thread (max %d)", N, h->phase.mpt, MD_MAX_MPT);
            goto bad;
        }
        h->phase.threads = (((N + h->phase.mpt - 1) / h->phase.mpt) + 31) & ~31;
        layout(h->phase, 1, best_cap(1, 0, 200 * 1024), false, false);
        // (2) run kernel (fused loop): threads for the monomers that can move
        {
            std::vector<uint16_t> amap, fmap;
            for (int i = 0; i < N; i++) (top->fixed[i] ? fmap : amap).push_back((uint16_t)i);
            if (fmap.size() > 256 || getenv("MADDY_NO_COMPACT")) { // unusual: give every monomer a thread
                amap.resize(N);
                for (int i = 0; i < N; i++) amap[i] = (uint16_t)i;
                fmap.clear();
            }
            h->amap = amap;
            h->fmap = fmap;
            a.n_active = (int)amap.size();
            a.n_fixed = (int)fmap.size();
            const int na = a.n_active > 0 ? a.n_active : 1;
            h->run.mpt = (na + MD_RUN_THREADS - 1) / MD_RUN_THREADS;
            if (h->run.mpt > MD_RUN_MAX_MPT) {
                rc = fail(h, MADDY_EINVAL, "n_tot=%d needs %d monomers per thread in the fused loop (max %d)", N, h->run.mpt, MD_RUN_MAX_MPT);
                goto bad;
            }
            h->run.threads = (((na + h->run.mpt - 1) / h->run.mpt) + 31) & ~31;
            if (h->run.threads < a.n_fixed) h->run.threads = (a.n_fixed + 31) & ~31;
            int n_sm = 148;
            cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, par->device);
            const char *force = getenv("MADDY_CTAS_PER_SM");
            // a second CTA per SM pays when there are more trajectories than SMs (it halves the register budget)
            const bool want2 = h->run.mpt == 1 && (force ? force[0] == '2' : ntr > n_sm);
            h->run.ctas = 1;
            if (want2) {
                // per CTA: double-buffered stage, near list of <= 16 rows, RNG streams and topology words in SMEM
                const size_t per_cta = (227 * 1024) / 2 - 1792 - 1024;
                const int cap = best_cap(2, (size_t)48 * N, per_cta);
                if (cap >= 12 || no_near) {
                    layout(h->run, 2, cap > 16 ? 16 : cap, true, true);
                    if (h->run.smem <= per_cta) h->run.ctas = 2;
                }
            }
            if (h->run.ctas == 1) {
                const int nbuf = ((size_t)2 * 64 * N <= 100 * 1024) ? 2 : 1;
                const bool topo = (size_t)nbuf * 64 * N + tiles + (size_t)12 * 2 * N + (size_t)17 * N + 2048 <= 200 * 1024;
                layout(h->run, nbuf, best_cap(nbuf, topo ? (size_t)16 * N : 0, 200 * 1024), false, topo);
            }
        }
        } // !wide
        h->cut_pairs = make_cut(par->ljpairscutoff);
        h->cut_force = make_cut(MD_LJ_FORCE_CUTOFF);

        std::vector<PoolReq> reqs;
        g_pool_reqs = &reqs;
        pool_req(&a.pos, n);
        pool_req(&a.ang, n);
        pool_req(&a.fpos, n);
        pool_req(&a.fang, n);
        pool_req(&a.rng_xyz, n);
        pool_req(&a.rng_ang, n);
        pool_req(const_cast<int **>(&a.harm), (size_t)N * a.maxH);
        pool_req(const_cast<int **>(&a.harm_count), (size_t)N);
        pool_req(const_cast<uint8_t **>(&a.sflags), (size_t)N);
        pool_req(&a.extra, n);
        pool_req(&a.gtp, n);
        pool_req(&a.ontub, n);
        pool_req(&a.bl, (size_t)ntr * (a.capLong + a.capLat) * a.Npad);
        pool_req(&a.bcnt, (size_t)ntr * 2 * a.Npad);
        pool_req(&a.lj, par->lj_on ? (size_t)ntr * MADDY_LJ_CAPACITY * a.Npad : 1);
        pool_req(&a.ljcnt, (size_t)ntr * a.Npad);
        pool_req(&a.cand, (h->run.near_cap > 0 || h->phase.near_cap > 0) ? (size_t)ntr * MD_CAND_CAPACITY * a.Npad : 1);
        pool_req(const_cast<uint16_t **>(&a.amap), h->amap.size());
        pool_req(const_cast<uint16_t **>(&a.fmap), h->fmap.size());
        pool_req(&a.candcnt, (size_t)ntr * a.Npad);
        pool_req(&a.ncand, (h->run.near_cap > 0 || h->phase.near_cap > 0) ? (size_t)ntr * MD_NCAND_CAPACITY * a.Npad : 1);
        pool_req(&a.ncandcnt, (size_t)ntr * a.Npad);
        pool_req(&a.rpos, n);
        pool_req(&a.lj_stale, (size_t)ntr);
        pool_req(&a.cpos, n);
        pool_req(&a.cand_valid, (size_t)ntr);
        pool_req(&a.en_mono, n * 7);
        pool_req(&a.en_traj, (size_t)ntr * 7);
        pool_req(&a.status, 1);
        pool_req(&a.guard, 1);
        pool_req(&a.stats, 4);
        if (par->tea_on) {
            pool_req(&a.tea_ci, n);
            pool_req(&a.tea_eps, n);
            pool_req(&a.tea_beta, (size_t)ntr);
            pool_req(&a.tea_co, n);
            pool_req(&a.tea_mf, n);
            pool_req(&a.tea_rf, n);
            pool_req(&a.tea_part, tea_partner_segments(N) > 1 ? n * tea_partner_segments(N) : 1);
            pool_req(&a.tea_cnt, (size_t)ntr * ((N + 31) / 32));
        }
        if (wide) pool_req(&a.gstage, 8 * n);
        if (wide) pool_req(&a.wbar, (size_t)ntr * 4);
        if (wide && !getenv("MADDY_WIDE_ALL_PAIRS")) // WGrid header + count/start/cursor[32768] + members[Npad] per trajectory (maddy_wide.cuh)
            pool_req(reinterpret_cast<char **>(&a.wgrid), (size_t)ntr * (64 + (size_t)3 * 32768 * 4 + (size_t)a.Npad * 2));
        CK(pool_commit(h, reqs));
        if (par->tea_on) CUK(cudaMemsetAsync(a.tea_cnt, 0, (size_t)ntr * ((N + 31) / 32) * sizeof(unsigned), h->stream));
        CUK(cudaMemsetAsync(a.cand_valid, 0, (size_t)ntr * sizeof(int), h->stream));
        CUK(cudaMemsetAsync(a.lj_stale, 0, (size_t)ntr * sizeof(int), h->stream));
        CUK(cudaMemsetAsync(a.status, 0, sizeof(int), h->stream));
        CUK(cudaMemsetAsync(a.guard, 0, sizeof(int), h->stream));
        CUK(cudaMemsetAsync(a.stats, 0, 4 * sizeof(unsigned long long), h->stream));
        CUK(cudaMemsetAsync(a.fpos, 0, n * sizeof(float4), h->stream));
        CUK(cudaMemsetAsync(a.fang, 0, n * sizeof(float4), h->stream));
        CUK(cudaMemsetAsync(a.bcnt, 0, (size_t)ntr * 2 * a.Npad, h->stream));
        CUK(cudaMemsetAsync(a.bl, 0, (size_t)ntr * (a.capLong + a.capLat) * a.Npad * sizeof(uint16_t), h->stream)); // rows beyond the counts are never read; zeroed so that the read-modify-write of maddy_upload_list moves defined bytes
        CUK(cudaMemsetAsync(a.ljcnt, 0, (size_t)ntr * a.Npad * sizeof(uint16_t), h->stream));
        CUK(cudaMemsetAsync(a.en_mono, 0, n * 7 * sizeof(double), h->stream));

        // coordinates
        std::vector<float4> pos, ang;
        aos_to_soa(coords, n, pos, ang, true);
        CUK(cudaMemcpyAsync(a.pos, pos.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CUK(cudaMemcpyAsync(a.ang, ang.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
        CUK(cudaMemcpyAsync(a.rpos, a.pos, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
        // Lazy Verlet list in the fused loop: every pair inside the near radius must be certain to be listed, whatever
        // the monomers did since the candidate list was built (7.0 + 4 x 0.74 < cut-off), and both kernels need candidates.
        h->lazy = par->lj_on && par->ljpairscutoff >= 10.0f && h->run.near_cap > 0 && h->phase.near_cap > 0 && !getenv("MADDY_NO_LAZY");

        // static topology
        std::vector<uint8_t> sflags(N);
        for (int i = 0; i < N; i++) sflags[i] = (uint8_t)((top->fixed[i] ? 1 : 0) | ((top->mon_type[i] & 0x7f) << 1));
        CUK(cudaMemcpyAsync((void *)a.sflags, sflags.data(), N, cudaMemcpyHostToDevice, h->stream));
        if (!h->amap.empty()) CUK(cudaMemcpyAsync((void *)a.amap, h->amap.data(), h->amap.size() * 2, cudaMemcpyHostToDevice, h->stream));
        if (!h->fmap.empty()) CUK(cudaMemcpyAsync((void *)a.fmap, h->fmap.data(), h->fmap.size() * 2, cudaMemcpyHostToDevice, h->stream));
        CUK(cudaMemcpyAsync((void *)a.harm, top->harmonic, (size_t)N * a.maxH * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        CUK(cudaMemcpyAsync((void *)a.harm_count, top->harmonic_count, (size_t)N * sizeof(int), cudaMemcpyHostToDevice, h->stream));
        for (int i = 0; i < N; i++) {
            for (int kk = 0; kk < top->harmonic_count[i]; kk++) {
                int j = abs(top->harmonic[(size_t)i * a.maxH + kk]);
                if (j >= N || top->harmonic_count[i] > a.maxH) {
                    rc = fail(h, MADDY_EINVAL, "harmonic list of monomer %d is out of range", i);
                    goto bad;
                }
            }
        }
        CUK(cudaStreamSynchronize(h->stream));

        // per-trajectory flags
        CK(maddy_upload_extra(h, top->extra));
        CK(maddy_upload_gtp(h, top->gtp));
        CK(maddy_upload_on_tubule(h, top->on_tubule_cur));
        CK(cls_pair_ready(h)); // hydrolysis on the device reads the last two classifications: both start as these flags
        if (a.capLong > 0 && top->longitudinal) CK(maddy_upload_list(h, MADDY_LIST_LONGITUDINAL, top->longitudinal_count, top->longitudinal));
        if (a.capLat > 0 && top->lateral) CK(maddy_upload_list(h, MADDY_LIST_LATERAL, top->lateral_count, top->lateral));

        // per-run constants, evaluated on the device (fast-math division / sqrt)
        {
            StepConsts *d_c = reinterpret_cast<StepConsts *>(a.en_traj); // scratch: not yet in use
            cudaError_t ec = launch_consts_kernel(*par, d_c, h->stream);
            if (ec != cudaSuccess) {
                rc = fail(h, MADDY_ECUDA, "consts kernel: %s", cudaGetErrorString(ec));
                goto bad;
            }
            CUK(cudaMemcpyAsync(&h->consts, d_c, sizeof(StepConsts), cudaMemcpyDeviceToHost, h->stream));
            CUK(cudaStreamSynchronize(h->stream));
        }

        // RNG: the GLOBAL table of 2*Ntot*Ntr states, sliced (HybridTaus.cu:21-31, compute_cuda.cu:1097)
        {
            // (generated beside the allocations and uploads above: seed_future, started at the top of maddy_create)
            std::shared_ptr<std::vector<unsigned>> table = seed_future.get();
            const std::vector<unsigned> &seeds = *table;
            const size_t off_xyz = (size_t)par->traj_first * N;
            const size_t off_ang = (size_t)N * par->n_tr + off_xyz;
            CUK(cudaMemcpyAsync(a.rng_xyz, seeds.data() + off_xyz * 4, n * sizeof(uint4), cudaMemcpyHostToDevice, h->stream));
            CUK(cudaMemcpyAsync(a.rng_ang, seeds.data() + off_ang * 4, n * sizeof(uint4), cudaMemcpyHostToDevice, h->stream));
            CUK(cudaStreamSynchronize(h->stream));
        }
    }
    {
        // exact on-tubule rule of this process's cosf (MADDY_SNAP_ONTUBULE): bisected once
        h->ontub_rule_ok = process_ontub_rule(h->ontub_rule);
        if (const char *e = getenv("MADDY_ONTUB_AMAX")) { // (the cached rule was made once per process: the hook is per handle)
            const float v = (float)atof(e);
            if (v > 0.f && v < h->ontub_rule.a_max) h->ontub_rule.a_max = v;
        }
    }
    *out = h;
    return MADDY_OK;
bad:
    g_create_error = h->err;
    maddy_destroy(h);
    return rc;
#undef CK
#undef CUK
}

// ------------------------------------------------------------------ step-granular entry points
// The fused loop may leave the Verlet list of its last list-update step unwritten (KArgs::lazy); anything that reads
// the list, or invalidates the candidates it would be derived from, calls this first.
static int ensure_lj(maddy_handle *h)
{
    if (!h->lj_maybe_stale) return MADDY_OK;
    h->lj_maybe_stale = false;
    return launch(h, kargs(h, OP_MATERIALISE));
}

extern "C" int maddy_rebuild_lj(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    if (!h->p.lj_on) return MADDY_OK;
    h->lj_maybe_stale = false; // the kernel writes every trajectory's list and clears its flag
    return launch(h, kargs(h, OP_REBUILD_LJ));
}
extern "C" int maddy_rebuild_bonds(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    return launch(h, kargs(h, OP_REBUILD_BONDS));
}
extern "C" int maddy_force(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    int rc = ensure_lj(h);
    if (rc) return rc;
    return launch(h, kargs(h, OP_FORCE));
}
extern "C" int maddy_integrate(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    cudaError_t e = launch_integrate_kernel(kargs(h, 0), h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "integrate kernel launch: %s", cudaGetErrorString(e));
    h->launches++;
    return MADDY_OK;
}

extern "C" int maddy_tea_update(maddy_handle *h, long long step);
extern "C" int maddy_tea_integrate(maddy_handle *h);
extern "C" int maddy_run(maddy_handle *h, long long first_step, long long n_steps, unsigned flags)
{
    if (!h || n_steps < 0) return MADDY_EINVAL;
    if (n_steps == 0) return MADDY_OK;
    const long long freq = h->p.ljpairsupdatefreq > 0 ? h->p.ljpairsupdatefreq : 1;
    const bool rebuilds = h->p.lj_on || h->p.is_assembly;
    const bool skip_first = (flags & MADDY_RUN_SKIP_FIRST_REBUILD) != 0;
    if (h->p.tea_on) {
        // TEA window: the step is a chain of GPU-wide launches (force -> [epsilon/beta] -> prepare -> pair kernel), queued
        // here back to back; the only host round trip is the capricious check every tea_epsilon_freq steps.
        for (long long step = first_step; step < first_step + n_steps; step++) {
            int rc = MADDY_OK;
            if (rebuilds && step % freq == 0 && !(step == first_step && skip_first)) {
                h->lj_maybe_stale = false;
                rc = launch(h, kargs(h, (h->p.lj_on ? OP_REBUILD_LJ : 0u) | (h->p.is_assembly ? OP_REBUILD_BONDS : 0u)));
            }
            if (!rc) rc = ensure_lj(h);
            if (!rc) rc = launch(h, kargs(h, OP_FORCE | OP_TEA_PREP)); // force + integrateTea_prepare in one launch
            if (!rc) rc = maddy_tea_update(h, step);
            if (!rc) {
                cudaError_t e = launch_tea_kernels(kargs(h, 0), 2, 0, h->stream);
                if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "TEA pair kernel: %s", cudaGetErrorString(e));
                h->launches++;
            }
            if (rc) return rc;
        }
        return MADDY_OK;
    }
    if (!(rebuilds && first_step % freq == 0 && !skip_first)) { // the window starts from the lists as they stand
        int rc = ensure_lj(h);
        if (rc) return rc;
    }
    KArgs k = kargs(h, OP_RUN);
    k.first_step = first_step;
    k.n_steps = n_steps;
    k.run_flags = flags;
    if (h->plan_wait && k.sched_slots > 0 && first_step + n_steps > k.sched_first) { // the window reaches an event planned on the aux stream
        CU(h, cudaStreamWaitEvent(h->stream, h->hyd_staged, 0));
        h->plan_wait = false;
    }
    if (k.lazy && rebuilds) { // does the window contain a list-update step?
        long long m = (first_step + freq - 1) / freq * freq;
        if (m == first_step && skip_first) m += freq;
        if (m < first_step + n_steps) h->lj_maybe_stale = true;
    }
    int rc = launch(h, k);
    if (!rc && k.sched_slots > 0) { // a later schedule upload into this buffer has to wait for this run
        const int b = h->sched_cur;
        if (!h->sched_reader[b]) CU(h, cudaEventCreateWithFlags(&h->sched_reader[b], cudaEventDisableTiming));
        CU(h, cudaEventRecord(h->sched_reader[b], h->stream));
    }
    return rc;
}

static int energies_impl(maddy_handle *h, unsigned ops, double *out_per_traj, double *out_per_monomer);
extern "C" int maddy_energies(maddy_handle *h, double *out_per_traj, double *out_per_monomer)
{
    if (!h) return MADDY_EINVAL;
    return energies_impl(h, OP_ENERGY, out_per_traj, out_per_monomer);
}
extern "C" int maddy_rebuild_and_energies(maddy_handle *h, double *out_per_traj, double *out_per_monomer)
{
    if (!h) return MADDY_EINVAL;
    const unsigned ops = OP_ENERGY | (h->p.lj_on ? OP_REBUILD_LJ : 0u) | (h->p.is_assembly ? OP_REBUILD_BONDS : 0u);
    return energies_impl(h, ops, out_per_traj, out_per_monomer);
}
static int energies_impl(maddy_handle *h, unsigned ops, double *out_per_traj, double *out_per_monomer)
{
    int rc = (ops & OP_REBUILD_LJ) ? MADDY_OK : ensure_lj(h);
    if (rc) return rc;
    if (ops & OP_REBUILD_LJ) h->lj_maybe_stale = false;
    rc = launch(h, kargs(h, ops));
    if (rc) return rc;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (out_per_traj)
        CU(h, cudaMemcpyAsync(out_per_traj, h->a.en_traj, (size_t)h->a.ntr * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (out_per_monomer)
        CU(h, cudaMemcpyAsync(out_per_monomer, h->a.en_mono, n * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    if (out_per_traj || out_per_monomer) CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}
extern "C" void *maddy_energies_device(maddy_handle *h) { return h ? (void *)h->a.en_traj : nullptr; }

// ------------------------------------------------------------------ state transfer
static void soa_to_aos_raw(const float4 *pos, const float4 *ang, size_t n, float *aos);
extern "C" int maddy_download_coords(maddy_handle *h, float *aos)
{
    if (!h || !aos) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (!h->snap_what) { // through the pinned mirrors (a pageable destination halves the copy rate)
        if (!h->snap_pos) {
            CU(h, cudaMallocHost(&h->snap_pos, n * sizeof(float4)));
            CU(h, cudaMallocHost(&h->snap_ang, n * sizeof(float4)));
        }
        CU(h, cudaMemcpyAsync(h->snap_pos, h->a.pos, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(h->snap_ang, h->a.ang, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
        int rc = sync_and_check(h);
        if (rc) return rc;
        soa_to_aos_raw(h->snap_pos, h->snap_ang, n, aos);
        return MADDY_OK;
    }
    std::vector<float4> pos(n), ang(n);
    CU(h, cudaMemcpyAsync(pos.data(), h->a.pos, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(ang.data(), h->a.ang, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    int rc = sync_and_check(h);
    if (rc) return rc;
    soa_to_aos(pos, ang, aos);
    return MADDY_OK;
}
extern "C" int maddy_download_forces(maddy_handle *h, float *aos)
{
    if (!h || !aos) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    std::vector<float4> pos(n), ang(n);
    CU(h, cudaMemcpyAsync(pos.data(), h->a.fpos, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(ang.data(), h->a.fang, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    int rc = sync_and_check(h);
    if (rc) return rc;
    soa_to_aos(pos, ang, aos);
    return MADDY_OK;
}
// ------------------------------------------------------------------ exact on-tubule rule
// `cosf(theta) > cosf(ANG_THRES)` (updater.cpp:161) as intervals of a = |theta| (cosf is even).  The authority is the C
// library's cosf of THIS process (the one the host's mt_length() calls), so the crossing on every monotone branch of the
// cosine is bisected over the float bit patterns with that function, and the neighbourhood of each crossing is then
// checked value by value: the predicate must switch exactly once there, otherwise the rule is refused (the device never
// guesses) and MADDY_SNAP_ONTUBULE reports MADDY_EINVAL.
static bool make_ontub_rule(OnTubRule &r)
{
    const float ang_thres = 1.0f, r_mt = 8.12f, r_thres = 2.0f * 8; // mt.h:23-30
    volatile float one = ang_thres;
    const float thr = cosf(ang_thres);
    if (thr != cosf(one)) return false; // compile-time and run-time evaluation of the threshold must agree
    r.rad_hi = r_mt + r_thres;
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto flt = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    auto pred = [&](float a) { return cosf(a) > thr; };
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < ONTUB_EDGES; k++) {
        // branch k: [k pi, (k+1) pi]; cos decreasing for even k (on -> off), increasing for odd k (off -> on)
        uint32_t lo = bits((float)(k * pi)), hi = bits((float)((k + 1) * pi));
        const bool on_at_lo = (k % 2 == 0);
        if (pred(flt(lo)) != on_at_lo || pred(flt(hi)) == on_at_lo) return false;
        while (hi - lo > 1) {
            const uint32_t mid = lo + (hi - lo) / 2;
            if (pred(flt(mid)) == on_at_lo) lo = mid;
            else hi = mid;
        }
        // exactly one switch within +-4096 ulps of the crossing
        for (uint32_t u = lo - 4096; u <= lo; u++)
            if (pred(flt(u)) != on_at_lo) return false;
        for (uint32_t u = hi; u <= hi + 4096; u++)
            if (pred(flt(u)) == on_at_lo) return false;
        // even k: on iff a < edge (edge = first "off" value); odd k: on iff a > edge (edge = last "off" value)
        r.edge[k] = on_at_lo ? flt(hi) : flt(lo);
    }
    r.a_max = (float)(ONTUB_EDGES * pi); // the end of the last branch that was bisected
    return true;
}
// the rule of this process, made once (the MADDY_ONTUB_AMAX test hook narrows a handle's copy, never this one)
static bool process_ontub_rule(OnTubRule &r)
{
    static std::once_flag once;
    static OnTubRule rule;
    static bool rule_ok = false;
    std::call_once(once, [] { rule_ok = make_ontub_rule(rule); });
    r = rule;
    return rule_ok;
}
static_assert(ONTUB_EDGES == MADDY_ON_TUBULE_EDGES, "maddy_b200.h and OnTubRule disagree");
extern "C" int maddy_on_tubule_rule(float *rad_hi, float *a_max, float *edges)
{
    OnTubRule r;
    if (!process_ontub_rule(r)) return MADDY_EINVAL;
    if (rad_hi) *rad_hi = r.rad_hi;
    if (a_max) *a_max = r.a_max;
    if (edges) memcpy(edges, r.edge, sizeof r.edge);
    return MADDY_OK;
}

// the two on-tubule classification buffers; both start as the flags the handle was created with (called by maddy_create)
static int cls_pair_ready(maddy_handle *h)
{
    if (h->d_cls_pair[0]) return MADDY_OK;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    for (int b = 0; b < 2; b++) {
        CU(h, cudaMalloc(&h->d_cls_pair[b], n));
        CU(h, cudaMemcpyAsync(h->d_cls_pair[b], h->a.ontub, n, cudaMemcpyDeviceToDevice, h->stream));
    }
    return MADDY_OK;
}

// ------------------------------------------------------------------ hydrolysis events of a stride on the device
static const LfibPoly *lfib_host_table()
{
    static LfibPoly table[LFIB_POW2];
    static std::once_flag once;
    std::call_once(once, [] { lfib_table(table); });
    return table;
}
extern "C" void maddy_rand_discard(unsigned *window31, unsigned long long n)
{
    if (!window31 || n == 0) return;
    uint32_t out[LFIB_DEG];
    lfib_window(out, window31, lfib_host_table(), n);
    memcpy(window31, out, sizeof out);
}
static int sched_reserve(maddy_handle *h, size_t bytes);

// buffers every hydrolysis call needs (first use)
static int hyd_ready(maddy_handle *h)
{
    if (h->d_hyd_own) return MADDY_OK;
    const int nd = h->a.N / 2;
    const size_t cells = (size_t)nd * h->a.ntr;
    CU(h, cudaMalloc(&h->d_hyd_own, 2 * cells));
    CU(h, cudaMalloc(&h->d_hyd_window, LFIB_DEG * sizeof(uint32_t)));
    CU(h, cudaMallocHost(&h->h_hyd_window, LFIB_DEG * sizeof(uint32_t)));
    CU(h, cudaMalloc(&h->d_lfib_table, sizeof(LfibPoly) * LFIB_POW2));
    CU(h, cudaMemcpyAsync(h->d_lfib_table, lfib_host_table(), sizeof(LfibPoly) * LFIB_POW2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMalloc(&h->d_hyd_mats, hyd_matrices_bytes()));
    {
        cudaError_t e = launch_hyd_matrices(h->d_lfib_table, h->d_hyd_mats, h->stream); // once per handle
        if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "hydrolysis jump matrices: %s", cudaGetErrorString(e));
    }
    CU(h, cudaMalloc(&h->d_hyd_status, sizeof(int)));
    CU(h, cudaMallocHost(&h->h_hyd_status, sizeof(int)));
    CU(h, cudaEventCreateWithFlags(&h->hyd_staged, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->hyd_done, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&h->hyd_in, cudaEventDisableTiming));
    CU(h, cudaStreamCreateWithFlags(&h->aux_stream, cudaStreamNonBlocking));
    if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    return MADDY_OK;
}
static HydArgs hyd_args(maddy_handle *h, int shards, int shard)
{
    HydArgs a;
    memset(&a, 0, sizeof a);
    a.gtp = h->a.gtp;
    a.extra = h->a.extra;
    a.cur = h->d_cls_pair[h->cls_cur];
    a.prev = h->d_cls_pair[h->cls_cur ^ 1];
    a.own = h->d_hyd_own;
    a.all = shards > 1 ? h->d_hyd_all : h->d_hyd_own;
    a.rowcount = h->d_hyd_rowcount;
    a.rowstart = h->d_hyd_rowstart;
    a.cursor = h->d_hyd_counters;
    a.event_start = h->d_hyd_counters + 1;
    a.stream = h->d_hyd_stream;
    a.status = h->d_hyd_status;
    a.guard = h->a.guard;
    a.N = h->a.N;
    a.nd = h->a.N / 2;
    a.ntr_l = h->a.ntr;
    a.shards = shards;
    a.shard = shard;
    a.ntr = shards * h->a.ntr;
    a.seg = a.ntr < 256 ? a.ntr : 256; // a warp takes 256 trajectories of a dimer row (8 rounds), whatever the ensemble's size
    a.nseg = (a.ntr + a.seg - 1) / a.seg;
    a.nrows = a.nd * a.nseg;
    return a;
}

extern "C" int maddy_hydrolysis_inputs(maddy_handle *h, void **device_buffer, unsigned long *bytes)
{
    if (!h || !device_buffer) return MADDY_EINVAL;
    if (h->a.N & 1) return fail(h, MADDY_EINVAL, "hydrolysis on the device needs an even n_tot");
    CU(h, cudaSetDevice(h->p.device));
    int rc = cls_pair_ready(h);
    if (!rc) rc = hyd_ready(h);
    if (rc) return rc;
    cudaError_t e = launch_hyd_prepare(hyd_args(h, 1, 0), h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "hydrolysis prepare kernel: %s", cudaGetErrorString(e));
    h->launches++;
    *device_buffer = h->d_hyd_own;
    if (bytes) *bytes = (unsigned long)((size_t)h->a.N * h->a.ntr); // 2 x (N / 2) x n_tr_local
    return MADDY_OK;
}

// gathered == nullptr: one shard, the plan prepares its inputs itself
static int hyd_plan_impl(maddy_handle *h, const void *gathered, int shards, const unsigned *window31, long long first_event, long long period, int n_events,
                         unsigned flags)
{
    if (!h || !window31 || n_events < 1 || period <= 0 || shards < 1) return MADDY_EINVAL;
    if (h->a.N & 1) return fail(h, MADDY_EINVAL, "hydrolysis on the device needs an even n_tot");
    if ((long long)shards * h->p.n_tr_local != h->p.n_tr || h->p.traj_first % h->p.n_tr_local != 0)
        return fail(h, MADDY_EINVAL, "hydrolysis plan: %d shard(s) of %d trajectories do not make the ensemble of %d (draw positions are global: "
                                     "every shard must take part, in equal contiguous blocks)", shards, h->p.n_tr_local, h->p.n_tr);
    if (h->hyd_pending) return fail(h, MADDY_EINVAL, "maddy_hydrolysis_plan: the previous plan's result has not been collected");
    const int shard = h->p.traj_first / h->p.n_tr_local;
    CU(h, cudaSetDevice(h->p.device));
    const int N = h->a.N, ntr = h->a.ntr, nd = N / 2;
    const size_t n = (size_t)ntr * N, cells_l = (size_t)nd * ntr, cells = cells_l * shards;
    int rc = cls_pair_ready(h);
    if (!rc) rc = hyd_ready(h);
    if (rc) return rc;
    if (shards > 1 && 2 * cells > h->hyd_all_cap) {
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaStreamSynchronize(h->aux_stream));
        if (h->d_hyd_all && h->d_hyd_all != h->d_hyd_own) cudaFree(h->d_hyd_all);
        h->d_hyd_all = nullptr;
        h->hyd_all_cap = 0;
        CU(h, cudaMalloc(&h->d_hyd_all, 2 * cells));
        h->hyd_all_cap = 2 * cells;
    }
    if (n_events + 1 > h->hyd_counters_cap) {
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaStreamSynchronize(h->aux_stream));
        if (h->d_hyd_counters) cudaFree(h->d_hyd_counters);
        if (h->h_hyd_counters) cudaFreeHost(h->h_hyd_counters);
        h->hyd_counters_cap = n_events + 16;
        CU(h, cudaMalloc(&h->d_hyd_counters, (size_t)h->hyd_counters_cap * sizeof(unsigned long long)));
        CU(h, cudaMallocHost(&h->h_hyd_counters, (size_t)h->hyd_counters_cap * sizeof(unsigned long long)));
    }
    {
        const HydArgs probe = hyd_args(h, shards, shard);
        if (probe.nrows > h->hyd_rows_cap) {
            CU(h, cudaStreamSynchronize(h->stream));
            CU(h, cudaStreamSynchronize(h->aux_stream));
            if (h->d_hyd_rowcount) cudaFree(h->d_hyd_rowcount);
            if (h->d_hyd_rowstart) cudaFree(h->d_hyd_rowstart);
            h->d_hyd_rowcount = nullptr;
            h->d_hyd_rowstart = nullptr;
            h->hyd_rows_cap = 0;
            CU(h, cudaMalloc(&h->d_hyd_rowcount, (size_t)probe.nrows * sizeof(unsigned)));
            CU(h, cudaMalloc(&h->d_hyd_rowstart, (size_t)probe.nrows * sizeof(unsigned long long)));
            h->hyd_rows_cap = probe.nrows;
        }
    }
    // worst case: every dimer of every trajectory of the ENSEMBLE draws at every event
    const unsigned long long need = (unsigned long long)n_events * cells;
    if (need > hyd_stream_capacity()) return fail(h, MADDY_EINVAL, "hydrolysis plan: %llu draws exceed the stream kernel's range", need);
    if (need > h->hyd_stream_cap) {
        CU(h, cudaStreamSynchronize(h->stream));
        CU(h, cudaStreamSynchronize(h->aux_stream));
        if (h->d_hyd_stream) cudaFree(h->d_hyd_stream);
        h->d_hyd_stream = nullptr;
        h->hyd_stream_cap = 0;
        cudaError_t e = cudaMalloc(&h->d_hyd_stream, need * sizeof(uint32_t));
        if (e != cudaSuccess) return fail(h, MADDY_ENOMEM, "%llu bytes for the hydrolysis draws: %s", need * 4ull, cudaGetErrorString(e));
        h->hyd_stream_cap = need;
    }
    rc = sched_reserve(h, n * (size_t)n_events);
    if (rc) return rc;
    const int b = h->sched_cur ^ 1; // the buffer the last scheduled run was NOT given (same stream: no reader left behind)
    if (h->sched_copied[b]) CU(h, cudaStreamWaitEvent(h->aux_stream, h->sched_copied[b], 0));
    // threshold of updater.cpp:236-237 in this process's own double arithmetic
    static const unsigned threshold = [] {
        int v = (int)(0.02 * (double)RAND_MAX) + 2;
        while (!((double)v / (double)RAND_MAX < 0.02)) v--;
        return (unsigned)v;
    }();
    // The kernels run on a stream of their own, behind everything queued so far (classification, GTP state, the gather) and
    // beside what is queued next - the window up to the first event, which needs none of it (the small CTAs fit on the SMs
    // the fused loop leaves half empty).  The first maddy_run that reaches an event of the plan waits for it.
    cudaStream_t ax = h->aux_stream;
    CU(h, cudaEventRecord(h->hyd_in, h->stream));
    CU(h, cudaStreamWaitEvent(ax, h->hyd_in, 0));
    memcpy(h->h_hyd_window, window31, LFIB_DEG * sizeof(uint32_t));
    CU(h, cudaMemcpyAsync(h->d_hyd_window, h->h_hyd_window, LFIB_DEG * sizeof(uint32_t), cudaMemcpyHostToDevice, ax));
    CU(h, cudaMemsetAsync(h->d_hyd_counters, 0, sizeof(unsigned long long), ax));
    CU(h, cudaMemsetAsync(h->d_hyd_status, 0, sizeof(int), ax));
    uint8_t *work = shards > 1 ? h->d_hyd_all : h->d_hyd_own;
    if (gathered && gathered != work) CU(h, cudaMemcpyAsync(work, gathered, 2 * cells, cudaMemcpyDeviceToDevice, ax)); // the plan rewrites its copy
    cudaError_t e = launch_hyd_stream(h->d_hyd_window, h->d_hyd_mats, need, h->d_hyd_stream, h->a.guard, ax);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "hydrolysis stream kernel: %s", cudaGetErrorString(e));
    HydArgs a = hyd_args(h, shards, shard);
    a.all = work;
    a.stream_count = need;
    a.threshold = threshold;
    e = launch_hyd_plan(a, n_events, h->d_sched_buf[b], gathered != nullptr, ax);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "hydrolysis plan kernels: %s", cudaGetErrorString(e));
    h->launches += 2 + 3LL * n_events; // stream + prepare + (count, scan, apply) per event
    h->sched_cur = b;
    h->d_sched = h->d_sched_buf[b];
    h->sched_first = first_event;
    h->sched_period = period;
    h->sched_slots = n_events;
    // counters (and, on request, the slots) travel beside the windows queued next
    CU(h, cudaEventRecord(h->hyd_staged, ax));
    h->plan_wait = true;
    CU(h, cudaStreamWaitEvent(h->copy_stream, h->hyd_staged, 0));
    CU(h, cudaMemcpyAsync(h->h_hyd_counters, h->d_hyd_counters, (size_t)(n_events + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaMemcpyAsync(h->h_hyd_status, h->d_hyd_status, sizeof(int), cudaMemcpyDeviceToHost, h->copy_stream));
    h->hyd_keep = (flags & MADDY_HYD_KEEP_SLOTS) != 0;
    if (h->hyd_keep) {
        const size_t bytes = n * (size_t)n_events;
        if (bytes > h->hyd_slots_cap) {
            if (h->h_hyd_slots) cudaFreeHost(h->h_hyd_slots);
            h->h_hyd_slots = nullptr;
            h->hyd_slots_cap = 0;
            CU(h, cudaMallocHost(&h->h_hyd_slots, bytes));
            h->hyd_slots_cap = bytes;
        }
        CU(h, cudaMemcpyAsync(h->h_hyd_slots, h->d_sched_buf[b], bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    }
    CU(h, cudaEventRecord(h->hyd_done, h->copy_stream));
    // a later plan must not rewrite this buffer while the copy stream still reads it
    if (!h->sched_copied[b]) CU(h, cudaEventCreateWithFlags(&h->sched_copied[b], cudaEventDisableTiming));
    CU(h, cudaEventRecord(h->sched_copied[b], h->copy_stream));
    h->hyd_events = n_events;
    h->hyd_pending = true;
    return MADDY_OK;
}

extern "C" int maddy_hydrolysis_plan(maddy_handle *h, const unsigned *window31, long long first_event, long long period, int n_events, unsigned flags)
{
    if (h && h->p.n_tr_local != h->p.n_tr)
        return fail(h, MADDY_EINVAL, "maddy_hydrolysis_plan needs the whole ensemble on one handle (a shard: maddy_hydrolysis_plan_sharded / _all)");
    return hyd_plan_impl(h, nullptr, 1, window31, first_event, period, n_events, flags);
}
extern "C" int maddy_hydrolysis_plan_sharded(maddy_handle *h, const void *gathered_device, int n_shards, const unsigned *window31, long long first_event,
                                             long long period, int n_events, unsigned flags)
{
    if (!gathered_device) return MADDY_EINVAL;
    return hyd_plan_impl(h, gathered_device, n_shards, window31, first_event, period, n_events, flags);
}

extern "C" int maddy_hydrolysis_result(maddy_handle *h, unsigned long long *draws_total, unsigned long long *event_first_draw, int *gtp_slots)
{
    if (!h) return MADDY_EINVAL;
    if (!h->hyd_pending) return fail(h, MADDY_EINVAL, "maddy_hydrolysis_result without maddy_hydrolysis_plan");
    CU(h, cudaSetDevice(h->p.device));
    CU(h, cudaEventSynchronize(h->hyd_done));
    h->hyd_pending = false;
    if (*h->h_hyd_status) return fail(h, MADDY_EOVERFLOW, "hydrolysis plan: the pre-drawn rand() stream was too short");
    if (draws_total) *draws_total = h->h_hyd_counters[0];
    if (event_first_draw) memcpy(event_first_draw, h->h_hyd_counters + 1, (size_t)h->hyd_events * sizeof(unsigned long long));
    if (gtp_slots) {
        if (!h->hyd_keep) return fail(h, MADDY_EINVAL, "maddy_hydrolysis_result: the plan was made without MADDY_HYD_KEEP_SLOTS");
        const size_t bytes = (size_t)h->a.ntr * h->a.N * (size_t)h->hyd_events;
        const uint8_t *v = h->h_hyd_slots;
        HOST_PARALLEL_FOR(bytes)
        for (long long q = 0; q < (long long)bytes; q++) gtp_slots[q] = v[q];
    }
    return MADDY_OK;
}

extern "C" int maddy_apply_scheduled_gtp(maddy_handle *h, long long step)
{
    if (!h) return MADDY_EINVAL;
    if (h->sched_slots <= 0 || step < h->sched_first || (step - h->sched_first) % h->sched_period != 0) return MADDY_OK;
    const long long slot = (step - h->sched_first) / h->sched_period;
    if (slot >= h->sched_slots) return MADDY_OK;
    CU(h, cudaSetDevice(h->p.device));
    if (h->plan_wait) {
        CU(h, cudaStreamWaitEvent(h->stream, h->hyd_staged, 0));
        h->plan_wait = false;
    }
    const size_t n = (size_t)h->a.ntr * h->a.N;
    CU(h, cudaMemcpyAsync(h->a.gtp, h->d_sched + (size_t)slot * n, n, cudaMemcpyDeviceToDevice, h->stream));
    return MADDY_OK;
}

extern "C" int maddy_clear_guard(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    CU(h, cudaMemsetAsync(h->a.guard, 0, sizeof(int), h->stream));
    return MADDY_OK;
}

extern "C" int maddy_snapshot_gtp(maddy_handle *h, int *gtp)
{
    if (!h || !gtp) return MADDY_EINVAL;
    if (!(h->cls_what & MADDY_SNAP_GTP)) return fail(h, MADDY_EINVAL, "maddy_snapshot_gtp: the last collected snapshot did not carry the GTP flags");
    const size_t n = (size_t)h->a.ntr * h->a.N;
    const uint8_t *v = h->h_snap_gtp;
    HOST_PARALLEL_FOR(n)
    for (size_t q = 0; q < n; q++) gtp[q] = v[q];
    return MADDY_OK;
}

// ------------------------------------------------------------------ asynchronous stride snapshot
extern "C" int maddy_snapshot_begin(maddy_handle *h, unsigned what)
{
    if (!h || !(what & (MADDY_SNAP_COORDS | MADDY_SNAP_FORCES | MADDY_SNAP_ENERGIES | MADDY_SNAP_ONTUBULE | MADDY_SNAP_GTP))) return MADDY_EINVAL;
    if (h->snap_what) return fail(h, MADDY_EINVAL, "maddy_snapshot_begin: the previous snapshot has not been collected");
    if ((what & MADDY_SNAP_ONTUBULE_APPLY) && !(what & MADDY_SNAP_ONTUBULE))
        return fail(h, MADDY_EINVAL, "maddy_snapshot_begin: MADDY_SNAP_ONTUBULE_APPLY needs MADDY_SNAP_ONTUBULE");
    if ((what & MADDY_SNAP_ONTUBULE) && !h->ontub_rule_ok)
        return fail(h, MADDY_EINVAL, "maddy_snapshot_begin: the C library's cosf is not monotone around cosf(ANG_THRES); classify on the host");
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (!h->snap_done) CU(h, cudaEventCreateWithFlags(&h->snap_done, cudaEventDisableTiming));
    if (what & MADDY_SNAP_ENERGIES) {
        unsigned ops = OP_ENERGY;
        if (what & MADDY_SNAP_REBUILD) ops |= (h->p.lj_on ? OP_REBUILD_LJ : 0u) | (h->p.is_assembly ? OP_REBUILD_BONDS : 0u);
        int rc = (ops & OP_REBUILD_LJ) ? MADDY_OK : ensure_lj(h);
        if (rc) return rc;
        if (ops & OP_REBUILD_LJ) h->lj_maybe_stale = false;
        rc = launch(h, kargs(h, ops));
        if (rc) return rc;
        if (!h->snap_en) {
            CU(h, cudaMallocHost(&h->snap_en, (size_t)h->a.ntr * 7 * sizeof(double)));
            CU(h, cudaMalloc(&h->d_snap_en, (size_t)h->a.ntr * 7 * sizeof(double)));
        }
        CU(h, cudaMemcpyAsync(h->d_snap_en, h->a.en_traj, (size_t)h->a.ntr * 7 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
    } else if (what & MADDY_SNAP_REBUILD) {
        return fail(h, MADDY_EINVAL, "maddy_snapshot_begin: MADDY_SNAP_REBUILD needs MADDY_SNAP_ENERGIES");
    }
    cudaEvent_t c0 = nullptr, c1 = nullptr, c2 = nullptr;
    if (h->gpu_prof) {
        cudaEventCreate(&c0);
        cudaEventCreate(&c1);
        cudaEventCreate(&c2);
        cudaEventRecord(c0, h->stream);
    }
    // state -> AoS-7 staging on the device (a few microseconds on the main stream), then the PCIe leg on a second
    // stream so that the window queued next does not wait for it
    const size_t aos_bytes = n * MADDY_COORD_STRIDE * sizeof(float);
    // each resource under its own check: maddy_schedule_gtp may have created the copy stream already
    if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (!h->snap_staged) CU(h, cudaEventCreateWithFlags(&h->snap_staged, cudaEventDisableTiming));
    if (!h->d_snap_status) CU(h, cudaMalloc(&h->d_snap_status, sizeof(int)));
    if (what & MADDY_SNAP_ONTUBULE) {
        // mt_length()'s classification of the state being snapshotted, after the energies above (which read the flags of
        // the previous stride, as in the reference: compute_cuda.cu:1165 precedes :1186-1190).  The counts travel first, on
        // their own event, so a caller that needs them before it may queue the next window waits microseconds, not for the
        // coordinates.
        if (!h->d_cls_count) {
            CU(h, cudaMalloc(&h->d_cls_count, ((size_t)h->a.ntr + 1) * sizeof(int)));
            CU(h, cudaMallocHost(&h->h_cls_flags, n));
            CU(h, cudaMallocHost(&h->h_cls_count, ((size_t)h->a.ntr + 1) * sizeof(int)));
            CU(h, cudaEventCreateWithFlags(&h->cls_staged, cudaEventDisableTiming));
            CU(h, cudaEventCreateWithFlags(&h->cls_done, cudaEventDisableTiming));
        }
        CU(h, cudaMemsetAsync(h->d_cls_count + h->a.ntr, 0, sizeof(int), h->stream));
        int rcp = cls_pair_ready(h);
        if (rcp) return rcp;
        h->cls_cur ^= 1; // the older of the two classifications is overwritten and becomes the current one
        cudaError_t e = launch_ontubule(h->a.pos, h->a.ang, h->a.ntr, h->a.N, h->ontub_rule, h->d_cls_pair[h->cls_cur], h->a.ontub,
                                        (what & MADDY_SNAP_ONTUBULE_APPLY) ? 1 : 0, h->d_cls_count, h->d_cls_count + h->a.ntr,
                                        (what & MADDY_SNAP_ONTUBULE_GUARD) ? h->a.guard : nullptr, h->stream);
        if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "ontubule kernel launch: %s", cudaGetErrorString(e));
        h->launches++;
        CU(h, cudaEventRecord(h->cls_staged, h->stream));
        CU(h, cudaStreamWaitEvent(h->copy_stream, h->cls_staged, 0));
        CU(h, cudaMemcpyAsync(h->h_cls_count, h->d_cls_count, ((size_t)h->a.ntr + 1) * sizeof(int), cudaMemcpyDeviceToHost, h->copy_stream));
        CU(h, cudaEventRecord(h->cls_done, h->copy_stream));
    }
    if (what & MADDY_SNAP_COORDS) {
        if (!h->snap_r) {
            CU(h, cudaMallocHost(&h->snap_r, aos_bytes));
            CU(h, cudaMalloc(&h->d_snap_r, aos_bytes));
        }
        cudaError_t e = launch_snapshot_kernel(h->a.pos, h->a.ang, h->d_snap_r, n, h->stream);
        if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "snapshot kernel launch: %s", cudaGetErrorString(e));
        h->launches++;
    }
    if (what & MADDY_SNAP_FORCES) {
        if (!h->snap_f) {
            CU(h, cudaMallocHost(&h->snap_f, aos_bytes));
            CU(h, cudaMalloc(&h->d_snap_f, aos_bytes));
        }
        cudaError_t e = launch_snapshot_kernel(h->a.fpos, h->a.fang, h->d_snap_f, n, h->stream);
        if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "snapshot kernel launch: %s", cudaGetErrorString(e));
        h->launches++;
    }
    if (what & MADDY_SNAP_GTP) {
        if (!h->d_snap_gtp) {
            CU(h, cudaMalloc(&h->d_snap_gtp, n));
            CU(h, cudaMallocHost(&h->h_snap_gtp, n));
        }
        CU(h, cudaMemcpyAsync(h->d_snap_gtp, h->a.gtp, n, cudaMemcpyDeviceToDevice, h->stream)); // the next window may rewrite a.gtp
    }
    if (h->gpu_prof) cudaEventRecord(c1, h->stream);
    CU(h, cudaMemcpyAsync(h->d_snap_status, h->a.status, sizeof(int), cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaEventRecord(h->snap_staged, h->stream));
    CU(h, cudaStreamWaitEvent(h->copy_stream, h->snap_staged, 0));
    if (what & MADDY_SNAP_COORDS) CU(h, cudaMemcpyAsync(h->snap_r, h->d_snap_r, aos_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    if (what & MADDY_SNAP_FORCES) CU(h, cudaMemcpyAsync(h->snap_f, h->d_snap_f, aos_bytes, cudaMemcpyDeviceToHost, h->copy_stream));
    if (what & MADDY_SNAP_ENERGIES)
        CU(h, cudaMemcpyAsync(h->snap_en, h->d_snap_en, (size_t)h->a.ntr * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
    if (what & MADDY_SNAP_ONTUBULE) CU(h, cudaMemcpyAsync(h->h_cls_flags, h->d_cls_pair[h->cls_cur], n, cudaMemcpyDeviceToHost, h->copy_stream));
    if (what & MADDY_SNAP_GTP) CU(h, cudaMemcpyAsync(h->h_snap_gtp, h->d_snap_gtp, n, cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaMemcpyAsync(h->h_status + 1, h->d_snap_status, sizeof(int), cudaMemcpyDeviceToHost, h->copy_stream));
    CU(h, cudaEventRecord(h->snap_done, h->copy_stream));
    if (h->gpu_prof) {
        cudaEventRecord(c2, h->stream);
        h->prof_copy.push_back({c0, c1});
        h->prof_copy.push_back({c1, c2});
    }
    h->snap_what = what;
    return MADDY_OK;
}

static void host_copy(void *dst, const void *src, size_t bytes)
{
    const size_t chunk = 1 << 18, nchunk = (bytes + chunk - 1) / chunk;
    HOST_PARALLEL_FOR(bytes)
    for (size_t c = 0; c < nchunk; c++) {
        const size_t o = c * chunk;
        memcpy((char *)dst + o, (const char *)src + o, bytes - o < chunk ? bytes - o : chunk);
    }
}
static void soa_to_aos_raw(const float4 *pos, const float4 *ang, size_t n, float *aos)
{
    HOST_PARALLEL_FOR(n)
    for (size_t q = 0; q < n; q++) {
        float *c = aos + q * MADDY_COORD_STRIDE;
        c[0] = pos[q].x; c[1] = pos[q].y; c[2] = pos[q].z;
        c[3] = ang[q].x; c[4] = ang[q].z; c[5] = ang[q].y;
        c[6] = 0.f;
    }
}

extern "C" int maddy_snapshot_end(maddy_handle *h, float *coords_aos7, float *forces_aos7, double *energies_per_traj)
{
    if (!h) return MADDY_EINVAL;
    if (!h->snap_what) return fail(h, MADDY_EINVAL, "maddy_snapshot_end without maddy_snapshot_begin");
    const unsigned what = h->snap_what;
    h->snap_what = 0;
    h->cls_what = what & (MADDY_SNAP_ONTUBULE | MADDY_SNAP_ONTUBULE_APPLY | MADDY_SNAP_GTP);
    CU(h, cudaSetDevice(h->p.device));
    CU(h, cudaEventSynchronize(h->snap_done));
    // the status word copied with the snapshot covers everything queued before it; later launches may already be
    // running, so the device word is cleared by whoever synchronises the stream next (check_status)
    if (h->h_status[1] & (ST_LJ_OVERFLOW | ST_LONG_OVERFLOW | ST_LAT_OVERFLOW)) {
        const int st = h->h_status[1];
        return fail(h, MADDY_EOVERFLOW, "neighbour list overflow:%s%s%s (capacities LJ %d, longitudinal %d, lateral %d)",
                    (st & ST_LJ_OVERFLOW) ? " LJ" : "", (st & ST_LONG_OVERFLOW) ? " longitudinal" : "",
                    (st & ST_LAT_OVERFLOW) ? " lateral" : "", MADDY_LJ_CAPACITY, h->a.capLong, h->a.capLat);
    }
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if ((what & MADDY_SNAP_COORDS) && coords_aos7) host_copy(coords_aos7, h->snap_r, n * MADDY_COORD_STRIDE * sizeof(float));
    if ((what & MADDY_SNAP_FORCES) && forces_aos7) host_copy(forces_aos7, h->snap_f, n * MADDY_COORD_STRIDE * sizeof(float));
    if ((what & MADDY_SNAP_ENERGIES) && energies_per_traj) memcpy(energies_per_traj, h->snap_en, (size_t)h->a.ntr * 7 * sizeof(double));
    return MADDY_OK;
}

extern "C" int maddy_has_exact_on_tubule(const maddy_handle *h) { return h && h->ontub_rule_ok ? 1 : 0; }

extern "C" int maddy_snapshot_tubule_lengths(maddy_handle *h, int *mt_len, int *undecided)
{
    if (!h) return MADDY_EINVAL;
    if (!((h->snap_what | h->cls_what) & MADDY_SNAP_ONTUBULE))
        return fail(h, MADDY_EINVAL, "maddy_snapshot_tubule_lengths: neither the snapshot in flight nor the one collected last classified");
    CU(h, cudaSetDevice(h->p.device));
    CU(h, cudaEventSynchronize(h->cls_done));
    if (mt_len) memcpy(mt_len, h->h_cls_count, (size_t)h->a.ntr * sizeof(int));
    if (undecided) *undecided = h->h_cls_count[h->a.ntr];
    return MADDY_OK;
}

extern "C" int maddy_snapshot_on_tubule(maddy_handle *h, int *on_tubule_cur, int *mt_len)
{
    if (!h) return MADDY_EINVAL;
    if (!(h->cls_what & MADDY_SNAP_ONTUBULE)) return fail(h, MADDY_EINVAL, "maddy_snapshot_on_tubule: the last collected snapshot did not classify");
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (h->h_cls_count[h->a.ntr])
        return fail(h, MADDY_EINVAL, "on-tubule classification: some |theta| is beyond %g rad, outside the range of the exact rule; classify on the host",
                    (double)h->ontub_rule.a_max);
    if (on_tubule_cur) {
        const uint8_t *v = h->h_cls_flags;
        HOST_PARALLEL_FOR(n)
        for (size_t q = 0; q < n; q++) on_tubule_cur[q] = v[q];
    }
    if (mt_len) memcpy(mt_len, h->h_cls_count, (size_t)h->a.ntr * sizeof(int));
    return MADDY_OK;
}

extern "C" int maddy_upload_coords(maddy_handle *h, const float *aos)
{
    if (!h || !aos) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    std::vector<float4> pos, ang;
    aos_to_soa(aos, n, pos, ang, false);
    int rc = ensure_lj(h); // the Verlet list keeps referring to the old positions until the next list-update step
    if (rc) return rc;
    CU(h, cudaMemsetAsync(h->a.cand_valid, 0, (size_t)h->a.ntr * sizeof(int), h->stream)); // candidate lists refer to the old positions
    CU(h, cudaMemcpyAsync(h->a.pos, pos.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->a.ang, ang.data(), n * sizeof(float4), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->a.rpos, h->a.pos, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}

extern "C" int maddy_insert_dimers(maddy_handle *h, int n_insert, const int *index, const float *xyzz)
{
    if (!h || n_insert < 0 || (n_insert > 0 && (!index || !xyzz))) return MADDY_EINVAL;
    if (n_insert == 0) return MADDY_OK;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    for (int k = 0; k < n_insert; k++)
        if (index[k] < 0 || (size_t)index[k] + 1 >= n || index[k] % h->a.N == h->a.N - 1)
            return fail(h, MADDY_EINVAL, "maddy_insert_dimers: index[%d] = %d is not the first monomer of a dimer of this handle", k, index[k]);
    CU(h, cudaSetDevice(h->p.device));
    int rc = ensure_lj(h); // the Verlet list keeps referring to the old positions until the next list-update step
    if (rc) return rc;
    const size_t idx_bytes = ((size_t)n_insert * sizeof(int) + 15) & ~(size_t)15, bytes = idx_bytes + (size_t)n_insert * sizeof(float4);
    if (bytes > h->ins_capacity) {
        if (h->ins_done) CU(h, cudaEventSynchronize(h->ins_done));
        else CU(h, cudaEventCreateWithFlags(&h->ins_done, cudaEventDisableTiming));
        if (h->d_ins) cudaFree(h->d_ins);
        if (h->h_ins) cudaFreeHost(h->h_ins);
        h->d_ins = h->h_ins = nullptr;
        h->ins_capacity = 0;
        const size_t cap = bytes < 4096 ? 4096 : 2 * bytes;
        cudaError_t e = cudaMalloc(&h->d_ins, cap);
        if (e == cudaSuccess) e = cudaMallocHost(&h->h_ins, cap);
        if (e != cudaSuccess) return fail(h, MADDY_ENOMEM, "%zu bytes for the insertion records: %s", cap, cudaGetErrorString(e));
        h->ins_capacity = cap;
    } else {
        CU(h, cudaEventSynchronize(h->ins_done)); // the previous call's copy has left the pinned buffer
    }
    memcpy(h->h_ins, index, (size_t)n_insert * sizeof(int));
    memcpy(h->h_ins + idx_bytes, xyzz, (size_t)n_insert * sizeof(float4));
    CU(h, cudaMemcpyAsync(h->d_ins, h->h_ins, bytes, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaEventRecord(h->ins_done, h->stream));
    cudaError_t e = launch_insert_dimers(h->a.pos, h->a.rpos, h->a.extra, h->a.cand_valid, h->a.N, n_insert, (const int *)h->d_ins,
                                         (const float4 *)(h->d_ins + idx_bytes), h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "insertion kernel launch: %s", cudaGetErrorString(e));
    h->launches++;
    return MADDY_OK;
}

// next pinned staging buffer (waits only if the copy that last used it has not finished)
static int stage_acquire(maddy_handle *h, uint8_t **buf, int *slot)
{
    CU(h, cudaSetDevice(h->p.device));
    const int k = h->stage_next;
    h->stage_next = (k + 1) % maddy_handle::kStage;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (!h->stage[k]) {
        CU(h, cudaMallocHost(&h->stage[k], n));
        CU(h, cudaEventCreateWithFlags(&h->stage_done[k], cudaEventDisableTiming));
    } else {
        CU(h, cudaEventSynchronize(h->stage_done[k]));
    }
    *buf = h->stage[k];
    *slot = k;
    return MADDY_OK;
}
static int stage_submit(maddy_handle *h, uint8_t *dst, int slot)
{
    const size_t n = (size_t)h->a.ntr * h->a.N;
    CU(h, cudaMemcpyAsync(dst, h->stage[slot], n, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaEventRecord(h->stage_done[slot], h->stream));
    return MADDY_OK;
}
// the two device / pinned schedule buffers hold at least `bytes` (a resize waits for everything in flight)
static int sched_reserve(maddy_handle *h, size_t bytes)
{
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    if (bytes > h->sched_capacity) {
        CU(h, cudaStreamSynchronize(h->stream)); // a running window may still read the old buffers
        CU(h, cudaStreamSynchronize(h->copy_stream));
        if (h->aux_stream) CU(h, cudaStreamSynchronize(h->aux_stream)); // ... and a plan may still be writing them
        const size_t cap = bytes < 4 * n ? 4 * n : bytes; // room for a few events without another resize
        for (int b = 0; b < 2; b++) {
            if (h->d_sched_buf[b]) cudaFree(h->d_sched_buf[b]);
            if (h->h_sched_buf[b]) cudaFreeHost(h->h_sched_buf[b]);
            h->d_sched_buf[b] = h->h_sched_buf[b] = nullptr;
        }
        h->sched_capacity = 0;
        h->d_sched = nullptr;
        for (int b = 0; b < 2; b++) {
            cudaError_t e = cudaMalloc(&h->d_sched_buf[b], cap);
            if (e == cudaSuccess) e = cudaMallocHost(&h->h_sched_buf[b], cap);
            if (e != cudaSuccess) return fail(h, MADDY_ENOMEM, "%zu bytes for the GTP schedule: %s", cap, cudaGetErrorString(e));
            if (!h->sched_copied[b]) CU(h, cudaEventCreateWithFlags(&h->sched_copied[b], cudaEventDisableTiming));
        }
        h->sched_capacity = cap;
    }
    return MADDY_OK;
}
extern "C" int maddy_schedule_gtp(maddy_handle *h, long long first_event, long long period, int n_slots, const int *gtp_slots)
{
    if (!h || n_slots < 0 || (n_slots > 0 && (!gtp_slots || period <= 0))) return MADDY_EINVAL;
    h->sched_slots = 0;
    if (n_slots == 0) return MADDY_OK;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N, bytes = n * (size_t)n_slots;
    if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    {
        int rcr = sched_reserve(h, bytes);
        if (rcr) return rcr;
    }
    // Fill the buffer the last scheduled run was NOT given (that run may still be executing).  The copy waits for the
    // last run that did read this buffer (sched_reader), the next run for the copy (sched_copied).
    const int b = h->sched_cur ^ 1;
    CU(h, cudaEventSynchronize(h->sched_copied[b])); // the pinned buffer is free again (its last copy has completed)
    uint8_t *v = h->h_sched_buf[b];
    HOST_PARALLEL_FOR(bytes)
    for (long long q = 0; q < (long long)bytes; q++) v[q] = (uint8_t)(gtp_slots[q] == 1 ? 1 : (gtp_slots[q] == 0 ? 0 : 2));
    if (h->sched_reader[b]) CU(h, cudaStreamWaitEvent(h->copy_stream, h->sched_reader[b], 0)); // last run that read buffer b
    CU(h, cudaMemcpyAsync(h->d_sched_buf[b], v, bytes, cudaMemcpyHostToDevice, h->copy_stream));
    CU(h, cudaEventRecord(h->sched_copied[b], h->copy_stream));
    CU(h, cudaStreamWaitEvent(h->stream, h->sched_copied[b], 0)); // the next run starts after its schedule has landed
    h->sched_cur = b;
    h->d_sched = h->d_sched_buf[b];
    h->sched_first = first_event;
    h->sched_period = period;
    h->sched_slots = n_slots;
    return MADDY_OK;
}

extern "C" int maddy_upload_gtp(maddy_handle *h, const int *gtp)
{
    if (!h || !gtp) return MADDY_EINVAL;
    h->sched_slots = 0; // an explicit upload supersedes any schedule
    const size_t n = (size_t)h->a.ntr * h->a.N;
    uint8_t *v;
    int slot, rc = stage_acquire(h, &v, &slot);
    if (rc) return rc;
    HOST_PARALLEL_FOR(n)
    for (size_t q = 0; q < n; q++) v[q] = (uint8_t)(gtp[q] == 1 ? 1 : (gtp[q] == 0 ? 0 : 2)); // kernels test == 1
    return stage_submit(h, h->a.gtp, slot);
}
extern "C" int maddy_upload_on_tubule(maddy_handle *h, const int *on)
{
    if (!h || !on) return MADDY_EINVAL;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    uint8_t *v;
    int slot, rc = stage_acquire(h, &v, &slot);
    if (rc) return rc;
    HOST_PARALLEL_FOR(n)
    for (size_t q = 0; q < n; q++) v[q] = on[q] != 0;
    return stage_submit(h, h->a.ontub, slot);
}
extern "C" int maddy_upload_extra(maddy_handle *h, const unsigned char *extra)
{
    if (!h || !extra) return MADDY_EINVAL;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    uint8_t *v;
    int slot, rc = stage_acquire(h, &v, &slot);
    if (rc) return rc;
    for (size_t q = 0; q < n; q++) v[q] = extra[q] != 0;
    rc = ensure_lj(h);
    if (rc) return rc;
    if (h->a.cand_valid) CU(h, cudaMemsetAsync(h->a.cand_valid, 0, (size_t)h->a.ntr * sizeof(int), h->stream)); // rows of former extras are empty
    return stage_submit(h, h->a.extra, slot);
}

// native bond code (j<<1 | neg) <-> reference encodings
static inline int long_to_ref(unsigned code)
{
    int j = (int)(code >> 1);
    return (code & 1u) ? -j : j;
}
// dynamic lists (pairs_kernel) write the ZERO sentinel for +-0; the static host builder
// (preparator.cpp:550-554) stores a plain 0, which compute_kernel then decodes on its `j <= 0` branch
static inline int lat_to_ref(unsigned code, bool dynamic)
{
    int j = (int)(code >> 1);
    if (j == 0 && !dynamic) return 0;
    if (code & 1u) return j ? -j : -MADDY_ZERO_SENTINEL;
    return j ? j : MADDY_ZERO_SENTINEL;
}

extern "C" int maddy_download_list(maddy_handle *h, int kind, int *counts, int *entries)
{
    if (!h || !counts || !entries) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const DevSys &a = h->a;
    const int N = a.N, Npad = a.Npad, ntr = a.ntr;
    if (kind == MADDY_LIST_LJ) {
        if (!h->p.lj_on) return fail(h, MADDY_EINVAL, "LJ list requested but LJ_on is off");
        int rce = ensure_lj(h);
        if (rce) return rce;
        std::vector<uint16_t> lj((size_t)ntr * MADDY_LJ_CAPACITY * Npad), cnt((size_t)ntr * Npad);
        CU(h, cudaMemcpyAsync(lj.data(), a.lj, lj.size() * 2, cudaMemcpyDeviceToHost, h->stream));
        CU(h, cudaMemcpyAsync(cnt.data(), a.ljcnt, cnt.size() * 2, cudaMemcpyDeviceToHost, h->stream));
        int rc = sync_and_check(h);
        if (rc) return rc;
        for (int t = 0; t < ntr; t++)
            for (int i = 0; i < N; i++) {
                const int c = cnt[(size_t)t * Npad + i];
                counts[(size_t)t * N + i] = c;
                int *o = entries + ((size_t)t * N + i) * MADDY_LJ_CAPACITY;
                for (int k = 0; k < MADDY_LJ_CAPACITY; k++)
                    o[k] = k < c ? lj[((size_t)t * MADDY_LJ_CAPACITY + k) * Npad + i] : 0;
            }
        return MADDY_OK;
    }
    if (kind != MADDY_LIST_LONGITUDINAL && kind != MADDY_LIST_LATERAL) return fail(h, MADDY_EINVAL, "unknown list kind %d", kind);
    const int rows = a.capLong + a.capLat;
    std::vector<uint16_t> bl((size_t)ntr * rows * Npad);
    std::vector<uint8_t> bc((size_t)ntr * 2 * Npad);
    CU(h, cudaMemcpyAsync(bl.data(), a.bl, bl.size() * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(bc.data(), a.bcnt, bc.size(), cudaMemcpyDeviceToHost, h->stream));
    int rc = sync_and_check(h);
    if (rc) return rc;
    const bool lat = kind == MADDY_LIST_LATERAL;
    const int cap = lat ? a.capLat : a.capLong, row0 = lat ? a.capLong : 0;
    for (int t = 0; t < ntr; t++)
        for (int i = 0; i < N; i++) {
            const int c = bc[((size_t)t * 2 + (lat ? 1 : 0)) * Npad + i];
            counts[(size_t)t * N + i] = c;
            int *o = entries + ((size_t)t * N + i) * cap;
            for (int k = 0; k < cap; k++) {
                if (k >= c) {
                    o[k] = 0;
                    continue;
                }
                const unsigned code = bl[((size_t)t * rows + row0 + k) * Npad + i];
                o[k] = lat ? lat_to_ref(code, h->p.is_assembly != 0) : long_to_ref(code);
            }
        }
    return MADDY_OK;
}

extern "C" int maddy_upload_list(maddy_handle *h, int kind, const int *counts, const int *entries)
{
    if (!h || !counts || !entries) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const DevSys &a = h->a;
    const int N = a.N, Npad = a.Npad, ntr = a.ntr;
    if (kind == MADDY_LIST_LJ) {
        if (!h->p.lj_on) return fail(h, MADDY_EINVAL, "LJ list upload but LJ_on is off");
        h->lj_maybe_stale = false; // the uploaded list is the list
        CU(h, cudaMemsetAsync(h->a.lj_stale, 0, (size_t)ntr * sizeof(int), h->stream));
        std::vector<uint16_t> lj((size_t)ntr * MADDY_LJ_CAPACITY * Npad, 0), cnt((size_t)ntr * Npad, 0);
        for (int t = 0; t < ntr; t++)
            for (int i = 0; i < N; i++) {
                const int c = counts[(size_t)t * N + i];
                if (c < 0 || c > MADDY_LJ_CAPACITY) return fail(h, MADDY_EINVAL, "LJ count %d out of range", c);
                cnt[(size_t)t * Npad + i] = (uint16_t)c;
                for (int k = 0; k < c; k++) {
                    const int j = entries[((size_t)t * N + i) * MADDY_LJ_CAPACITY + k];
                    if (j < 0 || j >= N) return fail(h, MADDY_EINVAL, "LJ entry %d out of range", j);
                    lj[((size_t)t * MADDY_LJ_CAPACITY + k) * Npad + i] = (uint16_t)j;
                }
            }
        CU(h, cudaMemcpyAsync(a.lj, lj.data(), lj.size() * 2, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaMemcpyAsync(a.ljcnt, cnt.data(), cnt.size() * 2, cudaMemcpyHostToDevice, h->stream));
        CU(h, cudaStreamSynchronize(h->stream));
        return MADDY_OK;
    }
    if (kind != MADDY_LIST_LONGITUDINAL && kind != MADDY_LIST_LATERAL) return fail(h, MADDY_EINVAL, "unknown list kind %d", kind);
    const bool lat = kind == MADDY_LIST_LATERAL;
    const int rows = a.capLong + a.capLat, cap = lat ? a.capLat : a.capLong, row0 = lat ? a.capLong : 0;
    // read-modify-write of the interleaved bond table
    std::vector<uint16_t> bl((size_t)ntr * rows * Npad);
    std::vector<uint8_t> bc((size_t)ntr * 2 * Npad);
    CU(h, cudaMemcpyAsync(bl.data(), a.bl, bl.size() * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(bc.data(), a.bcnt, bc.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    for (int t = 0; t < ntr; t++)
        for (int i = 0; i < N; i++) {
            const int c = counts[(size_t)t * N + i];
            if (c < 0 || c > cap) return fail(h, MADDY_EINVAL, "bond count %d exceeds capacity %d", c, cap);
            bc[((size_t)t * 2 + (lat ? 1 : 0)) * Npad + i] = (uint8_t)c;
            for (int k = 0; k < c; k++) {
                int v = entries[((size_t)t * N + i) * cap + k];
                unsigned neg;
                int j;
                if (lat) { // compute_cuda.cu:307-324: `j <= 0` selects the swapped site set, ZERO stands for 0
                    neg = v <= 0;
                    j = abs(v);
                    if (j == MADDY_ZERO_SENTINEL) j = 0;
                } else { // compute_cuda.cu:191-197
                    neg = v < 0;
                    j = abs(v);
                }
                if (j >= N) return fail(h, MADDY_EINVAL, "bond entry %d out of range", v);
                bl[((size_t)t * rows + row0 + k) * Npad + i] = (uint16_t)((j << 1) | neg);
            }
        }
    CU(h, cudaMemcpyAsync(a.bl, bl.data(), bl.size() * 2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(a.bcnt, bc.data(), bc.size(), cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}

extern "C" int maddy_download_rng(maddy_handle *h, unsigned *state)
{
    if (!h || !state) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    CU(h, cudaMemcpyAsync(state, h->a.rng_xyz, n * 16, cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaMemcpyAsync(state + n * 4, h->a.rng_ang, n * 16, cudaMemcpyDeviceToHost, h->stream));
    return sync_and_check(h);
}
extern "C" int maddy_upload_rng(maddy_handle *h, const unsigned *state)
{
    if (!h || !state) return MADDY_EINVAL;
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    CU(h, cudaMemcpyAsync(h->a.rng_xyz, state, n * 16, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->a.rng_ang, state + n * 4, n * 16, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}

// ------------------------------------------------------------------ TEA (step-granular)
extern "C" int maddy_tea_update(maddy_handle *h, long long step)
{
    if (!h) return MADDY_EINVAL;
    if (!h->p.tea_on) return fail(h, MADDY_EINVAL, "maddy_tea_update: tea_on is off");
    if (step % h->p.tea_epsilon_freq != 0) return MADDY_OK;
    CU(h, cudaSetDevice(h->p.device));
    cudaError_t e = launch_tea_kernels(kargs(h, OP_TEA_EPS), 0, step, h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "TEA epsilon kernel: %s", cudaGetErrorString(e));
    h->launches += 3;
    // the reference aborts the process on capricious violations (bdhitea.cu:89-107): surface them here
    CU(h, cudaMemcpyAsync(h->h_status, h->a.status, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    if (*h->h_status & 0x100) {
        *h->h_status = 0;
        cudaMemsetAsync(h->a.status, 0, sizeof(int), h->stream);
        return fail(h, MADDY_ETEA, "TEA: hydrodynamic tensor outside the capricious bounds at step %lld", step);
    }
    return check_status(h);
}
extern "C" int maddy_download_tea(maddy_handle *h, float *ci4, float *epsilon, float *beta)
{
    if (!h) return MADDY_EINVAL;
    if (!h->p.tea_on) return fail(h, MADDY_EINVAL, "maddy_download_tea: tea_on is off");
    CU(h, cudaSetDevice(h->p.device));
    const size_t n = (size_t)h->a.ntr * h->a.N;
    if (ci4) CU(h, cudaMemcpyAsync(ci4, h->a.tea_ci, n * sizeof(float4), cudaMemcpyDeviceToHost, h->stream));
    if (epsilon) CU(h, cudaMemcpyAsync(epsilon, h->a.tea_eps, n * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (beta) CU(h, cudaMemcpyAsync(beta, h->a.tea_beta, (size_t)h->a.ntr * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}
extern "C" int maddy_tea_integrate(maddy_handle *h)
{
    if (!h) return MADDY_EINVAL;
    if (!h->p.tea_on) return fail(h, MADDY_EINVAL, "maddy_tea_integrate: tea_on is off");
    CU(h, cudaSetDevice(h->p.device));
    cudaError_t e = launch_tea_kernels(kargs(h, 0), 1, 0, h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "TEA integrate kernel: %s", cudaGetErrorString(e));
    h->launches += 2;
    return MADDY_OK;
}

// ------------------------------------------------------------------ in-situ analysis
extern "C" int maddy_analysis_setup(maddy_handle *h, const int *chain, const int *resid, const char *name1, int n_pf)
{
    if (!h || !chain || !resid || !name1) return MADDY_EINVAL;
    if (n_pf < 1 || n_pf > 32) return fail(h, MADDY_EINVAL, "maddy_analysis_setup: n_pf=%d outside [1,32]", n_pf);
    if (h->an.temp) return fail(h, MADDY_EINVAL, "maddy_analysis_setup: already set up");
    CU(h, cudaSetDevice(h->p.device));
    const int N = h->a.N, ntr = h->a.ntr;
    const size_t n = (size_t)ntr * N;
    std::vector<short> c(N), r(N);
    for (int i = 0; i < N; i++) {
        if (resid[i] < -32768 || resid[i] > 32767) return fail(h, MADDY_EINVAL, "maddy_analysis_setup: residue number %d of atom %d", resid[i], i);
        c[i] = (short)((chain[i] >= 0 && chain[i] < n_pf) ? chain[i] : -1);
        r[i] = (short)resid[i];
    }
    AnalysisArgs &a = h->an;
    std::vector<PoolReq> reqs;
    g_pool_reqs = &reqs;
    pool_req(&a.ppos, n);
    pool_req(&a.pang, n);
    pool_req(const_cast<short **>(&a.chain), (size_t)N);
    pool_req(const_cast<short **>(&a.resid), (size_t)N);
    pool_req(const_cast<char **>(&a.name1), (size_t)N);
    pool_req(&a.temp, (size_t)ntr * 8);
    pool_req(&a.proj, n * 3);
    pool_req(&a.pf, (size_t)ntr * n_pf * 3);
    int rc = pool_commit(h, reqs);
    if (rc) {
        a = AnalysisArgs{};
        return rc;
    }
    a.pos = h->a.pos;
    a.ang = h->a.ang;
    a.N = N;
    a.ntr = ntr;
    a.n_pf = n_pf;
    CU(h, cudaMemcpyAsync((void *)a.chain, c.data(), (size_t)N * 2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync((void *)a.resid, r.data(), (size_t)N * 2, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaMemcpyAsync((void *)a.name1, name1, (size_t)N, cudaMemcpyHostToDevice, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}
static int analysis_ready(maddy_handle *h, const char *who)
{
    if (!h) return MADDY_EINVAL;
    if (!h->an.temp) return fail(h, MADDY_EINVAL, "%s: call maddy_analysis_setup first", who);
    CU(h, cudaSetDevice(h->p.device));
    return MADDY_OK;
}
extern "C" int maddy_analysis_reference(maddy_handle *h)
{
    int rc = analysis_ready(h, "maddy_analysis_reference");
    if (rc) return rc;
    const size_t n = (size_t)h->a.ntr * h->a.N;
    CU(h, cudaMemcpyAsync(h->an.ppos, h->a.pos, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    CU(h, cudaMemcpyAsync(h->an.pang, h->a.ang, n * sizeof(float4), cudaMemcpyDeviceToDevice, h->stream));
    h->an_has_reference = true;
    return MADDY_OK;
}
static int analysis_launch(maddy_handle *h, int which, const char *who)
{
    cudaError_t e = launch_analysis(which, h->an, h->stream);
    if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "%s: %s", who, cudaGetErrorString(e));
    h->launches++;
    return MADDY_OK;
}
extern "C" int maddy_analysis_temperature(maddy_handle *h, double *sums)
{
    int rc = analysis_ready(h, "maddy_analysis_temperature");
    if (rc) return rc;
    if (!sums) return MADDY_EINVAL;
    if (!h->an_has_reference) return fail(h, MADDY_EINVAL, "maddy_analysis_temperature: no previous frame (maddy_analysis_reference)");
    rc = analysis_launch(h, 0, "analysis_temperature_kernel");
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(sums, h->an.temp, (size_t)h->a.ntr * 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}
extern "C" int maddy_analysis_project(maddy_handle *h, float *out)
{
    int rc = analysis_ready(h, "maddy_analysis_project");
    if (rc) return rc;
    if (!out) return MADDY_EINVAL;
    rc = analysis_launch(h, 1, "analysis_project_kernel");
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(out, h->an.proj, (size_t)h->a.ntr * h->a.N * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}
extern "C" int maddy_analysis_protofilaments(maddy_handle *h, int *out)
{
    int rc = analysis_ready(h, "maddy_analysis_protofilaments");
    if (rc) return rc;
    if (!out) return MADDY_EINVAL;
    rc = analysis_launch(h, 2, "analysis_protofilament_kernel");
    if (rc) return rc;
    CU(h, cudaMemcpyAsync(out, h->an.pf, (size_t)h->a.ntr * h->an.n_pf * 3 * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CU(h, cudaStreamSynchronize(h->stream));
    return MADDY_OK;
}

// ------------------------------------------------------------------ NCCL ensemble reduction
// NCCL is loaded lazily so that the library has no link-time dependency on a particular libnccl
// (a Python process may already carry torch's copy).
namespace {
typedef struct ncclComm *ncclComm_t;
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load()
    {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) return false;
        CommInitAll = (decltype(CommInitAll))dlsym(lib, "ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
        GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        return CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
    }
};
Nccl g_nccl;
std::vector<ncclComm_t> g_comms;
std::vector<int> g_comm_devs;
} // namespace

extern "C" int maddy_ensemble_allreduce(maddy_handle **hs, int n, double **values, int count)
{
    if (!hs || n <= 0 || !values || count <= 0) return MADDY_EINVAL;
    maddy_handle *h0 = hs[0];
    if (n == 1) return MADDY_OK;
    if (!g_nccl.load()) return fail(h0, MADDY_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
    std::vector<int> devs(n);
    for (int g = 0; g < n; g++) devs[g] = hs[g]->p.device;
    if (g_comm_devs != devs) {
        for (ncclComm_t c : g_comms) g_nccl.CommDestroy(c);
        g_comms.assign(n, nullptr);
        int r = g_nccl.CommInitAll(g_comms.data(), n, devs.data());
        if (r != 0) {
            g_comms.clear();
            g_comm_devs.clear();
            return fail(h0, MADDY_ENCCL, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        }
        g_comm_devs = devs;
    }
    std::vector<double *> dbuf(n, nullptr);
    for (int g = 0; g < n; g++) {
        CU(hs[g], cudaSetDevice(devs[g]));
        CU(hs[g], cudaMalloc(&dbuf[g], count * sizeof(double)));
        CU(hs[g], cudaMemcpyAsync(dbuf[g], values[g], count * sizeof(double), cudaMemcpyHostToDevice, hs[g]->stream));
    }
    g_nccl.GroupStart();
    int rr = 0;
    for (int g = 0; g < n; g++) {
        cudaSetDevice(devs[g]);
        int r = g_nccl.AllReduce(dbuf[g], dbuf[g], (size_t)count, /*ncclDouble*/ 8, /*ncclSum*/ 0, g_comms[g], hs[g]->stream);
        if (r) rr = r;
    }
    int r2 = g_nccl.GroupEnd();
    if (rr || r2) return fail(h0, MADDY_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rr ? rr : r2) : "error");
    for (int g = 0; g < n; g++) {
        CU(hs[g], cudaSetDevice(devs[g]));
        CU(hs[g], cudaMemcpyAsync(values[g], dbuf[g], count * sizeof(double), cudaMemcpyDeviceToHost, hs[g]->stream));
        CU(hs[g], cudaStreamSynchronize(hs[g]->stream));
        cudaFree(dbuf[g]);
    }
    return MADDY_OK;
}

// Ensemble statistics of the per-trajectory energies over ALL handles, device-resident: every handle reduces the energies it
// evaluated last (maddy_energies / maddy_rebuild_and_energies / maddy_snapshot_begin with MADDY_SNAP_ENERGIES) to
// [sum(7), sum of squares(7), count, 0] with one small kernel on its own stream, the 16 doubles are all-reduced with
// ncclAllReduce (one group call, in stream order, so nothing waits on the host), and the result travels to pinned memory on
// the copy stream beside whatever the handle's stream runs next.  _end blocks only until that copy has landed.
extern "C" int maddy_ensemble_stats_begin(maddy_handle **hs, int n)
{
    if (!hs || n <= 0) return MADDY_EINVAL;
    maddy_handle *h0 = hs[0];
    for (int g = 0; g < n; g++)
        if (!hs[g]) return MADDY_EINVAL;
    if (n > 1) {
        if (!g_nccl.load()) return fail(h0, MADDY_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
        std::vector<int> devs(n);
        for (int g = 0; g < n; g++) devs[g] = hs[g]->p.device;
        if (g_comm_devs != devs) {
            for (ncclComm_t c : g_comms) g_nccl.CommDestroy(c);
            g_comms.assign(n, nullptr);
            int r = g_nccl.CommInitAll(g_comms.data(), n, devs.data());
            if (r != 0) {
                g_comms.clear();
                g_comm_devs.clear();
                return fail(h0, MADDY_ENCCL, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
            }
            g_comm_devs = devs;
        }
    }
    for (int g = 0; g < n; g++) {
        maddy_handle *h = hs[g];
        CU(h, cudaSetDevice(h->p.device));
        if (!h->d_ens) CU(h, cudaMalloc(&h->d_ens, 16 * sizeof(double)));
        if (!h->h_ens) CU(h, cudaMallocHost(&h->h_ens, 16 * sizeof(double)));
        if (!h->ens_ready) CU(h, cudaEventCreateWithFlags(&h->ens_ready, cudaEventDisableTiming));
        if (!h->ens_done) CU(h, cudaEventCreateWithFlags(&h->ens_done, cudaEventDisableTiming));
        if (!h->copy_stream) CU(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        cudaError_t e = launch_ensemble_stats(h->a.en_traj, h->a.ntr, h->d_ens, h->stream);
        if (e != cudaSuccess) return fail(h, MADDY_ECUDA, "ensemble_stats_kernel launch: %s", cudaGetErrorString(e));
        h->launches++;
    }
    if (n > 1) {
        g_nccl.GroupStart();
        int rr = 0;
        for (int g = 0; g < n; g++) {
            cudaSetDevice(hs[g]->p.device);
            int r = g_nccl.AllReduce(hs[g]->d_ens, hs[g]->d_ens, 16, /*ncclDouble*/ 8, /*ncclSum*/ 0, g_comms[g], hs[g]->stream);
            if (r) rr = r;
        }
        int r2 = g_nccl.GroupEnd();
        if (rr || r2) return fail(h0, MADDY_ENCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rr ? rr : r2) : "error");
    }
    for (int g = 0; g < n; g++) {
        maddy_handle *h = hs[g];
        CU(h, cudaSetDevice(h->p.device));
        CU(h, cudaEventRecord(h->ens_ready, h->stream));
        CU(h, cudaStreamWaitEvent(h->copy_stream, h->ens_ready, 0));
        CU(h, cudaMemcpyAsync(h->h_ens, h->d_ens, 16 * sizeof(double), cudaMemcpyDeviceToHost, h->copy_stream));
        CU(h, cudaEventRecord(h->ens_done, h->copy_stream));
        h->ens_pending = true;
    }
    return MADDY_OK;
}
extern "C" int maddy_ensemble_stats_end(maddy_handle **hs, int n, double *out16)
{
    if (!hs || n <= 0 || !out16) return MADDY_EINVAL;
    for (int g = 0; g < n; g++) {
        maddy_handle *h = hs[g];
        if (!h) return MADDY_EINVAL;
        if (!h->ens_pending) return fail(h, MADDY_EINVAL, "maddy_ensemble_stats_end without maddy_ensemble_stats_begin");
        CU(h, cudaSetDevice(h->p.device));
        CU(h, cudaEventSynchronize(h->ens_done));
        h->ens_pending = false;
    }
    memcpy(out16, hs[0]->h_ens, 16 * sizeof(double)); // every handle holds the same all-reduced record
    return MADDY_OK;
}

// communicators for these handles' devices (created once per device set)
static int nccl_comms_for(maddy_handle **hs, int n)
{
    maddy_handle *h0 = hs[0];
    if (!g_nccl.load() || !g_nccl.AllGather) return fail(h0, MADDY_ENCCL, "cannot load libnccl.so.2: %s", dlerror());
    std::vector<int> devs(n);
    for (int g = 0; g < n; g++) devs[g] = hs[g]->p.device;
    if (g_comm_devs != devs) {
        for (ncclComm_t c : g_comms) g_nccl.CommDestroy(c);
        g_comms.assign(n, nullptr);
        int r = g_nccl.CommInitAll(g_comms.data(), n, devs.data());
        if (r != 0) {
            g_comms.clear();
            g_comm_devs.clear();
            return fail(h0, MADDY_ENCCL, "ncclCommInitAll: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "error");
        }
        g_comm_devs = devs;
    }
    return MADDY_OK;
}

// The plan of a SHARDED ensemble, all handles in this process (one per GPU, shard order): every shard prepares its transposed
// inputs, one ncclAllGather (a byte per dimer-trajectory cell and array, once per stride - not per event) hands every GPU the
// whole ensemble's, and each evaluates the global plan on its own copy, keeping the slots of its own trajectories.  The draws
// and their order are those of the single-GPU plan and of the reference's host loop.
extern "C" int maddy_hydrolysis_plan_all(maddy_handle **hs, int n, const unsigned *window31, long long first_event, long long period, int n_events,
                                         unsigned flags)
{
    if (!hs || n <= 0 || !window31) return MADDY_EINVAL;
    for (int g = 0; g < n; g++)
        if (!hs[g]) return MADDY_EINVAL;
    if (n == 1) return maddy_hydrolysis_plan(hs[0], window31, first_event, period, n_events, flags);
    maddy_handle *h0 = hs[0];
    int rc = nccl_comms_for(hs, n);
    if (rc) return rc;
    const size_t bytes = (size_t)h0->a.N * h0->a.ntr; // per shard: 2 x (N / 2) x n_tr_local
    for (int g = 0; g < n; g++) {
        maddy_handle *h = hs[g];
        if (h->a.ntr != h0->a.ntr || h->a.N != h0->a.N) return fail(h0, MADDY_EINVAL, "maddy_hydrolysis_plan_all: the shards must be equal blocks");
        void *own = nullptr;
        rc = maddy_hydrolysis_inputs(h, &own, nullptr);
        if (rc) return rc;
        CU(h, cudaSetDevice(h->p.device));
        if ((size_t)n * bytes > h->hyd_all_cap) {
            CU(h, cudaStreamSynchronize(h->stream));
            CU(h, cudaStreamSynchronize(h->aux_stream));
            if (h->d_hyd_all) cudaFree(h->d_hyd_all);
            h->d_hyd_all = nullptr;
            h->hyd_all_cap = 0;
            CU(h, cudaMalloc(&h->d_hyd_all, (size_t)n * bytes));
            h->hyd_all_cap = (size_t)n * bytes;
        }
    }
    g_nccl.GroupStart();
    int rr = 0;
    for (int g = 0; g < n; g++) {
        cudaSetDevice(hs[g]->p.device);
        int r = g_nccl.AllGather(hs[g]->d_hyd_own, hs[g]->d_hyd_all, bytes, /*ncclUint8*/ 1, g_comms[g], hs[g]->stream);
        if (r) rr = r;
    }
    int r2 = g_nccl.GroupEnd();
    if (rr || r2) return fail(h0, MADDY_ENCCL, "ncclAllGather: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rr ? rr : r2) : "error");
    for (int g = 0; g < n; g++) {
        rc = hyd_plan_impl(hs[g], hs[g]->d_hyd_all, n, window31, first_event, period, n_events, flags);
        if (rc) return rc;
    }
    return MADDY_OK;
}
