/*
 * events.cpp — host-side per-stride events and the step loop of the drop-in host.
 *
 * Host events keep the reference's semantics and its use of the single libc rand() stream
 * (src/updater.cpp:74-257): tubule length / on-tubule flags, constant-concentration insertion,
 * hydrolysis, energy and force printing, DCD frames.  The step loop is compute()
 * (src/compute_cuda.cu:1125-1260) re-expressed over the C-ABI: all steps between two host
 * events run inside ONE fused maddy_run() launch per GPU; trajectories are sharded in
 * contiguous blocks over `n_gpus` devices driven by this single host thread, so the order in
 * which host events consume rand() is unchanged.
 */
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <map>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <thread>
#include "mt_host.hpp"

namespace mt {

static const float R_MT = 8.12f, R_MON = 2.0f, ANG_THRES = 1.0f, R_THRES = R_MON * 8;
static const int PF_NUMBER = 13;

// Threads for the host-side loops over the ensemble (flag classification, hydrolysis tests, conversions).  One GPU:
// OMP_NUM_THREADS if it asks for several threads; launchers that pin it to 1 per rank (torchrun) still get a share of a
// many-core host (one sixteenth of the hardware threads), at most 8.  When ONE host thread drives G GPUs (n_gpus = G) the
// per-stride host work is G times larger while the GPU time per stride stays the same, so the share grows with G, up to
// half the hardware threads (at most 32).  MADDY_HOST_THREADS overrides.
static int g_host_gpus = 1;
static int host_threads()
{
#ifdef _OPENMP
    if (const char *e = getenv("MADDY_HOST_THREADS")) {
        const int k = atoi(e);
        if (k > 0) return k;
    }
    const int hw = (int)std::thread::hardware_concurrency();
    int m = omp_get_max_threads();
    if (m < hw / 16) m = hw / 16;
    if (m > 8) m = 8;
    if (m < 1) m = 1;
    if (g_host_gpus > 1) {
        int cap = hw / 2 < 32 ? hw / 2 : 32;
        if (cap < m) cap = m;
        m = m * g_host_gpus < cap ? m * g_host_gpus : cap;
    }
    return m;
#else
    return 1;
#endif
}

// ---------------------------------------------------------------- timers (timer.cpp)
void init_timer(System &s) { s.initial_clock = s.last_clock = (long)clock(); }

static void print_formatted_time(float timer)
{
    int days = (int)(timer / (3600.0f * 24.0f));
    int hours = (int)(timer / 3600.0f - days * 24.0f);
    int minutes = (int)(timer / 60.0f - hours * 60.0f - days * 24.0f * 60.0f);
    int seconds = (int)(timer - hours * 3600.0f - days * 24.0f * 3600.0f - minutes * 60.0f);
    printf("%dd %dh %dm %ds", days, hours, minutes, seconds);
}
static void print_time(System &s, long long step)
{
    float timer = ((float)(clock() - s.initial_clock)) / ((float)CLOCKS_PER_SEC);
    printf("Computation time: ");
    print_formatted_time(timer);
    timer = ((float)(clock() - s.last_clock)) / ((float)CLOCKS_PER_SEC);
    if (step != 0) printf(" (~%f steps/sec)\n", ((float)step) / timer);
    else printf("\n");
}
static void print_estimated_timeleft(System &s, float fraction)
{
    if (fraction != 0.0f) {
        float timer = ((float)(clock() - s.initial_clock)) / ((float)CLOCKS_PER_SEC) * (1.0f / fraction - 1.0f);
        printf("Estimated time left: ");
        print_formatted_time(timer);
        printf(" (%3.1f%% completed)\n", fraction * 100.0f);
    }
}

// ---------------------------------------------------------------- outputs (updater.cpp:3-72)
void output_all_energies(System &s, long long)
{
    for (int t = 0; t < s.par.n_tr; t++) {
        const double *e = &s.energies[(size_t)t * 7];
        if (!std::isfinite(e[0]) || !std::isfinite(e[1]) || !std::isfinite(e[2]) || !std::isfinite(e[6]))
            printf("Some energy in %d trajectory is NaN. NOT Exit program\n", t);
        printf("Energies[%d]:\t%f\t%f\t%f\t%f\t%f\t%f\t%f\n", t, e[0], e[1], e[2], e[3], e[4], e[5], e[6]);
    }
}
void output_sum_force(System &s)
{
    const int N = s.par.n_tot;
    for (int t = 0; t < s.par.n_tr; t++) {
        float sx = 0, sy = 0, sz = 0;
        for (int i = t * N; i < (t + 1) * N; i++) {
            sx += s.f[(size_t)i * 7 + 0];
            sy += s.f[(size_t)i * 7 + 1];
            sz += s.f[(size_t)i * 7 + 2];
        }
        printf("SF for %d traj %.16f\n", t, sqrt(sx * sx + sy * sy + sz * sz));
    }
}
void output_forces(System &s)
{
    const size_t n = (size_t)s.par.n_tot * s.par.n_tr;
    for (size_t i = 0; i < n; i++) {
        const float *f = &s.f[i * 7], *r = &s.r[i * 7];
        printf("Force[%d].theta = %f, fi = %f, psi = %f\n", (int)i, f[4], f[3], f[5]);
        printf("Angle[%d].theta = %f, fi = %f, psi = %f\n", (int)i, r[4], r[3], r[5]);
    }
}

void update(System &s, long long step, std::vector<int> &)
{
    if (!s.quiet) {
        printf("Saving coordinates at step %lld\n", step);
        print_time(s, step);
        print_estimated_timeleft(s, (float)step / (float)s.hp.steps);
    }
    if (s.write_files) {
        save_coord_dcd(s);
        if (s.hp.hydrolysis && (step % (s.hp.stride * 10) == 0)) {
            if (step == 0) {
                if (s.writer) s.writer->drain();
                FILE *first = fopen("dcd/hydrolysis.pdb", "w");
                if (first) fclose(first);
            }
            append_coord_pdb(s);
        }
    }
    if (s.hp.out_force && !s.quiet) {
        output_sum_force(s);
        output_forces(s);
    }
    if (s.hp.out_energy && !s.quiet) output_all_energies(s, step);
}

// ---------------------------------------------------------------- tubule length (updater.cpp:154-227)
void mt_length_classify(System &s, std::vector<int> &mt_len)
{
    const int N = s.par.n_tot;
    // Same predicate as the reference, `rad < r_mt + r_thres && rad > 1 && cosf(theta) > cosf(ang_thres)`; the libm
    // calls are only made where their rounding could matter (cos is even and monotone on [0, pi], sqrt is monotone), so
    // the classification is bit-identical at a fraction of the cost.  Trajectories are independent: split over threads.
    const float cos_thr = cosf(ANG_THRES), rad_hi = R_MT + R_THRES;
    const int Ntr = s.par.n_tr;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if ((size_t)N * Ntr > 65536)
    for (int t = 0; t < Ntr; t++) {
        int sum = 0;
        for (size_t i = (size_t)t * N; i < (size_t)(t + 1) * N; i++) {
            const float *c = &s.r[i * 7];
            const float r2 = c[0] * c[0] + c[1] * c[1];
            bool in_r;
            if (r2 > 1.1f && r2 < 0.99f * rad_hi * rad_hi) in_r = true;
            else {
                float rad = sqrt(r2);
                in_r = (rad < rad_hi) && (rad > 1.0);
            }
            bool on = false;
            if (in_r) {
                const float a = fabsf(c[4]);
                if (a < ANG_THRES - 0.01f) on = true;
                else if (a > ANG_THRES + 0.01f && a < 6.2f - ANG_THRES) on = false;
                else on = cosf(c[4]) > cos_thr;
            }
            s.on_tubule_cur[i] = on ? 1 : 0;
            sum += on;
        }
        mt_len[t] = sum;
    }
}
void mt_length_output(System &s, long long step, const std::vector<int> &mt_len)
{
    const int Ntr = s.par.n_tr;
    if (!s.quiet)
        for (int t = 0; t < Ntr; t++) printf("tubule[%d]: %d\n", t, mt_len[t]);
    if (s.write_files) {
        FILE *f = fopen("mt_len.dat", "a");
        if (f) {
            fprintf(f, "%lld\t", step);
            for (int t = 0; t < s.par.n_tr; t++) fprintf(f, "%f\t", 2 * R_MON * (float)mt_len[t] / PF_NUMBER);
            fprintf(f, "\n");
            fclose(f);
        }
    }
}
void mt_length(System &s, long long step, std::vector<int> &mt_len)
{
    if (step == 0 && s.write_files) {
        FILE *first = fopen("mt_len.dat", "w");
        if (first) fclose(first);
    }
    mt_length_classify(s, mt_len);
    mt_length_output(s, step, mt_len);
}

// ---------------------------------------------------------------- constant concentration (updater.cpp:97-152)
int change_conc(System &s, std::vector<int> &, std::vector<int> &mt_len, std::vector<Insertion> *log)
{
    const maddy_params &par = s.par;
    const int N = par.n_tot;
    int flag = 0;
    for (int tr = 0; tr < par.n_tr; tr++) {
        int num_of_extra = 0;
        for (int i = 0; i < N; i++)
            if (s.extra[i + (size_t)tr * N]) num_of_extra++;
        const float zs = par.rep_h;
        float Vol = float(3.14 * par.rep_r * par.rep_r * zs);
        int NFreeDimers = (N - mt_len[tr] - num_of_extra) / 2;
        while (1.0e7 * NFreeDimers / 6.0 < s.hp.conc * Vol) {
            for (int i = 0; i < N; i += 2) {
                const size_t q = i + (size_t)tr * N;
                if (s.extra[q] && s.mon_type[i] == 0) {
                    s.extra[q] = 0;
                    s.extra[q + 1] = 0;
                    float x, y;
                    for (;;) {
                        x = par.rep_r - 2 * (s.rng.next() % int(par.rep_r));
                        y = par.rep_r - 2 * (s.rng.next() % int(par.rep_r));
                        if (x * x + y * y <= par.rep_r * par.rep_r) { // inside the cylinder
                            if (!s.quiet) printf("New x,y coordinates for extra particle: %f  %f index: %d\n", x, y, (int)q);
                            num_of_extra -= 2;
                            break;
                        }
                    }
                    const float z = zs + par.rep_leftborder + 3 * 2 * R_MON, z2 = z + 2 * R_MON;
                    if (log) {
                        // the caller applies the record (device: maddy_insert_dimers; host frame: once the snapshot of THIS
                        // stride has been collected - s.r may still hold the previous stride's frame, whose output is pending)
                        log->push_back({q, x, y, z, z2});
                    } else {
                        s.r[q * 7 + 0] = x;
                        s.r[q * 7 + 1] = y;
                        s.r[q * 7 + 2] = z;
                        s.r[(q + 1) * 7 + 0] = x;
                        s.r[(q + 1) * 7 + 1] = y;
                        s.r[(q + 1) * 7 + 2] = z2;
                    }
                    flag++;
                    break;
                }
            }
            NFreeDimers += 1;
            if (num_of_extra == 0) {
                if (!s.quiet) printf("No more extra particles for trajectory[%d]!\n", tr);
                break;
            }
        }
        if (!s.quiet)
            printf("Concentration for tajectory[%d]: %f [muMole / L],\t %f [Dimers / nm^3],\t %d [Dimers / Volume],\t  Volume: %f [nm^3]\n",
                   tr, 1.0e7 * NFreeDimers / (6.0 * Vol), NFreeDimers / Vol, NFreeDimers, Vol);
    }
    return flag;
}

static void event_message(System &s, const char *fmt, int a, int b)
{
    if (!s.event_log) {
        printf(fmt, a, b);
        return;
    }
    char buf[128];
    snprintf(buf, sizeof buf, fmt, a, b);
    *s.event_log += buf;
}

// ---------------------------------------------------------------- hydrolysis (updater.cpp:229-257)
void hydrolyse(System &s)
{
    // Same events, same draw order (dimer-outer / trajectory-inner, one draw per eligible dimer) as the reference.  A draw
    // never changes the eligibility of ANOTHER dimer, so the number of draws and the position of every dimer's draw in
    // the stream follow from the eligibility table alone: the table is built in parallel (memory order), the stream of the
    // whole event is pre-drawn by one thread (HostRand::fill), and the 2 % tests run in parallel over the dimer rows.
    const int N = s.par.n_tot, Ntr = s.par.n_tr, nd = N / 2;
    std::vector<unsigned char> elig((size_t)nd * Ntr);
    const bool par = (size_t)N * Ntr > 65536;
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (int tr = 0; tr < Ntr; tr++) {
        const size_t o = (size_t)tr * N;
        for (int d = 0; d < nd; d++) {
            const size_t q = o + 2 * d;
            elig[(size_t)d * Ntr + tr] = s.gtp[q] == 1 && !s.extra[q] && s.on_tubule_cur[q] * s.on_tubule_prev[q] == 1;
        }
    }
    std::vector<size_t> row_first((size_t)nd + 1, 0);
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (int d = 0; d < nd; d++) {
        size_t c = 0;
        const unsigned char *e = &elig[(size_t)d * Ntr];
        for (int tr = 0; tr < Ntr; tr++) c += e[tr];
        row_first[d + 1] = c;
    }
    for (int d = 0; d < nd; d++) row_first[d + 1] += row_first[d];
    std::vector<uint32_t> draws(row_first[nd]);
    s.rng.fill(draws.data(), draws.size());
    std::vector<std::vector<int>> hits(s.quiet ? 0 : nd); // trajectories hydrolysed per dimer, for the messages
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
    for (int d = 0; d < nd; d++) {
        const unsigned char *e = &elig[(size_t)d * Ntr];
        const uint32_t *v = &draws[row_first[d]];
        for (int tr = 0; tr < Ntr; tr++) {
            if (!e[tr]) continue;
            const double prob = (int)*v++ / (double)RAND_MAX;
            if (prob < 0.02) {
                const size_t q = 2 * d + (size_t)tr * N;
                s.gtp[q] = 0;
                s.gtp[q + 1] = 0;
                if (!s.quiet) hits[d].push_back(tr);
            }
        }
    }
    if (!s.quiet)
        for (int d = 0; d < nd; d++)
            for (int tr : hits[d]) event_message(s, "*** Hydrolysis occured to dimer # %d trajectory #%d ***\n", d, tr);
    // GDP dimers that are off the tubule now and at the previous stride return to GTP (no draw)
    if (s.quiet) {
#pragma omp parallel for schedule(static) num_threads(host_threads()) if (par)
        for (int tr = 0; tr < Ntr; tr++) {
            const size_t o = (size_t)tr * N;
            for (int d = 0; d < nd; d++) {
                const size_t q = o + 2 * d;
                if (s.gtp[q] == 0 && !s.extra[q] && s.on_tubule_cur[q] == 0 && s.on_tubule_prev[q] == 0) s.gtp[q] = s.gtp[q + 1] = 1;
            }
        }
    } else {
        for (int i = 0; i < N - 1; i += 2)
            for (int tr = 0; tr < Ntr; tr++) {
                const size_t q = i + (size_t)tr * N;
                if (s.gtp[q] == 0 && !s.extra[q] && s.on_tubule_cur[q] == 0 && s.on_tubule_prev[q] == 0) {
                    s.gtp[q] = 1;
                    s.gtp[q + 1] = 1;
                    event_message(s, "*** Transition to GTP occured to dimer # %d trajectory #%d ***\n", i / 2, tr);
                }
            }
    }
}

// ---------------------------------------------------------------- the step loop
namespace {
struct Shard {
    maddy_handle *h = nullptr;
    int first = 0, count = 0;
};
struct Shards {
    std::vector<Shard> v;
    ~Shards()
    {
        for (Shard &s : v)
            if (s.h) maddy_destroy(s.h);
    }
};
void ck(int rc, maddy_handle *h, const char *what)
{
    if (rc != MADDY_OK) die("%s failed (%d): %s", what, rc, maddy_last_error(h));
}
} // namespace

namespace {
// MADDY_HOST_PROFILE=1: wall-clock breakdown of the host side of the step loop on stderr
struct Prof {
    bool on = getenv("MADDY_HOST_PROFILE") != nullptr;
    std::map<std::string, double> acc;
    double t0 = 0;
    static double now()
    {
        timespec ts;
        clock_gettime(CLOCK_MONOTONIC, &ts);
        return ts.tv_sec + 1e-9 * ts.tv_nsec;
    }
    void begin() { if (on) t0 = now(); }
    void end(const char *what) { if (on) acc[what] += now() - t0; }
    ~Prof()
    {
        if (on)
            for (auto &kv : acc) fprintf(stderr, "[maddy host profile] %-18s %8.3f ms\n", kv.first.c_str(), kv.second * 1e3);
    }
};
} // namespace

void compute(System &s, bool fused, ComputeStats *stats)
{
    Prof prof;
    const bool trace = getenv("MADDY_HOST_TRACE") != nullptr; // wall-clock marks of the first strides on stderr
    const double trace_t0 = Prof::now();
    auto mark = [&](const char *what, long long at) {
        if (trace && at <= 3 * s.hp.stride) fprintf(stderr, "[trace] %9.3f ms  step %lld  %s\n", (Prof::now() - trace_t0) * 1e3, at, what);
    };
    prof.begin();
    const maddy_params &par = s.par;
    const HostParams &hp = s.hp;
    const int N = par.n_tot, Ntr = par.n_tr;
    const size_t n = (size_t)N * Ntr;
    ComputeStats st;

    // initIntegration: forces zeroed; the angle wrap happens inside maddy_create on the device copy
    // AND on the host copy (compute_cuda.cu:995-1010 modifies r in place)
    std::fill(s.f.begin(), s.f.end(), 0.f);
    CheckpointState resume_state;
    if (hp.resume) checkpoint_load(s, hp.checkpoint, resume_state); // coordinates, flags, host rand(): exactly as they stood
    else
        for (size_t q = 0; q < n; q++) {
            float *c = &s.r[q * 7];
            c[3] -= (2 * M_PI) * (int)(c[3] / (2 * M_PI));
            c[5] -= (2 * M_PI) * (int)(c[5] / (2 * M_PI));
            c[4] -= (2 * M_PI) * (int)(c[4] / (2 * M_PI));
        }

    // contiguous trajectory blocks per GPU
    int G = hp.n_gpus < 1 ? 1 : hp.n_gpus;
    if (G > Ntr) G = Ntr;
    Shards sh;
    sh.v.resize(G);
    g_host_gpus = G;
    for (int g = 0; g < G; g++) {
        Shard &d = sh.v[g];
        d.first = (int)((long long)Ntr * g / G);
        d.count = (int)((long long)Ntr * (g + 1) / G) - d.first;
        maddy_params p = par;
        p.traj_first = d.first;
        p.n_tr_local = d.count;
        p.device = par.device + g;
        maddy_topology top = s.topology_view(d.first);
        int rc = maddy_create(&p, &top, &s.r[(size_t)d.first * N * 7], nullptr, &d.h);
        if (rc != MADDY_OK) die("maddy_create failed (%d): %s", rc, maddy_last_error(nullptr));
        st.h2d_bytes += (double)d.count * N * (7 * 4 + 32 + 3);
        if (hp.resume) { // raw coordinate bits (maddy_create re-wraps angles) and the RNG streams of this shard
            const size_t cnt = (size_t)d.count * N, off = (size_t)d.first * N;
            std::vector<unsigned> rs(cnt * 8);
            memcpy(rs.data(), &resume_state.rng[off * 4], cnt * 16);
            memcpy(rs.data() + cnt * 4, &resume_state.rng[(n + off) * 4], cnt * 16);
            ck(maddy_upload_coords(d.h, &s.r[off * 7]), d.h, "maddy_upload_coords");
            ck(maddy_upload_rng(d.h, rs.data()), d.h, "maddy_upload_rng");
        }
    }
    prof.end("create");
    if (!s.quiet) printf("Using %d device(s), first device %d\n", G, par.device);
    struct WriterScope { // background trajectory output for the duration of the loop
        System &s;
        explicit WriterScope(System &sys) : s(sys) { if (s.write_files) s.writer = std::make_shared<OutputWorker>(); }
        ~WriterScope()
        {
            s.writer.reset();       // joins the thread (pending frames are written first)
            s.writer_files.reset(); // closes the per-trajectory files
        }
    } writer_scope(s);
    s.energies.assign((size_t)Ntr * 7, 0.0);
    // periodic ensemble reduction: per-term sum / sum of squares / count of the per-trajectory energies, reduced
    // on the GPUs and all-reduced with NCCL in every stride block with energies; one line per stride in ensemble.dat
    const bool ens_on = hp.out_energy && (G > 1 || getenv("MADDY_ENSEMBLE_STATS"));
    std::vector<maddy_handle *> ens_handles;
    for (Shard &d : sh.v) ens_handles.push_back(d.h);
    FILE *ens_file = nullptr;
    struct EnsClose {
        FILE *&f;
        ~EnsClose() { if (f) fclose(f); }
    } ens_close{ens_file};
    if (ens_on) {
        s.ensemble_stats.assign(MADDY_ENSEMBLE_STATS, 0.0);
        if (s.write_files) ens_file = fopen("ensemble.dat", hp.resume ? "a" : "w");
    }
    auto ens_begin = [&] {
        if (!ens_on) return;
        int rc = maddy_ensemble_stats_begin(ens_handles.data(), G);
        if (rc != MADDY_OK) die("maddy_ensemble_stats_begin failed (%d): %s", rc, maddy_last_error(ens_handles[0]));
    };
    auto ens_end = [&](long long at) {
        if (!ens_on) return;
        int rc = maddy_ensemble_stats_end(ens_handles.data(), G, s.ensemble_stats.data());
        if (rc != MADDY_OK) die("maddy_ensemble_stats_end failed (%d): %s", rc, maddy_last_error(ens_handles[0]));
        st.d2h_bytes += (double)G * MADDY_ENSEMBLE_STATS * 8;
        if (!ens_file) return;
        const double *q = s.ensemble_stats.data(), cnt = q[14] > 0 ? q[14] : 1.0;
        fprintf(ens_file, "%lld\t%d", at, (int)q[14]); // step, trajectories, then mean and standard deviation of each energy term
        for (int k = 0; k < 7; k++) {
            const double mean = q[k] / cnt, var = q[7 + k] / cnt - mean * mean;
            fprintf(ens_file, "\t%f\t%f", mean, var > 0 ? sqrt(var) : 0.0);
        }
        fputc('\n', ens_file);
    };

    std::vector<int> mt_len(Ntr, 0), mt_len_prev(Ntr, 0);
    auto for_each = [&](auto fn) {
        for (Shard &d : sh.v) fn(d);
    };

    // Host events need the device only where the reference order makes it observable:
    //  * lists are rebuilt BEFORE the host events of a step (compute_cuda.cu:1140-1151); that only matters when an
    //    event can move particles afterwards (constant-concentration insertion at a stride step) — otherwise the
    //    rebuild is folded into the fused window launch;
    //  * hydrolyse() reads only host flags from the previous stride, so the draw for the NEXT event step is made
    //    while the GPU runs the current window (same rand() order: it follows every event of the current step).
    // Overlapped stride read-back: possible whenever the host results of a stride (on-tubule flags, insertion) cannot
    // change the forces of the following steps.
    // Where they CAN (on-tubule flags with a non-zero barrier amplitude, constant-concentration insertion), the events run
    // on the device side of the boundary: mt_length()'s classification is evaluated by the snapshot itself
    // (MADDY_SNAP_ONTUBULE, exact) and applied in place, the host waits only for the Ntr counts (microseconds) and hands
    // insertions over as sparse records (maddy_insert_dimers) - no full-coordinate round trip stands between two windows.
    const bool flags_matter = par.barrier && (par.a_barr_long != 0.f || par.a_barr_lat != 0.f);
    bool dev_classify = fused && hp.tub_length && !getenv("MADDY_HOST_EVENTS");
    for (Shard &d : sh.v) dev_classify = dev_classify && maddy_has_exact_on_tubule(d.h);
    const bool feedback = flags_matter || (hp.is_const_conc && hp.tub_length); // host results of a stride change the next forces
    const bool overlap_stride = fused && !par.tea_on && !hp.out_force && (!feedback || dev_classify) && !getenv("MADDY_NO_OVERLAP");
    // The next window waits for the Ntr counts (not the coordinates) where the host needs them first: for the insertions of
    // constant concentration, and with several shards whose flags feed back into the forces (a shard cannot run ahead on a
    // classification another shard could not decide).
    const bool hyd_possible = overlap_stride && dev_classify && hp.hydrolysis && hp.hydrostep > 0 && Ntr % G == 0 && N % 2 == 0 &&
                              !getenv("MADDY_HOST_HYDROLYSIS");
    const bool counts_first = overlap_stride && dev_classify && ((hp.is_const_conc && hp.tub_length) || ((feedback || hyd_possible) && G > 1));
    // hydrolyse() itself runs on the device when ONE handle holds the ensemble (draw positions are global): right after a
    // stride block every event up to the next stride step is evaluated in one go (maddy_hydrolysis_plan) from the 31 words
    // of the host generator and left as the GTP schedule of the fused loop - a window then spans the whole stride and the
    // host only advances its generator by the number of draws the device reports.  May be switched off mid-run (a
    // classification the device could not decide): the host then carries on from the synchronised state.
    // With several shards every GPU evaluates the ensemble's plan from the gathered inputs (maddy_hydrolysis_plan_all).
    bool hyd_dev = hyd_possible;
    // Everywhere else nothing is waited for: plan and window are queued right behind the snapshot, GUARDED - should the device
    // be unable to decide a classification, they return without touching the state and the host redoes the stride itself.
    const bool guarded = overlap_stride && dev_classify && !counts_first && G == 1 && (feedback || hyd_dev);
    bool plan_pending = false;           // a plan whose draw count has not been folded into s.rng yet
    long long plan_first = 0;            // step of its first event
    int plan_events = 0;
    auto plan_covers = [&](long long at) {
        return plan_events > 0 && at >= plan_first && (at - plan_first) % hp.hydrostep == 0 && (at - plan_first) / hp.hydrostep < plan_events;
    };
    // result of the pending plan: the host generator jumps over its draws, and the events' messages are printed in the
    // reference's order (the schedule holds the GTP state after every event; s.gtp the state the plan started from)
    auto collect_plan = [&] {
        if (!plan_pending) return;
        plan_pending = false;
        unsigned long long total = 0;
        std::vector<int> slots; // [event][global monomer]
        if (!s.quiet) slots.resize((size_t)plan_events * n);
        for_each([&](Shard &d) {
            const size_t cnt = (size_t)d.count * N, off = (size_t)d.first * N;
            std::vector<int> part;
            if (!s.quiet) part.resize((size_t)plan_events * cnt);
            unsigned long long t = 0;
            ck(maddy_hydrolysis_result(d.h, &t, nullptr, s.quiet ? nullptr : part.data()), d.h, "maddy_hydrolysis_result");
            total = t; // the ensemble's, identical on every shard
            for (int k = 0; k < plan_events && !s.quiet; k++) memcpy(&slots[(size_t)k * n + off], &part[(size_t)k * cnt], cnt * sizeof(int));
        });
        s.rng.discard(total);
        st.d2h_bytes += 8.0 * (plan_events + 1) + (s.quiet ? 0.0 : (double)plan_events * n);
        if (s.quiet) return;
        for (int k = 0; k < plan_events; k++) {
            const int *before = k == 0 ? s.gtp.data() : &slots[(size_t)(k - 1) * n], *after = &slots[(size_t)k * n];
            for (int dm = 0; dm < N / 2; dm++)
                for (int tr = 0; tr < Ntr; tr++) {
                    const size_t q = 2 * dm + (size_t)tr * N;
                    if (before[q] == 1 && after[q] == 0) printf("*** Hydrolysis occured to dimer # %d trajectory #%d ***\n", dm, tr);
                }
            for (int dm = 0; dm < N / 2; dm++)
                for (int tr = 0; tr < Ntr; tr++) {
                    const size_t q = 2 * dm + (size_t)tr * N;
                    if (before[q] == 0 && after[q] == 1) printf("*** Transition to GTP occured to dimer # %d trajectory #%d ***\n", dm, tr);
                }
        }
    };
    // the GTP flags as they stand on the device (hyd_dev mode keeps s.gtp current only at strides)
    auto download_gtp = [&] {
        for_each([&](Shard &d) {
            ck(maddy_snapshot_begin(d.h, MADDY_SNAP_GTP), d.h, "maddy_snapshot_begin");
            ck(maddy_snapshot_end(d.h, nullptr, nullptr, nullptr), d.h, "maddy_snapshot_end");
            ck(maddy_snapshot_gtp(d.h, &s.gtp[(size_t)d.first * N]), d.h, "maddy_snapshot_gtp");
        });
        st.d2h_bytes += (double)n;
    };
    int pending_output = 0; // an overlapped stride whose update() is still to be written
    long long pending_step = 0;
    std::string pending_log;
    auto flush_pending = [&] {
        update(s, pending_step, mt_len);
        s.gtp_for_output.reset();
        if (!pending_log.empty()) fputs(pending_log.c_str(), stdout);
        pending_log.clear();
        pending_output = 0;
    };
    long long step = 0;
    long long hydrolysed_for = -1; // event step whose hydrolysis has already been evaluated on the host
    const bool pending_snapshot_open = false; // (a checkpoint is taken at the top of a step: no snapshot is in flight there)
    // Windows that span several hydrolysis events (overlapped mode only).  Every event of a stride depends on the flags
    // classified at that stride alone, so the events of the NEXT window are drawn (same rand() order) while the GPU runs the
    // current one and handed over as a GTP schedule: window lengths grow by one event period per window inside a stride
    // (100, 100, 200, 300, 300 steps at the template's periods) - fewer window boundaries, the host stays ahead of the GPU.
    const bool multi_event = overlap_stride && hp.hydrolysis && hp.hydrostep > 0 && !getenv("MADDY_SINGLE_EVENT_WINDOWS");
    std::vector<std::vector<int>> sched_gtp; // GTP state after each pre-drawn event of the next window, first at sched_first
    long long sched_first = -1, sched_end = -1; // [sched_first, sched_end) = the window those events belong to
    if (hp.resume) {
        step = resume_state.step;
        hydrolysed_for = resume_state.hydrolysed_for;
        mt_len = resume_state.mt_len;
        mt_len_prev = resume_state.mt_len_prev;
    }
    const long long start_step = step;
    // checkpoint = the state at the TOP of step `at`, before any of that step's events
    // Invariant: checkpoints are taken only at window boundaries that are stride steps (system.cpp forces checkpoint_freq to be
    // a multiple of stride), where no pre-drawn hydrolysis schedule is pending; CheckpointState does not carry sched_gtp.
    auto write_checkpoint = [&](long long at) {
        if (!sched_gtp.empty()) die("checkpoint at step %lld with a pending GTP schedule (checkpoint_freq must be a multiple of stride)", at);
        if (pending_output) flush_pending();
        if (hyd_dev) {
            // the device has evaluated the events up to and including `at` already: fold them in (generator, messages) and
            // record `at` as done, so that a resumed run continues exactly where the uninterrupted one does
            if (pending_snapshot_open) die("checkpoint at step %lld with a snapshot in flight", at);
            collect_plan();
            if (plan_covers(at)) {
                for_each([&](Shard &d) { ck(maddy_apply_scheduled_gtp(d.h, at), d.h, "maddy_apply_scheduled_gtp"); });
                hydrolysed_for = at;
            }
            download_gtp();
        }
        CheckpointState out;
        out.step = at;
        out.hydrolysed_for = hydrolysed_for;
        out.mt_len = mt_len;
        out.mt_len_prev = mt_len_prev;
        out.coords.resize(n * 7);
        out.rng.resize(n * 8);
        for_each([&](Shard &d) {
            const size_t cnt = (size_t)d.count * N, off = (size_t)d.first * N;
            std::vector<unsigned> rs(cnt * 8);
            ck(maddy_download_coords(d.h, &out.coords[off * 7]), d.h, "maddy_download_coords");
            ck(maddy_download_rng(d.h, rs.data()), d.h, "maddy_download_rng");
            memcpy(&out.rng[off * 4], rs.data(), cnt * 16);
            memcpy(&out.rng[(n + off) * 4], rs.data() + cnt * 4, cnt * 16);
        });
        if (s.writer) s.writer->drain(); // the frames up to here are on disk before the checkpoint says so
        checkpoint_save(s, hp.checkpoint, out);
    };
    // ---- in-situ analysis (extension key insitu_analysis): per stride frame and trajectory, the line
    // scripts/temp_calc prints (main.cpp:94-107, with this run's gammaR / gammaTheta / dt / Ntot / k_B instead of the
    // tool's hard-coded ones) and the timeline line of scripts/disas_speed/disc (disc.cpp:160-166)
    const int n_pf = 13; // pf_number, disc.cpp:12
    long insitu_frames = 0;
    std::vector<FILE *> insitu_temp, insitu_disc;
    struct InsituClose {
        std::vector<FILE *> &a, &b;
        ~InsituClose()
        {
            for (FILE *f : a) if (f) fclose(f);
            for (FILE *f : b) if (f) fclose(f);
        }
    } insitu_close{insitu_temp, insitu_disc};
    if (hp.insitu) {
        std::vector<int> chain(N), resid(N);
        std::vector<char> name1(N);
        for (int i = 0; i < N; i++) {
            const PDBAtom &at = s.pdb.atoms[i];
            const int c = at.chain - 'A';
            chain[i] = (c >= 0 && c < n_pf) ? c : -1;
            resid[i] = at.resid;
            name1[i] = at.name[0] ? at.name[1] : ' ';
        }
        for_each([&](Shard &d) {
            ck(maddy_analysis_setup(d.h, chain.data(), resid.data(), name1.data(), n_pf), d.h, "maddy_analysis_setup");
            ck(maddy_analysis_reference(d.h), d.h, "maddy_analysis_reference");
        });
        if (s.write_files)
            for (int t = 0; t < Ntr; t++) {
                insitu_temp.push_back(fopen((hp.dcd_xyz[t] + ".temp.dat").c_str(), hp.resume ? "a" : "w"));
                insitu_disc.push_back(fopen((hp.dcd_xyz[t] + ".disc.dat").c_str(), hp.resume ? "a" : "w"));
            }
    }
    auto insitu_frame = [&](long long at) {
        insitu_frames++;
        std::vector<double> sums((size_t)Ntr * 8);
        std::vector<int> pf((size_t)Ntr * n_pf * 3);
        for_each([&](Shard &d) {
            if (insitu_frames > 1) ck(maddy_analysis_temperature(d.h, &sums[(size_t)d.first * 8]), d.h, "maddy_analysis_temperature");
            else ck(maddy_analysis_reference(d.h), d.h, "maddy_analysis_reference");
            ck(maddy_analysis_protofilaments(d.h, &pf[(size_t)d.first * n_pf * 3]), d.h, "maddy_analysis_protofilaments");
        });
        st.d2h_bytes += (double)Ntr * (8 * 8 + n_pf * 12);
        if (insitu_temp.empty()) return;
        const double kB = 0.0019872041; // kcal/(mol K), mt.h:39
        const double six = 6.0 * (double)hp.stride * par.dt * N * kB, two = six / 3.0;
        for (int t = 0; t < Ntr; t++) {
            if (insitu_frames > 1 && insitu_temp[t]) {
                const double *q = &sums[(size_t)t * 8];
                fprintf(insitu_temp[t], "%ld %f %f %15f %f %f %15f %f %f\n", insitu_frames, q[0] * par.gammaR / six, q[1] * par.gammaTheta / six,
                        q[2] * par.gammaR / two, q[3] * par.gammaR / two, q[4] * par.gammaR / two, q[5] * par.gammaTheta / two,
                        q[6] * par.gammaTheta / two, q[7] * par.gammaTheta / two);
            }
            if (insitu_disc[t]) {
                float lt = 0;
                for (int c = 0; c < n_pf; c++) lt += (float)pf[((size_t)t * n_pf + c) * 3 + 2] / 13.0;
                fprintf(insitu_disc[t], "%ld %f", insitu_frames, 2 * lt);
                for (int c = 0; c < n_pf; c++) {
                    const int *o = &pf[((size_t)t * n_pf + c) * 3];
                    fprintf(insitu_disc[t], " %d:%d:%d", o[0], o[1], o[2]); // pf_end_number : curled_start : mt_end_number
                }
                fputc('\n', insitu_disc[t]);
            }
        }
        (void)at;
    };
    while (step < hp.steps) {
        if (!hp.checkpoint.empty() && hp.checkpoint_freq > 0 && step != start_step && step % hp.checkpoint_freq == 0) write_checkpoint(step);
        const bool rebuild_now = step % par.ljpairsupdatefreq == 0;
        const bool stride_now = step % hp.stride == 0;
        const bool may_teleport = stride_now && hp.tub_length && step != 0 && hp.is_const_conc;
        // energies printed at a stride step are evaluated on the lists rebuilt at that step (compute_cuda.cu:1140-1170)
        const bool explicit_rebuild = rebuild_now && ((stride_now && hp.out_energy) || may_teleport || par.tea_on || !fused);
        prof.begin();
        const bool fused_stride_energy = explicit_rebuild && stride_now && hp.out_energy; // rebuild + energies in one launch below
        if (explicit_rebuild && !fused_stride_energy) {
            for_each([&](Shard &d) {
                if (par.lj_on) ck(maddy_rebuild_lj(d.h), d.h, "maddy_rebuild_lj");
                if (par.is_assembly) ck(maddy_rebuild_bonds(d.h), d.h, "maddy_rebuild_bonds");
            });
        }
        prof.end("explicit rebuild");
        prof.begin();
        // ---- hydrolysis (compute_cuda.cu:1153-1160)
        long long scheduled_end = -1;
        if (hyd_dev && hp.hydrolysis && step % hp.hydrostep == 0 && step != 0 && plan_covers(step)) {
            // evaluated on the device at the last stride; inside a window the fused loop applies the slot itself, at a
            // stride step it has to be current before the stride block evaluates the energies
            if (stride_now) for_each([&](Shard &d) { ck(maddy_apply_scheduled_gtp(d.h, step), d.h, "maddy_apply_scheduled_gtp"); });
            hydrolysed_for = step;
        } else if (hp.hydrolysis && step % hp.hydrostep == 0 && step != 0) {
            if (!sched_gtp.empty() && sched_first == step) {
                // the events of this window were drawn beside the previous one: slot k becomes current at step + k periods
                const int slots = (int)sched_gtp.size();
                for_each([&](Shard &d) {
                    const size_t cnt = (size_t)d.count * N, off = (size_t)d.first * N;
                    std::vector<int> buf((size_t)slots * cnt);
                    for (int k = 0; k < slots; k++) memcpy(&buf[(size_t)k * cnt], &sched_gtp[k][off], cnt * sizeof(int));
                    ck(maddy_schedule_gtp(d.h, step, hp.hydrostep, slots, buf.data()), d.h, "maddy_schedule_gtp");
                });
                st.h2d_bytes += (double)n * slots;
                scheduled_end = sched_end;
                sched_gtp.clear();
                mark("gtp schedule uploaded", step);
            } else {
                if (hydrolysed_for != step) hydrolyse(s);
                for_each([&](Shard &d) { ck(maddy_upload_gtp(d.h, &s.gtp[(size_t)d.first * N]), d.h, "maddy_upload_gtp"); });
                st.h2d_bytes += (double)n;
                mark("gtp uploaded", step);
            }
        }
        prof.end("upload gtp");
        prof.begin();
        // ---- stride block (compute_cuda.cu:1163-1226).  Everything the device needs (energies, downloads, on-tubule
        // and insertion uploads) happens here; the output part (update(): stdout, DCD frames) is deferred until the next
        // window has been queued, so the GPU does not wait for host formatting.
        if (hp.insitu && stride_now) insitu_frame(step);
        int deferred_output = 0;
        const bool overlapped = stride_now && overlap_stride;
        const bool classify = overlapped && dev_classify && step != 0; // mt_length()'s classification rides on the snapshot
        bool collected = false;                                        // the snapshot has been collected before the window launch
        std::vector<Insertion> inserted;                               // this stride's insertions (re-applied to the collected frame)
        // the host events of a stride once its coordinates are in s.r (compute_cuda.cu:1181-1219); returns deferred_output
        auto host_events = [&]() -> int {
            if (!hp.tub_length) return 1;
            s.on_tubule_prev = s.on_tubule_cur;
            if (step == 0) return 2; // step 0: update() first, then mt_length() (host flags only, nothing is uploaded)
            mt_len_prev = mt_len;
            mt_length(s, step, mt_len);
            if (par.barrier) {
                for_each([&](Shard &d) { ck(maddy_upload_on_tubule(d.h, &s.on_tubule_cur[(size_t)d.first * N]), d.h, "maddy_upload_on_tubule"); });
                st.h2d_bytes += (double)n;
            }
            if (hp.is_const_conc) {
                for (int t = 0; t < Ntr; t++) mt_len_prev[t] = mt_len[t] - mt_len_prev[t];
                if (change_conc(s, mt_len_prev, mt_len)) {
                    for_each([&](Shard &d) {
                        ck(maddy_upload_extra(d.h, &s.extra[(size_t)d.first * N]), d.h, "maddy_upload_extra");
                        ck(maddy_upload_coords(d.h, &s.r[(size_t)d.first * N * 7]), d.h, "maddy_upload_coords");
                    });
                    st.h2d_bytes += (double)n * 33;
                }
            }
            return 1;
        };
        auto collect = [&] {
            for_each([&](Shard &d) {
                ck(maddy_snapshot_end(d.h, &s.r[(size_t)d.first * N * 7], nullptr, hp.out_energy ? &s.energies[(size_t)d.first * 7] : nullptr),
                   d.h, "maddy_snapshot_end");
            });
            if (hp.out_energy) ens_end(step);
        };
        const bool classify0 = overlapped && hyd_dev && step == 0; // the device needs the step-0 classification too (hydrolysis flags only)
        if (overlapped) {
            // The read-back is only QUEUED here; it is collected after the next window has been launched.
            if (hyd_dev) collect_plan(); // the generator passes the events of the last stride BEFORE this stride's insertion draws
            const unsigned what = MADDY_SNAP_COORDS | (hp.out_energy ? MADDY_SNAP_ENERGIES : 0u) | (fused_stride_energy ? MADDY_SNAP_REBUILD : 0u) |
                                  (classify ? MADDY_SNAP_ONTUBULE | (par.barrier ? MADDY_SNAP_ONTUBULE_APPLY : 0u) : 0u) |
                                  (classify0 ? MADDY_SNAP_ONTUBULE : 0u) | (hyd_dev ? MADDY_SNAP_GTP : 0u) |
                                  (guarded && (classify || classify0) ? MADDY_SNAP_ONTUBULE_GUARD : 0u);
            for_each([&](Shard &d) { ck(maddy_snapshot_begin(d.h, what), d.h, "maddy_snapshot_begin"); });
            if (hp.out_energy) ens_begin();
            st.d2h_bytes += (double)n * 32 + (hp.out_energy ? (double)Ntr * 7 * 8 : 0.0) + (classify || classify0 ? (double)n + 4.0 * Ntr : 0.0) +
                            (hyd_dev ? (double)n : 0.0);
            if (classify0 && !guarded) { // several shards: nobody runs ahead of a classification that may be undecided
                int undecided = 0;
                for_each([&](Shard &d) {
                    int u = 0;
                    ck(maddy_snapshot_tubule_lengths(d.h, nullptr, &u), d.h, "maddy_snapshot_tubule_lengths");
                    undecided |= u;
                });
                if (undecided) hyd_dev = false; // s.gtp and the generator are still the host's own
            }
            if (classify && counts_first) {
                // What the host derives from this stride feeds back into the next forces (non-zero barrier: the flags, already
                // applied on the device; constant concentration: the insertions).  Only the Ntr counts are waited for.
                std::vector<int> counts(Ntr);
                int undecided = 0;
                for_each([&](Shard &d) {
                    int u = 0;
                    ck(maddy_snapshot_tubule_lengths(d.h, &counts[d.first], &u), d.h, "maddy_snapshot_tubule_lengths");
                    undecided |= u;
                });
                mark("tubule lengths collected", step);
                if (undecided) { // some |theta| outside the exact rule's range: this stride takes the reference's serial block
                    collect();
                    collected = true;
                    if (hyd_dev) { // ... and hydrolysis returns to the host for the rest of the run, from the synchronised state
                        for_each([&](Shard &d) { ck(maddy_snapshot_gtp(d.h, &s.gtp[(size_t)d.first * N]), d.h, "maddy_snapshot_gtp"); });
                        hyd_dev = false;
                        plan_events = 0;
                    }
                    deferred_output = host_events();
                } else {
                    mt_len_prev = mt_len;
                    mt_len = counts;
                    mt_length_output(s, step, mt_len);
                    if (hp.is_const_conc) {
                        for (int t = 0; t < Ntr; t++) mt_len_prev[t] = mt_len[t] - mt_len_prev[t];
                        if (change_conc(s, mt_len_prev, mt_len, &inserted)) {
                            for_each([&](Shard &d) {
                                std::vector<int> idx;
                                std::vector<float> xyzz;
                                const size_t lo = (size_t)d.first * N, hi = lo + (size_t)d.count * N;
                                for (const Insertion &e : inserted)
                                    if (e.q >= lo && e.q < hi) {
                                        idx.push_back((int)(e.q - lo));
                                        xyzz.insert(xyzz.end(), {e.x, e.y, e.z, e.z2});
                                    }
                                ck(maddy_insert_dimers(d.h, (int)idx.size(), idx.data(), xyzz.data()), d.h, "maddy_insert_dimers");
                            });
                            st.h2d_bytes += (double)inserted.size() * 20;
                        }
                    }
                }
            }
            if (hyd_dev) {
                // every hydrolysis event up to (and including) the next stride step, on the device, from the 31 words of
                // the host generator as they stand after this stride's insertion draws
                const long long h = hp.hydrostep;
                const long long first = (step / h + 1) * h, last = std::min((step / hp.stride + 1) * hp.stride, hp.steps - 1);
                plan_events = first <= last ? (int)((last - first) / h + 1) : 0;
                plan_first = first;
                if (plan_events > 0) {
                    uint32_t w[31];
                    s.rng.window(w);
                    int rcp = maddy_hydrolysis_plan_all(ens_handles.data(), G, w, first, h, plan_events, s.quiet ? 0u : MADDY_HYD_KEEP_SLOTS);
                    if (rcp != MADDY_OK) die("maddy_hydrolysis_plan_all failed (%d): %s", rcp, maddy_last_error(ens_handles[0]));
                    plan_pending = true;
                    st.h2d_bytes += 124.0;
                    mark("hydrolysis plan queued", step);
                }
            }
        } else if (stride_now) {
            if (hp.out_energy) {
                for_each([&](Shard &d) {
                    double *out = &s.energies[(size_t)d.first * 7];
                    if (fused_stride_energy) ck(maddy_rebuild_and_energies(d.h, out, nullptr), d.h, "maddy_rebuild_and_energies");
                    else ck(maddy_energies(d.h, out, nullptr), d.h, "maddy_energies");
                });
                st.d2h_bytes += (double)Ntr * 7 * 8;
                ens_begin();
                ens_end(step);
            }
            if (hp.out_force) {
                for_each([&](Shard &d) { ck(maddy_download_forces(d.h, &s.f[(size_t)d.first * N * 7]), d.h, "maddy_download_forces"); });
                st.d2h_bytes += (double)n * 32;
            }
            for_each([&](Shard &d) { ck(maddy_download_coords(d.h, &s.r[(size_t)d.first * N * 7]), d.h, "maddy_download_coords"); });
            st.d2h_bytes += (double)n * 32;
            deferred_output = host_events();
        }
        prof.end("stride block");
        prof.begin();
        // ---- steps up to the next host event
        long long next = hp.steps, count = 0, split_at = 0;
        const bool stepwise = !fused; // TEA windows are queued by maddy_run as well (force + prepare in one launch)
        const bool hydro = hp.hydrolysis && hp.hydrostep > 0;
        auto window_end = [&] {
        next = std::min(hp.steps, (step / hp.stride + 1) * hp.stride);
        // One window per hydrolysis period: hydrolyse() for the NEXT event is evaluated on the host while the GPU runs
        // the current window (see below), so the events cost no GPU idle time.  (maddy_schedule_gtp can fold several
        // events into one launch, but their evaluation would then sit between two windows instead of beside one.)
        // Events evaluated on the device are applied inside the window.  The window that follows a stride block stops at the
        // first of them: the plan queued with the stride block is evaluated on a stream of its own BESIDE that window, and
        // the next one - which runs to the next stride step - waits for it.
        const long long first_event = (step / hp.hydrostep + 1) * (long long)hp.hydrostep;
        split_at = 0;
        if (hydro && hyd_dev && plan_covers(first_event)) {
            if (stride_now && first_event < next) split_at = first_event; // both windows are queued now: the host's stride output must not delay the second
        } else if (hydro) {
            next = std::min(next, scheduled_end > step ? scheduled_end : first_event);
        }
        count = next - step;
        };
        window_end();
        if (stepwise) {
            for (long long q = step; q < next; q++) {
                for_each([&](Shard &d) {
                    if (q != step && q % par.ljpairsupdatefreq == 0) {
                        if (par.lj_on) ck(maddy_rebuild_lj(d.h), d.h, "maddy_rebuild_lj");
                        if (par.is_assembly) ck(maddy_rebuild_bonds(d.h), d.h, "maddy_rebuild_bonds");
                    }
                    ck(maddy_force(d.h), d.h, "maddy_force");
                    if (par.tea_on) {
                        ck(maddy_tea_update(d.h, q), d.h, "maddy_tea_update");
                        ck(maddy_tea_integrate(d.h), d.h, "maddy_tea_integrate");
                    } else {
                        ck(maddy_integrate(d.h), d.h, "maddy_integrate");
                    }
                });
            }
        } else {
            auto launch_windows = [&] {
                const unsigned fl = explicit_rebuild ? MADDY_RUN_SKIP_FIRST_REBUILD : 0u;
                if (split_at > step) {
                    for_each([&](Shard &d) { ck(maddy_run(d.h, step, split_at - step, fl), d.h, "maddy_run"); });
                    for_each([&](Shard &d) { ck(maddy_run(d.h, split_at, next - split_at, 0u), d.h, "maddy_run"); });
                } else {
                    for_each([&](Shard &d) { ck(maddy_run(d.h, step, count, fl), d.h, "maddy_run"); });
                }
            };
            launch_windows();
            mark("window launched", step);
        }
        prof.end("launch window");
        prof.begin();
        // Output of the PREVIOUS overlapped stride (it reads s.r / s.energies, which the collect below overwrites): by now
        // the window after that stride and this one are both queued.
        if (pending_output) {
            flush_pending();
            mark("pending output flushed", step);
        }
        prof.end("stride output");
        prof.begin();
        if (overlapped && !collected) {
            collect();
            if (hyd_dev) // state after the event of this step, if any
                for_each([&](Shard &d) { ck(maddy_snapshot_gtp(d.h, &s.gtp[(size_t)d.first * N]), d.h, "maddy_snapshot_gtp"); });
            prof.end("stride collect (wait + transpose)");
            mark("snapshot collected", step);
            prof.begin();
            // A guarded classification the device could not decide: plan and window returned without touching the state.  The
            // host takes the stride over - guard cleared, hydrolysis back on the host for good (from the synchronised state:
            // s.gtp was just read back, the generator has not moved), host classification, and the window queued again.
            int undecided = 0;
            if ((classify && !counts_first) || classify0)
                for_each([&](Shard &d) {
                    int u = 0;
                    ck(maddy_snapshot_tubule_lengths(d.h, nullptr, &u), d.h, "maddy_snapshot_tubule_lengths");
                    undecided |= u;
                });
            const bool redo = undecided && guarded;
            if (redo) {
                for_each([&](Shard &d) { ck(maddy_clear_guard(d.h), d.h, "maddy_clear_guard"); });
                if (hyd_dev) {
                    if (plan_pending) ck(maddy_hydrolysis_result(sh.v[0].h, nullptr, nullptr, nullptr), sh.v[0].h, "maddy_hydrolysis_result");
                    ck(maddy_schedule_gtp(sh.v[0].h, 0, 1, 0, nullptr), sh.v[0].h, "maddy_schedule_gtp");
                    plan_pending = false;
                    plan_events = 0;
                    hyd_dev = false;
                }
            }
            if (!classify) {
                deferred_output = host_events(); // flags cannot change a force here: the upload only keeps the device copy current
            } else {
                s.on_tubule_prev = s.on_tubule_cur;
                if (!counts_first) mt_len_prev = mt_len;
                if (undecided) { // classify this stride on the host
                    mt_length(s, step, mt_len);
                    if (par.barrier) {
                        for_each([&](Shard &d) { ck(maddy_upload_on_tubule(d.h, &s.on_tubule_cur[(size_t)d.first * N]), d.h, "maddy_upload_on_tubule"); });
                        st.h2d_bytes += (double)n;
                    }
                } else {
                    for_each([&](Shard &d) {
                        ck(maddy_snapshot_on_tubule(d.h, &s.on_tubule_cur[(size_t)d.first * N], &mt_len[d.first]), d.h, "maddy_snapshot_on_tubule");
                    });
                    if (!counts_first) mt_length_output(s, step, mt_len);
                }
                // the frame that was read back predates this stride's insertions; the reference writes the frame after them
                for (const Insertion &e : inserted) {
                    float *a0 = &s.r[e.q * 7], *a1 = &s.r[(e.q + 1) * 7];
                    a0[0] = a1[0] = e.x;
                    a0[1] = a1[1] = e.y;
                    a0[2] = e.z;
                    a1[2] = e.z2;
                }
                deferred_output = 1;
            }
            if (redo) { // the window of this stride, again, now in host mode (it ends at the next host event)
                window_end();
                for_each([&](Shard &d) { ck(maddy_run(d.h, step, count, explicit_rebuild ? MADDY_RUN_SKIP_FIRST_REBUILD : 0u), d.h, "maddy_run"); });
                mark("window launched again (host events)", step);
            }
        }
        prof.end("stride collect");
        prof.begin();
        if (deferred_output) {
            if (overlapped && deferred_output == 1 && !hyd_dev) {
                // The next hydrolysis event only needs the flags just classified; evaluate it and queue its window
                // BEFORE the stride's own output is formatted.  Messages are held back so stdout keeps the reference's
                // order, and the GTP state the output refers to is kept aside.
                pending_output = 1;
                pending_step = step;
                if (s.write_files && hp.hydrolysis && step % (hp.stride * 10) == 0) s.gtp_for_output = std::make_shared<std::vector<int>>(s.gtp);
            } else {
                update(s, step, mt_len);
                if (deferred_output == 2) mt_length(s, step, mt_len);
            }
        }
        prof.end("stride output");
        prof.begin();
        // overlap with the asynchronous window: evaluate the hydrolysis event AT the window end on the host now
        if (hydro && next < hp.steps && next % hp.hydrostep == 0 && next != 0 && !(hyd_dev && plan_covers(next))) {
            if (pending_output) s.event_log = &pending_log;
            hydrolyse(s);
            hydrolysed_for = next;
            // further events of the next window: only inside a stride (an event AT a stride step precedes that stride's
            // energies and is uploaded on its own); the window is one event period longer than this one, in whole periods
            // (not beside the window that follows a stride step: the host is busy with that stride's read-back there)
            if (multi_event && next % hp.stride != 0 && step % hp.stride != 0) {
                const long long h = hp.hydrostep;
                const long long room = std::min((next / hp.stride + 1) * hp.stride, hp.steps) - next; // > 0
                const long long n_ev = std::min(count / h + 1, (room + h - 1) / h);
                sched_gtp.clear();
                if (n_ev > 1) {
                    sched_gtp.push_back(s.gtp);
                    for (long long k = 1; k < n_ev; k++) {
                        hydrolyse(s);
                        hydrolysed_for = next + k * h;
                        sched_gtp.push_back(s.gtp);
                    }
                    sched_first = next;
                    sched_end = next + std::min(n_ev * h, room);
                }
            }
            s.event_log = nullptr;
            mark("next hydrolysis evaluated", step);
        }
        prof.end("hydrolyse (host)");
        step = next;
        st.steps += count;
    }
    if (pending_output) flush_pending();
    if (hyd_dev) { // the host copies catch up with the device: generator past the last events, their messages, final GTP flags
        collect_plan();
        download_gtp();
    }
    if (!hp.checkpoint.empty() && hp.steps > start_step) write_checkpoint(hp.steps);
    prof.begin();
    if (s.writer) {
        s.writer->drain();
        s.writer_files.reset();
    }
    prof.end("drain writer");
    prof.begin();
    for_each([&](Shard &d) {
        ck(maddy_sync(d.h), d.h, "maddy_sync");
        st.launches += maddy_launch_count(d.h);
    });
    prof.end("final sync");
    if (stats) *stats = st;
}

} // namespace mt
