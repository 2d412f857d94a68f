/*
 * system.cpp — model preparation of the drop-in host: config -> Parameters, PDB -> topology.
 *
 * Behaviour follows the reference's initParameters / read_PDB / AssemblyInit
 * (src/preparator.cpp:4-246, :249-561, :717-731) because these define the inputs the kernels
 * see: key names and defaults, the order config -> forcefield -> conditions (each parse
 * REPLACES the table), derived constants (gamma, var, barrier widths, hydrostep) evaluated with
 * the reference's float/double mix, and the topology rules.
 */
#include <cerrno>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sys/resource.h>
#include <thread>
#include "mt_host.hpp"

namespace mt {

static const float R_MT = 8.12f, R_MON = 2.0f, KB = 0.0019872041f;

maddy_topology System::topology_view(int traj_first) const
{
    maddy_topology t{};
    const size_t o = (size_t)traj_first * par.n_tot;
    t.harmonic_count = harmonic_count.data();
    t.harmonic = harmonic.data();
    t.longitudinal_count = longitudinal_count.data() + o;
    t.longitudinal = longitudinal.empty() ? nullptr : longitudinal.data() + o * par.max_longitudinal;
    t.lateral_count = lateral_count.data() + o;
    t.lateral = lateral.empty() ? nullptr : lateral.data() + o * par.max_lateral;
    t.fixed = fixed.data();
    t.extra = extra.data() + o;
    t.mon_type = mon_type.data();
    t.gtp = gtp.data() + o;
    t.on_tubule_cur = on_tubule_cur.data() + o;
    return t;
}

// ---- orientation helpers for the static lateral list (host libm, as the reference host)
struct V3 { float x, y, z; };
static V3 site_offset(float fi, float psi, float theta, float xp, float yp, float zp)
{
    const float sf = sinf(fi), cf = cosf(fi), sp = sinf(psi), cp = cosf(psi), st = sinf(theta), ct = cosf(theta);
    V3 o;
    o.x = xp * cp * ct + yp * (cp * sf * st - cf * sp) + zp * (sf * sp + cf * cp * st);
    o.y = xp * ct * sp + yp * (cf * cp + sf * sp * st) + zp * (cf * sp * st - cp * sf);
    o.z = -xp * st + yp * ct * sf + zp * cf * ct;
    return o;
}

static void read_structure(System &s)
{
    maddy_params &par = s.par;
    read_pdb(s.hp.coord_xyz, s.pdb, s.quiet);
    if (!s.quiet) printf("Building topology....\n");
    const int N = par.n_tot = (int)s.pdb.atoms.size();
    const int Ntr = par.n_tr;
    if (N == 0) die("no ATOM records in '%s'", s.hp.coord_xyz.c_str());
    const size_t n = (size_t)N * Ntr;
    s.r.assign(n * 7, 0.f);
    s.f.assign(n * 7, 0.f);

    s.mon_type.assign(N, 0);
    for (int i = 0; i < N; i++) {
        if (s.pdb.atoms[i].name[1] == 'A') s.mon_type[i] = 0;
        else if (s.pdb.atoms[i].name[1] == 'B') s.mon_type[i] = 1;
    }
    s.gtp.assign(n, 0);
    s.on_tubule_cur.assign(n, 0);
    s.on_tubule_prev.assign(n, 0);
    s.fixed.assign(N, 0);
    for (int i = 0; i < N; i++) s.fixed[i] = s.pdb.atoms[i].resid <= s.hp.fix;
    s.extra.assign(n, 0);
    for (int t = 0; t < Ntr; t++)
        for (int i = 0; i < N; i++) s.extra[(size_t)t * N + i] = s.pdb.atoms[i].chain == 'X';

    read_pdb(s.hp.coord_ang, s.pdb_ang, s.quiet);
    if ((int)s.pdb_ang.atoms.size() < N) die("'%s' has fewer atoms than '%s'", s.hp.coord_ang.c_str(), s.hp.coord_xyz.c_str());
    for (int t = 0; t < Ntr; t++)
        for (int i = 0; i < N; i++) {
            float *c = &s.r[((size_t)t * N + i) * 7];
            c[0] = (float)s.pdb.atoms[i].x;
            c[1] = (float)s.pdb.atoms[i].y;
            c[2] = (float)s.pdb.atoms[i].z;
            c[3] = (float)s.pdb_ang.atoms[i].x; // fi
            c[5] = (float)s.pdb_ang.atoms[i].y; // psi
            c[4] = (float)s.pdb_ang.atoms[i].z; // theta
        }

    // intra-dimer ("harmonic") bonds: same residue and chain, adjacent records; the partner
    // with the smaller serial is stored as -j  (preparator.cpp:326-355)
    const auto &at = s.pdb.atoms;
    auto bonded = [&](int i, int j) {
        return at[i].resid == at[j].resid && at[i].chain == at[j].chain && i != j && abs(i - j) < 2;
    };
    s.harmonic_count.assign(N, 0);
    int maxH = 0;
    for (int i = 0; i < N; i++) {
        int c = 0;
        for (int j = (i > 0 ? i - 1 : 0); j <= i + 1 && j < N; j++)
            if (bonded(i, j)) c++;
        if (c > maxH) maxH = c;
    }
    par.max_harmonic = maxH > 0 ? maxH : 1; // keep one (unused, -1) slot so entry [i*maxH] exists
    s.harmonic.assign((size_t)par.max_harmonic * N, -1);
    for (int i = 0; i < N; i++)
        for (int j = (i > 0 ? i - 1 : 0); j <= i + 1 && j < N; j++)
            if (bonded(i, j)) {
                s.harmonic[(size_t)par.max_harmonic * i + s.harmonic_count[i]] = at[i].id > at[j].id ? j : -j;
                s.harmonic_count[i]++;
            }

    s.longitudinal_count.assign(n, 0);
    s.lateral_count.assign(n, 0);
    s.longitudinal.clear();
    s.lateral.clear();
    par.max_longitudinal = 0;
    par.max_lateral = 0;

    if (!par.is_assembly) {
        // static longitudinal list: neighbours along the chain in adjacent residues (preparator.cpp:357-407)
        std::vector<std::vector<int>> lng(N), lat(N);
        for (int i = 0; i < N; i++) {
            for (int j = 0; j < N; j++)
                if (abs(at[i].resid - at[j].resid) == 1 && at[i].chain == at[j].chain && abs(at[i].id - at[j].id) == 1)
                    lng[i].push_back(at[i].id > at[j].id ? j : -j);
        }
        // static lateral list: site distance < r_mon on the initial structure, -j for the
        // (i:p1, j:p2) pairing and +j for (i:p2, j:p1); +-0 is NOT replaced by the ZERO sentinel
        // here (preparator.cpp:409-561)
        const float a13 = (float)(2.0f * M_PI / 13.0f);
        const float xp = 0.5f * R_MT * (cosf(a13) - 1.0f), yp = 0.5f * R_MT * sinf(a13), zp = -3.0f * R_MON / 13.0f;
        std::vector<V3> o1(N), o2(N);
        for (int i = 0; i < N; i++) {
            const float *c = &s.r[(size_t)i * 7];
            o1[i] = site_offset(c[3], c[5], c[4], xp, yp, zp);
            o2[i] = site_offset(c[3], c[5], c[4], xp, -yp, -zp);
        }
        auto dist = [&](int i, const V3 &oi, int j, const V3 &oj) {
            const float *ci = &s.r[(size_t)i * 7], *cj = &s.r[(size_t)j * 7];
            double dx = (double)(ci[0] - cj[0] + oi.x - oj.x), dy = (double)(ci[1] - cj[1] + oi.y - oj.y),
                   dz = (double)(ci[2] - cj[2] + oi.z - oj.z);
            return sqrtf((float)(dz * dz + dx * dx + dy * dy));
        };
        for (int i = 0; i < N; i++)
            for (int j = 0; j < N; j++) {
                if (dist(i, o1[i], j, o2[j]) < R_MON) lat[i].push_back(-j);
                if (dist(i, o2[i], j, o1[j]) < R_MON) lat[i].push_back(j);
            }
        for (int t = 0; t < Ntr; t++)
            for (int i = 0; i < N; i++) {
                if (s.extra[(size_t)t * N + i]) continue;
                par.max_longitudinal = std::max(par.max_longitudinal, (int)lng[i].size());
                par.max_lateral = std::max(par.max_lateral, (int)lat[i].size());
            }
        s.longitudinal.assign((size_t)par.max_longitudinal * n, 0);
        s.lateral.assign((size_t)par.max_lateral * n, 0);
        for (int t = 0; t < Ntr; t++)
            for (int i = 0; i < N; i++) {
                const size_t q = (size_t)t * N + i;
                if (s.extra[q]) continue;
                s.longitudinal_count[q] = (int)lng[i].size();
                for (size_t k = 0; k < lng[i].size(); k++) s.longitudinal[q * par.max_longitudinal + k] = lng[i][k];
                s.lateral_count[q] = (int)lat[i].size();
                for (size_t k = 0; k < lat[i].size(); k++) s.lateral[q * par.max_lateral + k] = lat[i][k];
            }
    }
    if (!s.quiet) printf("done building topology without LJ.\n");
}

void assembly_init(System &s)
{
    maddy_params &par = s.par;
    par.max_lateral = 16;
    par.max_longitudinal = 8;
    const size_t n = (size_t)par.n_tot * par.n_tr;
    s.lateral.assign((size_t)par.max_lateral * n, 0);
    s.longitudinal.assign((size_t)par.max_longitudinal * n, 0);
    s.longitudinal_count.assign(n, 0);
    s.lateral_count.assign(n, 0);
}

void init_parameters(System &s, const std::string &config, const std::vector<std::string> &overrides)
{
    maddy_params &par = s.par;
    HostParams &hp = s.hp;
    ParamTable &tb = s.table;
    tb.quiet = s.quiet;
    s.overrides = overrides;
    memset(&par, 0, sizeof par);
    par.abi_version = MADDY_ABI_VERSION;

    tb.parse(config, overrides);
    par.device = tb.integer("device");
    par.dt = tb.real("dt");
    par.rseed = tb.integer("rseed");
    hp.steps = tb.long_integer("steps", -1);
    hp.stride = tb.long_integer("stride", -1);
    par.ljpairscutoff = tb.real("LJPairsCutoff");
    par.ljpairsupdatefreq = tb.integer("LJPairsUpdateFreq");
    hp.coord_ang = tb.masked("coordinates_ang");
    hp.coord_xyz = tb.masked("coordinates_xyz");
    hp.ff_file = tb.masked("forcefield");
    hp.cond_file = tb.masked("conditions");
    hp.fix = tb.integer("fix", 1);
    par.n_tr = tb.integer("runnum", 1);
    par.is_assembly = tb.yesno("is_assembly", 1);
    hp.tub_length = tb.yesno("tubule_length", 1);
    hp.out_energy = tb.yesno("output_energy", 1);
    hp.out_force = tb.yesno("output_force", 0);
    par.tea_on = tb.yesno("tea_on", 0);
    if (par.tea_on) {
        par.tea_a = tb.real("tea_a", R_MON);
        par.tea_capricious = tb.yesno("tea_capricious", 1);
        par.tea_epsilon_freq = tb.integer("tea_epsilon_freq"); // mandatory (allowDefault = 0)
        par.tea_epsmax = tb.real("tea_epsmax", 999.f);
    }
    hp.n_gpus = tb.integer("n_gpus", 1); // extension of this host, absent in the reference
    hp.restartkey = tb.masked("restartkey");
    hp.is_restart = tb.yesno("is_restart", 0);
    // extension keys, absent in the reference: exact checkpoints (checkpoint.cpp)
    hp.checkpoint = tb.masked("checkpoint", "");
    hp.checkpoint_freq = tb.long_integer("checkpoint_freq", 0);
    hp.insitu = tb.yesno("insitu_analysis", 0);
    long long resume_step = 0;
    hp.resume = hp.is_restart && !hp.checkpoint.empty() && checkpoint_peek(hp.checkpoint, &resume_step);
    // every modulus of the step loop, checked before anything divides by it (the reference would die with SIGFPE)
    if (hp.stride <= 0) die("stride must be positive");
    if (par.ljpairsupdatefreq <= 0) die("LJPairsUpdateFreq must be positive");
    if (!hp.checkpoint.empty()) {
        if (hp.steps % par.ljpairsupdatefreq != 0)
            die("checkpoint: steps (%lld) must be a multiple of LJPairsUpdateFreq (%d)", hp.steps, par.ljpairsupdatefreq);
        if (hp.checkpoint_freq < 0 || hp.checkpoint_freq % par.ljpairsupdatefreq != 0 || hp.checkpoint_freq % hp.stride != 0)
            die("checkpoint_freq (%lld) must be a multiple of stride and of LJPairsUpdateFreq", hp.checkpoint_freq);
    }
    if (!hp.is_restart || hp.resume) hp.firststep = hp.resume ? resume_step : tb.long_integer("firststep", 0);
    else {
        FILE *k = fopen(hp.restartkey.c_str(), "r");
        if (!k) die("Opening file '%s'", hp.restartkey.c_str());
        if (fscanf(k, "%lld", &hp.firststep) != 1) die("Reading restartkey %s: unable to get firststep", hp.restartkey.c_str());
        fclose(k);
    }
    if (par.n_tr <= 0) die("runnum must be positive");
    if (hp.stride <= 0) die("stride must be positive");
    par.traj_first = 0;
    par.n_tr_local = par.n_tr;

    read_structure(s);

    // per-trajectory output names and the DCD headers (preparator.cpp:89-115)
    hp.dcd_xyz.resize(par.n_tr);
    hp.dcd_ang.resize(par.n_tr);
    hp.restart_xyz.resize(par.n_tr);
    hp.restart_ang.resize(par.n_tr);
    s.dcd = make_dcd_header(par.n_tot, (int)(hp.steps / hp.stride), 1, par.dt, (int)hp.stride);
    for (int t = 0; t < par.n_tr; t++) {
        const std::string run = std::to_string(t);
        hp.dcd_ang[t] = tb.masked_replace("dcd_ang", run, "<run>");
        hp.dcd_xyz[t] = tb.masked_replace("dcd_xyz", run, "<run>");
        if (s.write_files && !hp.resume) {
            for (const std::string *name : {&hp.dcd_xyz[t], &hp.dcd_ang[t]}) {
                FILE *f = fopen(name->c_str(), "w");
                if (!f) die("Opening file '%s'", name->c_str());
                dcd_write_header(f, s.dcd);
                fclose(f);
            }
        }
        hp.restart_xyz[t] = tb.masked_replace("restart_xyz", run, "<run>");
        hp.restart_ang[t] = tb.masked_replace("restart_ang", run, "<run>");
    }
    if (hp.is_restart && !hp.resume) read_restart(s);
    if (hp.resume && s.write_files) checkpoint_trim_outputs(s, resume_step);

    // ---- force field
    tb.parse(hp.ff_file, overrides);
    par.A_long = tb.real("A_long");
    par.A_lat = tb.real("A_lat");
    par.D_long = tb.real("D_long");
    par.D_lat = tb.real("D_lat");
    par.seam_coeff = tb.real("seam_coeff");
    if (tb.yesno("barrier", 1, false)) {
        par.barrier = 1;
        par.a_barr_long = tb.real("a_barr_long");
        par.r_barr_long = tb.real("r_barr_long");
        par.w_barr_long = (float)(tb.real("w_barr_long") / (2 * (sqrt(-2 * log(0.5))))); // FWHM -> sigma
        par.a_barr_lat = tb.real("a_barr_lat");
        par.r_barr_lat = tb.real("r_barr_lat");
        par.w_barr_lat = (float)(tb.real("w_barr_lat") / (2 * (sqrt(-2 * log(0.5)))));
    } else par.barrier = 0;
    par.C = tb.real("C");
    par.B_fi = tb.real("B_fi");
    par.B_psi = tb.real("B_psi");
    par.B_theta = tb.real("B_theta");
    par.fi_0 = tb.real("fi0");
    par.psi_0 = tb.real("psi0");
    par.theta0_gdp = tb.real("theta0_gdp");
    par.theta0_gtp = tb.real("theta0_gtp");
    if (tb.yesno("LJ_on", 1, false)) {
        par.lj_on = 1;
        par.ljscale = tb.real("LJScale");
        par.ljsigma6 = (float)pow((double)tb.real("LJSigma"), 6);
    } else par.lj_on = 0;
    par.is_wall = tb.yesno("repulsive_walls", 1, false);
    par.rep_leftborder = tb.real("rep_leftborder");
    par.rep_h = tb.real("rep_h");
    par.rep_r = tb.real("rep_r");
    par.rep_eps = tb.real("rep_eps");

    // ---- conditions
    tb.parse(hp.cond_file, overrides);
    if (tb.yesno("is_const_conc", 1, false)) {
        hp.is_const_conc = true;
        hp.conc = tb.real("conc");
    } else hp.is_const_conc = false;
    if (tb.yesno("hydrolysis", 1, false)) {
        hp.hydrolysis = true;
        hp.khydro = tb.real("khydro");
        hp.hydrostep = (long)(0.02 * 1000000000000 / (par.dt * hp.khydro)); // 2 % probability threshold
        if (hp.hydrostep <= 0) die("hydrolysis: dt * khydro = %g gives a hydrolysis period of %lld steps (must be >= 1)", (double)(par.dt * hp.khydro), (long long)hp.hydrostep);
    } else hp.hydrolysis = false;
    for (size_t q = 0; q < s.gtp.size(); q++) s.gtp[q] = 1;

    par.Temp = tb.real("Temp");
    hp.viscosity = tb.real("viscosity");
    if (par.tea_on) par.gammaR = (float)(6 * M_PI * hp.viscosity * par.tea_a);
    else par.gammaR = (float)(6 * M_PI * hp.viscosity * R_MON);
    par.gammaTheta = (float)(8 * M_PI * hp.viscosity * pow((double)R_MON, 3));
    par.varR = sqrtf(2.0f * KB * par.Temp * par.dt / par.gammaR);
    par.varTheta = sqrtf(2.0f * KB * par.Temp * par.dt / par.gammaTheta);
    par.alpha = tb.real("alpha");
    par.freeze_temp = tb.real("freeze_temp");

    int extra_counter = 0;
    for (int i = 0; i < par.n_tot; i++)
        if (s.extra[i]) extra_counter++;
    if (extra_counter == 0 && hp.is_const_conc)
        die("Error! You want constant concentration! Load structure with extra particles (chain X) to support constant concentration!");

    if (hp.hydrolysis) {
        s.coordspdb.atoms.resize((size_t)par.n_tot * par.n_tr);
        for (int t = 0; t < par.n_tr; t++)
            for (int i = 0; i < par.n_tot; i++) s.coordspdb.atoms[(size_t)t * par.n_tot + i] = s.pdb.atoms[i];
    }
    if (par.n_tot > MADDY_MAX_NTOT) die("n_tot = %d exceeds the %d monomers per trajectory this build supports", par.n_tot, MADDY_MAX_NTOT);
}

// ------------------------------------------------------------------ coordinate output
void save_coord_pdb(System &s, const std::string &xyz, const std::string &ang)
{
    // trajectory 0 only, from the host copy of the coordinates (preparator.cpp:566-580)
    PDB out = s.pdb;
    for (int i = 0; i < s.par.n_tot; i++) {
        const float *c = &s.r[(size_t)i * 7];
        out.atoms[i].x = c[0];
        out.atoms[i].y = c[1];
        out.atoms[i].z = c[2];
    }
    write_pdb(xyz, out, s.quiet);
    for (int i = 0; i < s.par.n_tot; i++) {
        const float *c = &s.r[(size_t)i * 7];
        out.atoms[i].x = c[3];
        out.atoms[i].y = c[5];
        out.atoms[i].z = c[4];
    }
    write_pdb(ang, out, s.quiet);
}

// ------------------------------------------------------------------ background output
OutputWorker::OutputWorker() : thread_([this] { loop(); }) {}
OutputWorker::~OutputWorker()
{
    {
        std::lock_guard<std::mutex> g(m_);
        stop_ = true;
    }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
}
void OutputWorker::loop()
{
    for (;;) {
        std::function<void()> job;
        {
            std::unique_lock<std::mutex> g(m_);
            cv_.wait(g, [this] { return stop_ || !q_.empty(); });
            if (q_.empty()) return;
            job = std::move(q_.front());
            q_.pop_front();
            busy_ = true;
        }
        try {
            job();
        } catch (const std::exception &e) {
            std::lock_guard<std::mutex> g(m_);
            if (error_.empty()) error_ = e.what();
        }
        {
            std::lock_guard<std::mutex> g(m_);
            busy_ = false;
        }
        cv_.notify_all();
    }
}
void OutputWorker::submit(std::function<void()> job)
{
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return q_.size() < 16; });
    q_.push_back(std::move(job));
    cv_.notify_all();
}
void OutputWorker::drain()
{
    std::unique_lock<std::mutex> g(m_);
    cv_.wait(g, [this] { return q_.empty() && !busy_; });
    if (!error_.empty()) {
        std::string e = error_;
        error_.clear();
        throw Fatal(e);
    }
}

// Append-mode file handles kept open across strides (the reference re-opens and closes 2*Ntr files per stride:
// at 256 trajectories that is ~15 ms of syscalls per stride).  Falls back to open/close when descriptors run out.
class AppendFiles {
  public:
    ~AppendFiles() { close_all(); }
    void append(const std::string &name, const void *data, size_t bytes)
    {
        FILE *f = nullptr;
        auto it = open_.find(name);
        if (it != open_.end()) f = it->second;
        else {
            f = fopen(name.c_str(), "a");
            if (!f && (errno == EMFILE || errno == ENFILE)) {
                close_all();
                f = fopen(name.c_str(), "a");
            }
            if (!f) die("Opening file '%s'", name.c_str());
            if (open_.size() < kMaxOpen) open_[name] = f;
            else transient_ = f;
        }
        const bool ok = fwrite(data, 1, bytes, f) == bytes && fflush(f) == 0;
        if (transient_) {
            fclose(transient_);
            transient_ = nullptr;
        }
        if (!ok) die("Writing file '%s'", name.c_str());
    }
    void close_all()
    {
        for (auto &kv : open_) fclose(kv.second);
        open_.clear();
    }

    // Handles kept open: all the process may have (the soft descriptor limit is raised to the hard one once), minus a
    // reserve for everything else.  2 files per trajectory: an ensemble of 2048 needs 4096.
    static size_t max_open()
    {
        static const size_t k = [] {
            struct rlimit rl;
            if (getrlimit(RLIMIT_NOFILE, &rl) != 0) return (size_t)768;
            if (rl.rlim_cur < rl.rlim_max) {
                struct rlimit want = rl;
                want.rlim_cur = rl.rlim_max == RLIM_INFINITY ? 65536 : (rl.rlim_max > 65536 ? 65536 : rl.rlim_max);
                if (want.rlim_cur > rl.rlim_cur && setrlimit(RLIMIT_NOFILE, &want) == 0) rl = want;
            }
            const size_t cur = rl.rlim_cur == RLIM_INFINITY ? 65536 : (size_t)rl.rlim_cur;
            return cur > 512 ? cur - 256 : cur / 2;
        }();
        return k;
    }
    size_t kMaxOpen = 768; // per partition (see FramePartitions)

  private:
    std::map<std::string, FILE *> open_;
    FILE *transient_ = nullptr;
};
// The per-trajectory files of an ensemble, partitioned over a few writer threads (trajectory t belongs to partition
// t % parts): one host driving eight GPUs appends 4096 frames per stride, which one thread does not keep up with.
struct FramePartitions {
    std::vector<AppendFiles> part;
    explicit FramePartitions(int n) : part(n)
    {
        for (AppendFiles &f : part) f.kMaxOpen = AppendFiles::max_open() / (size_t)n;
    }
};

// one frame per trajectory appended to <dcd_xyz> (x,y,z) and <dcd_ang> (fi,psi,theta); each frame is assembled in
// memory and written with a single fwrite (the reference does open / 9 small fwrites / close per file)
static void write_dcd_frames_part(const std::vector<float> &r, int N, int Ntr, const std::vector<std::string> &xyz,
                                  const std::vector<std::string> &ang, AppendFiles *files, int first, int step)
{
    const size_t frame_bytes = 3 * ((size_t)N * 4 + 8);
    std::vector<char> buf(frame_bytes);
    const int len = N * 4;
    for (int t = first; t < Ntr; t += step) {
        for (int pass = 0; pass < 2; pass++) {
            static const int col[2][3] = {{0, 1, 2}, {3, 5, 4}};
            char *o = buf.data();
            for (int k = 0; k < 3; k++) {
                memcpy(o, &len, 4);
                o += 4;
                float *dst = reinterpret_cast<float *>(o);
                const float *src = &r[(size_t)t * N * 7 + col[pass][k]];
                for (int i = 0; i < N; i++) dst[i] = src[(size_t)i * 7];
                o += (size_t)N * 4;
                memcpy(o, &len, 4);
                o += 4;
            }
            files->append(pass ? ang[t] : xyz[t], buf.data(), frame_bytes);
        }
    }
}
static void write_dcd_frames(const std::vector<float> &r, int N, int Ntr, const std::vector<std::string> &xyz,
                             const std::vector<std::string> &ang, FramePartitions *files = nullptr)
{
    if (!files) {
        AppendFiles local;
        write_dcd_frames_part(r, N, Ntr, xyz, ang, &local, 0, 1);
        return;
    }
    const int P = (int)files->part.size();
    if (P == 1) {
        write_dcd_frames_part(r, N, Ntr, xyz, ang, &files->part[0], 0, 1);
        return;
    }
    std::vector<std::thread> th;
    std::vector<std::string> err(P);
    for (int p = 0; p < P; p++)
        th.emplace_back([&, p] {
            try {
                write_dcd_frames_part(r, N, Ntr, xyz, ang, &files->part[p], p, P);
            } catch (const std::exception &e) {
                err[p] = e.what();
            }
        });
    for (std::thread &t : th) t.join();
    for (const std::string &e : err)
        if (!e.empty()) die("%s", e.c_str());
}

void save_coord_dcd(System &s)
{
    if (s.writer) {
        // snapshot for the writer thread: buffers are recycled (a fresh 30 MB vector per stride is mostly page faults at
        // 2048 trajectories) and filled by a few threads
        static thread_local std::vector<std::shared_ptr<std::vector<float>>> pool;
        std::shared_ptr<std::vector<float>> snap;
        for (auto &b : pool)
            if (b.use_count() == 1 && b->size() == s.r.size()) {
                snap = b;
                break;
            }
        if (!snap) {
            snap = std::make_shared<std::vector<float>>(s.r.size());
            if (pool.size() >= 4) pool.erase(pool.begin());
            pool.push_back(snap);
        }
        {
            const size_t bytes = s.r.size() * sizeof(float), chunk = (size_t)1 << 20, nchunk = (bytes + chunk - 1) / chunk;
            const char *src = reinterpret_cast<const char *>(s.r.data());
            char *dst = reinterpret_cast<char *>(snap->data());
            int nt = (int)(std::thread::hardware_concurrency() / 4);
            nt = nt < 1 ? 1 : (nt > 8 ? 8 : nt);
#pragma omp parallel for schedule(static) num_threads(nt) if (bytes > ((size_t)8 << 20))
            for (long long c = 0; c < (long long)nchunk; c++) {
                const size_t o = (size_t)c * chunk;
                memcpy(dst + o, src + o, bytes - o < chunk ? bytes - o : chunk);
            }
        }
        const int N = s.par.n_tot, Ntr = s.par.n_tr;
        const HostParams *hp = &s.hp;
        auto files = std::static_pointer_cast<FramePartitions>(s.writer_files);
        if (!files) {
            // writer threads: one per ~512 trajectories, at most 8 (and never more than half the hardware threads)
            int parts = (Ntr + 511) / 512;
            const int hw = (int)std::thread::hardware_concurrency() / 2;
            parts = parts > 8 ? 8 : parts;
            parts = parts > hw ? (hw < 1 ? 1 : hw) : parts;
            files = std::make_shared<FramePartitions>(parts);
            s.writer_files = files;
        }
        s.writer->submit([snap, N, Ntr, hp, files] { write_dcd_frames(*snap, N, Ntr, hp->dcd_xyz, hp->dcd_ang, files.get()); });
    } else {
        write_dcd_frames(s.r, s.par.n_tot, s.par.n_tr, s.hp.dcd_xyz, s.hp.dcd_ang);
    }
}

// dcd/hydrolysis.pdb: every monomer of every trajectory with occupancy = GTP flag, beta = trajectory
// (preparator.cpp:604-636); formatted from snapshots so that it can run on the output thread
static void write_hydrolysis_pdb(const PDB &tmpl, const std::vector<float> &r, const std::vector<int> &gtp, int N, int Ntr, bool quiet)
{
    PDB out;
    out.atoms.resize((size_t)N * Ntr);
    for (int t = 0; t < Ntr; t++)
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            PDBAtom a = tmpl.atoms[i];
            a.x = r[q * 7 + 0];
            a.y = r[q * 7 + 1];
            a.z = r[q * 7 + 2];
            a.id = (int)q;
            a.beta = (double)t;
            a.occupancy = (double)gtp[q];
            out.atoms[q] = a;
        }
    append_pdb("dcd/hydrolysis.pdb", out, quiet);
}

void append_coord_pdb(System &s)
{
    const int N = s.par.n_tot, Ntr = s.par.n_tr;
    if (s.writer) {
        auto r = std::make_shared<std::vector<float>>(s.r);
        auto g = s.gtp_for_output ? s.gtp_for_output : std::make_shared<std::vector<int>>(s.gtp);
        const PDB *tmpl = &s.pdb;
        const bool quiet = s.quiet;
        s.writer->submit([r, g, tmpl, N, Ntr, quiet] { write_hydrolysis_pdb(*tmpl, *r, *g, N, Ntr, quiet); });
    } else {
        write_hydrolysis_pdb(s.pdb, s.r, s.gtp_for_output ? *s.gtp_for_output : s.gtp, N, Ntr, s.quiet);
    }
}

void read_restart(System &s)
{
    const int N = s.par.n_tot;
    std::vector<XYZAtom> atoms;
    for (int t = 0; t < s.par.n_tr; t++) {
        read_xyz(s.hp.restart_xyz[t], atoms, s.quiet);
        if ((int)atoms.size() < N) die("restart file '%s' has too few atoms", s.hp.restart_xyz[t].c_str());
        for (int i = 0; i < N; i++) {
            float *c = &s.r[((size_t)t * N + i) * 7];
            c[0] = (float)atoms[i].x;
            c[1] = (float)atoms[i].y;
            c[2] = (float)atoms[i].z;
        }
        read_xyz(s.hp.restart_ang[t], atoms, s.quiet);
        if ((int)atoms.size() < N) die("restart file '%s' has too few atoms", s.hp.restart_ang[t].c_str());
        for (int i = 0; i < N; i++) {
            float *c = &s.r[((size_t)t * N + i) * 7];
            c[3] = (float)atoms[i].x;
            c[5] = (float)atoms[i].y;
            c[4] = (float)atoms[i].z;
        }
    }
}

void write_restart(System &s, long long step)
{
    const int N = s.par.n_tot;
    std::vector<XYZAtom> atoms(N);
    for (int t = 0; t < s.par.n_tr; t++) {
        for (int pass = 0; pass < 2; pass++) {
            for (int i = 0; i < N; i++) {
                const float *c = &s.r[((size_t)t * N + i) * 7];
                atoms[i].x = pass ? c[3] : c[0];
                atoms[i].y = pass ? c[5] : c[1];
                atoms[i].z = pass ? c[4] : c[2];
                atoms[i].name = s.pdb.atoms[i].name[0];
            }
            write_xyz(pass ? s.hp.restart_ang[t] : s.hp.restart_xyz[t], atoms, s.quiet);
        }
    }
    FILE *k = fopen(s.hp.restartkey.c_str(), "w");
    if (!k) die("Opening file '%s'", s.hp.restartkey.c_str());
    fprintf(k, "%lld", step);
    fclose(k);
}

} // namespace mt
