/*
 * mt_host.hpp — drop-in compatible C++ host of MADDY around the B200 C-ABI (internal API).
 *
 * The reference host (src/main.cpp, preparator.cpp, updater.cpp, configreader.cpp, pdbio.cpp,
 * dcdio.cpp, xyzio.cpp) is a set of free functions over file-scope globals.  This host keeps
 * the same observable behaviour — config.conf / forcefield / conditions keys, PDB/XYZ inputs,
 * DCD + stdout outputs, the order of host events inside the step loop — but is re-entrant:
 * all state lives in `mt::System`, and fatal conditions throw `mt::Fatal` (the `mt` executable
 * turns that into the reference's "Fatal error!" + exit(-1), the Python binding into a status).
 */
#pragma once
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <stdexcept>
#include <string>
#include <cstddef>
#include <vector>
#include "maddy_b200.h"

namespace mt {

struct Fatal : std::runtime_error {
    using std::runtime_error::runtime_error;
};
[[noreturn]] void die(const char *fmt, ...);

// ---------------------------------------------------------------- config (configreader.cpp)
class ParamTable {
  public:
    // replaces the table (configreader.cpp:31) and re-applies name=value overrides (:62-75)
    void parse(const std::string &filename, const std::vector<std::string> &overrides);
    void set(const std::string &name, const std::string &value, bool force);
    bool raw(const std::string &name, std::string &value) const;

    std::string masked(const std::string &name);                               // mandatory
    std::string masked(const std::string &name, const std::string &def);       // with default
    std::string masked_replace(const std::string &name, const std::string &replacement, const std::string &what);
    int integer(const std::string &name);
    int integer(const std::string &name, int def);
    long long long_integer(const std::string &name, long def);
    float real(const std::string &name);
    float real(const std::string &name, float def);
    int yesno(const std::string &name, int def, bool allow_default = true);
    bool quiet = false;

  private:
    bool lookup(const std::string &name, const std::string *def, std::string &out);
    std::string apply_mask(const std::string &token) const;
    std::vector<std::pair<std::string, std::string>> items_;
};

// ---------------------------------------------------------------- PDB (pdbio.cpp)
struct PDBAtom {
    int id = 0;
    char name[5] = {0}, chain = ' ', resName[4] = {0}, altLoc = ' ';
    int resid = 0;
    double x = 0, y = 0, z = 0, occupancy = 0, beta = 0;
};
struct PDB {
    std::vector<PDBAtom> atoms;
};
void read_pdb(const std::string &filename, PDB &pdb, bool quiet = false);
void write_pdb(const std::string &filename, const PDB &pdb, bool quiet = false);
void append_pdb(const std::string &filename, const PDB &pdb, bool quiet = false);

// ---------------------------------------------------------------- DCD (dcdio.cpp)
struct DCDHeader {
    int n_atoms = 0, nfile = 0, npriv = 0, nsavc = 0;
    float delta = 0;
    std::string remark1, remark2;
};
DCDHeader make_dcd_header(int n_atoms, int frame_count, int first_frame, float timestep, int dcd_freq);
void dcd_write_header(FILE *f, const DCDHeader &h);
void dcd_write_frame(FILE *f, int n, const float *x, const float *y, const float *z);
bool dcd_read_header(FILE *f, DCDHeader &h);
bool dcd_read_frame(FILE *f, int n, float *x, float *y, float *z);

// ---------------------------------------------------------------- XYZ (xyzio.cpp)
struct XYZAtom {
    char name;
    double x, y, z;
};
void read_xyz(const std::string &filename, std::vector<XYZAtom> &atoms, bool quiet = false);
void write_xyz(const std::string &filename, const std::vector<XYZAtom> &atoms, bool quiet = false);

// ---------------------------------------------------------------- system (preparator.cpp, globals)
struct HostParams { // scalars of `Parameters` that only the host uses
    long long steps = 0, firststep = 0, stride = 0;
    int fix = 1;
    bool tub_length = true, out_energy = true, out_force = false, is_restart = false;
    bool is_const_conc = false, hydrolysis = false;
    float conc = 0, khydro = 0, viscosity = 0;
    long hydrostep = 0;
    std::string coord_xyz, coord_ang, ff_file, cond_file, restartkey;
    std::vector<std::string> dcd_xyz, dcd_ang, restart_xyz, restart_ang;
    int n_gpus = 1; // extension key `n_gpus` (default 1): shard trajectories over this many devices
    // extension keys `checkpoint <file>` / `checkpoint_freq <steps>`: exact checkpoints (checkpoint.cpp); with
    // `is_restart yes` and an existing checkpoint file the run resumes from it instead of the XYZ restart files
    std::string checkpoint;
    long long checkpoint_freq = 0;
    bool resume = false;
    // extension key `insitu_analysis yes`: the reference's offline DCD tools (scripts/temp_calc, scripts/disas_speed)
    // evaluated on the device at every stride, one line per frame into <dcd_xyz>.temp.dat / <dcd_xyz>.disc.dat
    bool insitu = false;
};

// In-order background writer for trajectory output: the step loop hands over a snapshot of
// the host coordinates and keeps the GPU busy while frames are formatted and appended to the per-trajectory files.
class OutputWorker {
  public:
    OutputWorker();
    ~OutputWorker();
    void submit(std::function<void()> job); // blocks while 16 jobs are already pending
    void drain();                           // waits for all jobs; rethrows the first failure
  private:
    void loop();
    std::thread thread_;
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::function<void()>> q_;
    bool stop_ = false, busy_ = false;
    std::string error_;
};

// The reference's host events draw from libc rand() after srand(rseed) (main.cpp:67, updater.cpp:118,235).  This is
// the same generator (glibc TYPE_3 additive feedback, identical sequence — checked in tests/test_events.py) held in
// a private state: no global lock per draw (hydrolysis makes ~66 k draws per event at 256 trajectories) and no
// interference from other users of rand() in the process.
class HostRand {
  public:
    HostRand() { seed(1); }
    // glibc srandom_r for the 31-word additive-feedback generator (TYPE_3, x^31 + x^3 + 1) that rand() uses
    void seed(unsigned s)
    {
        int32_t word = s ? (int32_t)s : 1;
        r_[0] = word;
        for (int i = 1; i < 31; i++) { // 16807 * word mod (2^31 - 1), Schrage
            const long hi = word / 127773, lo = word % 127773;
            word = (int32_t)(16807 * lo - 2836 * hi);
            if (word < 0) word += 2147483647;
            r_[i] = word;
        }
        f_ = 3;
        b_ = 0;
        for (int i = 0; i < 310; i++) next();
    }
    inline int next()
    {
        const uint32_t v = (uint32_t)r_[f_] + (uint32_t)r_[b_];
        r_[f_] = (int32_t)v;
        if (++f_ >= 31) f_ = 0;
        if (++b_ >= 31) b_ = 0;
        return (int)(v >> 1);
    }

    // The next `count` raw outputs (what `count` calls of next() would return, in order) written to out[], and the
    // generator advanced past them.  The recurrence x[n] = x[n-31] + x[n-3] is unrolled over a linear buffer (no cursor
    // wrap, no store-to-load stall on a 31-word ring): ~0.4 ns per draw, so one thread can pre-draw the stream of a whole
    // hydrolysis event and the per-dimer tests can then run in parallel with the reference's draw order preserved.
    void fill(uint32_t *out, size_t count)
    {
        if (count == 0) return;
        std::vector<uint32_t> x(31 + count);
        for (int i = 0; i < 31; i++) x[i] = (uint32_t)r_[(f_ + i) % 31]; // oldest first: r_[f_] is x[n-31] (overwritten next), r_[b_] is x[n-3]
        uint32_t *p = x.data() + 31;
        for (size_t n = 0; n < count; n++) p[n] = p[(ptrdiff_t)n - 31] + p[(ptrdiff_t)n - 3];
        for (size_t n = 0; n < count; n++) out[n] = p[n] >> 1;
        // the ring afterwards: the newest 31 values, cursors advanced by count
        const int nb = (int)((b_ + count) % 31), nf = (int)((f_ + count) % 31);
        for (int i = 0; i < 31; i++) r_[(nf + i) % 31] = (int32_t)x[count + i];
        b_ = nb;
        f_ = nf;
    }

    // The 31 words of the generator, oldest first: the next draw is (w[0] + w[28]) >> 1.  With discard() this is all the
    // device needs to produce the stream itself (maddy_hydrolysis_plan) and all the host needs to follow it afterwards.
    void window(uint32_t w[31]) const
    {
        for (int i = 0; i < 31; i++) w[i] = (uint32_t)r_[(f_ + i) % 31];
    }
    void set_window(const uint32_t w[31])
    {
        for (int i = 0; i < 31; i++) r_[i] = (int32_t)w[i];
        f_ = 0;
        b_ = 28;
    }
    // advance by n draws without making them (polynomial jump-ahead, maddy_rand_discard)
    void discard(unsigned long long n)
    {
        if (n == 0) return;
        uint32_t w[31];
        window(w);
        maddy_rand_discard(w, n);
        set_window(w);
    }

    // full generator state, for checkpoints: 31 words + the two cursors
    void get_state(int32_t out[33]) const
    {
        for (int i = 0; i < 31; i++) out[i] = r_[i];
        out[31] = f_;
        out[32] = b_;
    }
    void set_state(const int32_t in[33])
    {
        for (int i = 0; i < 31; i++) r_[i] = in[i];
        f_ = in[31] % 31;
        b_ = in[32] % 31;
    }

  private:
    int32_t r_[31];
    int f_ = 3, b_ = 0;
};

struct System {
    HostRand rng;
    std::shared_ptr<OutputWorker> writer; // set by compute() while the loop runs; null = synchronous output
    std::shared_ptr<void> writer_files;   // append-mode handles the writer keeps open during the loop
    maddy_params par{};
    HostParams hp;
    ParamTable table;
    std::vector<std::string> overrides;
    PDB pdb, pdb_ang, coordspdb;
    DCDHeader dcd;
    // coordinates / forces, AoS7 {x,y,z,fi,theta,psi,w} per (traj, monomer)   (globals r, f)
    std::vector<float> r, f;
    // topology (global `top`)
    std::vector<int> harmonic_count, harmonic, longitudinal_count, longitudinal, lateral_count, lateral;
    std::vector<unsigned char> fixed, extra;
    std::vector<int> mon_type, gtp, on_tubule_cur, on_tubule_prev;
    std::vector<double> energies; // [Ntr][7] per-trajectory sums of the last energy evaluation
    // ensemble statistics of the last stride with energies (maddy_ensemble_stats: sum(7), sum of squares(7), count, 0),
    // all-reduced over the GPUs with NCCL; filled when the run is sharded over more than one GPU (or MADDY_ENSEMBLE_STATS=1)
    std::vector<double> ensemble_stats;
    bool quiet = false;           // suppress the reference's stdout chatter (tests / bench)
    std::string *event_log = nullptr; // when set, hydrolyse() appends its messages here instead of printing them (the
                                      // step loop evaluates events ahead of time and prints them in the reference's order)
    std::shared_ptr<std::vector<int>> gtp_for_output; // GTP state as of the pending stride output (see compute())
    bool write_files = true;      // DCD / mt_len.dat / hydrolysis.pdb side effects
    std::string workdir;          // relative output paths are taken relative to the cwd, like the reference
    // timers (timer.cpp)
    long initial_clock = 0, last_clock = 0;

    maddy_topology topology_view(int traj_first = 0) const;
};

// initParameters (+ read_PDB, DCD headers, restart) — preparator.cpp:4-246
void init_parameters(System &s, const std::string &config, const std::vector<std::string> &overrides);
void assembly_init(System &s);                 // preparator.cpp:717-731
void save_coord_pdb(System &s, const std::string &xyz, const std::string &ang); // :566-580
void save_coord_dcd(System &s);                // :582-602
void append_coord_pdb(System &s);              // :604-636
void read_restart(System &s);                  // :664-684
void write_restart(System &s, long long step); // :686-714

// updater.cpp
void mt_length(System &s, long long step, std::vector<int> &mt_len);
// the two halves of mt_length(): the classification (on_tubule_cur, counts) and its output (tubule[] lines, mt_len.dat);
// the step loop uses the second half alone when the device classified (MADDY_SNAP_ONTUBULE)
void mt_length_classify(System &s, std::vector<int> &mt_len);
void mt_length_output(System &s, long long step, const std::vector<int> &mt_len);
struct Insertion { // one dimer placed by change_conc(): global monomer index of its first monomer, position, z of the second
    size_t q;
    float x, y, z, z2;
};
int change_conc(System &s, std::vector<int> &delta, std::vector<int> &mt_len, std::vector<Insertion> *log = nullptr);
void hydrolyse(System &s);
void output_all_energies(System &s, long long step);
void output_sum_force(System &s);
void output_forces(System &s);
void update(System &s, long long step, std::vector<int> &mt_len);
void init_timer(System &s);

// exact checkpoint / restart (checkpoint.cpp)
struct CheckpointState {
    long long step = 0, hydrolysed_for = -1;
    std::vector<float> coords;  // [Ntr*Ntot][7], device state at `step`
    std::vector<uint32_t> rng;  // [Ntr*Ntot][8]: xyz stream, angular stream per monomer
    std::vector<int> mt_len, mt_len_prev;
};
bool checkpoint_peek(const std::string &name, long long *step);
void checkpoint_save(System &s, const std::string &name, const CheckpointState &st);
void checkpoint_load(System &s, const std::string &name, CheckpointState &st);
void checkpoint_trim_outputs(System &s, long long step);

// the step loop: compute() of compute_cuda.cu:1125-1260 over the C-ABI.
// fused = false uses one C-ABI call per reference launch (force; integrate) — same results.
struct ComputeStats {
    long long steps = 0, launches = 0;
    double h2d_bytes = 0, d2h_bytes = 0;
};
void compute(System &s, bool fused = true, ComputeStats *stats = nullptr);

} // namespace mt
