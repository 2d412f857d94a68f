/* host_capi.cpp — C-ABI of the drop-in host declared in include/maddy_host.h */
#include <cstring>
#include "maddy_host.h"
#include "mt_host.hpp"

struct mt_system {
    mt::System sys;
};
static thread_local std::string g_err;
const char *mt_host_last_error(void) { return g_err.c_str(); }

#define GUARD(...)                       \
    try {                                \
        __VA_ARGS__;                     \
        return 0;                        \
    } catch (const std::exception &e) {  \
        g_err = e.what();                \
        return 1;                        \
    }

int mt_system_load(const char *config, int n_over, const char *const *over, unsigned flags, mt_system **out)
{
    if (!config || !out) {
        g_err = "null argument";
        return 1;
    }
    mt_system *s = new mt_system;
    try {
        s->sys.quiet = flags & MT_LOAD_QUIET;
        s->sys.write_files = !(flags & MT_LOAD_NO_FILES);
        std::vector<std::string> ov;
        for (int i = 0; i < n_over; i++) ov.emplace_back(over[i]);
        mt::init_parameters(s->sys, config, ov);
        if (s->sys.par.is_assembly) mt::assembly_init(s->sys);
        *out = s;
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        delete s;
        return 1;
    }
}
void mt_system_free(mt_system *s) { delete s; }

int mt_system_params(const mt_system *s, maddy_params *par, mt_host_params *host)
{
    if (!s) return 1;
    if (par) *par = s->sys.par;
    if (host) {
        const mt::HostParams &h = s->sys.hp;
        host->steps = h.steps;
        host->firststep = h.firststep;
        host->stride = h.stride;
        host->hydrostep = h.hydrostep;
        host->fix = h.fix;
        host->tub_length = h.tub_length;
        host->out_energy = h.out_energy;
        host->out_force = h.out_force;
        host->is_restart = h.is_restart;
        host->is_const_conc = h.is_const_conc;
        host->hydrolysis = h.hydrolysis;
        host->n_gpus = h.n_gpus;
        host->conc = h.conc;
        host->khydro = h.khydro;
        host->viscosity = h.viscosity;
    }
    return 0;
}
int mt_system_topology(const mt_system *s, maddy_topology *top)
{
    if (!s || !top) return 1;
    *top = s->sys.topology_view(0);
    return 0;
}
float *mt_system_coords(mt_system *s) { return s->sys.r.data(); }
int *mt_system_gtp(mt_system *s) { return s->sys.gtp.data(); }
int *mt_system_on_tubule(mt_system *s, int prev) { return prev ? s->sys.on_tubule_prev.data() : s->sys.on_tubule_cur.data(); }
unsigned char *mt_system_extra(mt_system *s) { return s->sys.extra.data(); }
double *mt_system_energies(mt_system *s) { return s->sys.energies.data(); }
const double *mt_system_ensemble_stats(const mt_system *s) { return s->sys.ensemble_stats.empty() ? nullptr : s->sys.ensemble_stats.data(); }
int mt_system_srand(mt_system *s, unsigned seed)
{
    s->sys.rng.seed(seed);
    return 0;
}
int mt_system_rand_window(mt_system *s, unsigned *w31)
{
    s->sys.rng.window(w31);
    return 0;
}
int mt_system_rand_discard(mt_system *s, unsigned long long n)
{
    s->sys.rng.discard(n);
    return 0;
}
int mt_system_rand_next(mt_system *s) { return s->sys.rng.next(); }
int mt_system_set_ngpus(mt_system *s, int n)
{
    s->sys.hp.n_gpus = n;
    return 0;
}
int mt_system_set_steps(mt_system *s, long long steps)
{
    s->sys.hp.steps = steps;
    return 0;
}

int mt_system_compute(mt_system *s, int fused, double *stats4)
{
    GUARD({
        mt::ComputeStats st;
        mt::compute(s->sys, fused != 0, &st);
        if (stats4) {
            stats4[0] = (double)st.steps;
            stats4[1] = (double)st.launches;
            stats4[2] = st.h2d_bytes;
            stats4[3] = st.d2h_bytes;
        }
    });
}
int mt_system_mt_length(mt_system *s, long long step, int *mt_len)
{
    GUARD({
        std::vector<int> v(s->sys.par.n_tr);
        mt::mt_length(s->sys, step, v);
        if (mt_len) memcpy(mt_len, v.data(), v.size() * sizeof(int));
    });
}
int mt_system_hydrolyse(mt_system *s) { GUARD(mt::hydrolyse(s->sys)); }
int mt_system_change_conc(mt_system *s, int *delta, int *mt_len, int *changed)
{
    GUARD({
        std::vector<int> d(delta, delta + s->sys.par.n_tr), m(mt_len, mt_len + s->sys.par.n_tr);
        int c = mt::change_conc(s->sys, d, m);
        if (changed) *changed = c;
    });
}
int mt_system_save_pdb(mt_system *s, const char *xyz, const char *ang) { GUARD(mt::save_coord_pdb(s->sys, xyz, ang)); }

int mt_dcd_read(const char *path, int *n_atoms, int *n_frames, float *out, long long cap)
{
    FILE *f = fopen(path, "rb");
    if (!f) {
        g_err = std::string("cannot open ") + path;
        return 1;
    }
    mt::DCDHeader h;
    if (!mt::dcd_read_header(f, h)) {
        fclose(f);
        g_err = "bad DCD header";
        return 1;
    }
    const int N = h.n_atoms;
    std::vector<float> x(N), y(N), z(N);
    int frames = 0;
    while (mt::dcd_read_frame(f, N, x.data(), y.data(), z.data())) {
        if (out && (long long)(frames + 1) * N * 3 <= cap) {
            float *o = out + (size_t)frames * N * 3;
            for (int i = 0; i < N; i++) {
                o[3 * i] = x[i];
                o[3 * i + 1] = y[i];
                o[3 * i + 2] = z[i];
            }
        }
        frames++;
    }
    fclose(f);
    if (n_atoms) *n_atoms = N;
    if (n_frames) *n_frames = frames;
    return 0;
}
int mt_pdb_count(const char *path)
{
    try {
        mt::PDB p;
        mt::read_pdb(path, p, true);
        return (int)p.atoms.size();
    } catch (...) {
        return -1;
    }
}
