/*
 * compute_b200_stub.cpp — the reference-side binding of the C-ABI, complete.
 * TEST INFRASTRUCTURE: proves the drop-in boundary from the reference's side.
 *
 * mt_b200/build.py::build_reference compiles the reference's OWN host sources where they lie
 * (/root/reference/src/{main,preparator,updater,configreader,parameters,globals,wrapper,timer,dcdio,pdbio,xyzio}.cpp)
 * together with this file INSTEAD of compute_cuda.cu, bdhitea.cu, bdhitea_kernel.cu and HybridTaus.cu, and links
 * libmaddy_b200.so: oracle/_ref/mt_stub.  The reference's main() calls compute() (main.cpp:79) exactly as before; every
 * device-side statement of its loop (compute_cuda.cu:1125-1260) is one C-ABI call here, and the host events it calls
 * back into (hydrolyse, mt_length, change_conc, update; updater.h:12-15) are the reference's own, untouched.
 *
 * Layout facts relied upon (checked below): Coord = 7 floats {x,y,z,fi,theta,psi,w} (mt.h:63-71), Energies = 7 doubles
 * (mt.h:94-102), sizeof(bool) == 1.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include "globals.h"
#include "updater.h"
#include "wrapper.h"
#include "maddy_b200.h"

static_assert(sizeof(Coord) == MADDY_COORD_STRIDE * sizeof(float), "Coord is not 7 floats");
static_assert(sizeof(Energies) == MADDY_ENERGY_TERMS * sizeof(double), "Energies is not 7 doubles");
static_assert(sizeof(bool) == 1, "bool arrays are passed as unsigned char");

static maddy_params params_from_reference(const Parameters &p, const Topology &t)
{
    maddy_params m;
    memset(&m, 0, sizeof m);
    m.abi_version = MADDY_ABI_VERSION;
    m.n_tot = p.Ntot;
    m.n_tr = p.Ntr;
    m.traj_first = 0;
    m.n_tr_local = p.Ntr;
    m.device = p.device;
    m.rseed = p.rseed;
    m.dt = p.dt;
    m.Temp = p.Temp;
    m.gammaR = p.gammaR;
    m.gammaTheta = p.gammaTheta;
    m.varR = p.varR;
    m.varTheta = p.varTheta;
    m.alpha = p.alpha;
    m.freeze_temp = p.freeze_temp;
    m.C = p.C;
    m.B_psi = p.B_psi;
    m.B_fi = p.B_fi;
    m.B_theta = p.B_theta;
    m.psi_0 = p.psi_0;
    m.fi_0 = p.fi_0;
    m.theta0_gtp = p.theta0_gtp;
    m.theta0_gdp = p.theta0_gdp;
    m.A_long = p.A_long;
    m.D_long = p.D_long;
    m.A_lat = p.A_lat;
    m.D_lat = p.D_lat;
    m.seam_coeff = p.seam_coeff;
    m.barrier = p.barrier;
    m.a_barr_long = p.a_barr_long;
    m.r_barr_long = p.r_barr_long;
    m.w_barr_long = p.w_barr_long;
    m.a_barr_lat = p.a_barr_lat;
    m.r_barr_lat = p.r_barr_lat;
    m.w_barr_lat = p.w_barr_lat;
    m.lj_on = p.lj_on;
    m.ljscale = p.ljscale;
    m.ljsigma6 = p.ljsigma6;
    m.ljpairscutoff = p.ljpairscutoff;
    m.ljpairsupdatefreq = p.ljpairsupdatefreq;
    m.is_wall = p.is_wall;
    m.rep_leftborder = p.rep_leftborder;
    m.rep_r = p.rep_r;
    m.rep_eps = p.rep_eps;
    m.rep_h = p.rep_h;
    m.is_assembly = p.is_assembly;
    m.tea_on = p.hdi_on;
    if (p.hdi_on) { // preparator.cpp:61-65 fills `tea` only then
        m.tea_a = tea.a;
        m.tea_epsilon_freq = tea.epsilon_freq;
        m.tea_capricious = tea.capricious;
        m.tea_epsmax = tea.epsmax;
    }
    m.max_harmonic = t.maxHarmonicPerMonomer;
    m.max_longitudinal = t.maxLongitudinalPerMonomer;
    m.max_lateral = t.maxLateralPerMonomer;
    return m;
}

static maddy_handle *g_handle;
#define CK(call)                                                                    \
    do {                                                                            \
        int rc_ = (call);                                                           \
        if (rc_) DIE("%s failed (%d): %s", #call, rc_, maddy_last_error(g_handle)); \
    } while (0)

void compute(Coord *r, Coord *f, Parameters &par, Topology &top, Energies *energies)
{
    const maddy_params mp = params_from_reference(par, top);
    maddy_topology mt;
    mt.harmonic_count = top.harmonicCount;
    mt.harmonic = top.harmonic;
    mt.longitudinal_count = top.longitudinalCount;
    mt.longitudinal = top.longitudinal;
    mt.lateral_count = top.lateralCount;
    mt.lateral = top.lateral;
    mt.fixed = (const unsigned char *)top.fixed;
    mt.extra = (const unsigned char *)top.extra;
    mt.mon_type = top.mon_type;
    mt.gtp = top.gtp;
    mt.on_tubule_cur = top.on_tubule_cur;
    // initIntegration + initRand (+ initTeaIntegrator): compute_cuda.cu:977-1098
    if (maddy_create(&mp, &mt, (const float *)r, NULL, &g_handle)) DIE("maddy_create: %s", maddy_last_error(NULL));
    maddy_handle *h = g_handle;
    const size_t n = (size_t)par.Ntot * par.Ntr;
    memset(f, 0, n * sizeof(Coord)); // :995-1002

    int *mt_len = (int *)malloc(par.Ntr * sizeof(int));
    int *mt_len_prev = (int *)malloc(par.Ntr * sizeof(int));
    double *per_traj = (double *)malloc((size_t)par.Ntr * MADDY_ENERGY_TERMS * sizeof(double));

    long long step = 0;
    while (step < par.steps) {
        // :1140-1151  lists are rebuilt before the host events of the step
        const bool rebuilt = step % par.ljpairsupdatefreq == 0;
        if (rebuilt) {
            if (par.lj_on) CK(maddy_rebuild_lj(h));
            if (par.is_assembly) CK(maddy_rebuild_bonds(h));
        }
        // :1153-1160
        if (par.hydrolysis && step % par.hydrostep == 0 && step != 0) {
            hydrolyse();
            CK(maddy_upload_gtp(h, top.gtp));
        }
        // :1163-1226  stride block
        if (step % par.stride == 0) {
            if (par.out_energy) CK(maddy_energies(h, per_traj, (double *)energies)); // per-monomer array for OutputAllEnergies
            if (par.out_force) CK(maddy_download_forces(h, (float *)f));
            CK(maddy_download_coords(h, (float *)r));
            if (par.tub_length) {
                memcpy(top.on_tubule_prev, top.on_tubule_cur, n * sizeof(int));
                if (step != 0) {
                    memcpy(mt_len_prev, mt_len, par.Ntr * sizeof(int));
                    mt_length(step, mt_len);
                    if (par.barrier) CK(maddy_upload_on_tubule(h, top.on_tubule_cur));
                    if (par.is_const_conc) {
                        for (int t = 0; t < par.Ntr; t++) mt_len_prev[t] = mt_len[t] - mt_len_prev[t];
                        if (change_conc(mt_len_prev, mt_len)) {
                            CK(maddy_upload_extra(h, (const unsigned char *)top.extra));
                            CK(maddy_upload_coords(h, (const float *)r));
                        }
                    }
                    update(step, mt_len);
                } else {
                    update(step, mt_len);
                    mt_length(step, mt_len);
                }
            } else {
                update(step, mt_len);
            }
        }
        // :1228-1238  force + integration (TEA included) for every step up to the next host event, one fused call
        long long next = std::min<long long>(par.steps, (step / par.stride + 1) * par.stride);
        if (par.hydrolysis && par.hydrostep > 0) next = std::min<long long>(next, (step / par.hydrostep + 1) * par.hydrostep);
        CK(maddy_run(h, step, next - step, rebuilt ? MADDY_RUN_SKIP_FIRST_REBUILD : 0u));
        step = next;
    }
    CK(maddy_sync(h));
    // deleteIntegration: the reference leaves r as of the last stride; so does this
    maddy_destroy(h);
    g_handle = NULL;
    free(mt_len);
    free(mt_len_prev);
    free(per_traj);
}
