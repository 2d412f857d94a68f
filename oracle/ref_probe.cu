/*
 * ref_probe.cu — instrumented driver around the UNMODIFIED reference kernels.  TEST INFRASTRUCTURE.
 *
 * Compiled by mt_b200/build.py together with the reference's own translation units from
 * /root/reference/src (in place, nothing copied) into oracle/_ref/ref_probe.  It replaces only
 * the reference's main()/compute() driver: the reference's initParameters(), AssemblyInit(),
 * initIntegration(), its five __global__ kernels, its TEA integrator and its host events are
 * called as they are, in the order compute() calls them (src/compute_cuda.cu:1137-1238), and the
 * device state is dumped after every phase so that the B200 kernels of this repo can be compared
 * with the reference's own kernels on the same GPU, phase by phase.
 *
 *   ref_probe <config.conf> <out.bin> <window_steps> <detail_steps> [name=value ...]
 *
 * Host events (hydrolysis, stride block) are NOT run inside the window; probe-only keys:
 *   probe_gdp_every=K   before step 0 mark every K-th dimer GDP (gtp = 0) and upload, to reach the
 *                       theta0_gdp branches
 *   probe_ontub=1       before step 0 run the reference's mt_length() and upload on_tubule_cur, to
 *                       reach the barrier on/off branches
 * Records: {char tag[8]; int64 step; int64 nbytes; payload}.  The unmodified `mt` binary run with
 * stride 1 validates this driver: its DCD frames must equal the "coords" records bit for bit.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "compute_cuda.cuh"
#include "configreader.h"
#include "preparator.h"
#include "updater.h"
#include "bdhitea.cuh"
#include "ht.cuh"

// file-scope state of HybridTaus.cu (not exported by ht.cuh): same layout, declared here to read d_seeds
struct HybTau {
    uint4 *h_seeds;
    uint4 *d_seeds;
    uint4 mseed;
};
extern HybTau ht;

static FILE *g_out;

static void rec(const char *tag, long long step, const void *data, long long nbytes)
{
    char t[8] = {0};
    strncpy(t, tag, 8);
    fwrite(t, 8, 1, g_out);
    fwrite(&step, 8, 1, g_out);
    fwrite(&nbytes, 8, 1, g_out);
    if (nbytes) fwrite(data, 1, (size_t)nbytes, g_out);
}

template <typename T> static void rec_dev(const char *tag, long long step, const T *dptr, size_t count)
{
    std::vector<T> h(count);
    cudaMemcpy(h.data(), dptr, count * sizeof(T), cudaMemcpyDeviceToHost);
    checkCUDAError(tag);
    rec(tag, step, h.data(), (long long)(count * sizeof(T)));
}

int main(int argc, char *argv[])
{
    if (argc < 5) {
        fprintf(stderr, "usage: %s <config.conf> <out.bin> <window_steps> <detail_steps> [name=value ...]\n", argv[0]);
        return 2;
    }
    const long long window = atoll(argv[3]), detail = atoll(argv[4]);
    // the reference's argv override convention: argv[0]=binary, argv[1]=config, argv[2..]=name=value
    std::vector<char *> av;
    av.push_back(argv[0]);
    av.push_back(argv[1]);
    for (int i = 5; i < argc; i++) av.push_back(argv[i]);
    int ac = (int)av.size();

    parseParametersFile(av[1], ac, av.data());
    const int gdp_every = getIntegerParameter("probe_gdp_every", 0);
    const int ontub = getIntegerParameter("probe_ontub", 0);
    initParameters(ac, av.data());
    srand(par.rseed);
    if (par.is_assembly) AssemblyInit();
    if (par.out_energy) energies = (Energies *)malloc(par.Ntot * par.Ntr * sizeof(Energies));

    g_out = fopen(argv[2], "wb");
    if (!g_out) {
        perror(argv[2]);
        return 2;
    }
    const size_t n = (size_t)par.Ntot * par.Ntr;
    {
        long long hdr[8] = {par.Ntot, par.Ntr, window, detail, top.maxHarmonicPerMonomer, top.maxLongitudinalPerMonomer,
                            top.maxLateralPerMonomer, (long long)sizeof(Parameters)};
        rec("header", -1, hdr, sizeof hdr);
        rec("params", -1, &par, sizeof(Parameters));
        rec("harm", -1, top.harmonic, (long long)(sizeof(int) * par.Ntot * top.maxHarmonicPerMonomer));
        rec("harmcnt", -1, top.harmonicCount, (long long)(sizeof(int) * par.Ntot));
        rec("montype", -1, top.mon_type, (long long)(sizeof(int) * par.Ntot));
        rec("fixed", -1, top.fixed, (long long)(sizeof(bool) * par.Ntot));
        rec("extra", -1, top.extra, (long long)(sizeof(bool) * n));
        rec("r_host", -1, r, (long long)(sizeof(Coord) * n));
    }

    initIntegration(r, f, par, top, energies);
    if (par.hdi_on) initTeaIntegrator();
    const int grid = par.Ntot * par.Ntr / BLOCK_SIZE + 1;

    if (gdp_every > 0) {
        for (int tr = 0; tr < par.Ntr; tr++)
            for (int i = 0; i < par.Ntot; i += 2)
                if ((i / 2) % gdp_every == 0) top.gtp[i + tr * par.Ntot] = top.gtp[i + 1 + tr * par.Ntot] = 0;
        cudaMemcpy(topGPU.gtp, top.gtp, n * sizeof(int), cudaMemcpyHostToDevice);
        checkCUDAError("probe gtp");
    }
    if (ontub) {
        std::vector<int> len(par.Ntr);
        mt_length(1, len.data()); // step != 0: no file truncation
        cudaMemcpy(topGPU.on_tubule_cur, top.on_tubule_cur, n * sizeof(int), cudaMemcpyHostToDevice);
        checkCUDAError("probe on_tubule");
    }
    rec("gtp", -1, top.gtp, (long long)(sizeof(int) * n));
    rec("ontub", -1, top.on_tubule_cur, (long long)(sizeof(int) * n));
    rec_dev("seeds", -1, ht.d_seeds, 2 * n);

    for (long long step = 0; step < window; step++) {
        const bool full = step < detail;
        if (step % par.ljpairsupdatefreq == 0) {
            if (par.lj_on) {
                LJ_kernel<<<grid, BLOCK_SIZE>>>(d_r);
                checkCUDAError("lj_kernel");
                rec_dev("ljcnt", step, topGPU.LJCount, n);
                if (full || step == 0) rec_dev("lj", step, topGPU.LJ, n * topGPU.maxLJPerMonomer);
            }
            if (par.is_assembly) {
                pairs_kernel<<<grid, BLOCK_SIZE>>>(d_r);
                checkCUDAError("pairs_kernel");
                rec_dev("longcnt", step, topGPU.longitudinalCount, n);
                rec_dev("long", step, topGPU.longitudinal, n * topGPU.maxLongitudinalPerMonomer);
                rec_dev("latcnt", step, topGPU.lateralCount, n);
                rec_dev("lat", step, topGPU.lateral, n * topGPU.maxLateralPerMonomer);
            }
        }
        if (step == 0 && !par.is_assembly) {
            rec_dev("longcnt", step, topGPU.longitudinalCount, n);
            rec_dev("long", step, topGPU.longitudinal, n * topGPU.maxLongitudinalPerMonomer);
            rec_dev("latcnt", step, topGPU.lateralCount, n);
            rec_dev("lat", step, topGPU.lateral, n * topGPU.maxLateralPerMonomer);
        }
        rec_dev("coords", step, d_r, n);
        if (full && par.out_energy) {
            energy_kernel<<<grid, BLOCK_SIZE>>>(d_r, d_energies);
            checkCUDAError("energy_kernel");
            rec_dev("energy", step, d_energies, n);
        }
        compute_kernel<<<grid, BLOCK_SIZE>>>(d_r, d_f);
        checkCUDAError("compute_kernel");
        if (full) rec_dev("forces", step, d_f, n);
        if (par.hdi_on) {
            updateTea(step);
            if (full && step % tea.epsilon_freq == 0) {
                rec_dev("tea_ci", step, tea.d_ci, n);
                rec_dev("tea_eps", step, tea.d_epsilon, n);
                rec_dev("tea_beta", step, tea.d_beta_ij, (size_t)par.Ntr);
            }
            integrateTea();
        } else {
            integrate_kernel<<<grid, BLOCK_SIZE>>>(d_r, d_f);
            checkCUDAError("integrate_kernel");
        }
    }
    rec_dev("coords", window, d_r, n);
    rec_dev("seeds", window, ht.d_seeds, 2 * n);
    rec("end", window, nullptr, 0);
    fclose(g_out);
    return 0;
}
