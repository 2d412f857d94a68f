/*
 * maddy_oracle.h — CPU restatement of the MADDY hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (mt_b200/) never does and has no CPU fallback.
 *
 * Parity status: the reference ships NO tests, golden vectors or fixtures for this path
 * ("parity unpinned" by the reference's own tests).  The oracle is pinned
 * instead against (a) known answers produced by the reference's own host code
 * (generateSeeds, tests/golden/seeds_*.json, script oracle/gen_golden_seeds.cu) and (b) dumps
 * of the reference's own CUDA kernels run on a B200 through oracle/ref_probe.cu
 * (tests/golden/ref_*.npz).  Integer work (seed table, HybridTaus stream, LJ lists on given
 * coordinates) is exact; float work uses libm where the reference's -use_fast_math build uses
 * MUFU approximations, so forces/energies/coordinates agree to a stated tolerance only.
 *
 * Layouts are the reference's own: coordinates AoS {x,y,z,fi,theta,psi,w} (mt.h:63-71), lists
 * [traj][i][capacity] with the signed / ZERO-sentinel encodings, energies 7 doubles per monomer
 * in the field order of `Energies` (mt.h:94-102).
 */
#ifndef MADDY_ORACLE_H_
#define MADDY_ORACLE_H_
#include "maddy_b200.h" /* maddy_params / maddy_topology: plain data definitions only */

#ifdef __cplusplus
extern "C" {
#endif

/* mutable list storage in the reference layout, for n_tr_local trajectories */
typedef struct oracle_lists {
    int *longitudinal_count, *longitudinal; /* capacity par->max_longitudinal */
    int *lateral_count, *lateral;           /* capacity par->max_lateral      */
    int *lj_count, *lj;                     /* capacity MADDY_LJ_CAPACITY     */
} oracle_lists;

/* ran2 + generateSeeds (ran2.h:18-56, HybridTaus.cu:32-48), fresh static state */
void oracle_generate_seeds(unsigned *seeds, int rseed, long long np);
/* one HybridTaus draw (HybridTaus.cu:63-76); state[4] updated */
unsigned oracle_hybrid_taus(unsigned *state);
/* rforce (HybridTaus.cu:85-98) with libm; out[4]; state[4] updated */
void oracle_rforce(unsigned *state, float *out);

/* LJ_kernel (compute_cuda.cu:913-940) */
void oracle_lj_lists(const maddy_params *par, const maddy_topology *top, const float *coords, oracle_lists *l);
/* pairs_kernel (compute_cuda.cu:527-674) */
void oracle_pair_lists(const maddy_params *par, const maddy_topology *top, const float *coords, oracle_lists *l);
/* compute_kernel (compute_cuda.cu:32-525): forces AoS7 [n_tr_local*n_tot*7]; extras left untouched */
void oracle_forces(const maddy_params *par, const maddy_topology *top, const oracle_lists *l, const float *coords, float *forces);
/* energy_kernel (compute_cuda.cu:676-911): energies [n_tr_local*n_tot][7] doubles */
void oracle_energies(const maddy_params *par, const maddy_topology *top, const oracle_lists *l, const float *coords, double *energies);
/* integrate_kernel (compute_cuda.cu:943-975): rng = [2][n_tr_local*n_tot][4] (xyz streams, angular streams) */
void oracle_integrate(const maddy_params *par, const maddy_topology *top, float *coords, float *forces, unsigned *rng);
/* steps [first, first+n) of the loop body: rebuild at step % freq == 0 (unless skip_first), force, integrate */
void oracle_run(const maddy_params *par, const maddy_topology *top, oracle_lists *l, float *coords, float *forces, unsigned *rng,
                long long first_step, long long n_steps, int skip_first_rebuild);

/* beta from the epsilon sum (bdhitea.cu:79-113); MADDY_ETEA where the reference exits */
int oracle_tea_beta(double epsilon_sum, int n_noextra, int capricious, float tea_a, float epsmax, float *beta, double *eps_out);
/* TEA (bdhitea_kernel.cu:16-213, bdhitea.cu:57-118): ci [n][4] floats, eps [n] floats, beta [n_tr_local] */
int oracle_tea_update(const maddy_params *par, const maddy_topology *top, const float *coords, float *ci, float *eps, float *beta);
void oracle_tea_integrate(const maddy_params *par, const maddy_topology *top, float *coords, float *forces, unsigned *rng,
                          const float *ci, const float *beta);

#ifdef __cplusplus
}
#endif
#endif
