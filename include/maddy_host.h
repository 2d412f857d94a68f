/*
 * maddy_host.h — C-ABI over the drop-in C++ host (mt_b200/host), for bindings and tests.
 *
 * The host side of the boundary is what the reference does around compute():
 * parse config.conf / forcefield / conditions (src/preparator.cpp:4-246), build the topology
 * from the PDB pair (src/preparator.cpp:249-561), run the step loop with its host events
 * (src/compute_cuda.cu:1125-1260, src/updater.cpp) and write DCD / PDB outputs.
 * Every function returns 0 on success, non-zero on a fatal condition (the reference would
 * print "Fatal error!" and exit(-1)); mt_host_last_error() has the message.
 */
#ifndef MADDY_HOST_H_
#define MADDY_HOST_H_
#include "maddy_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mt_system mt_system;

/* flags for mt_system_load */
#define MT_LOAD_QUIET 1u     /* no stdout chatter */
#define MT_LOAD_NO_FILES 2u  /* do not create/append DCD, mt_len.dat, hydrolysis.pdb */

typedef struct mt_host_params {
    long long steps, firststep, stride;
    long long hydrostep;
    int fix, tub_length, out_energy, out_force, is_restart, is_const_conc, hydrolysis, n_gpus;
    float conc, khydro, viscosity;
} mt_host_params;

const char *mt_host_last_error(void);

/* initParameters (+ AssemblyInit when is_assembly) for `mt <config> [name=value ...]` */
int mt_system_load(const char *config_path, int n_overrides, const char *const *overrides, unsigned flags, mt_system **out);
void mt_system_free(mt_system *s);

/* views into the system (valid until mt_system_free; the arrays are the live host state) */
int mt_system_params(const mt_system *s, maddy_params *par, mt_host_params *host);
int mt_system_topology(const mt_system *s, maddy_topology *top);
float *mt_system_coords(mt_system *s);            /* [n_tr*n_tot*7] AoS Coord */
int *mt_system_gtp(mt_system *s);
int *mt_system_on_tubule(mt_system *s, int prev);
unsigned char *mt_system_extra(mt_system *s);
double *mt_system_energies(mt_system *s);         /* [n_tr][7] after a compute with output_energy */
/* ensemble statistics of the last stride with energies: MADDY_ENSEMBLE_STATS doubles (maddy_ensemble_stats_end), or NULL when the
 * run did not reduce them (one GPU and MADDY_ENSEMBLE_STATS unset) */
const double *mt_system_ensemble_stats(const mt_system *s);
/* srand(seed) of the reference main (main.cpp:67) for this system's host events (same sequence as libc rand()) */
int mt_system_srand(mt_system *s, unsigned seed);
/* the system's rand() stream (glibc TYPE_3, same sequence as libc rand()): its 31-word window (oldest first, what
 * maddy_hydrolysis_plan takes), a jump over n draws, and one draw */
int mt_system_rand_window(mt_system *s, unsigned *window31);
int mt_system_rand_discard(mt_system *s, unsigned long long n);
int mt_system_rand_next(mt_system *s);
int mt_system_set_ngpus(mt_system *s, int n_gpus);
int mt_system_set_steps(mt_system *s, long long steps);

/* the drop-in compute(): full step loop with host events. fused=0 issues one C-ABI call per
 * reference kernel launch (same results). stats (may be NULL): [steps, launches, h2d bytes, d2h bytes] */
int mt_system_compute(mt_system *s, int fused, double *stats4);

/* host events, individually (updater.cpp) */
int mt_system_mt_length(mt_system *s, long long step, int *mt_len);
int mt_system_hydrolyse(mt_system *s);
int mt_system_change_conc(mt_system *s, int *delta, int *mt_len, int *changed);
int mt_system_save_pdb(mt_system *s, const char *xyz, const char *ang);

/* file formats */
int mt_dcd_read(const char *path, int *n_atoms, int *n_frames, float *xyz_out, long long capacity_floats);
int mt_pdb_count(const char *path);

#ifdef __cplusplus
}
#endif
#endif
