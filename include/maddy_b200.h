/*
 * maddy_b200.h — C-ABI of the B200-native Langevin/BD step loop of MADDY (klyshko/MT).
 *
 * This is the drop-in boundary for ONE path of the reference: the body of
 *   void compute(Coord* r, Coord* f, Parameters &par, Topology &top, Energies* energies)
 * (reference src/compute_cuda.cu:1125-1260, declared src/compute_cuda.cuh:17, called once
 * from src/main.cpp:79) and the kernels it launches.  Plain C: pointers and sizes only,
 * no CUDA, torch or C++ types.  Every call returns 0 on success or a negative
 * MADDY_E* code; it never calls exit().  maddy_last_error() gives the message.
 *
 * Conventions
 *  - The caller owns every host array it passes; the library copies what it needs and
 *    owns all device memory (as compute() owns d_r/d_f/topGPU in the reference,
 *    src/compute_cuda.cu:989-1081).
 *  - Coordinates and forces cross the boundary in the reference's AoS "Coord" layout:
 *    7 floats per monomer {x, y, z, fi, theta, psi, w} (src/mt.h:63-71 — note theta
 *    BEFORE psi).  On the device the library keeps two float4 SoA arrays.
 *  - Bond/LJ lists cross the boundary in the reference encoding (signed index, ZERO
 *    sentinel 999999 for +-0 in dynamic lateral lists; src/compute_cuda.cu:588-592,
 *    :646-658, src/preparator.cpp:550-554).
 *  - One handle = one GPU = a contiguous block of trajectories
 *    [traj_first, traj_first + n_tr_local) of a global ensemble of n_tr trajectories.
 *    RNG stream ids are the GLOBAL ones (xyz: traj*n_tot + i, angular:
 *    n_tot*n_tr + traj*n_tot + i; src/compute_cuda.cu:952-953), so a sharded run is
 *    bit-identical to a single-GPU run of the whole ensemble.
 *  - Calls on one handle must be serialised by the caller (the reference is
 *    single-threaded); several handles may coexist (no global state).
 */
#ifndef MADDY_B200_H_
#define MADDY_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define MADDY_ABI_VERSION 1

#define MADDY_COORD_STRIDE 7      /* floats per monomer in the AoS boundary layout   */
#define MADDY_ENERGY_TERMS 7      /* harm,long,lat,psi,fi,teta,lj (updater.cpp:35-36) */
#define MADDY_ZERO_SENTINEL 999999 /* src/mt.h:35 (ZERO)                              */
#define MADDY_LJ_CAPACITY 256     /* src/compute_cuda.cu:1076                         */
#define MADDY_MAX_NTOT 32767      /* bond codes are (index << 1 | sign) in 16 bits    */
#define MADDY_MAX_NTOT_CTA 3200   /* above this a trajectory no longer fits one CTA's shared memory: wide path */

/* status codes */
#define MADDY_OK 0
#define MADDY_EINVAL (-1)   /* bad argument / unsupported configuration   */
#define MADDY_ECUDA (-2)    /* CUDA runtime error (message has the detail) */
#define MADDY_ENOMEM (-3)
#define MADDY_EOVERFLOW (-4) /* a neighbour list exceeded its capacity (UB in the reference) */
#define MADDY_ETEA (-5)     /* TEA "capricious" abort (bdhitea.cu:89-107 would exit(-1)) */
#define MADDY_ENCCL (-6)

/* flags for maddy_run */
#define MADDY_RUN_SKIP_FIRST_REBUILD 1u /* caller already rebuilt lists for first_step */

/* list kinds for maddy_download_list / maddy_upload_list */
#define MADDY_LIST_LONGITUDINAL 0
#define MADDY_LIST_LATERAL 1
#define MADDY_LIST_LJ 2

/*
 * Scalars of the reference's `Parameters` (src/parameters.h:246-318) that the hot path
 * reads, plus the shard description.  Derived constants (gammaR, varR, ... ;
 * src/preparator.cpp:209-221) are computed by the host exactly as the reference does and
 * passed in, so that both sides see the same float values.
 */
typedef struct maddy_params {
    int abi_version;          /* = MADDY_ABI_VERSION */
    int n_tot;                /* par.Ntot: monomers per trajectory */
    int n_tr;                 /* par.Ntr: trajectories in the GLOBAL ensemble */
    int traj_first;           /* first global trajectory held by this handle */
    int n_tr_local;           /* trajectories held by this handle */
    int device;               /* par.device (CUDA ordinal) */
    int rseed;                /* par.rseed */

    float dt, Temp;
    float gammaR, gammaTheta, varR, varTheta, alpha, freeze_temp;

    float C;                                  /* harmonic (intra-dimer) */
    float B_psi, B_fi, B_theta;               /* bending */
    float psi_0, fi_0, theta0_gtp, theta0_gdp;
    float A_long, D_long, A_lat, D_lat, seam_coeff; /* Morse */

    int barrier;
    float a_barr_long, r_barr_long, w_barr_long;
    float a_barr_lat, r_barr_lat, w_barr_lat;   /* w_* already FWHM-converted (preparator.cpp:139,142) */

    int lj_on;
    float ljscale, ljsigma6, ljpairscutoff;
    int ljpairsupdatefreq;

    int is_wall;
    float rep_leftborder, rep_r, rep_eps, rep_h; /* zs[traj] == rep_h for every trajectory (preparator.cpp:176-178) */

    int is_assembly;          /* dynamic bond lists rebuilt by pairs_kernel */

    int tea_on;               /* par.hdi_on */
    float tea_a;
    int tea_epsilon_freq;
    int tea_capricious;
    float tea_epsmax;

    int max_harmonic;         /* top.maxHarmonicPerMonomer */
    int max_longitudinal;     /* top.maxLongitudinalPerMonomer (8 in assembly mode)  */
    int max_lateral;          /* top.maxLateralPerMonomer (16 in assembly mode)      */
} maddy_params;

/*
 * Host copy of the reference's `Topology` (src/mt.h:73-92) for the LOCAL trajectories.
 * bool arrays are passed as unsigned char (sizeof(bool) == 1 in the reference build).
 */
typedef struct maddy_topology {
    const int *harmonic_count;      /* [n_tot]                                  */
    const int *harmonic;            /* [n_tot * max_harmonic], signed           */
    const int *longitudinal_count;  /* [n_tr_local * n_tot]                     */
    const int *longitudinal;        /* [n_tr_local * n_tot * max_longitudinal]  */
    const int *lateral_count;       /* [n_tr_local * n_tot]                     */
    const int *lateral;             /* [n_tr_local * n_tot * max_lateral]       */
    const unsigned char *fixed;     /* [n_tot]                                  */
    const unsigned char *extra;     /* [n_tr_local * n_tot]                     */
    const int *mon_type;            /* [n_tot]                                  */
    const int *gtp;                 /* [n_tr_local * n_tot]                     */
    const int *on_tubule_cur;       /* [n_tr_local * n_tot]                     */
} maddy_topology;

typedef struct maddy_handle maddy_handle;

/* ---- lifetime: replaces initIntegration + initRand + initTeaIntegrator
 *      (compute_cuda.cu:977-1098, HybridTaus.cu:21-31, bdhitea.cu:13-35) and
 *      deleteTeaIntegrator + deleteIntegration (compute_cuda.cu:1100-1123).
 * coords_aos7: [n_tr_local * n_tot * 7] floats; angles are wrapped like
 * compute_cuda.cu:1004-1010 before upload (the caller's array is not modified).
 * stream: a cudaStream_t to enqueue on, or NULL for a library-owned stream. */
int maddy_create(const maddy_params *par, const maddy_topology *top, const float *coords_aos7,
                 void *stream, maddy_handle **out);
int maddy_destroy(maddy_handle *h);
/* message of the last failed call on h (h == NULL: last failed maddy_create) */
const char *maddy_last_error(const maddy_handle *h);
/* the cudaStream_t all work of this handle is enqueued on */
void *maddy_stream(const maddy_handle *h);
int maddy_sync(maddy_handle *h);

/* ---- step-granular entry points (each replaces one launch of the reference loop) */
int maddy_rebuild_lj(maddy_handle *h);     /* LJ_kernel     launch, compute_cuda.cu:1143 */
int maddy_rebuild_bonds(maddy_handle *h);  /* pairs_kernel  launch, compute_cuda.cu:1148 */
int maddy_force(maddy_handle *h);          /* compute_kernel launch, compute_cuda.cu:1228 */
int maddy_integrate(maddy_handle *h);      /* integrate_kernel launch, compute_cuda.cu:1236 */
int maddy_tea_update(maddy_handle *h, long long step); /* updateTea, bdhitea.cu:57-118  */
int maddy_tea_integrate(maddy_handle *h);              /* integrateTea, bdhitea.cu:37-42 */

/* TEA state as the reference's updateTea() sees it (bdhitea.cu:59-60,72-76,115: d_ci, d_epsilon, d_beta_ij): ci4 [n][4]
 * floats (per-bead C_i in .x/.y/.z), epsilon [n], beta [n_tr_local]; any pointer may be NULL. */
int maddy_download_tea(maddy_handle *h, float *ci4, float *epsilon, float *beta);

/* ---- fused execution of steps [first_step, first_step + n_steps): the body of the
 * for(step) loop of compute() between two host events (compute_cuda.cu:1137-1238 minus
 * the hydrolysis / stride blocks).  Lists are rebuilt in-kernel at every step with
 * step % ljpairsupdatefreq == 0 (LJ if lj_on, bonds if is_assembly); pass
 * MADDY_RUN_SKIP_FIRST_REBUILD when the caller already called maddy_rebuild_* for
 * first_step (the reference rebuilds BEFORE its host events, compute_cuda.cu:1140-1151).
 * Observationally identical to n_steps x (rebuild?; force; integrate). Asynchronous. */
int maddy_run(maddy_handle *h, long long first_step, long long n_steps, unsigned flags);

/* ---- GTP schedule (folds the hydrolysis uploads into the fused window).
 * hydrolyse() (updater.cpp:229-257) reads only host flags that change at stride steps, so the host can evaluate
 * every hydrolysis event up to the next stride in advance (same rand() order) and hand the resulting GTP flags over
 * in one call: slot k (gtp_slots + k * n_tr_local * n_tot ints, same layout as maddy_upload_gtp) becomes the
 * current GTP state at the START of step first_event + k * period, exactly as if maddy_upload_gtp had been called
 * before that step.  A following maddy_run may then span those steps in ONE launch.  The schedule is consumed by
 * the runs that cover it; maddy_upload_gtp clears it.  n_slots = 0 clears it.  The call does not wait for the GPU: the
 * slots are copied out of gtp_slots before it returns and uploaded on a second stream into a buffer of their own, so it
 * may be issued while the previous window (and its schedule) is still running. */
int maddy_schedule_gtp(maddy_handle *h, long long first_event, long long period, int n_slots, const int *gtp_slots);

/* ---- hydrolyse() (updater.cpp:229-257) for ALL events of one output stride, on the device.
 * Inside a stride a dimer's eligibility changes only through hydrolysis itself (on-tubule and reserve flags change at
 * stride steps), so right after the stride block the events at first_event, first_event + period, ... (n_events of them,
 * none beyond the next stride step) are evaluated in one go: same draws in the same order as the reference's host loop
 * (dimer-outer / trajectory-inner, one rand() per eligible dimer, hydrolysis iff rand() / (double)RAND_MAX < 0.02, then
 * the return of off-tubule GDP dimers to GTP).  rand_window31: the 31 words of the caller's libc-rand()-compatible
 * generator (glibc TYPE_3), oldest first, such that the next draw is (w[0] + w[28]) >> 1; the device produces the stream
 * behind it by jump-ahead.  The on-tubule flags are the ones of the last two MADDY_SNAP_ONTUBULE classifications (before
 * the first: the flags passed to maddy_create).  The result becomes the handle's GTP schedule exactly as if
 * maddy_schedule_gtp had been called with the n_events states (a following maddy_run may span all of them).
 * Asynchronous; maddy_hydrolysis_result blocks until the counters have landed: draws_total = rand() calls the events
 * consumed (the caller advances its generator by that many, e.g. with maddy_rand_discard), event_first_draw[k] = draws
 * consumed before event k (may be NULL), gtp_slots ([n_events][n_tr_local * n_tot] ints, may be NULL; only available
 * when the plan was made with MADDY_HYD_KEEP_SLOTS) = GTP state after every event, for the caller's messages.
 * Requires one handle holding the whole ensemble (n_tr_local == n_tr) and an even n_tot.
 *
 * SHARDED ensembles (equal contiguous blocks, one handle each): the draw positions are global, so every shard evaluates
 * the plan of the WHOLE ensemble and keeps the slots of its own trajectories.  maddy_hydrolysis_inputs queues the
 * preparation of this shard's inputs (its GTP state and eligibility masks, transposed: [2][n_tot / 2][n_tr_local] bytes)
 * and returns the device buffer; the caller gathers the buffers of all shards in shard order into one device array on
 * every GPU (ncclAllGather, torch.distributed.all_gather_into_tensor, ... in stream order with the handle's stream) and
 * passes it to maddy_hydrolysis_plan_sharded.  maddy_hydrolysis_plan_all does all of that for handles living in this
 * process (NCCL).  draws_total of maddy_hydrolysis_result is the ensemble's, identical on every shard; gtp_slots holds
 * this shard's trajectories. */
#define MADDY_HYD_KEEP_SLOTS 1u
int maddy_hydrolysis_plan(maddy_handle *h, const unsigned *rand_window31, long long first_event, long long period, int n_events, unsigned flags);
int maddy_hydrolysis_inputs(maddy_handle *h, void **device_buffer, unsigned long *bytes);
int maddy_hydrolysis_plan_sharded(maddy_handle *h, const void *gathered_device, int n_shards, const unsigned *rand_window31, long long first_event,
                                  long long period, int n_events, unsigned flags);
int maddy_hydrolysis_plan_all(maddy_handle **handles, int n, const unsigned *rand_window31, long long first_event, long long period, int n_events,
                              unsigned flags);
int maddy_hydrolysis_result(maddy_handle *h, unsigned long long *draws_total, unsigned long long *event_first_draw, int *gtp_slots);
/* The scheduled GTP state of `step` (maddy_schedule_gtp / maddy_hydrolysis_plan), if there is one, becomes current NOW
 * (the fused loop applies slots at the start of their step; a caller that evaluates energies at that step before it
 * launches the window - the stride block, compute_cuda.cu:1153-1170 - needs them earlier).  No-op without such a slot. */
int maddy_apply_scheduled_gtp(maddy_handle *h, long long step);
/* window31 (oldest word first) of a glibc TYPE_3 generator advanced by n draws, in place.  Needs no GPU. */
void maddy_rand_discard(unsigned *window31, unsigned long long n);

/* ---- energies: energy_kernel + OutputAllEnergies (compute_cuda.cu:676-911,
 * updater.cpp:3-43).  out_per_traj: [n_tr_local][7] doubles (harm,long,lat,psi,fi,teta,lj);
 * out_per_monomer (may be NULL): [n_tr_local*n_tot][7] doubles in the reference's
 * `Energies` field order (U_harm,U_long,U_lat,U_psi,U_fi,U_teta,U_lj; mt.h:94-102). */
int maddy_energies(maddy_handle *h, double *out_per_traj, double *out_per_monomer);
/* The stride block of the reference loop rebuilds the lists and then evaluates the energies (compute_cuda.cu:1140-1170):
 * same as maddy_rebuild_lj + maddy_rebuild_bonds + maddy_energies, in ONE launch. */
int maddy_rebuild_and_energies(maddy_handle *h, double *out_per_traj, double *out_per_monomer);

/* ---- asynchronous stride snapshot.  What the reference loop reads back at a stride step (energies :1163-1170,
 * coordinates :1173, forces :1177) without stalling the stream: maddy_snapshot_begin enqueues [list rebuild +] energy
 * evaluation and device -> pinned-host copies behind the work already queued and returns at once; work queued afterwards
 * (the next maddy_run window) runs while the host waits in maddy_snapshot_end, which blocks only until THOSE copies
 * have landed and converts them into the caller's arrays (any of which may be NULL).  One snapshot in flight per handle. */
#define MADDY_SNAP_COORDS 1u
#define MADDY_SNAP_FORCES 2u
#define MADDY_SNAP_ENERGIES 4u
#define MADDY_SNAP_REBUILD 8u /* with ENERGIES: rebuild the lists first, as maddy_rebuild_and_energies does */
/* MADDY_SNAP_ONTUBULE: mt_length()'s classification (updater.cpp:154-227, the predicate of :161) of the state being
 * snapshotted, evaluated on the device AFTER the energies (which read the flags of the previous stride, as
 * compute_cuda.cu:1165 precedes :1186-1190).  Exact, not approximate: the radius test is IEEE float arithmetic and the
 * `cosf(theta) > cosf(ANG_THRES)` test is applied as intervals of |theta| whose end points were bisected with the
 * calling process's own cosf at maddy_create.  MADDY_SNAP_ONTUBULE_APPLY additionally makes the result the handle's
 * current on-tubule flags, as maddy_upload_on_tubule of the same values would (compute_cuda.cu:1190) - without the
 * host round trip.  Results: maddy_snapshot_tubule_lengths (counts only, available first) and
 * maddy_snapshot_on_tubule (flags, after maddy_snapshot_end). */
#define MADDY_SNAP_ONTUBULE 16u
#define MADDY_SNAP_ONTUBULE_APPLY 32u
#define MADDY_SNAP_GTP 64u /* the GTP flags as of the snapshot (maddy_snapshot_gtp after maddy_snapshot_end) */
/* With MADDY_SNAP_ONTUBULE: a caller that does not want to wait for the verdict may queue its hydrolysis plan and its next
 * window right behind the snapshot.  Should the classification turn out undecided (see maddy_snapshot_tubule_lengths),
 * every maddy_run window and maddy_hydrolysis_plan queued behind it on this handle returns at once WITHOUT touching the
 * state, until maddy_clear_guard(): the caller then classifies on the host, uploads, and queues the window again.
 * Nothing ever runs on a guessed flag. */
#define MADDY_SNAP_ONTUBULE_GUARD 128u
int maddy_snapshot_begin(maddy_handle *h, unsigned what);
int maddy_snapshot_end(maddy_handle *h, float *coords_aos7, float *forces_aos7, double *energies_per_traj);
/* Per-trajectory on-tubule counts (`sum` of updater.cpp:164-170) of the snapshot in flight: blocks only until the
 * (n_tr_local + 1)-int record at the head of the read-back has landed, i.e. a few microseconds behind the classification
 * kernel, long before the coordinates (also valid for the snapshot collected last).  *undecided != 0: some |theta| lies beyond the range the exact rule covers
 * (several turns away from zero); the caller must then classify on the host and upload (nothing is guessed). */
int maddy_snapshot_tubule_lengths(maddy_handle *h, int *mt_len, int *undecided);
/* 1 when the exact on-tubule rule could be derived from this process's cosf (it then is, for every libm seen so far);
 * 0: MADDY_SNAP_ONTUBULE is refused and the caller classifies on the host. */
int maddy_has_exact_on_tubule(const maddy_handle *h);
/* The rule itself (host-only call, no device needed; what the classification kernel is given): with rad =
 * sqrtf(x*x + y*y) in float and a = |theta|, a monomer is on the tubule iff rad < *rad_hi && (double)rad > 1.0 && a in
 * [0, e0) U (e1, e2) U (e3, e4) U (e5, e6); a >= *a_max (or NaN) is undecided.  edges[MADDY_ON_TUBULE_EDGES] are the float
 * values at which this process's `cosf(theta) > cosf(ANG_THRES)` (updater.cpp:166-168, mt.h:23-30) switches, found by
 * bisection over the float bit patterns.  MADDY_EINVAL when no such rule could be derived. */
#define MADDY_ON_TUBULE_EDGES 7
int maddy_on_tubule_rule(float *rad_hi, float *a_max, float *edges);
/* on_tubule_cur[n_tr_local * n_tot] / mt_len[n_tr_local] of the snapshot collected last by maddy_snapshot_end. */
int maddy_snapshot_on_tubule(maddy_handle *h, int *on_tubule_cur, int *mt_len);
int maddy_snapshot_gtp(maddy_handle *h, int *gtp);
int maddy_clear_guard(maddy_handle *h);
/* device-resident result of the last maddy_energies call: [n_tr_local][7] doubles */
void *maddy_energies_device(maddy_handle *h);

/* ---- state transfer (compute_cuda.cu:1157,1173,1177,1190,1202-1204) */
int maddy_download_coords(maddy_handle *h, float *coords_aos7);
int maddy_download_forces(maddy_handle *h, float *forces_aos7);
int maddy_upload_coords(maddy_handle *h, const float *coords_aos7); /* no angle wrapping (matches :1204) */
int maddy_upload_gtp(maddy_handle *h, const int *gtp);
int maddy_upload_on_tubule(maddy_handle *h, const int *on_tubule_cur);
int maddy_upload_extra(maddy_handle *h, const unsigned char *extra);
/* change_conc()'s insertions (updater.cpp:118-135) as a sparse update: for k < n_insert the reserve dimer
 * (index[k], index[k] + 1) - LOCAL monomer indices traj_local * n_tot + i - leaves the reserve; its first monomer is
 * placed at {xyzz[4k], xyzz[4k+1], xyzz[4k+2]}, its second at {xyzz[4k], xyzz[4k+1], xyzz[4k+3]} (angles untouched; the
 * caller evaluates z + 2 r_mon itself, updater.cpp:133).  Same device state afterwards as maddy_upload_extra +
 * maddy_upload_coords of the host arrays change_conc() modified (compute_cuda.cu:1202-1204), without moving the whole
 * ensemble over PCIe twice.  Asynchronous (the arrays are copied before the call returns). */
int maddy_insert_dimers(maddy_handle *h, int n_insert, const int *index, const float *xyzz);
/* lists in the reference layout [traj][i][capacity] / [traj][i], capacity =
 * max_longitudinal / max_lateral / MADDY_LJ_CAPACITY */
int maddy_download_list(maddy_handle *h, int kind, int *counts, int *entries);
int maddy_upload_list(maddy_handle *h, int kind, const int *counts, const int *entries);
/* RNG state: [2][n_tr_local*n_tot][4] uint32 (xyz streams then angular streams) */
int maddy_download_rng(maddy_handle *h, unsigned *state);
int maddy_upload_rng(maddy_handle *h, const unsigned *state);

/* list-maintenance statistics of the fused loop since creation / the last reset, summed over the shard's trajectories:
 * out[0] near-list refreshes forced by the displacement guard, out[1] candidate-list re-scans, out[2] rebuilds that fell
 * back to the all-pairs path, out[3] near-list overflows.  (The lists themselves are exact on every path.) */
int maddy_list_stats(maddy_handle *h, unsigned long long out[4], int reset);

/* ---- in-situ analysis: the reference's offline DCD post-processing tools as reductions over the state
 * already on the device, so a run needs no DCD round trip for them.  All of them read the CURRENT state of the handle.
 *
 * maddy_analysis_setup: per-monomer PDB labels of the structure (shared by all trajectories): chain index
 *   (chain - 'A', or -1 for atoms outside the protofilaments, e.g. the reserve chain 'X'), residue number and the
 *   second character of the atom name ('A' / 'B' of CA / CB); n_pf protofilaments (pf_number, disc.cpp:12).
 * maddy_analysis_reference: the state becomes the "previous frame" of the displacement statistics.
 * maddy_analysis_temperature: scripts/temp_calc/main.cpp:70-92 between the previous frame and the current state, then
 *   the current state becomes the previous frame.  sums[traj][8] = raw sums over the monomers
 *   {dx2+dy2+dz2, dfi2+dpsi2+dtheta2 - 2 dfi dtheta cos2(psi), dx2, dy2, dz2, dfi2, dpsi2, dtheta2} (the tool then scales
 *   them with constants of its own, main.cpp:94-106).
 * maddy_analysis_project: scripts/disas_speed/3d22d.cpp:42-51: out[traj][i] = {sqrtf(x*x + y*y), z, theta}.
 * maddy_analysis_protofilaments: scripts/disas_speed/disc.cpp:62-117 on that projection: out[traj][pf] =
 *   {pf_end_number, curled_start, mt_end_number}: the dimer where the protofilament breaks off (radial-axial gap
 *   > 5 nm between consecutive dimers), where its curl starts (theta > 0.2) and the dimer number of its straight tip. */
int maddy_analysis_setup(maddy_handle *h, const int *chain, const int *resid, const char *name1, int n_pf);
int maddy_analysis_reference(maddy_handle *h);
int maddy_analysis_temperature(maddy_handle *h, double *sums);
int maddy_analysis_project(maddy_handle *h, float *out);
int maddy_analysis_protofilaments(maddy_handle *h, int *out);

/* ---- host-side pieces of the path that need no GPU (usable without a device) */
/* generateSeeds (HybridTaus.cu:32-48) on a FRESH ran2 state: fills seeds[np*4]. */
void maddy_generate_seeds(unsigned *seeds, int rseed, long long np);
/* TEA beta from the per-trajectory epsilon sum (bdhitea.cu:79-113). Returns 0, or
 * MADDY_ETEA when the reference would exit(-1). */
int maddy_tea_beta(double epsilon_sum, int n_noextra, int capricious, float tea_a, float epsmax,
                   float *beta_out, double *epsilon_out);

/* ---- multi-GPU ensemble statistics (new; NCCL over NVLink).  handles[n] live on n
 * different devices of this process; values[g] points to `count` doubles on the HOST per
 * handle; on return every values[g] holds the element-wise sum over handles.  The
 * reduction itself runs on the GPUs with ncclAllReduce. */
int maddy_ensemble_allreduce(maddy_handle **handles, int n, double **values, int count);

/* Device-resident form for the stride block of the step loop: each handle reduces the per-trajectory energies it
 * evaluated last (maddy_energies, maddy_rebuild_and_energies or maddy_snapshot_begin with MADDY_SNAP_ENERGIES) to
 * MADDY_ENSEMBLE_STATS doubles {sum of each of the 7 terms, sum of squares of each, trajectory count, 0} on its own
 * stream, the records are summed over the handles with ncclAllReduce in stream order (n == 1: no collective), and the
 * result is copied to pinned memory beside the work queued next.  _begin returns at once; _end blocks until the copy
 * has landed and returns the all-reduced record (identical on every handle).  Replaces nothing in the reference (its MPI
 * ranks never exchange data, main.cpp:27-50); it is the periodic ensemble reduction of SURVEY 8e. */
#define MADDY_ENSEMBLE_STATS 16
int maddy_ensemble_stats_begin(maddy_handle **handles, int n);
int maddy_ensemble_stats_end(maddy_handle **handles, int n, double *out16);

/* number of kernels this handle has launched since creation (bench bookkeeping) */
long long maddy_launch_count(const maddy_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* MADDY_B200_H_ */
