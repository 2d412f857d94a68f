/*
 * maddy_kernels.cuh — kernel argument structures shared by the kernels and the C-ABI host code.
 */
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "maddy_b200.h"
#include "maddy_device.cuh"

namespace maddy {

#define MD_MAX_THREADS 576 // threads per trajectory CTA (18 warps)
#define MD_MAX_MPT 6
#define MD_RUN_THREADS 512 // fused loop: 16 warps per CTA (registers are allocated per 4-warp group)
#define MD_RUN_MAX_MPT 7

// operation mask of the trajectory kernel
enum : unsigned {
    OP_REBUILD_LJ = 1u,
    OP_REBUILD_BONDS = 2u,
    OP_FORCE = 4u,   // evaluate forces and store them (step-granular maddy_force)
    OP_ENERGY = 8u,  // per-monomer energies + per-trajectory reduction
    OP_RUN = 16u,    // fused multi-step loop
    OP_TEA_EPS = 32u, // TEA epsilon / C_i statistics (integrateTea_epsilon_unlisted)
    OP_MATERIALISE = 64u, // write the exact Verlet list of the last (lazy) list-update step of the fused loop
    OP_TEA_PREP = 128u    // with OP_FORCE: also do integrateTea_prepare for the bead (TEA windows of maddy_run: one launch fewer per step)
};

// status bits written by kernels
enum : int { ST_LJ_OVERFLOW = 1, ST_LONG_OVERFLOW = 2, ST_LAT_OVERFLOW = 4 };

/*
 * Data layout in HBM (per handle; n = n_tr_local * N, Npad = N rounded up to 32):
 *   pos[n], ang[n]        float4 {x,y,z,0}, {fi,psi,theta,0}      coalesced 16-B loads
 *   fpos[n], fang[n]      float4 forces (step-granular API only; the fused loop keeps them in registers)
 *   rng_xyz[n], rng_ang[n] uint4 HybridTaus state (global stream ids, see maddy_b200.h)
 *   bl[traj][capLong+capLat][Npad]  uint16 bond codes (j<<1 | negative-sign), k-major so
 *                                    that thread i and i+1 read adjacent entries
 *   bcnt[traj][2][Npad]   uint8 counts (longitudinal, lateral)
 *   lj[traj][256][Npad]   uint16 neighbour index, k-major; ljcnt[traj][Npad] uint16
 *   extra/gtp/ontub[n]    uint8 flags
 */
struct DevSys {
    int N, Npad, ntr, maxH, capLong, capLat;
    float4 *pos, *ang, *fpos, *fang;
    uint4 *rng_xyz, *rng_ang;
    const int *harm;       // [N*maxH] signed, reference encoding
    const int *harm_count; // [N]
    const uint8_t *sflags; // [N] bit0 fixed, bits1.. mon_type
    const uint16_t *amap;  // [n_active] monomers that own a thread in the fused loop (all non-fixed ones)
    const uint16_t *fmap;  // [n_fixed]  fixed monomers looked after by threads 0..n_fixed-1
    int n_active, n_fixed;
    uint8_t *extra, *gtp, *ontub;
    const uint8_t *gtp_sched; // scheduled GTP states (maddy_schedule_gtp), or nullptr
    uint16_t *bl;
    uint8_t *bcnt;
    uint16_t *lj;
    uint16_t *ljcnt;
    double *en_mono; // [n][7]
    double *en_traj; // [ntr][7]
    int *status;
    int *guard; // non-zero: fused windows and hydrolysis plans queued behind return at once (MADDY_SNAP_ONTUBULE_GUARD)
    unsigned long long *stats; // [4] list-maintenance events: near refresh on guard trip, candidate re-scan, all-pairs fallback, near overflow
    uint16_t *cand;    // [ntr][MD_CAND_CAPACITY][Npad] candidate list, k-major
    uint16_t *candcnt; // [ntr][Npad]
    uint16_t *ncand;   // [ntr][MD_NCAND_CAPACITY][Npad] the candidates within MD_NEAR_R + MD_CAND_SKIN (subset of cand, same order)
    uint8_t *ncandcnt; // [ntr][Npad]; MD_NEAR_FULL = more than the capacity: use cand
    float4 *rpos;      // [n] positions at the last list-update step of the fused loop (what the Verlet list refers to)
    int *lj_stale;     // [ntr] 1: lj/ljcnt predate that step; the exact rows follow from rpos + cand (materialise_row)
    float4 *cpos;      // [n] positions when the candidate list was built (.w unused)
    int *cand_valid;   // [ntr] 1 = cand/candcnt/cpos describe the current extra flags
    // TEA (bdhitea): per-bead sum of squared tensor rows (d_ci), per-bead epsilon sums, per-traj beta
    float4 *tea_ci;
    float *tea_eps;
    float *tea_beta;
    float4 *tea_co, *tea_mf, *tea_rf; // per-step snapshots: coordinates (+extra in .w), molecular force, random force
    unsigned *wbar;    // [ntr][4] wide path, persistent window: barrier tickets + displacement-guard words of each trajectory
    float4 *tea_part;  // [ntr][segments][N] partial sums of the partner segments (long trajectories only)
    unsigned *tea_cnt; // [ntr][bead blocks] tickets of the segment CTAs (self-resetting)
    float4 *gstage; // wide path only (maddy_wide.cuh): [2][4][ntr*N] stage in HBM, or nullptr
    void *wgrid;    // wide path only: per-trajectory cell grid of the list rebuild (see wcells_at), or nullptr
};

// Near list (shared memory, fused loop only): all j != i with centre distance < MD_NEAR_R at the last
// rebuild / refresh, ascending j, bit 15 = "is in the reference LJ list".  It caches the part of the
// 15-nm Verlet list that can reach the 6-nm force cut-off before any monomer has moved MD_NEAR_GUARD:
// a pair at >= 7.0 needs a relative displacement > 1.0 to get inside 6.0, i.e. one endpoint > 0.5.
// The guard is checked every step (folded into the step barrier); a trip refreshes the cache from the
// full list in HBM, so forces are bit-identical to walking the full list.
#define MD_NEAR_R2 49.0f
#define MD_NEAR_GUARD2 0.2401f // 0.49^2 (< 0.5^2: margin for float rounding)
#define MD_NEAR_LJ_FLAG 0x8000u
#define MD_TILE 8              // consecutive monomers per culling tile (bounding boxes rebuilt with the lists)
// Candidate list (HBM): all j != i within (cut-off + MD_CAND_SKIN) when it was built.  The O(N^2)
// tile-culled scan only runs when some monomer has moved more than MD_CAND_SKIN / 2 since then
// (checked at every list-update step); otherwise the reference's Verlet list is obtained by
// re-testing the ~60 candidates with the exact cut-off test — identical output, ~30x less work.
#define MD_CAND_SKIN 1.5f
#define MD_CAND_GUARD2 0.5476f // 0.74^2 (< (MD_CAND_SKIN/2)^2)
#define MD_CAND_CAPACITY 320
#define MD_NCAND_CAPACITY 64
#define MD_NCAND_R2 72.25f // (MD_NEAR_R + MD_CAND_SKIN)^2 = 8.5^2
#define MD_NEAR_FULL 255 // near.cnt value of a monomer whose near list overflowed: it walks its full Verlet list instead
#define MD_NEAR_FULL_EXACT 254 // ... of a monomer escalated by the displacement guard: it walks its EXACT Verlet row (materialised if the list was lazy)
#define MD_NEAR_PENDING 253    // ... marked for escalation by a partner that tripped the guard in this step
#define MD_FILTER_BATCH 8 // candidate indices fetched per round trip in filter_candidates

// Per-run constants evaluated ONCE on the device (consts_kernel) with the same fast-math float expressions the
// reference evaluates in every thread of every step (approximate division / sqrt), then passed by value.
struct StepConsts {
    float aR;         // dt / gammaR
    float aA;         // dt / (gammaTheta * alpha)
    float aT;         // dt / gammaTheta
    float vA;         // varTheta * sqrt(freeze_temp / alpha)
    float D_lat_seam; // D_lat / seam_coeff
};

struct KArgs {
    maddy_params p;
    StepConsts c;
    int barr_long_on, barr_lat_on; // barrier term present AND its amplitude non-zero (a zero amplitude adds exactly -0)
    DevSys a;
    long long first_step, n_steps;
    // GTP schedule: slot k of gtp_sched ([n_slots][n] bytes) becomes current at the start of step sched_first + k*sched_period
    long long sched_first, sched_period;
    int sched_slots;
    unsigned ops;
    unsigned run_flags;
    int nbuf;          // 1 or 2 shared-memory stage buffers
    int near_cap;      // rows of the shared-memory near list (0: fast path disabled)
    int rng_smem_offset; // byte offset of the shared-memory RNG area (two-CTA shape only)
    int topo_smem_offset; // byte offset of the packed per-monomer topology words, or -1
    int lazy;          // fused loop: list-update steps only refresh the near/bond lists; the exact Verlet list is materialised on demand
    int near_all_listed; // LJ pairs cut-off > near radius: every near-list entry is LJ-listed (no per-entry flag test in the force loop)
    float band_mid, band_hw; // |sf - band_mid| <= band_hw covers [cut_force.lo, cut_force.hi]: pairs that need the fp64 tie-break
    float rcand2;      // squared candidate radius: (max(LJ pairs cut-off, near radius) + MD_CAND_SKIN)^2
    CutTest cut_pairs; // LJ list cut-off (ljpairscutoff)
    CutTest cut_force; // LJ force cut-off (6.0)
};

// exact on-tubule classification on the device (ontubule_kernel, maddy_analysis.cu): thresholds bisected on the host
#define ONTUB_EDGES 7
struct OnTubRule {
    float rad_hi;            // R_MT + R_THRES (mt.h:23,30)
    float a_max;             // |theta| at or beyond this is reported as undecided
    float edge[ONTUB_EDGES]; // on iff |theta| in [0, e0) U (e1, e2) U (e3, e4) U (e5, e6)
};

// hydrolysis events of one stride on the device (maddy_events.cu)
struct HydArgs {
    const uint8_t *gtp, *extra, *cur, *prev; // [ntr * N]: GTP state at the stride, reserve flags, on-tubule flags now / previous stride
    uint8_t *own;                            // [2][nd][ntr_l] this shard's transposed inputs: GTP state, then static mask (hyd_prepare_kernel)
    uint8_t *all;                            // [shards][2][nd][ntr_l] the inputs of EVERY shard (== own for one shard); the plan works on this copy
    int ntr_l, shards, shard;                // trajectories per shard, number of shards, this shard: ntr = shards * ntr_l is the GLOBAL count
    int seg, nseg, nrows;                    // a dimer row is cut into nseg segments of seg trajectories: nrows = nd * nseg units of work (one warp each)
    unsigned *rowcount;                      // [nrows] draws of the current event per row segment
    unsigned long long *rowstart;            // [nrows] index of each segment's first draw in the plan's stream
    unsigned long long *cursor;              // [1] draws consumed so far by the plan
    unsigned long long *event_start;         // [n_events] first draw of every event
    const uint32_t *stream;                  // the plan's draws (rand() values), stream_count of them
    unsigned long long stream_count;
    unsigned threshold;                      // largest rand() value v with v / (double)RAND_MAX < 0.02
    int *status;                             // bit 0: the stream was too short (cannot happen: sized for the worst case)
    const int *guard;                        // see DevSys::guard
    int N, ntr, nd;
};

// in-situ analysis (maddy_analysis.cu)
struct AnalysisArgs {
    const float4 *pos, *ang;   // current state
    float4 *ppos, *pang;       // previous frame of the displacement statistics
    const short *chain, *resid; // [N] PDB labels (chain - 'A' or -1, residue number)
    const char *name1;         // [N] second character of the atom name
    int N, ntr, n_pf;
    double *temp; // [ntr][8]
    float *proj;  // [ntr*N][3]
    int *pf;      // [ntr][n_pf][3]
};

} // namespace maddy
