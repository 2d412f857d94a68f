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

// operation mask of the trajectory kernel
enum : unsigned {
    OP_REBUILD_LJ = 1u,
    OP_REBUILD_BONDS = 2u,
    OP_FORCE = 4u,   // evaluate forces and store them (step-granular maddy_force)
    OP_ENERGY = 8u,  // per-monomer energies + per-trajectory reduction
    OP_RUN = 16u,    // fused multi-step loop
    OP_TEA_EPS = 32u // TEA epsilon / C_i statistics (integrateTea_epsilon_unlisted)
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
    uint8_t *extra, *gtp, *ontub;
    uint16_t *bl;
    uint8_t *bcnt;
    uint16_t *lj;
    uint16_t *ljcnt;
    double *en_mono; // [n][7]
    double *en_traj; // [ntr][7]
    int *status;
    // TEA (bdhitea): per-bead sum of squared tensor rows (d_ci), per-bead epsilon sums, per-traj beta
    float4 *tea_ci;
    float *tea_eps;
    float *tea_beta;
};

struct KArgs {
    maddy_params p;
    DevSys a;
    long long first_step, n_steps;
    unsigned ops;
    unsigned run_flags;
    int nbuf;          // 1 or 2 shared-memory stage buffers
    CutTest cut_pairs; // LJ list cut-off (ljpairscutoff)
    CutTest cut_force; // LJ force cut-off (6.0)
};

} // namespace maddy
