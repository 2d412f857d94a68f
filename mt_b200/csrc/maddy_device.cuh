/*
 * maddy_device.cuh — device-side physics of the MADDY Langevin/BD step, written for sm_100a.
 *
 * Everything here is a from-scratch formulation of WHAT the reference kernels compute
 * (src/compute_cuda.cu, src/HybridTaus.cu), organised around one idea the reference does
 * not use: every monomer's orientation frame (e1,e2,e3 = columns of Rz(psi)Ry(theta)Rx(fi))
 * is evaluated ONCE per step by its owning thread and the three site-offset vectors every
 * neighbour needs (r_mon*e3, R*p1, R*p2) are staged in shared memory, so a bonded
 * neighbour costs three LDS.128 instead of six MUFU + ~60 FMAs.
 *
 * Arithmetic contract: float state; MUFU sin/cos/ex2/lg2 exactly where the
 * reference's -use_fast_math build uses them; the squared site distance is summed in fp64
 * from float components and rounded to float before the (approximate) sqrt, as the
 * reference's `sqrtf(pow(..,2)+pow(..,2)+pow(..,2))` does; cut-off decisions on the LJ
 * lists are taken on the exact fp64 sum against a precomputed fp64 threshold that is
 * equivalent to the reference's `float(sqrt(double)) < cutoff` test.
 *
 * This translation unit is compiled with -use_fast_math so that sinf/cosf/expf/logf/sqrtf
 * and `/` lower to the same MUFU sequences as in the reference build.
 */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "maddy_b200.h"

namespace maddy {

// geometry constants of the model (reference src/mt.h:23-50)
#define MD_R_MT 8.12f
#define MD_R_MON 2.0f
#define MD_ANGLE_CUTOFF 1.0f
#define MD_PAIR_CUTOFF 2.5f
#define MD_LJ_FORCE_CUTOFF 6.0f
// centre-distance prefilter for bond candidates: every interaction site lies within
// r_mon (=|p1|=|p2| up to rounding) of its monomer centre, so a site distance < 2.5
// needs a centre distance < 2.5 + 2*2.0; 6.6^2 leaves a 0.1 nm margin.
#define MD_BOND_PREFILTER2 43.56f

struct F3 { float x, y, z; };
__device__ __forceinline__ F3 mk3(float x, float y, float z) { F3 r; r.x = x; r.y = y; r.z = z; return r; }
// Arithmetic in the kernels that share these functions is pinned: maddy_kernels.cu is compiled with -fmad=false and
// every fused multiply-add is written out as fmaf(), so the fused loop and the step-granular kernels (different
// instantiations, different register budgets) round identically and stay bit-for-bit interchangeable.
__device__ __forceinline__ float dot3(const F3 &a, const F3 &b) { return fmaf(a.z, b.z, fmaf(a.y, b.y, a.x * b.x)); }

// generalized force / coordinate of one monomer: x,y,z,fi,psi,theta
struct G6 { float x, y, z, fi, psi, theta; };

// lateral site local coordinates.  The reference evaluates these float expressions on the
// device at run time under fast-math (MUFU sin/cos of 2*pi/13, see its PTX), so we do the
// same: p1 = (xp, yp, zp), p2 = (xp, -yp, -zp)   (src/mt.h:44-50)
struct LatSite { float xp, yp, zp; };
__device__ __forceinline__ LatSite lateral_site()
{
    LatSite s;
    const float a = (float)(2.0f * M_PI / 13.0f);
    s.xp = 0.5f * MD_R_MT * (cosf(a) - 1.0f);
    s.yp = 0.5f * MD_R_MT * sinf(a);
    s.zp = -3.0f * MD_R_MON / 13.0f;
    return s;
}

// What the force evaluation needs of a monomer's orientation R = Rz(psi) Ry(theta) Rx(fi): the body x axis
// e1 = R x^ and (sin psi, cos psi).  Every derivative of a site offset o = R p is a cross product with the
// instantaneous rotation axis of the angle:   d o/d fi = e1 x o,   d o/d psi = z^ x o,   d o/d theta = (-sp, cp, 0) x o,
// so no other trigonometric value has to stay alive between publish() and the bond loops (five registers).
struct Frame {
    float e1x, e1y, e1z, sp, cp;
};

// Evaluates the frame once per step: the six MUFU sin/cos, the site offsets staged for neighbours
// (a = r_mon*e3: "end" sites are +-a;  l1 = R*p1, l2 = R*p2) and the five values the force loops keep.
__device__ __forceinline__ Frame make_frame(float fi, float psi, float theta, const LatSite &ls, F3 &a, F3 &l1, F3 &l2)
{
    const float sf = sinf(fi), cf = cosf(fi);
    const float sp = sinf(psi), cp = cosf(psi);
    const float st = sinf(theta), ct = cosf(theta);
    const F3 e1 = mk3(cp * ct, sp * ct, -st);
    const F3 e2 = mk3(fmaf(cp * sf, st, -(cf * sp)), fmaf(sf * sp, st, cf * cp), ct * sf);
    const F3 e3 = mk3(fmaf(cf * cp, st, sf * sp), fmaf(cf * sp, st, -(cp * sf)), cf * ct);
    a = mk3(MD_R_MON * e3.x, MD_R_MON * e3.y, MD_R_MON * e3.z);
    // u = xp*e1, v = yp*e2 + zp*e3 ; l1 = u + v, l2 = u - v
    const F3 u = mk3(ls.xp * e1.x, ls.xp * e1.y, ls.xp * e1.z);
    const F3 v = mk3(fmaf(ls.yp, e2.x, ls.zp * e3.x), fmaf(ls.yp, e2.y, ls.zp * e3.y), fmaf(ls.yp, e2.z, ls.zp * e3.z));
    l1 = mk3(u.x + v.x, u.y + v.y, u.z + v.z);
    l2 = mk3(u.x - v.x, u.y - v.y, u.z - v.z);
    Frame f;
    f.e1x = e1.x; f.e1y = e1.y; f.e1z = e1.z; f.sp = sp; f.cp = cp;
    // keep these five in registers: under pressure the compiler would otherwise re-issue the MUFU sin/cos at every
    // use (the XU pipe is the scarcest one in this kernel)
    asm volatile("" : "+f"(f.e1x), "+f"(f.e1y), "+f"(f.e1z), "+f"(f.sp), "+f"(f.cp));
    return f;
}

// |d| as the reference forms it: fp64 sum of squares (order z,x,y), rounded to float,
// approximate float sqrt.  (compute_cuda.cu:89-96, :208-215, :337-350)
__device__ __forceinline__ float site_distance(const F3 &d)
{
    // (products of two floats are exact in fp64, so fma() rounds exactly like the reference's multiply-then-add)
    double s = (double)d.z * (double)d.z;
    s = fma((double)d.x, (double)d.x, s);
    s = fma((double)d.y, (double)d.y, s);
    return sqrtf((float)s);
}
// variant used by energy_kernel: fp64 sqrt, then rounded to float (compute_cuda.cu:731,:791,:858)
__device__ __forceinline__ float site_distance_d(const F3 &d)
{
    double s = (double)d.z * (double)d.z;
    s += (double)d.x * (double)d.x;
    s += (double)d.y * (double)d.y;
    return (float)sqrt(s);
}

// potentials (compute_cuda.cu:16-30)
__device__ __forceinline__ float dmorse(float D, float a, float x)
{
    float e = expf(-a * x);
    return 2 * a * D * (1 - e) * e;
}
__device__ __forceinline__ float morse_en(float D, float a, float x)
{
    float e = expf(-a * x);
    return D * (1 - e) * (1 - e) - D;
}
__device__ __forceinline__ float dbarr(float a, float r, float w, float x)
{
    return -a * expf(-(x - r) * (x - r) / (2 * w * w)) * (x - r) / (w * w);
}
__device__ __forceinline__ float barr(float a, float r, float w, float x)
{
    return a * expf(-(x - r) * (x - r) / (2 * w * w));
}

// Accumulate the generalized force of one bond on monomer i.
//   d = site_j - site_i,  k = U'(dr)/dr:   F_xyz += k d,   F_q += k d . d(site_i)/dq,   o = own site offset R_i p
__device__ __forceinline__ void bond_accumulate(G6 &f, float k, const F3 &d, const F3 &o, const Frame &fr)
{
    f.x = fmaf(k, d.x, f.x);
    f.y = fmaf(k, d.y, f.y);
    f.z = fmaf(k, d.z, f.z);
    // d . (e1 x o)
    const float tfi = fmaf(d.z, fmaf(fr.e1x, o.y, -(fr.e1y * o.x)),
                           fmaf(d.y, fmaf(fr.e1z, o.x, -(fr.e1x * o.z)), d.x * fmaf(fr.e1y, o.z, -(fr.e1z * o.y))));
    // d . ((-sp, cp, 0) x o) = o.z (cp d.x + sp d.y) - d.z (sp o.y + cp o.x)
    const float tth = fmaf(o.z, fmaf(fr.cp, d.x, fr.sp * d.y), -(d.z * fmaf(fr.sp, o.y, fr.cp * o.x)));
    f.fi = fmaf(k, tfi, f.fi);
    f.psi = fmaf(k, fmaf(d.y, o.x, -(d.x * o.y)), f.psi);
    f.theta = fmaf(k, tth, f.theta);
}

// ---------------------------------------------------------------- HybridTaus RNG
// Integer stream of GPU Gems 3 ch.37 as used by the reference (HybridTaus.cu:63-76):
// three Tausworthe steps and one LCG step XOR-ed; state = uint4.
__device__ __forceinline__ unsigned taus_step(unsigned &z, int s1, int s2, int s3, unsigned m)
{
    unsigned b = (((z << s1) ^ z) >> s2);
    return z = (((z & m) << s3) ^ b);
}
__device__ __forceinline__ unsigned hybrid_taus(uint4 &s)
{
    unsigned r = taus_step(s.x, 13, 19, 12, 4294967294u);
    r ^= taus_step(s.y, 2, 25, 4, 4294967288u);
    r ^= taus_step(s.z, 3, 11, 17, 4294967280u);
    s.w = 1664525u * s.w + 1013904223u;
    return r ^ s.w;
}
// low 23 bits as the mantissa of a float in [1,2), minus 1; 0 -> 1e-8 (HybridTaus.cu:53-61)
__device__ __forceinline__ float uint_to_unit_float(unsigned u)
{
    // r is 0 or >= 2^-23 > 1e-8, so the reference's `r == 0 ? 1e-8 : r` is a maximum
    const float r = __uint_as_float(0x3f800000u | (0x007fffffu & u)) - 1.0f;
    return fmaxf(r, 1.0e-8f);
}
// Two Box-Muller pairs from four draws; .w is produced and discarded by the callers, as in
// the reference (HybridTaus.cu:85-98): r = sqrtf(-2 logf(u1)), angle = float(2*pi (double) * u2).
__device__ __forceinline__ float4 rforce(uint4 &s)
{
    float4 n;
    float r = sqrtf(-2.0f * logf(uint_to_unit_float(hybrid_taus(s))));
    float t = 2.0f * M_PI * uint_to_unit_float(hybrid_taus(s));
    n.x = r * __sinf(t);
    n.y = r * __cosf(t);
    r = sqrtf(-2.0f * logf(uint_to_unit_float(hybrid_taus(s))));
    t = 2.0f * M_PI * uint_to_unit_float(hybrid_taus(s));
    n.z = r * __sinf(t);
    n.w = r * __cosf(t);
    return n;
}

// exact fp64 squared distance of float differences (order x,y,z as LJ_kernel, compute_cuda.cu:929-931)
__device__ __forceinline__ double dist2_exact(float dx, float dy, float dz)
{
    double s = (double)dx * (double)dx;
    s += (double)dy * (double)dy;
    s += (double)dz * (double)dz;
    return s;
}

// Cut-off test equivalent to `float(sqrt(double s)) < c`: thr.t is the smallest double s for which
// the reference's test fails; lo/hi bracket it in float so fp64 is only touched inside the band.
struct CutTest { double t; float lo, hi; };
__device__ __forceinline__ bool inside_cut(const CutTest &c, float dx, float dy, float dz, float sf)
{
    if (sf < c.lo) return true;
    if (sf > c.hi) return false;
    return dist2_exact(dx, dy, dz) < c.t;
}

} // namespace maddy
