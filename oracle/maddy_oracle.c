/*
 * maddy_oracle.c — CPU restatement of the MADDY Langevin/BD step.  TEST INFRASTRUCTURE ONLY
 * (see maddy_oracle.h for who may use it and for the parity status).
 *
 * Written independently of the CUDA kernels: where the kernels use the closed-form frame
 * identities (d e2/d fi = e3, ...), this file builds R = Rz(psi) Ry(theta) Rx(fi) and its three
 * angle derivatives from elementary 3x3 matrices and multiplies them out numerically, so an
 * algebra slip on either side shows up as a parity failure.
 *
 * Each function cites the reference lines it follows (paths under /root/reference/src).
 */
#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "maddy_oracle.h"

#define R_MT 8.12f
#define R_MON 2.0f
#define KB 0.0019872041f
#define ANGLE_CUTOFF 1.0f
#define PAIR_CUTOFF 2.5f
#define ZERO MADDY_ZERO_SENTINEL

/* --------------------------------------------------------------------------------------------
 * ran2 (ran2.h:18-56) called with a positive seed: the initialisation branch never runs, so the
 * generator starts from idum2 = 123456789, iy = 0 and an all-zero shuffle table.
 * generateSeeds (HybridTaus.cu:32-48): four components per state, each re-drawn while < 128. */
typedef struct {
    int idum, idum2, iy, iv[32];
} ran2_state;

static double ran2_next(ran2_state *s)
{
    const int IM1 = 2147483563, IM2 = 2147483399, IA1 = 40014, IA2 = 40692, IQ1 = 53668, IQ2 = 52774, IR1 = 12211, IR2 = 3791;
    const int IMM1 = IM1 - 1, NDIV = 1 + IMM1 / 32;
    int k = s->idum / IQ1;
    s->idum = IA1 * (s->idum - k * IQ1) - k * IR1;
    if (s->idum < 0) s->idum += IM1;
    k = s->idum2 / IQ2;
    s->idum2 = IA2 * (s->idum2 - k * IQ2) - k * IR2;
    if (s->idum2 < 0) s->idum2 += IM2;
    int j = s->iy / NDIV;
    s->iy = s->iv[j] - s->idum2;
    s->iv[j] = s->idum;
    if (s->iy < 1) s->iy += IMM1;
    double t = (1.0 / IM1) * s->iy;
    return t > (1.0 - 1.2e-7) ? (1.0 - 1.2e-7) : t;
}

void oracle_generate_seeds(unsigned *seeds, int rseed, long long np)
{
    ran2_state s;
    memset(&s, 0, sizeof s);
    s.idum = rseed;
    s.idum2 = 123456789;
    for (long long q = 0; q < 4 * np; q++) {
        do {
            seeds[q] = (unsigned)(ran2_next(&s) * UINT_MAX);
        } while (seeds[q] < 128);
    }
}

/* HybridTaus.cu:63-76 */
static unsigned taus(unsigned *z, int s1, int s2, int s3, unsigned m)
{
    unsigned b = (((*z << s1) ^ *z) >> s2);
    return *z = (((*z & m) << s3) ^ b);
}
unsigned oracle_hybrid_taus(unsigned *st)
{
    unsigned a = taus(&st[0], 13, 19, 12, 4294967294u);
    unsigned b = taus(&st[1], 2, 25, 4, 4294967288u);
    unsigned c = taus(&st[2], 3, 11, 17, 4294967280u);
    st[3] = 1664525u * st[3] + 1013904223u;
    return a ^ b ^ c ^ st[3];
}
/* HybridTaus.cu:53-61 */
static float uint_to_float(unsigned u)
{
    union { unsigned i; float f; } v;
    v.i = 0x3f800000u | (0x007fffffu & u);
    float r = v.f - 1.0f;
    return r == 0 ? 1.0e-8f : r;
}
/* HybridTaus.cu:85-98 (libm in place of the MUFU intrinsics) */
void oracle_rforce(unsigned *st, float *out)
{
    float r = sqrtf(-2.0f * logf(uint_to_float(oracle_hybrid_taus(st))));
    float th = (float)(2.0f * M_PI * uint_to_float(oracle_hybrid_taus(st)));
    out[0] = r * sinf(th);
    out[1] = r * cosf(th);
    r = sqrtf(-2.0f * logf(uint_to_float(oracle_hybrid_taus(st))));
    th = (float)(2.0f * M_PI * uint_to_float(oracle_hybrid_taus(st)));
    out[2] = r * sinf(th);
    out[3] = r * cosf(th);
}

/* -------------------------------------------------------------------------------------------- geometry */
typedef struct { float m[3][3]; } mat3;

static mat3 mul3(mat3 a, mat3 b)
{
    mat3 r;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return r;
}
static mat3 rot_x(float a, int deriv)
{
    float s = sinf(a), c = cosf(a);
    mat3 r = {{{1, 0, 0}, {0, c, -s}, {0, s, c}}};
    mat3 d = {{{0, 0, 0}, {0, -s, -c}, {0, c, -s}}};
    return deriv ? d : r;
}
static mat3 rot_y(float a, int deriv)
{
    float s = sinf(a), c = cosf(a);
    mat3 r = {{{c, 0, s}, {0, 1, 0}, {-s, 0, c}}};
    mat3 d = {{{-s, 0, c}, {0, 0, 0}, {-c, 0, -s}}};
    return deriv ? d : r;
}
static mat3 rot_z(float a, int deriv)
{
    float s = sinf(a), c = cosf(a);
    mat3 r = {{{c, -s, 0}, {s, c, 0}, {0, 0, 1}}};
    mat3 d = {{{-s, -c, 0}, {c, -s, 0}, {0, 0, 0}}};
    return deriv ? d : r;
}
/* R = Rz(psi) Ry(theta) Rx(fi)   (read off compute_cuda.cu:337-350) */
typedef struct { mat3 R, dfi, dpsi, dtheta; } frame;
static frame make_frame(const float *c /* AoS7 */)
{
    const float fi = c[3], theta = c[4], psi = c[5];
    frame f;
    f.R = mul3(mul3(rot_z(psi, 0), rot_y(theta, 0)), rot_x(fi, 0));
    f.dfi = mul3(mul3(rot_z(psi, 0), rot_y(theta, 0)), rot_x(fi, 1));
    f.dpsi = mul3(mul3(rot_z(psi, 1), rot_y(theta, 0)), rot_x(fi, 0));
    f.dtheta = mul3(mul3(rot_z(psi, 0), rot_y(theta, 1)), rot_x(fi, 0));
    return f;
}
static void apply(const mat3 *m, const float p[3], float out[3])
{
    for (int i = 0; i < 3; i++) out[i] = m->m[i][0] * p[0] + m->m[i][1] * p[1] + m->m[i][2] * p[2];
}
/* lateral site local coordinates (mt.h:44-50): p1, p2 */
static void lateral_points(float p1[3], float p2[3])
{
    const float a = (float)(2.0f * M_PI / 13.0f);
    p1[0] = 0.5f * R_MT * (cosf(a) - 1.0f);
    p1[1] = 0.5f * R_MT * sinf(a);
    p1[2] = -3.0f * R_MON / 13.0f;
    p2[0] = p1[0];
    p2[1] = -p1[1];
    p2[2] = -p1[2];
}
/* d = site_j - site_i with site = r + R p; squared length summed in double as the reference's
 * pow(.,2)+pow(.,2)+pow(.,2) does (order z, x, y: compute_cuda.cu:89-96) */
static double site_delta(const float *ci, const frame *fi, const float pi[3], const float *cj, const frame *fj, const float pj[3],
                         float d[3])
{
    float oi[3], oj[3];
    apply(&fi->R, pi, oi);
    apply(&fj->R, pj, oj);
    for (int k = 0; k < 3; k++) d[k] = (cj[k] - ci[k]) - oi[k] + oj[k];
    return (double)d[2] * d[2] + (double)d[0] * d[0] + (double)d[1] * d[1];
}
/* F_i += k * [ d ; d . d(site_i)/d(angle) ]   — generalized force of U(|d|) with k = U'(dr)/dr
 * (the gradx..gradtheta blocks, compute_cuda.cu:98-138, :217-270, :352-464) */
static void add_bond_force(float *F /* AoS7 */, float k, const float d[3], const frame *fi, const float pi[3])
{
    float t[3];
    F[0] += k * d[0];
    F[1] += k * d[1];
    F[2] += k * d[2];
    apply(&fi->dfi, pi, t);
    F[3] += k * (d[0] * t[0] + d[1] * t[1] + d[2] * t[2]);
    apply(&fi->dtheta, pi, t);
    F[4] += k * (d[0] * t[0] + d[1] * t[1] + d[2] * t[2]);
    apply(&fi->dpsi, pi, t);
    F[5] += k * (d[0] * t[0] + d[1] * t[1] + d[2] * t[2]);
}
/* compute_cuda.cu:16-30 */
static float dmorse(float D, float a, float x) { return 2 * a * D * (1 - expf(-a * x)) * expf(-a * x); }
static float morse_en(float D, float a, float x) { return D * (1 - expf(-a * x)) * (1 - expf(-a * x)) - D; }
static float dbarr(float a, float r, float w, float x) { return -a * expf(-(x - r) * (x - r) / (2 * w * w)) * (x - r) / (w * w); }
static float barr(float a, float r, float w, float x) { return a * expf(-(x - r) * (x - r) / (2 * w * w)); }

static frame *all_frames(const float *coords, int N)
{
    frame *f = (frame *)malloc(sizeof(frame) * (size_t)N);
    for (int i = 0; i < N; i++) f[i] = make_frame(coords + (size_t)i * 7);
    return f;
}

/* decode a lateral list entry (compute_cuda.cu:307-324): returns j, *swapped = 1 for `j <= 0` */
static int lateral_decode(int v, int *swapped)
{
    *swapped = v <= 0;
    int j = abs(v);
    return j == ZERO ? 0 : j;
}

/* -------------------------------------------------------------------------------------------- LJ list */
void oracle_lj_lists(const maddy_params *par, const maddy_topology *top, const float *coords, oracle_lists *l)
{
    const int N = par->n_tot;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < par->n_tr_local; t++) {
        const float *c = coords + (size_t)t * N * 7;
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            l->lj_count[q] = 0;
            if (top->extra[q]) continue; /* built only for non-extra i, over ALL j (:925-936) */
            for (int j = 0; j < N; j++) {
                const float dx = c[i * 7] - c[j * 7], dy = c[i * 7 + 1] - c[j * 7 + 1], dz = c[i * 7 + 2] - c[j * 7 + 2];
                const float dr = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
                if (dr < par->ljpairscutoff && i != j) {
                    if (l->lj_count[q] < MADDY_LJ_CAPACITY) l->lj[q * MADDY_LJ_CAPACITY + l->lj_count[q]] = j;
                    l->lj_count[q]++;
                }
            }
            if (l->lj_count[q] > MADDY_LJ_CAPACITY) l->lj_count[q] = MADDY_LJ_CAPACITY;
        }
    }
}

/* -------------------------------------------------------------------------------------------- bond lists */
void oracle_pair_lists(const maddy_params *par, const maddy_topology *top, const float *coords, oracle_lists *l)
{
    const int N = par->n_tot, capL = par->max_longitudinal, capT = par->max_lateral, maxH = par->max_harmonic;
    float p1[3], p2[3];
    lateral_points(p1, p2);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < par->n_tr_local; t++) {
        const float *c = coords + (size_t)t * N * 7;
        frame *fr = all_frames(c, N);
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            int nl = 0, nt = 0;
            l->longitudinal_count[q] = l->lateral_count[q] = 0;
            if (top->extra[q]) continue;
            const int h = top->harmonic[maxH * i];
            const float rm = h < 0 ? -R_MON : R_MON; /* :548-551 */
            const float pe_i[3] = {0, 0, rm}, pe_j[3] = {0, 0, -rm};
            float d[3];
            for (int j = 0; j < N; j++) { /* longitudinal candidates :564-596 */
                if (top->mon_type[i] != top->mon_type[j] && abs(h) != j) {
                    const float dr2 = (float)site_delta(c + i * 7, &fr[i], pe_i, c + j * 7, &fr[j], pe_j, d);
                    if (sqrtf(dr2) < PAIR_CUTOFF) {
                        if (nl < capL) l->longitudinal[q * capL + nl] = h < 0 ? j : -j;
                        nl++;
                    }
                }
            }
            for (int j = 0; j < N; j++) { /* lateral candidates :598-669 */
                if (i != j && abs(h) != j) {
                    if (sqrtf((float)site_delta(c + i * 7, &fr[i], p1, c + j * 7, &fr[j], p2, d)) < PAIR_CUTOFF) {
                        if (nt < capT) l->lateral[q * capT + nt] = j != 0 ? -j : -ZERO;
                        nt++;
                    }
                    if (sqrtf((float)site_delta(c + i * 7, &fr[i], p2, c + j * 7, &fr[j], p1, d)) < PAIR_CUTOFF) {
                        if (nt < capT) l->lateral[q * capT + nt] = j != 0 ? j : ZERO;
                        nt++;
                    }
                }
            }
            l->longitudinal_count[q] = nl < capL ? nl : capL;
            l->lateral_count[q] = nt < capT ? nt : capT;
        }
        free(fr);
    }
}

/* -------------------------------------------------------------------------------------------- forces */
static void bending(const maddy_params *p, float *F, const float *ci, const float *cj, float rm, float theta0)
{
    /* compute_cuda.cu:141-180, :272-297: +B sin(q_j - q_i - q0) on the R_MON > 0 side, -B sin(q_i - q_j - q0) on the other */
    const float psiji = cj[5] - ci[5], thetaji = cj[4] - ci[4], fiji = cj[3] - ci[3];
    if (rm > 0) {
        F[5] += p->B_psi * sinf(psiji - p->psi_0);
        F[3] += p->B_fi * sinf(fiji - p->fi_0);
        F[4] += p->B_theta * sinf(thetaji - theta0);
    } else {
        F[5] -= p->B_psi * sinf(-psiji - p->psi_0);
        F[3] -= p->B_fi * sinf(-fiji - p->fi_0);
        F[4] -= p->B_theta * sinf(-thetaji - theta0);
    }
}

void oracle_forces(const maddy_params *par, const maddy_topology *top, const oracle_lists *l, const float *coords, float *forces)
{
    const int N = par->n_tot, capL = par->max_longitudinal, capT = par->max_lateral, maxH = par->max_harmonic;
    float p1[3], p2[3];
    lateral_points(p1, p2);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < par->n_tr_local; t++) {
        const float *c = coords + (size_t)t * N * 7;
        frame *fr = all_frames(c, N);
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            if (top->extra[q]) continue; /* :55 */
            float F[7] = {0, 0, 0, 0, 0, 0, 0};
            const float *ci = c + i * 7;
            float d[3];
            for (int k = 0; k < top->harmonic_count[i]; k++) { /* harmonic :70-182 */
                int j = top->harmonic[maxH * i + k];
                const float rm = j < 0 ? R_MON : -R_MON;
                j = abs(j);
                const float pe_i[3] = {0, 0, rm}, pe_j[3] = {0, 0, -rm};
                const float dr = sqrtf((float)site_delta(ci, &fr[i], pe_i, c + j * 7, &fr[j], pe_j, d));
                add_bond_force(F, par->C, d, &fr[i], pe_i);
                if (dr < ANGLE_CUTOFF) bending(par, F, ci, c + j * 7, rm, top->gtp[q] == 1 ? par->theta0_gtp : par->theta0_gdp);
            }
            for (int k = 0; k < l->longitudinal_count[q]; k++) { /* longitudinal :189-299 */
                int j = l->longitudinal[q * capL + k];
                const float rm = j < 0 ? R_MON : -R_MON;
                j = abs(j);
                const float *cj = c + j * 7;
                const float pe_i[3] = {0, 0, rm}, pe_j[3] = {0, 0, -rm};
                const float dr = sqrtf((float)site_delta(ci, &fr[i], pe_i, cj, &fr[j], pe_j, d));
                float dUdr = dr == 0 ? 0.0f : dmorse(par->D_long, par->A_long, dr) / dr;
                if (par->barrier && top->on_tubule_cur[q] == 0 && top->on_tubule_cur[(size_t)t * N + j] == 0 && dr != 0.0f)
                    dUdr += dbarr(par->a_barr_long, par->r_barr_long, par->w_barr_long, dr) / dr;
                add_bond_force(F, dUdr, d, &fr[i], pe_i);
                if (dr < ANGLE_CUTOFF) {
                    const size_t last = ci[2] > cj[2] ? q : (size_t)t * N + j; /* :282-283 */
                    bending(par, F, ci, cj, rm, top->gtp[last] == 1 ? par->theta0_gtp : par->theta0_gdp);
                }
            }
            for (int k = 0; k < l->lateral_count[q]; k++) { /* lateral :304-466 */
                int sw;
                const int j = lateral_decode(l->lateral[q * capT + k], &sw);
                const float *pi_ = sw ? p1 : p2, *pj_ = sw ? p2 : p1;
                const float dr = sqrtf((float)site_delta(ci, &fr[i], pi_, c + j * 7, &fr[j], pj_, d));
                float dUdr;
                if (dr == 0) dUdr = 0.0f;
                else if (top->mon_type[i] != top->mon_type[j]) dUdr = dmorse(par->D_lat / par->seam_coeff, par->A_lat, dr) / dr;
                else dUdr = dmorse(par->D_lat, par->A_lat, dr) / dr;
                if (par->barrier && top->on_tubule_cur[q] == 0 && top->on_tubule_cur[(size_t)t * N + j] == 0 && dr != 0.0f)
                    dUdr += dbarr(par->a_barr_lat, par->r_barr_lat, par->w_barr_lat, dr) / dr;
                add_bond_force(F, dUdr, d, &fr[i], pi_);
            }
            if (par->lj_on) { /* :470-495 */
                for (int k = 0; k < l->lj_count[q]; k++) {
                    const float *cj = c + l->lj[q * MADDY_LJ_CAPACITY + k] * 7;
                    const float dx = ci[0] - cj[0], dy = ci[1] - cj[1], dz = ci[2] - cj[2];
                    const float dr = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
                    if (dr < 6.0) {
                        const float df = (float)(6 / pow((double)dr, 8));
                        F[0] += par->ljscale * par->ljsigma6 * df * dx;
                        F[1] += par->ljscale * par->ljsigma6 * df * dy;
                        F[2] += par->ljscale * par->ljsigma6 * df * dz;
                    }
                }
            }
            if (par->is_wall) { /* :497-517, zs[traj] == rep_h */
                if (ci[2] < par->rep_leftborder) F[2] += par->rep_eps * fabsf(ci[2] - par->rep_leftborder);
                else if (ci[2] > par->rep_h + par->rep_leftborder) F[2] += -par->rep_eps * fabsf(ci[2] - (par->rep_h + par->rep_leftborder));
                const float rad2 = ci[0] * ci[0] + ci[1] * ci[1];
                if (rad2 > par->rep_r * par->rep_r) {
                    const float coeff = -par->rep_eps * (sqrtf(rad2) - par->rep_r);
                    F[0] += ci[0] / sqrtf(rad2) * coeff;
                    F[1] += ci[1] / sqrtf(rad2) * coeff;
                }
            }
            memcpy(forces + q * 7, F, sizeof F);
        }
        free(fr);
    }
}

/* -------------------------------------------------------------------------------------------- energies */
void oracle_energies(const maddy_params *par, const maddy_topology *top, const oracle_lists *l, const float *coords, double *energies)
{
    const int N = par->n_tot, capL = par->max_longitudinal, capT = par->max_lateral, maxH = par->max_harmonic;
    float p1[3], p2[3];
    lateral_points(p1, p2);
#pragma omp parallel for schedule(static)
    for (int t = 0; t < par->n_tr_local; t++) {
        const float *c = coords + (size_t)t * N * 7;
        frame *fr = all_frames(c, N);
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            float U_lat = 0, U_long = 0, U_harm = 0, U_fi = 0, U_psi = 0, U_teta = 0, U_lj = 0;
            const float *ci = c + i * 7;
            float d[3];
            if (!top->extra[q]) {
                for (int k = 0; k < top->harmonic_count[i]; k++) { /* :713-761 */
                    int j = top->harmonic[maxH * i + k];
                    const float rm = j < 0 ? R_MON : -R_MON;
                    j = abs(j);
                    const float *cj = c + j * 7;
                    const float pe_i[3] = {0, 0, rm}, pe_j[3] = {0, 0, -rm};
                    const float dr = (float)sqrt(site_delta(ci, &fr[i], pe_i, cj, &fr[j], pe_j, d));
                    U_harm = (float)(U_harm + (double)(par->C / 2) * pow((double)dr, 2));
                    if (dr < ANGLE_CUTOFF) { /* psi, fi use (i - j); theta uses (j - i) — :751-757 */
                        U_psi += par->B_psi * (1 - cosf((ci[5] - cj[5]) - par->psi_0));
                        U_fi += par->B_fi * (1 - cosf((ci[3] - cj[3]) - par->fi_0));
                        U_teta += par->B_theta * (1 - cosf((cj[4] - ci[4]) - (top->gtp[q] == 1 ? par->theta0_gtp : par->theta0_gdp)));
                    }
                }
                for (int k = 0; k < l->longitudinal_count[q]; k++) { /* :764-819 */
                    int j = l->longitudinal[q * capL + k];
                    const float rm = j < 0 ? R_MON : -R_MON;
                    j = abs(j);
                    const float *cj = c + j * 7;
                    const float pe_i[3] = {0, 0, rm}, pe_j[3] = {0, 0, -rm};
                    const float dr = sqrtf((float)site_delta(ci, &fr[i], pe_i, cj, &fr[j], pe_j, d));
                    U_long += morse_en(par->D_long, par->A_long, dr);
                    if (par->barrier && top->on_tubule_cur[q] == 0 && top->on_tubule_cur[(size_t)t * N + j] == 0)
                        U_long += barr(par->a_barr_long, par->r_barr_long, par->w_barr_long, dr);
                    if (dr < ANGLE_CUTOFF) {
                        const size_t last = ci[2] > cj[2] ? q : (size_t)t * N + j;
                        const float theta0 = top->gtp[last] == 1 ? par->theta0_gtp : par->theta0_gdp;
                        U_psi += par->B_psi * (1 - cosf((ci[5] - cj[5]) - par->psi_0));
                        U_fi += par->B_fi * (1 - cosf((ci[3] - cj[3]) - par->fi_0));
                        U_teta += par->B_theta * (1 - cosf((ci[4] - cj[4]) - theta0));
                    }
                }
                for (int k = 0; k < l->lateral_count[q]; k++) { /* :823-886 */
                    int sw;
                    const int j = lateral_decode(l->lateral[q * capT + k], &sw);
                    const float dr = (float)sqrt(site_delta(ci, &fr[i], sw ? p1 : p2, c + j * 7, &fr[j], sw ? p2 : p1, d));
                    if (top->mon_type[i] != top->mon_type[j]) U_lat += morse_en(par->D_lat / par->seam_coeff, par->A_lat, dr);
                    else U_lat += morse_en(par->D_lat, par->A_lat, dr);
                    if (par->barrier && top->on_tubule_cur[q] == 0 && top->on_tubule_cur[(size_t)t * N + j] == 0)
                        U_lat += barr(par->a_barr_lat, par->r_barr_lat, par->w_barr_lat, dr);
                }
                if (par->lj_on) { /* :889-898 */
                    for (int k = 0; k < l->lj_count[q]; k++) {
                        const float *cj = c + l->lj[q * MADDY_LJ_CAPACITY + k] * 7;
                        const float dx = ci[0] - cj[0], dy = ci[1] - cj[1], dz = ci[2] - cj[2];
                        const float dr = (float)sqrt((double)dx * dx + (double)dy * dy + (double)dz * dz);
                        if (dr < 6.0) U_lj = (float)(U_lj + (double)(par->ljscale * par->ljsigma6) / pow((double)dr, 6));
                    }
                }
            }
            double *e = energies + q * 7; /* Energies field order: harm,long,lat,psi,fi,teta,lj (mt.h:94-102) */
            e[0] = U_harm / 2;
            e[1] = U_long / 2;
            e[2] = U_lat / 2;
            e[3] = U_psi / 2;
            e[4] = U_fi / 2;
            e[5] = U_teta / 2;
            e[6] = U_lj / 2;
        }
        free(fr);
    }
}

/* -------------------------------------------------------------------------------------------- integrator */
void oracle_integrate(const maddy_params *par, const maddy_topology *top, float *coords, float *forces, unsigned *rng)
{
    const int N = par->n_tot;
    const size_t n = (size_t)par->n_tr_local * N;
#pragma omp parallel for schedule(static)
    for (long long qq = 0; qq < (long long)n; qq++) {
        const size_t q = (size_t)qq;
        float *r = coords + q * 7, *f = forces + q * 7;
        if (!top->fixed[q % N] && !top->extra[q]) { /* :949 */
            float nx[4], na[4];
            oracle_rforce(rng + q * 4, nx);
            oracle_rforce(rng + (n + q) * 4, na);
            r[0] += (par->dt / par->gammaR) * f[0] + par->varR * nx[0];
            r[1] += (par->dt / par->gammaR) * f[1] + par->varR * nx[1];
            r[2] += (par->dt / par->gammaR) * f[2] + par->varR * nx[2];
            r[3] += (par->dt / (par->gammaTheta * par->alpha)) * f[3] + (par->varTheta * sqrtf(par->freeze_temp / par->alpha)) * na[0];
            r[5] += (par->dt / (par->gammaTheta * par->alpha)) * f[5] + (par->varTheta * sqrtf(par->freeze_temp / par->alpha)) * na[1];
            r[4] += (par->dt / par->gammaTheta) * f[4] + par->varTheta * na[2];
        }
        for (int k = 0; k < 6; k++) f[k] = 0.0f; /* :966-972 */
    }
}

void oracle_run(const maddy_params *par, const maddy_topology *top, oracle_lists *l, float *coords, float *forces, unsigned *rng,
                long long first_step, long long n_steps, int skip_first_rebuild)
{
    for (long long step = first_step; step < first_step + n_steps; step++) { /* compute_cuda.cu:1137-1238 */
        if (step % par->ljpairsupdatefreq == 0 && !(step == first_step && skip_first_rebuild)) {
            if (par->lj_on) oracle_lj_lists(par, top, coords, l);
            if (par->is_assembly) oracle_pair_lists(par, top, coords, l);
        }
        oracle_forces(par, top, l, coords, forces);
        oracle_integrate(par, top, coords, forces, rng);
    }
}

/* -------------------------------------------------------------------------------------------- TEA */
typedef struct { float xx, xy, xz, yy, yz, zz; } sym6;
/* bdhitea_kernel.cu:38-58 */
static sym6 rpy(float x, float y, float z, float w, float a)
{
    const float ra = w / a;
    float crr, cii;
    if (ra > 2.f) {
        crr = 0.75f / ra * (1.f - 2.f / ra / ra);
        cii = 0.75f / ra * (1.f + 2.f / 3.f / ra / ra);
    } else {
        crr = 3.f * ra / 32.f;
        cii = 1.f - 9.f * ra / 32.f;
    }
    sym6 d = {x * x * crr + cii, x * y * crr, x * z * crr, y * y * crr + cii, y * z * crr, z * z * crr + cii};
    return d;
}

/* integrateTea_epsilon_unlisted (:84-101) + host part of updateTea (bdhitea.cu:57-118) */
int oracle_tea_update(const maddy_params *par, const maddy_topology *top, const float *coords, float *ci, float *eps, float *beta)
{
    const int N = par->n_tot;
    int rc = MADDY_OK;
    for (int t = 0; t < par->n_tr_local; t++) {
        const float *c = coords + (size_t)t * N * 7;
        int nnoextra = 0;
        for (int i = 0; i < N; i++)
            if (!top->extra[(size_t)t * N + i]) nnoextra++;
        double epsilon = 0.0;
        for (int i = 0; i < N; i++) {
            const size_t q = (size_t)t * N + i;
            float sx = 0, sy = 0, sz = 0, sw = 0;
            for (int j = 0; j < N; j++) {
                if (j == i || top->extra[(size_t)t * N + j] || top->extra[q]) continue;
                float dx = c[j * 7] - c[i * 7], dy = c[j * 7 + 1] - c[i * 7 + 1], dz = c[j * 7 + 2] - c[i * 7 + 2];
                const float w = sqrtf(dx * dx + dy * dy + dz * dz);
                dx /= w;
                dy /= w;
                dz /= w;
                const sym6 d = rpy(dx, dy, dz, w, par->tea_a);
                sw += d.xx + 2 * d.xy + 2 * d.xz + d.yy + 2 * d.yz + d.zz;
                sx += d.xx * d.xx + d.xy * d.xy + d.xz * d.xz;
                sy += d.xy * d.xy + d.yy * d.yy + d.yz * d.yz;
                sz += d.xz * d.xz + d.yz * d.yz + d.zz * d.zz;
            }
            ci[q * 4] = sx;
            ci[q * 4 + 1] = sy;
            ci[q * 4 + 2] = sz;
            ci[q * 4 + 3] = 0.f;
            eps[q] = sw;
            epsilon += sw;
        }
        double e;
        int r = oracle_tea_beta(epsilon, nnoextra, par->tea_capricious, par->tea_a, par->tea_epsmax, &beta[t], &e);
        if (r) rc = r;
    }
    return rc;
}

/* bdhitea.cu:79-113 */
int oracle_tea_beta(double epsilon_sum, int n, int capricious, float tea_a, float epsmax, float *beta, double *eps_out)
{
    int rc = MADDY_OK;
    double epsilon = epsilon_sum / (3. * n * (3. * n - 3.));
    if (epsilon > 1.0) {
        if (capricious) rc = MADDY_ETEA;
        epsilon = 1.0;
    }
    if (epsilon > epsmax) rc = MADDY_ETEA;
    double a = (3. * n - 1.) * epsilon * epsilon - (3. * n - 2.) * epsilon;
    if (fabs(a) < 1e-7) {
        *beta = .5f;
        if (capricious && tea_a > 0.0f) rc = MADDY_ETEA;
    } else {
        *beta = (float)((1. - sqrt(1. - a)) / a);
    }
    if (eps_out) *eps_out = epsilon;
    return rc;
}

/* integrateTea_prepare + integrateTea_kernel_unlisted (bdhitea_kernel.cu:16-36, :148-213) */
void oracle_tea_integrate(const maddy_params *par, const maddy_topology *top, float *coords, float *forces, unsigned *rng,
                          const float *ci_raw, const float *beta)
{
    const int N = par->n_tot;
    const size_t n = (size_t)par->n_tr_local * N;
    float *rf = (float *)malloc(n * 3 * sizeof(float)), *mf = (float *)malloc(n * 3 * sizeof(float)),
          *co = (float *)malloc(n * 3 * sizeof(float));
    const float var = sqrtf(2.0f * KB * par->Temp * par->gammaR / par->dt);
    for (size_t q = 0; q < n; q++) { /* prepare: every bead draws */
        float nx[4];
        oracle_rforce(rng + q * 4, nx);
        for (int k = 0; k < 3; k++) {
            rf[q * 3 + k] = nx[k] * var;
            mf[q * 3 + k] = forces[q * 7 + k];
            forces[q * 7 + k] = 0.f;
            co[q * 3 + k] = coords[q * 7 + k];
        }
    }
    const float mult = par->dt / par->gammaR;
    for (size_t q = 0; q < n; q++) {
        const int t = (int)(q / N);
        const float b = beta[t], b2 = b * b;
        float cx = 1.f / sqrtf(1.f + b2 * ci_raw[q * 4]), cy = 1.f / sqrtf(1.f + b2 * ci_raw[q * 4 + 1]),
              cz = 1.f / sqrtf(1.f + b2 * ci_raw[q * 4 + 2]);
        float fx = mf[q * 3] + rf[q * 3] * cx, fy = mf[q * 3 + 1] + rf[q * 3 + 1] * cy, fz = mf[q * 3 + 2] + rf[q * 3 + 2] * cz;
        cx *= b;
        cy *= b;
        cz *= b;
        for (size_t j = (size_t)t * N; j < (size_t)(t + 1) * N; j++) {
            if (j == q || top->extra[q] || top->extra[j]) continue;
            float dx = co[j * 3] - co[q * 3], dy = co[j * 3 + 1] - co[q * 3 + 1], dz = co[j * 3 + 2] - co[q * 3 + 2];
            const float w = sqrtf(dx * dx + dy * dy + dz * dz);
            dx /= w;
            dy /= w;
            dz /= w;
            const float gx = mf[j * 3] + rf[j * 3] * cx, gy = mf[j * 3 + 1] + rf[j * 3 + 1] * cy, gz = mf[j * 3 + 2] + rf[j * 3 + 2] * cz;
            const sym6 d = rpy(dx, dy, dz, w, par->tea_a);
            fx += d.xx * gx + d.xy * gy + d.xz * gz;
            fy += d.xy * gx + d.yy * gy + d.yz * gz;
            fz += d.xz * gx + d.yz * gy + d.zz * gz;
        }
        float na[4];
        oracle_rforce(rng + (n + q) * 4, na); /* angular stream advances for every bead (:194) */
        if (!top->fixed[q % N] && !top->extra[q]) {
            float *r = coords + q * 7, *f = forces + q * 7;
            r[0] = co[q * 3] + mult * fx;
            r[1] = co[q * 3 + 1] + mult * fy;
            r[2] = co[q * 3 + 2] + mult * fz;
            r[3] += (par->dt / (par->gammaTheta * par->alpha)) * f[3] + (par->varTheta * sqrtf(par->freeze_temp / par->alpha)) * na[0];
            r[5] += (par->dt / (par->gammaTheta * par->alpha)) * f[5] + (par->varTheta * sqrtf(par->freeze_temp / par->alpha)) * na[1];
            r[4] += (par->dt / par->gammaTheta) * f[4] + par->varTheta * na[2];
        }
    }
    free(rf);
    free(mf);
    free(co);
}
