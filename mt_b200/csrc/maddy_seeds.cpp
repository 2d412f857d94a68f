/*
 * maddy_seeds.cpp — host-side, GPU-free pieces of the path: the HybridTaus seed table and
 * the TEA beta formula.  Exact integer / IEEE-double work; must be bit-reproducible.
 *
 * Seed table (reference src/HybridTaus.cu:32-48 + src/ran2.h:18-56): the reference fills
 * uint4[Np] sequentially with (unsigned)(ran2(&seed) * UINT_MAX), re-drawing any component
 * below 128.  ran2 is the Numerical-Recipes L'Ecuyer combined generator with a Bays-Durham
 * shuffle; the reference calls it with a POSITIVE seed, so its initialisation branch never
 * runs and the generator starts from its static state: second stream 123456789, shuffle
 * table all zero, previous output 0.  That quirk defines the stream and is reproduced here.
 */
#include <limits.h>
#include <math.h>
#include <stdint.h>
#include "maddy_b200.h"

namespace {

struct LEcuyerShuffle {
    // moduli / multipliers / Schrage factors of the two L'Ecuyer streams
    static constexpr int32_t M1 = 2147483563, M2 = 2147483399;
    static constexpr int32_t A1 = 40014, A2 = 40692;
    static constexpr int32_t Q1 = 53668, Q2 = 52774;
    static constexpr int32_t R1 = 12211, R2 = 3791;
    static constexpr int TAB = 32;
    static constexpr int32_t DIV = 1 + (M1 - 1) / TAB;

    int32_t s1;             // first stream  (the caller's seed variable in the reference)
    int32_t s2 = 123456789; // second stream (static initial value)
    int32_t last = 0;       // previous shuffled output
    int32_t table[TAB] = {0};

    explicit LEcuyerShuffle(int32_t seed) : s1(seed) {}

    double next()
    {
        int32_t k = s1 / Q1;
        s1 = A1 * (s1 - k * Q1) - k * R1;
        if (s1 < 0) s1 += M1;
        k = s2 / Q2;
        s2 = A2 * (s2 - k * Q2) - k * R2;
        if (s2 < 0) s2 += M2;
        const int slot = last / DIV;
        last = table[slot] - s2;
        table[slot] = s1;
        if (last < 1) last += M1 - 1;
        const double v = (1.0 / M1) * last;
        const double vmax = 1.0 - 1.2e-7;
        return v > vmax ? vmax : v;
    }
};

} // namespace

extern "C" void maddy_generate_seeds(unsigned *seeds, int rseed, long long np)
{
    LEcuyerShuffle g(rseed);
    for (long long q = 0; q < 4 * np; q++) {
        unsigned v;
        do {
            v = (unsigned)(g.next() * UINT_MAX);
        } while (v < 128);
        seeds[q] = v;
    }
}

/*
 * TEA coupling coefficient (reference src/bdhitea.cu:79-113; Geyer & Winter 2009 eq. 22, 26):
 *   eps = sum / (3N (3N - 3)),  a = (3N-1) eps^2 - (3N-2) eps,  beta = (1 - sqrt(1-a)) / a
 * with N the number of non-extra beads.  Returns MADDY_ETEA where the reference exits.
 */
extern "C" int maddy_tea_beta(double epsilon_sum, int n_noextra, int capricious, float tea_a, float epsmax,
                              float *beta_out, double *epsilon_out)
{
    const double n3 = 3. * n_noextra;
    double eps = epsilon_sum / (n3 * (n3 - 3.));
    int rc = MADDY_OK;
    if (eps > 1.0) {
        if (capricious) rc = MADDY_ETEA;
        eps = 1.0;
    }
    if (eps > epsmax) rc = MADDY_ETEA;
    const double a = (n3 - 1.) * eps * eps - (n3 - 2.) * eps;
    float beta;
    if (fabs(a) < 1e-7) {
        beta = .5f;
        if (capricious && tea_a > 0.0f) rc = MADDY_ETEA;
    } else {
        beta = (float)((1. - sqrt(1. - a)) / a);
    }
    if (beta_out) *beta_out = beta;
    if (epsilon_out) *epsilon_out = eps;
    return rc;
}
