/*
 * maddy_lfib.h — jump-ahead for the generator behind libc rand() (glibc TYPE_3 additive feedback, the stream
 * hydrolyse() and change_conc() consume: updater.cpp:118-119, :235):
 *
 *     x[n] = x[n-31] + x[n-3]   (mod 2^32),      rand() = x[n] >> 1.
 *
 * The recurrence is linear over Z/2^32, so with p(z) = z^31 - z^28 - 1 and  z^m mod p = sum_k c_k z^k  one has
 * x[m] = sum_k c_k x[k] for ANY m: a block of the stream can be produced from the 31-word window at its start, and that
 * window from the base window with ~log2(m) polynomial products.  This is what lets the device draw the numbers of a
 * hydrolysis event itself, in the reference's order, from nothing but the host generator's 31 words.
 * Plain C++ usable from host and device code; test infrastructure checks it against libc rand() (tests/test_events.py).
 */
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define LFIB_HD __host__ __device__ inline
#else
#define LFIB_HD inline
#endif

#define LFIB_DEG 31
#define LFIB_POW2 48 /* z^(2^k) tabulated for k < 48: offsets below 2^48 draws */

struct LfibPoly {
    uint32_t c[LFIB_DEG];
};

// r = a * b mod p(z)   (z^d = z^(d-3) + z^(d-31) for d >= 31)
LFIB_HD void lfib_mul(LfibPoly &r, const LfibPoly &a, const LfibPoly &b)
{
    uint32_t t[2 * LFIB_DEG - 1];
    for (int i = 0; i < 2 * LFIB_DEG - 1; i++) t[i] = 0;
    for (int i = 0; i < LFIB_DEG; i++) {
        const uint32_t ai = a.c[i];
        if (ai == 0) continue;
        for (int j = 0; j < LFIB_DEG; j++) t[i + j] += ai * b.c[j];
    }
    for (int d = 2 * LFIB_DEG - 2; d >= LFIB_DEG; d--) {
        t[d - 3] += t[d];
        t[d - LFIB_DEG] += t[d];
    }
    for (int i = 0; i < LFIB_DEG; i++) r.c[i] = t[i];
}

// q = z * q mod p(z)
LFIB_HD void lfib_shift(LfibPoly &q)
{
    const uint32_t top = q.c[LFIB_DEG - 1];
    for (int i = LFIB_DEG - 1; i > 0; i--) q.c[i] = q.c[i - 1];
    q.c[0] = top;
    q.c[LFIB_DEG - 3] += top; // z^31 = z^28 + 1
}

// table[k] = z^(2^k) mod p, k < LFIB_POW2
LFIB_HD void lfib_table(LfibPoly *table)
{
    for (int i = 0; i < LFIB_DEG; i++) table[0].c[i] = 0;
    table[0].c[1] = 1;
    for (int k = 1; k < LFIB_POW2; k++) lfib_mul(table[k], table[k - 1], table[k - 1]);
}

// q = z^m mod p
LFIB_HD void lfib_power(LfibPoly &q, const LfibPoly *table, unsigned long long m)
{
    for (int i = 0; i < LFIB_DEG; i++) q.c[i] = 0;
    q.c[0] = 1;
    for (int k = 0; k < LFIB_POW2 && (m >> k) != 0; k++)
        if ((m >> k) & 1ull) {
            LfibPoly t;
            lfib_mul(t, q, table[k]);
            q = t;
        }
}

// out[i] = x[m + i], i < 31, from the base window base[k] = x[k]
LFIB_HD void lfib_window(uint32_t *out, const uint32_t *base, const LfibPoly *table, unsigned long long m)
{
    LfibPoly q;
    lfib_power(q, table, m);
    for (int i = 0; i < LFIB_DEG; i++) {
        uint32_t v = 0;
        for (int k = 0; k < LFIB_DEG; k++) v += q.c[k] * base[k];
        out[i] = v;
        lfib_shift(q);
    }
}
