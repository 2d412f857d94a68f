/*
 * ref_events_probe.cu — TEST INFRASTRUCTURE.  Drives the reference's OWN host callbacks of the stride block,
 *   mt_length() (updater.cpp:154-227), hydrolyse() (:229-257), change_conc() (:97-152),
 * on inputs prepared by tests/golden/make_events_golden.py and dumps what they did.  The reference's updater.cpp and
 * globals.cpp are compiled from where they lie (mt_b200/build.py::build_reference); nothing of them is copied.  Host
 * code only: runs without a GPU, so the golden fixtures (tests/golden/ref_events.npz) are made in the build container.
 *
 * case file (little endian): int32 Ntot, Ntr, seed, n_hydrolyse; float rep_r, rep_h, rep_leftborder, conc;
 *   float r[Ntr*Ntot][7]; int32 gtp[], on_cur[], on_prev[], extra[] (each Ntr*Ntot); int32 mon_type[Ntot]; int32 mt_len_prev[Ntr]
 * output: after srand(seed): mt_length(1000) -> on_cur, mt_len; n_hydrolyse x hydrolyse() -> gtp after each;
 *   change_conc(mt_len - mt_len_prev, mt_len) -> flag, extra, r; then the next four rand() values.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "updater.h"

// what update() would call (not exercised here)
void saveCoordDCD() {}
void appendCoordPDB() {}
void printTime(long long int) {}
void printEstimatedTimeleft(float) {}

template <class T> static void rd(FILE *f, T *p, size_t n)
{
    if (fread(p, sizeof(T), n, f) != n) {
        fprintf(stderr, "ref_events_probe: short read\n");
        exit(2);
    }
}
template <class T> static void wr(FILE *f, const T *p, size_t n) { fwrite(p, sizeof(T), n, f); }

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *in = fopen(argv[1], "rb");
    if (!in) return 1;
    int hdr[4];
    float fl[4];
    rd(in, hdr, 4);
    rd(in, fl, 4);
    const int N = hdr[0], Ntr = hdr[1], seed = hdr[2], nhyd = hdr[3];
    const size_t n = (size_t)N * Ntr;
    par.Ntot = N;
    par.Ntr = Ntr;
    par.rep_r = fl[0];
    for (int t = 0; t < Ntr && t < 100; t++) par.zs[t] = fl[1];
    par.rep_leftborder = fl[2];
    par.conc = fl[3];
    r = (Coord *)calloc(n, sizeof(Coord));
    rd(in, (float *)r, n * 7);
    std::vector<int> ex(n), len_prev(Ntr);
    top.gtp = (int *)calloc(n, sizeof(int));
    top.on_tubule_cur = (int *)calloc(n, sizeof(int));
    top.on_tubule_prev = (int *)calloc(n, sizeof(int));
    top.extra = (bool *)calloc(n, sizeof(bool));
    top.mon_type = (int *)calloc(N, sizeof(int));
    rd(in, top.gtp, n);
    rd(in, top.on_tubule_cur, n);
    rd(in, top.on_tubule_prev, n);
    rd(in, ex.data(), n);
    for (size_t q = 0; q < n; q++) top.extra[q] = ex[q] != 0;
    rd(in, top.mon_type, N);
    rd(in, len_prev.data(), Ntr);
    fclose(in);

    FILE *out = fopen(argv[2], "wb");
    std::vector<int> mt_len(Ntr), delta(Ntr);
    srand(seed);
    freopen("/dev/null", "w", stdout); // the callbacks print; the dump is the record
    mt_length(1000, mt_len.data());   // (appends a line to ./mt_len.dat, as the reference does)
    wr(out, top.on_tubule_cur, n);
    wr(out, mt_len.data(), Ntr);
    for (int k = 0; k < nhyd; k++) {
        hydrolyse();
        wr(out, top.gtp, n);
    }
    for (int t = 0; t < Ntr; t++) delta[t] = mt_len[t] - len_prev[t];
    const int flag = change_conc(delta.data(), mt_len.data());
    wr(out, &flag, 1);
    for (size_t q = 0; q < n; q++) ex[q] = top.extra[q] ? 1 : 0;
    wr(out, ex.data(), n);
    wr(out, (float *)r, n * 7);
    int nxt[4];
    for (int k = 0; k < 4; k++) nxt[k] = rand();
    wr(out, nxt, 4);
    fclose(out);
    return 0;
}
