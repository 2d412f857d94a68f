/*
 * checkpoint.cpp — exact checkpoint / restart of a run.
 *
 * The reference's restart path reads coordinates from XYZ text files (preparator.cpp:664-684) and its writer
 * (writeRestart, :686-714) is never called; a run resumed that way starts a different random trajectory.  A
 * checkpoint here holds everything the step loop's future depends on, so that
 *     run(0 .. T)   ==   run(0 .. S), checkpoint, new process, resume(S .. T)        bit for bit
 * in the state, the DCD trajectories and mt_len.dat (both trimmed back to the checkpoint step on resume).  NOT covered:
 * dcd/hydrolysis.pdb and the in-situ analysis files are appended as they stand (frames a crashed run wrote after its last
 * checkpoint stay, in-situ frame numbers restart at 1 and the first resumed frame has no temperature line).  Held:
 * coordinates as raw float bits (no text round trip, no angle re-wrapping), both HybridTaus streams of every
 * monomer, GTP / reserve / on-tubule flags (current and previous stride), the tubule lengths, the state of the host
 * rand() generator that drives hydrolysis and insertion, the step, and whether the hydrolysis event of that step has
 * already been drawn.  Checkpoints are taken at list-update steps only, so no neighbour list has to be stored: the
 * first resumed step rebuilds them from the same positions the uninterrupted run would have used.
 *
 * File: little-endian, header + arrays + FNV-1a checksum, written to <name>.tmp and renamed.
 */
#include <cstdio>
#include <unistd.h>
#include <cstring>
#include <string>
#include <cstdlib>
#include "mt_host.hpp"

namespace mt {

namespace {
struct CkHeader {
    char magic[8];
    int32_t version, n_tot, n_tr, rseed;
    int64_t step, hydrolysed_for;
    float dt;
    int32_t rand_state[33];
};
const char kMagic[8] = {'M', 'A', 'D', 'D', 'Y', 'C', 'K', '1'};

struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void add(const void *p, size_t n)
    {
        const unsigned char *c = (const unsigned char *)p;
        for (size_t i = 0; i < n; i++) h = (h ^ c[i]) * 1099511628211ull;
    }
};
struct Out {
    FILE *f;
    Fnv sum;
    const std::string &name;
    void put(const void *p, size_t n)
    {
        if (n && fwrite(p, 1, n, f) != n) die("Writing checkpoint '%s'", name.c_str());
        sum.add(p, n);
    }
};
struct In {
    FILE *f;
    Fnv sum;
    const std::string &name;
    void get(void *p, size_t n)
    {
        if (n && fread(p, 1, n, f) != n) die("Checkpoint '%s' is truncated", name.c_str());
        sum.add(p, n);
    }
};
} // namespace

bool checkpoint_peek(const std::string &name, long long *step)
{
    FILE *f = fopen(name.c_str(), "rb");
    if (!f) return false;
    CkHeader h;
    const bool ok = fread(&h, sizeof h, 1, f) == 1 && memcmp(h.magic, kMagic, 8) == 0 && h.version == 1;
    fclose(f);
    if (ok && step) *step = h.step;
    return ok;
}

void checkpoint_save(System &s, const std::string &name, const CheckpointState &st)
{
    const size_t n = (size_t)s.par.n_tot * s.par.n_tr;
    if (st.coords.size() != n * 7 || st.rng.size() != n * 8) die("checkpoint_save: state arrays have the wrong size");
    const std::string tmp = name + ".tmp";
    FILE *f = fopen(tmp.c_str(), "wb");
    if (!f) die("Opening file '%s'", tmp.c_str());
    Out o{f, Fnv(), tmp};
    CkHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, kMagic, 8);
    h.version = 1;
    h.n_tot = s.par.n_tot;
    h.n_tr = s.par.n_tr;
    h.rseed = s.par.rseed;
    h.step = st.step;
    h.hydrolysed_for = st.hydrolysed_for;
    h.dt = s.par.dt;
    s.rng.get_state(h.rand_state);
    o.put(&h, sizeof h);
    o.put(st.coords.data(), n * 7 * sizeof(float));
    o.put(st.rng.data(), n * 8 * sizeof(uint32_t));
    o.put(s.gtp.data(), n * sizeof(int));
    o.put(s.extra.data(), n);
    o.put(s.on_tubule_cur.data(), n * sizeof(int));
    o.put(s.on_tubule_prev.data(), n * sizeof(int));
    o.put(st.mt_len.data(), (size_t)s.par.n_tr * sizeof(int));
    o.put(st.mt_len_prev.data(), (size_t)s.par.n_tr * sizeof(int));
    const uint64_t sum = o.sum.h;
    if (fwrite(&sum, sizeof sum, 1, f) != 1 || fclose(f) != 0) die("Writing checkpoint '%s'", tmp.c_str());
    if (rename(tmp.c_str(), name.c_str()) != 0) die("Renaming '%s' to '%s'", tmp.c_str(), name.c_str());
    if (!s.quiet) printf("Checkpoint written at step %lld: %s\n", (long long)st.step, name.c_str());
}

void checkpoint_load(System &s, const std::string &name, CheckpointState &st)
{
    const size_t n = (size_t)s.par.n_tot * s.par.n_tr;
    FILE *f = fopen(name.c_str(), "rb");
    if (!f) die("Opening file '%s'", name.c_str());
    In in{f, Fnv(), name};
    CkHeader h;
    in.get(&h, sizeof h);
    if (memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) die("'%s' is not a checkpoint of this program", name.c_str());
    if (h.n_tot != s.par.n_tot || h.n_tr != s.par.n_tr)
        die("Checkpoint '%s' holds %d x %d monomers, the configuration asks for %d x %d", name.c_str(), h.n_tr, h.n_tot, s.par.n_tr, s.par.n_tot);
    if (h.rseed != s.par.rseed || h.dt != s.par.dt) die("Checkpoint '%s' was written with another rseed / dt", name.c_str());
    st.step = h.step;
    st.hydrolysed_for = h.hydrolysed_for;
    st.coords.resize(n * 7);
    st.rng.resize(n * 8);
    st.mt_len.resize(s.par.n_tr);
    st.mt_len_prev.resize(s.par.n_tr);
    in.get(st.coords.data(), n * 7 * sizeof(float));
    in.get(st.rng.data(), n * 8 * sizeof(uint32_t));
    in.get(s.gtp.data(), n * sizeof(int));
    in.get(s.extra.data(), n);
    in.get(s.on_tubule_cur.data(), n * sizeof(int));
    in.get(s.on_tubule_prev.data(), n * sizeof(int));
    in.get(st.mt_len.data(), (size_t)s.par.n_tr * sizeof(int));
    in.get(st.mt_len_prev.data(), (size_t)s.par.n_tr * sizeof(int));
    uint64_t sum = 0;
    const bool ok = fread(&sum, sizeof sum, 1, f) == 1 && sum == in.sum.h;
    fclose(f);
    if (!ok) die("Checkpoint '%s' is corrupt (checksum mismatch)", name.c_str());
    s.rng.set_state(h.rand_state);
    s.r = st.coords;
    if (!s.quiet) printf("Resuming from checkpoint %s at step %lld\n", name.c_str(), (long long)st.step);
}

// DCD files of a resumed run: keep the header and the frames written for strides before `step`, drop anything later
// (frames a crashed run wrote after its last checkpoint), so the resumed run appends where the checkpoint stands.
void checkpoint_trim_outputs(System &s, long long step)
{
    const long long frames = step <= 0 ? 0 : (step - 1) / s.hp.stride + 1;
    const long long header = 276, frame_bytes = 3 * ((long long)s.par.n_tot * 4 + 8);
    for (int t = 0; t < s.par.n_tr; t++)
        for (const std::string *name : {&s.hp.dcd_xyz[t], &s.hp.dcd_ang[t]}) {
            FILE *f = fopen(name->c_str(), "rb");
            if (!f) die("Resuming: trajectory file '%s' is missing", name->c_str());
            fseek(f, 0, SEEK_END);
            const long long size = ftell(f);
            fclose(f);
            const long long want = header + frames * frame_bytes;
            if (size < want) die("Resuming at step %lld: '%s' holds fewer than %lld frames", step, name->c_str(), frames);
            if (size > want && truncate(name->c_str(), want) != 0) die("Truncating '%s'", name->c_str());
        }
    // mt_len.dat (one line per stride step, first field = step; updater.cpp:213-219): lines of strides at or after `step` go
    {
        FILE *f = fopen("mt_len.dat", "rb");
        if (f) {
            std::string keep;
            char *line = nullptr;
            size_t cap = 0;
            ssize_t len;
            while ((len = getline(&line, &cap, f)) > 0) {
                char *end = nullptr;
                const long long at = strtoll(line, &end, 10);
                if (end == line || at >= step) break;
                keep.append(line, (size_t)len);
            }
            free(line);
            fclose(f);
            f = fopen("mt_len.dat", "wb");
            if (!f) die("Rewriting mt_len.dat");
            fwrite(keep.data(), 1, keep.size(), f);
            fclose(f);
        }
    }
}

} // namespace mt
