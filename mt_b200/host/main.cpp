/*
 * main.cpp — `mt <config.conf> [name=value ...]`: drop-in replacement of the reference
 * executable (src/main.cpp:25-115).  Same call order: parse config, initParameters,
 * srand(rseed), timers, AssemblyInit, compute, result_{xyz,ang}.pdb.  MPI is gone: the
 * reference only used it to pick a device per rank (main.cpp:27-50); here one process drives
 * `n_gpus` devices and shards the trajectories itself.
 */
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#include "mt_host.hpp"

int main(int argc, char *argv[])
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s <config.conf> [name=value ...]\n", argv[0]);
        return -1;
    }
    try {
        mt::System s;
        std::vector<std::string> overrides;
        for (int i = 2; i < argc; i++) overrides.emplace_back(argv[i]);
        mt::init_parameters(s, argv[1], overrides);
        s.rng.seed(s.par.rseed); // srand(par.rseed) of the reference (main.cpp:67), private state
        mt::init_timer(s);
        if (s.par.is_assembly) mt::assembly_init(s);
        mt::compute(s, true, nullptr);
        mt::save_coord_pdb(s, "result_xyz.pdb", "result_ang.pdb");
    } catch (const mt::Fatal &e) {
        fflush(stdout);
        fprintf(stderr, "=====================================\nFatal error!\n%s\n", e.what());
        void *bt[30];
        int n = backtrace(bt, 30);
        fprintf(stderr, "Stack trace (%d frames):\n", n);
        backtrace_symbols_fd(bt, n, 2);
        exit(-1);
    }
    return 0;
}
