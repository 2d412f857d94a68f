/*
 * gen_golden_seeds.cu — known answers of the reference's own seed generator.  TEST INFRASTRUCTURE.
 *
 * Links the reference's src/HybridTaus.cu in place and calls its HOST function generateSeeds()
 * (HybridTaus.cu:32-48; no GPU needed).  ran2 keeps static state, so each case runs in a fresh
 * process:   gen_golden_seeds <rseed> <Np>   prints a JSON object with selected states.
 *
 *   nvcc -O2 -arch=sm_100 -I/root/reference/src -o oracle/_ref/gen_golden_seeds \
 *        oracle/gen_golden_seeds.cu /root/reference/src/HybridTaus.cu
 *   python tests/golden/make_seed_golden.py      (writes tests/golden/seeds.json)
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "ht.cuh"

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    const int seed = atoi(argv[1]);
    const int np = atoi(argv[2]);
    std::vector<uint4> s(np);
    generateSeeds(s.data(), seed, np);
    unsigned long long sum = 0, x = 0;
    for (int i = 0; i < np; i++) {
        sum += (unsigned long long)s[i].x + s[i].y + s[i].z + s[i].w;
        x ^= ((unsigned long long)s[i].x << 32 | s[i].y) ^ ((unsigned long long)s[i].z << 32 | s[i].w) * 0x9E3779B97F4A7C15ull;
        x = (x << 7) | (x >> 57);
    }
    printf("{\"rseed\": %d, \"np\": %d, \"sum\": %llu, \"mix\": %llu, \"states\": {", seed, np, sum, x);
    const int picks[] = {0, 1, 2, np / 2 - 1, np / 2, np - 1};
    for (int k = 0; k < 6; k++)
        printf("%s\"%d\": [%u, %u, %u, %u]", k ? ", " : "", picks[k], s[picks[k]].x, s[picks[k]].y, s[picks[k]].z, s[picks[k]].w);
    printf("}}\n");
    return 0;
}
