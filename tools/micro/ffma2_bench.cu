// Microbenchmark: FP32 throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ffma2_bench tools/micro/ffma2_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int PACKED>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float2 v[8];
    for (int q = 0; q < 8; q++) v[q] = make_float2(threadIdx.x * 1e-3f + q, blockIdx.x * 1e-4f - q);
    const float2 A = make_float2(a, a * 1.01f), B = make_float2(b, b * 0.99f);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) {
            if (PACKED) v[q] = __ffma2_rn(v[q], A, B);
            else {
                v[q].x = fmaf(v[q].x, A.x, B.x);
                v[q].y = fmaf(v[q].y, A.y, B.y);
            }
        }
    }
    float s = 0;
    for (int q = 0; q < 8; q++) s += v[q].x + v[q].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float *out;
    const int blocks = 148 * 8, iters = 20000;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int packed = 0; packed < 2; packed++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (packed) k<1><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
            else k<0><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            const double flop = 2.0 * 16 * iters * (double)blocks * 256;
            printf("%s: %.3f ms, %.1f TFLOP/s\n", packed ? "FFMA2 (f32x2)" : "FFMA scalar  ", ms, flop / ms / 1e9);
        }
    }
    return 0;
}
