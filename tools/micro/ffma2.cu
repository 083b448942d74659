// microbenchmark: issue-slot cost of packed f32x2 math (FFMA2/FADD2/FMUL2) vs scalar FFMA on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float2* out, int iters, float2 seed) {
    float2 a[8], b = seed, c = make_float2(seed.y, seed.x);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
            else if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
            else if (MODE == 2) { a[i].x = a[i].x + b.x; a[i].y = a[i].y + b.y; }
            else a[i] = __fadd2_rn(a[i], b);
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int i = 0; i < 8; ++i) { s.x += a[i].x; s.y += a[i].y; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(float2* d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(d, 10, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(d, iters, make_float2(1.0001f, 0.9999f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    float2* d; cudaMalloc(&d, 148 * 4 * 256 * sizeof(float2));
    const int iters = 20000;
    const double flops = 148.0 * 4 * 256 * iters * 8 * 2;   // float ops (fma counted as 1 op x 2 lanes)
    float t0 = run<0>(d, iters), t1 = run<1>(d, iters), t2 = run<2>(d, iters), t3 = run<3>(d, iters);
    printf("scalar FFMA  : %.3f ms  %.2f Tfma/s\n", t0, flops / t0 / 1e9);
    printf("packed FFMA2 : %.3f ms  %.2f Tfma/s\n", t1, flops / t1 / 1e9);
    printf("scalar FADD  : %.3f ms  %.2f Tadd/s\n", t2, flops / t2 / 1e9);
    printf("packed FADD2 : %.3f ms  %.2f Tadd/s\n", t3, flops / t3 / 1e9);
    return 0;
}
