// K4 — tiny dense heads on the device: out = act(x W + b).
//
// Covers the shapes that today force a host hop in the reference: the geomodel forward
// ([1,3] -> [1,12012] sigmoid once per run: src/inference/classifier.rs:117-188 via
// birdnet_onnx::RangeFilter; fixture weights tests/fixtures/make_fixture_geomodel.py:20-28) and the
// bat head over backbone embeddings ([B,1024] -> [B,11..38]: src/pipeline/processor.rs:323-360).
// f32 throughout, one warp per output element with a shuffle reduction over K (left-to-right partial
// sums per lane, so results are deterministic).  Latency-bound by design (kilobytes of work).
#include "common.cuh"
#include <algorithm>
#include <cfloat>

namespace bb {
namespace {

template <int ACT>
__global__ void __launch_bounds__(256)
dense_kernel(const float* __restrict__ x, uint32_t B, uint32_t K, const float* __restrict__ W, const float* __restrict__ bias,
             uint32_t N, float* __restrict__ out) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint64_t total = (uint64_t)B * N;
    if (warp >= total) return;
    const uint32_t r = warp / N, n = warp - r * N;
    float acc = 0.f;
    for (uint32_t k = lane; k < K; k += 32) acc = fmaf(__ldg(x + (uint64_t)r * K + k), __ldg(W + (uint64_t)k * N + n), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        float v = acc + (bias ? bias[n] : 0.f);
        if (ACT == BB_ACT_SIGMOID) v = 1.0f / (1.0f + expf(-v));
        out[(uint64_t)r * N + n] = v;
    }
}

// row-wise softmax in place (small N): one warp per row
__global__ void __launch_bounds__(256)
softmax_rows_kernel(float* __restrict__ out, uint32_t B, uint32_t N) {
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= B) return;
    float* p = out + (uint64_t)row * N;
    float m = -FLT_MAX;
    for (uint32_t i = lane; i < N; i += 32) m = fmaxf(m, p[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float s = 0.f;
    for (uint32_t i = lane; i < N; i += 32) s += expf(p[i] - m);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    for (uint32_t i = lane; i < N; i += 32) p[i] = expf(p[i] - m) / s;
}

// per-class affine map of the scores: out[b, c] = a[c] * x[b, c] + b[c]
__global__ void __launch_bounds__(256)
affine_classes_kernel(const float* __restrict__ x, uint64_t total, uint32_t C, const float* __restrict__ a, const float* __restrict__ bias,
                      float* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i % C);
        out[i] = fmaf(__ldg(a + c), x[i], bias ? __ldg(bias + c) : 0.f);
    }
}

}  // namespace

cudaError_t launch_affine_classes(cudaStream_t st, const float* d_x, uint32_t B, uint32_t C, const float* d_a, const float* d_b, float* d_out) {
    const uint64_t total = (uint64_t)B * C;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)std::min<uint64_t>((total + 255) / 256, 148ull * 16);
    affine_classes_kernel<<<blocks, 256, 0, st>>>(d_x, total, C, d_a, d_b, d_out);
    return cudaGetLastError();
}

cudaError_t launch_dense(cudaStream_t st, const float* d_x, uint32_t B, uint32_t K, const float* d_W, const float* d_b,
                         uint32_t N, int activation, float* d_out, int* launches) {
    if (launches) *launches = 0;
    const uint64_t total = (uint64_t)B * N;
    if (total == 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total * 32 + 255) / 256);
    if (activation == BB_ACT_SIGMOID) dense_kernel<BB_ACT_SIGMOID><<<blocks, 256, 0, st>>>(d_x, B, K, d_W, d_b, N, d_out);
    else dense_kernel<BB_ACT_NONE><<<blocks, 256, 0, st>>>(d_x, B, K, d_W, d_b, N, d_out);
    if (launches) *launches = 1;
    if (activation == BB_ACT_SOFTMAX) {
        softmax_rows_kernel<<<(unsigned)(((uint64_t)B * 32 + 255) / 256), 256, 0, st>>>(d_out, B, N);
        if (launches) *launches = 2;
    }
    return cudaGetLastError();
}

}  // namespace bb
