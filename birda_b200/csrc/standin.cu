// Stand-in classifier for benches and tests (product library, but NOT a model and not on the reference's path):
// BirdClassifier::predict_batch's I/O contract ([B, samples] f32 windows -> [B, C] f32 logits, device to device;
// src/inference/classifier.rs:469-582) with trivial arithmetic, so that batching, padding, the post step, the
// multi-GPU pool and "front end + inference" timings can be exercised without ONNX Runtime (absent from this image).
// Per window: log10 of the mean square of 48 equal frames, times a fixed pseudo-random [48, C] matrix, plus a bias
// around -6 (a few classes per window clear a 0.1 confidence threshold after the sigmoid).
#include "common.cuh"
#include "guard.hpp"
#include <cmath>
#include <vector>

namespace {
constexpr int kBands = 48;

__global__ void __launch_bounds__(256) standin_energy_kernel(const float* __restrict__ x, uint32_t samples, uint32_t frame, float* __restrict__ feat) {
    const uint32_t b = blockIdx.y, f = blockIdx.x;
    const float* p = x + (size_t)b * samples + (size_t)f * frame;
    float acc = 0.f;
    for (uint32_t i = threadIdx.x; i < frame; i += blockDim.x) { const float v = __ldg(p + i); acc = fmaf(v, v, acc); }
    __shared__ float red[8];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        feat[(size_t)b * kBands + f] = log10f(s / (float)frame + 1e-6f);
    }
}

__global__ void __launch_bounds__(256) standin_project_kernel(const float* __restrict__ feat, const float* __restrict__ W, const float* __restrict__ bias,
                                                              uint32_t C, float* __restrict__ out) {
    __shared__ float f[kBands];
    const uint32_t b = blockIdx.y;
    if (threadIdx.x < kBands) f[threadIdx.x] = feat[(size_t)b * kBands + threadIdx.x];
    __syncthreads();
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = __ldg(bias + c);
#pragma unroll 8
    for (int k = 0; k < kBands; ++k) acc = fmaf(f[k], __ldg(W + (size_t)k * C + c), acc);
    out[(size_t)b * C + c] = acc;
}
}  // namespace

struct bb_standin {
    int device = 0; uint32_t samples = 0, classes = 0, max_batch = 0;
    float *d_W = nullptr, *d_bias = nullptr, *d_feat = nullptr, *d_out = nullptr;
    std::vector<float> W, bias;
    // default: run on `own` and finish before returning (what bb_pool asks of its callback); bb_standin_use_stream:
    // queue asynchronously on the caller's stream (a pipeline on its own context: bb_ctx_stream)
    cudaStream_t own = nullptr, stream = nullptr; bool async = false;
    cudaEvent_t done = nullptr;                        // blocking-sync event: the calling worker thread sleeps, it does not spin
    uint64_t launches = 0;
};

extern "C" {

void bb_standin_destroy(bb_standin* s) {
    if (!s) return;
    bb::DeviceGuard g(s->device);
    for (float* p : {s->d_W, s->d_bias, s->d_feat, s->d_out}) if (p) cudaFree(p);
    if (s->own) cudaStreamDestroy(s->own);
    if (s->done) cudaEventDestroy(s->done);
    delete s;
}

int32_t bb_standin_create(int32_t device, uint32_t samples, uint32_t classes, uint32_t max_batch, uint64_t seed, bb_standin** out) {
    BB_TRY
    if (!out || samples < (uint32_t)kBands || classes == 0 || max_batch == 0) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "bad stand-in shape");
    *out = nullptr;
    bb_standin* s = new bb_standin();
    s->device = device; s->samples = samples; s->classes = classes; s->max_batch = max_batch;
    // weights: sums of uniforms from a 64-bit LCG (deterministic across hosts); W ~ N(0, 0.05^2), bias ~ N(-6, 1)
    uint64_t st = seed * 0x9E3779B97F4A7C15ull + 0x632BE59BD9B4E019ull;
    auto uni = [&]() { st = st * 6364136223846793005ull + 1442695040888963407ull; return (double)(st >> 11) * (1.0 / 9007199254740992.0); };
    auto gauss = [&]() { double a = 0; for (int i = 0; i < 12; ++i) a += uni(); return a - 6.0; };
    s->W.resize((size_t)kBands * classes); s->bias.resize(classes);
    for (auto& w : s->W) w = (float)(0.05 * gauss());
    for (auto& b : s->bias) b = (float)(gauss() - 6.0);
    bb::DeviceGuard g(device);
    auto fail = [&](cudaError_t e) { const std::string m = std::string("stand-in classifier: ") + cudaGetErrorString(e); bb_standin_destroy(s); bb::set_tls_error(m); return (int32_t)BB_ERR_CUDA; };
    if (g.err != cudaSuccess) return fail(g.err);
    cudaError_t e;
    if ((e = cudaMalloc(&s->d_W, s->W.size() * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&s->d_bias, s->bias.size() * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&s->d_feat, (size_t)max_batch * kBands * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMalloc(&s->d_out, (size_t)max_batch * classes * 4)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpy(s->d_W, s->W.data(), s->W.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    if ((e = cudaMemcpy(s->d_bias, s->bias.data(), s->bias.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return fail(e);
    if ((e = cudaStreamCreateWithFlags(&s->own, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    if ((e = cudaEventCreateWithFlags(&s->done, cudaEventBlockingSync | cudaEventDisableTiming)) != cudaSuccess) return fail(e);
    *out = s;
    return BB_OK;
    BB_CATCH(nullptr)
}

void bb_standin_use_stream(bb_standin* s, void* cuda_stream, int32_t on) {
    if (!s) return;
    s->async = on != 0; s->stream = on ? (cudaStream_t)cuda_stream : nullptr;
}

int32_t bb_standin_weights(const bb_standin* s, float* W, float* bias) {
    if (!s) return BB_ERR_INVALID_ARG;
    if (W) std::copy(s->W.begin(), s->W.end(), W);
    if (bias) std::copy(s->bias.begin(), s->bias.end(), bias);
    return BB_OK;
}

uint64_t bb_standin_launches(const bb_standin* s) { return s ? s->launches : 0; }

// a bb_classify_fn (user = the bb_standin*): 0 on success
int32_t bb_standin_classify(void* user, const float* d_segments, uint32_t batch, uint32_t samples, const float** d_scores, uint32_t* classes) {
    bb_standin* s = static_cast<bb_standin*>(user);
    if (!s || !d_segments || !d_scores || !classes || batch == 0 || batch > s->max_batch || samples != s->samples) return 1;
    bb::DeviceGuard g(s->device);
    if (g.err != cudaSuccess) return 2;
    cudaStream_t st = s->async ? s->stream : s->own;
    const uint32_t frame = samples / kBands;
    standin_energy_kernel<<<dim3(kBands, batch), 256, 0, st>>>(d_segments, samples, frame, s->d_feat);
    standin_project_kernel<<<dim3((s->classes + 255) / 256, batch), 256, 0, st>>>(s->d_feat, s->d_W, s->d_bias, s->classes, s->d_out);
    if (cudaGetLastError() != cudaSuccess) return 3;
    s->launches += 2;
    if (!s->async && (cudaEventRecord(s->done, st) != cudaSuccess || cudaEventSynchronize(s->done) != cudaSuccess)) return 4;
    *d_scores = s->d_out; *classes = s->classes;
    return 0;
}

}  // extern "C"
