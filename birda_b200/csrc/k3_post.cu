// K3 — post-inference scoring: activation -> top-k with confidence threshold -> range / species
// mask (keep / drop / rerank) -> second confidence threshold, one CTA per score row.
//
// Replaces the tail of birdnet_onnx::Classifier::predict_batch* (top_k / min_confidence set at
// src/inference/classifier.rs:269-273), BirdClassifier::apply_range_filter
// (src/inference/classifier.rs:587-645 -> src/inference/geomodel_filter.rs:45-79) and the
// `pred.confidence >= min_confidence` test at src/pipeline/processor.rs:374.  Order of the steps
// is the reference's: the mask sees only the already truncated top-k list.
//
// B*C*4 bytes read once, B*top_k*8 + B*4 written.  The scan is load + compare against a coarse
// threshold in the score domain; survivors go to a shared-memory candidate list, are activated there
// and ranked against each other under the total order (confidence desc, class index asc); a list
// overflow falls back to top_k rounds of block-wide arg-max (warp shuffles).
#include "common.cuh"
#include <cfloat>

namespace bb {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps   = kThreads / 32;
constexpr int K        = BB_MAX_TOP_K;

struct Cand { float conf; uint32_t idx; };

__device__ __forceinline__ bool better(float ca, uint32_t ia, float cb, uint32_t ib) {
    return ca > cb || (ca == cb && ia < ib);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

constexpr int kCap = 1024;     // candidate list capacity per row (scores that clear the coarse threshold)

// block-wide arg-max under the total order (conf desc, idx asc)
__device__ __forceinline__ void block_argmax(float& c, uint32_t& ix, float* s_wc, uint32_t* s_wi, int lane, int warp) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oc = __shfl_xor_sync(0xffffffffu, c, o);
        const uint32_t oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (better(oc, oi, c, ix)) { c = oc; ix = oi; }
    }
    if (lane == 0) { s_wc[warp] = c; s_wi[warp] = ix; }
    __syncthreads();
    c = s_wc[0]; ix = s_wi[0];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) if (better(s_wc[w], s_wi[w], c, ix)) { c = s_wc[w]; ix = s_wi[w]; }
    __syncthreads();
}

template <int ACT>
__global__ void __launch_bounds__(kThreads)
post_kernel(const float* __restrict__ scores, uint32_t C, bb_post_cfg cfg,
            const float* __restrict__ mask, const uint8_t* __restrict__ keep,
            uint32_t* __restrict__ o_index, float* __restrict__ o_conf, uint32_t* __restrict__ o_count) {
    __shared__ float    s_red[kWarps];
    __shared__ float    s_wc[kWarps];
    __shared__ uint32_t s_wi[kWarps];
    __shared__ Cand     s_win[K];
    __shared__ Cand     s_list[kCap];
    __shared__ int      s_count;

    const uint32_t row = blockIdx.x;
    const float* __restrict__ x = scores + (uint64_t)row * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t topk = cfg.top_k;
    const float min_conf = cfg.min_confidence;
    if (tid == 0) s_count = 0;
    if (tid < K) { s_win[tid].conf = 0.f; s_win[tid].idx = 0xFFFFFFFFu; }

    float row_max = 0.f, inv_sum = 1.f;
    if (ACT == BB_ACT_SOFTMAX) {
        float m = -FLT_MAX;
        for (uint32_t i = tid; i < C; i += kThreads) m = fmaxf(m, x[i]);
        m = warp_max(m);
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
        m = s_red[0];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) m = fmaxf(m, s_red[w]);
        __syncthreads();
        float s = 0.f;
        for (uint32_t i = tid; i < C; i += kThreads) s += expf(x[i] - m);
        s = warp_sum(s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        s = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += s_red[w];
        row_max = m; inv_sum = 1.0f / s;
    }
    auto act = [&](float xv) -> float {
        if (ACT == BB_ACT_SIGMOID) return sigmoidf_(xv);
        if (ACT == BB_ACT_SOFTMAX) return expf(xv - row_max) * inv_sum;
        return xv;
    };
    // coarse threshold in the SCORE domain: anything below cannot reach min_conf (activations are monotone),
    // so the hot loop is load + compare and only survivors are appended to the candidate list.
    float coarse = -FLT_MAX;
    if (ACT == BB_ACT_SIGMOID && min_conf > 0.f && min_conf < 1.f) coarse = logf(min_conf / (1.0f - min_conf)) - 0.01f;
    if (ACT == BB_ACT_SOFTMAX && min_conf > 0.f) coarse = row_max + logf(min_conf / inv_sum) - 0.01f;
    if (ACT == BB_ACT_NONE) coarse = min_conf;
    __syncthreads();

    constexpr int U = 8;
    for (uint32_t base = tid; base < C; base += kThreads * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = base + u * kThreads;
            v[u] = i < C ? __ldg(x + i) : -FLT_MAX;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t i = base + u * kThreads;
            if (i < C && v[u] >= coarse) {
                const int pos = atomicAdd(&s_count, 1);
                if (pos < kCap) { s_list[pos].conf = v[u]; s_list[pos].idx = i; }
            }
        }
    }
    __syncthreads();
    const int n = s_count;
    if (n <= kCap) {
        // activate the survivors, then rank each one against the others (total order: conf desc, idx asc)
        for (int t = tid; t < n; t += kThreads) s_list[t].conf = act(s_list[t].conf);
        __syncthreads();
        for (int t = tid; t < n; t += kThreads) {
            const float c = s_list[t].conf; const uint32_t ix = s_list[t].idx;
            if (!(c >= min_conf)) continue;
            uint32_t rank = 0;
            for (int o = 0; o < n; ++o) rank += better(s_list[o].conf, s_list[o].idx, c, ix) ? 1u : 0u;
            if (rank < topk) { s_win[rank].conf = c; s_win[rank].idx = ix; }
        }
    } else {
        // list overflow (tiny min_conf): top_k rounds of block-wide arg-max over the whole row, each round
        // restricted to candidates strictly after the previous winner in the total order
        float pc = FLT_MAX; uint32_t pi = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            float bc = -FLT_MAX; uint32_t bi = 0xFFFFFFFFu;
            for (uint32_t i = tid; i < C; i += kThreads) {
                const float xv = __ldg(x + i);
                if (!(xv >= coarse)) continue;
                const float c = act(xv);
                if (!(c >= min_conf)) continue;
                if (r > 0 && !better(pc, pi, c, i)) continue;
                if (better(c, i, bc, bi)) { bc = c; bi = i; }
            }
            block_argmax(bc, bi, s_wc, s_wi, lane, warp);
            if (bi == 0xFFFFFFFFu) break;
            if (tid == 0) { s_win[r].conf = bc; s_win[r].idx = bi; }
            pc = bc; pi = bi;
        }
    }
    __syncthreads();

    if (tid == 0) {
        Cand out[K];
        uint32_t n = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            Cand w = s_win[r];
            if (w.idx == 0xFFFFFFFFu) break;
            if (mask != nullptr) {                                  // geomodel_filter.rs:54-71
                const float s = mask[w.idx];
                if (isnan(s)) { if (!(cfg.keep_unmatched && !cfg.rerank)) continue; }
                else if (s >= cfg.range_threshold) { if (cfg.rerank) w.conf = __fmul_rn(w.conf, s); }
                else continue;
            } else if (keep != nullptr) {                           // classifier.rs:616-641
                if (!keep[w.idx]) continue;
            }
            out[n++] = w;
        }
        if (mask != nullptr && cfg.rerank) {                        // geomodel_filter.rs:74-76 (stable here)
            for (uint32_t a = 1; a < n; ++a) {
                Cand t = out[a]; int b = (int)a - 1;
                while (b >= 0 && out[b].conf < t.conf) { out[b + 1] = out[b]; --b; }
                out[b + 1] = t;
            }
        }
        uint32_t m = 0;
        for (uint32_t a = 0; a < n; ++a) {                          // processor.rs:374
            if (out[a].conf >= min_conf) {
                o_index[(uint64_t)row * topk + m] = out[a].idx;
                o_conf[(uint64_t)row * topk + m]  = out[a].conf;
                ++m;
            }
        }
        o_count[row] = m;
        for (; m < topk; ++m) { o_index[(uint64_t)row * topk + m] = 0xFFFFFFFFu; o_conf[(uint64_t)row * topk + m] = 0.f; }
    }
}

}  // namespace

cudaError_t launch_post(cudaStream_t st, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                        const bb_post_cfg& cfg, const float* d_mask, const uint8_t* d_keep,
                        uint32_t* d_index, float* d_conf, uint32_t* d_count) {
    (void)B;
    if (valid_B == 0) return cudaSuccess;
    switch (cfg.activation) {
        case BB_ACT_NONE:    post_kernel<BB_ACT_NONE><<<valid_B, kThreads, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); break;
        case BB_ACT_SIGMOID: post_kernel<BB_ACT_SIGMOID><<<valid_B, kThreads, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); break;
        case BB_ACT_SOFTMAX: post_kernel<BB_ACT_SOFTMAX><<<valid_B, kThreads, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace bb
