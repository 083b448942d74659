// K3 — post-inference scoring: activation -> top-k with confidence threshold -> range / species
// mask (keep / drop / rerank) -> second confidence threshold, one CTA per score row.
//
// Replaces the tail of birdnet_onnx::Classifier::predict_batch* (top_k / min_confidence set at
// src/inference/classifier.rs:269-273), BirdClassifier::apply_range_filter
// (src/inference/classifier.rs:587-645 -> src/inference/geomodel_filter.rs:45-79) and the
// `pred.confidence >= min_confidence` test at src/pipeline/processor.rs:374.  Order of the steps
// is the reference's: the mask sees only the already truncated top-k list.
//
// HBM-bound: B*C*4 bytes read once, B*top_k*8 + B*4 written.  Per-thread register top-k lists
// (each thread scans its elements in increasing class index, so a tie never displaces an
// earlier entry), merged with warp-shuffle arg-max rounds; ties resolve to the lower index.
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

template <int ACT>
__global__ void __launch_bounds__(kThreads)
post_kernel(const float* __restrict__ scores, uint32_t C, bb_post_cfg cfg,
            const float* __restrict__ mask, const uint8_t* __restrict__ keep,
            uint32_t* __restrict__ o_index, float* __restrict__ o_conf, uint32_t* __restrict__ o_count) {
    __shared__ float    s_red[kWarps];
    __shared__ float    s_wc[kWarps];
    __shared__ uint32_t s_wi[kWarps];
    __shared__ float    s_bc;
    __shared__ uint32_t s_bi;
    __shared__ Cand     s_win[K];

    const uint32_t row = blockIdx.x;
    const float* __restrict__ x = scores + (uint64_t)row * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t topk = cfg.top_k;
    const float min_conf = cfg.min_confidence;

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
    // coarse reject in the score domain (sigmoid only): anything this far below logit(min_conf)
    // cannot reach min_conf, so its activation is never evaluated.
    float coarse = -FLT_MAX;
    if (ACT == BB_ACT_SIGMOID && min_conf > 0.f && min_conf < 1.f)
        coarse = logf(min_conf / (1.0f - min_conf)) - 0.01f;
    if (ACT == BB_ACT_NONE) coarse = min_conf;     // conf == score; NaN min_conf rejects all below

    Cand loc[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { loc[k].conf = -FLT_MAX; loc[k].idx = 0xFFFFFFFFu; }
    int nloc = 0;
    float kth = -FLT_MAX;          // confidence of the list's last slot (loc[topk-1]) kept in a register

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
            if (i >= C) continue;
            const float xv = v[u];
            if (!(xv >= coarse)) continue;
            float c;
            if (ACT == BB_ACT_SIGMOID) c = sigmoidf_(xv);
            else if (ACT == BB_ACT_SOFTMAX) c = expf(xv - row_max) * inv_sum;
            else c = xv;
            if (!(c >= min_conf)) continue;
            // later index never beats an equal confidence already in the list
            if (nloc == (int)topk && !(c > kth)) continue;
            // insert keeping the list sorted descending (stable)
            Cand cur{c, i};
#pragma unroll
            for (int k = 0; k < K; ++k) {
                if (k < (int)topk && cur.conf > loc[k].conf) { Cand t = loc[k]; loc[k] = cur; cur = t; }
            }
            if (nloc < (int)topk) ++nloc;
#pragma unroll
            for (int k = 0; k < K; ++k) if (k == (int)topk - 1) kth = loc[k].conf;
        }
    }

    // merge: top_k rounds of block-wide arg-max over the list heads
    int head = 0;
    for (uint32_t r = 0; r < topk; ++r) {
        float c = -FLT_MAX; uint32_t ix = 0xFFFFFFFFu;
#pragma unroll
        for (int k = 0; k < K; ++k) if (k == head && k < nloc) { c = loc[k].conf; ix = loc[k].idx; }
        float bc = c; uint32_t bi = ix;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (better(oc, oi, bc, bi)) { bc = oc; bi = oi; }
        }
        if (lane == 0) { s_wc[warp] = bc; s_wi[warp] = bi; }
        __syncthreads();
        if (tid == 0) {
            float gc = s_wc[0]; uint32_t gi = s_wi[0];
#pragma unroll
            for (int w = 1; w < kWarps; ++w) if (better(s_wc[w], s_wi[w], gc, gi)) { gc = s_wc[w]; gi = s_wi[w]; }
            s_bc = gc; s_bi = gi;
            s_win[r].conf = gc; s_win[r].idx = gi;
        }
        __syncthreads();
        if (s_bi != 0xFFFFFFFFu && ix == s_bi) ++head;     // the owner pops its head
    }

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
