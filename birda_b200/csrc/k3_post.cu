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

constexpr int kMaxWarps = 8;    // widest CTA: 256 threads
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
template <int NW>
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
    for (int w = 1; w < NW; ++w) if (better(s_wc[w], s_wi[w], c, ix)) { c = s_wc[w]; ix = s_wi[w]; }
    __syncthreads();
}

// Row scan with 128-bit loads: scalars up to the first 16-byte boundary, float4 body (kVec loads in flight per
// thread), scalar tail.  f(i, v, valid) is called by ALL threads the same number of times (warp-convergent: the
// callers use ballots), with valid == false for the slots past the end.
constexpr int kVec = 8;
template <int NT, class F>
__device__ __forceinline__ void scan_row(const float* __restrict__ x, uint32_t C, int tid, F&& f) {
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(x) >> 2) & 3u);
    const uint32_t head = min(C, (4u - mis) & 3u);
    const uint32_t nvec = (C - head) >> 2;
    const uint32_t tail0 = head + (nvec << 2);
    {   // head (< 4) and tail (< 4) scalars: one convergent slot per thread
        const uint32_t i = (uint32_t)tid < head ? (uint32_t)tid : tail0 + ((uint32_t)tid - head);
        const bool valid = (uint32_t)tid < head || ((uint32_t)tid >= head && i < C && (uint32_t)tid < head + 4u);
        f(i, valid ? __ldg(x + i) : 0.f, valid);
    }
    const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + head);
    for (uint32_t v0 = 0; v0 < nvec; v0 += NT * kVec) {
        float4 q[kVec];
#pragma unroll
        for (int u = 0; u < kVec; ++u) {
            const uint32_t vi = v0 + u * NT + tid;
            q[u] = vi < nvec ? __ldg(xv + vi) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kVec; ++u) {
            const uint32_t vi = v0 + u * NT + tid;
            const bool valid = vi < nvec;
            const uint32_t i = head + (vi << 2);
            f(i, q[u].x, valid); f(i + 1, q[u].y, valid); f(i + 2, q[u].z, valid); f(i + 3, q[u].w, valid);
        }
    }
}

template <int ACT, int NT>
__global__ void __launch_bounds__(NT)
post_kernel(const float* __restrict__ scores, uint32_t C, bb_post_cfg cfg,
            const float* __restrict__ mask, const uint8_t* __restrict__ keep,
            uint32_t* __restrict__ o_index, float* __restrict__ o_conf, uint32_t* __restrict__ o_count) {
    constexpr int NW = NT / 32;
    __shared__ float    s_red[NW];
    __shared__ float    s_wc[NW];
    __shared__ uint32_t s_wi[NW];
    __shared__ Cand     s_win[K];
    __shared__ Cand     s_list[kCap];
    __shared__ Cand     s_fin[NW * K];
    __shared__ float    s_mask[K];
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
        scan_row<NT>(x, C, tid, [&](uint32_t, float v, bool valid) { if (valid) m = fmaxf(m, v); });
        m = warp_max(m);
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
        m = s_red[0];
#pragma unroll
        for (int w = 1; w < NW; ++w) m = fmaxf(m, s_red[w]);
        __syncthreads();
        float s = 0.f;
        scan_row<NT>(x, C, tid, [&](uint32_t, float v, bool valid) { if (valid) s += expf(v - m); });
        s = warp_sum(s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_red[w];
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

    // Survivors of one batch of loads are counted per thread, placed by a warp prefix sum and ONE shared-memory
    // atomic per warp, then written from the registers they were loaded into.
    {
        const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(x) >> 2) & 3u);
        const uint32_t head = min(C, (4u - mis) & 3u);
        const uint32_t nvec = (C - head) >> 2;
        const uint32_t tail0 = head + (nvec << 2);
        const uint32_t hi = (uint32_t)tid < head ? (uint32_t)tid : tail0 + ((uint32_t)tid - head);
        const bool hvalid = (uint32_t)tid < head || ((uint32_t)tid < head + 4u && hi < C);
        const float hv = hvalid ? __ldg(x + hi) : 0.f;
        const float4* __restrict__ xv = reinterpret_cast<const float4*>(x + head);
        for (uint32_t v0 = 0; v0 == 0 || v0 < nvec; v0 += NT * kVec) {
            float4 q[kVec];
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                const uint32_t vi = v0 + u * NT + tid;
                q[u] = vi < nvec ? __ldg(xv + vi) : make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
            }
            const bool hp = v0 == 0 && hvalid && hv >= coarse;
            int cnt = hp ? 1 : 0;
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                const bool in = v0 + u * NT + tid < nvec;
                cnt += (in && q[u].x >= coarse) + (in && q[u].y >= coarse) + (in && q[u].z >= coarse) + (in && q[u].w >= coarse);
            }
            int inc = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
            const int wtotal = __shfl_sync(0xffffffffu, inc, 31);
            if (wtotal == 0) continue;                       // warp-uniform
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_count, wtotal);
            base = __shfl_sync(0xffffffffu, base, 31);
            int pos = base + inc - cnt;
            if (cnt == 0) continue;
            if (hp) { if (pos < kCap) { s_list[pos].conf = hv; s_list[pos].idx = hi; } ++pos; }
#pragma unroll
            for (int u = 0; u < kVec; ++u) {
                const uint32_t vi = v0 + u * NT + tid;
                if (vi >= nvec) continue;
                const uint32_t i = head + (vi << 2);
                const float e[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4)
                    if (e[c4] >= coarse) { if (pos < kCap) { s_list[pos].conf = e[c4]; s_list[pos].idx = i + c4; } ++pos; }
            }
        }
    }
    __syncthreads();
    const int n = s_count;
    if (n <= kCap) {
        // activate the survivors in place; each warp then extracts the top_k of ITS share of the list (entries
        // tid, tid + NT, ...) by arg-max rounds, each round restricted to what comes after the previous winner in
        // the total order (conf desc, idx asc); the NW * top_k finalists are ranked against each other
        for (int t = tid; t < n; t += NT) s_list[t].conf = act(s_list[t].conf);
        __syncthreads();
        float pc = FLT_MAX; uint32_t pi = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            float bc = -FLT_MAX; uint32_t bi = 0xFFFFFFFFu;
            for (int t = tid; t < n; t += NT) {
                const float c = s_list[t].conf; const uint32_t i = s_list[t].idx;
                if (c >= min_conf && better(pc, pi, c, i) && better(c, i, bc, bi)) { bc = c; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float oc = __shfl_xor_sync(0xffffffffu, bc, o);
                const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (better(oc, oi, bc, bi)) { bc = oc; bi = oi; }
            }
            if (lane == 0) { s_fin[warp * K + r].conf = bc; s_fin[warp * K + r].idx = bi; }
            pc = bc; pi = bi;
        }
        __syncthreads();
        const int nf = NW * (int)topk;
        if (tid < nf) {
            const Cand me = s_fin[(tid / (int)topk) * K + tid % (int)topk];
            if (me.idx != 0xFFFFFFFFu) {
                uint32_t rank = 0;
                for (int o = 0; o < nf; ++o) {
                    const Cand ot = s_fin[(o / (int)topk) * K + o % (int)topk];
                    rank += (ot.idx != 0xFFFFFFFFu && better(ot.conf, ot.idx, me.conf, me.idx)) ? 1u : 0u;
                }
                if (rank < topk) s_win[rank] = me;
            }
        }
    } else {
        // list overflow (tiny min_conf): top_k rounds of block-wide arg-max over the whole row, each round
        // restricted to candidates strictly after the previous winner in the total order
        float pc = FLT_MAX; uint32_t pi = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            float bc = -FLT_MAX; uint32_t bi = 0xFFFFFFFFu;
            for (uint32_t i = tid; i < C; i += NT) {
                const float xv = __ldg(x + i);
                if (!(xv >= coarse)) continue;
                const float c = act(xv);
                if (!(c >= min_conf)) continue;
                if (r > 0 && !better(pc, pi, c, i)) continue;
                if (better(c, i, bc, bi)) { bc = c; bi = i; }
            }
            block_argmax<NW>(bc, bi, s_wc, s_wi, lane, warp);
            if (bi == 0xFFFFFFFFu) break;
            if (tid == 0) { s_win[r].conf = bc; s_win[r].idx = bi; }
            pc = bc; pi = bi;
        }
    }
    __syncthreads();
    // the winners' mask entries are fetched in parallel; one thread then walks the short list
    if (tid < K && mask != nullptr && s_win[tid].idx != 0xFFFFFFFFu) s_mask[tid] = __ldg(mask + s_win[tid].idx);
    __syncthreads();

    if (tid == 0) {
        Cand out[K];
        uint32_t n = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            Cand w = s_win[r];
            if (w.idx == 0xFFFFFFFFu) break;
            if (mask != nullptr) {                                  // geomodel_filter.rs:54-71
                const float s = s_mask[r];
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
    // one CTA per row.  Many rows: 64-thread CTAs (16+ rows resident per SM, the whole grid in one wave, enough
    // loads in flight chip-wide); few rows (one inference batch): 256-thread CTAs so that a row's loads are all
    // in flight at once.
    const bool narrow = valid_B >= 1024;
#define BB_POST(ACT)                                                                                               \
    if (narrow) post_kernel<ACT, 64><<<valid_B, 64, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); \
    else post_kernel<ACT, 256><<<valid_B, 256, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count);
    switch (cfg.activation) {
        case BB_ACT_NONE:    BB_POST(BB_ACT_NONE) break;
        case BB_ACT_SIGMOID: BB_POST(BB_ACT_SIGMOID) break;
        case BB_ACT_SOFTMAX: BB_POST(BB_ACT_SOFTMAX) break;
        default: return cudaErrorInvalidValue;
    }
#undef BB_POST
    return cudaGetLastError();
}

}  // namespace bb
