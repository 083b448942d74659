// K3 — post-inference scoring: activation -> top-k with confidence threshold -> range / species
// mask (keep / drop / rerank) -> second confidence threshold, one CTA per score row.
//
// Replaces the tail of birdnet_onnx::Classifier::predict_batch* (top_k / min_confidence set at
// src/inference/classifier.rs:269-273), BirdClassifier::apply_range_filter
// (src/inference/classifier.rs:587-645 -> src/inference/geomodel_filter.rs:45-79) and the
// `pred.confidence >= min_confidence` test at src/pipeline/processor.rs:374.  Order of the steps
// is the reference's: the mask sees only the already truncated top-k list.
//
// B*C*4 bytes read once (softmax: twice), B*top_k*8 + B*4 written.  The row is filtered in the SCORE domain against a
// threshold derived from the row itself: the top_k-th largest per-thread maximum of the first batch of loads (softmax: of
// the max pass) is a score that top_k elements reach, lowered past the activation's rounding plateau (`lowered`) and
// never below logit(min_conf); a float4 is looked into only when its maximum clears the threshold, so the scan costs about
// one instruction per element.  The handful of survivors go to a shared-memory list, are activated there and ranked
// against each other under the total order (confidence desc, class index asc); a list overflow (no usable threshold)
// falls back to top_k rounds of block-wide arg-max.
#include "common.cuh"
#include <cfloat>
#include <cstdlib>

namespace bb {
namespace {

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

constexpr int kCap = 1024;     // candidate list capacity per row (scores that clear the row's threshold)

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

// Row geometry for 128-bit loads: up to 3 scalars before the first 16-byte boundary, nvec float4, up to 3 scalars after.
// The scalars are ONE slot per thread (threads [0, head) take the head, [head, head + 4) the tail).
constexpr int kVec = 8;         // float4 loads in flight per thread
struct RowGeom {
    const float4* xv; uint32_t head, nvec; uint32_t hi; bool hvalid;
    __device__ __forceinline__ RowGeom(const float* x, uint32_t C, int tid) {
        const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(x) >> 2) & 3u);
        head = min(C, (4u - mis) & 3u);
        nvec = (C - head) >> 2;
        const uint32_t tail0 = head + (nvec << 2);
        hi = (uint32_t)tid < head ? (uint32_t)tid : tail0 + ((uint32_t)tid - head);
        hvalid = (uint32_t)tid < head || ((uint32_t)tid < head + 4u && hi < C);
        xv = reinterpret_cast<const float4*>(x + head);
    }
};
// slots past the end of the row read as -inf: never a maximum and exp() of it is 0; `collect` checks the bounds of the
// (rare) vectors it looks into, so a -inf threshold (everything survives, real -inf scores included) stays exact
// FULL: the whole batch lies inside the row (CTA-uniform): plain loads, no bounds arithmetic
template <int NT, bool FULL>
__device__ __forceinline__ void load_vecs(float4 (&q)[kVec], const RowGeom& g, uint32_t v0, int tid) {
#pragma unroll
    for (int u = 0; u < kVec; ++u) {
        const uint32_t vi = v0 + u * NT + tid;
        if (FULL) q[u] = __ldg(g.xv + vi);
        else q[u] = vi < g.nvec ? __ldg(g.xv + vi) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
}
template <int NT>
__device__ __forceinline__ void load_vecs(float4 (&q)[kVec], const RowGeom& g, uint32_t v0, int tid, bool full) {
    if (full) load_vecs<NT, true>(q, g, v0, tid); else load_vecs<NT, false>(q, g, v0, tid);
}

__device__ __forceinline__ float max4(const float4& q) { return fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w)); }

// The k-th largest and the largest of the CTA's per-thread values.  Lanes that tie with a round's maximum leave together,
// so the "k-th largest" may come out LOWER than the true one, never higher: at least k threads hold a value >= kth.
template <int NW>
__device__ __forceinline__ void block_kth_max(float v, int k, float* s_k, float* s_out, int tid, int lane, int warp,
                                              float& kth, float& top) {
    float cur = v;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        if (r < k) {                                          // k is CTA-uniform
            const float m = warp_max(cur);
            if (lane == 0) s_k[warp * k + r] = m;
            if (cur == m) cur = -FLT_MAX;
        }
    }
    __syncthreads();
    const int ne = NW * k;
    if (tid < ne) {
        const float me = s_k[tid];
        int rank = 0;
        for (int o = 0; o < ne; ++o) {
            const float ot = s_k[o];
            rank += (ot > me || (ot == me && o < tid)) ? 1 : 0;
        }
        if (rank == k - 1) s_out[0] = me;
        if (rank == 0) s_out[1] = me;
    }
    __syncthreads();
    kth = s_out[0]; top = s_out[1];
}

// Survivor threshold from a score t that at least top_k elements of the row reach: every element whose CONFIDENCE can
// tie with or beat act(t) must survive, so t is lowered past the plateau the activation rounds to around it (0.02 in the
// score domain moves a sigmoid below 8 by > 100 ulp and a softmax term by 2 %); where the confidence underflows towards
// 0 the order is decided by the class index alone and no threshold is safe: -inf (everything survives; a long row's list
// overflows and the row takes the arg-max path).
template <int ACT>
__device__ __forceinline__ float lowered(float t, float row_max) {
    if (ACT == BB_ACT_SIGMOID) return t < -80.f ? -INFINITY : fminf(t, 8.f) - 0.02f;
    if (ACT == BB_ACT_SOFTMAX) return t - row_max < -60.f ? -INFINITY : t - 0.02f;
    return t;
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
    __shared__ float    s_k[NW * K];
    __shared__ float    s_kout[2];
    __shared__ Cand     s_win[K];
    __shared__ Cand     s_list[kCap];
    __shared__ Cand     s_fin[NW * K];
    __shared__ int      s_count;

    const uint32_t row = blockIdx.x;
    const float* __restrict__ x = scores + (uint64_t)row * C;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t topk = cfg.top_k;
    const float min_conf = cfg.min_confidence;
    if (tid == 0) s_count = 0;
    if (tid < K) { s_win[tid].conf = 0.f; s_win[tid].idx = 0xFFFFFFFFu; }
    __syncthreads();

    const RowGeom g(x, C, tid);
    const float hv = g.hvalid ? __ldg(x + g.hi) : -INFINITY;
    // Survivors are rare once the threshold is row-derived (a handful per row): each is appended with its own
    // shared-memory atomic, and a float4 is looked into only when its maximum clears the threshold.
    auto push = [&](float v, uint32_t i) {
        const int pos = atomicAdd(&s_count, 1);
        if (pos < kCap) { s_list[pos].conf = v; s_list[pos].idx = i; }
    };
    auto collect = [&](const float4 (&q)[kVec], uint32_t v0, float thr, bool full) {
        unsigned hm = 0;
#pragma unroll
        for (int u = 0; u < kVec; ++u) hm |= (max4(q[u]) >= thr ? 1u : 0u) << u;
        if (hm == 0) return;
#pragma unroll
        for (int u = 0; u < kVec; ++u) {
            if (!((hm >> u) & 1u) || (!full && v0 + u * NT + tid >= g.nvec)) continue;
            const uint32_t i = g.head + ((v0 + u * NT + tid) << 2);
            if (q[u].x >= thr) push(q[u].x, i);
            if (q[u].y >= thr) push(q[u].y, i + 1);
            if (q[u].z >= thr) push(q[u].z, i + 2);
            if (q[u].w >= thr) push(q[u].w, i + 3);
        }
    };
    constexpr uint32_t kStep = NT * kVec;

    float row_max = 0.f, inv_sum = 1.f;
    // coarse threshold in the SCORE domain: anything below cannot reach min_conf (activations are monotone)
    float coarse = -INFINITY;
    if (ACT == BB_ACT_SIGMOID && min_conf > 0.f && min_conf < 1.f) coarse = logf(min_conf / (1.0f - min_conf)) - 0.01f;
    if (ACT == BB_ACT_NONE) coarse = min_conf;
    if (ACT == BB_ACT_SOFTMAX) {
        // pass 1: row maximum; the per-thread maxima also give a score that top_k elements reach
        float tm = fmaxf(-FLT_MAX, hv);
        for (uint32_t v0 = 0; v0 < g.nvec; v0 += kStep) {
            float4 q[kVec];
            load_vecs<NT>(q, g, v0, tid, v0 + kStep <= g.nvec);
#pragma unroll
            for (int u = 0; u < kVec; ++u) tm = fmaxf(tm, max4(q[u]));
        }
        float kth, m;
        block_kth_max<NW>(tm, (int)topk, s_k, s_kout, tid, lane, warp, kth, m);
        const float thr = kth <= -FLT_MAX ? -INFINITY : lowered<ACT>(kth, m);     // fewer than top_k threads saw a finite score
        // pass 2: sum of exp(x - max), survivors collected on the way (min_conf is applied to the confidences below)
        float s = g.hvalid ? expf(hv - m) : 0.f;
        if (g.hvalid && hv >= thr) push(hv, g.hi);
        for (uint32_t v0 = 0; v0 < g.nvec; v0 += kStep) {
            float4 q[kVec];
            const bool full = v0 + kStep <= g.nvec;
            load_vecs<NT>(q, g, v0, tid, full);
#pragma unroll
            for (int u = 0; u < kVec; ++u) s += (expf(q[u].x - m) + expf(q[u].y - m)) + (expf(q[u].z - m) + expf(q[u].w - m));
            collect(q, v0, thr, full);
        }
        s = warp_sum(s);
        if (lane == 0) s_red[warp] = s;
        __syncthreads();
        s = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += s_red[w];
        row_max = m; inv_sum = 1.0f / s;
        if (min_conf > 0.f) coarse = row_max + logf(min_conf / inv_sum) - 0.01f;
    } else {
        // one pass: the first batch of loads (a third of a BirdNET row with 64 threads, the whole row with 256) yields the
        // score that top_k of ITS elements reach, the batch is then filtered from the registers it sits in
        float4 q[kVec];
        float thr = 0.f;
        for (uint32_t v0 = 0; v0 == 0 || v0 < g.nvec; v0 += kStep) {      // ONE copy of the loop body: the kernel stays small
            const bool full = v0 + kStep <= g.nvec;                        // CTA-uniform
            load_vecs<NT>(q, g, v0, tid, full);
            if (v0 == 0) {
                float tm = fmaxf(-FLT_MAX, hv);
#pragma unroll
                for (int u = 0; u < kVec; ++u) tm = fmaxf(tm, max4(q[u]));
                float kth, top;
                block_kth_max<NW>(tm, (int)topk, s_k, s_kout, tid, lane, warp, kth, top);
                // (sentinel: fewer than top_k threads saw a finite score)
                thr = fmaxf(coarse, kth <= -FLT_MAX ? -INFINITY : lowered<ACT>(kth, 0.f));
                if (g.hvalid && hv >= thr) push(hv, g.hi);
            }
            collect(q, v0, thr, full);
        }
    }
    auto act = [&](float xv) -> float {
        if (ACT == BB_ACT_SIGMOID) return sigmoidf_(xv);
        if (ACT == BB_ACT_SOFTMAX) return expf(xv - row_max) * inv_sum;
        return xv;
    };
    __syncthreads();
    const int n = s_count;
    if (n <= NT) {
        // short list (the usual case): one entry per thread, ranked against the others under the total order
        Cand me; me.conf = 0.f; me.idx = 0xFFFFFFFFu;
        if (tid < n) {
            me = s_list[tid];
            me.conf = act(me.conf);
            if (!(me.conf >= min_conf)) me.idx = 0xFFFFFFFFu;
            s_list[tid] = me;
        }
        __syncthreads();
        if (me.idx != 0xFFFFFFFFu) {
            uint32_t rank = 0;
            for (int o = 0; o < n; ++o) {
                const Cand ot = s_list[o];
                rank += (ot.idx != 0xFFFFFFFFu && better(ot.conf, ot.idx, me.conf, me.idx)) ? 1u : 0u;
            }
            if (rank < topk) s_win[rank] = me;
        }
    } else if (n <= kCap) {
        // activate the survivors in place; each warp then extracts the top_k of ITS share of the list (entries
        // tid, tid + NT, ...) by arg-max rounds, each round restricted to what comes after the previous winner in
        // the total order (conf desc, idx asc); the NW * top_k finalists are ranked against each other
        for (int t = tid; t < n; t += NT) s_list[t].conf = act(s_list[t].conf);
        __syncthreads();
        float pc = FLT_MAX; uint32_t pi = 0;
        for (uint32_t r = 0; r < topk; ++r) {
            float bc = -INFINITY; uint32_t bi = 0xFFFFFFFFu;
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
            float bc = -INFINITY; uint32_t bi = 0xFFFFFFFFu;
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
    // Tail, one lane per winner (the list is at most BB_MAX_TOP_K long): mask / species test, rerank, second threshold.
    if (warp == 0) {
        const bool in = (uint32_t)lane < topk;
        Cand w; w.conf = 0.f; w.idx = 0xFFFFFFFFu;
        if (in) w = s_win[lane];
        bool kept = w.idx != 0xFFFFFFFFu;                           // winners are contiguous from slot 0
        const bool rerank = mask != nullptr && cfg.rerank;
        if (kept) {
            if (mask != nullptr) {                                  // geomodel_filter.rs:54-71
                const float sc = __ldg(mask + w.idx);
                if (isnan(sc)) kept = cfg.keep_unmatched && !cfg.rerank;
                else if (sc >= cfg.range_threshold) { if (cfg.rerank) w.conf = __fmul_rn(w.conf, sc); }
                else kept = false;
            } else if (keep != nullptr) {                           // classifier.rs:616-641
                kept = keep[w.idx] != 0;
            }
        }
        // position among the kept entries: list order, or (rerank) a stable sort by confidence, geomodel_filter.rs:74-76
        const unsigned km = __ballot_sync(0xffffffffu, kept);
        int pos = __popc(km & ((1u << lane) - 1u));
        if (rerank) {
            pos = 0;
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const float cj = __shfl_sync(0xffffffffu, w.conf, j);
                pos += (((km >> j) & 1u) && (cj > w.conf || (cj == w.conf && j < lane))) ? 1 : 0;
            }
        }
        const bool pass = kept && w.conf >= min_conf;               // processor.rs:374, order preserved
        const unsigned pm = __ballot_sync(0xffffffffu, pass);
        int slot = 0;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            const int pj = __shfl_sync(0xffffffffu, pos, j);
            slot += (((pm >> j) & 1u) && pj < pos) ? 1 : 0;
        }
        const uint32_t cnt = (uint32_t)__popc(pm);
        if (pass) { o_index[(uint64_t)row * topk + slot] = w.idx; o_conf[(uint64_t)row * topk + slot] = w.conf; }
        if (in && (uint32_t)lane >= cnt) { o_index[(uint64_t)row * topk + lane] = 0xFFFFFFFFu; o_conf[(uint64_t)row * topk + lane] = 0.f; }
        if (lane == 0) o_count[row] = cnt;
    }
}

}  // namespace

cudaError_t launch_post(cudaStream_t st, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                        const bb_post_cfg& cfg, const float* d_mask, const uint8_t* d_keep,
                        uint32_t* d_index, float* d_conf, uint32_t* d_count) {
    (void)B;
    if (valid_B == 0) return cudaSuccess;
    // one CTA per row.  Many rows (a whole file's windows): 64-thread CTAs — 16+ rows resident per SM, the whole grid in
    // one wave, enough loads in flight chip-wide; a few hundred rows: 128 threads; one inference batch: 256 threads, so
    // that a row's loads are all in flight at once (measured: tools/prof_k3.py, profiles/r02_k3_*).
    // BIRDA_K3_THREADS=64|128|256 overrides the choice (A/B runs)
    const char* env = getenv("BIRDA_K3_THREADS");
    const int forced = env ? atoi(env) : 0;
    const int nt = forced == 64 || forced == 128 || forced == 256 ? forced : (valid_B >= 1024 ? 64 : valid_B >= 128 ? 128 : 256);
#define BB_POST(ACT)                                                                                               \
    if (nt == 64) post_kernel<ACT, 64><<<valid_B, 64, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); \
    else if (nt == 128) post_kernel<ACT, 128><<<valid_B, 128, 0, st>>>(d_scores, C, cfg, d_mask, d_keep, d_index, d_conf, d_count); \
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
