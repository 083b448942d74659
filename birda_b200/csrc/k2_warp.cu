// K2 — convert + downmix + window gather + per-window block-FFT resample + pack (src/audio/resample.rs:10-105 over
// src/pipeline/processor.rs:61-100 and src/audio/decode.rs:150-202, 353-411), the device kernels and their launcher.
// The transform core is k2_warp.cuh; here: PCM staging (cp.async into the forward buffer's own slots), loaders with
// the reference's exact sample conversion, the work-item loop (atomic counter), sinks, and three kernels over the
// core — resample_plan2_kernel (compile-time plan, two windows per thread group, packed f32x2 math: the fast path),
// resample_plan_kernel (compile-time plan, one window) and resample_warp_kernel (runtime plan, any supported rate pair).
#include "common.cuh"
#include "k2_warp.cuh"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

namespace bb {
namespace {

using namespace bb::k2w;

constexpr int kMaxThreads = 576;     // 18 warps: register cap 112 per thread
constexpr int kMaxGroups = 15;       // one named barrier per group (ids 1..15)
constexpr size_t kSmemMax = 227 * 1024;

struct WarpParams {
    RtPlan plan;
    const void* pcm; int fmt; uint32_t channels; uint64_t total_frames;
    uint64_t src_seg, hop, nseg, last_start, rows_total, row_first, seg;
    uint32_t out_len; float* out;
    const float2 *twf, *twi, *WI; const uint4* sidx; const float4 *pq1, *pq2;     // tables in global memory
    const uint32_t *ordf, *ordi; uint32_t ordf_len, ordi_len, off_ordf, off_ordi;
    WorkItems wi; uint32_t nblk; uint64_t nitems;     // work-item plan (k2_warp.cuh: plan_work_items)
    unsigned long long* counter;
    // shared memory layout (bytes)
    uint32_t off_twi, off_sidx, off_pq1, off_pq2, off_WI, off_items, tables, per_group, off_B, off_carry;
    int groups, gw;                    // thread groups per CTA, warps per group (a group owns one block at a time)
    int dual;                          // 1: two rows per group in lockstep (packed f32x2 math), tables are float4
};

struct DevExec {
    int glane, nl, bar_id;
    __device__ __forceinline__ void sync() const {
        if (nl == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nl) : "memory");
    }
    template <class F> BB_HD void each(F&& f) const {
#ifdef __CUDA_ARCH__
        f(glane, nl);
        sync();
#endif
    }
};
// the same with the group size a compile-time constant: the stage loops of a compile-time plan then have constant trip
// counts (one predicated butterfly per lane where a stage has fewer butterflies than the group has lanes)
template <int NLC>
struct DevExecC {
    int glane, bar_id;
    static constexpr int nl = NLC;
    __device__ __forceinline__ void sync() const {
        if (NLC == 32) __syncwarp();
        else asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(NLC) : "memory");
    }
    template <class F> BB_HD void each(F&& f) const {
#ifdef __CUDA_ARCH__
        f(glane, NLC);
        sync();
#endif
    }
};

// Tail of a buffer: copy exactly the bytes that exist.  (cp.async with src-size < cp-size zero-fills the rest and reads
// only src-size bytes, but it still NAMES cp-size bytes at the source — compute-sanitizer flags that at the last bytes
// of an allocation; staged bytes past `valid` are never used, so nothing needs zero-filling.)
__device__ __forceinline__ void stage_tail(void* smem_dst, const char* gsrc, long long room, int want) {
    char* d = static_cast<char*>(smem_dst);
    int off = 0;
    if (want >= 8 && room >= 8 && ((reinterpret_cast<uintptr_t>(gsrc) | reinterpret_cast<uintptr_t>(d)) & 7) == 0) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(d);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(a), "l"(gsrc) : "memory");
        return;
    }
    for (; off + 4 <= want && off + 4 <= room; off += 4) {
        const unsigned a = (unsigned)__cvta_generic_to_shared(d + off);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(a), "l"(gsrc + off) : "memory");
    }
    if (off < want && room - off >= 2)              // one 16-bit sample left (mono s16 at the very end of the buffer)
        *reinterpret_cast<short*>(d + off) = __ldg(reinterpret_cast<const short*>(gsrc + off));
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// Sample conversion + downmix of one frame, bit-identical to decode.rs:353-411.  For S16 the integer
// channel sum is exact (|sum| < 2^24), so (sum * 2^-15) / C equals the reference's f32 sequence.
struct Source {
    const void* pcm; const char* pcm_end; int fmt; uint32_t ch; float fch;
    int kind;          // 0: S16 stereo staged in A   1: S16 mono staged in A   2: direct loads
    __device__ __noinline__ float direct(uint64_t f) const {
        if (fmt == BB_S16) {
            const short* p = reinterpret_cast<const short*>(pcm) + f * ch;
            int s = 0;
            for (uint32_t c = 0; c < ch; ++c) s += (int)__ldg(p + c);
            const float v = __fmul_rn(__int2float_rn(s), 1.0f / 32768.0f);
            return ch == 1 ? v : __fdiv_rn(v, fch);
        } else if (fmt == BB_S24) {              // packed 24-bit: (s24 << 8) through the S32 arm (decode.rs:386-402)
            const unsigned char* p = reinterpret_cast<const unsigned char*>(pcm) + f * ch * 3;
            auto ld24 = [](const unsigned char* q) {
                return (int)(((unsigned)__ldg(q) << 8) | ((unsigned)__ldg(q + 1) << 16) | ((unsigned)__ldg(q + 2) << 24));
            };
            if (ch == 1) return __fmul_rn(__int2float_rn(ld24(p)), 1.0f / 2147483648.0f);
            float s = 0.0f;
            for (uint32_t c = 0; c < ch; ++c) s = __fadd_rn(s, __fmul_rn(__int2float_rn(ld24(p + 3 * c)), 1.0f / 2147483648.0f));
            return __fdiv_rn(s, fch);
        } else if (fmt == BB_S32) {
            const int* p = reinterpret_cast<const int*>(pcm) + f * ch;
            if (ch == 1) return __fmul_rn(__int2float_rn(__ldg(p)), 1.0f / 2147483648.0f);
            float s = 0.0f;
            for (uint32_t c = 0; c < ch; ++c) s = __fadd_rn(s, __fmul_rn(__int2float_rn(__ldg(p + c)), 1.0f / 2147483648.0f));
            return __fdiv_rn(s, fch);
        } else {
            const float* p = reinterpret_cast<const float*>(pcm) + f * ch;
            if (ch == 1) return __ldg(p);
            float s = 0.0f;
            for (uint32_t c = 0; c < ch; ++c) s = __fadd_rn(s, __ldg(p + c));
            return __fdiv_rn(s, fch);
        }
    }
};

// z[n] = x[2n] + i x[2n+1] of one block; `valid` = real samples in the block (zero above)
template <class AM>
struct BlockLoader {
    const Source* src; const float2* A; uint64_t base; int valid; int shift;
    __device__ __forceinline__ float2 operator()(int n) const {
        const int i0 = 2 * n;
        float re = 0.f, im = 0.f;
        if (src->kind == 0) {                // slot n holds frames (2n, 2n+1) as two packed L|R words
            const int2 v = *reinterpret_cast<const int2*>(A + AM::at(n));
            re = __fmul_rn(__int2float_rn((int)(short)(v.x & 0xffff) + (v.x >> 16)), 1.0f / 65536.0f);
            im = __fmul_rn(__int2float_rn((int)(short)(v.y & 0xffff) + (v.y >> 16)), 1.0f / 65536.0f);
        } else if (src->kind == 1) {         // slot n holds the aligned word(s) covering frames (2n, 2n+1)
            const int2 v = *reinterpret_cast<const int2*>(A + AM::at(n));
            const int w = __funnelshift_r(v.x, v.y, shift);
            re = __fmul_rn(__int2float_rn((int)(short)(w & 0xffff)), 1.0f / 32768.0f);
            im = __fmul_rn(__int2float_rn(w >> 16), 1.0f / 32768.0f);
        } else {
            if (i0 < valid) re = src->direct(base + i0);
            if (i0 + 1 < valid) im = src->direct(base + i0 + 1);
            return make_float2(re, im);
        }
        if (i0 >= valid) re = 0.f;
        if (i0 + 1 >= valid) im = 0.f;
        return make_float2(re, im);
    }
};

// stage the raw PCM of one block into A with cp.async (A is dead between the split pass of the
// previous block and the first stage of this one).  Returns the funnel shift for mono.
template <class AM>
__device__ __forceinline__ int prefetch_block(const Source& s, float2* A, uint64_t base, int valid, int half_in, int lane, int nl) {
    int shift = 0;
    if (s.kind == 0) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + base * 4;
        for (int n = lane; n < half_in; n += nl) {
            if (2 * n >= valid) break;
            const char* g = g0 + (size_t)n * 8;
            stage_tail(A + AM::at(n), g, s.pcm_end - g, 8);
        }
    } else if (s.kind == 1) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + base * 2;
        const bool odd = (reinterpret_cast<uintptr_t>(g0) & 3) != 0;    // frame pair starts mid-word
        shift = odd ? 16 : 0;
        g0 -= odd ? 2 : 0;
        for (int n = lane; n < half_in; n += nl) {
            if (2 * n >= valid) break;
            const char* g = g0 + (size_t)n * 4;
            const long long room = s.pcm_end - g;
            stage_tail(A + AM::at(n), g, room, odd ? 8 : 4);
        }
    }
    return shift;
}

struct DevSink {
    float* p; int lim; bool vec;
    __device__ __forceinline__ void operator()(int n, float2 y) const {
        const int o = 2 * n;
        if (vec && o + 1 < lim) *reinterpret_cast<float2*>(p + o) = y;
        else { if (o < lim) p[o] = y.x; if (o + 1 < lim) p[o + 1] = y.y; }
    }
};

// plan views: the runtime plan reads sizes from the stage table, a compile-time plan folds them
template <class AM> struct RtView {
    using MapA = AM; using MapB = AM;
    static __device__ __forceinline__ int n(const RtPlan& p) { return p.N; }
    static __device__ __forceinline__ int m(const RtPlan& p) { return p.M; }
    static __device__ __forceinline__ int half_in(const RtPlan& p) { return p.half_in; }
    static __device__ __forceinline__ int adv_in(const RtPlan& p) { return p.adv_in; }
    static __device__ __forceinline__ int adv_out(const RtPlan& p) { return p.adv_out; }
    static __device__ __forceinline__ int carry_slots(const RtPlan& p) { return p.carry_slots; }
    static constexpr bool kCompactTw = false;
    template <class C, class Exec, class Loader, class Sink, class After, class TB>
    static __device__ __forceinline__ void block(const Exec& ex, const RtPlan& p, const TB& T, typename Mem<C>::T* A,
                                                 typename Mem<C>::T* B, typename Mem<C>::T* carry, const Loader& ld,
                                                 const Sink& sink, After&& after) {
        process_block<C, AM>(ex, p, T, A, B, carry, ld, sink, after);
    }
};
template <class PL> struct CtView {
    using MapA = typename PL::MapA; using MapB = typename PL::MapB;
    static __device__ __forceinline__ constexpr int n(const RtPlan&) { return PL::N; }
    static __device__ __forceinline__ constexpr int m(const RtPlan&) { return PL::M; }
    static __device__ __forceinline__ constexpr int half_in(const RtPlan&) { return PL::HALF_IN; }
    static __device__ __forceinline__ constexpr int adv_in(const RtPlan&) { return PL::ADV_IN; }
    static __device__ __forceinline__ constexpr int adv_out(const RtPlan&) { return PL::ADV_OUT; }
    static __device__ __forceinline__ constexpr int carry_slots(const RtPlan&) { return PL::CARRY; }
    static constexpr bool kCompactTw = PL::COMPACT_TW;      // two-stream kernel only
    template <class C, class Exec, class Loader, class Sink, class After, class TB>
    static __device__ __forceinline__ void block(const Exec& ex, const RtPlan& p, const TB& T, typename Mem<C>::T* A,
                                                 typename Mem<C>::T* B, typename Mem<C>::T* carry, const Loader& ld,
                                                 const Sink& sink, After&& after) {
        process_block_ct<PL, C>(ex, T, A, B, carry, ld, sink, after);
    }
};

template <class PV>
__device__ __forceinline__ void resample_body(const WarpParams& P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const RtPlan& PL = P.plan;
    float2* s_twf = reinterpret_cast<float2*>(smem);
    float2* s_twi = reinterpret_cast<float2*>(smem + P.off_twi);
    uint4* s_sidx = reinterpret_cast<uint4*>(smem + P.off_sidx);
    float4* s_pq1 = reinterpret_cast<float4*>(smem + P.off_pq1);
    float4* s_pq2 = reinterpret_cast<float4*>(smem + P.off_pq2);
    float2* s_WI = reinterpret_cast<float2*>(smem + P.off_WI);
    const int NT = blockDim.x;
    for (int i = threadIdx.x; i < PL.twf_len; i += NT) s_twf[i] = P.twf[i];
    for (int i = threadIdx.x; i < PL.twi_len; i += NT) s_twi[i] = P.twi[i];
    for (int i = threadIdx.x; i < PL.split_len; i += NT) { s_sidx[i] = P.sidx[i]; s_pq1[i] = P.pq1[i]; s_pq2[i] = P.pq2[i]; s_WI[i] = P.WI[i]; }
    uint32_t* s_ordf = reinterpret_cast<uint32_t*>(smem + P.off_ordf);
    uint32_t* s_ordi = reinterpret_cast<uint32_t*>(smem + P.off_ordi);
    for (uint32_t i = threadIdx.x; i < P.ordf_len; i += NT) s_ordf[i] = P.ordf[i];
    for (uint32_t i = threadIdx.x; i < P.ordi_len; i += NT) s_ordi[i] = P.ordi[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    const int group = warp / P.gw;
    DevExec ex;
    ex.nl = P.gw * 32;
    ex.glane = (warp - group * P.gw) * 32 + (int)(threadIdx.x & 31);
    ex.bar_id = 1 + group;
    const int lane = ex.glane, nl = ex.nl;
    unsigned long long* s_items = reinterpret_cast<unsigned long long*>(smem + P.off_items);
    unsigned char* gbase = smem + P.tables + (size_t)group * P.per_group;
    float2* A = reinterpret_cast<float2*>(gbase);
    float2* B = reinterpret_cast<float2*>(gbase + P.off_B);
    float2* carry = reinterpret_cast<float2*>(gbase + P.off_carry);
    const Tables<float2> T{{s_twf}, {s_twi}, s_sidx, s_pq1, s_pq2, s_WI, P.ordf_len ? s_ordf : nullptr, P.ordi_len ? s_ordi : nullptr};
    // N / M: input samples consumed / output samples emitted per block (the blocking; rubato's: the transform lengths)
    const int N = PV::adv_in(PL), M = PV::adv_out(PL), HALF_IN = PV::half_in(PL), CARRY = PV::carry_slots(PL);

    Source src;
    src.pcm = P.pcm; src.fmt = P.fmt; src.ch = P.channels; src.fch = (float)P.channels;
    const uint32_t bps = P.fmt == BB_S16 ? 2u : P.fmt == BB_S24 ? 3u : 4u;
    src.pcm_end = reinterpret_cast<const char*>(P.pcm) + P.total_frames * P.channels * bps;
    src.kind = 2;
    if (P.fmt == BB_S16 && P.channels == 2 && (reinterpret_cast<uintptr_t>(P.pcm) & 3) == 0) src.kind = 0;
    if (P.fmt == BB_S16 && P.channels == 1 && (reinterpret_cast<uintptr_t>(P.pcm) & 1) == 0) src.kind = 1;

    for (;;) {
        if (lane == 0) s_items[group] = atomicAdd(P.counter, 1ull);
        ex.sync();
        const unsigned long long item = s_items[group];
        ex.sync();
        if (item >= P.nitems) break;
        uint64_t lrow; uint32_t b0, b1; bool last_item;
        decode_work_item(P.wi, item, lrow, b0, b1, last_item);
        const uint64_t row = P.row_first + lrow;
        float* __restrict__ orow = P.out + row * P.seg;
        const uint32_t o_lo = min(b0 * (uint32_t)M, P.out_len);
        const uint32_t o_hi = last_item ? P.out_len : min(b1 * (uint32_t)M, P.out_len);
        if (row >= P.nseg) {                         // batch-padding row: zeros (processor.rs:239-260)
            const uint64_t z_hi = last_item ? P.seg : o_hi;
            for (uint64_t j = o_lo + lane; j < z_hi; j += nl) orow[j] = 0.0f;
            continue;
        }
        if (last_item) for (uint64_t j = P.out_len + lane; j < P.seg; j += nl) orow[j] = 0.0f;
        if (b0 >= b1) continue;
        const uint64_t start = (row + 1 == P.nseg) ? P.last_start : row * P.hop;
        const uint64_t take = P.total_frames - start < P.src_seg ? P.total_frames - start : P.src_seg;
        auto valid_of = [&](uint32_t b) -> int {
            const uint64_t q0 = (uint64_t)b * N;
            return q0 < take ? (int)(take - q0 < (uint64_t)N ? take - q0 : (uint64_t)N) : 0;
        };

        for (int j = lane; j < CARRY; j += nl) carry[j] = make_float2(0.f, 0.f);
        const bool vec = ((reinterpret_cast<uintptr_t>(orow) & 7) == 0);      // M is even: b*M keeps 8-byte alignment
        const uint32_t bfirst = b0 > 0 ? b0 - 1 : 0;       // recompute the block before the run for its carry
        int shift = prefetch_block<typename PV::MapA>(src, A, start + (uint64_t)bfirst * N, valid_of(bfirst), HALF_IN, lane, nl);
        for (uint32_t b = bfirst; b < b1; ++b) {
            cp_async_wait_all();
            ex.sync();
            BlockLoader<typename PV::MapA> ld{&src, A, start + (uint64_t)b * N, valid_of(b), shift};
            DevSink sink;
            const int64_t lim = (int64_t)o_hi - (int64_t)b * M;
            sink.p = orow + (size_t)b * M;
            sink.lim = b < b0 ? 0 : (int)(lim < 0 ? 0 : (lim > M ? M : lim));   // recomputed block: carry only
            sink.vec = vec;
            int next_shift = 0;
            PV::template block<float2>(ex, PL, T, A, B, carry, ld, sink, [&] {
                if (b + 1 < b1) next_shift = prefetch_block<typename PV::MapA>(src, A, start + (uint64_t)(b + 1) * N, valid_of(b + 1), HALF_IN, lane, nl);
            });
            shift = next_shift;
        }
    }
}



// ------------------------------------------------------------------------------------------------
// Two-stream variant: a group runs TWO rows (windows r, r+1; same block range) in lockstep.  Every
// shared-memory element is a float4 (re0, re1, im0, im1) and every butterfly instruction is a packed
// f32x2 op, so index math, loads/stores and FP issue slots are shared by the two rows.
struct DualStream { uint64_t base; int valid; int shift; bool active; };

template <int KIND>
__device__ __forceinline__ void conv_pair(int a, int b, int shift, float& re, float& im) {
    if constexpr (KIND == 0) {           // two packed L|R words: exact integer channel sum (one dp2a: L*1 + R*1), then * 2^-15 / 2
        re = __fmul_rn(__int2float_rn(__dp2a_lo(a, 0x0101, 0)), 1.0f / 65536.0f);
        im = __fmul_rn(__int2float_rn(__dp2a_lo(b, 0x0101, 0)), 1.0f / 65536.0f);
    } else {                             // mono: the aligned word(s) covering the frame pair
        const int w = __funnelshift_r(a, b, shift);
        re = __fmul_rn(__int2float_rn((int)(short)(w & 0xffff)), 1.0f / 32768.0f);
        im = __fmul_rn(__int2float_rn(w >> 16), 1.0f / 32768.0f);
    }
}

// Loader of one block of the two streams.  KIND: 0 = S16 stereo staged in A, 1 = S16 mono staged in A,
// 2 = direct global loads (other formats).  FULL: both streams hold `n` valid frames (interior block),
// so only the odd tail sample needs zeroing.
template <int KIND, bool FULL, class AM>
struct DualLoaderT {
    const Source* src; const float4* A; DualStream s0, s1; int n;
    __device__ __forceinline__ cx2 operator()(int e) const {
        const int i0 = 2 * e;
        float r0 = 0.f, m0 = 0.f, r1 = 0.f, m1 = 0.f;
        if constexpr (KIND != 2) {
            const int4 v = *reinterpret_cast<const int4*>(A + AM::at(e));
            conv_pair<KIND>(v.x, v.y, s0.shift, r0, m0);
            conv_pair<KIND>(v.z, v.w, s1.shift, r1, m1);
            if constexpr (FULL) {
                if (i0 + 1 >= n) { m0 = 0.f; m1 = 0.f; }
            } else {
                if (i0 >= s0.valid) r0 = 0.f;
                if (i0 + 1 >= s0.valid) m0 = 0.f;
                if (i0 >= s1.valid) r1 = 0.f;
                if (i0 + 1 >= s1.valid) m1 = 0.f;
            }
        } else {
            if (i0 < s0.valid) r0 = src->direct(s0.base + i0);
            if (i0 + 1 < s0.valid) m0 = src->direct(s0.base + i0 + 1);
            if (i0 < s1.valid) r1 = src->direct(s1.base + i0);
            if (i0 + 1 < s1.valid) m1 = src->direct(s1.base + i0 + 1);
        }
        cx2 c; c.re = make_float2(r0, r1); c.im = make_float2(m0, m1);
        return c;
    }
};
// picks the interior-block loader once per block (k2_warp.cuh: loader_pick); the test is uniform over the group
template <int KIND, class AM>
struct DualLoaderSet {
    const Source* src; const float4* A; DualStream s0, s1; int n;
    template <class F> __device__ __forceinline__ void pick(F&& f) const {
        if constexpr (KIND == 2) f(DualLoaderT<2, false, AM>{src, A, s0, s1, n});
        else {
            if (s0.valid == n && s1.valid == n) f(DualLoaderT<KIND, true, AM>{src, A, s0, s1, n});
            else f(DualLoaderT<KIND, false, AM>{src, A, s0, s1, n});
        }
    }
};

__device__ __forceinline__ void cp_async4_full(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8_full(void* smem_dst, const void* gsrc) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gsrc) : "memory");
}

// stage one stream's raw PCM of a block into its 8-byte half of the float4 slots.  Interior blocks (all
// `n_full` frames valid and the rounded-up copy inside the buffer) take a loop without bounds arithmetic.
template <int KIND, class AM>
__device__ __forceinline__ int prefetch_half(const Source& s, float4* A, int half, uint64_t base, int valid, int half_in, int n_full,
                                             int lane, int nl) {
    int shift = 0;
    char* dst0 = reinterpret_cast<char*>(A) + half * 8;
    if constexpr (KIND == 0) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + base * 4;
        const bool al8 = (reinterpret_cast<uintptr_t>(g0) & 7) == 0;
        if (valid == n_full && g0 + (size_t)half_in * 8 <= s.pcm_end) {
            const char* g = g0 + (size_t)lane * 8;
            if (al8) for (int n = lane; n < half_in; n += nl, g += (size_t)nl * 8) cp_async8_full(dst0 + (size_t)AM::at(n) * 16, g);
            else for (int n = lane; n < half_in; n += nl, g += (size_t)nl * 8) { char* d = dst0 + (size_t)AM::at(n) * 16; cp_async4_full(d, g); cp_async4_full(d + 4, g + 4); }
            return 0;
        }
        for (int n = lane; n < half_in; n += nl) {
            if (2 * n >= valid) break;
            const char* g = g0 + (size_t)n * 8;
            char* d = dst0 + (size_t)AM::at(n) * 16;
            const long long room = s.pcm_end - g;
            stage_tail(d, g, room, 8);
        }
    } else if constexpr (KIND == 1) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + base * 2;
        const bool odd = (reinterpret_cast<uintptr_t>(g0) & 3) != 0;
        shift = odd ? 16 : 0;
        g0 -= odd ? 2 : 0;
        if (valid == n_full && g0 + (size_t)half_in * 4 + 4 <= s.pcm_end) {
            const char* g = g0 + (size_t)lane * 4;
            if (!odd) for (int n = lane; n < half_in; n += nl, g += (size_t)nl * 4) cp_async4_full(dst0 + (size_t)AM::at(n) * 16, g);
            else for (int n = lane; n < half_in; n += nl, g += (size_t)nl * 4) { char* d = dst0 + (size_t)AM::at(n) * 16; cp_async4_full(d, g); cp_async4_full(d + 4, g + 4); }
            return shift;
        }
        for (int n = lane; n < half_in; n += nl) {
            if (2 * n >= valid) break;
            const char* g = g0 + (size_t)n * 4;
            char* d = dst0 + (size_t)AM::at(n) * 16;
            const long long room = s.pcm_end - g;
            stage_tail(d, g, room, odd ? 8 : 4);
        }
    }
    return shift;
}

// One lane's share of a staging pass: slots n0, n0 + dn, ... < half_in; the piece of slot n comes from g + k * dn *
// SRC_STRIDE (k-th step) and goes to d + map(n) * 16.  When the thread-group size is the one the kernel was built for
// (NLH lanes, HALF slots: compile-time), the loop unrolls into cp.async instructions with immediate offsets.
template <int BYTES, int SRC_STRIDE, int LANES_PER_SLOT, class AM, int NLH, int HALF>
__device__ __forceinline__ void stage_loop(char* d, const char* g, int n0, int half_in, int nl) {
    if constexpr (NLH > 0 && HALF > 0) {
        if (nl == NLH && half_in == HALF) {
            constexpr int DN = NLH / LANES_PER_SLOT, IT = (HALF + DN - 1) / DN;
#pragma unroll
            for (int k = 0; k < IT; ++k) {
                const int n = n0 + k * DN;
                if (k + 1 < IT || n < HALF) {
                    if (BYTES == 8) cp_async8_full(d + (size_t)AM::at(n) * 16, g + (size_t)k * DN * SRC_STRIDE);
                    else cp_async4_full(d + (size_t)AM::at(n) * 16, g + (size_t)k * DN * SRC_STRIDE);
                }
            }
            return;
        }
    }
    const int dn = nl / LANES_PER_SLOT;
    for (int n = n0; n < half_in; n += dn, g += (size_t)dn * SRC_STRIDE) {
        if (BYTES == 8) cp_async8_full(d + (size_t)AM::at(n) * 16, g);
        else cp_async4_full(d + (size_t)AM::at(n) * 16, g);
    }
}

// Stage the raw PCM of one block of BOTH streams.  Interior blocks use a lane mapping in which consecutive lanes
// fill adjacent 4- or 8-byte pieces of the same float4 slot, so a warp's cp.async writes cover contiguous shared
// memory (no bank conflicts); anything else goes stream by stream through prefetch_half.
template <int KIND, class AM, int NLH, int HALF>
__device__ __forceinline__ void prefetch_pair(const Source& s, float4* A, DualStream& a, DualStream& b, int half_in, int n_full,
                                              int lane, int nl) {
    if constexpr (KIND == 0) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + a.base * 4;
        const char* g1 = reinterpret_cast<const char*>(s.pcm) + b.base * 4;
        if (a.valid == n_full && b.valid == n_full && g0 + (size_t)half_in * 8 <= s.pcm_end && g1 + (size_t)half_in * 8 <= s.pcm_end) {
            a.shift = 0; b.shift = 0;
            if (((reinterpret_cast<uintptr_t>(g0) | reinterpret_cast<uintptr_t>(g1)) & 7) == 0) {
                const int st = lane & 1, n0 = lane >> 1;
                stage_loop<8, 8, 2, AM, NLH, HALF>(reinterpret_cast<char*>(A) + st * 8, (st ? g1 : g0) + (size_t)n0 * 8, n0, half_in, nl);
            } else {
                const int qd = lane & 3, n0 = lane >> 2;
                stage_loop<4, 8, 4, AM, NLH, HALF>(reinterpret_cast<char*>(A) + qd * 4, ((qd >> 1) ? g1 : g0) + (size_t)n0 * 8 + (qd & 1) * 4,
                                                   n0, half_in, nl);
            }
            return;
        }
    } else if constexpr (KIND == 1) {
        const char* g0 = reinterpret_cast<const char*>(s.pcm) + a.base * 2;
        const char* g1 = reinterpret_cast<const char*>(s.pcm) + b.base * 2;
        const bool odd0 = (reinterpret_cast<uintptr_t>(g0) & 3) != 0, odd1 = (reinterpret_cast<uintptr_t>(g1) & 3) != 0;
        g0 -= odd0 ? 2 : 0; g1 -= odd1 ? 2 : 0;
        if (a.valid == n_full && b.valid == n_full && g0 + (size_t)half_in * 4 + 4 <= s.pcm_end && g1 + (size_t)half_in * 4 + 4 <= s.pcm_end) {
            a.shift = odd0 ? 16 : 0; b.shift = odd1 ? 16 : 0;
            if (!odd0 && !odd1) {                       // one aligned word per frame pair and stream
                const int st = lane & 1, n0 = lane >> 1;
                stage_loop<4, 4, 2, AM, NLH, HALF>(reinterpret_cast<char*>(A) + st * 8, (st ? g1 : g0) + (size_t)n0 * 4, n0, half_in, nl);
            } else {                                    // two words (the funnel shift picks the pair)
                const int qd = lane & 3, n0 = lane >> 2;
                stage_loop<4, 4, 4, AM, NLH, HALF>(reinterpret_cast<char*>(A) + qd * 4, ((qd >> 1) ? g1 : g0) + (size_t)n0 * 4 + (qd & 1) * 4,
                                                   n0, half_in, nl);
            }
            return;
        }
    }
    a.shift = prefetch_half<KIND, AM>(s, A, 0, a.base, a.valid, half_in, n_full, lane, nl);
    b.shift = prefetch_half<KIND, AM>(s, A, 1, b.base, b.valid, half_in, n_full, lane, nl);
}

struct DualSink {
    float* p0; float* p1; int lim0, lim1; bool vec0, vec1;
    bool full;          // both streams take every sample of the block and both rows are 8-byte aligned: plain vector stores
    __device__ __forceinline__ void put(float* p, int lim, bool vec, int o, float a, float b) const {
        if (vec && o + 1 < lim) *reinterpret_cast<float2*>(p + o) = make_float2(a, b);
        else { if (o < lim) p[o] = a; if (o + 1 < lim) p[o + 1] = b; }
    }
    __device__ __forceinline__ void operator()(int n, const cx2& y) const {
        const int o = 2 * n;
        if (full) {
            *reinterpret_cast<float2*>(p0 + o) = make_float2(y.re.x, y.im.x);
            *reinterpret_cast<float2*>(p1 + o) = make_float2(y.re.y, y.im.y);
            return;
        }
        put(p0, lim0, vec0, o, y.re.x, y.im.x);
        put(p1, lim1, vec1, o, y.re.y, y.im.y);
    }
};

template <class PV, int KIND, int NLH>
__device__ __forceinline__ void resample_body_dual(const WarpParams& P) {
    extern __shared__ __align__(16) unsigned char smem[];
    const RtPlan& PL = P.plan;
    constexpr bool CTW = PV::kCompactTw;
    using TwE = typename Tw<cx2, CTW>::E;                 // float4 (x, x, y, y), or float2 when the plan keeps its twiddles compact
    TwE* s_twf = reinterpret_cast<TwE*>(smem);
    TwE* s_twi = reinterpret_cast<TwE*>(smem + P.off_twi);
    uint4* s_sidx = reinterpret_cast<uint4*>(smem + P.off_sidx);
    float4* s_pq1 = reinterpret_cast<float4*>(smem + P.off_pq1);
    float4* s_pq2 = reinterpret_cast<float4*>(smem + P.off_pq2);
    float2* s_WI = reinterpret_cast<float2*>(smem + P.off_WI);
    const int NT = blockDim.x;
    auto bc = [](float2 w) { if constexpr (CTW) return w; else return make_float4(w.x, w.x, w.y, w.y); };
    for (int i = threadIdx.x; i < PL.twf_len; i += NT) s_twf[i] = bc(P.twf[i]);
    for (int i = threadIdx.x; i < PL.twi_len; i += NT) s_twi[i] = bc(P.twi[i]);
    for (int i = threadIdx.x; i < PL.split_len; i += NT) { s_sidx[i] = P.sidx[i]; s_pq1[i] = P.pq1[i]; s_pq2[i] = P.pq2[i]; s_WI[i] = P.WI[i]; }
    uint32_t* s_ordf = reinterpret_cast<uint32_t*>(smem + P.off_ordf);
    uint32_t* s_ordi = reinterpret_cast<uint32_t*>(smem + P.off_ordi);
    for (uint32_t i = threadIdx.x; i < P.ordf_len; i += NT) s_ordf[i] = P.ordf[i];
    for (uint32_t i = threadIdx.x; i < P.ordi_len; i += NT) s_ordi[i] = P.ordi[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5;
    // the group size is a compile-time constant of this kernel (the launcher checks gw * 32 == NLH and takes the
    // single-stream kernel otherwise): constant trip counts in every stage loop — C2 1.94 -> 1.86 ms, C3 1.09 -> 1.05
    static_assert(NLH >= 32 && NLH % 32 == 0, "two-stream kernels are built for one group size");
    constexpr int GW = NLH / 32;
    const int group = warp / GW;
    DevExecC<NLH> ex;
    ex.glane = (warp - group * GW) * 32 + (int)(threadIdx.x & 31);
    ex.bar_id = 1 + group;
    const int lane = ex.glane; constexpr int nl = NLH;
    unsigned long long* s_items = reinterpret_cast<unsigned long long*>(smem + P.off_items);
    unsigned char* gbase = smem + P.tables + (size_t)group * P.per_group;
    float4* A = reinterpret_cast<float4*>(gbase);
    float4* B = reinterpret_cast<float4*>(gbase + P.off_B);
    float4* carry = reinterpret_cast<float4*>(gbase + P.off_carry);
    const Tables<cx2, CTW> T{{s_twf}, {s_twi}, s_sidx, s_pq1, s_pq2, s_WI, P.ordf_len ? s_ordf : nullptr, P.ordi_len ? s_ordi : nullptr};
    // N / M: input samples consumed / output samples emitted per block (the blocking; rubato's: the transform lengths)
    const int N = PV::adv_in(PL), M = PV::adv_out(PL), HALF_IN = PV::half_in(PL), CARRY = PV::carry_slots(PL);

    Source src;
    src.pcm = P.pcm; src.fmt = P.fmt; src.ch = P.channels; src.fch = (float)P.channels;
    const uint32_t bps = P.fmt == BB_S16 ? 2u : P.fmt == BB_S24 ? 3u : 4u;
    src.pcm_end = reinterpret_cast<const char*>(P.pcm) + P.total_frames * P.channels * bps;
    src.kind = KIND;                                   // chosen by the launcher (source_kind)
    const uint64_t row_end = P.row_first + P.rows_total;

    for (;;) {
        if (lane == 0) s_items[group] = atomicAdd(P.counter, 1ull);
        ex.sync();
        const unsigned long long item = s_items[group];
        ex.sync();
        if (item >= P.nitems) break;
        uint64_t lpair; uint32_t b0, b1; bool last_item;
        decode_work_item(P.wi, item, lpair, b0, b1, last_item);
        const uint32_t o_lo = min(b0 * (uint32_t)M, P.out_len);
        const uint32_t o_hi = last_item ? P.out_len : min(b1 * (uint32_t)M, P.out_len);
        uint64_t start[2], take[2]; float* orow[2]; bool active[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint64_t row = P.row_first + 2 * lpair + s;
            orow[s] = P.out + row * P.seg;
            active[s] = row < row_end && row < P.nseg;
            start[s] = 0; take[s] = 0;
            if (row < row_end && row >= P.nseg) {            // batch-padding row: zeros (processor.rs:239-260)
                const uint64_t z_hi = last_item ? P.seg : o_hi;
                for (uint64_t j = o_lo + lane; j < z_hi; j += nl) orow[s][j] = 0.0f;
            }
            if (active[s]) {
                if (last_item) for (uint64_t j = P.out_len + lane; j < P.seg; j += nl) orow[s][j] = 0.0f;
                start[s] = (row + 1 == P.nseg) ? P.last_start : row * P.hop;
                take[s] = P.total_frames - start[s] < P.src_seg ? P.total_frames - start[s] : P.src_seg;
            }
        }
        if ((!active[0] && !active[1]) || b0 >= b1) continue;
        auto stream_of = [&](int s, uint32_t b) -> DualStream {
            DualStream d; d.active = active[s]; d.shift = 0;
            const uint64_t q0 = (uint64_t)b * N;
            d.valid = (active[s] && q0 < take[s]) ? (int)(take[s] - q0 < (uint64_t)N ? take[s] - q0 : (uint64_t)N) : 0;
            d.base = start[s] + q0;
            return d;
        };
        for (int j = lane; j < CARRY; j += nl) carry[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t bfirst = b0 > 0 ? b0 - 1 : 0;       // recompute the block before the run for its carry
        DualStream c0 = stream_of(0, bfirst), c1 = stream_of(1, bfirst);
        prefetch_pair<KIND, typename PV::MapA, NLH, PV::half_in(RtPlan{})>(src, A, c0, c1, HALF_IN, N, lane, nl);
        for (uint32_t b = bfirst; b < b1; ++b) {
            cp_async_wait_all();
            ex.sync();
            DualLoaderSet<KIND, typename PV::MapA> ld{&src, A, c0, c1, N};
            DualSink sink;
            const int64_t lim = (int64_t)o_hi - (int64_t)b * M;
            const int l = b < b0 ? 0 : (int)(lim < 0 ? 0 : (lim > M ? M : lim));   // recomputed block: carry only
            sink.p0 = orow[0] + (size_t)b * M; sink.p1 = orow[1] + (size_t)b * M;
            sink.lim0 = active[0] ? l : 0; sink.lim1 = active[1] ? l : 0;
            sink.vec0 = (reinterpret_cast<uintptr_t>(orow[0]) & 7) == 0; sink.vec1 = (reinterpret_cast<uintptr_t>(orow[1]) & 7) == 0;
            sink.full = sink.vec0 && sink.vec1 && sink.lim0 == M && sink.lim1 == M;
            DualStream n0 = c0, n1 = c1;
            PV::template block<cx2>(ex, PL, T, A, B, carry, ld, sink, [&] {
                if (b + 1 < b1) {
                    n0 = stream_of(0, b + 1); n1 = stream_of(1, b + 1);
                    prefetch_pair<KIND, typename PV::MapA, NLH, PV::half_in(RtPlan{})>(src, A, n0, n1, HALF_IN, N, lane, nl);
                }
            });
            c0 = n0; c1 = n1;
        }
    }
}

template <class AM>
__global__ void __launch_bounds__(kMaxThreads, 1)
resample_warp_kernel(const __grid_constant__ WarpParams P) { resample_body<RtView<AM>>(P); }

template <class PL>
__global__ void __launch_bounds__(kMaxThreads, 1)
resample_plan_kernel(const __grid_constant__ WarpParams P) { resample_body<CtView<PL>>(P); }

// two-stream kernel; THREADS is the plan's CTA size (BB_K2_CT_PLANS): register budget and warps per group
template <class PL, int THREADS, int KIND>
__global__ void __launch_bounds__(THREADS, 1)
resample_plan2_kernel(const __grid_constant__ WarpParams P) {
    // expected thread-group size (the staging loops unroll for it; any other size takes their generic form)
    resample_body_dual<CtView<PL>, KIND, PL::dual_group_lanes(THREADS, (int)kSmemMax, kMaxGroups)>(P);
}

// how the kernels read the PCM: 0 = S16 stereo staged with cp.async, 1 = S16 mono staged, 2 = direct loads
int source_kind(const void* pcm, int fmt, uint32_t channels) {
    if (fmt == BB_S16 && channels == 2 && (reinterpret_cast<uintptr_t>(pcm) & 3) == 0) return 0;
    if (fmt == BB_S16 && channels == 1 && (reinterpret_cast<uintptr_t>(pcm) & 1) == 0) return 1;
    return 2;
}

template <class PL, int THREADS>
cudaError_t launch_plan2(int kind, unsigned ctas, unsigned threads, size_t smem, cudaStream_t st, const WarpParams& P) {
    cudaError_t e = cudaSuccess;
#define BB_L2(K)                                                                                                          \
    e = cudaFuncSetAttribute(resample_plan2_kernel<PL, THREADS, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return e;                                                                                        \
    resample_plan2_kernel<PL, THREADS, K><<<ctas, threads, smem, st>>>(P);
    if (kind == 0) { BB_L2(0) } else if (kind == 1) { BB_L2(1) } else { BB_L2(2) }
#undef BB_L2
    return e;
}

// lanes per thread group the two-stream kernel of a compile-time plan is built for
int ct_plan_dual_lanes(int ct_index) {
    int i = 0;
#define BB_CT(NAME, NI, NO, TH, ...) if (i == ct_index) return __VA_ARGS__::dual_group_lanes(TH, (int)kSmemMax, kMaxGroups); ++i;
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    return 0;
}

int ct_plan_dual_threads(int ct_index) {
    int i = 0;
#define BB_CT(NAME, NI, NO, TH, ...) if (i == ct_index) return TH; ++i;
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    return 384;
}

bool ct_plan_compact_tw(int ct_index) {
    int i = 0;
#define BB_CT(NAME, NI, NO, TH, ...) if (i == ct_index) return __VA_ARGS__::COMPACT_TW; ++i;
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    return false;
}

// index of the compile-time plan for (n_in, n_out), -1 when only the runtime plan applies
// BIRDA_K2_BLOCKING=rubato skips the plans that use their own blocking (A/B runs, tests)
int ct_plan_index(uint32_t n_in, uint32_t n_out) {
    if (const char* g = std::getenv("BIRDA_K2_RUNTIME_PLAN")) if (g[0] == '1') return -1;
    bool own_blocking = true;
    if (const char* g = std::getenv("BIRDA_K2_BLOCKING")) if (g[0] == 'r') own_blocking = false;
    int i = 0;
#define BB_CT(NAME, NI, NO, TH, ...) if (n_in == NI && n_out == NO && (own_blocking || __VA_ARGS__::ADV_IN == NI)) return i; ++i;
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    return -1;
}

template <class PL> void ct_radices(std::vector<int>* f, std::vector<int>* v, int* adv_in, int* adv_out) {
    f->clear(); v->clear();
    for (int i = 0; i < PL::Fwd::count; ++i) f->push_back(PL::Fwd::at(i));
    for (int i = 0; i < PL::Inv::count; ++i) v->push_back(PL::Inv::at(i));
    *adv_in = PL::ADV_IN; *adv_out = PL::ADV_OUT;
}

// Spectrum of the reference's taps for a transform of 2N real points (N != n_in: a plan with its own blocking):
// bins [0, nkeep) of the taps zero-padded to 2N, with the forward transform's normalisation 1 / (2N) in place of
// the reference's 1 / (2 n_in) (rules.cpp: make_resampler_spec applies the latter to spec.taps).  f64, exact index
// reduction, rounded once.
void blocked_filter_spectrum(const ResamplerSpec& spec, int N, int nkeep, std::vector<float>* re, std::vector<float>* im) {
    const uint32_t L = 2u * (uint32_t)N, n = spec.n_in;
    const double scale = (double)(2u * n) / (double)L;
    std::vector<double> cs(L), sn(L);
    for (uint32_t j = 0; j < L; ++j) {
        const double a = -2.0 * 3.14159265358979323846 * (double)j / (double)L;
        cs[j] = std::cos(a); sn[j] = std::sin(a);
    }
    re->assign(nkeep, 0.f); im->assign(nkeep, 0.f);
    for (int k = 0; k < nkeep; ++k) {
        double r = 0, i = 0;
        uint32_t idx = 0;
        for (uint32_t x = 0; x < n; ++x) {
            r += (double)spec.taps[x] * cs[idx];
            i += (double)spec.taps[x] * sn[idx];
            idx += (uint32_t)k; if (idx >= L) idx -= L;
        }
        (*re)[k] = (float)(r * scale); (*im)[k] = (float)(i * scale);
    }
}

}  // namespace

bool warp_plan_available(const ResamplerSpec& spec) {
    if (const char* g = std::getenv("BIRDA_K2_GENERIC")) if (g[0] == '1') return false;
    RtPlan P; std::vector<int> f, i;
    if (!build_plan((int)spec.n_in, (int)spec.n_out, (int)spec.n_keep, &P, &f, &i)) return false;
    rt_plan_pads(&P);
    const size_t per_group = (size_t)phys_len(P.N, P.pad_a) * 8 + (size_t)phys_len(P.M, P.pad_b) * 8 + (size_t)spec.n_out * 4 + 48;
    const size_t tables = ((size_t)P.twf_len + P.twi_len) * 8 + (size_t)P.split_len * 56 + ((size_t)P.N + P.M) * 4 + 512;
    return per_group + tables < kSmemMax;        // at least one block in flight next to the tables
}

cudaError_t warp_tables_init(const ResamplerSpec& spec, ResamplerDev* rs) {
    RtPlan P; std::vector<int> fwd, inv;
    rs->ct_index = ct_plan_index(spec.n_in, spec.n_out);
    std::vector<float> own_re, own_im;
    const float* filt_re = spec.filt_re.data(); const float* filt_im = spec.filt_im.data();
    if (rs->ct_index >= 0) {
        int i = 0, adv_in = 0, adv_out = 0;
#define BB_CT(NAME, NI, NO, TH, ...) if (i == rs->ct_index) ct_radices<__VA_ARGS__>(&fwd, &inv, &adv_in, &adv_out); ++i;
        BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
        int N = 1, M = 1;
        for (int r : fwd) N *= r;
        for (int r : inv) M *= r;
        if (N == (int)spec.n_in && M == (int)spec.n_out) {
            if (!build_plan_from_radices(N, M, (int)spec.n_keep, &P, &fwd, &inv)) return cudaErrorInvalidConfiguration;
        } else {
            // a plan with its own blocking: longer transforms over the same taps.  Only where the spectrum is extended
            // (up-sampling): truncating it ties the result to rubato's block length.
            const uint64_t g = std::gcd((uint64_t)spec.n_in, (uint64_t)spec.n_out);
            const uint64_t a = spec.n_in / g, b = spec.n_out / g;
            if (spec.n_in >= spec.n_out || (uint64_t)N * b != (uint64_t)M * a || (uint64_t)adv_in * b != (uint64_t)adv_out * a ||
                (uint64_t)adv_in + spec.n_in - 1 > 2ull * N) return cudaErrorInvalidConfiguration;
            if (!build_plan_from_radices(N, M, N + 1, &P, &fwd, &inv, adv_in, adv_out)) return cudaErrorInvalidConfiguration;
            blocked_filter_spectrum(spec, N, P.nkeep, &own_re, &own_im);
            filt_re = own_re.data(); filt_im = own_im.data();
        }
    } else if (!build_plan((int)spec.n_in, (int)spec.n_out, (int)spec.n_keep, &P, &fwd, &inv)) return cudaErrorInvalidConfiguration;
    if (rs->ct_index >= 0) ct_plan_pads(&P); else rt_plan_pads(&P);
    // order tables: groups of 8 lanes for the two-stream kernels of compile-time plans (their offsets are compile-time,
    // the single-stream kernels of those plans ignore the tables), groups of 16 for the runtime-plan kernel
    std::vector<uint32_t> ordf, ordi;
    build_stage_orders(&P, rs->ct_index >= 0 ? 8 : 16, rs->ct_index < 0, &ordf, &ordi);
    std::vector<uint16_t> pf(P.N), pi_(P.M);
    build_pos_tables(fwd, inv, P.N, P.M, pf.data(), pi_.data());
    std::vector<float2> Pt(P.nkeep), Qt(P.nkeep), WI(P.M / 2 + 1), twf(P.twf_len), twi(P.twi_len);
    build_split_tables(P.N, P.M, P.nkeep, filt_re, filt_im, Pt.data(), Qt.data(), WI.data());
    build_twiddles(P, twf.data(), twi.data());
    // bank groups: 8 lanes of 16-byte elements in two-stream mode (compile-time plans), 16 lanes of 8-byte ones otherwise
    SplitLayout SL;
    build_split_layout(P.N, P.M, P.nkeep, pf.data(), pi_.data(), Pt.data(), Qt.data(), WI.data(), rs->ct_index >= 0 ? 8 : 16, &SL,
                       P.pad_a, P.pad_b);
    auto up = [](const void* h, size_t bytes, void** d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e;
    if ((e = up(twf.data(), twf.size() * 8, (void**)&rs->f_twf)) != cudaSuccess) return e;
    if ((e = up(twi.data(), twi.size() * 8, (void**)&rs->f_twi)) != cudaSuccess) return e;
    if ((e = up(SL.sidx.data(), SL.sidx.size() * 16, (void**)&rs->f_sidx)) != cudaSuccess) return e;
    if ((e = up(SL.pq1.data(), SL.pq1.size() * 16, (void**)&rs->f_pq1)) != cudaSuccess) return e;
    if ((e = up(SL.pq2.data(), SL.pq2.size() * 16, (void**)&rs->f_pq2)) != cudaSuccess) return e;
    if ((e = up(SL.wi.data(), SL.wi.size() * 8, (void**)&rs->f_WI)) != cudaSuccess) return e;
    rs->ordf_len = (uint32_t)ordf.size(); rs->ordi_len = (uint32_t)ordi.size();
    if (!ordf.empty() && (e = up(ordf.data(), ordf.size() * 4, (void**)&rs->f_ordf)) != cudaSuccess) return e;
    if (!ordi.empty() && (e = up(ordi.data(), ordi.size() * 4, (void**)&rs->f_ordi)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&rs->f_counter, sizeof(unsigned long long))) != cudaSuccess) return e;
    static_assert(sizeof(RtPlan) <= sizeof(rs->plan_blob), "plan blob too small");
    memcpy(rs->plan_blob, &P, sizeof(P));
    rs->fast = true;
    return cudaSuccess;
}

std::string warp_plan_describe(const ResamplerDev& rs) {
    if (!rs.fast) return "resample_kernel (CTA-cooperative fallback) blocks " + std::to_string(rs.n_in) + "/" + std::to_string(rs.n_out);
    RtPlan P; memcpy(&P, rs.plan_blob, sizeof(P));
    std::string s = rs.ct_index >= 0 ? "resample_plan2_kernel (two-stream compile-time plan" : "resample_warp_kernel (runtime plan";
    s += ", transforms " + std::to_string(2 * P.N) + " -> " + std::to_string(2 * P.M) + " real points, " + std::to_string(P.adv_in) + " in / " +
         std::to_string(P.adv_out) + " out per block";
    s += (P.adv_in == (int)rs.n_in) ? ", the reference's blocking" : ", own blocking over the reference's " + std::to_string(rs.n_in) + " taps";
    if (rs.ct_index >= 0) s += ", " + std::to_string(ct_plan_dual_threads(rs.ct_index)) + " threads";
    return s + ")";
}

void warp_tables_free(ResamplerDev* rs) {
    void* ptrs[] = {rs->f_twf, rs->f_twi, rs->f_sidx, rs->f_pq1, rs->f_pq2, rs->f_WI, rs->f_counter, rs->f_ordf, rs->f_ordi};
    for (void* p : ptrs) if (p) cudaFree(p);
    rs->f_twf = rs->f_twi = rs->f_WI = nullptr; rs->f_sidx = nullptr; rs->f_pq1 = rs->f_pq2 = nullptr; rs->f_counter = nullptr; rs->f_ordf = rs->f_ordi = nullptr; rs->ordf_len = rs->ordi_len = 0;
    rs->fast = false;
}

static cudaError_t launch_warp_impl(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                                    uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                                    uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                                    uint64_t resampled_len, float* d_out, int* launches, bool allow_dual);

cudaError_t launch_resample_warp(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                                 uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                                 uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                                 uint64_t resampled_len, float* d_out, int* launches) {
    return launch_warp_impl(st, sm_count, rs, d_pcm, fmt, channels, total_frames, src_seg, hop, nseg, last_start, row_first,
                            rows_total, seg, resampled_len, d_out, launches, true);
}

static cudaError_t launch_warp_impl(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                                    uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                                    uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                                    uint64_t resampled_len, float* d_out, int* launches, bool allow_dual) {
    if (launches) *launches = 0;
    if (rows_total == 0) return cudaSuccess;
    WarpParams P{};
    memcpy(&P.plan, rs.plan_blob, sizeof(RtPlan));
    const RtPlan& PL = P.plan;
    P.pcm = d_pcm; P.fmt = fmt; P.channels = channels; P.total_frames = total_frames;
    P.src_seg = src_seg; P.hop = hop; P.nseg = nseg; P.last_start = last_start; P.rows_total = rows_total; P.row_first = row_first; P.seg = seg;
    P.out_len = (uint32_t)(resampled_len < seg ? resampled_len : seg);
    P.out = d_out;
    P.twf = rs.f_twf; P.twi = rs.f_twi; P.WI = rs.f_WI; P.sidx = rs.f_sidx; P.pq1 = rs.f_pq1; P.pq2 = rs.f_pq2;
    P.ordf = rs.f_ordf; P.ordi = rs.f_ordi; P.ordf_len = rs.ordf_len; P.ordi_len = rs.ordi_len;
    P.counter = rs.f_counter;
    P.nblk = (P.out_len + (uint32_t)PL.adv_out - 1) / (uint32_t)PL.adv_out;
    if (P.nblk == 0) P.nblk = 1;
    auto a16 = [](size_t x) { return (uint32_t)((x + 15) & ~(size_t)15); };
    // two-stream mode: compile-time plans only, at least two rows
    bool dual = allow_dual && rs.ct_index >= 0 && rows_total >= 2;
    if (const char* g = std::getenv("BIRDA_K2_DUAL")) if (g[0] == '0') dual = false;
    const size_t esz = dual ? 16 : 8;                   // bytes per complex element in shared memory
    const int max_threads = dual ? ct_plan_dual_threads(rs.ct_index) : kMaxThreads;
    P.dual = dual ? 1 : 0;
    const size_t tw_esz = (dual && ct_plan_compact_tw(rs.ct_index)) ? 8 : esz;
    P.off_twi = a16((size_t)PL.twf_len * tw_esz);
    P.off_sidx = P.off_twi + a16((size_t)PL.twi_len * tw_esz);
    P.off_pq1 = P.off_sidx + a16((size_t)PL.split_len * 16);       // split tables have the same layout in both modes
    P.off_pq2 = P.off_pq1 + a16((size_t)PL.split_len * 16);
    P.off_WI = P.off_pq2 + a16((size_t)PL.split_len * 16);
    P.off_ordf = P.off_WI + a16((size_t)PL.split_len * 8);
    P.off_ordi = P.off_ordf + a16((size_t)P.ordf_len * 4);
    P.off_items = P.off_ordi + a16((size_t)P.ordi_len * 4);
    P.tables = P.off_items + a16((size_t)kMaxGroups * 8);
    P.off_B = a16((size_t)phys_len(PL.N, PL.pad_a) * esz);
    P.off_carry = P.off_B + a16((size_t)phys_len(PL.M, PL.pad_b) * esz);
    P.per_group = P.off_carry + a16((size_t)PL.carry_slots * esz);
    if (P.tables + P.per_group > kSmemMax) {
        if (!dual) return cudaErrorInvalidConfiguration;
        // does not fit with two streams: one stream per group
        return launch_warp_impl(st, sm_count, rs, d_pcm, fmt, channels, total_frames, src_seg, hop, nseg, last_start,
                                row_first, rows_total, seg, resampled_len, d_out, launches, false);
    }
    int groups = (int)((kSmemMax - P.tables) / P.per_group);
    if (groups > kMaxGroups) groups = kMaxGroups;
    if (groups > max_threads / 32) groups = max_threads / 32;
    int gw = (max_threads / 32) / groups;               // warps cooperating on one block (pair)
    if (gw < 1) gw = 1;
    if (gw > 16) gw = 16;
    if (const char* g = std::getenv("BIRDA_K2_GROUP_WARPS")) { int v = atoi(g); if (v >= 1 && v <= 16 && v * groups * 32 <= max_threads) gw = v; }
    if (const char* g = std::getenv("BIRDA_K2_DEBUG")) if (g[0] == '1')
        fprintf(stderr, "[k2] plan N=%d M=%d adv %d/%d dual=%d tables=%u per_group=%u groups=%d gw=%d smem=%zu nblk=%u\n", PL.N, PL.M, PL.adv_in, PL.adv_out,
                (int)dual, P.tables, P.per_group, groups, gw, (size_t)P.tables + (size_t)groups * P.per_group, P.nblk);
    P.groups = groups; P.gw = gw;
    if (dual && gw * 32 != ct_plan_dual_lanes(rs.ct_index))       // not the group size the two-stream kernel was built for
        return launch_warp_impl(st, sm_count, rs, d_pcm, fmt, channels, total_frames, src_seg, hop, nseg, last_start,
                                row_first, rows_total, seg, resampled_len, d_out, launches, false);
    const size_t smem = P.tables + (size_t)groups * P.per_group;
    const uint64_t units = dual ? (rows_total + 1) / 2 : rows_total;      // rows or row pairs
    const uint64_t total_groups = (uint64_t)sm_count * groups;
    int force_blocks = 0;
    if (const char* g = std::getenv("BIRDA_K2_ITEM_BLOCKS")) force_blocks = atoi(g);       // tests: runs of this many blocks everywhere
    P.wi = plan_work_items(units, total_groups, P.nblk, hop < src_seg, force_blocks);
    P.nitems = P.wi.nitems;
    cudaError_t e = cudaMemsetAsync(P.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t ctas = (P.nitems + groups - 1) / groups;
    if (ctas > (uint64_t)sm_count) ctas = sm_count;
    const unsigned threads = (unsigned)(groups * gw * 32);
    bool launched = false;
    int i = 0;
#define BB_CT(NAME, NI, NO, TH, ...)                                                                                    \
    if (!launched && i == rs.ct_index) {                                                                                 \
        if (dual) {                                                                                                      \
            e = launch_plan2<__VA_ARGS__, TH>(source_kind(d_pcm, fmt, channels), (unsigned)ctas, threads, smem, st, P);  \
            if (e != cudaSuccess) return e;                                                                              \
        } else {                                                                                                         \
            e = cudaFuncSetAttribute(resample_plan_kernel<__VA_ARGS__>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                              \
            resample_plan_kernel<__VA_ARGS__><<<(unsigned)ctas, threads, smem, st>>>(P);                                 \
        }                                                                                                                \
        launched = true;                                                                                                 \
    }                                                                                                                    \
    ++i;
    BB_K2_CT_PLANS(BB_CT)
#undef BB_CT
    if (!launched) {
        if (PL.pad_a) {
            e = cudaFuncSetAttribute(resample_warp_kernel<MapPad8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            resample_warp_kernel<MapPad8><<<(unsigned)ctas, threads, smem, st>>>(P);
        } else {
            e = cudaFuncSetAttribute(resample_warp_kernel<MapId>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
            resample_warp_kernel<MapId><<<(unsigned)ctas, threads, smem, st>>>(P);
        }
    }
    e = cudaGetLastError();
    if (e == cudaSuccess && launches) *launches = 1;
    return e;
}

}  // namespace bb
