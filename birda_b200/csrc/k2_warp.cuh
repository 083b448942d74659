// K2 resampler core (product code): everything the per-window block-FFT resampler does between the staged PCM of a
// block and its output samples, written once for the device kernels (k2_warp.cu) and for the host self check
// (tests/host_k2_check.cu runs the same templates lane by lane on the CPU).
//
// A thread group (1-6 warps) owns a run of consecutive resampler blocks of one window — or of TWO adjacent windows in
// lockstep (element type cx2: float4 slots, packed f32x2 math):
//   * forward half-length complex FFT in place (decimation in frequency: natural order in, digit-reversed out) in
//     buffer A; inverse in place (decimation in time: digit-reversed in, natural order out) in buffer B;
//   * the fused split / filter / re-bin / inverse-pack pass between them is driven by a host-built layout
//     (build_split_layout): per (k, M-k) pair the six positions it touches, flags, and the filter tables in visiting
//     order, the order chosen to keep the scattered accesses off each other's banks;
//   * stages come from a small stage table (RtPlan, one code body per radix and direction: the kernel stays in the
//     instruction cache) or from compile-time plans (CtPlan: strides are immediates);
//   * buffers of power-of-two-like transforms use a padded address map (MapPad8), middle stages whose default
//     lane -> butterfly mapping conflicts read an order table (build_stage_orders);
//   * twiddles: one conflict-free table read (w^1) per butterfly plus an in-register power chain, full tables for the
//     big radices;
//   * the overlap-add carry is per-group shared memory; the last inverse stage adds it and hands the samples to a sink.
#pragma once
#include <cmath>
#include <cstdint>
#include <type_traits>
#include <utility>
#include <vector>
#include "fft_butterflies.cuh"

namespace bb {
namespace k2w {

constexpr int kMaxStages = 8;
#ifndef BB_K2W_UNROLL
#define BB_K2W_UNROLL 1
#endif
// radices >= this read every twiddle power from a [k][p] table; smaller ones read w^1 and run a
// power chain in registers (the chain's pw[] array would spill for the big butterflies)
#ifndef BB_K2W_TABLE_RADIX
#define BB_K2W_TABLE_RADIX 9
#endif
BB_HD constexpr bool tw_table_mode(int radix) { return radix >= BB_K2W_TABLE_RADIX; }
#ifndef BB_K2W_SPLIT_UNROLL
#define BB_K2W_SPLIT_UNROLL BB_K2W_UNROLL
#endif
#define BB_PRAGMA(x) _Pragma(#x)
#define BB_UNROLL_N(n) BB_PRAGMA(unroll n)

template <class C> struct Mem;

struct RtStage {
    int radix;      // butterfly size
    int span;       // DIF: sub-transform length n_t;  DIT: n_t = m * radix
    int m;          // DIF: n_t / radix;               DIT: product of the previous radices
    int nbf;        // butterflies in the stage = total / radix
    int tw_off;     // offset of this stage's compact twiddle table (entries p in [0, m)); -1 = none
    uint32_t magic; // ceil(2^32 / m) for q / m  (0 when m == 1)
    int nsb;        // sub-transforms in the stage = nbf / m
    uint32_t magic_nsb;
    int sbfast;     // 1: consecutive lanes walk sub-transforms (stride = span, odd) instead of p (see decompose)
    int ord_off;    // offset of the stage's order table (build_stage_orders), -1 = default mapping
    template <class C> BB_HD int ordoff() const { return ord_off; }
    BB_HD int M_() const { return m; }
    BB_HD int SPAN_() const { return span; }
    BB_HD int NBF_() const { return nbf; }
    BB_HD int TWOFF_() const { return tw_off; }
    BB_HD int div(int q) const {
#ifdef __CUDA_ARCH__
        return m == 1 ? q : (int)__umulhi((unsigned)q, magic);
#else
        return q / m;
#endif
    }
    // butterfly index q -> (sub-transform sb, position p).  Short sub-transforms (m < 32) with an odd span are
    // walked sub-transform-first: the lanes of a warp then hit addresses `span` apart (odd stride in 8-byte
    // units: bank-conflict free) instead of several short runs that collide.
    BB_HD void decompose(int q, int& sb, int& p) const {
        if (sbfast) {
#ifdef __CUDA_ARCH__
            p = nsb == 1 ? q : (int)__umulhi((unsigned)q, magic_nsb);
#else
            p = q / nsb;
#endif
            sb = q - p * nsb;
        } else { sb = div(q); p = q - sb * m; }
    }
};

// ---- lane -> butterfly mapping of a middle stage and its bank behaviour
// Default mapping of butterfly q: position p fastest (sub-transform sb = q / m), or sub-transform fastest for short
// odd-span stages (SBFAST).  A butterfly touches slots e + j*m, e = sb*span + p, one j per instruction, so the lanes
// served together (8 with 16-byte elements, 16 with 8-byte ones: `group`) collide exactly when their e share a
// residue mod `group`.  stage_default_conflicts counts the extra wavefronts of one pass over the stage; stages that
// lose more than 10 % that way read their (e, p) from a host-built order table instead (build_stage_orders).
BB_HD constexpr bool stage_sbfast(int span, int m, int ntot) { return m < 16 && (span % 2) == 1 && ntot / span >= 8; }
BB_HD constexpr int stage_default_conflicts(int span, int m, int ntot, int group, bool pad) {
    const int radix = span / m, nbf = ntot / radix, nsb = ntot / span;
    const bool sbf = stage_sbfast(span, m, ntot);
    int cost = 0;
    for (int q0 = 0; q0 < nbf; q0 += group) {
        int cnt[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, mx = 0;
        for (int q = q0; q < q0 + group && q < nbf; ++q) {
            int sb = 0, p = 0;
            if (sbf) { p = q / nsb; sb = q - p * nsb; } else { sb = q / m; p = q - sb * m; }
            int e = sb * span + p;
            if (pad) e += e >> 3;
            const int r = ++cnt[e % group];
            mx = r > mx ? r : mx;
        }
        cost += mx - 1;
    }
    return cost;
}
BB_HD constexpr bool stage_wants_order(int span, int m, int ntot, int group, bool pad) {
    const int nbf = ntot / (span / m);
    return stage_default_conflicts(span, m, ntot, group, pad) * 10 * group > nbf;
}

// the same stage with every quantity a compile-time constant (strides become immediate offsets).
// ORDOFF: offset of the stage's order table for the two-stream element type (16-byte slots), -1 = default mapping
template <int RADIX, int SPAN, int NTOT, int TWOFF, int ORDOFF = -1>
struct CtStage {
    static constexpr int radix = RADIX;
    static BB_HD constexpr int M_() { return SPAN / RADIX; }
    static BB_HD constexpr int SPAN_() { return SPAN; }
    static BB_HD constexpr int NBF_() { return NTOT / RADIX; }
    static BB_HD constexpr int TWOFF_() { return TWOFF; }
    template <class C> static BB_HD constexpr int ordoff() { return sizeof(typename Mem<C>::T) == 16 ? ORDOFF : -1; }
    static BB_HD constexpr int div(int q) { return q / (SPAN / RADIX); }
    static constexpr int NSB = NTOT / SPAN;
    static constexpr bool SBFAST = stage_sbfast(SPAN, SPAN / RADIX, NTOT);
    static BB_HD void decompose(int q, int& sb, int& p) {
        if (SBFAST) { p = q / NSB; sb = q - p * NSB; }
        else { sb = q / (SPAN / RADIX); p = q - sb * (SPAN / RADIX); }
    }
};

struct RtPlan {
    int N, M, nkeep, half_in;    // half_in = ceil(N/2): non-zero inputs of the forward transform
    int nf, ni;
    RtStage f[kMaxStages], i[kMaxStages];
    int twf_len, twi_len;        // compact twiddle table sizes (float2 entries)
    int pad_a, pad_b;            // 1: the forward (A) / inverse (B) buffer uses MapPad8 (one slot of padding after every eight)
    int split_len;               // pairs (k, M-k) the split pass visits = M/2 + 1 (order and addresses come from a host-built table)
    int n_regular;               // the first n_regular pairs of the visiting order need no flag tests (split_regular_count)
    // Blocking (overlap-add geometry).  The transforms are 2N real points in, 2M real points out; a block consumes
    // adv_in input samples and emits adv_out output samples.  rubato's own blocking is adv_in = N, adv_out = M (the
    // filter is N taps long); up-sampling plans may use LONGER transforms over the same taps (adv_in + taps - 1 <= 2N):
    // the result is the same linear map of the window with less transform work per sample.
    int adv_in, adv_out;
    int carry_slots;             // complex slots of overlap-add carry = M - adv_out/2  (<= adv_out/2)
    int emit_k;                  // of the last inverse stage's R outputs per butterfly, the first emit_k are emitted
};

// shared-memory storage of one element: float2 for a single stream, float4 (re0, re1, im0, im1) for two
template <class C> struct Mem;
template <> struct Mem<float2> {
    using T = float2;
    static BB_HD float2 ld(const float2* p) { return *p; }
    static BB_HD void st(float2* p, float2 v) { *p = v; }
    static BB_HD float2 bcast(float2 w) { return w; }                      // twiddle / table entry for the stream(s)
};
template <> struct Mem<cx2> {
    using T = float4;
    static BB_HD cx2 ld(const float4* p) { const float4 v = *p; cx2 c; c.re = make_float2(v.x, v.y); c.im = make_float2(v.z, v.w); return c; }
    static BB_HD void st(float4* p, const cx2& v) { *p = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
    static BB_HD cx2 bcast(float2 w) { cx2 c; c.re = make_float2(w.x, w.x); c.im = make_float2(w.y, w.y); return c; }
};

// Twiddle table pointer.  Full form: entries in the element's own storage layout (two streams: float4 (x, x, y, y), one
// LDS.128 and no moves).  Compact form: float2 entries broadcast to the stream(s) at load time — half the shared memory,
// for plans whose buffers leave no room (CtPlan::COMPACT_TW).
template <class C, bool COMPACT> struct Tw {
    using E = typename std::conditional<COMPACT, float2, typename Mem<C>::T>::type;
    const E* p;
    BB_HD C operator[](int i) const {
        if constexpr (COMPACT) return Mem<C>::bcast(p[i]);
        else return Mem<C>::ld(p + i);
    }
    BB_HD Tw operator+(int o) const { return Tw{p + o}; }
};

// Address map of a transform buffer: logical slot -> physical slot.  MapPad8 leaves one slot of padding after every
// eight: a power-of-two plan's stride-8 stage (span 8, m = 1: lanes 8 slots = 128 bytes apart, an 8-way bank conflict
// with 16-byte elements) becomes a stride-9 walk, and the runs of 8 consecutive slots the other stages touch stay
// contiguous.
struct MapId  { static BB_HD constexpr int at(int i) { return i; } };
struct MapPad8 { static BB_HD constexpr int at(int i) { return i + (i >> 3); } };

// a[k] *= w_span^(p k), k = 1..R-1.  Table layout: chain mode [p] holds w^p; table mode [(k-1)*m + p]
template <int R, class C, class TW> BB_HD void apply_twiddles(C (&a)[R], const TW tw, int m, int p) {
    if constexpr (tw_table_mode(R)) {
#pragma unroll
        for (int k = 1; k < R; ++k) a[k] = cmul(a[k], tw[(k - 1) * m + p]);
    } else {
        C pw[R];
        pw[1] = tw[p];
#pragma unroll
        for (int k = 2; k < R; ++k) pw[k] = cmul(pw[k / 2], pw[k - k / 2]);
#pragma unroll
        for (int k = 1; k < R; ++k) a[k] = cmul(a[k], pw[k]);
    }
}

// ---- forward DIF stage, in place
template <int R, class C, class AM = MapId, class S, class TW>
BB_HD void dif_stage(typename Mem<C>::T* __restrict__ buf, const TW tw, const S& s, int lane, int nl,
                     const uint32_t* __restrict__ order = nullptr) {
    const int m = s.M_();
    const int ooff = order != nullptr ? s.template ordoff<C>() : -1;
BB_UNROLL_N(BB_K2W_UNROLL)
    for (int q = lane; q < s.NBF_(); q += nl) {
        int e, p;
        if (ooff >= 0) { const uint32_t w = order[ooff + q]; e = (int)(w & 0xffffu); p = (int)(w >> 16); }
        else { int sb; s.decompose(q, sb, p); e = sb * s.SPAN_() + p; }
        C a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = Mem<C>::ld(buf + AM::at(e + j * m));
        Dft<R, false>::run(a);
        if (s.TWOFF_() >= 0) apply_twiddles<R, C>(a, tw + s.TWOFF_(), m, p);
#pragma unroll
        for (int k = 0; k < R; ++k) Mem<C>::st(buf + AM::at(e + k * m), a[k]);
    }
}

// first forward stage: input z[n] comes from `ld(n)` for n < half_in, zero above
template <int R, class C, class AM = MapId, class S, class Loader, class TW>
BB_HD void dif_first(typename Mem<C>::T* __restrict__ buf, const TW tw, const S& s, int half_in,
                     const Loader& ld, int lane, int nl) {
    const int m = s.M_();
BB_UNROLL_N(BB_K2W_UNROLL)
    for (int q = lane; q < m; q += nl) {
        C a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int n = q + j * m;
            // j * m >= half_in: this input lies in the zero padding for every q (a compile-time fact for compile-time
            // plans: the load, its conversion and the butterfly terms it feeds fold away)
            a[j] = (j * m < half_in && n < half_in) ? ld(n) : czero<C>();
        }
        Dft<R, false>::run(a);
        if (s.TWOFF_() >= 0) apply_twiddles<R, C>(a, tw + s.TWOFF_(), m, q);
#pragma unroll
        for (int k = 0; k < R; ++k) Mem<C>::st(buf + AM::at(q + k * m), a[k]);
    }
}

// ---- inverse DIT stage, in place
template <int R, class C, class AM = MapId, class S, class TW>
BB_HD void dit_stage(typename Mem<C>::T* __restrict__ buf, const TW tw, const S& s, int lane, int nl,
                     const uint32_t* __restrict__ order = nullptr) {
    const int m = s.M_();
    const int ooff = order != nullptr ? s.template ordoff<C>() : -1;
BB_UNROLL_N(BB_K2W_UNROLL)
    for (int q = lane; q < s.NBF_(); q += nl) {
        int e, p;
        if (ooff >= 0) { const uint32_t w = order[ooff + q]; e = (int)(w & 0xffffu); p = (int)(w >> 16); }
        else { int sb; s.decompose(q, sb, p); e = sb * s.SPAN_() + p; }
        C a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = Mem<C>::ld(buf + AM::at(e + j * m));
        if (s.TWOFF_() >= 0) apply_twiddles<R, C>(a, tw + s.TWOFF_(), m, p);
        Dft<R, true>::run(a);
#pragma unroll
        for (int k = 0; k < R; ++k) Mem<C>::st(buf + AM::at(e + k * m), a[k]);
    }
}

// last inverse stage (span == M): output z'[n] = (y[2n], y[2n+1]), n = p + k*m.  The first `ke` values of k are this
// block's output samples (n < adv_out/2), the rest the overlap-add carry of the next block; the old carry covers the
// first R - ke values.  adv_out/2 is a multiple of m, so a carry slot is read (k = j) and rewritten (k = j + ke) by
// the same lane: all reads first, then the writes (no cross-lane hazard).  rubato's blocking: ke = R/2.
template <int R, class C, class AM = MapId, class S, class Sink, class TW>
BB_HD void dit_last(const typename Mem<C>::T* __restrict__ buf, const TW tw, const S& s,
                    typename Mem<C>::T* __restrict__ carry, const Sink& sink, int lane, int nl, const int ke) {
    const int m = s.M_();
BB_UNROLL_N(BB_K2W_UNROLL)
    for (int p = lane; p < m; p += nl) {
        C a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = Mem<C>::ld(buf + AM::at(p + j * m));
        if (s.TWOFF_() >= 0) apply_twiddles<R, C>(a, tw + s.TWOFF_(), m, p);
        Dft<R, true>::run(a);
#pragma unroll
        for (int k = 0; k < R; ++k) if (k < R - ke) a[k] = cadd(a[k], Mem<C>::ld(carry + p + k * m));
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if (k < ke) sink(p + k * m, a[k]);
            else Mem<C>::st(carry + p + (k - ke) * m, a[k]);
        }
    }
}

// Pairs (k, M-k) with both spectra present, two outputs and no DC special case: k in [max(1, M - nkeep + 1), M/2 - 1].
// They come first in the visiting order and take a branch-free path.
BB_HD constexpr int split_regular_count(int M, int nkeep) {
    const int lo = (M - nkeep + 1) > 1 ? (M - nkeep + 1) : 1, hi = M / 2 - 1;
    return hi >= lo ? hi - lo + 1 : 0;
}

// Split-pass entry of one (k, M-k) pair, built on the host (build_split_layout): the positions of the four forward
// results it reads and of the two inverse inputs it writes (digit reversal folded in), plus flags.
//   x = offA(Z[k]) | offA(Z[N-k]) << 16     y = offA(Z[k2]) | offA(Z[N-k2]) << 16     z = offB(k) | offB(k2) << 16
//   (offsets in units of 8 bytes: position for single-stream elements, scaled by sizeof(element) / 8 at the access —
//    one LEA; below 64 K for every plan the warp kernels take)
//   w = flags: 1 = Y(k) present (k < nkeep), 2 = Y(k2) present, 4 = k == 0 (DC / Nyquist are real), 8 = write Z'(k2)
constexpr unsigned kSplitHasK = 1u, kSplitHasK2 = 2u, kSplitDc = 4u, kSplitStore2 = 8u;

template <class C, bool COMPACT_TW = false> struct Tables {
    using T = typename Mem<C>::T;
    Tw<C, COMPACT_TW> twf, twi;
    const uint4* sidx;                       // [split_len]
    const float4* pq1; const float4* pq2;    // [split_len] (P[k], Q[k]) and (P[k2], Q[k2]); broadcast to the stream(s) at load time
    const float2* WI;                        // [split_len] exp(+i pi k / M)
    const uint32_t* order_f; const uint32_t* order_i;   // per-stage order tables (e | p << 16), may be null
};

// ---- fused split / filter / re-bin / inverse pack:  A (digit-reversed forward result) -> B
//   Y(k)  = P[k] Z[k] + Q[k] conj(Z[N-k])                       (k < nkeep, else 0)
//   Z'(k) = Y(k) + conj(Y(M-k)) + i wi[k] (Y(k) - conj(Y(M-k)))
// one 128-bit shared-memory load of a split-table entry (left to itself the compiler splits it into two 64-bit loads
// because the halves are consumed at different times: twice the wavefronts at a 16-byte lane stride)
BB_HD uint4 ld_entry128(const uint4* p) {
#ifdef __CUDA_ARCH__
    uint4 v;
    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
#else
    return *p;
#endif
}

template <class C> BB_HD const typename Mem<C>::T* at_off(const typename Mem<C>::T* base, unsigned off) {
    return reinterpret_cast<const typename Mem<C>::T*>(reinterpret_cast<const char*>(base) + (size_t)off * (sizeof(typename Mem<C>::T) / 8));
}
template <class C> BB_HD typename Mem<C>::T* at_off(typename Mem<C>::T* base, unsigned off) {
    return reinterpret_cast<typename Mem<C>::T*>(reinterpret_cast<char*>(base) + (size_t)off * (sizeof(typename Mem<C>::T) / 8));
}

template <class C, class TB>
BB_HD void split_pass(const typename Mem<C>::T* __restrict__ A, typename Mem<C>::T* __restrict__ B, const TB& T,
                      int L, int lane, int nl, int n_regular) {
    // regular pairs: both Y(k) and Y(M-k) exist, both outputs are written, nothing is forced real
BB_UNROLL_N(BB_K2W_SPLIT_UNROLL)
    for (int idx = lane; idx < n_regular; idx += nl) {
        const uint4 e = ld_entry128(T.sidx + idx);
        const float4 pa = T.pq1[idx], pb = T.pq2[idx];
        const C zk = Mem<C>::ld(at_off<C>(A, e.x & 0xffffu)), zn = cconj(Mem<C>::ld(at_off<C>(A, e.x >> 16)));
        const C zk2 = Mem<C>::ld(at_off<C>(A, e.y & 0xffffu)), zn2 = cconj(Mem<C>::ld(at_off<C>(A, e.y >> 16)));
        const C yk = cadd(cmul(Mem<C>::bcast(make_float2(pa.x, pa.y)), zk), cmul(Mem<C>::bcast(make_float2(pa.z, pa.w)), zn));
        const C yk2 = cadd(cmul(Mem<C>::bcast(make_float2(pb.x, pb.y)), zk2), cmul(Mem<C>::bcast(make_float2(pb.z, pb.w)), zn2));
        const C wi = Mem<C>::bcast(T.WI[idx]);
        {
            const C ev = cadd(yk, cconj(yk2)), o = cmul(wi, csub(yk, cconj(yk2)));
            Mem<C>::st(at_off<C>(B, e.z & 0xffffu), Cx<C>::make(ksub(Cx<C>::re(ev), Cx<C>::im(o)), kadd(Cx<C>::im(ev), Cx<C>::re(o))));
        }
        {
            const C wi2 = Cx<C>::make(kneg(Cx<C>::re(wi)), Cx<C>::im(wi));    // exp(i pi (M-k)/M) = -conj(wi)
            const C ev = cadd(yk2, cconj(yk)), o = cmul(wi2, csub(yk2, cconj(yk)));
            Mem<C>::st(at_off<C>(B, e.z >> 16), Cx<C>::make(ksub(Cx<C>::re(ev), Cx<C>::im(o)), kadd(Cx<C>::im(ev), Cx<C>::re(o))));
        }
    }
    // the rest (k = 0, k = M/2, pairs whose Y(M-k) lies beyond the kept bins): flags decide
    for (int idx = n_regular + lane; idx < L; idx += nl) {
        const uint4 e = ld_entry128(T.sidx + idx);      // the tables live in shared memory in every kernel that runs this pass
        const unsigned fl = e.w;
        C yk = czero<C>(), yk2 = czero<C>();
        if (fl & kSplitHasK) {
            const float4 pq = T.pq1[idx];
            const C zk = Mem<C>::ld(at_off<C>(A, e.x & 0xffffu)), zn = cconj(Mem<C>::ld(at_off<C>(A, e.x >> 16)));
            yk = cadd(cmul(Mem<C>::bcast(make_float2(pq.x, pq.y)), zk), cmul(Mem<C>::bcast(make_float2(pq.z, pq.w)), zn));
        }
        if (fl & kSplitHasK2) {
            const float4 pq = T.pq2[idx];
            const C zk = Mem<C>::ld(at_off<C>(A, e.y & 0xffffu)), zn = cconj(Mem<C>::ld(at_off<C>(A, e.y >> 16)));
            yk2 = cadd(cmul(Mem<C>::bcast(make_float2(pq.x, pq.y)), zk), cmul(Mem<C>::bcast(make_float2(pq.z, pq.w)), zn));
        }
        if (fl & kSplitDc) {                            // DC / Nyquist are real (realfft ignores their imag)
            yk = Cx<C>::make(Cx<C>::re(yk), kzero(typename Cx<C>::K()));
            yk2 = Cx<C>::make(Cx<C>::re(yk2), kzero(typename Cx<C>::K()));
        }
        const C wi = Mem<C>::bcast(T.WI[idx]);
        {
            const C ev = cadd(yk, cconj(yk2)), o = cmul(wi, csub(yk, cconj(yk2)));
            Mem<C>::st(at_off<C>(B, e.z & 0xffffu), Cx<C>::make(ksub(Cx<C>::re(ev), Cx<C>::im(o)), kadd(Cx<C>::im(ev), Cx<C>::re(o))));
        }
        if (fl & kSplitStore2) {
            const C wi2 = Cx<C>::make(kneg(Cx<C>::re(wi)), Cx<C>::im(wi));    // exp(i pi (M-k)/M) = -conj(wi)
            const C ev = cadd(yk2, cconj(yk)), o = cmul(wi2, csub(yk2, cconj(yk)));
            Mem<C>::st(at_off<C>(B, e.z >> 16), Cx<C>::make(ksub(Cx<C>::re(ev), Cx<C>::im(o)), kadd(Cx<C>::im(ev), Cx<C>::re(o))));
        }
    }
}

// radix dispatch: one code body per (radix, direction) in the whole kernel
#define BB_K2W_RADIX_SWITCH(R_EXPR, CALL)                                                       \
    switch (R_EXPR) {                                                                           \
        case 2: { constexpr int R = 2; CALL; } break;   case 3: { constexpr int R = 3; CALL; } break;   \
        case 4: { constexpr int R = 4; CALL; } break;   case 5: { constexpr int R = 5; CALL; } break;   \
        case 6: { constexpr int R = 6; CALL; } break;   case 7: { constexpr int R = 7; CALL; } break;   \
        case 8: { constexpr int R = 8; CALL; } break;   case 9: { constexpr int R = 9; CALL; } break;   \
        case 16: { constexpr int R = 16; CALL; } break; case 19: { constexpr int R = 19; CALL; } break; \
        case 11: { constexpr int R = 11; CALL; } break; case 13: { constexpr int R = 13; CALL; } break; \
        default: break;                                                                         \
    }
#define BB_K2W_EVEN_RADIX_SWITCH(R_EXPR, CALL)                                                  \
    switch (R_EXPR) {                                                                           \
        case 2: { constexpr int R = 2; CALL; } break;   case 4: { constexpr int R = 4; CALL; } break;   \
        case 6: { constexpr int R = 6; CALL; } break;   case 8: { constexpr int R = 8; CALL; } break;   \
        case 16: { constexpr int R = 16; CALL; } break;                                         \
        default: break;                                                                         \
    }

// One block, in two halves.  ex.each(f) runs f(lane, nlanes) for every lane of the thread group that owns
// the block and orders memory between calls (device: __syncwarp or a named barrier; host harness: a loop).
// forward half: loader -> in-place DIF in A -> (before_split) -> split / filter / re-bin into B
template <class C, class AM = MapId, class Exec, class Loader, class BeforeSplit, class TB>
BB_HD void forward_half(const Exec& ex, const RtPlan& P, const TB& T, typename Mem<C>::T* A, typename Mem<C>::T* B,
                        const Loader& ld, BeforeSplit&& before_split) {
    ex.each([&](int lane, int nl) {
        BB_K2W_RADIX_SWITCH(P.f[0].radix, (dif_first<R, C, AM>(A, T.twf, P.f[0], P.half_in, ld, lane, nl)))
    });
    for (int t = 1; t < P.nf; ++t)
        ex.each([&](int lane, int nl) { BB_K2W_RADIX_SWITCH(P.f[t].radix, (dif_stage<R, C, AM>(A, T.twf, P.f[t], lane, nl, T.order_f))) });
    before_split();
    ex.each([&](int lane, int nl) { split_pass<C>(A, B, T, P.split_len, lane, nl, P.n_regular); });
}
// inverse half: in-place DIT in B -> overlap-add with the carry -> sink
template <class C, class AM = MapId, class Exec, class Sink, class TB>
BB_HD void inverse_half(const Exec& ex, const RtPlan& P, const TB& T, typename Mem<C>::T* B, typename Mem<C>::T* carry,
                        const Sink& sink) {
    for (int t = 0; t + 1 < P.ni; ++t)
        ex.each([&](int lane, int nl) { BB_K2W_RADIX_SWITCH(P.i[t].radix, (dit_stage<R, C, AM>(B, T.twi, P.i[t], lane, nl, T.order_i))) });
    ex.each([&](int lane, int nl) {
        BB_K2W_EVEN_RADIX_SWITCH(P.i[P.ni - 1].radix, (dit_last<R, C, AM>(B, T.twi, P.i[P.ni - 1], carry, sink, lane, nl, P.emit_k)))
    });
}
// runtime plans pad both buffers or neither (AM); the plan's pad_a / pad_b say which (rt_plan_pads)
template <class C, class AM = MapId, class Exec, class Loader, class Sink, class AfterSplit, class TB>
BB_HD void process_block(const Exec& ex, const RtPlan& P, const TB& T, typename Mem<C>::T* A, typename Mem<C>::T* B,
                         typename Mem<C>::T* carry, const Loader& ld, const Sink& sink, AfterSplit&& after_split) {
    forward_half<C, AM>(ex, P, T, A, B, ld, [] {});
    after_split();
    inverse_half<C, AM>(ex, P, T, B, carry, sink);
}


// ------------------------------------------------------------------ work items
// The kernels hand out work items in order through an atomic counter.  Phase 1: while there are at least as many rows
// (row pairs in two-stream mode) left as thread groups, every group gets the same number of rows — equal items,
// (almost) nothing recomputed.  With overlapping windows a row is cut in two halves: the second half of window w and
// the first half of window w + 1 read the same PCM, and as half-row items they run at the same time on different groups,
// so the second reader hits L2 (whole-row items doubled the DRAM reads of C2).  Phase 2: the remaining rows are cut
// into runs of R blocks (each run recomputes the block before it for its carry); R minimises the modelled makespan of
// that phase, rounds x (R + 1) block times.
struct WorkItems {
    uint32_t nblk, R, items_per_row, R1, parts1;
    uint64_t nitems, n1_items, n1_rows;
};
inline WorkItems plan_work_items(uint64_t units, uint64_t groups, uint32_t nblk, bool overlapping, int force_blocks = 0) {
    WorkItems w{};
    w.nblk = nblk;
    uint64_t n1_rows = groups ? (units / groups) * groups : 0;
    w.parts1 = (overlapping && nblk >= 32) ? 2u : 1u;
    const uint64_t rem = units - n1_rows;
    uint32_t R = nblk, ipr = 1;
    if (rem > 0) {
        uint64_t best = ~0ull;
        for (uint32_t pieces = 1; pieces <= nblk; ++pieces) {
            const uint32_t r = (nblk + pieces - 1) / pieces;
            if (pieces > 1 && r < 4) break;
            const uint32_t ip = (nblk + r - 1) / r;
            const uint64_t rounds = (rem * ip + groups - 1) / groups;
            const uint64_t t = rounds * (r + (ip > 1 ? 1u : 0u));
            if (t < best) { best = t; R = r; ipr = ip; }
        }
    }
    if (force_blocks >= 1) {                       // tests: runs of this many blocks everywhere
        n1_rows = 0; R = (uint32_t)force_blocks < nblk ? (uint32_t)force_blocks : nblk; ipr = (nblk + R - 1) / R;
    }
    w.R = R; w.items_per_row = ipr;
    w.R1 = (nblk + w.parts1 - 1) / w.parts1;
    w.n1_rows = n1_rows; w.n1_items = n1_rows * w.parts1;
    w.nitems = w.n1_items + (units - n1_rows) * ipr;
    return w;
}
// item -> (row, [b0, b1), last run of its row)
BB_HD void decode_work_item(const WorkItems& w, uint64_t item, uint64_t& row, uint32_t& b0, uint32_t& b1, bool& last) {
    if (item < w.n1_items) {
        row = item / w.parts1;
        const uint32_t it = (uint32_t)(item - row * w.parts1);
        b0 = it * w.R1; b1 = b0 + w.R1 < w.nblk ? b0 + w.R1 : w.nblk; last = it + 1 == w.parts1;
    } else {
        const uint64_t j = item - w.n1_items, q = j / w.items_per_row;
        const uint32_t it = (uint32_t)(j - q * w.items_per_row);
        row = w.n1_rows + q; b0 = it * w.R; b1 = b0 + w.R < w.nblk ? b0 + w.R : w.nblk; last = it + 1 == w.items_per_row;
    }
}

// ------------------------------------------------------------------ compile-time plans
template <int... Rs> struct RSeq {
    static constexpr int count = sizeof...(Rs);
    static constexpr int at(int i) { constexpr int v[] = {Rs...}; return v[i]; }
    static constexpr int prod_upto(int t) { int p = 1; for (int i = 0; i < t; ++i) p *= at(i); return p; }
    static constexpr int total() { return prod_upto(count); }
};

// FWD: DIF radices in application order (product N).  INV: DIT radices in application order (product M,
// last one even).  Offsets follow build_plan_from_radices exactly (checked on the host at init).
// ADV_IN_T / ADV_OUT_T: the blocking (RtPlan::adv_in / adv_out); 0 = rubato's own (N, M).
// COMPACT_TW_T: the two-stream kernel keeps its twiddles as float2 (Tw<C, true>).
template <class FWD, class INV_, int ADV_IN_T = 0, int ADV_OUT_T = 0, bool COMPACT_TW_T = false>
struct CtPlan {
    static constexpr bool COMPACT_TW = COMPACT_TW_T;
    using Fwd = FWD; using Inv = INV_;
    static constexpr int N = FWD::total();
    static constexpr int M = INV_::total();
    static constexpr int NKEEP = N < M ? N + 1 : M;
    static constexpr int ADV_IN = ADV_IN_T ? ADV_IN_T : N;
    static constexpr int ADV_OUT = ADV_OUT_T ? ADV_OUT_T : M;
    static constexpr int HALF_IN = (ADV_IN + 1) / 2;
    static constexpr int CARRY = M - ADV_OUT / 2;
    static constexpr int EMIT_K = INV_::at(INV_::count - 1) * (ADV_OUT / 2) / M;
    static_assert(ADV_OUT % 2 == 0 && (ADV_OUT / 2) % (M / INV_::at(INV_::count - 1)) == 0, "adv_out/2 must be a multiple of the last inverse stage's m");
    static_assert(CARRY <= ADV_OUT / 2 && ADV_IN <= 2 * N, "the carry may only reach the next block");
    // buffers whose length is a multiple of 64 have stride-8 stages: pad them (MapPad8) — unless the transform has
    // two or more odd radices: their odd-stride walks are conflict-free only in the unpadded layout (measured:
    // padding the 2240-point buffer of 1029/2240 = 7 5 8 8 costs 37 %, padding 1536 = 3 8 8 8 gains 25 %)
    static constexpr int odd_radices(bool inverse) {
        int c = 0;
        if (inverse) { for (int i = 0; i < INV_::count; ++i) c += INV_::at(i) & 1; }
        else { for (int i = 0; i < FWD::count; ++i) c += FWD::at(i) & 1; }
        return c;
    }
    static constexpr bool PAD_A = N % 64 == 0 && odd_radices(false) <= 1, PAD_B = M % 64 == 0 && odd_radices(true) <= 1;
    using MapA = typename std::conditional<PAD_A, MapPad8, MapId>::type;
    using MapB = typename std::conditional<PAD_B, MapPad8, MapId>::type;
    static constexpr int tw_entries(int radix, int m) { return tw_table_mode(radix) ? m * (radix - 1) : m; }
    static constexpr int fwd_span(int t) { return N / FWD::prod_upto(t); }
    static constexpr int fwd_twoff(int t) {
        if (t + 1 >= FWD::count) return -1;
        int off = 0;
        for (int u = 0; u < t; ++u) off += tw_entries(FWD::at(u), fwd_span(u) / FWD::at(u));
        return off;
    }
    static constexpr int inv_span(int t) { return INV_::prod_upto(t + 1); }
    static constexpr int inv_twoff(int t) {
        if (t == 0) return -1;
        int off = 0;
        for (int u = 1; u < t; ++u) off += tw_entries(INV_::at(u), INV_::prod_upto(u));
        return off;
    }
    // order tables of the two-stream kernel (16-byte slots: groups of 8 lanes); same rule as build_stage_orders
    static constexpr bool fwd_wants(int t) { return t >= 1 && stage_wants_order(fwd_span(t), fwd_span(t) / FWD::at(t), N, 8, PAD_A); }
    static constexpr bool inv_wants(int t) { return t + 1 < INV_::count && stage_wants_order(inv_span(t), inv_span(t) / INV_::at(t), M, 8, PAD_B); }
    static constexpr int fwd_ordoff(int t) {
        if (!fwd_wants(t)) return -1;
        int off = 0;
        for (int u = 1; u < t; ++u) if (fwd_wants(u)) off += N / FWD::at(u);
        return off;
    }
    static constexpr int inv_ordoff(int t) {
        if (!inv_wants(t)) return -1;
        int off = 0;
        for (int u = 0; u < t; ++u) if (inv_wants(u)) off += M / INV_::at(u);
        return off;
    }
    // shared-memory footprint of the two-stream kernel (16-byte elements), the launcher's formula at compile time:
    // the kernel needs the lanes-per-group it will be launched with to unroll its staging loops
    static constexpr int a16(int x) { return (x + 15) & ~15; }
    static constexpr int twf_len() { int off = 0; for (int t = 0; t + 1 < FWD::count; ++t) off += tw_entries(FWD::at(t), fwd_span(t) / FWD::at(t)); return off > 0 ? off : 1; }
    static constexpr int twi_len() { int off = 0; for (int t = 1; t < INV_::count; ++t) off += tw_entries(INV_::at(t), INV_::prod_upto(t)); return off > 0 ? off : 1; }
    static constexpr int ordf_len() { int n = 0; for (int t = 1; t < FWD::count; ++t) if (fwd_wants(t)) n += N / FWD::at(t); return n; }
    static constexpr int ordi_len() { int n = 0; for (int t = 0; t + 1 < INV_::count; ++t) if (inv_wants(t)) n += M / INV_::at(t); return n; }
    static constexpr int dual_tables_bytes(int max_groups) {
        return a16(twf_len() * (COMPACT_TW ? 8 : 16)) + a16(twi_len() * (COMPACT_TW ? 8 : 16)) + 3 * a16((M / 2 + 1) * 16) + a16((M / 2 + 1) * 8) + a16(ordf_len() * 4) + a16(ordi_len() * 4) + a16(max_groups * 8);
    }
    static constexpr int dual_group_bytes() {
        return a16((PAD_A ? N + (N + 7) / 8 : N) * 16) + a16((PAD_B ? M + (M + 7) / 8 : M) * 16) + a16(CARRY * 16);
    }
    static constexpr int dual_group_lanes(int threads, int smem_max, int max_groups) {
        int groups = (smem_max - dual_tables_bytes(max_groups)) / dual_group_bytes();
        if (groups > max_groups) groups = max_groups;
        if (groups > threads / 32) groups = threads / 32;
        if (groups < 1) return 0;
        return (threads / 32) / groups * 32;
    }
    template <int T> using FwdStage = CtStage<FWD::at(T), fwd_span(T), N, fwd_twoff(T), fwd_ordoff(T)>;
    template <int T> using InvStage = CtStage<INV_::at(T), inv_span(T), M, inv_twoff(T), inv_ordoff(T)>;
    static_assert(INV_::at(INV_::count - 1) % 2 == 0, "last inverse radix must be even");
};

// A loader may offer `pick(f)`: it calls f with one of several specialised loaders (chosen by a test that is
// uniform over the thread group), so the choice is made once per block instead of once per element.
template <class L, class F> BB_HD auto loader_pick(const L& l, F&& f, int) -> decltype(l.pick(f)) { return l.pick(f); }
template <class L, class F> BB_HD void loader_pick(const L& l, F&& f, long) { f(l); }

template <class PL, class C, class Exec, int T> struct CtFwdRest {
    template <class TW> static BB_HD void run(const Exec& ex, typename Mem<C>::T* A, const TW twf, const uint32_t* order) {
        if constexpr (T < PL::Fwd::count) {
            using S = typename PL::template FwdStage<T>;
            ex.each([&](int lane, int nl) { dif_stage<S::radix, C, typename PL::MapA>(A, twf, S{}, lane, nl, order); });
            CtFwdRest<PL, C, Exec, T + 1>::run(ex, A, twf, order);
        }
    }
};
template <class PL, class C, class Exec, int T> struct CtInvMid {
    template <class TW> static BB_HD void run(const Exec& ex, typename Mem<C>::T* B, const TW twi, const uint32_t* order) {
        if constexpr (T + 1 < PL::Inv::count) {
            using S = typename PL::template InvStage<T>;
            ex.each([&](int lane, int nl) { dit_stage<S::radix, C, typename PL::MapB>(B, twi, S{}, lane, nl, order); });
            CtInvMid<PL, C, Exec, T + 1>::run(ex, B, twi, order);
        }
    }
};

template <class PL, class C, class Exec, class Loader, class BeforeSplit, class TB>
BB_HD void forward_half_ct(const Exec& ex, const TB& T, typename Mem<C>::T* A, typename Mem<C>::T* B, const Loader& ld,
                           BeforeSplit&& before_split) {
    using S0 = typename PL::template FwdStage<0>;
    loader_pick(ld, [&](const auto& l) {
        ex.each([&](int lane, int nl) { dif_first<S0::radix, C, typename PL::MapA>(A, T.twf, S0{}, PL::HALF_IN, l, lane, nl); });
    }, 0);
    CtFwdRest<PL, C, Exec, 1>::run(ex, A, T.twf, T.order_f);
    before_split();
    ex.each([&](int lane, int nl) { split_pass<C>(A, B, T, PL::M / 2 + 1, lane, nl, split_regular_count(PL::M, PL::NKEEP)); });
}
template <class PL, class C, class Exec, class Sink, class TB>
BB_HD void inverse_half_ct(const Exec& ex, const TB& T, typename Mem<C>::T* B, typename Mem<C>::T* carry, const Sink& sink) {
    CtInvMid<PL, C, Exec, 0>::run(ex, B, T.twi, T.order_i);
    using SL = typename PL::template InvStage<PL::Inv::count - 1>;
    ex.each([&](int lane, int nl) { dit_last<SL::radix, C, typename PL::MapB>(B, T.twi, SL{}, carry, sink, lane, nl, PL::EMIT_K); });
}
template <class PL, class C, class Exec, class Loader, class Sink, class AfterSplit, class TB>
BB_HD void process_block_ct(const Exec& ex, const TB& T, typename Mem<C>::T* A, typename Mem<C>::T* B,
                            typename Mem<C>::T* carry, const Loader& ld, const Sink& sink, AfterSplit&& after_split) {
    forward_half_ct<PL, C>(ex, T, A, B, ld, [] {});
    after_split();
    inverse_half_ct<PL, C>(ex, T, B, carry, sink);
}

// X(NAME, N_IN, N_OUT, DUAL_THREADS, CtPlan<RSeq<forward DIF radices>, RSeq<inverse DIT radices, last one even>>)
// DUAL_THREADS: CTA size of the two-stream kernel = register budget (65536 / threads) and warps per group;
// 640 (5 warps per window pair, 102 registers) suits the radix <= 8 plans, the radix-19 / 16 plans need 384.
// the two headline plans can be overridden from the command line for A/B builds (tools/build_variant.sh)
#ifndef BB_P1029_1120_TH
#define BB_P1029_1120_TH 640
#define BB_P1029_1120_F 7, 7, 7, 3
#define BB_P1029_1120_I 7, 5, 4, 8
#endif
#ifndef BB_P1026_684_TH
#define BB_P1026_684_TH 384
#define BB_P1026_684_F 6, 19, 9
#define BB_P1026_684_I 19, 9, 4
#endif
// 44.1k -> 48k with its own blocking: transforms of 4116 -> 4480 real points over rubato's 1029 taps, 3087 samples in /
// 3360 out per block (3/4 of the transform is payload instead of 1/2)
#ifndef BB_P2058_2240_TH
#define BB_P2058_2240_TH 640
#define BB_P2058_2240_F 7, 6, 7, 7
#define BB_P2058_2240_I 5, 7, 8, 8
#endif
#define BB_K2_CT_PLANS(X)                                                                                        \
    X(p2058_2240, 1029, 1120, BB_P2058_2240_TH, bb::k2w::CtPlan<bb::k2w::RSeq<BB_P2058_2240_F>, bb::k2w::RSeq<BB_P2058_2240_I>, 3087, 3360, true>)  /* 44.1k -> 48k, own blocking */ \
    X(p1029_1120, 1029, 1120, BB_P1029_1120_TH, bb::k2w::CtPlan<bb::k2w::RSeq<BB_P1029_1120_F>, bb::k2w::RSeq<BB_P1029_1120_I>>)  /* 44.1k -> 48k */ \
    X(p1026_684, 1026, 684, BB_P1026_684_TH, bb::k2w::CtPlan<bb::k2w::RSeq<BB_P1026_684_F>, bb::k2w::RSeq<BB_P1026_684_I>>)        /* 48k -> 32k */   \
    X(p1029_2240, 1029, 2240, 384, bb::k2w::CtPlan<bb::k2w::RSeq<7, 7, 7, 3>, bb::k2w::RSeq<7, 5, 8, 8>>)  /* 22.05k -> 48k */ \
    X(p1024_512, 1024, 512, 384, bb::k2w::CtPlan<bb::k2w::RSeq<16, 8, 8>, bb::k2w::RSeq<8, 8, 8>>)         /* 96k -> 48k */   \
    X(p1024_1536, 1024, 1536, 384, bb::k2w::CtPlan<bb::k2w::RSeq<16, 8, 8>, bb::k2w::RSeq<3, 8, 8, 8>>)    /* 32k -> 48k */   \
    X(p1323_960, 1323, 960, 384, bb::k2w::CtPlan<bb::k2w::RSeq<7, 7, 9, 3>, bb::k2w::RSeq<5, 3, 8, 8>>)    /* 44.1k -> 32k */

// ------------------------------------------------------------------ host-side plan construction
inline bool radix_supported(int r) {
    switch (r) { case 2: case 3: case 4: case 5: case 6: case 7: case 8: case 9: case 11: case 13: case 16: case 19: return true; default: return false; }
}

// factorise into supported radices; `last_even`: the final entry must be even (inverse plan).
// Order: DIF wants large contiguous runs first and odd strides last; DIT is the mirror image.
inline bool choose_radices(int n, bool inverse, std::vector<int>* out) {
    out->clear();
    std::vector<int> odd, even;
    int m = n;
    while (m % 19 == 0) { odd.push_back(19); m /= 19; }
    while (m % 13 == 0) { odd.push_back(13); m /= 13; }
    while (m % 11 == 0) { odd.push_back(11); m /= 11; }
    while (m % 7 == 0) { odd.push_back(7); m /= 7; }
    while (m % 5 == 0) { odd.push_back(5); m /= 5; }
    while (m % 9 == 0) { odd.push_back(9); m /= 9; }
    while (m % 3 == 0) { odd.push_back(3); m /= 3; }
    int e = 0;
    while (m % 2 == 0) { ++e; m /= 2; }
    if (m != 1) return false;
    if (e > 0) {                         // 2^e as ceil(e/4) radices of 2..16, bits spread evenly
        const int ns = (e + 3) / 4;
        for (int t = 0; t < ns; ++t) { int bits = e / ns + (t < e % ns ? 1 : 0); even.push_back(1 << bits); }
    }
    // fuse a lone 2 with a 3 into a 6 (one pass less)
    for (size_t i = 0; i < even.size(); ++i)
        if (even[i] == 2) {
            for (size_t j = 0; j < odd.size(); ++j)
                if (odd[j] == 3) { odd.erase(odd.begin() + j); even[i] = 6; break; }
            break;
        }
    if (inverse) {                       // DIT: odd radices first (odd strides), even ones last
        if (even.empty()) return false;  // last stage must be even
        // largest even radix last
        for (size_t i = 0; i + 1 < even.size(); ++i) if (even[i] > even.back()) std::swap(even[i], even.back());
        *out = odd; out->insert(out->end(), even.begin(), even.end());
    } else {                             // DIF: even radices first, odd ones last
        *out = even; out->insert(out->end(), odd.begin(), odd.end());
    }
    if ((int)out->size() > kMaxStages) return false;
    return true;
}

inline bool build_plan_from_radices(int N, int M, int nkeep, RtPlan* P, const std::vector<int>* fwd, const std::vector<int>* inv,
                                    int adv_in = 0, int adv_out = 0);

inline bool build_plan(int N, int M, int nkeep, RtPlan* P, std::vector<int>* fwd, std::vector<int>* inv) {
    if (M % 2 != 0 || N >= 65536 || M >= 65536) return false;
    if (!choose_radices(N, false, fwd) || !choose_radices(M, true, inv)) return false;
    return build_plan_from_radices(N, M, nkeep, P, fwd, inv);
}

inline bool build_plan_from_radices(int N, int M, int nkeep, RtPlan* P, const std::vector<int>* fwd, const std::vector<int>* inv,
                                    int adv_in, int adv_out) {
    if (M % 2 != 0 || N >= 65536 || M >= 65536) return false;
    if ((int)fwd->size() > kMaxStages || (int)inv->size() > kMaxStages || inv->empty() || inv->back() % 2) return false;
    if (adv_in <= 0) adv_in = N;
    if (adv_out <= 0) adv_out = M;
    const int m_last = M / inv->back();
    if (adv_out % 2 != 0 || (adv_out / 2) % m_last != 0 || adv_in > 2 * N || M - adv_out / 2 > adv_out / 2 || adv_out / 2 > M) return false;
    P->adv_in = adv_in; P->adv_out = adv_out; P->carry_slots = M - adv_out / 2; P->emit_k = (adv_out / 2) / m_last;
    P->N = N; P->M = M; P->nkeep = nkeep; P->half_in = (adv_in + 1) / 2; P->split_len = M / 2 + 1;
    P->n_regular = split_regular_count(M, nkeep);
    P->pad_a = 0; P->pad_b = 0;            // set by the caller: ct_plan_pads / rt_plan_pads
    P->nf = (int)fwd->size(); P->ni = (int)inv->size();
    int span = N, off = 0;
    for (int t = 0; t < P->nf; ++t) {
        RtStage& s = P->f[t];
        s.radix = (*fwd)[t]; s.span = span; s.m = span / s.radix; s.nbf = N / s.radix;
        s.tw_off = (t + 1 < P->nf) ? off : -1;
        if (s.tw_off >= 0) off += tw_table_mode(s.radix) ? s.m * (s.radix - 1) : s.m;
        s.magic = s.m <= 1 ? 0u : (uint32_t)(((1ull << 32) + s.m - 1) / s.m);
        s.nsb = N / s.span; s.magic_nsb = s.nsb <= 1 ? 0u : (uint32_t)(((1ull << 32) + s.nsb - 1) / s.nsb);
        s.sbfast = (s.m < 16 && (s.span % 2) == 1 && s.nsb >= 8) ? 1 : 0; s.ord_off = -1;
        span = s.m;
    }
    P->twf_len = off > 0 ? off : 1;
    int prev = 1; off = 0;
    for (int t = 0; t < P->ni; ++t) {
        RtStage& s = P->i[t];
        s.radix = (*inv)[t]; s.m = prev; s.span = prev * s.radix; s.nbf = M / s.radix;
        s.tw_off = t > 0 ? off : -1;
        if (s.tw_off >= 0) off += tw_table_mode(s.radix) ? s.m * (s.radix - 1) : s.m;
        s.magic = s.m <= 1 ? 0u : (uint32_t)(((1ull << 32) + s.m - 1) / s.m);
        s.nsb = M / s.span; s.magic_nsb = s.nsb <= 1 ? 0u : (uint32_t)(((1ull << 32) + s.nsb - 1) / s.nsb);
        s.sbfast = (s.m < 16 && (s.span % 2) == 1 && s.nsb >= 8) ? 1 : 0; s.ord_off = -1;
        prev = s.span;
    }
    P->twi_len = off > 0 ? off : 1;
    return true;
}

// padding rules.  Compile-time plans pad each buffer on its own (CtPlan::PAD_A / PAD_B); the runtime-plan kernel has
// one padded and one unpadded instantiation, so it pads both buffers or neither.
inline int odd_radix_count(const RtStage* st, int n) { int c = 0; for (int t = 0; t < n; ++t) c += st[t].radix & 1; return c; }
inline void ct_plan_pads(RtPlan* P) {                 // = CtPlan::PAD_A / PAD_B
    P->pad_a = P->N % 64 == 0 && odd_radix_count(P->f, P->nf) <= 1;
    P->pad_b = P->M % 64 == 0 && odd_radix_count(P->i, P->ni) <= 1;
}
inline void rt_plan_pads(RtPlan* P) {                 // positions stay 16-bit
    ct_plan_pads(P);
    P->pad_a = P->pad_b = (P->pad_a && P->pad_b && P->N < 50000 && P->M < 50000) ? 1 : 0;
}
inline int phys_len(int n, int pad) { return pad ? n + (n + 7) / 8 : n; }

// compact per-stage twiddle tables: forward stage t holds exp(-2 pi i p / span_t), p < m_t;
// inverse stage t holds exp(+2 pi i p / span_t), p < m_t
inline void build_twiddles(const RtPlan& P, float2* twf, float2* twi) {
    const double pi = 3.14159265358979323846;
    auto fill = [&](const RtStage& s, float2* tw, double sign) {
        if (s.tw_off < 0) return;
        const int nk = tw_table_mode(s.radix) ? s.radix - 1 : 1;
        for (int k = 1; k <= nk; ++k)
            for (int p = 0; p < s.m; ++p) {
                const double a = sign * 2 * pi * (double)p * k / s.span;
                tw[s.tw_off + (k - 1) * s.m + p] = make_float2((float)cos(a), (float)sin(a));
            }
    };
    for (int t = 0; t < P.nf; ++t) fill(P.f[t], twf, -1.0);
    for (int t = 0; t < P.ni; ++t) fill(P.i[t], twi, 1.0);
}

inline void build_pos_tables(const std::vector<int>& fwd, const std::vector<int>& inv, int N, int M,
                             uint16_t* pos_f, uint16_t* pos_i) {
    for (int k = 0; k < N; ++k) {           // DIF: k = k1 + r1 (k2 + r2 (...)), pos = sum k_t * N/(r1..rt)
        int rem = k, pos = 0, prod = 1;
        for (size_t t = 0; t < fwd.size(); ++t) { int d = rem % fwd[t]; rem /= fwd[t]; prod *= fwd[t]; pos += d * (N / prod); }
        pos_f[k] = (uint16_t)pos;
    }
    for (int n = 0; n < M; ++n) {           // DIT: n = j_S + q_S (j_{S-1} + ...), pos = sum j_t * (q1..q_{t-1})
        int rem = n, pos = 0;
        for (int t = (int)inv.size() - 1; t >= 0; --t) {
            int d = rem % inv[t]; rem /= inv[t];
            int prev = 1; for (int u = 0; u < t; ++u) prev *= inv[u];
            pos += d * prev;
        }
        pos_i[n] = (uint16_t)pos;
    }
}

// ---- per-stage order tables
// For every middle stage whose default mapping loses more than 10 % to bank conflicts (stage_wants_order), the
// butterflies are regrouped so that the lanes served together start at distinct residues mod `group` (and read distinct
// twiddle residues where possible).  Entry = e | p << 16 with e = sb*span + p (logical slot) and p the twiddle index.
// Offsets follow CtPlan::fwd_ordoff / inv_ordoff; `set_offsets` also records them in the plan (runtime-plan kernels).
inline void build_stage_orders(RtPlan* P, int group, bool set_offsets, std::vector<uint32_t>* of, std::vector<uint32_t>* oi) {
    auto build = [&](RtStage& s, int ntot, bool pad, std::vector<uint32_t>* out) {
        const int nbf = s.nbf, G = group;
        std::vector<int> e(nbf), p(nbf), r0(nbf), r1(nbf);
        for (int q = 0; q < nbf; ++q) {
            int sb, pp; s.decompose(q, sb, pp);
            e[q] = sb * s.span + pp; p[q] = pp;
            const int phys = pad ? MapPad8::at(e[q]) : e[q];
            r0[q] = phys % G; r1[q] = pp % G;
        }
        (void)ntot;
        std::vector<char> used(nbf, 0);
        std::vector<int> order; order.reserve(nbf);
        int left = nbf;
        while (left > 0) {
            bool s0[16] = {false}, s1[16] = {false};
            for (int n = 0; n < G && left > 0; ++n) {
                int best = -1, bc = 99;
                for (int q = 0; q < nbf && bc > 0; ++q) {
                    if (used[q]) continue;
                    const int c = (s0[r0[q]] ? 4 : 0) + (s1[r1[q]] ? 1 : 0);      // data conflicts cost R accesses, twiddle ones 1
                    if (c < bc) { bc = c; best = q; }
                }
                used[best] = 1; --left; order.push_back(best);
                s0[r0[best]] = true; s1[r1[best]] = true;
            }
        }
        const int off = (int)out->size();
        for (int q : order) out->push_back((uint32_t)e[q] | ((uint32_t)p[q] << 16));
        return off;
    };
    of->clear(); oi->clear();
    for (int t = 1; t < P->nf; ++t) {
        RtStage& s = P->f[t];
        const bool want = stage_wants_order(s.span, s.m, P->N, group, P->pad_a != 0);
        const int off = want ? build(s, P->N, P->pad_a != 0, of) : -1;
        if (set_offsets) s.ord_off = off;
    }
    for (int t = 0; t + 1 < P->ni; ++t) {
        RtStage& s = P->i[t];
        const bool want = stage_wants_order(s.span, s.m, P->M, group, P->pad_b != 0);
        const int off = want ? build(s, P->M, P->pad_b != 0, oi) : -1;
        if (set_offsets) s.ord_off = off;
    }
}

// ---- split-pass layout: visiting order of the (k, M-k) pairs and their addresses
// The six scattered accesses of a pair (four reads of A through the forward digit reversal, two writes of B through
// the inverse one) are served `group` lanes at a time (128 bytes of banks: 8 lanes of 16-byte elements, 16 lanes of
// 8-byte ones); a group is conflict-free when, for each of the six, its lanes hit distinct residues mod `group`.
// The order is built greedily group by group and polished by pair swaps (deterministic); a plain stride walk left
// ~40 % extra wavefronts on 1029/1120, this leaves ~14 %.
#ifndef BB_SPLIT_POLISH_ITERS_PER_PAIR
#define BB_SPLIT_POLISH_ITERS_PER_PAIR 100
#endif
struct SplitLayout {
    std::vector<uint4> sidx; std::vector<float4> pq1, pq2; std::vector<float2> wi;
    int extra_wavefronts = 0;      // residual conflicts of the chosen order (diagnostic)
    int n_regular = 0;             // leading entries that take the branch-free path (== split_regular_count(M, nkeep))
};

inline void build_split_layout(int N, int M, int nkeep, const uint16_t* pos_f_log, const uint16_t* pos_i_log,
                               const float2* Pt, const float2* Qt, const float2* WI, int group, SplitLayout* out,
                               int pad_a = 0, int pad_b = 0) {
    const int L = M / 2 + 1;
    // physical positions (the buffers' address maps folded in)
    std::vector<uint16_t> pfv(N), piv(M);
    for (int i = 0; i < N; ++i) pfv[i] = (uint16_t)(pad_a ? MapPad8::at(pos_f_log[i]) : pos_f_log[i]);
    for (int i = 0; i < M; ++i) piv[i] = (uint16_t)(pad_b ? MapPad8::at(pos_i_log[i]) : pos_i_log[i]);
    const uint16_t* pos_f = pfv.data(); const uint16_t* pos_i = piv.data();
    struct Item { int v[6]; unsigned flags; };
    std::vector<Item> it(L);
    for (int k = 0; k < L; ++k) {
        const int k2 = M - k;
        Item& e = it[k];
        for (int& x : e.v) x = -1;
        e.flags = 0;
        if (k < nkeep) { e.v[0] = pos_f[k == N ? 0 : k]; e.v[1] = pos_f[k == 0 ? 0 : N - k]; e.flags |= kSplitHasK; }
        if (k2 < nkeep) { e.v[2] = pos_f[k2 == N ? 0 : k2]; e.v[3] = pos_f[N - k2]; e.flags |= kSplitHasK2; }
        e.v[4] = pos_i[k];
        if (k == 0) e.flags |= kSplitDc;
        if (k != 0 && k2 != k) { e.v[5] = pos_i[k2]; e.flags |= kSplitStore2; }
    }
    const int G = group;
    // residues of the six positions of every pair (255 = the pair does not make that access)
    std::vector<unsigned char> res((size_t)L * 6);
    for (int k = 0; k < L; ++k) for (int a = 0; a < 6; ++a) res[(size_t)k * 6 + a] = it[k].v[a] >= 0 ? (unsigned char)(it[k].v[a] % G) : 255;
    auto group_cost = [&](const int* g, int n) {
        int c = 0;
        for (int a = 0; a < 6; ++a) {
            unsigned char cnt[16] = {0};
            int mx = 0;
            for (int i = 0; i < n; ++i) { const unsigned r = res[(size_t)g[i] * 6 + a]; if (r != 255) { const int v = ++cnt[r]; mx = v > mx ? v : mx; } }
            if (mx > 1) c += mx - 1;
        }
        return c;
    };
    // The visiting order of one list of pairs: greedy (fill one group at a time with the pair that collides least with
    // what the group already holds), then polish (a colliding pair of a conflicting group is offered to every seat of
    // a random other group; the best exchange is taken when it does not raise the cost — ties keep the walk moving).
    uint64_t rng = 0x9E3779B97F4A7C15ull;
    auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); };
    auto optimise = [&](const std::vector<int>& ids, std::vector<int>* order_out) -> int {
        const int Ln = (int)ids.size();
        std::vector<int> order; order.reserve(Ln);
        if (Ln == 0) return 0;
        std::vector<char> used(Ln, 0);
        int left = Ln;
        while (left > 0) {
            bool seen[6][16] = {{false}};
            for (int n = 0; n < G && left > 0; ++n) {
                int best = -1, bc = 99;
                for (int i = 0; i < Ln && bc > 0; ++i) {
                    if (used[i]) continue;
                    int c = 0;
                    for (int a = 0; a < 6; ++a) { const unsigned r = res[(size_t)ids[i] * 6 + a]; if (r != 255 && seen[a][r]) ++c; }
                    if (c < bc) { bc = c; best = i; }
                }
                used[best] = 1; --left; order.push_back(ids[best]);
                for (int a = 0; a < 6; ++a) { const unsigned r = res[(size_t)ids[best] * 6 + a]; if (r != 255) seen[a][r] = true; }
            }
        }
        const int ng = (Ln + G - 1) / G;
        auto gsize = [&](int g) { return g + 1 < ng ? G : Ln - g * G; };
        std::vector<int> cost(ng);
        int total = 0;
        for (int g = 0; g < ng; ++g) { cost[g] = group_cost(&order[(size_t)g * G], gsize(g)); total += cost[g]; }
        const int polish_iters = BB_SPLIT_POLISH_ITERS_PER_PAIR * Ln;
        for (int iter = 0; iter < polish_iters && total > 0 && ng > 1; ++iter) {
            int g1 = (int)(next() % ng);
            for (int tries = 0; tries < 16 && cost[g1] == 0; ++tries) g1 = (int)(next() % ng);
            if (cost[g1] == 0) continue;
            const int n1 = gsize(g1);
            int cand[16], nc = 0;
            for (int a = 0; a < 6; ++a) {
                int cnt[16] = {0};
                for (int i = 0; i < n1; ++i) { const unsigned r = res[(size_t)order[(size_t)g1 * G + i] * 6 + a]; if (r != 255) ++cnt[r]; }
                for (int i = 0; i < n1 && nc < 16; ++i) { const unsigned r = res[(size_t)order[(size_t)g1 * G + i] * 6 + a]; if (r != 255 && cnt[r] > 1) cand[nc++] = i; }
            }
            if (nc == 0) continue;
            const int i1 = g1 * G + cand[next() % nc];
            const int g2 = (int)(next() % ng);
            if (g1 == g2) continue;
            int best_j = -1, best_c = cost[g1] + cost[g2] + 1, best_c1 = 0, best_c2 = 0;
            for (int j = 0; j < gsize(g2); ++j) {
                const int i2 = g2 * G + j;
                std::swap(order[i1], order[i2]);
                const int c1 = group_cost(&order[(size_t)g1 * G], n1), c2 = group_cost(&order[(size_t)g2 * G], gsize(g2));
                std::swap(order[i1], order[i2]);
                if (c1 + c2 < best_c) { best_c = c1 + c2; best_j = j; best_c1 = c1; best_c2 = c2; }
            }
            if (best_j >= 0 && best_c <= cost[g1] + cost[g2]) {
                std::swap(order[i1], order[g2 * G + best_j]);
                total += best_c - cost[g1] - cost[g2]; cost[g1] = best_c1; cost[g2] = best_c2;
            }
        }
        order_out->insert(order_out->end(), order.begin(), order.end());
        return total;
    };
    // regular pairs (both spectra present, two outputs, not DC) first: the device runs them without flag tests; each
    // list is visited by its own loop (lane 0 starts at the head of the list), so each is laid out on its own
    std::vector<int> regular, rest;
    for (int k = 0; k < L; ++k) ((it[k].flags == (kSplitHasK | kSplitHasK2 | kSplitStore2)) ? regular : rest).push_back(k);
    std::vector<int> order; order.reserve(L);
    int total = optimise(regular, &order);
    total += optimise(rest, &order);
    out->n_regular = (int)regular.size();
    out->extra_wavefronts = total;
    out->sidx.resize(L); out->pq1.resize(L); out->pq2.resize(L); out->wi.resize(L);
    for (int idx = 0; idx < L; ++idx) {
        const int k = order[idx], k2 = M - k;
        const Item& e = it[k];
        auto u = [&](int p) { return (unsigned)(p < 0 ? 0 : p) * 8u; };   // units of 8 bytes (at_off scales by the element size)
        out->sidx[idx] = make_uint4(u(e.v[0]) | (u(e.v[1]) << 16), u(e.v[2]) | (u(e.v[3]) << 16), u(e.v[4]) | (u(e.v[5]) << 16), e.flags);
        out->pq1[idx] = (e.flags & kSplitHasK) ? make_float4(Pt[k].x, Pt[k].y, Qt[k].x, Qt[k].y) : make_float4(0.f, 0.f, 0.f, 0.f);
        out->pq2[idx] = (e.flags & kSplitHasK2) ? make_float4(Pt[k2].x, Pt[k2].y, Qt[k2].x, Qt[k2].y) : make_float4(0.f, 0.f, 0.f, 0.f);
        out->wi[idx] = WI[k];
    }
}

// P = A + B, Q = A - B with A = 0.5 Hf, B = -0.5 i exp(-i pi k / N) Hf;  WI[k] = exp(+i pi k / M)
inline void build_split_tables(int N, int M, int nkeep, const float* filt_re, const float* filt_im,
                               float2* Pt, float2* Qt, float2* WI) {
    const double pi = 3.14159265358979323846;
    for (int k = 0; k < nkeep; ++k) {
        const double hr = filt_re[k], hi = filt_im[k];
        const double ar = 0.5 * hr, ai = 0.5 * hi;
        const double wr = cos(-pi * k / N), wi = sin(-pi * k / N);
        const double xr = wr * hr - wi * hi, xi = wr * hi + wi * hr;      // w H
        const double br = 0.5 * xi, bi = -0.5 * xr;                        // -0.5 i (w H)
        Pt[k] = make_float2((float)(ar + br), (float)(ai + bi));
        Qt[k] = make_float2((float)(ar - br), (float)(ai - bi));
    }
    for (int k = 0; k <= M / 2; ++k) WI[k] = make_float2((float)cos(pi * k / M), (float)sin(pi * k / M));
}

// expand a float2 table into the storage type of the element (float2 as is; two streams: (x, x, y, y))
template <class C> inline std::vector<typename Mem<C>::T> expand_table(const std::vector<float2>& t);
template <> inline std::vector<float2> expand_table<float2>(const std::vector<float2>& t) { return t; }
template <> inline std::vector<float4> expand_table<cx2>(const std::vector<float2>& t) {
    std::vector<float4> o(t.size());
    for (size_t i = 0; i < t.size(); ++i) o[i] = make_float4(t[i].x, t[i].x, t[i].y, t[i].y);
    return o;
}

}  // namespace k2w
}  // namespace bb
