// K2 fast path — compile-time resampler plans, one WARP per resampler block.
//
// Same arithmetic contract as k2_resample.cu (the generic kernel); what changes is the mapping:
//   * a warp owns a run of consecutive blocks of one window; the overlap-add carry never leaves
//     its registers, and there is no CTA-wide barrier after the table load;
//   * both half-length complex FFTs run IN PLACE in shared memory: forward decimation in
//     frequency (natural order in, digit-reversed out), inverse decimation in time
//     (digit-reversed in, natural order out); the fused split/filter/re-bin pass between them
//     absorbs both permutations through two small index tables;
//   * the first forward stage reads PCM straight from global memory (convert + downmix fused),
//     the last inverse stage writes the overlap-added samples straight to the output tensor;
//   * radices, sub-transform sizes and strides are template constants; twiddles are one table
//     read (w^1) plus an in-register power chain per butterfly.
// Everything here is __host__ __device__ so tests/host_k2_check.cu can execute the same code
// lane by lane on the CPU at build time.
#pragma once
#include <cmath>
#include <cstdint>
#include "fft_butterflies.cuh"

namespace bb {
namespace k2f {

template <int... Rs> struct RSeq {
    static constexpr int count = sizeof...(Rs);
    static constexpr int at(int i) { constexpr int v[] = {Rs...}; return v[i]; }
    static constexpr int prod_upto(int t) { int p = 1; for (int i = 0; i < t; ++i) p *= at(i); return p; }
    static constexpr int total() { return prod_upto(count); }
};

// FWD: DIF radices in application order (product N).  INV: DIT radices in application order
// (product M); the last one must be even (its two output halves are "store now" / "carry").
template <class FWD, class INV_>
struct Plan {
    using Fwd = FWD; using Inv = INV_;
    static constexpr int N = FWD::total();
    static constexpr int M = INV_::total();
    static constexpr int NKEEP = N < M ? N + 1 : M;
    static constexpr int QL = INV_::at(INV_::count - 1);
    static constexpr int ML = M / QL;                       // butterflies of the last inverse stage
    static constexpr int CARRY_ITERS = (ML + 31) / 32;
    static constexpr int twf_len() { int m = 1; for (int t = 0; t + 1 < FWD::count; ++t) { int v = N / FWD::at(t); if (v > m) m = v; } return m; }
    static constexpr int twi_len() { int m = 1; for (int t = 1; t < INV_::count; ++t) { int v = M / INV_::at(t); if (v > m) m = v; } return m; }
    static constexpr int TWF = twf_len();
    static constexpr int TWI = twi_len();
    static constexpr int HALFM = M / 2;
    static_assert(QL % 2 == 0, "last inverse radix must be even");
    static_assert(M % 2 == 0, "M must be even");
};

// a[k] *= w^k for k = 1..R-1 with a log-depth power chain
template <int R> BB_HD void apply_twiddle_chain(float2 (&a)[R], float2 w) {
    float2 pw[R];
    pw[1 % R] = w;
#pragma unroll
    for (int k = 2; k < R; ++k) pw[k] = cmul(pw[k / 2], pw[k - k / 2]);
#pragma unroll
    for (int k = 1; k < R; ++k) a[k] = cmul(a[k], pw[k]);
}

// ---- forward DIF, in place.  Stage with sub-transform length NSUB inside a length-NTOT array.
template <int R, int NSUB, int NTOT, bool TW>
BB_HD void dif_stage_lane(float2* __restrict__ buf, const float2* __restrict__ tw, int lane) {
    constexpr int m = NSUB / R, NBF = NTOT / R, STRIDE = NTOT / NSUB, ITERS = (NBF + 31) / 32;
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
        const int q = it * 32 + lane;
        if (q < NBF) {
            const int sb = q / m, p = q - sb * m;
            float2* __restrict__ e = buf + sb * NSUB + p;
            float2 a[R];
#pragma unroll
            for (int j = 0; j < R; ++j) a[j] = e[j * m];
            Dft<R, false>::run(a);
            if (TW) apply_twiddle_chain<R>(a, tw[p * STRIDE]);
#pragma unroll
            for (int k = 0; k < R; ++k) e[k * m] = a[k];
        }
    }
}

// first forward stage: inputs come from the loader (z[n] = x[2n] + i x[2n+1], zero for n >= HALF)
template <int R, int NTOT, bool TW, class Loader>
BB_HD void dif_first_lane(const Loader& ld, float2* __restrict__ buf, const float2* __restrict__ tw, int lane) {
    constexpr int m = NTOT / R, ITERS = (m + 31) / 32, HALF = (NTOT + 1) / 2;
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
        const int q = it * 32 + lane;
        if (q < m) {
            float2 a[R];
#pragma unroll
            for (int j = 0; j < R; ++j) {
                a[j] = make_float2(0.f, 0.f);
                if (j * m < HALF) { if (q + j * m < HALF) a[j] = ld(q + j * m); }
            }
            Dft<R, false>::run(a);
            if (TW) apply_twiddle_chain<R>(a, tw[q]);
            float2* __restrict__ e = buf + q;
#pragma unroll
            for (int k = 0; k < R; ++k) e[k * m] = a[k];
        }
    }
}

// ---- inverse DIT, in place.  Stage whose sub-transforms have length MPREV*Q.
template <int Q, int MPREV, int MTOT, bool TW>
BB_HD void dit_stage_lane(float2* __restrict__ buf, const float2* __restrict__ tw, int lane) {
    constexpr int NT = MPREV * Q, NBF = MTOT / Q, STRIDE = MTOT / NT, ITERS = (NBF + 31) / 32;
#pragma unroll 2
    for (int it = 0; it < ITERS; ++it) {
        const int q = it * 32 + lane;
        if (q < NBF) {
            const int sb = q / MPREV, p = q - sb * MPREV;
            float2* __restrict__ e = buf + sb * NT + p;
            float2 a[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) a[j] = e[j * MPREV];
            if (TW) apply_twiddle_chain<Q>(a, tw[p * STRIDE]);
            Dft<Q, true>::run(a);
#pragma unroll
            for (int k = 0; k < Q; ++k) e[k * MPREV] = a[k];
        }
    }
}

// last inverse stage: outputs z'[n] = (y[2n], y[2n+1]); n < M/2 -> overlap-add with the carry and
// hand to the sink, n >= M/2 -> becomes the carry of the next block (same lane, same slot).
template <int Q, int MTOT, class Sink>
BB_HD void dit_last_lane(const float2* __restrict__ buf, const float2* __restrict__ tw, int lane,
                         float (&carry)[(MTOT / Q + 31) / 32][Q], Sink& sink) {
    constexpr int m = MTOT / Q, ITERS = (m + 31) / 32;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
        const int p = it * 32 + lane;
        if (p < m) {
            float2 a[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) a[j] = buf[p + j * m];
            if (MTOT != Q) apply_twiddle_chain<Q>(a, tw[p]);
            Dft<Q, true>::run(a);
#pragma unroll
            for (int k = 0; k < Q / 2; ++k) {
                sink(p + k * m, make_float2(a[k].x + carry[it][2 * k], a[k].y + carry[it][2 * k + 1]));
                carry[it][2 * k] = a[k + Q / 2].x; carry[it][2 * k + 1] = a[k + Q / 2].y;
            }
        }
    }
}

// ---- fused split / filter / re-bin / inverse pack:  A (forward result, digit-reversed) -> B
//   Y(k)  = P[k] Z[k] + Q[k] conj(Z[N-k])                       (k < NKEEP, else 0)
//   Z'(k) = Y(k) + conj(Y(M-k)) + i wi[k] (Y(k) - conj(Y(M-k)))
template <class PL>
BB_HD void split_lane(const float2* __restrict__ A, float2* __restrict__ B, const uint16_t* __restrict__ pos_f,
                      const uint16_t* __restrict__ pos_i, const float2* __restrict__ Pt, const float2* __restrict__ Qt,
                      const float2* __restrict__ WI, int lane) {
    constexpr int N = PL::N, M = PL::M, NKEEP = PL::NKEEP, HALF = PL::HALFM;
#pragma unroll 2
    for (int k = lane; k <= HALF; k += 32) {
        const int k2 = M - k;
        float2 yk = make_float2(0.f, 0.f), yk2 = make_float2(0.f, 0.f);
        if (k < NKEEP) {
            const float2 zk = A[pos_f[k == N ? 0 : k]], zn = cconj(A[pos_f[k == 0 ? 0 : N - k]]);
            yk = cadd(cmul(Pt[k], zk), cmul(Qt[k], zn));
        }
        if (k2 < NKEEP) {
            const float2 zk = A[pos_f[k2 == N ? 0 : k2]], zn = cconj(A[pos_f[N - k2]]);
            yk2 = cadd(cmul(Pt[k2], zk), cmul(Qt[k2], zn));
        }
        if (k == 0) { yk.y = 0.f; yk2.y = 0.f; }           // DC / Nyquist are real (realfft ignores their imag)
        const float2 wi = WI[k];
        {
            const float2 e = cadd(yk, cconj(yk2)), o = cmul(wi, csub(yk, cconj(yk2)));
            B[pos_i[k]] = make_float2(e.x - o.y, e.y + o.x);
        }
        if (k != 0 && k2 != k) {
            const float2 wi2 = make_float2(-wi.x, wi.y);    // exp(i pi (M-k)/M) = -conj(wi)
            const float2 e = cadd(yk2, cconj(yk)), o = cmul(wi2, csub(yk2, cconj(yk)));
            B[pos_i[k2]] = make_float2(e.x - o.y, e.y + o.x);
        }
    }
}

// ---- stage drivers (compile-time recursion over the radix lists)
template <class PL, class Exec, int T> struct FwdRest {
    static BB_HD void run(float2* A, const float2* twf) {
        if constexpr (T < PL::Fwd::count) {
            constexpr int R = PL::Fwd::at(T), NSUB = PL::N / PL::Fwd::prod_upto(T);
            constexpr bool TW = T + 1 < PL::Fwd::count;
            Exec::each([&](int lane, int) { dif_stage_lane<R, NSUB, PL::N, TW>(A, twf, lane); });
            FwdRest<PL, Exec, T + 1>::run(A, twf);
        }
    }
};
template <class PL, class Exec, int T> struct InvMid {
    static BB_HD void run(float2* B, const float2* twi) {
        if constexpr (T + 1 < PL::Inv::count) {
            constexpr int Q = PL::Inv::at(T), MPREV = PL::Inv::prod_upto(T);
            Exec::each([&](int lane, int) { dit_stage_lane<Q, MPREV, PL::M, (T > 0)>(B, twi, lane); });
            InvMid<PL, Exec, T + 1>::run(B, twi);
        }
    }
};

template <class PL> struct Tables {          // shared by all warps of a CTA (device: shared memory)
    const float2* twf;       // [TWF]  exp(-2 pi i k / N)
    const float2* twi;       // [TWI]  exp(+2 pi i k / M)
    const uint16_t* pos_f;   // [N]    position of forward bin k after the in-place DIF
    const uint16_t* pos_i;   // [M]    position the in-place DIT expects inverse bin k at
    const float2* Pt;        // [NKEEP]
    const float2* Qt;        // [NKEEP]
    const float2* WI;        // [M/2+1] exp(+i pi k / M)
};

template <class PL> struct LaneCarry { float c[PL::CARRY_ITERS][PL::QL]; };

// One resampler block: loader -> forward FFT in A -> split into B -> inverse FFT -> sink (+carry)
template <class PL, class Exec, class Loader, class SinkFactory>
BB_HD void process_block(const Tables<PL>& T, float2* A, float2* B, const Loader& ld,
                         LaneCarry<PL>* carry /* [Exec::kSlots] */, SinkFactory& sinks) {
    constexpr int R0 = PL::Fwd::at(0);
    Exec::each([&](int lane, int) { dif_first_lane<R0, PL::N, (PL::Fwd::count > 1)>(ld, A, T.twf, lane); });
    FwdRest<PL, Exec, 1>::run(A, T.twf);
    Exec::each([&](int lane, int) { split_lane<PL>(A, B, T.pos_f, T.pos_i, T.Pt, T.Qt, T.WI, lane); });
    InvMid<PL, Exec, 0>::run(B, T.twi);
    Exec::each([&](int lane, int slot) {
        auto sink = sinks(lane);
        dit_last_lane<PL::QL, PL::M>(B, T.twi, lane, carry[slot].c, sink);
    });
}

// Host-side table construction (double precision, rounded once)
inline void build_pos_tables(const int* fwd, int nf, const int* inv, int ni, int N, int M,
                             uint16_t* pos_f, uint16_t* pos_i) {
    for (int k = 0; k < N; ++k) {           // DIF: k = k1 + r1 (k2 + r2 (...)), pos = sum k_t * N/(r1..rt)
        int rem = k, pos = 0, prod = 1;
        for (int t = 0; t < nf; ++t) { int d = rem % fwd[t]; rem /= fwd[t]; prod *= fwd[t]; pos += d * (N / prod); }
        pos_f[k] = (uint16_t)pos;
    }
    for (int n = 0; n < M; ++n) {           // DIT: n = j_S + q_S (j_{S-1} + ...), pos = sum j_t * (q1..q_{t-1})
        int rem = n, pos = 0;
        for (int t = ni - 1; t >= 0; --t) {
            int d = rem % inv[t]; rem /= inv[t];
            int prev = 1; for (int u = 0; u < t; ++u) prev *= inv[u];
            pos += d * prev;
        }
        pos_i[n] = (uint16_t)pos;
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
        // B = -0.5 i w H  ->  (w H) = (wr hr - wi hi) + i (wr hi + wi hr);  -0.5 i (x + i y) = 0.5 y - 0.5 i x
        const double xr = wr * hr - wi * hi, xi = wr * hi + wi * hr;
        const double br = 0.5 * xi, bi = -0.5 * xr;
        Pt[k] = make_float2((float)(ar + br), (float)(ai + bi));
        Qt[k] = make_float2((float)(ar - br), (float)(ai - bi));
    }
    for (int k = 0; k <= M / 2; ++k) WI[k] = make_float2((float)cos(pi * k / M), (float)sin(pi * k / M));
}

}  // namespace k2f
}  // namespace bb
