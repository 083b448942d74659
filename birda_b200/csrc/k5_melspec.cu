// K5 — spectrogram prefix of a classifier graph on the device (north_star: "an STFT +
// mel-filterbank kernel whose mel projection runs on tcgen05 tensor cores").
//
//   packed windows [rows, samples] f32 (what K1 / K2 wrote)
//     -> K5a  stft_power2_kernel: frame t = samples [t*hop, t*hop + n_fft) * window, real FFT of length n_fft as a
//                                 half-length complex FFT in shared memory (the K2 butterflies and stage tables),
//                                 |X|^power of the bins the mel filters touch -> P [frames, Kpad] f32 (L2-sized chunks)
//     -> K5b  mel_gemm_kernel   : D[frame, mel] = sum_k P[frame, k] * W[mel, k] on the 5th-gen tensor cores
//                                 (tcgen05.mma kind::tf32, accumulator in TMEM), run as THREE tf32 products of an
//                                 error-free hi/lo split (P_hi W_hi + P_lo W_hi + P_hi W_lo, "3xTF32") so that the
//                                 result carries f32 accuracy; epilogue (tcgen05.ld) applies the log scaling and
//                                 stores [rows, n_mels, n_frames].
//
// The reference has no code for this step (the spectrogram lives inside the ONNX graph, and no file in the
// reference pins frame length, hop, mel edges or scaling — SURVEY.md §8f-3), so the layer is a general one: window,
// mel weights and scaling are inputs; parity is against the float64 statement of the layer kept with the tests.
#include "common.cuh"
#include "guard.hpp"
#include "k2_warp.cuh"
#include <cuda.h>               // CUtensorMap and its enums only: the encoder is looked up at run time (cudaGetDriverEntryPoint)
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

struct bb_melspec {
    bb_ctx* ctx = nullptr;
    bb_melspec_cfg cfg{};
    int N = 0;                          // complex FFT length = n_fft / 2
    uint32_t bin_lo = 0, nb = 0, kpad = 0;    // mel support [bin_lo, bin_lo + nb), padded to a multiple of 32
    bb::k2w::RtPlan plan{};
    float2* d_twf = nullptr; uint16_t* d_posf = nullptr; float2* d_wk = nullptr;   // stage twiddles, digit reversal, exp(-i pi k / N)
    float* d_window = nullptr;
    float* d_w = nullptr;                                                            // [n_mels, 2 * kpad]: tf32 hi | lo parts per 32-bin chunk
    float* d_P = nullptr; uint64_t p_capacity_frames = 0;                            // [frames (multiple of 128), kpad]
};

namespace bb {
namespace {

using namespace bb::k2w;

// ------------------------------------------------------------------------------------------ K5a
constexpr int kFftGroupWarps = 2;        // warps per thread group

struct FftExec {
    int glane, nl, bar_id;
    __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nl) : "memory"); }
    template <class F> __device__ __forceinline__ void each(F&& f) const { f(glane, nl); sync(); }
};

// error-free split for the tf32 pipe: hi keeps the top 19 bits (sign, exponent, 10 mantissa bits), lo = v - hi exactly
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    lo = v - hi;
}
// P row layout: per chunk of 32 bins, 32 hi values then 32 lo values (one 256-byte run per frame and chunk, so the
// GEMM's producers copy straight into its hi and lo operand tiles)
__device__ __forceinline__ void store_split(float* __restrict__ row, uint32_t j, float v) {
    float hi, lo;
    split_tf32(v, hi, lo);
    float* q = row + (size_t)(j >> 5) * 64 + (j & 31u);
    q[0] = hi; q[32] = lo;
}

struct StftParams {
    RtPlan plan;
    const float* x; const float* window; const float2* twf; const uint16_t* posf; const float2* wk;
    float* P;
    uint32_t samples, hop, n_frames, n_fft, bin_lo, nb, kpad;
    uint64_t frame0, nframes;            // flattened frame range of this launch (frame = row * n_frames + t)
    float power;
    uint32_t off_posf, off_wk, off_win, tables, per_group;
};

// A thread group transforms TWO frames at once (all power-of-two frame lengths 256..4096): every element is a float4
// (re0, re1, im0, im1) and every butterfly instruction a packed f32x2 op (the K2 trick), stage constants are
// compile-time.  6 groups of 2 warps per CTA: 64 lanes = the butterfly count of the radix-16 stage of N = 1024.
constexpr int kFft2Threads = 384;
template <class FWD> struct CtFwdPlan {              // forward half of a k2w::CtPlan (the inverse half is not used here)
    using Plan = CtPlan<FWD, RSeq<2>>;
    static constexpr int N = FWD::total();
};
template <class PL, int T> struct Fft2Rest {
    template <class Ex> static __device__ __forceinline__ void run(const Ex& ex, float4* A, const float4* twf) {
        if constexpr (T < PL::Fwd::count) {
            using S = typename PL::template FwdStage<T>;
            ex.each([&](int l, int n_l) { dif_stage<S::radix, cx2, MapPad8>(A, Tw<cx2, false>{twf}, S{}, l, n_l); });
            Fft2Rest<PL, T + 1>::run(ex, A, twf);
        }
    }
};

template <class FWD>
__global__ void __launch_bounds__(kFft2Threads, 1)
stft_power2_kernel(const __grid_constant__ StftParams p) {
    using PL = typename CtFwdPlan<FWD>::Plan;
    constexpr int N = CtFwdPlan<FWD>::N;
    extern __shared__ __align__(16) unsigned char smem[];
    const RtPlan& RP = p.plan;
    float4* s_twf = reinterpret_cast<float4*>(smem);
    uint16_t* s_posf = reinterpret_cast<uint16_t*>(smem + p.off_posf);
    float2* s_wk = reinterpret_cast<float2*>(smem + p.off_wk);
    float2* s_win = reinterpret_cast<float2*>(smem + p.off_win);
    const int NT = blockDim.x;                   // 64 threads per group; as many groups as fit next to the tables (<= 6)
    for (int i = threadIdx.x; i < RP.twf_len; i += NT) { const float2 w = p.twf[i]; s_twf[i] = make_float4(w.x, w.x, w.y, w.y); }
    for (int i = threadIdx.x; i < N; i += NT) s_posf[i] = p.posf[i];
    for (int i = threadIdx.x; i <= N; i += NT) s_wk[i] = p.wk[i];
    for (int i = threadIdx.x; i < N; i += NT) s_win[i] = make_float2(p.window[2 * i], p.window[2 * i + 1]);
    __syncthreads();

    const int warp = threadIdx.x >> 5, group = warp / kFftGroupWarps;
    const int kGroups = (int)blockDim.x / 32 / kFftGroupWarps;
    FftExec ex;
    ex.nl = kFftGroupWarps * 32;
    ex.glane = (warp - group * kFftGroupWarps) * 32 + (int)(threadIdx.x & 31);
    ex.bar_id = 1 + group;
    const int lane = ex.glane, nl = ex.nl;
    float4* A = reinterpret_cast<float4*>(smem + p.tables + (size_t)group * p.per_group);
    const uint64_t npairs = (p.nframes + 1) / 2;

    for (uint64_t pi = (uint64_t)blockIdx.x * kGroups + group; pi < npairs; pi += (uint64_t)gridDim.x * kGroups) {
        const float* xr[2]; uint32_t s0[2]; bool act[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint64_t fi = 2 * pi + s;
            act[s] = fi < p.nframes;
            const uint64_t f = p.frame0 + (act[s] ? fi : 0);
            const uint64_t row = f / p.n_frames;
            xr[s] = p.x + row * p.samples;
            s0[s] = (uint32_t)(f - row * p.n_frames) * p.hop;
        }
        const bool fast = act[0] && act[1] && s0[0] + 2u * N <= p.samples && s0[1] + 2u * N <= p.samples &&
                          ((reinterpret_cast<uintptr_t>(xr[0] + s0[0]) | reinterpret_cast<uintptr_t>(xr[1] + s0[1])) & 7u) == 0;
        if (fast) {
            const float2* __restrict__ x0 = reinterpret_cast<const float2*>(xr[0] + s0[0]);
            const float2* __restrict__ x1 = reinterpret_cast<const float2*>(xr[1] + s0[1]);
            auto ld = [&](int n) -> cx2 {
                const float2 a = __ldg(x0 + n), b = __ldg(x1 + n), w = s_win[n];
                cx2 c; c.re = make_float2(a.x * w.x, b.x * w.x); c.im = make_float2(a.y * w.y, b.y * w.y);
                return c;
            };
            using S0 = typename PL::template FwdStage<0>;
            ex.each([&](int l, int n_l) { dif_first<S0::radix, cx2, MapPad8>(A, Tw<cx2, false>{s_twf}, S0{}, N, ld, l, n_l); });
        } else {
            auto ld = [&](int n) -> cx2 {
                float v[2][2];
#pragma unroll
                for (int s = 0; s < 2; ++s) {
                    const uint32_t i0 = s0[s] + 2u * (uint32_t)n;
                    v[s][0] = act[s] && i0 < p.samples ? __ldg(xr[s] + i0) : 0.f;
                    v[s][1] = act[s] && i0 + 1u < p.samples ? __ldg(xr[s] + i0 + 1u) : 0.f;
                }
                const float2 w = s_win[n];
                cx2 c; c.re = make_float2(v[0][0] * w.x, v[1][0] * w.x); c.im = make_float2(v[0][1] * w.y, v[1][1] * w.y);
                return c;
            };
            using S0 = typename PL::template FwdStage<0>;
            ex.each([&](int l, int n_l) { dif_first<S0::radix, cx2, MapPad8>(A, Tw<cx2, false>{s_twf}, S0{}, N, ld, l, n_l); });
        }
        Fft2Rest<PL, 1>::run(ex, A, s_twf);
        float* __restrict__ P0 = p.P + (2 * pi) * (2ull * p.kpad);
        float* __restrict__ P1 = P0 + 2ull * p.kpad;
        for (uint32_t j = lane; j < p.kpad; j += nl) {
            float2 v = make_float2(0.f, 0.f);
            if (j < p.nb) {
                const int k = (int)(p.bin_lo + j);
                const cx2 zk = Mem<cx2>::ld(A + MapPad8::at(s_posf[k == N ? 0 : k])), zn = cconj(Mem<cx2>::ld(A + MapPad8::at(s_posf[k == 0 ? 0 : N - k])));
                const cx2 e = cadd(zk, zn), d = csub(zk, zn);
                const cx2 o = rot90<false>(cmul(Mem<cx2>::bcast(s_wk[k]), d));      // -i w d
                const float2 xr_ = kmulc(kadd(e.re, o.re), 0.5f), xi_ = kmulc(kadd(e.im, o.im), 0.5f);
                const float2 pw = kfma(xr_, xr_, kmul(xi_, xi_));
                if (p.power == 2.0f) v = pw;
                else if (p.power == 1.0f) v = make_float2(sqrtf(pw.x), sqrtf(pw.y));
                else v = make_float2(powf(pw.x, 0.5f * p.power), powf(pw.y, 0.5f * p.power));
            }
            store_split(P0, j, v.x);
            if (act[1]) store_split(P1, j, v.y);
        }
        ex.sync();
    }
}

// ------------------------------------------------------------------------------------------ K5b
constexpr int kBM = 128;                 // frames per tile = UMMA M
constexpr int kBK = 32;                  // tf32 elements per K chunk = one 128-byte swizzle row
constexpr int kGemmMaxStages = 4;
constexpr int kGemmThreads = 160;        // warps 0-3: producers, then epilogue (TMEM lane quarters); warp 4: TMEM owner + MMA issuer

struct GemmParams {
    const float* P; const float* W; float* out;      // both in the chunked hi | lo layout (store_split)
    uint32_t kpad, n_mels, n_frames, tmem_cols, idesc, stages;
    uint64_t frame0, nframes;            // flattened frames of this launch; P row 0 = frame0
    int32_t log_mode; float log_eps;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        uint32_t done;
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
// K-major, 128-byte swizzle, rows of 128 bytes, 8-row groups 1024 bytes apart (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                 ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 2)
mel_gemm_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment for the 128-byte swizzle atoms
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = kBM * 128, b_bytes = p.n_mels * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    __shared__ __align__(8) uint64_t s_full[kGemmMaxStages], s_empty[kGemmMaxStages], s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nk = p.kpad / kBK;

    if (tid == 0) {
        for (int s = 0; s < kGemmMaxStages; ++s) { mbar_init(&s_full[s], 128); mbar_init(&s_empty[s], 1); }
        mbar_init(&s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kBM;         // first frame of the tile, relative to frame0

    if (warp < 4) {
        // ===== producers: a classic multi-stage cp.async pipeline.  A frame's (or mel row's) chunk is 256 contiguous
        // bytes in global memory: 8 pieces of 16 bytes for the hi tile, 8 for the lo tile, each copied straight to its
        // place in the canonical K-major 128B-swizzled layout (piece c of row r lands at r*128 + ((c ^ (r & 7)) << 4)).
        const uint32_t S = p.stages;
        auto issue = [&](uint32_t kc) {
            unsigned char* st = smem + (size_t)(kc % S) * stage_bytes;
            const uint32_t a_pieces = kBM * 16u, w_pieces = p.n_mels * 16u;
            for (uint32_t piece = tid; piece < a_pieces; piece += 128u) {
                const uint32_t r = piece >> 4, c16 = piece & 15u, c = c16 & 7u, half = c16 >> 3;
                const uint64_t fr = tile0 + r;
                const bool in = fr < p.nframes;
                const float* g = p.P + (in ? fr : 0) * (2ull * p.kpad) + (size_t)kc * 64 + c16 * 4;
                const uint32_t d = smem_u32(st + half * a_bytes + r * 128u + ((c ^ (r & 7u)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(g), "r"(in ? 16 : 0) : "memory");
            }
            for (uint32_t piece = tid; piece < w_pieces; piece += 128u) {
                const uint32_t r = piece >> 4, c16 = piece & 15u, c = c16 & 7u, half = c16 >> 3;
                const float* g = p.W + (size_t)r * (2ull * p.kpad) + (size_t)kc * 64 + c16 * 4;
                const uint32_t d = smem_u32(st + 2 * a_bytes + half * b_bytes + r * 128u + ((c ^ (r & 7u)) << 4));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory");
            }
        };
        // prologue: S - 1 chunks in flight
        for (uint32_t kc = 0; kc + 1 < S; ++kc) {
            if (kc < nk) issue(kc);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        for (uint32_t kc = 0; kc < nk; ++kc) {
            const uint32_t nxt = kc + S - 1;
            if (nxt < nk) {
                mbar_wait(&s_empty[nxt % S], ((nxt / S) & 1u) ^ 1u);        // the MMAs that read this slot have retired
                issue(nxt);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            // chunk kc has landed once at most S - 1 newer groups are pending
            if (S == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (S == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
            else asm volatile("cp.async.wait_group 3;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
            mbar_arrive(&s_full[kc % S]);
        }
        // ===== epilogue: TMEM -> registers -> scaling -> [rows, n_mels, n_frames]
        mbar_wait(&s_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t fr = tile0 + (uint32_t)(warp * 32 + lane);             // this thread's frame = TMEM lane
        const bool live = fr < p.nframes;
        const uint64_t f = p.frame0 + fr;
        const uint64_t row = f / p.n_frames;
        const uint32_t t = (uint32_t)(f - row * p.n_frames);
        float* __restrict__ o = p.out + (row * p.n_mels) * (uint64_t)p.n_frames + t;
        for (uint32_t c0 = 0; c0 < p.n_mels; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v = __uint_as_float(r[j]);
                    if (p.log_mode == 1) v = logf(v + p.log_eps);
                    else if (p.log_mode == 2) v = 10.0f * log10f(fmaxf(v, p.log_eps));
                    o[(uint64_t)(c0 + j) * p.n_frames] = v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===== MMA issuer: one elected lane
        if (lane == 0) {
            for (uint32_t kc = 0; kc < nk; ++kc) {
                const uint32_t s = kc % p.stages, ph = (kc / p.stages) & 1u;
                mbar_wait(&s_full[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
                for (uint32_t k4 = 0; k4 < 4; ++k4) {          // 4 x (K = 8 tf32 = 32 bytes) inside the 128-byte row
                    const uint64_t a_hi = umma_desc(base + k4 * 32u), a_lo = umma_desc(base + a_bytes + k4 * 32u);
                    const uint64_t b_hi = umma_desc(base + 2 * a_bytes + k4 * 32u), b_lo = umma_desc(base + 2 * a_bytes + b_bytes + k4 * 32u);
                    umma_tf32(tmem, a_hi, b_hi, p.idesc, (kc | k4) != 0u ? 1u : 0u);
                    umma_tf32(tmem, a_lo, b_hi, p.idesc, 1u);
                    umma_tf32(tmem, a_hi, b_lo, p.idesc, 1u);
                }
                umma_commit(&s_empty[s]);                      // the stage is free once these MMAs have read it
            }
            umma_commit(&s_done);                              // accumulator complete
        }
        __syncwarp();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ K5b, TMA-fed
// The same GEMM with its operands moved by the TMA unit: one elected thread issues four cp.async.bulk.tensor.2d boxes
// per K chunk (P hi, P lo, W hi, W lo: [rows, 32 tf32] with the 128-byte swizzle the UMMA descriptors expect) and
// arms the stage's mbarrier with the byte count; frames past the end of the launch are zero-filled by the unit.  No
// thread touches operand data, the four epilogue warps only wait for the accumulator, and stages are as deep as
// shared memory allows.  SASS: UTMALDG.
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int32_t x, int32_t y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(tm), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__global__ void __launch_bounds__(kGemmThreads, 2)
mel_gemm_tma_kernel(const __grid_constant__ GemmParams p, const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmW) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t a_bytes = kBM * 128, b_bytes = p.n_mels * 128;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    __shared__ __align__(8) uint64_t s_full[kGemmMaxStages], s_empty[kGemmMaxStages], s_done;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t nk = p.kpad / kBK;

    if (tid == 0) {
        for (int s = 0; s < kGemmMaxStages; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], 1); }
        mbar_init(&s_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kBM;

    if (warp < 4) {
        if (tid == 0) {
            // ===== producer: one thread, TMA
            const uint32_t S = p.stages;
            for (uint32_t kc = 0; kc < nk; ++kc) {
                const uint32_t s = kc % S;
                if (kc >= S) mbar_wait(&s_empty[s], ((kc / S) & 1u) ^ 1u);       // the MMAs that read this slot have retired
                const uint32_t st = smem_u32(smem + (size_t)s * stage_bytes);
                mbar_expect_tx(&s_full[s], stage_bytes);
                tma_load_2d(st, &tmP, (int32_t)(kc * 64), (int32_t)tile0, &s_full[s]);
                tma_load_2d(st + a_bytes, &tmP, (int32_t)(kc * 64 + 32), (int32_t)tile0, &s_full[s]);
                tma_load_2d(st + 2 * a_bytes, &tmW, (int32_t)(kc * 64), 0, &s_full[s]);
                tma_load_2d(st + 2 * a_bytes + b_bytes, &tmW, (int32_t)(kc * 64 + 32), 0, &s_full[s]);
            }
        }
        __syncwarp();
        // ===== epilogue: TMEM -> registers -> scaling -> [rows, n_mels, n_frames]
        mbar_wait(&s_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint64_t fr = tile0 + (uint32_t)(warp * 32 + lane);             // this thread's frame = TMEM lane
        const bool live = fr < p.nframes;
        const uint64_t f = p.frame0 + fr;
        const uint64_t row = f / p.n_frames;
        const uint32_t t = (uint32_t)(f - row * p.n_frames);
        float* __restrict__ o = p.out + (row * p.n_mels) * (uint64_t)p.n_frames + t;
        for (uint32_t c0 = 0; c0 < p.n_mels; c0 += 16) {
            uint32_t r[16];
            const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                         : "r"(taddr) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (live) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v = __uint_as_float(r[j]);
                    if (p.log_mode == 1) v = logf(v + p.log_eps);
                    else if (p.log_mode == 2) v = 10.0f * log10f(fmaxf(v, p.log_eps));
                    o[(uint64_t)(c0 + j) * p.n_frames] = v;
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else {
        // ===== MMA issuer: one elected lane
        if (lane == 0) {
            for (uint32_t kc = 0; kc < nk; ++kc) {
                const uint32_t s = kc % p.stages, ph = (kc / p.stages) & 1u;
                mbar_wait(&s_full[s], ph);                     // the unit has written all four boxes of the stage
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t base = smem_u32(smem + (size_t)s * stage_bytes);
#pragma unroll
                for (uint32_t k4 = 0; k4 < 4; ++k4) {
                    const uint64_t a_hi = umma_desc(base + k4 * 32u), a_lo = umma_desc(base + a_bytes + k4 * 32u);
                    const uint64_t b_hi = umma_desc(base + 2 * a_bytes + k4 * 32u), b_lo = umma_desc(base + 2 * a_bytes + b_bytes + k4 * 32u);
                    umma_tf32(tmem, a_hi, b_hi, p.idesc, (kc | k4) != 0u ? 1u : 0u);
                    umma_tf32(tmem, a_lo, b_hi, p.idesc, 1u);
                    umma_tf32(tmem, a_hi, b_lo, p.idesc, 1u);
                }
                umma_commit(&s_empty[s]);
            }
            umma_commit(&s_done);
        }
        __syncwarp();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
    }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links cudart statically, not libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// [rows, cols] f32 row-major, boxes of [box_rows, 32] with the 128-byte swizzle; rows past the end read as zeros
bool make_tile_map(CUtensorMap* tm, const float* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || box_rows == 0 || box_rows > 256) return false;
    const cuuint64_t dims[2] = {cols, rows}, strides[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {32, box_rows}, estr[2] = {1, 1};
    return fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace
}  // namespace bb

// ------------------------------------------------------------------------------------------ C ABI
using namespace bb;

extern "C" {

void bb_melspec_destroy(bb_melspec* m) {
    if (!m) return;
    bb::DeviceGuard dev_guard(m->ctx->device);
    cudaStreamSynchronize(m->ctx->stream);
    void* ptrs[] = {m->d_twf, m->d_posf, m->d_wk, m->d_window, m->d_w, m->d_P};
    for (void* q : ptrs) if (q) cudaFree(q);
    delete m;
}

int32_t bb_melspec_create(bb_ctx* c, const bb_melspec_cfg* cfg, const float* window, const float* mel_weights, bb_melspec** out) {
    BB_TRY
    if (!c || !cfg || !window || !mel_weights || !out) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    const uint32_t nfft = cfg->n_fft;
    if (nfft < 256 || nfft > 4096 || (nfft & (nfft - 1)) != 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "n_fft must be a power of two in [256, 4096]");
    if (cfg->hop == 0 || cfg->n_frames == 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "hop and n_frames must be > 0");
    if (cfg->n_mels < 16 || cfg->n_mels > 256 || cfg->n_mels % 16 != 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "n_mels must be a multiple of 16 in [16, 256]");
    if (!(cfg->power > 0.f)) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "power must be > 0");
    if (cfg->log_mode < 0 || cfg->log_mode > 2) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "log_mode must be 0, 1 or 2");
    bb_melspec* m = new (std::nothrow) bb_melspec();
    if (!m) BB_SET_ERR(c, BB_ERR_OOM, "out of host memory");
    struct Owner { bb_melspec* m; ~Owner() { if (m) bb_melspec_destroy(m); } } owner{m};      // released on success
    m->ctx = c; m->cfg = *cfg; m->N = (int)nfft / 2;
    const uint32_t bins = nfft / 2 + 1;
    // support of the mel filters: only those bins are produced and multiplied
    uint32_t lo = bins, hi = 0;
    for (uint32_t r = 0; r < cfg->n_mels; ++r)
        for (uint32_t k = 0; k < bins; ++k)
            if (mel_weights[(size_t)r * bins + k] != 0.0f) { lo = k < lo ? k : lo; hi = k + 1 > hi ? k + 1 : hi; }
    if (hi <= lo) { lo = 0; hi = 1; }
    m->bin_lo = lo; m->nb = hi - lo; m->kpad = (m->nb + 31u) & ~31u;
    std::vector<int> fwd, inv;
    if (!k2w::choose_radices(m->N, false, &fwd)) BB_SET_ERR(c, BB_ERR_INTERNAL, "no radix plan");
    // only the forward half of the plan is used; the inverse half is filled in to keep the plan well formed
    if (!k2w::choose_radices(m->N, true, &inv) || !k2w::build_plan_from_radices(m->N, m->N, m->N, &m->plan, &fwd, &inv)) {
        BB_SET_ERR(c, BB_ERR_INTERNAL, "no radix plan");
    }
    m->plan.half_in = m->N;                                              // no zero padding: every input slot is live
    std::vector<float2> twf(m->plan.twf_len), twi(m->plan.twi_len), wk(m->N + 1);
    k2w::build_twiddles(m->plan, twf.data(), twi.data());
    std::vector<uint16_t> pf(m->N), pi_(m->N);
    k2w::build_pos_tables(fwd, inv, m->N, m->N, pf.data(), pi_.data());
    const double pi = 3.14159265358979323846;
    for (int k = 0; k <= m->N; ++k) wk[k] = make_float2((float)cos(-pi * k / m->N), (float)sin(-pi * k / m->N));
    std::vector<float> w2((size_t)cfg->n_mels * m->kpad * 2, 0.f);        // same chunked hi | lo layout as P
    for (uint32_t r = 0; r < cfg->n_mels; ++r)
        for (uint32_t j = 0; j < m->nb; ++j) {
            const float w = mel_weights[(size_t)r * bins + lo + j];
            uint32_t b; memcpy(&b, &w, 4); b &= 0xFFFFE000u;
            float h; memcpy(&h, &b, 4);
            float* q = &w2[(size_t)r * m->kpad * 2 + (size_t)(j >> 5) * 64 + (j & 31u)];
            q[0] = h; q[32] = w - h;
        }
    bb::DeviceGuard dev_guard(c->device);
    auto up = [&](const void* h, size_t bytes, void** d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes);
        return e != cudaSuccess ? e : cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = up(twf.data(), twf.size() * 8, (void**)&m->d_twf);
    if (e == cudaSuccess) e = up(pf.data(), pf.size() * 2, (void**)&m->d_posf);
    if (e == cudaSuccess) e = up(wk.data(), wk.size() * 8, (void**)&m->d_wk);
    if (e == cudaSuccess) e = up(window, (size_t)nfft * 4, (void**)&m->d_window);
    if (e == cudaSuccess) e = up(w2.data(), w2.size() * 4, (void**)&m->d_w);
    if (e != cudaSuccess) {
        std::string msg = std::string("melspec init: ") + cudaGetErrorString(e);
        BB_SET_ERR(c, e == cudaErrorMemoryAllocation ? BB_ERR_OOM : BB_ERR_CUDA, msg);
    }
    owner.m = nullptr;
    *out = m;
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

int32_t bb_melspec_info(const bb_melspec* m, uint32_t* bin_lo, uint32_t* n_bins, uint32_t* k_padded) {
    if (!m) return BB_ERR_INVALID_ARG;
    if (bin_lo) *bin_lo = m->bin_lo;
    if (n_bins) *n_bins = m->nb;
    if (k_padded) *k_padded = m->kpad;
    return BB_OK;
}

int32_t bb_melspec_run(bb_melspec* m, const float* d_segments, uint32_t rows, uint32_t samples, float* d_out) {
    BB_TRY
    if (!m) return BB_ERR_INVALID_ARG;
    bb_ctx* c = m->ctx;
    if (rows == 0) return BB_OK;
    if (!d_segments || !d_out || samples == 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null argument");
    BB_DEVICE(c, c->device);
    const bb_melspec_cfg& cfg = m->cfg;
    // P is produced and consumed in chunks of whole rows: L2-sized (<= 64 MB) when that still fills the GPU
    const uint64_t row_bytes = (uint64_t)cfg.n_frames * m->kpad * 8;       // hi and lo parts
    uint64_t rows_per_chunk = (64ull << 20) / row_bytes;
    // ... but never so few that the GEMM grid (one CTA per 128 frames, two CTAs per SM) drops below ~4 waves
    const uint64_t min_rows = ((uint64_t)c->sm_count * 2 * 4 * kBM + cfg.n_frames - 1) / cfg.n_frames;
    if (rows_per_chunk < min_rows) rows_per_chunk = min_rows;
    if (rows_per_chunk == 0) rows_per_chunk = 1;
    if (rows_per_chunk > rows) rows_per_chunk = rows;
    const uint64_t cap_frames = ((rows_per_chunk * cfg.n_frames + kBM - 1) / kBM) * kBM;
    if (cap_frames > m->p_capacity_frames) {
        if (m->d_P) { BB_CUDA_OK(c, cudaStreamSynchronize(c->stream)); cudaFree(m->d_P); m->d_P = nullptr; m->p_capacity_frames = 0; }
        BB_CUDA_OK(c, cudaMalloc((void**)&m->d_P, cap_frames * m->kpad * 8));
        m->p_capacity_frames = cap_frames;
    }
    StftParams sp{};
    sp.plan = m->plan; sp.x = d_segments; sp.window = m->d_window; sp.twf = m->d_twf; sp.posf = m->d_posf; sp.wk = m->d_wk;
    sp.P = m->d_P; sp.samples = samples; sp.hop = cfg.hop; sp.n_frames = cfg.n_frames; sp.n_fft = cfg.n_fft;
    sp.bin_lo = m->bin_lo; sp.nb = m->nb; sp.kpad = m->kpad; sp.power = cfg.power;
    auto a16 = [](size_t x) { return (uint32_t)((x + 15) & ~(size_t)15); };
    GemmParams gp{};
    gp.P = m->d_P; gp.W = m->d_w; gp.out = d_out;
    gp.kpad = m->kpad; gp.n_mels = cfg.n_mels; gp.n_frames = cfg.n_frames;
    gp.tmem_cols = cfg.n_mels <= 32 ? 32 : cfg.n_mels <= 64 ? 64 : cfg.n_mels <= 128 ? 128 : 256;
    // cute::UMMA::InstrDescriptor: c = F32 (1 << 4), a = b = TF32 (2 << 7, 2 << 10), K-major both, N >> 3 at bit 17, M >> 4 at bit 24
    gp.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((cfg.n_mels >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
    gp.log_mode = cfg.log_mode; gp.log_eps = cfg.log_eps;
    // two stages when two CTAs then share an SM (their main loops and epilogues overlap), else as many as fit
    const size_t stage_bytes = 2 * (size_t)kBM * 128 + 2 * (size_t)cfg.n_mels * 128;
    gp.stages = 2;
    if (2 * (2 * stage_bytes + 2048) > 227 * 1024) { gp.stages = (uint32_t)((220 * 1024) / stage_bytes); if (gp.stages > kGemmMaxStages) gp.stages = kGemmMaxStages; }
    // operands by TMA (default) when the W tiles keep their 1024-byte alignment; BIRDA_K5_CPASYNC=1 keeps the cp.async producers
    bool use_tma = cfg.n_mels % 8 == 0 && encode_tiled_fn() != nullptr;
    if (const char* e = std::getenv("BIRDA_K5_CPASYNC")) if (e[0] == '1') use_tma = false;
    CUtensorMap tmW;
    if (use_tma && !make_tile_map(&tmW, m->d_w, cfg.n_mels, 2ull * m->kpad, cfg.n_mels)) use_tma = false;
    if (use_tma) {                         // nobody stages by hand: as many stages as two CTAs per SM allow, at least 3 when one fits
        gp.stages = (uint32_t)((113 * 1024 - 1024) / stage_bytes);
        if (gp.stages < 3) gp.stages = (uint32_t)((220 * 1024) / stage_bytes);
        if (gp.stages > kGemmMaxStages) gp.stages = kGemmMaxStages;
        if (gp.stages < 2) use_tma = false;
    }
    const size_t gemm_smem = (size_t)gp.stages * stage_bytes + 1024;
    if (use_tma) BB_CUDA_OK(c, cudaFuncSetAttribute(mel_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
    else BB_CUDA_OK(c, cudaFuncSetAttribute(mel_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem));
    for (uint64_t r0 = 0; r0 < rows; r0 += rows_per_chunk) {
        const uint64_t nr = rows - r0 < rows_per_chunk ? rows - r0 : rows_per_chunk;
        const uint64_t nframes = nr * cfg.n_frames;
        sp.frame0 = r0 * cfg.n_frames; sp.nframes = nframes;
        {
            StftParams s2 = sp;
            s2.off_posf = a16((size_t)m->plan.twf_len * 16);
            s2.off_wk = s2.off_posf + a16((size_t)m->N * 2);
            s2.off_win = s2.off_wk + a16((size_t)(m->N + 1) * 8);
            s2.tables = s2.off_win + a16((size_t)cfg.n_fft * 4);
            s2.per_group = a16((size_t)(m->N + m->N / 8) * 16);      // MapPad8: one slot of padding after every eight
            int groups2 = (int)((227 * 1024 - s2.tables) / s2.per_group);
            if (groups2 > kFft2Threads / 32 / kFftGroupWarps) groups2 = kFft2Threads / 32 / kFftGroupWarps;
            if (groups2 < 1) BB_SET_ERR(c, BB_ERR_INTERNAL, "STFT tables do not fit in shared memory");
            const unsigned threads2 = (unsigned)groups2 * 32 * kFftGroupWarps;
            const size_t smem2 = s2.tables + (size_t)groups2 * s2.per_group;
            const uint64_t want2 = ((nframes + 1) / 2 + groups2 - 1) / groups2;
            const unsigned grid2 = (unsigned)(want2 < (uint64_t)c->sm_count ? want2 : (uint64_t)c->sm_count);
#define BB_STFT2(...)                                                                                                              \
            { BB_CUDA_OK(c, cudaFuncSetAttribute(stft_power2_kernel<RSeq<__VA_ARGS__>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2)); \
              stft_power2_kernel<RSeq<__VA_ARGS__>><<<grid2, threads2, smem2, c->stream>>>(s2); }
            // radices as choose_radices picks them for these lengths (the plan's tables are laid out for them)
            if (m->N == 2048) BB_STFT2(16, 16, 8)
            else if (m->N == 1024) BB_STFT2(16, 8, 8)
            else if (m->N == 512) BB_STFT2(8, 8, 8)
            else if (m->N == 256) BB_STFT2(16, 16)
            else BB_STFT2(16, 8)
#undef BB_STFT2
        }
        gp.frame0 = sp.frame0; gp.nframes = nframes;
        CUtensorMap tmP;
        if (use_tma && make_tile_map(&tmP, m->d_P, nframes, 2ull * m->kpad, kBM))
            mel_gemm_tma_kernel<<<(unsigned)((nframes + kBM - 1) / kBM), kGemmThreads, gemm_smem, c->stream>>>(gp, tmP, tmW);
        else {
            if (use_tma) BB_SET_ERR(c, BB_ERR_INTERNAL, "tensor map for the power spectra could not be encoded");
            mel_gemm_kernel<<<(unsigned)((nframes + kBM - 1) / kBM), kGemmThreads, gemm_smem, c->stream>>>(gp);
        }
        BB_CUDA_OK(c, cudaGetLastError());
        c->launches += 2;
    }
    return BB_OK;
    BB_CATCH((m && m->ctx ? &m->ctx->last_error : nullptr))
}

}  // extern "C"
