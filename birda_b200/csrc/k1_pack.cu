// K1 — sample conversion + channel downmix + overlapped window gather + pack into the model
// input tensor, for the paths that do not resample (source rate == model rate, and bat mode).
//
// Replaces, in one pass over HBM: append_samples (src/audio/decode.rs:353-411),
// StreamingDecoder::next_segment's copy + zero pad (src/audio/decode.rs:175-181) and the
// [B, sample_count] tensor pack inside birdnet_onnx::Classifier::predict_batch*.
//
// HBM-bound streaming kernel: algorithmic bytes = frames*channels*sizeof(sample) read once +
// rows*segment*4 written once.  Overlapped windows re-read their shared
// frames from L2.  128-bit stores always; 64/128-bit loads whenever the window start is
// aligned (warp-uniform test per row), scalar loads otherwise.
#include "common.cuh"
#include <cstdlib>
#include <type_traits>

namespace bb {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll  = 4;                       // float4 stores per thread per tile
constexpr int kSub     = kThreads * kUnroll * 4;  // output floats per pass of the general path
constexpr int kTile    = kSub * 2;                // output floats per tile

template <int FMT> struct SampleT;
template <> struct SampleT<BB_S16> { using type = int16_t; };
template <> struct SampleT<BB_S32> { using type = int32_t; };
template <> struct SampleT<BB_F32> { using type = float; };

// exact conversions of src/audio/decode.rs:371-398 (division by a power of two == exact scaling)
__device__ __forceinline__ float conv(int16_t s) { return __fmul_rn(__int2float_rn((int)s), 1.0f / 32768.0f); }
__device__ __forceinline__ float conv(int32_t s) { return __fmul_rn(__int2float_rn(s), 1.0f / 2147483648.0f); }
__device__ __forceinline__ float conv(float s)   { return s; }

// mono value of one frame: sum = 0; sum += conv(ch) left to right; sum / channels   (decode.rs:362-367)
template <typename S>
__device__ __forceinline__ float downmix_frame(const S* __restrict__ p, uint32_t channels, float fch) {
    if (channels == 1) return conv(p[0]);
    float sum = 0.0f;
    for (uint32_t c = 0; c < channels; ++c) sum = __fadd_rn(sum, conv(p[c]));
    return __fdiv_rn(sum, fch);
}

template <typename S, int CH> struct Vec4Frames;   // 4 consecutive frames of CH channels as one aligned load
template <> struct Vec4Frames<int16_t, 1> { using V = short4; static constexpr int n = 1; };
template <> struct Vec4Frames<int16_t, 2> { using V = int4;   static constexpr int n = 1; };
template <> struct Vec4Frames<int32_t, 1> { using V = int4;   static constexpr int n = 1; };
template <> struct Vec4Frames<int32_t, 2> { using V = int4;   static constexpr int n = 2; };
template <> struct Vec4Frames<float, 1>   { using V = float4; static constexpr int n = 1; };
template <> struct Vec4Frames<float, 2>   { using V = float4; static constexpr int n = 2; };

template <typename S, int CH>
__device__ __forceinline__ float4 load4_aligned(const S* __restrict__ p) {
    using VF = Vec4Frames<S, CH>;
    typename VF::V v[VF::n];
    const typename VF::V* vp = reinterpret_cast<const typename VF::V*>(p);
#pragma unroll
    for (int i = 0; i < VF::n; ++i) v[i] = __ldg(vp + i);
    const S* s = reinterpret_cast<const S*>(v);
    float4 o;
    if (CH == 1) {
        o.x = conv(s[0]); o.y = conv(s[1]); o.z = conv(s[2]); o.w = conv(s[3]);
    } else {
        // sum = 0 + l + r ; / 2  — same operation order as the reference
        o.x = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv(s[0])), conv(s[1])), 2.0f);
        o.y = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv(s[2])), conv(s[3])), 2.0f);
        o.z = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv(s[4])), conv(s[5])), 2.0f);
        o.w = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv(s[6])), conv(s[7])), 2.0f);
    }
    return o;
}

// One aligned raw load per thread -> E mono samples in the reference's operation order.  The load is sized so
// that E <= 4: a thread then issues ONE 16-byte (or 8-byte) store per load and a warp's store instruction covers
// contiguous memory.  (A 16-byte load of mono s16 gives 8 samples = two float4 per thread; the warp's stores then
// hit 16-byte pieces 32 bytes apart, and that pattern tops out at ~4.2-4.6 TB/s against ~6.2 TB/s for contiguous
// stores — tools/micro/rw.cu, profiles/r01_hbm_mix_micro.txt.)
template <typename S, int CH> struct SlotRaw {
    static constexpr int kBytes = (CH * (int)sizeof(S) * 4 < 16) ? CH * (int)sizeof(S) * 4 : 16;
    using Raw = typename std::conditional<kBytes == 8, int2, int4>::type;
    static constexpr int E = kBytes / (CH * (int)sizeof(S));
    static __device__ __forceinline__ void convert(const Raw& raw, float (&o)[E]) {
        const S* s = reinterpret_cast<const S*>(&raw);
#pragma unroll
        for (int e = 0; e < E; ++e) {
            if (CH == 1) o[e] = conv(s[e]);
            else o[e] = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv(s[2 * e])), conv(s[2 * e + 1])), 2.0f);
        }
    }
};

// CH_T: 1 or 2 = compile-time channel count with vector loads; 0 = runtime channel count.
template <int FMT, int CH_T>
__global__ void __launch_bounds__(kThreads)
pack_kernel(const void* __restrict__ pcm_v, uint32_t channels, uint64_t total_frames,
            uint64_t seg, uint64_t hop, uint64_t nseg, uint64_t last_start, uint64_t row_first,
            float* __restrict__ out, uint32_t tiles_per_row, uint64_t ntiles) {
    using S = typename SampleT<FMT>::type;
    const S* __restrict__ pcm = static_cast<const S*>(pcm_v);
    const float fch = (float)channels;
    const bool row_vec = (seg % 4 == 0);           // rows are 16-byte aligned
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t lrow = tile / tiles_per_row;
        const uint64_t row = row_first + lrow;
        const uint64_t j0  = (tile - lrow * tiles_per_row) * (uint64_t)kTile;
        uint64_t start = 0, take = 0;
        if (row < nseg) {
            start = (row + 1 == nseg) ? last_start : row * hop;
            take  = total_frames - start < seg ? total_frames - start : seg;
        }
        float* __restrict__ orow = out + row * seg;
        if (CH_T != 0) {
            // fast path: the whole tile lies inside the window's real samples and the window start is 16-byte
            // aligned -> all loads of the tile are issued back to back (kUnroll x 16 B in flight per thread),
            // no per-element control flow between a load and the next one
            constexpr int ch = CH_T == 0 ? 1 : CH_T;
            using SL = SlotRaw<S, ch>;
            using Raw = typename SL::Raw;
            constexpr int E = SL::E;
            constexpr int UN = kUnroll * (16 / SL::kBytes);       // the same bytes in flight per thread for 8-byte loads
            constexpr int kPass = kThreads * UN * E;              // floats per pass
            static_assert(kTile % kPass == 0, "tile must be a whole number of passes");
            const bool aligned16 = ((reinterpret_cast<uintptr_t>(pcm) + start * ch * sizeof(S)) & 15) == 0 && (seg % 4) == 0 &&
                                   ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
            if (aligned16 && j0 + kTile <= take) {
#pragma unroll 1
                for (uint64_t jp = j0; jp < j0 + kTile; jp += kPass) {
                    Raw raw[UN];
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const uint64_t j = jp + ((uint64_t)u * kThreads + threadIdx.x) * E;
                        raw[u] = __ldg(reinterpret_cast<const Raw*>(pcm + (start + j) * ch));
                    }
#pragma unroll
                    for (int u = 0; u < UN; ++u) {
                        const uint64_t j = jp + ((uint64_t)u * kThreads + threadIdx.x) * E;
                        float o[E];
                        SL::convert(raw[u], o);
                        if (E == 4) *reinterpret_cast<float4*>(orow + j) = make_float4(o[0], o[1], o[2], o[3]);
                        else *reinterpret_cast<float2*>(orow + j) = make_float2(o[0], o[1]);
                    }
                }
                continue;
            }
        }
        for (uint64_t js = j0; js < j0 + kTile; js += kSub) {
        if (js >= seg) break;
        if (row_vec) {
            bool in_vec = false;
            if (CH_T != 0) {
                constexpr int ch = CH_T == 0 ? 1 : CH_T;
                constexpr size_t align = 4 * ch * sizeof(S) > 16 ? 16 : 4 * ch * sizeof(S);
                in_vec = ((reinterpret_cast<uintptr_t>(pcm) + start * ch * sizeof(S)) % align) == 0;
            }
            float4 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t j = js + ((uint64_t)u * kThreads + threadIdx.x) * 4;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < seg) {
                    if (CH_T != 0 && in_vec && j + 4 <= take) {
                        constexpr int ch = CH_T == 0 ? 1 : CH_T;
                        v[u] = load4_aligned<S, ch>(pcm + (start + j) * ch);
                    } else {
                        float t[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            t[e] = (j + e < take) ? downmix_frame(pcm + (start + j + e) * channels, channels, fch) : 0.0f;
                        v[u] = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const uint64_t j = js + ((uint64_t)u * kThreads + threadIdx.x) * 4;
                if (j < seg) __stcs(reinterpret_cast<float4*>(orow + j), v[u]);
            }
        } else {
            for (uint32_t e = threadIdx.x; e < (uint32_t)kSub; e += kThreads) {
                const uint64_t j = js + e;
                if (j < seg)
                    orow[j] = (j < take) ? downmix_frame(pcm + (start + j) * channels, channels, fch) : 0.0f;
            }
        }
        }
    }
}

// ---- 24-bit packed PCM (BB_S24).  symphonia's PCM decoder presents a 24-bit sample as S32 `sample << 8`, which
// append_samples converts through its S32 arm (decode.rs:386-402): value = ((s24 << 8) as f32) / 2^31.
// Four frames per thread: 12 * CH contiguous bytes, fetched as aligned 32-bit words (one extra word when the group
// starts mid-word, merged by a funnel shift); the window tail and the last bytes of the buffer go byte by byte.
__device__ __forceinline__ float conv24(int s_shl8) { return __fmul_rn(__int2float_rn(s_shl8), 1.0f / 2147483648.0f); }
__device__ __forceinline__ int load24_bytes(const unsigned char* __restrict__ p) {
    return (int)(((unsigned)__ldg(p) << 8) | ((unsigned)__ldg(p + 1) << 16) | ((unsigned)__ldg(p + 2) << 24));
}
__device__ __forceinline__ float downmix24(const unsigned char* __restrict__ p, uint32_t channels, float fch) {
    if (channels == 1) return conv24(load24_bytes(p));
    float sum = 0.0f;
    for (uint32_t c = 0; c < channels; ++c) sum = __fadd_rn(sum, conv24(load24_bytes(p + 3 * c)));
    return __fdiv_rn(sum, fch);
}
// 12 bytes (three little-endian words) -> four samples, each already shifted left by 8
__device__ __forceinline__ void unpack24x4(unsigned x0, unsigned x1, unsigned x2, int (&s)[4]) {
    s[0] = (int)(x0 << 8);
    s[1] = (int)(__funnelshift_r(x0, x1, 24) << 8);
    s[2] = (int)(__funnelshift_r(x1, x2, 16) << 8);
    s[3] = (int)(x2 & 0xffffff00u);
}

template <int CH_T>
__global__ void __launch_bounds__(kThreads)
pack24_kernel(const unsigned char* __restrict__ pcm, uint32_t channels, uint64_t total_frames,
              uint64_t seg, uint64_t hop, uint64_t nseg, uint64_t last_start, uint64_t row_first,
              float* __restrict__ out, uint32_t tiles_per_row, uint64_t ntiles) {
    const float fch = (float)channels;
    const unsigned char* pcm_end = pcm + total_frames * channels * 3;
    for (uint64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint64_t lrow = tile / tiles_per_row;
        const uint64_t row = row_first + lrow;
        const uint64_t j0 = (tile - lrow * tiles_per_row) * (uint64_t)kTile;
        uint64_t start = 0, take = 0;
        if (row < nseg) {
            start = (row + 1 == nseg) ? last_start : row * hop;
            take = total_frames - start < seg ? total_frames - start : seg;
        }
        float* __restrict__ orow = out + row * seg;
        const bool vec_store = (seg % 4 == 0) && ((reinterpret_cast<uintptr_t>(orow) & 15) == 0);
        for (uint64_t j = j0 + (uint64_t)threadIdx.x * 4; j < j0 + kTile && j < seg; j += (uint64_t)kThreads * 4) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            bool done = false;
            if (CH_T != 0 && j + 4 <= take) {
                constexpr int ch = CH_T == 0 ? 1 : CH_T;
                constexpr int NW = 3 * ch;                                     // words of payload
                const unsigned char* g = pcm + (start + j) * ch * 3;
                const uintptr_t ga = reinterpret_cast<uintptr_t>(g);
                const unsigned char* ga4 = reinterpret_cast<const unsigned char*>(ga & ~(uintptr_t)3);
                if (ga4 >= pcm && ga4 + (NW + 1) * 4 <= pcm_end) {
                    const unsigned sh = (unsigned)(ga & 3) * 8;
                    unsigned w[NW + 1];
#pragma unroll
                    for (int i = 0; i <= NW; ++i) w[i] = __ldg(reinterpret_cast<const unsigned*>(ga4) + i);
                    unsigned x[NW];
#pragma unroll
                    for (int i = 0; i < NW; ++i) x[i] = __funnelshift_r(w[i], w[i + 1], sh);
                    if (ch == 1) {
                        int sv[4]; unpack24x4(x[0], x[1], x[2], sv);
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = conv24(sv[e]);
                    } else {
                        int a[4], b[4];
                        unpack24x4(x[0], x[1], x[2], a); unpack24x4(x[NW - 3], x[NW - 2], x[NW - 1], b);
                        // frames (L0 R0) (L1 R1) | (L2 R2) (L3 R3): sum = 0 + l + r ; / 2 in the reference's order
                        o[0] = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv24(a[0])), conv24(a[1])), 2.0f);
                        o[1] = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv24(a[2])), conv24(a[3])), 2.0f);
                        o[2] = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv24(b[0])), conv24(b[1])), 2.0f);
                        o[3] = __fdiv_rn(__fadd_rn(__fadd_rn(0.0f, conv24(b[2])), conv24(b[3])), 2.0f);
                    }
                    done = true;
                }
            }
            if (!done) {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (j + e < take) o[e] = downmix24(pcm + (start + j + e) * channels * 3, channels, fch);
            }
            if (vec_store && j + 4 <= seg) *reinterpret_cast<float4*>(orow + j) = make_float4(o[0], o[1], o[2], o[3]);
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) if (j + e < seg) orow[j + e] = o[e];
            }
        }
    }
}

cudaError_t launch_s24(cudaStream_t st, const void* d_pcm, uint32_t channels, uint64_t total_frames, uint64_t seg,
                       uint64_t hop, uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, float* d_out) {
    const uint32_t tiles_per_row = (uint32_t)((seg + kTile - 1) / kTile);
    const uint64_t ntiles = rows_total * tiles_per_row;
    if (ntiles == 0) return cudaSuccess;
    const unsigned grid = (unsigned)(ntiles > 0x7fffffffull ? 0x7fffffffull : ntiles);
    const unsigned char* p = static_cast<const unsigned char*>(d_pcm);
    if (channels == 1) pack24_kernel<1><<<grid, kThreads, 0, st>>>(p, channels, total_frames, seg, hop, nseg, last_start, row_first, d_out, tiles_per_row, ntiles);
    else if (channels == 2) pack24_kernel<2><<<grid, kThreads, 0, st>>>(p, channels, total_frames, seg, hop, nseg, last_start, row_first, d_out, tiles_per_row, ntiles);
    else pack24_kernel<0><<<grid, kThreads, 0, st>>>(p, channels, total_frames, seg, hop, nseg, last_start, row_first, d_out, tiles_per_row, ntiles);
    return cudaGetLastError();
}

template <int FMT>
cudaError_t launch_fmt(cudaStream_t st, int sm_count, const void* d_pcm, uint32_t channels,
                       uint64_t total_frames, uint64_t seg, uint64_t hop, uint64_t nseg,
                       uint64_t last_start, uint64_t row_first, uint64_t rows_total, float* d_out) {
    const uint32_t tiles_per_row = (uint32_t)((seg + kTile - 1) / kTile);
    const uint64_t ntiles = rows_total * tiles_per_row;
    if (ntiles == 0) return cudaSuccess;
    // One CTA per 8192-sample tile, in tile order: neighbouring tiles (and the overlapped re-reads of the next window)
    // run at the same time, so DRAM sees long sequential runs and the re-reads hit L2.  A persistent grid-stride
    // loop was slower the fewer CTAs it had: 5/SM 0.443 ms, 8/SM 0.413, 16/SM 0.395, 64/SM 0.361, one per tile 0.342 ms
    // on C4 20 min = 6.59 TB/s (BIRDA_K1_CTAS_PER_SM restores a persistent grid for experiments).
    auto launch = [&](auto kern) -> cudaError_t {
        uint64_t want = ntiles;
        if (const char* g = std::getenv("BIRDA_K1_CTAS_PER_SM")) { const int v = atoi(g); if (v >= 1 && v <= 4096) want = (uint64_t)sm_count * v; }
        const unsigned grid = (unsigned)(ntiles < want ? ntiles : (want > 0x7fffffffull ? 0x7fffffffull : want));
        kern<<<grid, kThreads, 0, st>>>(d_pcm, channels, total_frames, seg, hop, nseg, last_start, row_first, d_out, tiles_per_row, ntiles);
        return cudaGetLastError();
    };
    if (channels == 1) return launch(pack_kernel<FMT, 1>);
    if (channels == 2) return launch(pack_kernel<FMT, 2>);
    return launch(pack_kernel<FMT, 0>);
}

}  // namespace

cudaError_t launch_pack(cudaStream_t st, int sm_count, const void* d_pcm, int fmt, uint32_t channels,
                        uint64_t total_frames, uint64_t seg, uint64_t hop, uint64_t nseg,
                        uint64_t last_start, uint64_t row_first, uint64_t rows_total, float* d_out) {
    switch (fmt) {
        case BB_S16: return launch_fmt<BB_S16>(st, sm_count, d_pcm, channels, total_frames, seg, hop, nseg, last_start, row_first, rows_total, d_out);
        case BB_S32: return launch_fmt<BB_S32>(st, sm_count, d_pcm, channels, total_frames, seg, hop, nseg, last_start, row_first, rows_total, d_out);
        case BB_F32: return launch_fmt<BB_F32>(st, sm_count, d_pcm, channels, total_frames, seg, hop, nseg, last_start, row_first, rows_total, d_out);
        case BB_S24: return launch_s24(st, d_pcm, channels, total_frames, seg, hop, nseg, last_start, row_first, rows_total, d_out);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace bb
