// K2 — per-window FFT resampler fused with sample conversion / downmix / window gather in
// front and the model-tensor pack behind.
//
// Replaces, per window: resample_chunk (src/audio/resample.rs:97-105) over a NEW
// rubato::Fft<f32> (src/audio/resample.rs:19-25; overlap-add of zero-padded 2*N_in real FFTs,
// spectrum * filter spectrum, re-binned, 2*N_out inverse real FFT), `samples.resize(seg, 0.0)`
// (src/pipeline/processor.rs:87) and the loads K1 does (decode.rs:353-411, :175-181).  Every
// window starts from a zero carry and its last carry is dropped, as the reference builds a
// fresh resampler per segment.
//
// Layout: one CTA owns a run of consecutive blocks of one window (carry stays in shared
// memory; the block before the run is recomputed instead of exchanged), G blocks in flight per
// round.  Real transforms run as half-length complex Stockham FFTs in shared memory (radices
// 2,3,4,5,7,8 and odd primes <= 31), twiddles from f64-built tables.  The split / filter /
// re-bin / inverse-pack steps between the two transforms are fused into one pass.
//
// Bound: FP32 + shared-memory bandwidth, not HBM; algorithmic HBM
// bytes are the same as K1's.
#include "common.cuh"
#include "fft_butterflies.cuh"
#include <cmath>
#include <vector>

namespace bb {
namespace {

constexpr int kThreads   = 256;
constexpr int kMaxStages = 16;

struct FastDiv {          // exact for numerators and divisors < 65536
    uint32_t d, m;
    __device__ __forceinline__ uint32_t div(uint32_t t) const { return d == 1 ? t : __umulhi(t, m); }
};
static FastDiv make_fastdiv(uint32_t d) {
    FastDiv f; f.d = d; f.m = d <= 1 ? 0u : (uint32_t)(((1ull << 32) + d - 1) / d);
    return f;
}

struct Stage { int radix; uint32_t nb; uint32_t s; FastDiv div_nb, div_s; };

struct K2Params {
    const void* pcm; int fmt; uint32_t channels; uint64_t total_frames;
    uint64_t src_seg, hop, nseg, last_start, rows_total, row_first, seg;
    uint32_t out_len;        // min(resampled_len, seg): samples of each row that come from the resampler
    float* out;
    uint32_t n_in, n_out, n_keep;
    int nst_f, nst_i;
    Stage st_f[kMaxStages], st_i[kMaxStages];
    const float2 *tw_f, *tw_i, *split_f, *split_i, *filt;
    uint32_t buf_len, G, nblk, R, items_per_row;
    uint64_t nitems;
    unsigned char* ws; size_t ws_stride;       // global workspace mode (WS kernels): per-CTA slice in place of shared memory
};

// One Stockham DIF stage over the G blocks of a round:  y[q + s*(r*p + k)] = DFT_r(x[i + j*nb])_k * w^(s*p*k)
template <int R, bool INV>
__device__ __forceinline__ void run_stage(const Stage& st, const float2* __restrict__ src, float2* __restrict__ dst,
                                          const float2* __restrict__ tw, uint32_t buf_stride, uint32_t G, bool last) {
    const uint32_t nb = st.nb, s = st.s, total = G * nb;
    for (uint32_t t = threadIdx.x; t < total; t += kThreads) {
        const uint32_t g = st.div_nb.div(t);
        const uint32_t i = t - g * nb;
        const float2* __restrict__ x = src + (size_t)g * buf_stride;
        float2* __restrict__ y = dst + (size_t)g * buf_stride;
        float2 a[R];
#pragma unroll
        for (int j = 0; j < R; ++j) a[j] = x[i + j * nb];
        Dft<R, INV>::run(a);
        const uint32_t q = i - st.div_s.div(i) * s;
        const uint32_t sp = i - q;                 // s * p
        const uint32_t ob = i + sp * (R - 1);
        y[ob] = a[0];
        if (last || sp == 0) {
#pragma unroll
            for (int k = 1; k < R; ++k) y[ob + s * k] = a[k];
        } else {
#pragma unroll
            for (int k = 1; k < R; ++k) y[ob + s * k] = cmul(a[k], tw[sp * k]);
        }
    }
}

// The same stage for any prime radix (block lengths with a prime factor > 31): every output of a butterfly is a direct
// sum over its R inputs, W_R^(jk) read from the transform's full-length table (tw[m] = exp(-+2 pi i m / n), so
// W_R^e = tw[e * n / R]).  One thread per (butterfly, k); O(n * R) per stage.
__device__ __forceinline__ void run_stage_generic(const Stage& st, const float2* __restrict__ src, float2* __restrict__ dst,
                                                  const float2* __restrict__ tw, uint32_t n, uint32_t buf_stride, uint32_t G, bool last) {
    const uint32_t R = (uint32_t)st.radix, nb = st.nb, s = st.s, step = n / R;
    const uint32_t per_block = nb * R;                      // == n
    for (uint32_t t = threadIdx.x; t < G * per_block; t += kThreads) {
        const uint32_t g = t / per_block, r = t - g * per_block;
        const uint32_t k = r / nb, i = r - k * nb;
        const float2* __restrict__ x = src + (size_t)g * buf_stride;
        float2 acc = x[i];
        uint32_t e = 0;                                     // (j * k) mod R
        for (uint32_t j = 1; j < R; ++j) {
            e += k; if (e >= R) e -= R;
            acc = cadd(acc, cmul(x[i + j * nb], tw[e * step]));
        }
        const uint32_t q = i - st.div_s.div(i) * s;
        const uint32_t sp = i - q;
        const uint32_t ob = i + sp * (R - 1);
        if (!(last || sp == 0 || k == 0)) acc = cmul(acc, tw[sp * k]);
        dst[(size_t)g * buf_stride + ob + s * k] = acc;
    }
}

template <bool INV>
__device__ __forceinline__ void run_stage_dispatch(const Stage& st, const float2* src, float2* dst, const float2* tw, uint32_t n,
                                                   uint32_t buf_stride, uint32_t G, bool last) {
    switch (st.radix) {
        case 2:  run_stage<2, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 3:  run_stage<3, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 4:  run_stage<4, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 5:  run_stage<5, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 7:  run_stage<7, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 8:  run_stage<8, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 11: run_stage<11, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 13: run_stage<13, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 17: run_stage<17, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 19: run_stage<19, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 23: run_stage<23, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 29: run_stage<29, INV>(st, src, dst, tw, buf_stride, G, last); break;
        case 31: run_stage<31, INV>(st, src, dst, tw, buf_stride, G, last); break;
        default: run_stage_generic(st, src, dst, tw, n, buf_stride, G, last); break;      // the table carries the direction
    }
}

__device__ __forceinline__ float conv_s(int16_t s) { return __fmul_rn(__int2float_rn((int)s), 1.0f / 32768.0f); }
__device__ __forceinline__ float conv_s(int32_t s) { return __fmul_rn(__int2float_rn(s), 1.0f / 2147483648.0f); }
__device__ __forceinline__ float conv_s(float s)   { return s; }

template <typename S>
__device__ __forceinline__ float mono_at_t(const void* __restrict__ pcm_v, uint64_t frame, uint32_t channels, float fch) {
    const S* p = static_cast<const S*>(pcm_v) + frame * channels;
    if (channels == 1) return conv_s(__ldg(p));
    float sum = 0.0f;
    for (uint32_t c = 0; c < channels; ++c) sum = __fadd_rn(sum, conv_s(__ldg(p + c)));
    return __fdiv_rn(sum, fch);
}
// sample conversion + downmix of one frame (decode.rs:353-411); fmt is CTA-uniform
__device__ __forceinline__ float mono_at(const void* __restrict__ pcm, int fmt, uint64_t frame, uint32_t channels, float fch) {
    if (fmt == BB_S16) return mono_at_t<int16_t>(pcm, frame, channels, fch);
    if (fmt == BB_S32) return mono_at_t<int32_t>(pcm, frame, channels, fch);
    if (fmt == BB_S24) {                       // packed 24-bit: (s24 << 8) through the S32 arm (decode.rs:386-402)
        const unsigned char* p = static_cast<const unsigned char*>(pcm) + frame * channels * 3;
        auto ld24 = [](const unsigned char* q) {
            return (int)(((unsigned)__ldg(q) << 8) | ((unsigned)__ldg(q + 1) << 16) | ((unsigned)__ldg(q + 2) << 24));
        };
        if (channels == 1) return conv_s(ld24(p));
        float sum = 0.0f;
        for (uint32_t c = 0; c < channels; ++c) sum = __fadd_rn(sum, conv_s(ld24(p + 3 * c)));
        return __fdiv_rn(sum, fch);
    }
    return mono_at_t<float>(pcm, frame, channels, fch);
}

// spectrum bin k of the zero-padded 2*N real block from the half-length transform Z (k in [0, N])
__device__ __forceinline__ float2 real_bin(const float2* __restrict__ Z, uint32_t k, uint32_t N, float2 wf) {
    const float2 zk = Z[k == N ? 0 : k];
    const float2 zn = cconj(Z[k == 0 ? 0 : N - k]);
    const float2 fe = make_float2(0.5f * (zk.x + zn.x), 0.5f * (zk.y + zn.y));
    const float2 d  = make_float2(0.5f * (zk.x - zn.x), 0.5f * (zk.y - zn.y));
    const float2 fo = make_float2(d.y, -d.x);                 // -i * d
    return cadd(fe, cmul(wf, fo));
}

template <bool WS>
__global__ void __launch_bounds__(kThreads)
resample_kernel(const K2Params P) {
    extern __shared__ __align__(16) unsigned char smem_dyn[];
    unsigned char* smem_raw = WS ? P.ws + (size_t)blockIdx.x * P.ws_stride : smem_dyn;
    float2* tw_f = reinterpret_cast<float2*>(smem_raw);
    float2* tw_i = tw_f + P.n_in;
    float*  carry = reinterpret_cast<float*>(tw_i + P.n_out);
    float2* bufA = reinterpret_cast<float2*>(carry + ((P.n_out + 3) & ~3u));
    float2* bufB = bufA + (size_t)P.G * P.buf_len;

    const void* __restrict__ pcm = P.pcm;
    const float fch = (float)P.channels;
    const uint32_t N = P.n_in, M = P.n_out, G = P.G, BL = P.buf_len;

    for (uint32_t i = threadIdx.x; i < N; i += kThreads) tw_f[i] = __ldg(P.tw_f + i);
    for (uint32_t i = threadIdx.x; i < M; i += kThreads) tw_i[i] = __ldg(P.tw_i + i);

    for (uint64_t item = blockIdx.x; item < P.nitems; item += gridDim.x) {
        const uint64_t lrow = item / P.items_per_row;
        const uint64_t row = P.row_first + lrow;
        const uint32_t it  = (uint32_t)(item - lrow * P.items_per_row);
        float* __restrict__ orow = P.out + row * P.seg;
        const uint32_t b0 = it * P.R;
        const uint32_t b1 = min(b0 + P.R, P.nblk);
        // this item's share of the output row: [o_lo, o_hi) comes from the resampler, the last
        // item of a row also zero-fills [out_len, seg)
        const uint32_t o_lo = min(b0 * M, P.out_len);
        const uint32_t o_hi = (it + 1 == P.items_per_row) ? P.out_len : min(b1 * M, P.out_len);
        if (row >= P.nseg) {                     // batch-padding row: all zeros (processor.rs:239-260)
            const uint64_t z_hi = (it + 1 == P.items_per_row) ? P.seg : o_hi;
            for (uint64_t j = o_lo + threadIdx.x; j < z_hi; j += kThreads) orow[j] = 0.0f;
            continue;
        }
        if (it + 1 == P.items_per_row)
            for (uint64_t j = P.out_len + threadIdx.x; j < P.seg; j += kThreads) orow[j] = 0.0f;
        if (b0 >= b1) continue;

        const uint64_t start = (row + 1 == P.nseg) ? P.last_start : row * P.hop;
        const uint64_t take  = P.total_frames - start < P.src_seg ? P.total_frames - start : P.src_seg;

        __syncthreads();
        for (uint32_t j = threadIdx.x; j < M; j += kThreads) carry[j] = 0.0f;
        const uint32_t bfirst = b0 > 0 ? b0 - 1 : 0;     // recompute the block before the run for its carry

        for (uint32_t br = bfirst; br < b1; br += G) {
            const uint32_t g_n = min(G, b1 - br);
            __syncthreads();
            // ---- load: z[n] = x[2n] + i x[2n+1], zero beyond the block / the window's real samples
            for (uint32_t t = threadIdx.x; t < g_n * N; t += kThreads) {
                const uint32_t g = t / N, n = t - g * N;
                const uint64_t q0 = (uint64_t)(br + g) * N + 2 * n;      // sample index inside the window
                float re = 0.f, im = 0.f;
                if (2 * n < N && q0 < take) re = mono_at(pcm, P.fmt, start + q0, P.channels, fch);
                if (2 * n + 1 < N && q0 + 1 < take) im = mono_at(pcm, P.fmt, start + q0 + 1, P.channels, fch);
                bufA[(size_t)g * BL + n] = make_float2(re, im);
            }
            __syncthreads();
            // ---- forward half-length complex FFT
            float2* src = bufA; float2* dst = bufB;
            for (int s = 0; s < P.nst_f; ++s) {
                run_stage_dispatch<false>(P.st_f[s], src, dst, tw_f, N, BL, g_n, s + 1 == P.nst_f);
                __syncthreads();
                float2* t = src; src = dst; dst = t;
            }
            // ---- split -> X[k], * filter, re-bin to M+1 bins, pack for the inverse half-length FFT
            const uint32_t half = M / 2;
            for (uint32_t t = threadIdx.x; t < g_n * (half + 1); t += kThreads) {
                const uint32_t g = t / (half + 1), k = t - g * (half + 1);
                const uint32_t k2 = M - k;                                 // partner bin, in [M/2, M]
                const float2* __restrict__ Z = src + (size_t)g * BL;
                float2 yk = make_float2(0.f, 0.f), yk2 = make_float2(0.f, 0.f);
                if (k < P.n_keep)  yk  = cmul(real_bin(Z, k, N, __ldg(P.split_f + k)),  __ldg(P.filt + k));
                if (k2 < P.n_keep) yk2 = cmul(real_bin(Z, k2, N, __ldg(P.split_f + k2)), __ldg(P.filt + k2));
                if (k == 0) { yk.y = 0.f; yk2.y = 0.f; }                   // DC / Nyquist are real (realfft ignores imag)
                const float2 wi = __ldg(P.split_i + k);                    // exp(+i pi k / M)
                float2* __restrict__ Zp = dst + (size_t)g * BL;
                {   // Z'(k) = (Y(k) + conj Y(M-k)) + i wi (Y(k) - conj Y(M-k))
                    const float2 e = cadd(yk, cconj(yk2));
                    const float2 o = cmul(wi, csub(yk, cconj(yk2)));
                    Zp[k] = make_float2(e.x - o.y, e.y + o.x);
                }
                if (k != 0 && k2 != k) {   // Z'(M-k), with exp(i pi (M-k)/M) = -conj(wi)
                    const float2 wi2 = make_float2(-wi.x, wi.y);
                    const float2 e = cadd(yk2, cconj(yk));
                    const float2 o = cmul(wi2, csub(yk2, cconj(yk)));
                    Zp[k2] = make_float2(e.x - o.y, e.y + o.x);
                }
            }
            __syncthreads();
            { float2* t = src; src = dst; dst = t; }
            // ---- inverse half-length complex FFT (unnormalised)
            for (int s = 0; s < P.nst_i; ++s) {
                run_stage_dispatch<true>(P.st_i[s], src, dst, tw_i, M, BL, g_n, s + 1 == P.nst_i);
                __syncthreads();
                float2* t = src; src = dst; dst = t;
            }
            // ---- overlap-add and store: y_g[j] + y_{g-1}[M + j]  (carry for g == 0)
            const float* __restrict__ yb = reinterpret_cast<const float*>(src);
            for (uint32_t t = threadIdx.x; t < g_n * M; t += kThreads) {
                const uint32_t g = t / M, j = t - g * M;
                const uint32_t b = br + g;
                if (b < b0) continue;                                      // recomputed block: carry only
                const uint32_t o = b * M + j;
                if (o >= o_hi) continue;
                const float prev = g > 0 ? yb[(size_t)(g - 1) * 2 * BL + M + j] : carry[j];
                orow[o] = yb[(size_t)g * 2 * BL + j] + prev;
            }
            __syncthreads();
            for (uint32_t j = threadIdx.x; j < M; j += kThreads) carry[j] = yb[(size_t)(g_n - 1) * 2 * BL + M + j];
        }
    }
}

size_t smem_bytes(const K2Params& P) {
    return (size_t)(P.n_in + P.n_out) * sizeof(float2) + (size_t)((P.n_out + 3) & ~3u) * sizeof(float) +
           (size_t)2 * P.G * P.buf_len * sizeof(float2);
}

}  // namespace

cudaError_t resampler_dev_init(const ResamplerSpec& spec, ResamplerDev* rs) {
    rs->n_in = spec.n_in; rs->n_out = spec.n_out; rs->n_keep = spec.n_keep;
    if (spec.radix_fwd.size() > kMaxStages || spec.radix_inv.size() > kMaxStages) return cudaErrorInvalidConfiguration;
    if (spec.n_in >= 65536 || spec.n_out >= 65536) return cudaErrorInvalidConfiguration;
    rs->nstage_fwd = (int)spec.radix_fwd.size(); rs->nstage_inv = (int)spec.radix_inv.size();
    for (int i = 0; i < rs->nstage_fwd; ++i) rs->radix_fwd[i] = spec.radix_fwd[i];
    for (int i = 0; i < rs->nstage_inv; ++i) rs->radix_inv[i] = spec.radix_inv[i];
    rs->buf_len = (spec.n_in > spec.n_out ? spec.n_in : spec.n_out) + 1;
    if (rs->buf_len & 1) rs->buf_len += 1;        // keep every buffer 16-byte aligned
    // one block in flight must fit
    const size_t min_smem = (size_t)(spec.n_in + spec.n_out) * 8 + (size_t)(spec.n_out + 4) * 4 + (size_t)2 * rs->buf_len * 8;
    if (min_smem > 220 * 1024) {               // global workspace mode
        rs->ws_stride = (min_smem + 255) & ~(size_t)255;
        rs->ws_ctas = 304;
        cudaError_t we = cudaMalloc(&rs->d_ws, rs->ws_stride * rs->ws_ctas);
        if (we != cudaSuccess) return we;
    }
    const double pi = 3.14159265358979323846;
    auto upload = [](const std::vector<float2>& h, float2** d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, h.size() * sizeof(float2));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
    };
    std::vector<float2> h;
    cudaError_t e;
    h.resize(spec.n_in);
    for (uint32_t k = 0; k < spec.n_in; ++k) { double a = -2.0 * pi * k / spec.n_in; h[k] = make_float2((float)cos(a), (float)sin(a)); }
    if ((e = upload(h, &rs->d_tw_fwd)) != cudaSuccess) return e;
    h.resize(spec.n_out);
    for (uint32_t k = 0; k < spec.n_out; ++k) { double a = 2.0 * pi * k / spec.n_out; h[k] = make_float2((float)cos(a), (float)sin(a)); }
    if ((e = upload(h, &rs->d_tw_inv)) != cudaSuccess) return e;
    h.resize(spec.n_in + 1);
    for (uint32_t k = 0; k <= spec.n_in; ++k) { double a = -pi * k / spec.n_in; h[k] = make_float2((float)cos(a), (float)sin(a)); }
    if ((e = upload(h, &rs->d_split_fwd)) != cudaSuccess) return e;
    h.resize(spec.n_out + 1);
    for (uint32_t k = 0; k <= spec.n_out; ++k) { double a = pi * k / spec.n_out; h[k] = make_float2((float)cos(a), (float)sin(a)); }
    if ((e = upload(h, &rs->d_split_inv)) != cudaSuccess) return e;
    h.resize(spec.n_keep);
    for (uint32_t k = 0; k < spec.n_keep; ++k) h[k] = make_float2(spec.filt_re[k], spec.filt_im[k]);
    if ((e = upload(h, &rs->d_filt)) != cudaSuccess) return e;
    if (warp_plan_available(spec)) return warp_tables_init(spec, rs);
    return cudaSuccess;
}

void resampler_dev_free(ResamplerDev* rs) {
    if (rs->d_tw_fwd) cudaFree(rs->d_tw_fwd);
    if (rs->d_tw_inv) cudaFree(rs->d_tw_inv);
    if (rs->d_split_fwd) cudaFree(rs->d_split_fwd);
    if (rs->d_split_inv) cudaFree(rs->d_split_inv);
    if (rs->d_filt) cudaFree(rs->d_filt);
    if (rs->d_ws) cudaFree(rs->d_ws);
    rs->d_ws = nullptr; rs->ws_stride = 0; rs->ws_ctas = 0;
    warp_tables_free(rs);
    *rs = ResamplerDev();
}

cudaError_t launch_resample(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                            uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                            uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                            uint64_t resampled_len, float* d_out, int* launches) {
    if (rs.fast)
        return launch_resample_warp(st, sm_count, rs, d_pcm, fmt, channels, total_frames, src_seg, hop, nseg, last_start,
                                    row_first, rows_total, seg, resampled_len, d_out, launches);
    if (launches) *launches = 0;
    if (rows_total == 0) return cudaSuccess;
    K2Params P{};
    P.pcm = d_pcm; P.fmt = fmt; P.channels = channels; P.total_frames = total_frames;
    P.src_seg = src_seg; P.hop = hop; P.nseg = nseg; P.last_start = last_start; P.rows_total = rows_total; P.row_first = row_first; P.seg = seg;
    P.out_len = (uint32_t)(resampled_len < seg ? resampled_len : seg);
    P.out = d_out;
    P.n_in = rs.n_in; P.n_out = rs.n_out; P.n_keep = rs.n_keep;
    P.nst_f = rs.nstage_fwd; P.nst_i = rs.nstage_inv;
    auto fill = [](Stage* stg, const int* radix, int n, uint32_t N) {
        uint32_t s = 1;
        for (int i = 0; i < n; ++i) {
            stg[i].radix = radix[i]; stg[i].nb = N / radix[i]; stg[i].s = s;
            stg[i].div_nb = make_fastdiv(stg[i].nb); stg[i].div_s = make_fastdiv(s);
            s *= radix[i];
        }
    };
    fill(P.st_f, rs.radix_fwd, rs.nstage_fwd, rs.n_in);
    fill(P.st_i, rs.radix_inv, rs.nstage_inv, rs.n_out);
    P.tw_f = rs.d_tw_fwd; P.tw_i = rs.d_tw_inv; P.split_f = rs.d_split_fwd; P.split_i = rs.d_split_inv; P.filt = rs.d_filt;
    P.buf_len = rs.buf_len;
    // blocks that contribute output samples
    P.nblk = (P.out_len + rs.n_out - 1) / rs.n_out;
    if (P.nblk == 0) P.nblk = 1;
    // G: blocks in flight per round — as many as fit in ~108 KB (two CTAs per SM), at least 1
    const size_t fixed = (size_t)(rs.n_in + rs.n_out) * 8 + (size_t)((rs.n_out + 3) & ~3u) * 4;
    const size_t per_g = (size_t)2 * rs.buf_len * 8;
    size_t budget = 108 * 1024;
    uint32_t G = fixed + per_g <= budget ? (uint32_t)((budget - fixed) / per_g) : 1;
    if (G > 8) G = 8;
    if (G < 1 || rs.d_ws) G = 1;
    P.G = G;
    P.ws = rs.d_ws; P.ws_stride = rs.ws_stride;
    // R: blocks per work item.  Whole windows when there are plenty of them, shorter runs otherwise
    const uint64_t ctas = (uint64_t)sm_count * 2;
    uint32_t R = P.nblk;
    if (rows_total < ctas * 4) {
        uint64_t want_items = ctas * 4;
        uint64_t per_row = (want_items + rows_total - 1) / rows_total;
        R = (uint32_t)((P.nblk + per_row - 1) / per_row);
        const uint32_t minR = 4 * G > 16 ? 4 * G : 16;     // keep the recomputed block a small fraction
        if (R < minR) R = minR;
        if (R > P.nblk) R = P.nblk;
    }
    P.R = R;
    P.items_per_row = (P.nblk + R - 1) / R;
    P.nitems = rows_total * P.items_per_row;
    const size_t smem = smem_bytes(P);
    uint64_t grid64 = P.nitems < ctas * 8 ? P.nitems : ctas * 8;
    // persistent-ish: no more CTAs than items, a multiple of the SM count when there are many
    if (grid64 > ctas) grid64 = ctas * ((grid64 + ctas - 1) / ctas);
    if (grid64 > P.nitems) grid64 = P.nitems;
    const unsigned grid = (unsigned)grid64;
    cudaError_t e = cudaSuccess;
    if (fmt != BB_S16 && fmt != BB_S32 && fmt != BB_F32 && fmt != BB_S24) return cudaErrorInvalidValue;
    if (rs.d_ws) {
        if (smem > rs.ws_stride) return cudaErrorInvalidConfiguration;
        resample_kernel<true><<<grid < rs.ws_ctas ? grid : rs.ws_ctas, kThreads, 0, st>>>(P);
        if (launches) *launches = 1;
        return cudaGetLastError();
    }
    e = cudaFuncSetAttribute(resample_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    resample_kernel<false><<<grid, kThreads, smem, st>>>(P);
    if (launches) *launches = 1;
    return cudaGetLastError();
}

}  // namespace bb
