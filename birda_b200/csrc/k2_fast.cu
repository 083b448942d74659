// K2 fast path kernels (see k2_fast.cuh for the design) and their launcher.
#include "common.cuh"
#include "k2_plans.cuh"
#include <cstdlib>
#include <vector>

namespace bb {
namespace {

using namespace bb::k2f;

constexpr size_t kSmemMax = 227 * 1024;

template <class PL> struct DevLayout {
    static constexpr size_t a16(size_t x) { return (x + 15) & ~(size_t)15; }
    static constexpr size_t off_twf = 0;
    static constexpr size_t off_twi = off_twf + a16(sizeof(float2) * PL::TWF);
    static constexpr size_t off_posf = off_twi + a16(sizeof(float2) * PL::TWI);
    static constexpr size_t off_posi = off_posf + a16(sizeof(uint16_t) * PL::N);
    static constexpr size_t off_P = off_posi + a16(sizeof(uint16_t) * PL::M);
    static constexpr size_t off_Q = off_P + a16(sizeof(float2) * PL::NKEEP);
    static constexpr size_t off_WI = off_Q + a16(sizeof(float2) * PL::NKEEP);
    static constexpr size_t tables = off_WI + a16(sizeof(float2) * (PL::M / 2 + 1));
    static constexpr size_t per_warp = a16(sizeof(float2) * PL::N) + a16(sizeof(float2) * PL::M);
    static constexpr int warps_fit = (int)((kSmemMax - tables) / per_warp);
    static constexpr int WARPS = warps_fit > 16 ? 16 : (warps_fit < 1 ? 1 : warps_fit);
    static constexpr size_t smem = tables + (size_t)WARPS * per_warp;
};

struct FastParams {
    const void* pcm; int fmt; uint32_t channels; uint64_t total_frames;
    uint64_t src_seg, hop, nseg, last_start, rows_total, seg;
    uint32_t out_len; float* out;
    const float2 *twf, *twi, *Pt, *Qt, *WI; const uint16_t *pos_f, *pos_i;
    uint32_t nblk, R, items_per_row; uint64_t nitems;
    unsigned long long* counter;       // dynamic work distribution
};

struct DevExec {
    static constexpr int kSlots = 1;
    template <class F> static BB_HD void each(F&& f) {
#ifdef __CUDA_ARCH__
        f((int)(threadIdx.x & 31), 0);
        __syncwarp();
#endif
    }
};

// sample conversion + downmix of one frame, bit-identical to decode.rs:353-411.  For S16 the
// integer sum is exact (|sum| < 2^24), so (sum * 2^-15) / C equals the reference's f32 sequence.
struct DevLoader {
    const void* pcm; int kind; uint32_t ch; uint64_t base; int valid; float fch;
    __device__ __forceinline__ float mono(uint64_t f) const {
        if (kind == 0) {            // S16 stereo, 4-byte aligned frames
            const int v = __ldg(reinterpret_cast<const int*>(pcm) + f);
            return __fmul_rn(__int2float_rn((int)(short)(v & 0xffff) + (v >> 16)), 1.0f / 65536.0f);
        } else if (kind == 1) {     // S16 mono
            return __fmul_rn(__int2float_rn((int)__ldg(reinterpret_cast<const short*>(pcm) + f)), 1.0f / 32768.0f);
        } else if (kind == 2) {     // S16, any channel count
            const short* p = reinterpret_cast<const short*>(pcm) + f * ch;
            int s = 0;
            for (uint32_t c = 0; c < ch; ++c) s += (int)__ldg(p + c);
            return __fdiv_rn(__fmul_rn(__int2float_rn(s), 1.0f / 32768.0f), fch);
        } else if (kind == 3) {     // S32
            const int* p = reinterpret_cast<const int*>(pcm) + f * ch;
            if (ch == 1) return __fmul_rn(__int2float_rn(__ldg(p)), 1.0f / 2147483648.0f);
            float s = 0.0f;
            for (uint32_t c = 0; c < ch; ++c) s = __fadd_rn(s, __fmul_rn(__int2float_rn(__ldg(p + c)), 1.0f / 2147483648.0f));
            return __fdiv_rn(s, fch);
        } else {                    // F32
            const float* p = reinterpret_cast<const float*>(pcm) + f * ch;
            if (ch == 1) return __ldg(p);
            float s = 0.0f;
            for (uint32_t c = 0; c < ch; ++c) s = __fadd_rn(s, __ldg(p + c));
            return __fdiv_rn(s, fch);
        }
    }
    __device__ __forceinline__ float2 operator()(int n) const {
        const int i0 = 2 * n;
        float re = 0.f, im = 0.f;
        if (i0 < valid) re = mono(base + i0);
        if (i0 + 1 < valid) im = mono(base + i0 + 1);
        return make_float2(re, im);
    }
};

struct DevSink {
    float* p; int lim; bool vec;
    __device__ __forceinline__ void operator()(int n, float2 y) const {
        const int o = 2 * n;
        if (vec && o + 1 < lim) *reinterpret_cast<float2*>(p + o) = y;
        else { if (o < lim) p[o] = y.x; if (o + 1 < lim) p[o + 1] = y.y; }
    }
};
struct DevSinkFactory {
    DevSink s;
    __device__ __forceinline__ DevSink operator()(int) const { return s; }
};

template <class PL>
__global__ void __launch_bounds__(DevLayout<PL>::WARPS * 32, 1)
resample_fast_kernel(const FastParams P) {
    using L = DevLayout<PL>;
    extern __shared__ __align__(16) unsigned char smem[];
    float2* s_twf = reinterpret_cast<float2*>(smem + L::off_twf);
    float2* s_twi = reinterpret_cast<float2*>(smem + L::off_twi);
    uint16_t* s_posf = reinterpret_cast<uint16_t*>(smem + L::off_posf);
    uint16_t* s_posi = reinterpret_cast<uint16_t*>(smem + L::off_posi);
    float2* s_P = reinterpret_cast<float2*>(smem + L::off_P);
    float2* s_Q = reinterpret_cast<float2*>(smem + L::off_Q);
    float2* s_WI = reinterpret_cast<float2*>(smem + L::off_WI);
    constexpr int NT = L::WARPS * 32;
    for (int i = threadIdx.x; i < PL::TWF; i += NT) s_twf[i] = P.twf[i];
    for (int i = threadIdx.x; i < PL::TWI; i += NT) s_twi[i] = P.twi[i];
    for (int i = threadIdx.x; i < PL::N; i += NT) s_posf[i] = P.pos_f[i];
    for (int i = threadIdx.x; i < PL::M; i += NT) s_posi[i] = P.pos_i[i];
    for (int i = threadIdx.x; i < PL::NKEEP; i += NT) { s_P[i] = P.Pt[i]; s_Q[i] = P.Qt[i]; }
    for (int i = threadIdx.x; i <= PL::M / 2; i += NT) s_WI[i] = P.WI[i];
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2* A = reinterpret_cast<float2*>(smem + L::tables + (size_t)warp * L::per_warp);
    float2* B = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(A) + L::a16(sizeof(float2) * PL::N));
    const Tables<PL> T{s_twf, s_twi, s_posf, s_posi, s_P, s_Q, s_WI};
    constexpr int N = PL::N, M = PL::M;

    int kind;
    if (P.fmt == BB_S16) kind = P.channels == 2 && (reinterpret_cast<uintptr_t>(P.pcm) & 3) == 0 ? 0 : (P.channels == 1 ? 1 : 2);
    else kind = P.fmt == BB_S32 ? 3 : 4;

    for (;;) {
        unsigned long long item = 0;
        if (lane == 0) item = atomicAdd(P.counter, 1ull);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= P.nitems) break;
        const uint64_t row = item / P.items_per_row;
        const uint32_t it = (uint32_t)(item - row * P.items_per_row);
        float* __restrict__ orow = P.out + row * P.seg;
        const uint32_t b0 = it * P.R;
        const uint32_t b1 = min(b0 + P.R, P.nblk);
        const bool last_item = it + 1 == P.items_per_row;
        const uint32_t o_lo = min(b0 * (uint32_t)M, P.out_len);
        const uint32_t o_hi = last_item ? P.out_len : min(b1 * (uint32_t)M, P.out_len);
        if (row >= P.nseg) {                         // batch-padding row: zeros (processor.rs:239-260)
            const uint64_t z_hi = last_item ? P.seg : o_hi;
            for (uint64_t j = o_lo + lane; j < z_hi; j += 32) orow[j] = 0.0f;
            continue;
        }
        if (last_item) for (uint64_t j = P.out_len + lane; j < P.seg; j += 32) orow[j] = 0.0f;
        if (b0 >= b1) continue;
        const uint64_t start = (row + 1 == P.nseg) ? P.last_start : row * P.hop;
        const uint64_t take = P.total_frames - start < P.src_seg ? P.total_frames - start : P.src_seg;

        LaneCarry<PL> carry;
#pragma unroll
        for (int i = 0; i < PL::CARRY_ITERS; ++i)
#pragma unroll
            for (int j = 0; j < PL::QL; ++j) carry.c[i][j] = 0.f;

        const bool vec = ((reinterpret_cast<uintptr_t>(orow) & 7) == 0);      // M is even: b*M keeps 8-byte alignment
        for (uint32_t b = b0 > 0 ? b0 - 1 : 0; b < b1; ++b) {
            const uint64_t q0 = (uint64_t)b * N;
            int valid = 0;
            if (q0 < take) valid = (int)(take - q0 < (uint64_t)N ? take - q0 : (uint64_t)N);
            DevLoader ld{P.pcm, kind, P.channels, start + q0, valid, (float)P.channels};
            DevSinkFactory sf;
            const int64_t lim = (int64_t)o_hi - (int64_t)b * M;
            sf.s.p = orow + (size_t)b * M;
            sf.s.lim = b < b0 ? 0 : (int)(lim < 0 ? 0 : (lim > M ? M : lim));  // recomputed block: carry only
            sf.s.vec = vec;
            process_block<PL, DevExec>(T, A, B, ld, &carry, sf);
        }
    }
}

template <class PL>
cudaError_t launch_plan(cudaStream_t st, int sm_count, const FastParams& P0) {
    using L = DevLayout<PL>;
    FastParams P = P0;
    {   // per device and per plan instantiation; a cheap host call
        cudaError_t e = cudaFuncSetAttribute(resample_fast_kernel<PL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::smem);
        if (e != cudaSuccess) return e;
    }
    const uint64_t total_warps = (uint64_t)sm_count * L::WARPS;
    // blocks per work item: aim for >= 8 items per warp, keep the recomputed block <= ~6 %
    uint32_t R = P.nblk;
    const uint64_t want_items = total_warps * 8;
    if (P.rows_total < want_items) {
        uint64_t per_row = (want_items + P.rows_total - 1) / P.rows_total;
        R = (uint32_t)((P.nblk + per_row - 1) / per_row);
        if (R < 16) R = 16;
        if (R > P.nblk) R = P.nblk;
    }
    P.R = R;
    P.items_per_row = (P.nblk + R - 1) / R;
    P.nitems = P.rows_total * P.items_per_row;
    cudaError_t e = cudaMemsetAsync(P.counter, 0, sizeof(unsigned long long), st);
    if (e != cudaSuccess) return e;
    uint64_t ctas = (P.nitems + L::WARPS - 1) / L::WARPS;
    if (ctas > (uint64_t)sm_count) ctas = sm_count;
    resample_fast_kernel<PL><<<(unsigned)ctas, L::WARPS * 32, L::smem, st>>>(P);
    return cudaGetLastError();
}

}  // namespace

// Device tables for the fast path; returns false when no compile-time plan matches.
bool fast_plan_available(uint32_t n_in, uint32_t n_out) {
    if (const char* g = std::getenv("BIRDA_K2_GENERIC")) if (g[0] == '1') return false;
#define BB_PLAN(NAME, NI, NO, ...) if (n_in == NI && n_out == NO) return true;
    BB_K2_PLANS(BB_PLAN)
#undef BB_PLAN
    return false;
}

template <class PL>
static cudaError_t init_tables(const ResamplerSpec& spec, ResamplerDev* rs) {
    const double pi = 3.14159265358979323846;
    int fwd[16], inv[16];
    for (int i = 0; i < PL::Fwd::count; ++i) fwd[i] = PL::Fwd::at(i);
    for (int i = 0; i < PL::Inv::count; ++i) inv[i] = PL::Inv::at(i);
    std::vector<uint16_t> pf(PL::N), pi_(PL::M);
    build_pos_tables(fwd, PL::Fwd::count, inv, PL::Inv::count, PL::N, PL::M, pf.data(), pi_.data());
    std::vector<float2> Pt(PL::NKEEP), Qt(PL::NKEEP), WI(PL::M / 2 + 1), twf(PL::TWF), twi(PL::TWI);
    build_split_tables(PL::N, PL::M, PL::NKEEP, spec.filt_re.data(), spec.filt_im.data(), Pt.data(), Qt.data(), WI.data());
    for (int k = 0; k < PL::TWF; ++k) twf[k] = make_float2((float)cos(-2 * pi * k / PL::N), (float)sin(-2 * pi * k / PL::N));
    for (int k = 0; k < PL::TWI; ++k) twi[k] = make_float2((float)cos(2 * pi * k / PL::M), (float)sin(2 * pi * k / PL::M));
    auto up = [](const void* h, size_t bytes, void** d) -> cudaError_t {
        cudaError_t e = cudaMalloc(d, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e;
    if ((e = up(twf.data(), twf.size() * 8, (void**)&rs->f_twf)) != cudaSuccess) return e;
    if ((e = up(twi.data(), twi.size() * 8, (void**)&rs->f_twi)) != cudaSuccess) return e;
    if ((e = up(pf.data(), pf.size() * 2, (void**)&rs->f_pos_f)) != cudaSuccess) return e;
    if ((e = up(pi_.data(), pi_.size() * 2, (void**)&rs->f_pos_i)) != cudaSuccess) return e;
    if ((e = up(Pt.data(), Pt.size() * 8, (void**)&rs->f_P)) != cudaSuccess) return e;
    if ((e = up(Qt.data(), Qt.size() * 8, (void**)&rs->f_Q)) != cudaSuccess) return e;
    if ((e = up(WI.data(), WI.size() * 8, (void**)&rs->f_WI)) != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&rs->f_counter, sizeof(unsigned long long))) != cudaSuccess) return e;
    rs->fast = true;
    return cudaSuccess;
}

cudaError_t fast_tables_init(const ResamplerSpec& spec, ResamplerDev* rs) {
#define BB_PLAN(NAME, NI, NO, ...) if (spec.n_in == NI && spec.n_out == NO) return init_tables<__VA_ARGS__>(spec, rs);
    BB_K2_PLANS(BB_PLAN)
#undef BB_PLAN
    return cudaErrorInvalidConfiguration;
}

void fast_tables_free(ResamplerDev* rs) {
    void* ptrs[] = {rs->f_twf, rs->f_twi, rs->f_pos_f, rs->f_pos_i, rs->f_P, rs->f_Q, rs->f_WI, rs->f_counter};
    for (void* p : ptrs) if (p) cudaFree(p);
    rs->f_twf = rs->f_twi = rs->f_P = rs->f_Q = rs->f_WI = nullptr; rs->f_pos_f = rs->f_pos_i = nullptr; rs->f_counter = nullptr;
    rs->fast = false;
}

cudaError_t launch_resample_fast(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                                 uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                                 uint64_t nseg, uint64_t last_start, uint64_t rows_total, uint64_t seg,
                                 uint64_t resampled_len, float* d_out, int* launches) {
    if (launches) *launches = 0;
    if (rows_total == 0) return cudaSuccess;
    FastParams P{};
    P.pcm = d_pcm; P.fmt = fmt; P.channels = channels; P.total_frames = total_frames;
    P.src_seg = src_seg; P.hop = hop; P.nseg = nseg; P.last_start = last_start; P.rows_total = rows_total; P.seg = seg;
    P.out_len = (uint32_t)(resampled_len < seg ? resampled_len : seg);
    P.out = d_out;
    P.twf = rs.f_twf; P.twi = rs.f_twi; P.Pt = rs.f_P; P.Qt = rs.f_Q; P.WI = rs.f_WI; P.pos_f = rs.f_pos_f; P.pos_i = rs.f_pos_i;
    P.counter = rs.f_counter;
    P.nblk = (P.out_len + rs.n_out - 1) / rs.n_out;
    if (P.nblk == 0) P.nblk = 1;
    cudaError_t e = cudaErrorInvalidConfiguration;
#define BB_PLAN(NAME, NI, NO, ...) if (rs.n_in == NI && rs.n_out == NO) e = launch_plan<__VA_ARGS__>(st, sm_count, P);
    BB_K2_PLANS(BB_PLAN)
#undef BB_PLAN
    if (e == cudaSuccess && launches) *launches = 1;
    return e;
}

}  // namespace bb
