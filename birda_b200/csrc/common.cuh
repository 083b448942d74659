// Internal types of the birda_b200 library (product code).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include "rules.hpp"
#include "../../include/birda_b200.h"

struct bb_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    bool owns_stream = false;
    std::string last_error;
    uint64_t launches = 0;
    // bb_ctx_set_blocking_sync: waits sleep on an event instead of spinning (a directory run has several worker threads
    // per GPU and needs the host cores for file reads)
    bool blocking_sync = false; cudaEvent_t sync_event = nullptr;
    // scratch for bb_post_run (host-output variant)
    uint32_t* d_post_index = nullptr; float* d_post_conf = nullptr; uint32_t* d_post_count = nullptr;
    uint64_t post_capacity_rows = 0; uint32_t post_capacity_k = 0;
};

namespace bb {

// Device-side description of the resampler for one (from, to) pair.
struct ResamplerDev {
    uint32_t n_in = 0, n_out = 0, n_keep = 0;
    int nstage_fwd = 0, nstage_inv = 0;
    int radix_fwd[16] = {0}, radix_inv[16] = {0};
    float2* d_tw_fwd = nullptr;     // [n_in]   exp(-2*pi*i*k/n_in)
    float2* d_tw_inv = nullptr;     // [n_out]  exp(+2*pi*i*k/n_out)
    float2* d_split_fwd = nullptr;  // [n_in+1] exp(-2*pi*i*k/(2*n_in))
    float2* d_split_inv = nullptr;  // [n_out]  exp(+2*pi*i*k/(2*n_out))
    float2* d_filt = nullptr;       // [n_keep] filter spectrum
    uint32_t buf_len = 0;           // complex elements per ping/pong buffer
    // blocks too long for shared memory (rates with a small gcd, e.g. 8 001 Hz -> 48 kHz: 2667 / 16000): the fallback
    // kernel keeps its tables and buffers in a per-CTA slice of this global workspace instead (slow, but it resamples)
    unsigned char* d_ws = nullptr; size_t ws_stride = 0; uint32_t ws_ctas = 0;
    // warp-per-block path (runtime plan, k2_warp.cu)
    bool fast = false;
    alignas(8) unsigned char plan_blob[768] = {0};
    int ct_index = -1;                 // compile-time plan (k2_warp.cuh: BB_K2_CT_PLANS) or -1
    float2 *f_twf = nullptr, *f_twi = nullptr, *f_WI = nullptr;
    uint4* f_sidx = nullptr; float4 *f_pq1 = nullptr, *f_pq2 = nullptr;   // split-pass layout (k2_warp.cuh: build_split_layout)
    uint32_t *f_ordf = nullptr, *f_ordi = nullptr; uint32_t ordf_len = 0, ordi_len = 0;   // per-stage order tables (build_stage_orders)
    unsigned long long* f_counter = nullptr;
};

}  // namespace bb

struct bb_plan {
    bb_ctx* ctx = nullptr;
    uint32_t src_rate = 0, tgt_rate = 0, channels = 0;
    int fmt = 0;
    uint32_t bytes_per_sample = 0;
    uint64_t seg = 0, ovl = 0;            // target-rate counts
    uint64_t src_seg = 0, src_ovl = 0;    // source-rate window
    bool resample = false;
    bb::ResamplerSpec spec;
    bb::ResamplerDev rs;
    uint64_t resampled_len = 0;           // samples resample() returns for src_seg inputs
    // device buffers
    void* d_pcm = nullptr; uint64_t d_pcm_bytes = 0;
    // host-input pipeline: H2D pieces on a copy stream overlap the kernels of earlier pieces
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_piece[32] = {nullptr};
    cudaEvent_t ev_done = nullptr;      // last kernel of the previous run (the copy stream waits on it)
    bool ev_done_valid = false;
    float* d_out = nullptr; uint64_t d_out_rows = 0;
};

namespace bb {
// Makes `device` current for the scope and restores the caller's device on exit: a host that shares its thread
// with another CUDA user (ONNX Runtime, torch) keeps the device it had.
struct DeviceGuard {
    int prev = -1; bool changed = false; cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
        if (prev != device) { err = cudaSetDevice(device); changed = err == cudaSuccess && prev >= 0; }
    }
    ~DeviceGuard() { if (changed) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
}  // namespace bb
#define BB_DEVICE(ctx, device) bb::DeviceGuard _bb_dev_guard(device); BB_CUDA_OK(ctx, _bb_dev_guard.err)

namespace bb {
// wait for the context's stream: spinning cudaStreamSynchronize by default, a blocking-sync event when the context asks
inline cudaError_t ctx_stream_wait(bb_ctx* c) {
    if (!c->blocking_sync) return cudaStreamSynchronize(c->stream);
    if (!c->sync_event) { cudaError_t e = cudaEventCreateWithFlags(&c->sync_event, cudaEventBlockingSync | cudaEventDisableTiming); if (e != cudaSuccess) return e; }
    cudaError_t e = cudaEventRecord(c->sync_event, c->stream);
    if (e != cudaSuccess) return e;
    return cudaEventSynchronize(c->sync_event);
}
}  // namespace bb

#define BB_SET_ERR(ctx, code, msg) do { if (ctx) (ctx)->last_error = (msg); else bb::set_tls_error(msg); return (code); } while (0)
#define BB_CUDA_OK(ctx, expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { \
        std::string _m = std::string(#expr) + ": " + cudaGetErrorString(_e); \
        if (ctx) (ctx)->last_error = _m; else bb::set_tls_error(_m); \
        return _e == cudaErrorMemoryAllocation ? BB_ERR_OOM : BB_ERR_CUDA; } } while (0)

namespace bb {
void set_tls_error(const std::string& m);

// K1: convert + downmix + window gather + pack (no resampling).  k1_pack.cu
cudaError_t launch_pack(cudaStream_t st, int sm_count, const void* d_pcm, int fmt, uint32_t channels,
                        uint64_t total_frames, uint64_t seg, uint64_t hop, uint64_t nseg,
                        uint64_t last_start, uint64_t row_first, uint64_t rows_total, float* d_out);

// K2: convert + downmix + window gather + per-window block-FFT resample + pack.  k2_resample.cu
cudaError_t launch_resample(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                            uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                            uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                            uint64_t resampled_len, float* d_out, int* launches);
cudaError_t resampler_dev_init(const ResamplerSpec& spec, ResamplerDev* rs);
// K2 warp-per-block kernel (runtime plans; even N_out with supported radices).  k2_warp.cu
bool        warp_plan_available(const ResamplerSpec& spec);
std::string warp_plan_describe(const ResamplerDev& rs);     // which kernel / plan / blocking a launch will use
cudaError_t warp_tables_init(const ResamplerSpec& spec, ResamplerDev* rs);
void        warp_tables_free(ResamplerDev* rs);
cudaError_t launch_resample_warp(cudaStream_t st, int sm_count, const ResamplerDev& rs, const void* d_pcm, int fmt,
                                 uint32_t channels, uint64_t total_frames, uint64_t src_seg, uint64_t hop,
                                 uint64_t nseg, uint64_t last_start, uint64_t row_first, uint64_t rows_total, uint64_t seg,
                                 uint64_t resampled_len, float* d_out, int* launches);
void        resampler_dev_free(ResamplerDev* rs);

// K4: tiny dense heads (geomodel forward, bat head).  k4_dense.cu
cudaError_t launch_affine_classes(cudaStream_t st, const float* d_x, uint32_t B, uint32_t C, const float* d_a, const float* d_b, float* d_out);
cudaError_t launch_dense(cudaStream_t st, const float* d_x, uint32_t B, uint32_t K, const float* d_W, const float* d_b,
                         uint32_t N, int activation, float* d_out, int* launches);

// K3: activation + top-k + threshold + mask + threshold.  k3_post.cu
cudaError_t launch_post(cudaStream_t st, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                        const bb_post_cfg& cfg, const float* d_mask, const uint8_t* d_keep,
                        uint32_t* d_index, float* d_conf, uint32_t* d_count);
}  // namespace bb
