// extern "C" surface of birda_b200 (see include/birda_b200.h for the contract and the
// reference file:line each entry point replaces).  No exceptions cross this boundary.
#include "common.cuh"
#include "guard.hpp"
#include <cmath>
#include <cstring>
#include <new>

namespace bb {
static thread_local std::string g_tls_error;
static thread_local int32_t g_fault_countdown = 0;
void set_tls_error(const std::string& m) { g_tls_error = m; }
void fault_point() {
    if (g_fault_countdown > 0 && --g_fault_countdown == 0) throw std::bad_alloc();
}
}  // namespace bb

using namespace bb;

extern "C" {

// ------------------------------------------------------------------------------------ rules
int32_t bb_rule_segment_samples(float segment_duration, float overlap, uint32_t target_rate,
                                int32_t bat_mode, uint64_t* seg, uint64_t* ovl) {
    BB_TRY
    if (!seg || !ovl) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    segment_samples(segment_duration, overlap, target_rate, bat_mode != 0, seg, ovl);
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_source_window(uint64_t seg, uint64_t ovl, uint32_t sr, uint32_t tr, uint64_t* sseg, uint64_t* sovl) {
    BB_TRY
    if (!sseg || !sovl || sr == 0 || tr == 0) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "bad argument");
    source_window(seg, ovl, sr, tr, sseg, sovl);
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_segment_count(uint64_t total, uint64_t sseg, uint64_t sovl, uint64_t* nseg) {
    BB_TRY
    if (!nseg) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    WindowSeq w;
    if (!make_window_seq(total, sseg, sovl, false, &w))
        BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_OVERLAP_GE_SEGMENT,
                   "overlap_samples (" + std::to_string(sovl) + ") must be less than segment_samples (" + std::to_string(sseg) + ")");
    *nseg = w.nseg;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_segment_table(uint64_t total, uint64_t sseg, uint64_t sovl, uint64_t first, uint64_t capacity,
                              uint64_t* start_sample, uint64_t* take, uint64_t* written) {
    BB_TRY
    WindowSeq w;
    if (!make_window_seq(total, sseg, sovl, false, &w))
        BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_OVERLAP_GE_SEGMENT, "overlap_samples must be less than segment_samples");
    uint64_t n = 0;
    for (uint64_t i = first; i < w.nseg && n < capacity; ++i, ++n) {
        Window x = w.at(i);
        if (start_sample) start_sample[n] = x.start;
        if (take) take[n] = x.take;
    }
    if (written) *written = n;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_chunk_times(uint64_t start_sample, uint32_t sr, uint64_t seg, uint32_t tr, float* st, float* et) {
    BB_TRY
    if (!st || !et || sr == 0 || tr == 0) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "bad argument");
    chunk_times(start_sample, sr, seg, tr, st, et);
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_estimate_segment_count(double duration, int32_t has_duration, float seg_dur, float overlap, int64_t* est) {
    BB_TRY
    if (!est) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    *est = estimate_segment_count(duration, has_duration != 0, seg_dur, overlap);
    return BB_OK;
    BB_CATCH(nullptr)
}

uint32_t bb_rule_effective_batch_size(uint32_t batch, int64_t est) { return effective_batch_size(batch, est); }

int32_t bb_rule_resampler_blocks(uint32_t sr, uint32_t tr, uint32_t* n_in, uint32_t* n_out, uint32_t* n_keep, float* cutoff) {
    BB_TRY
    ResamplerSpec s; std::string err;
    if (!make_resampler_spec(sr, tr, false, &s, &err)) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_UNSUPPORTED_RATE, err);
    if (n_in) *n_in = s.n_in; if (n_out) *n_out = s.n_out; if (n_keep) *n_keep = s.n_keep; if (cutoff) *cutoff = s.cutoff;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_resampler_taps(uint32_t sr, uint32_t tr, float* taps, uint32_t n) {
    BB_TRY
    ResamplerSpec s; std::string err;
    if (!make_resampler_spec(sr, tr, false, &s, &err)) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_UNSUPPORTED_RATE, err);
    if (!taps || n != s.n_in) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "taps buffer must hold n_in floats");
    std::memcpy(taps, s.taps.data(), sizeof(float) * n);
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_rule_resampled_len(uint64_t src_len, uint32_t sr, uint32_t tr, uint64_t* out_len) {
    BB_TRY
    if (!out_len) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    if (sr == tr) { *out_len = src_len; return BB_OK; }
    ResamplerSpec s; std::string err;
    if (!make_resampler_spec(sr, tr, false, &s, &err)) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_UNSUPPORTED_RATE, err);
    *out_len = resampled_len(src_len, s);
    return BB_OK;
    BB_CATCH(nullptr)
}

uint32_t bb_rule_date_to_week(uint32_t month, uint32_t day) { return date_to_week(month, day); }
uint32_t bb_rule_week_to_start_day(uint32_t week) { return week_to_start_day(week); }
void     bb_rule_day_of_year_to_date(uint32_t doy, uint32_t* m, uint32_t* d) { uint32_t a, b; day_of_year_to_date(doy, &a, &b); if (m) *m = a; if (d) *d = b; }

// ---------------------------------------------------------------------------------- context
void bb_debug_inject_alloc_failure(int32_t nth) { bb::g_fault_countdown = nth > 0 ? nth : 0; }

uint32_t bb_version(void) { return (BB_VERSION_MAJOR << 16) | BB_VERSION_MINOR; }

int32_t bb_device_count(int32_t* count) {
    BB_TRY
    if (!count) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; cudaGetLastError(); BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_NO_DEVICE, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *count = n;
    return BB_OK;
    BB_CATCH(nullptr)
}

static int32_t ctx_create_impl(int32_t device, void* stream, bool have_stream, bb_ctx** out) {
    BB_TRY
    if (!out) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_NO_DEVICE,
                   "no CUDA device: birda_b200 has no CPU fallback (" + std::string(e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") + ")");
    }
    if (device < 0 || device >= n) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "device index out of range");
    BB_DEVICE((bb_ctx*)nullptr, device);
    bb_ctx* c = new (std::nothrow) bb_ctx();
    if (!c) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_OOM, "out of host memory");
    c->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) c->sm_count = prop.multiProcessorCount;
    if (have_stream) { c->stream = (cudaStream_t)stream; c->owns_stream = false; }
    else {
        e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete c; BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
        c->owns_stream = true;
    }
    *out = c;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_ctx_create(int32_t device, bb_ctx** out) { return ctx_create_impl(device, nullptr, false, out); }
int32_t bb_ctx_create_on_stream(int32_t device, void* stream, bb_ctx** out) { return ctx_create_impl(device, stream, true, out); }

void bb_ctx_destroy(bb_ctx* c) {
    if (!c) return;
    bb::DeviceGuard dev_guard(c->device);
    cudaStreamSynchronize(c->stream);
    if (c->sync_event) cudaEventDestroy(c->sync_event);
    if (c->d_post_index) cudaFree(c->d_post_index);
    if (c->d_post_conf) cudaFree(c->d_post_conf);
    if (c->d_post_count) cudaFree(c->d_post_count);
    if (c->owns_stream) cudaStreamDestroy(c->stream);
    delete c;
}

const char* bb_last_error(const bb_ctx* c) { return c ? c->last_error.c_str() : g_tls_error.c_str(); }
void* bb_ctx_stream(bb_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t bb_ctx_kernel_launches(const bb_ctx* c) { return c ? c->launches : 0; }
void bb_ctx_set_blocking_sync(bb_ctx* c, int32_t on) { if (c) c->blocking_sync = on != 0; }

int32_t bb_sync(bb_ctx* c) {
    BB_TRY
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    BB_DEVICE(c, c->device);
    BB_CUDA_OK(c, bb::ctx_stream_wait(c));
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

int32_t bb_host_alloc(uint64_t bytes, void** out) {
    BB_TRY
    if (!out) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null output");
    BB_CUDA_OK((bb_ctx*)nullptr, cudaHostAlloc(out, bytes, cudaHostAllocDefault));
    return BB_OK;
    BB_CATCH(nullptr)
}
void bb_host_free(void* p) { if (p) cudaFreeHost(p); }

int32_t bb_dev_alloc(bb_ctx* c, uint64_t bytes, void** out) {
    BB_TRY
    if (!c || !out) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "bad argument");
    BB_DEVICE(c, c->device);
    BB_CUDA_OK(c, cudaMalloc(out, bytes ? bytes : 1));
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}
void bb_dev_free(bb_ctx* c, void* p) { if (c && p) { bb::DeviceGuard dev_guard(c->device); cudaFree(p); } }

int32_t bb_memcpy_h2d(bb_ctx* c, void* dst, const void* src, uint64_t bytes) {
    BB_TRY
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    BB_CUDA_OK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}
int32_t bb_memcpy_d2h(bb_ctx* c, void* dst, const void* src, uint64_t bytes) {
    BB_TRY
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    BB_CUDA_OK(c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

// ------------------------------------------------------------------------------------- plan
int32_t bb_plan_create(bb_ctx* c, uint32_t src_rate, uint32_t channels, bb_sample_fmt fmt, uint32_t tgt_rate,
                       uint64_t seg, uint64_t ovl, bb_plan** out) {
    BB_TRY
    if (!c || !out) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "bad argument");
    *out = nullptr;
    if (src_rate == 0 || tgt_rate == 0 || channels == 0 || seg == 0)
        BB_SET_ERR(c, BB_ERR_INVALID_ARG, "rates, channels and segment_samples must be > 0");
    if (fmt != BB_S16 && fmt != BB_S32 && fmt != BB_F32 && fmt != BB_S24)
        BB_SET_ERR(c, BB_ERR_UNSUPPORTED_FORMAT, "unsupported sample format (S16, S24, S32, F32 are converted; decode.rs:353-411)");
    bb_plan* p = new (std::nothrow) bb_plan();
    if (!p) BB_SET_ERR(c, BB_ERR_OOM, "out of host memory");
    p->ctx = c; p->src_rate = src_rate; p->tgt_rate = tgt_rate; p->channels = channels; p->fmt = fmt;
    p->bytes_per_sample = sample_bytes(fmt);
    p->seg = seg; p->ovl = ovl;
    source_window(seg, ovl, src_rate, tgt_rate, &p->src_seg, &p->src_ovl);
    if (p->src_ovl >= p->src_seg) {
        std::string m = "overlap_samples (" + std::to_string(p->src_ovl) + ") must be less than segment_samples (" + std::to_string(p->src_seg) + ")";
        delete p;
        BB_SET_ERR(c, BB_ERR_OVERLAP_GE_SEGMENT, m);
    }
    p->resample = src_rate != tgt_rate;
    if (p->resample) {
        std::string err;
        if (!make_resampler_spec(src_rate, tgt_rate, true, &p->spec, &err)) { delete p; BB_SET_ERR(c, BB_ERR_UNSUPPORTED_RATE, err); }
        p->resampled_len = resampled_len(p->src_seg, p->spec);
        bb::DeviceGuard dev_guard(c->device);
        cudaError_t e = resampler_dev_init(p->spec, &p->rs);
        if (e != cudaSuccess) {
            resampler_dev_free(&p->rs); delete p;
            if (e == cudaErrorInvalidConfiguration)
                BB_SET_ERR(c, BB_ERR_UNSUPPORTED_RATE, "resampler blocks for " + std::to_string(src_rate) + " -> " + std::to_string(tgt_rate) + " do not fit in shared memory");
            BB_SET_ERR(c, e == cudaErrorMemoryAllocation ? BB_ERR_OOM : BB_ERR_CUDA, std::string("resampler init: ") + cudaGetErrorString(e));
        }
    }
    *out = p;
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

void bb_plan_destroy(bb_plan* p) {
    if (!p) return;
    bb::DeviceGuard dev_guard(p->ctx->device);
    cudaStreamSynchronize(p->ctx->stream);
    if (p->d_pcm) cudaFree(p->d_pcm);
    if (p->d_out) cudaFree(p->d_out);
    if (p->copy_stream) {
        cudaStreamSynchronize(p->copy_stream);
        for (auto& ev : p->ev_piece) if (ev) cudaEventDestroy(ev);
        if (p->ev_done) cudaEventDestroy(p->ev_done);
        cudaStreamDestroy(p->copy_stream);
    }
    resampler_dev_free(&p->rs);
    delete p;
}

int32_t bb_plan_source_window(const bb_plan* p, uint64_t* sseg, uint64_t* sovl) {
    BB_TRY
    if (!p) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null plan");
    if (sseg) *sseg = p->src_seg; if (sovl) *sovl = p->src_ovl;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_plan_segment_count(const bb_plan* p, uint64_t total_frames, uint64_t* nseg) {
    BB_TRY
    if (!p || !nseg) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "bad argument");
    WindowSeq w; make_window_seq(total_frames, p->src_seg, p->src_ovl, false, &w);
    *nseg = w.nseg;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_plan_describe(const bb_plan* p, char* buf, uint32_t buf_len) {
    BB_TRY
    if (!p || !buf || buf_len == 0) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "bad argument");
    const std::string s = p->resample ? "K2 " + warp_plan_describe(p->rs) : std::string("K1 pack_kernel (no resampling)");
    std::strncpy(buf, s.c_str(), buf_len - 1);
    buf[buf_len - 1] = 0;
    return BB_OK;
    BB_CATCH(nullptr)
}

int32_t bb_frontend_run(bb_plan* p, const void* pcm, uint64_t frames, int32_t pcm_is_device,
                        uint64_t first_start_sample, int32_t is_eof, uint32_t pad_to_batch,
                        float* d_out_user, uint64_t capacity_rows,
                        float** d_segments, uint64_t* start_sample, float* start_time, float* end_time,
                        uint64_t* nseg_out, uint64_t* nseg_padded, uint64_t* consumed_frames) {
    BB_TRY
    if (!p) BB_SET_ERR((bb_ctx*)nullptr, BB_ERR_INVALID_ARG, "null plan");
    bb_ctx* c = p->ctx;
    if (frames > 0 && !pcm) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null pcm");
    WindowSeq w;
    if (!make_window_seq(frames, p->src_seg, p->src_ovl, is_eof == 0, &w))
        BB_SET_ERR(c, BB_ERR_OVERLAP_GE_SEGMENT, "overlap_samples must be less than segment_samples");
    const uint64_t nseg = w.nseg;
    uint64_t rows = nseg;
    if (pad_to_batch > 1 && nseg % pad_to_batch) rows = (nseg / pad_to_batch + 1) * pad_to_batch;
    if (nseg_out) *nseg_out = nseg;
    if (nseg_padded) *nseg_padded = rows;
    if (consumed_frames) *consumed_frames = is_eof ? frames : nseg * w.hop;
    if ((start_sample || start_time || end_time || d_out_user) && capacity_rows < (d_out_user ? rows : nseg))
        BB_SET_ERR(c, BB_ERR_CAPACITY, "capacity_rows (" + std::to_string(capacity_rows) + ") < rows produced (" + std::to_string(rows) + ")");
    for (uint64_t i = 0; i < nseg; ++i) {
        const uint64_t st = first_start_sample + w.at(i).start;
        if (start_sample) start_sample[i] = st;
        if (start_time || end_time) {
            float a, b; chunk_times(st, p->src_rate, p->seg, p->tgt_rate, &a, &b);
            if (start_time) start_time[i] = a; if (end_time) end_time[i] = b;
        }
    }
    if (d_segments) *d_segments = nullptr;
    if (rows == 0) return BB_OK;

    BB_DEVICE(c, c->device);
    const uint64_t frame_bytes = (uint64_t)p->channels * p->bytes_per_sample;
    const uint64_t pcm_bytes = frames * frame_bytes;
    float* d_out = d_out_user;
    if (!d_out) {
        if (p->d_out_rows < rows) {
            BB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
            if (p->d_out) cudaFree(p->d_out);
            p->d_out = nullptr; p->d_out_rows = 0;
            BB_CUDA_OK(c, cudaMalloc(&p->d_out, rows * p->seg * sizeof(float)));
            p->d_out_rows = rows;
        }
        d_out = p->d_out;
    }
    auto launch_rows = [&](const void* d_pcm, uint64_t row_first, uint64_t row_count) -> cudaError_t {
        if (row_count == 0) return cudaSuccess;
        if (!p->resample) {
            c->launches += 1;
            return launch_pack(c->stream, c->sm_count, d_pcm, p->fmt, p->channels, frames, p->src_seg, w.hop, nseg,
                               w.last_start, row_first, row_count, d_out);
        }
        int n = 0;
        cudaError_t e = launch_resample(c->stream, c->sm_count, p->rs, d_pcm, p->fmt, p->channels, frames, p->src_seg, w.hop,
                                        nseg, w.last_start, row_first, row_count, p->seg, p->resampled_len, d_out, &n);
        c->launches += n;
        return e;
    };
    if (pcm_is_device) {
        BB_CUDA_OK(c, launch_rows(pcm, 0, rows));
    } else {
        // Host PCM: copy in pieces on a second stream so the H2D of piece k+1 overlaps the kernel
        // of piece k (the copy is the longer leg: PCIe moves ~55 GB/s, the kernels consume faster).
        if (p->d_pcm_bytes < pcm_bytes) {
            BB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
            if (p->copy_stream) BB_CUDA_OK(c, cudaStreamSynchronize(p->copy_stream));
            if (p->d_pcm) cudaFree(p->d_pcm);
            p->d_pcm = nullptr; p->d_pcm_bytes = 0;
            BB_CUDA_OK(c, cudaMalloc(&p->d_pcm, pcm_bytes ? pcm_bytes : 1));
            p->d_pcm_bytes = pcm_bytes;
        }
        constexpr uint64_t kMinPieceBytes = 8ull << 20;
        uint64_t npieces = pcm_bytes / kMinPieceBytes;
        if (npieces > 16) npieces = 16;
        if (npieces < 1 || nseg < 2) npieces = 1;
        if (npieces > nseg) npieces = nseg;
        if (npieces == 1) {
            if (pcm_bytes) BB_CUDA_OK(c, cudaMemcpyAsync(p->d_pcm, pcm, pcm_bytes, cudaMemcpyHostToDevice, c->stream));
            BB_CUDA_OK(c, launch_rows(p->d_pcm, 0, rows));
        } else {
            if (!p->copy_stream) {
                BB_CUDA_OK(c, cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking));
                for (auto& ev : p->ev_piece) BB_CUDA_OK(c, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
                BB_CUDA_OK(c, cudaEventCreateWithFlags(&p->ev_done, cudaEventDisableTiming));
            }
            // the copy stream must not overwrite PCM that kernels already queued on the ctx stream still read
            BB_CUDA_OK(c, cudaEventRecord(p->ev_done, c->stream));
            BB_CUDA_OK(c, cudaStreamWaitEvent(p->copy_stream, p->ev_done, 0));
            uint64_t copied = 0, row = 0;
            for (uint64_t k = 0; k < npieces; ++k) {
                const uint64_t row_end = (k + 1 == npieces) ? nseg : (nseg * (k + 1)) / npieces;
                // frames the windows [row, row_end) touch: up to the end of window row_end-1
                const Window last = w.at(row_end - 1);
                uint64_t need = last.start + last.take;
                if (k + 1 == npieces) need = frames;
                if (need > copied) {
                    BB_CUDA_OK(c, cudaMemcpyAsync(static_cast<char*>(p->d_pcm) + copied * frame_bytes,
                                                  static_cast<const char*>(pcm) + copied * frame_bytes,
                                                  (need - copied) * frame_bytes, cudaMemcpyHostToDevice, p->copy_stream));
                    copied = need;
                }
                BB_CUDA_OK(c, cudaEventRecord(p->ev_piece[k], p->copy_stream));
                BB_CUDA_OK(c, cudaStreamWaitEvent(c->stream, p->ev_piece[k], 0));
                const uint64_t rend = (k + 1 == npieces) ? rows : row_end;     // padding rows ride with the last piece
                BB_CUDA_OK(c, launch_rows(p->d_pcm, row, rend - row));
                row = rend;
            }
        }
    }
    if (d_segments) *d_segments = d_out;
    return BB_OK;
    BB_CATCH((p && p->ctx ? &p->ctx->last_error : nullptr))
}

// ------------------------------------------------------------------------------------- post
static int32_t post_check(bb_ctx* c, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B, const bb_post_cfg* cfg,
                          const float* d_mask, const uint8_t* d_keep) {
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    if (!cfg || (!d_scores && valid_B)) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null argument");
    if (valid_B > B) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "valid_B > B");
    if (C == 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "C must be > 0");
    if (cfg->top_k < 1 || cfg->top_k > BB_MAX_TOP_K) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "top_k must be in [1, BB_MAX_TOP_K]");
    if (cfg->activation < BB_ACT_NONE || cfg->activation > BB_ACT_SOFTMAX) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "unknown activation");
    if (d_mask && d_keep) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "range mask and species list are mutually exclusive (lib.rs:502-520)");
    return BB_OK;
}

int32_t bb_post_run_device(bb_ctx* c, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                           const bb_post_cfg* cfg, const float* d_mask, const uint8_t* d_keep,
                           uint32_t* d_index, float* d_conf, uint32_t* d_count) {
    BB_TRY
    int32_t rc = post_check(c, d_scores, B, C, valid_B, cfg, d_mask, d_keep);
    if (rc != BB_OK) return rc;
    if (valid_B && (!d_index || !d_conf || !d_count)) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null output");
    BB_DEVICE(c, c->device);
    BB_CUDA_OK(c, launch_post(c->stream, d_scores, B, C, valid_B, *cfg, d_mask, d_keep, d_index, d_conf, d_count));
    if (valid_B) c->launches += 1;
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

int32_t bb_post_run(bb_ctx* c, const float* d_scores, uint32_t B, uint32_t C, uint32_t valid_B,
                    const bb_post_cfg* cfg, const float* d_mask, const uint8_t* d_keep,
                    uint32_t* h_index, float* h_conf, uint32_t* h_count) {
    BB_TRY
    int32_t rc = post_check(c, d_scores, B, C, valid_B, cfg, d_mask, d_keep);
    if (rc != BB_OK) return rc;
    if (valid_B == 0) return BB_OK;
    if (!h_index || !h_conf || !h_count) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null output");
    BB_DEVICE(c, c->device);
    if (c->post_capacity_rows < valid_B || c->post_capacity_k < cfg->top_k) {
        BB_CUDA_OK(c, cudaStreamSynchronize(c->stream));
        if (c->d_post_index) cudaFree(c->d_post_index);
        if (c->d_post_conf) cudaFree(c->d_post_conf);
        if (c->d_post_count) cudaFree(c->d_post_count);
        c->d_post_index = nullptr; c->d_post_conf = nullptr; c->d_post_count = nullptr; c->post_capacity_rows = 0;
        const uint64_t rows = valid_B > 512 ? valid_B : 512;
        BB_CUDA_OK(c, cudaMalloc(&c->d_post_index, rows * BB_MAX_TOP_K * sizeof(uint32_t)));
        BB_CUDA_OK(c, cudaMalloc(&c->d_post_conf, rows * BB_MAX_TOP_K * sizeof(float)));
        BB_CUDA_OK(c, cudaMalloc(&c->d_post_count, rows * sizeof(uint32_t)));
        c->post_capacity_rows = rows; c->post_capacity_k = BB_MAX_TOP_K;
    }
    BB_CUDA_OK(c, launch_post(c->stream, d_scores, B, C, valid_B, *cfg, d_mask, d_keep, c->d_post_index, c->d_post_conf, c->d_post_count));
    c->launches += 1;
    const uint64_t n = (uint64_t)valid_B * cfg->top_k;
    BB_CUDA_OK(c, cudaMemcpyAsync(h_index, c->d_post_index, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    BB_CUDA_OK(c, cudaMemcpyAsync(h_conf, c->d_post_conf, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    BB_CUDA_OK(c, cudaMemcpyAsync(h_count, c->d_post_count, (uint64_t)valid_B * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    BB_CUDA_OK(c, bb::ctx_stream_wait(c));
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

// ------------------------------------------------------------------------------------ dense heads
int32_t bb_dense_run(bb_ctx* c, const float* d_x, uint32_t B, uint32_t K, const float* d_W, const float* d_b, uint32_t N,
                     int32_t activation, float* d_out) {
    BB_TRY
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    if ((uint64_t)B * N == 0) return BB_OK;
    if (!d_x || !d_W || !d_out || K == 0) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null argument");
    if (activation < BB_ACT_NONE || activation > BB_ACT_SOFTMAX) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "unknown activation");
    BB_DEVICE(c, c->device);
    int n = 0;
    BB_CUDA_OK(c, launch_dense(c->stream, d_x, B, K, d_W, d_b, N, activation, d_out, &n));
    c->launches += n;
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

int32_t bb_calibrate_run(bb_ctx* c, const float* d_scores, uint32_t B, uint32_t C, const float* d_a, const float* d_b, float* d_out) {
    BB_TRY
    if (!c) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null context");
    if ((uint64_t)B * C == 0) return BB_OK;
    if (!d_scores || !d_a || !d_out) BB_SET_ERR(c, BB_ERR_INVALID_ARG, "null argument");
    BB_DEVICE(c, c->device);
    BB_CUDA_OK(c, launch_affine_classes(c->stream, d_scores, B, C, d_a, d_b, d_out));
    c->launches += 1;
    return BB_OK;
    BB_CATCH((c ? &c->last_error : nullptr))
}

}  // extern "C"
