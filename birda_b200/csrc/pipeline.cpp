// Host pipeline over the C ABI (product code, C++): the per-file loop of the reference with the front
// end and the post step on the GPU and the classifier as a callback (ONNX Runtime with IoBinding in
// birda; any device function in tests).
//
// Follows process_file / run_streaming_inference / process_batch
// (src/pipeline/processor.rs:418-796, :114-190, :220-410):
//   estimate_segment_count -> effective batch size (:525-545) -> front end over the file in pieces ->
//   batches of `batch` rows, the last one padded with silence (:239-260) -> classifier -> post step on
//   the valid rows (:317, :363-385) -> detections sorted by (start_time asc, confidence desc) (:178-187).
#include "../../include/birda_b200.h"
#include "rules.hpp"
#include "guard.hpp"
#include "watchdog.hpp"
#include "pieces.hpp"
#include "pipeline_internal.hpp"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>


struct bb_pipeline {
    bb_ctx* ctx = nullptr;
    bb_pipeline_cfg cfg{};
    bb_classify_fn classify = nullptr;
    void* user = nullptr;
    bb_plan* plan = nullptr;                  // the plan of the file being processed (owned by `cache`)
    uint32_t plan_rate = 0, plan_channels = 0; int plan_fmt = 0;
    // A directory mixes sample rates and channel counts (BASELINE config 5): plans are kept per (rate, channels,
    // format), least recently used evicted, so a file only pays for plan creation (resampler tables, ~50 ms) the
    // first time its kind is seen.
    struct Cached { bb_plan* plan; uint32_t rate, channels; int fmt; uint64_t last_use; };
    std::vector<Cached> cache;
    uint64_t use_clock = 0;
    uint64_t plans_created = 0;
    bool sync_before_classify = false;        // bb_pipeline_set_sync_before_classify
    uint32_t read_threads = 4;                // bb_pipeline_set_read_threads
    // per-batch seam (processor.rs:263-277): host hooks and / or a library-kept watchdog
    bb_batch_hook before_batch = nullptr, after_batch = nullptr; void* hook_user = nullptr;
    uint64_t batch_timeout_ms = 0; bb_watchdog_fn on_timeout = nullptr; void* timeout_user = nullptr;
    std::unique_ptr<bb::Watchdog> watchdog;
    // two pinned staging buffers: a reader thread fills one with piece k+1 while the GPU works on piece k
    void* pinned[2] = {nullptr, nullptr}; uint64_t pinned_bytes[2] = {0, 0};
    // post-step results of one piece stay on the device and come back in ONE copy per piece (bb_post_run_device per
    // batch); with batch hooks or a watchdog set every batch is synchronous instead, as in the reference
    bb_flac* flac = nullptr;                  // FLAC files: decoder state (device buffers), created on first use
    uint32_t* d_index = nullptr; float* d_conf = nullptr; uint32_t* d_count = nullptr; uint64_t d_res_rows = 0; uint32_t d_res_k = 0;
    std::vector<uint32_t> h_index; std::vector<float> h_conf; std::vector<uint32_t> h_count;
    std::vector<float> st, et; std::vector<uint64_t> ss;
    std::string error;
};

namespace {

int fail(bb_pipeline* p, int code, const std::string& m) { if (p) p->error = m; bb::set_tls_error(m); return code; }

constexpr size_t kPlanCache = 8;

int ensure_plan(bb_pipeline* p, uint32_t rate, uint32_t channels, int fmt) {
    ++p->use_clock;
    for (auto& c : p->cache)
        if (c.rate == rate && c.channels == channels && c.fmt == fmt) {
            c.last_use = p->use_clock;
            p->plan = c.plan; p->plan_rate = rate; p->plan_channels = channels; p->plan_fmt = fmt;
            return BB_OK;
        }
    if (p->cache.size() >= kPlanCache) {
        size_t lru = 0;
        for (size_t i = 1; i < p->cache.size(); ++i) if (p->cache[i].last_use < p->cache[lru].last_use) lru = i;
        bb_plan_destroy(p->cache[lru].plan);
        p->cache.erase(p->cache.begin() + (long)lru);
    }
    p->plan = nullptr;
    const uint32_t target = p->cfg.bat_mode ? rate : p->cfg.target_rate;            // processor.rs:464-475
    uint64_t seg = 0, ovl = 0;
    bb_rule_segment_samples(p->cfg.segment_duration, p->cfg.overlap, target, p->cfg.bat_mode, &seg, &ovl);
    bb_plan* plan = nullptr;
    int rc = bb_plan_create(p->ctx, rate, channels, (bb_sample_fmt)fmt, target, seg, ovl, &plan);
    if (rc != BB_OK) return fail(p, rc, bb_last_error(p->ctx));
    p->cache.push_back({plan, rate, channels, fmt, p->use_clock});
    ++p->plans_created;
    p->plan = plan; p->plan_rate = rate; p->plan_channels = channels; p->plan_fmt = fmt;
    return BB_OK;
}

// detections go to the caller's array (count beyond its capacity: BB_ERR_CAPACITY) or to a growing vector (bb_pool)
struct Sink {
    bb_detection* out; uint64_t cap; uint64_t n; bool overflow; std::vector<bb_detection>* grow;
    void push(uint32_t segment, uint32_t index, float conf, float st, float et) {
        bb_detection d{}; d.segment = segment; d.index = index; d.confidence = conf; d.start_time = st; d.end_time = et;
        if (grow) grow->push_back(d);
        else if (n < cap) out[n] = d;
        else overflow = true;
        ++n;
    }
};

void free_results(bb_pipeline* p) {
    if (p->d_index) bb_dev_free(p->ctx, p->d_index);
    if (p->d_conf) bb_dev_free(p->ctx, p->d_conf);
    if (p->d_count) bb_dev_free(p->ctx, p->d_count);
    p->d_index = nullptr; p->d_conf = nullptr; p->d_count = nullptr; p->d_res_rows = 0; p->d_res_k = 0;
}
int ensure_results(bb_pipeline* p, uint64_t rows, uint32_t K) {
    if (p->d_res_rows >= rows && p->d_res_k == K) return BB_OK;
    if (bb_sync(p->ctx) != BB_OK) return BB_ERR_CUDA;
    free_results(p);
    void* a = nullptr; void* b = nullptr; void* c = nullptr;
    if (bb_dev_alloc(p->ctx, rows * K * 4, &a) != BB_OK || bb_dev_alloc(p->ctx, rows * K * 4, &b) != BB_OK ||
        bb_dev_alloc(p->ctx, rows * 4, &c) != BB_OK) {
        for (void* q : {a, b, c}) if (q) bb_dev_free(p->ctx, q);
        return BB_ERR_OOM;
    }
    p->d_index = static_cast<uint32_t*>(a); p->d_conf = static_cast<float*>(b); p->d_count = static_cast<uint32_t*>(c);
    p->d_res_rows = rows; p->d_res_k = K;
    return BB_OK;
}

// one piece of PCM already in host memory -> detections appended to the sink
int run_piece(bb_pipeline* p, const void* pcm, uint64_t frames, uint64_t first_start, bool eof, uint32_t B,
              uint64_t seg_base, Sink* sink, uint64_t* nseg_out, uint64_t* consumed, bool pcm_is_device = false) {
    uint64_t src_seg = 0, src_ovl = 0, nmax = 0;
    bb_plan_source_window(p->plan, &src_seg, &src_ovl);
    bb_rule_segment_count(frames, src_seg, src_ovl, &nmax);
    const uint64_t cap = (nmax / B + 1) * B;
    p->st.resize(cap); p->et.resize(cap); p->ss.resize(cap);
    float* d_seg = nullptr; uint64_t nseg = 0, rows = 0;
    int rc = bb_frontend_run(p->plan, pcm, frames, pcm_is_device ? 1 : 0, first_start, eof ? 1 : 0, B, nullptr, cap, &d_seg, p->ss.data(),
                             p->st.data(), p->et.data(), &nseg, &rows, consumed);
    if (rc != BB_OK) return fail(p, rc, bb_last_error(p->ctx));
    *nseg_out = nseg;
    uint64_t seg_samples = 0, dummy = 0;
    bb_rule_segment_samples(p->cfg.segment_duration, p->cfg.overlap, p->cfg.bat_mode ? p->plan_rate : p->cfg.target_rate,
                            p->cfg.bat_mode, &seg_samples, &dummy);
    const uint32_t K = p->cfg.post.top_k;
    // processor.rs:363-385: second threshold test (matters after rerank), one Detection per surviving prediction
    auto extract = [&](uint64_t first, uint32_t valid, const uint32_t* idx, const float* conf, const uint32_t* cnt) {
        for (uint32_t r = 0; r < valid; ++r)
            for (uint32_t j = 0; j < cnt[r] && j < K; ++j) {
                const float c = conf[(size_t)r * K + j];
                if (!(c >= p->cfg.post.min_confidence)) continue;
                sink->push((uint32_t)(seg_base + first + r), idx[(size_t)r * K + j], c, p->st[first + r], p->et[first + r]);
            }
    };
    const bool per_batch_sync = p->before_batch || p->after_batch || p->watchdog;
    if (per_batch_sync) {
        p->h_index.resize((size_t)B * K); p->h_conf.resize((size_t)B * K); p->h_count.resize(B);
        for (uint64_t first = 0; first < nseg; first += B) {
            const uint32_t valid = (uint32_t)std::min<uint64_t>(B, nseg - first);
            const float* d_scores = nullptr; uint32_t classes = 0;
            if (p->sync_before_classify && bb_sync(p->ctx) != BB_OK) return fail(p, BB_ERR_CUDA, bb_last_error(p->ctx));
            // the watchdog brackets the batch exactly as processor.rs:263-277 does: armed before the classifier is called,
            // dropped once its results are on the host (bb_post_run synchronises the stream)
            if (p->before_batch) p->before_batch(p->hook_user, B, valid, seg_base + first);
            if (p->watchdog) p->watchdog->arm(p->batch_timeout_ms, B);
            rc = p->classify(p->user, d_seg + first * seg_samples, B, (uint32_t)seg_samples, &d_scores, &classes);
            int prc = BB_OK;
            if (rc == 0 && d_scores && classes)
                prc = bb_post_run(p->ctx, d_scores, B, classes, valid, &p->cfg.post, p->cfg.d_mask, p->cfg.d_species_keep,
                                  p->h_index.data(), p->h_conf.data(), p->h_count.data());
            const bool fired = p->watchdog ? p->watchdog->disarm() : false;
            if (p->after_batch) p->after_batch(p->hook_user, B, valid, seg_base + first);
            if (fired) return fail(p, BB_ERR_TIMEOUT, "inference batch of " + std::to_string(B) + " outlived the watchdog (" +
                                                      std::to_string(p->batch_timeout_ms) + " ms)");
            if (rc != 0 || !d_scores || classes == 0) return fail(p, BB_ERR_INTERNAL, "classifier callback failed");   // Error::Inference
            if (prc != BB_OK) return fail(p, prc, bb_last_error(p->ctx));
            extract(first, valid, p->h_index.data(), p->h_conf.data(), p->h_count.data());
        }
        return BB_OK;
    }
    // no per-batch seam in use: every batch's classifier call and post kernel are queued back to back, the piece's
    // results come back in one copy and one synchronisation
    if (nseg == 0) return BB_OK;
    if (ensure_results(p, cap, K) != BB_OK) return fail(p, BB_ERR_OOM, "device allocation for the post-step results failed");
    for (uint64_t first = 0; first < nseg; first += B) {
        const uint32_t valid = (uint32_t)std::min<uint64_t>(B, nseg - first);
        const float* d_scores = nullptr; uint32_t classes = 0;
        if (p->sync_before_classify && bb_sync(p->ctx) != BB_OK) return fail(p, BB_ERR_CUDA, bb_last_error(p->ctx));
        rc = p->classify(p->user, d_seg + first * seg_samples, B, (uint32_t)seg_samples, &d_scores, &classes);
        if (rc != 0 || !d_scores || classes == 0) return fail(p, BB_ERR_INTERNAL, "classifier callback failed");       // Error::Inference
        rc = bb_post_run_device(p->ctx, d_scores, B, classes, valid, &p->cfg.post, p->cfg.d_mask, p->cfg.d_species_keep,
                                p->d_index + first * K, p->d_conf + first * K, p->d_count + first);
        if (rc != BB_OK) return fail(p, rc, bb_last_error(p->ctx));
    }
    p->h_index.resize((size_t)nseg * K); p->h_conf.resize((size_t)nseg * K); p->h_count.resize(nseg);
    if (bb_memcpy_d2h(p->ctx, p->h_index.data(), p->d_index, nseg * K * 4) != BB_OK ||
        bb_memcpy_d2h(p->ctx, p->h_conf.data(), p->d_conf, nseg * K * 4) != BB_OK ||
        bb_memcpy_d2h(p->ctx, p->h_count.data(), p->d_count, nseg * 4) != BB_OK || bb_sync(p->ctx) != BB_OK)
        return fail(p, BB_ERR_CUDA, bb_last_error(p->ctx));
    extract(0, (uint32_t)nseg, p->h_index.data(), p->h_conf.data(), p->h_count.data());
    return BB_OK;
}

void sort_detections(bb_detection* d, uint64_t n) {                                 // processor.rs:178-187
    std::stable_sort(d, d + n, [](const bb_detection& a, const bb_detection& b) {
        if (a.start_time != b.start_time) return a.start_time < b.start_time;
        return a.confidence > b.confidence;
    });
}

uint32_t effective_batch(const bb_pipeline* p, uint64_t frames, uint32_t rate) {
    const float seg_dur = p->cfg.bat_mode ? (float)144000 / (float)256000 : p->cfg.segment_duration;   // constants.rs:535
    int64_t est = -1;
    bb_rule_estimate_segment_count((double)frames / (double)rate, 1, seg_dur, p->cfg.overlap, &est);
    uint32_t B = bb_rule_effective_batch_size(p->cfg.batch_size, est);
    return B ? B : 1;
}

int ensure_pinned(bb_pipeline* p, int slot, uint64_t bytes) {
    if (p->pinned_bytes[slot] >= bytes) return BB_OK;
    if (p->pinned[slot]) bb_host_free(p->pinned[slot]);
    p->pinned[slot] = nullptr; p->pinned_bytes[slot] = 0;
    if (bb_host_alloc(bytes > 0 ? bytes : 1, &p->pinned[slot]) != BB_OK) return BB_ERR_OOM;
    p->pinned_bytes[slot] = bytes;
    return BB_OK;
}

// A WAV file through the pipeline in pieces of ~256 MB of PCM.  Pieces are cut so that every piece but the last holds a
// whole number of batches (only the file's LAST batch is padded, processor.rs:132-170).  A file of several pieces is
// read by a second thread into the other pinned buffer while the GPU works on the current piece (the reference's decode
// thread, processor.rs:20-47).  run_piece ends with the stream synchronised, so a buffer is free again two pieces later.
// A FLAC file: the compressed bytes are read into pinned memory and copied to the GPU, decoded there (K6) and handed to
// the front end as device-resident PCM — one piece, whatever the length (an hour of 44.1 kHz stereo is 0.6 GB decoded).
int process_flac(bb_pipeline* p, const char* path, Sink* sink, uint64_t* n_segments, uint32_t* batch_used) {
    bb_flac_info info;
    int rc = bb_flac_probe(path, &info);
    if (rc != BB_OK) return fail(p, rc, bb_last_error(nullptr));
    rc = ensure_plan(p, info.sample_rate, info.channels, info.fmt);
    if (rc != BB_OK) return rc;
    if (bb_sync(p->ctx) != BB_OK) return fail(p, BB_ERR_CUDA, bb_last_error(p->ctx));
    if (ensure_pinned(p, 0, info.file_bytes) != BB_OK) return fail(p, BB_ERR_OOM, "pinned staging allocation failed");
    rc = bb::read_range_parallel(path, 0, info.file_bytes, p->pinned[0], p->read_threads);
    if (rc != BB_OK) return fail(p, rc, std::string("cannot read ") + path);
    if (!p->flac && bb_flac_create(p->ctx, &p->flac) != BB_OK) return fail(p, BB_ERR_OOM, "FLAC decoder state");
    void* d_pcm = nullptr; uint64_t frames = 0;
    rc = bb_flac_decode(p->flac, p->pinned[0], info.file_bytes, &info, &d_pcm, &frames);
    if (rc != BB_OK) return fail(p, rc, bb_last_error(p->ctx));
    const uint32_t B = effective_batch(p, info.frames ? info.frames : frames, info.sample_rate);      // duration hint: STREAMINFO's count
    if (batch_used) *batch_used = B;
    uint64_t nseg = 0, consumed = 0;
    rc = run_piece(p, d_pcm, frames, 0, true, B, 0, sink, &nseg, &consumed, true);
    if (rc != BB_OK) return rc;
    if (n_segments) *n_segments = nseg;
    return BB_OK;
}

bool is_flac(const char* path) {
    unsigned char m[4] = {0, 0, 0, 0};
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    const size_t n = std::fread(m, 1, 4, f);
    std::fclose(f);
    return n == 4 && std::memcmp(m, "fLaC", 4) == 0;
}

int process_wav(bb_pipeline* p, const char* path, uint64_t piece_frames, Sink* sink, uint64_t* n_segments, uint32_t* batch_used) {
    if (is_flac(path)) return process_flac(p, path, sink, n_segments, batch_used);
    bb_wav_info info;
    int rc = bb_wav_probe(path, &info);
    if (rc != BB_OK) return fail(p, rc, bb_last_error(nullptr));
    rc = ensure_plan(p, info.sample_rate, info.channels, info.fmt);
    if (rc != BB_OK) return rc;
    const uint32_t B = effective_batch(p, info.frames, info.sample_rate);
    if (batch_used) *batch_used = B;
    const uint64_t fb = (uint64_t)info.channels * bb::sample_bytes(info.fmt);
    uint64_t src_seg = 0, src_ovl = 0;
    bb_plan_source_window(p->plan, &src_seg, &src_ovl);
    if (piece_frames == 0) piece_frames = (256ull << 20) / fb;
    if (piece_frames < 2 * src_seg) piece_frames = 2 * src_seg;
    const uint64_t hop = src_seg - src_ovl;
    // the piece table is a pure function of the header: (first frame, frames, last piece) — pieces.hpp
    using Piece = bb::Piece;
    const std::vector<Piece> pieces = bb::plan_pieces(info.frames, piece_frames, src_seg, hop, B);
    uint64_t max_bytes = 0;
    for (const Piece& pc : pieces) max_bytes = std::max(max_bytes, pc.frames * fb);
    const int nbuf = pieces.size() > 1 ? 2 : 1;
    if (bb_sync(p->ctx) != BB_OK) return fail(p, BB_ERR_CUDA, bb_last_error(p->ctx));   // earlier H2D copies out of the staging buffers are done
    for (int b = 0; b < nbuf; ++b) if (ensure_pinned(p, b, max_bytes) != BB_OK) return fail(p, BB_ERR_OOM, "pinned staging allocation failed");

    // reader: piece k goes to buffer k % nbuf once piece k - nbuf has been consumed
    std::mutex mu; std::condition_variable cv;
    size_t read_done = 0, consumed_n = 0, fail_at = ~(size_t)0; int read_rc = BB_OK; bool stop = false; std::string read_err;
    auto read_piece = [&](size_t k) { return bb_wav_read_parallel(path, &info, pieces[k].pos, pieces[k].frames, p->pinned[k % nbuf], p->read_threads); };
    std::thread reader;
    if (nbuf == 2)
        reader = std::thread([&] {
            for (size_t k = 0; k < pieces.size(); ++k) {
                { std::unique_lock<std::mutex> l(mu); cv.wait(l, [&] { return stop || k < consumed_n + 2; }); if (stop) return; }
                const int r = read_piece(k);
                { std::lock_guard<std::mutex> l(mu); if (r != BB_OK) { read_rc = r; fail_at = k; read_err = bb_last_error(nullptr); } read_done = k + 1; }
                cv.notify_all();
                if (r != BB_OK) return;
            }
        });
    struct Joiner { std::thread& t; std::mutex& mu; std::condition_variable& cv; bool& stop;
                    ~Joiner() { if (t.joinable()) { { std::lock_guard<std::mutex> l(mu); stop = true; } cv.notify_all(); t.join(); } } } joiner{reader, mu, cv, stop};

    // BIRDA_PIPE_TRACE=1: per-file wall times of the read and of the GPU leg on stderr (tools/prof_c5.py reads them)
    static const bool trace = [] { const char* e = std::getenv("BIRDA_PIPE_TRACE"); return e && e[0] == '1'; }();
    using clk = std::chrono::steady_clock;
    double t_read = 0, t_run = 0;
    uint64_t seg_base = 0;
    for (size_t k = 0; k < pieces.size(); ++k) {
        const auto t0 = clk::now();
        if (nbuf == 2) {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [&] { return read_done > k; });
            if (fail_at == k) return fail(p, read_rc, read_err);
        } else {
            rc = read_piece(k);
            if (rc != BB_OK) return fail(p, rc, bb_last_error(nullptr));
        }
        const auto t1 = clk::now();
        uint64_t nseg = 0, consumed = 0;
        rc = run_piece(p, p->pinned[k % nbuf], pieces[k].frames, pieces[k].pos, pieces[k].eof, B, seg_base, sink, &nseg, &consumed);
        if (rc != BB_OK) return rc;
        t_read += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t_run += std::chrono::duration<double, std::milli>(clk::now() - t1).count();
        seg_base += nseg;
        if (!pieces[k].eof && pieces[k].pos + consumed != pieces[k + 1].pos) return fail(p, BB_ERR_INTERNAL, "piece table and front end disagree");
        if (nbuf == 2) { { std::lock_guard<std::mutex> l(mu); consumed_n = k + 1; } cv.notify_all(); }
    }
    if (n_segments) *n_segments = seg_base;
    if (trace) std::fprintf(stderr, "[pipe] %s %.1f MB read %.2f ms gpu-leg %.2f ms\n", path, (double)info.frames * fb / 1e6, t_read, t_run);
    return BB_OK;
}

}  // namespace

namespace bb {
int32_t pipeline_process_wav_into(bb_pipeline* p, const char* path, uint64_t piece_frames, std::vector<bb_detection>* out,
                                  uint64_t* n_segments, uint32_t* batch_used) {
    if (!p || !path || !out) return fail(p, BB_ERR_INVALID_ARG, "bad argument");
    out->clear();
    Sink sink{nullptr, 0, 0, false, out};
    const int rc = process_wav(p, path, piece_frames, &sink, n_segments, batch_used);
    if (rc != BB_OK) return rc;
    sort_detections(out->data(), out->size());
    return BB_OK;
}
}  // namespace bb

extern "C" {

int32_t bb_pipeline_create(bb_ctx* ctx, const bb_pipeline_cfg* cfg, bb_classify_fn fn, void* user, bb_pipeline** out) {
    BB_TRY
    if (!ctx || !cfg || !fn || !out) return fail(nullptr, BB_ERR_INVALID_ARG, "null argument");
    if (cfg->batch_size < 1 || cfg->batch_size > 512) return fail(nullptr, BB_ERR_INVALID_ARG, "batch_size must be in [1, 512] (constants.rs:44-55)");
    if (cfg->post.top_k < 1 || cfg->post.top_k > BB_MAX_TOP_K) return fail(nullptr, BB_ERR_INVALID_ARG, "top_k out of range");
    bb_pipeline* p = new (std::nothrow) bb_pipeline();
    if (!p) return fail(nullptr, BB_ERR_OOM, "out of host memory");
    p->ctx = ctx; p->cfg = *cfg; p->classify = fn; p->user = user;
    if (const char* e = std::getenv("BIRDA_READ_THREADS")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) p->read_threads = (uint32_t)v; }
    *out = p;
    return BB_OK;
    BB_CATCH(nullptr)
}

void bb_pipeline_destroy(bb_pipeline* p) {
    if (!p) return;
    for (auto& c : p->cache) bb_plan_destroy(c.plan);
    for (void* b : p->pinned) if (b) bb_host_free(b);
    if (p->flac) bb_flac_destroy(p->flac);
    free_results(p);
    delete p;
}

const char* bb_pipeline_last_error(const bb_pipeline* p) { return p ? p->error.c_str() : ""; }
uint64_t bb_pipeline_plans_created(const bb_pipeline* p) { return p ? p->plans_created : 0; }
void bb_pipeline_set_sync_before_classify(bb_pipeline* p, int32_t on) { if (p) p->sync_before_classify = on != 0; }

void bb_pipeline_set_read_threads(bb_pipeline* p, uint32_t threads) { if (p) p->read_threads = threads ? threads : 1; }

void bb_pipeline_set_batch_hooks(bb_pipeline* p, bb_batch_hook before, bb_batch_hook after, void* user) {
    if (!p) return;
    p->before_batch = before; p->after_batch = after; p->hook_user = user;
}

void bb_pipeline_set_batch_timeout(bb_pipeline* p, uint64_t timeout_ms, bb_watchdog_fn on_fire, void* user) {
    if (!p) return;
    p->batch_timeout_ms = timeout_ms; p->on_timeout = on_fire; p->timeout_user = user;
    p->watchdog.reset();
    if (timeout_ms == 0) return;
    try { p->watchdog.reset(new bb::Watchdog(on_fire, user)); }
    catch (...) { p->batch_timeout_ms = 0; fail(p, BB_ERR_INTERNAL, "could not start the watchdog thread"); }
}

int32_t bb_pipeline_process_pcm(bb_pipeline* p, const void* pcm, uint64_t frames, uint32_t src_rate, uint32_t channels,
                                int32_t fmt, bb_detection* out, uint64_t capacity, uint64_t* n_detections,
                                uint64_t* n_segments, uint32_t* batch_used) {
    BB_TRY
    if (!p || (!pcm && frames) || src_rate == 0) return fail(p, BB_ERR_INVALID_ARG, "bad argument");
    int rc = ensure_plan(p, src_rate, channels, fmt);
    if (rc != BB_OK) return rc;
    const uint32_t B = effective_batch(p, frames, src_rate);
    if (batch_used) *batch_used = B;
    Sink sink{out, out ? capacity : 0, 0, false, nullptr};
    uint64_t nseg = 0, consumed = 0;
    rc = run_piece(p, pcm, frames, 0, true, B, 0, &sink, &nseg, &consumed);
    if (rc != BB_OK) return rc;
    if (n_segments) *n_segments = nseg;
    if (n_detections) *n_detections = sink.n;
    if (sink.overflow) return fail(p, BB_ERR_CAPACITY, "detection capacity too small (" + std::to_string(sink.n) + " needed)");
    sort_detections(out, sink.n);
    return BB_OK;
    BB_CATCH((p ? &p->error : nullptr))
}

int32_t bb_pipeline_process_wav(bb_pipeline* p, const char* path, uint64_t piece_frames, bb_detection* out, uint64_t capacity,
                                uint64_t* n_detections, uint64_t* n_segments, uint32_t* batch_used) {
    BB_TRY
    if (!p || !path) return fail(p, BB_ERR_INVALID_ARG, "bad argument");
    Sink sink{out, out ? capacity : 0, 0, false, nullptr};
    int rc = process_wav(p, path, piece_frames, &sink, n_segments, batch_used);
    if (rc != BB_OK) return rc;
    if (n_detections) *n_detections = sink.n;
    if (sink.overflow) return fail(p, BB_ERR_CAPACITY, "detection capacity too small (" + std::to_string(sink.n) + " needed)");
    sort_detections(out, sink.n);
    return BB_OK;
    BB_CATCH((p ? &p->error : nullptr))
}

}  // extern "C"
