// Directory batches across the GPUs of one box (product code): a host work queue of files, longest first, one worker
// thread + context + per-file pipeline per GPU, per-file detection lists handed back to the caller (the "merge" of
// the reference is per file: src/lib.rs:694-796 runs process_file once per file and writes its outputs; files are
// independent end to end, SURVEY.md 8e — no collective, no inter-GPU traffic).  The classifier callback is the same
// as bb_pipeline's; it is called from the worker thread of its device with that device's `user` pointer.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>
#include "rules.hpp"
#include "guard.hpp"
#include "pipeline_internal.hpp"
#include "../../include/birda_b200.h"


struct bb_pool {
    struct Worker { int32_t device; bb_ctx* ctx; bb_pipeline* pipe; };
    std::vector<Worker> workers;
};

namespace {
int pool_fail(int code, const std::string& m) { bb::set_tls_error(m); return code; }
}

extern "C" {

void bb_pool_destroy(bb_pool* p) {
    if (!p) return;
    for (auto& w : p->workers) {
        if (w.pipe) bb_pipeline_destroy(w.pipe);
        if (w.ctx) bb_ctx_destroy(w.ctx);
    }
    delete p;
}

int32_t bb_pool_create(const int32_t* devices, uint32_t n_devices, const bb_pipeline_cfg* cfgs, bb_classify_fn fn,
                       void* const* users, bb_pool** out) {
    BB_TRY
    if (!devices || n_devices == 0 || !cfgs || !fn || !out) return pool_fail(BB_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    bb_pool* p = new (std::nothrow) bb_pool();
    if (!p) return pool_fail(BB_ERR_OOM, "out of host memory");
    for (uint32_t i = 0; i < n_devices; ++i) {
        bb_pool::Worker w{devices[i], nullptr, nullptr};
        int rc = bb_ctx_create(devices[i], &w.ctx);
        if (rc == BB_OK) rc = bb_pipeline_create(w.ctx, &cfgs[i], fn, users ? users[i] : nullptr, &w.pipe);
        // the pool owns the contexts, so the callback cannot know their streams: the packed windows are complete
        // (stream synchronised) before it is called; it must in turn finish its own work before it returns
        if (rc == BB_OK) bb_pipeline_set_sync_before_classify(w.pipe, 1);
        if (rc == BB_OK) { const char* e = std::getenv("BIRDA_POOL_SPIN"); bb_ctx_set_blocking_sync(w.ctx, (e && e[0] == '1') ? 0 : 1); }   // workers sleep while the GPU works: the cores read files
        p->workers.push_back(w);
        if (rc != BB_OK) { bb_pool_destroy(p); return rc; }          // the failing call left the message
    }
    *out = p;
    return BB_OK;
    BB_CATCH(nullptr)
}

bb_ctx* bb_pool_worker_ctx(bb_pool* p, uint32_t worker) { return (p && worker < p->workers.size()) ? p->workers[worker].ctx : nullptr; }

void bb_pool_set_stream_ordered(bb_pool* p, int32_t on) {
    if (!p) return;
    for (auto& w : p->workers) bb_pipeline_set_sync_before_classify(w.pipe, on ? 0 : 1);
}

uint64_t bb_pool_kernel_launches(const bb_pool* p) {
    uint64_t n = 0;
    if (p) for (const auto& w : p->workers) n += bb_ctx_kernel_launches(w.ctx);
    return n;
}

void bb_pool_free_results(bb_pool_result* results, uint32_t n) {
    if (!results) return;
    for (uint32_t i = 0; i < n; ++i) { delete[] results[i].detections; results[i].detections = nullptr; results[i].n_detections = 0; }
}

int32_t bb_pool_process_wavs(bb_pool* p, const char* const* paths, uint32_t n_files, bb_pool_result* results) {
    BB_TRY
    if (!p || (!paths && n_files) || (!results && n_files)) return pool_fail(BB_ERR_INVALID_ARG, "null argument");
    // longest first: by duration (the reference's unit of work, src/lib.rs:694), ties by PCM bytes (what the host side
    // of this path pays for).  Unreadable files go last and report their error from the worker.
    std::vector<std::pair<double, uint32_t>> order(n_files);
    for (uint32_t i = 0; i < n_files; ++i) {
        std::memset(&results[i], 0, sizeof(results[i]));
        bb_wav_info info{};
        double key = -1.0;
        if (bb_wav_probe(paths[i], &info) == BB_OK && info.sample_rate)
            key = (double)info.frames / info.sample_rate + 1e-12 * (double)info.frames * info.channels * (info.bits_per_sample / 8);
        order[i] = {key, i};
    }
    std::stable_sort(order.begin(), order.end(), [](const auto& a, const auto& b) { return a.first > b.first; });
    std::atomic<uint32_t> next{0};
    auto work = [&](bb_pool::Worker& w) {
        std::vector<bb_detection> buf;
        for (;;) {
            const uint32_t k = next.fetch_add(1);
            if (k >= n_files) break;
            const uint32_t f = order[k].second;
            bb_pool_result& r = results[f];
            r.device = w.device;
            uint64_t ns = 0; uint32_t bu = 0;
            int rc;
            try { rc = bb::pipeline_process_wav_into(w.pipe, paths[f], 0, &buf, &ns, &bu); }      // the list grows as needed: one pass per file
            catch (...) { rc = bb::translate_exception("bb_pool worker", nullptr); }
            r.status = rc; r.n_segments = ns; r.batch_used = bu;
            if (rc != BB_OK) {
                const char* m = rc == BB_ERR_OOM || !*bb_pipeline_last_error(w.pipe) ? bb_last_error(nullptr) : bb_pipeline_last_error(w.pipe);
                std::strncpy(r.error, m ? m : "", sizeof(r.error) - 1);
                continue;
            }
            const uint64_t nd = buf.size();
            r.detections = new (std::nothrow) bb_detection[nd ? nd : 1];
            if (!r.detections) { r.status = BB_ERR_OOM; continue; }
            std::memcpy(r.detections, buf.data(), nd * sizeof(bb_detection));
            r.n_detections = nd;
        }
    };
    std::vector<std::thread> threads;
    for (size_t i = 1; i < p->workers.size(); ++i) threads.emplace_back(work, std::ref(p->workers[i]));
    work(p->workers[0]);                                              // the caller's thread serves the first device
    for (auto& t : threads) t.join();
    int32_t worst = BB_OK;
    for (uint32_t i = 0; i < n_files; ++i) if (results[i].status != BB_OK) worst = results[i].status;
    return worst;
    BB_CATCH(nullptr)
}

}  // extern "C"
