// Inference watchdog (product code): the native form of src/gpu/watchdog.rs:22-66 and of the timeout rule at
// src/pipeline/processor.rs:194-211.  The reference spawns a sleeping thread per batch and kills the process when
// the batch outlives the timeout; here one thread per watchdog object waits on a condition variable, so arming and
// disarming cost a lock, and the per-file pipeline (pipeline.cpp) keeps one for all its batches.
#include "watchdog.hpp"
#include "guard.hpp"
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>

namespace bb {

// watchdog.rs:31-49 — the text the reference prints before process::exit(1)
static void default_fire(void*, uint64_t timeout_secs, uint32_t batch_size) {
    const uint32_t suggested = batch_size / 2 ? batch_size / 2 : 1;
    std::fprintf(stderr,
                 "\n═══════════════════════════════════════════════════════════════\n"
                 "FATAL: Inference timeout after %llus (batch size: %u)\n"
                 "═══════════════════════════════════════════════════════════════\n\n"
                 "The GPU inference operation did not complete within the expected time.\n"
                 "This usually indicates GPU memory exhaustion causing the system to hang.\n\n"
                 "Recommendations:\n"
                 "  1. Reduce batch size: birda -b %u <input>\n"
                 "  2. Use CPU inference: birda --cpu <input>\n"
                 "  3. Close other GPU applications and try again\n\n"
                 "Terminating process to prevent system lockup.\n",
                 (unsigned long long)timeout_secs, batch_size, suggested);
    std::fflush(stderr);
    std::_Exit(1);
}

Watchdog::Watchdog(bb_watchdog_fn on_fire, void* user) : on_fire_(on_fire ? on_fire : default_fire), user_(user) {
    thread_ = std::thread([this] { run(); });
}

Watchdog::~Watchdog() {
    { std::lock_guard<std::mutex> l(mu_); quit_ = true; armed_ = false; }
    cv_.notify_all();
    if (thread_.joinable()) thread_.join();
}

void Watchdog::arm(uint64_t timeout_ms, uint32_t batch_size) {
    { std::lock_guard<std::mutex> l(mu_);
      armed_ = true; fired_ = false; ++generation_;
      timeout_ms_ = timeout_ms; batch_ = batch_size;
      deadline_ = std::chrono::steady_clock::now() + std::chrono::milliseconds(timeout_ms); }
    cv_.notify_all();
}

bool Watchdog::disarm() {
    std::lock_guard<std::mutex> l(mu_);
    armed_ = false;
    return fired_;
}

void Watchdog::run() {
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
        cv_.wait(l, [this] { return quit_ || armed_; });
        if (quit_) return;
        const uint64_t gen = generation_;
        // sleep until the deadline of THIS arming; a disarm or a re-arm wakes the wait early
        if (cv_.wait_until(l, deadline_, [&] { return quit_ || !armed_ || generation_ != gen; })) continue;
        armed_ = false; fired_ = true;
        const uint64_t secs = timeout_ms_ / 1000; const uint32_t b = batch_;
        l.unlock();
        on_fire_(user_, secs, b);          // default: does not return
        l.lock();
    }
}

uint64_t inference_timeout_secs(const char* v) {                  // processor.rs:194-211
    constexpr uint64_t kDefault = 10, kMin = 1, kMax = 3600;
    if (!v || !*v) return kDefault;
    // Rust's u64::from_str: an optional '+', then decimal digits only, no spaces, no overflow
    const char* p = v;
    if (*p == '+') ++p;
    if (!*p) return kDefault;
    uint64_t x = 0;
    for (; *p; ++p) {
        if (*p < '0' || *p > '9') return kDefault;
        const uint64_t d = (uint64_t)(*p - '0');
        if (x > (UINT64_MAX - d) / 10) return kDefault;
        x = x * 10 + d;
    }
    return (x >= kMin && x <= kMax) ? x : kDefault;
}

}  // namespace bb

struct bb_watchdog { bb::Watchdog w; bb_watchdog(bb_watchdog_fn f, void* u) : w(f, u) {} };

extern "C" {

uint64_t bb_rule_inference_timeout_secs(const char* env_value) { return bb::inference_timeout_secs(env_value); }

int32_t bb_watchdog_start(uint64_t timeout_ms, uint32_t batch_size, bb_watchdog_fn on_fire, void* user, bb_watchdog** out) {
    BB_TRY
        if (!out) { bb::set_tls_error("null output"); return BB_ERR_INVALID_ARG; }
        *out = nullptr;
        bb_watchdog* w = new bb_watchdog(on_fire, user);       // std::thread may throw system_error: guarded
        w->w.arm(timeout_ms, batch_size);
        *out = w;
        return BB_OK;
    BB_CATCH(nullptr)
}

void bb_watchdog_cancel(bb_watchdog* w) {
    if (!w) return;
    w->w.disarm();
    delete w;
}

}  // extern "C"
