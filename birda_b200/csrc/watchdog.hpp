// Inference watchdog (src/gpu/watchdog.rs:22-66), see watchdog.cpp.
#pragma once
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include "../../include/birda_b200.h"

namespace bb {

class Watchdog {
public:
    Watchdog(bb_watchdog_fn on_fire, void* user);      // on_fire == nullptr: the reference's message + exit(1)
    ~Watchdog();
    Watchdog(const Watchdog&) = delete;
    Watchdog& operator=(const Watchdog&) = delete;
    void arm(uint64_t timeout_ms, uint32_t batch_size);
    bool disarm();                                       // true when the timer fired since the last arm()
private:
    void run();
    bb_watchdog_fn on_fire_; void* user_;
    std::mutex mu_; std::condition_variable cv_; std::thread thread_;
    bool quit_ = false, armed_ = false, fired_ = false;
    uint64_t generation_ = 0, timeout_ms_ = 0; uint32_t batch_ = 0;
    std::chrono::steady_clock::time_point deadline_;
};

uint64_t inference_timeout_secs(const char* env_value);

}  // namespace bb
