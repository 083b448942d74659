// Watchdog (csrc/watchdog.cpp) under ThreadSanitizer: eight threads start / fire / cancel short timers (tools/fuzz/run.sh).
#include <atomic>
#include <cstdio>
#include <thread>
#include <vector>
#include <chrono>
#include "../../include/birda_b200.h"
static std::atomic<int> fired{0};
static void on_fire(void*, uint64_t, uint32_t) { fired++; }
int main() {
    // many short-lived watchdogs from several threads: start, maybe let it fire, cancel
    std::vector<std::thread> th;
    for (int t = 0; t < 8; ++t) th.emplace_back([t] {
        for (int i = 0; i < 300; ++i) {
            bb_watchdog* w = nullptr;
            if (bb_watchdog_start((i % 3 == 0) ? 1 : 50, 64, on_fire, nullptr, &w) != 0) { printf("start failed\n"); return; }
            if (i % 3 == 0) std::this_thread::sleep_for(std::chrono::milliseconds(3));
            bb_watchdog_cancel(w);
        }
    });
    for (auto& x : th) x.join();
    printf("fired %d of %d short timers\n", fired.load(), 8 * 100);
    return 0;
}
