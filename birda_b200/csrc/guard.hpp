// Exception firewall of the C ABI (product code).
//
// include/birda_b200.h promises that nothing throws or aborts across the boundary (the reference forbids panics:
// Cargo.toml:85-87).  The host code inside the library uses std::vector / std::string / std::thread, all of which
// can throw (bad_alloc, length_error, system_error), so every extern "C" entry point that can reach such code puts
// its body between BB_TRY and BB_CATCH(err): an exception becomes BB_ERR_OOM (bad_alloc) or BB_ERR_INTERNAL, with the
// message left where bb_last_error() / bb_pipeline_last_error() find it.
//
// bb_debug_inject_alloc_failure(n) (header: "debug hooks") arms a thread-local countdown; the n-th guarded entry on
// that thread throws std::bad_alloc from inside the guard, which is how tests/test_abi.py proves the firewall
// without exhausting memory.
#pragma once
#include <cstdint>
#include <exception>
#include <new>
#include <string>
#include "../../include/birda_b200.h"

namespace bb {
void set_tls_error(const std::string& m);
void fault_point();      // throws std::bad_alloc when the thread's injected-failure countdown reaches zero

// Call from a catch (...) handler: maps the exception in flight to a status code and records the message.
// `also`: the error string of the object the call belongs to (bb_ctx::last_error, bb_pipeline::error) or null.
inline int32_t translate_exception(const char* where, std::string* also) noexcept {
    int32_t code = BB_ERR_INTERNAL;
    const char* what = "unknown exception";
    std::string held;
    try { throw; }
    catch (const std::bad_alloc&) { code = BB_ERR_OOM; what = "out of host memory"; }
    catch (const std::exception& e) { try { held = e.what(); what = held.c_str(); } catch (...) {} }
    catch (...) {}
    try {
        const std::string m = std::string(where) + ": " + what;
        set_tls_error(m);
        if (also) *also = m;
    } catch (...) {}
    return code;
}
}  // namespace bb

#define BB_TRY try { bb::fault_point();
#define BB_CATCH(also) } catch (...) { return bb::translate_exception(__func__, (also)); }
