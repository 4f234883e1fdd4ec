// Miscellaneous C-ABI entry points (version, error strings).
#include <atomic>

#include "common.cuh"

static std::atomic<long long> g_launches{0};

namespace hupr {
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace hupr

extern "C" long long hupr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int hupr_version(void) { return 100; }

extern "C" const char* hupr_error_string(int code) {
    switch (code) {
        case HUPR_OK: return "ok";
        case HUPR_ERR_BAD_ARG: return "bad argument (null pointer, negative or unsupported size)";
        case HUPR_ERR_ALIGNMENT: return "pointer not aligned as required";
        case HUPR_ERR_CUDA: return "CUDA runtime error (see cudaGetLastError)";
        case HUPR_ERR_ARCH: return "device is not sm_100 (B200); there is no fallback path";
        case HUPR_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}
