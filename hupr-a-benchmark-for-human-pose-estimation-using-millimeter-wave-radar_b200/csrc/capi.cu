// Miscellaneous C-ABI entry points (version, error strings).
#include <atomic>
#include <stdlib.h>

#include "common.cuh"

static std::atomic<long long> g_launches{0};

namespace hupr {
void note_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
    int v = g_pdl.load(std::memory_order_relaxed);
    if (v < 0) {
        // default off: measured on B200 (profiles/r02_bench_*pdl*.json) it gains 0.7 % on the batch-1 forward (1.539 -> 1.528 ms) and costs
        // 0.6 % on the batch-32 stream (12.27 -> 12.35 ms); the Python side turns it on for small-batch forwards (ops.pdl)
        const char* e = getenv("HUPR_PDL");
        v = (e && e[0] == '1') ? 1 : 0;
        g_pdl.store(v, std::memory_order_relaxed);
    }
    return v != 0;
}
void set_pdl(int on) { g_pdl.store(on ? 1 : 0, std::memory_order_relaxed); }

int device_index() {
    int dev = -1;
    return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

// [device] -> 0 unknown, else 1 + (major * 100000 + SM count); written once per device (benign race: every writer stores the same value)
static std::atomic<int> g_dev_info[kMaxDevices];

static int device_info() {
    const int dev = device_index();
    if (dev < 0 || dev >= kMaxDevices) return 0;
    int v = g_dev_info[dev].load(std::memory_order_relaxed);
    if (v == 0) {
        int major = 0, sms = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
            return 0;
        v = 1 + major * 100000 + sms;
        g_dev_info[dev].store(v, std::memory_order_relaxed);
    }
    return v;
}

int device_check_sm100() {
    const int v = device_info();
    if (v == 0) return HUPR_ERR_CUDA;
    return (v - 1) / 100000 == 10 ? HUPR_OK : HUPR_ERR_ARCH;
}

int device_sm_count() {
    const int v = device_info();
    return v == 0 ? 0 : (v - 1) % 100000;
}
}  // namespace hupr

extern "C" long long hupr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int hupr_version(void) { return 200; }

namespace hupr { void set_pdl(int on); }
extern "C" int hupr_set_pdl(int on) {
    const int prev = hupr::pdl_enabled() ? 1 : 0;
    if (on >= 0) hupr::set_pdl(on);      // negative: query only
    return prev;
}

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
