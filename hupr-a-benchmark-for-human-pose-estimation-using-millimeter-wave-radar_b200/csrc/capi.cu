// Miscellaneous C-ABI entry points (version, error strings).
#include "common.cuh"

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
