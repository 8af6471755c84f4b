// Library identification and error strings for the C ABI (include/unidistill_b200.h).
#include "common.cuh"

unsigned long long g_ud_launches = 0ull;

extern "C" {

const char* ud_version(void) { return "unidistill_b200 sm_100a " __DATE__; }

const char* ud_error_string(int code) {
    switch (code) {
        case UD_OK: return "ok";
        case UD_ERR_INVALID_ARG: return "unidistill_b200: invalid argument";
        case UD_ERR_WORKSPACE: return "unidistill_b200: workspace missing or too small";
        case UD_ERR_UNSUPPORTED: return "unidistill_b200: unsupported shape or size";
        default: return cudaGetErrorString((cudaError_t)code);
    }
}

int ud_launch_count(int reset) {
    unsigned long long v = g_ud_launches;
    if (reset) g_ud_launches = 0ull;
    return v > 0x7FFFFFFFull ? 0x7FFFFFFF : (int)v;
}

}  // extern "C"
