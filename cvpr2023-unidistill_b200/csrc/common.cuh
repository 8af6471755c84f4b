// Shared helpers for the sm_100a kernels of the UniDistill BEV hot path.
// Torch-free on purpose: every launcher is `extern "C"`, takes raw device
// pointers + a cudaStream_t and returns a cudaError_t-style int (0 == ok).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define UD_OK 0
#define UD_ERR_INVALID_ARG 1001
#define UD_ERR_WORKSPACE 1002
#define UD_ERR_UNSUPPORTED 1003

#define UD_CUDA_TRY(expr)                          \
    do {                                           \
        cudaError_t _e = (expr);                   \
        if (_e != cudaSuccess) return (int)_e;     \
    } while (0)

// number of kernels this library launched since the last ud_launch_count(1) (bench.py's gpu_launches)
extern unsigned long long g_ud_launches;

#define UD_LAUNCH_CHECK()                          \
    do {                                           \
        cudaError_t _e = cudaGetLastError();       \
        if (_e != cudaSuccess) return (int)_e;     \
        ++g_ud_launches;                           \
    } while (0)

#define UD_MAX_BATCH 64
#define UD_NUM_SMS 148

static inline unsigned ud_blocks(long long n, int threads) {
    long long b = (n + threads - 1) / threads;
    if (b < 1) b = 1;
    return (unsigned)b;
}

static inline size_t ud_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct UdArena {
    char* base;
    size_t off;
    size_t cap;
    __host__ UdArena(void* p, size_t bytes) : base((char*)p), off(0), cap(bytes) {}
    template <typename T>
    __host__ T* take(size_t count) {
        size_t bytes = ud_align_up(count * sizeof(T));
        T* r = (T*)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    __host__ bool ok() const { return base != nullptr && off <= cap; }
};

__device__ __forceinline__ float ud_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float ud_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// 32-bit integer hash (murmur3 finalizer) used by the coordinate hash tables.
__host__ __device__ __forceinline__ uint32_t ud_hash32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
