// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map construction through the driver entry point (no -lcuda link) and
// the device-side 2-D tiled load.  A CUtensorMap is built per launch from the caller's pointer (a pure host computation,
// ~1 us) and passed to the kernel as a __grid_constant__ parameter.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "common.cuh"

typedef CUresult (*UdEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline UdEncodeTiledFn ud_encode_tiled_fn() {
    static UdEncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (UdEncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major tensor [rows][cols] of 4-byte elements (row pitch pitch_bytes, multiple of 16), box [box_rows][box_cols],
// no interleave, out-of-bounds elements read as zero.  base 16-byte aligned.  swizzle_128b: the box's rows must be <= 128
// bytes; the 16-byte chunks of shared-memory row r are XOR-ed with (r & 7) (destination 1024-byte aligned).
static inline int ud_make_tmap_2d_b32(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, uint64_t rows,
                                      uint64_t cols, uint64_t pitch_bytes, uint32_t box_rows, uint32_t box_cols,
                                      bool swizzle_128b = false) {
    UdEncodeTiledFn enc = ud_encode_tiled_fn();
    if (enc == nullptr) return UD_ERR_UNSUPPORTED;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           swizzle_128b ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? UD_OK : UD_ERR_INVALID_ARG;
}

#ifdef __CUDACC__
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// box at (col0, row0) -> shared memory `dst` (128-byte aligned), completion bytes on mbarrier `bar`
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int col0, int row0, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(col0), "r"(row0), "r"(bar)
        : "memory");
}
#endif
