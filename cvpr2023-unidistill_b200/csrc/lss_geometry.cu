// LSS frustum geometry -> BEV voxel index, one kernel.
//
// Replaces  reference unidistill/layers/blocks_3d/mmdet3d/lss_fpn.py:200-240 (LSSFPN.get_geometry: three batched 4x4
// mat-vec products over the (B, ncam, D, H, W) frustum, which torch runs as 87 chunked cuBLAS batched-GEMM launches of
// 1.9 M one-column products, 5.2 ms per bs-4 step) and :311-313 (`((geom - (voxel_coord - voxel_size / 2)) / voxel_size).int()`).
// Same operation order as the reference's tensor program, in fp32:
//     p = ida^-1 f                      f = frustum[d, h, w] = (u, v, depth, 1)
//     q = (p.x * p.z, p.y * p.z, p.z, p.w)
//     r = (sensor2ego intrin^-1) q
//     s = bda r                          (only when a bda matrix is given)
//     idx = trunc((s.xyz - lo) / voxel_size)        -- `.int()` truncates toward zero
// The tiny per-camera matrix inverses / products stay with the caller (torch, B*ncam 4x4 matrices).  A dot product is
// accumulated k = 0..3 with fused multiply-adds; BLAS libraries differ in exactly this (the reference's own CPU and GPU
// runs do not agree to the last bit), which matters only for points within ~1e-6 of a cell boundary -- the parity tests
// compare indices with an explicit borderline rule (tests/test_gpu_glue.py).
#include "common.cuh"

namespace {

struct Mat4 {
    float m[16];
};

__device__ __forceinline__ float4 matvec(const float* __restrict__ m, float4 v) {
    float4 r;
    r.x = fmaf(m[3], v.w, fmaf(m[2], v.z, fmaf(m[1], v.y, m[0] * v.x)));
    r.y = fmaf(m[7], v.w, fmaf(m[6], v.z, fmaf(m[5], v.y, m[4] * v.x)));
    r.z = fmaf(m[11], v.w, fmaf(m[10], v.z, fmaf(m[9], v.y, m[8] * v.x)));
    r.w = fmaf(m[15], v.w, fmaf(m[14], v.z, fmaf(m[13], v.y, m[12] * v.x)));
    return r;
}

// grid (ceil(P / 256), B * ncam), P = D*H*W frustum points of one camera
__global__ void __launch_bounds__(256)
k_lss_geometry(const float4* __restrict__ frustum, int P, int ncam, const float* __restrict__ ida_inv,
               const float* __restrict__ cam2ego, const float* __restrict__ bda, float lox, float loy, float loz,
               float vx, float vy, float vz, int* __restrict__ idx_out, float* __restrict__ geom_out) {
    __shared__ float sm[48];
    const int cam = blockIdx.y;                 // b * ncam + c
    if (threadIdx.x < 16) {
        sm[threadIdx.x] = ida_inv[(size_t)cam * 16 + threadIdx.x];
        sm[16 + threadIdx.x] = cam2ego[(size_t)cam * 16 + threadIdx.x];
        sm[32 + threadIdx.x] = bda ? bda[(size_t)(cam / ncam) * 16 + threadIdx.x] : 0.f;
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float4 v = matvec(sm, __ldg(frustum + p));
    v.x = __fmul_rn(v.x, v.z);
    v.y = __fmul_rn(v.y, v.z);
    v = matvec(sm + 16, v);
    if (bda) v = matvec(sm + 32, v);
    const size_t o = ((size_t)cam * P + p) * 3;
    if (geom_out) {
        geom_out[o + 0] = v.x;
        geom_out[o + 1] = v.y;
        geom_out[o + 2] = v.z;
    }
    if (idx_out) {
        // float -> int conversion with round-toward-zero == torch's .int(); NaN -> 0 like cvt.rzi saturating semantics
        idx_out[o + 0] = __float2int_rz(__fdiv_rn(__fsub_rn(v.x, lox), vx));
        idx_out[o + 1] = __float2int_rz(__fdiv_rn(__fsub_rn(v.y, loy), vy));
        idx_out[o + 2] = __float2int_rz(__fdiv_rn(__fsub_rn(v.z, loz), vz));
    }
}

}  // namespace

extern "C" {

// frustum (D*H*W, 4) f32 [u, v, depth, 1]; ida_inv, cam2ego (B*ncam, 4, 4) f32 row-major; bda (B, 4, 4) f32 or NULL;
// voxel_lo HOST float[3] = voxel_coord - voxel_size / 2 (computed by the caller in fp32, as the reference does),
// voxel_size HOST float[3].  idx (B, ncam, D*H*W, 3) int32 and/or geom (same shape, f32); either may be NULL.
int ud_lss_geometry(const float* frustum, int num_frustum_points, int batch_size, int num_cams, const float* ida_inv,
                    const float* cam2ego, const float* bda, const float* voxel_lo, const float* voxel_size, int* idx,
                    float* geom, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!frustum || !ida_inv || !cam2ego || !voxel_lo || !voxel_size || (!idx && !geom) || num_frustum_points < 1 ||
        batch_size < 1 || num_cams < 1)
        return UD_ERR_INVALID_ARG;
    if ((long long)batch_size * num_cams > 65535) return UD_ERR_UNSUPPORTED;
    dim3 grid(ud_blocks(num_frustum_points, 256), batch_size * num_cams);
    k_lss_geometry<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4*>(frustum), num_frustum_points, num_cams,
                                             ida_inv, cam2ego, bda, voxel_lo[0], voxel_lo[1], voxel_lo[2], voxel_size[0],
                                             voxel_size[1], voxel_size[2], idx, geom);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
