// Input side of the LiDAR branch: sweep merge + range filter + ragged collate on the GPU.
//
// Replaces, for the whole batch in one launch sequence,
//   CollectLidarSweeps.forward   reference unidistill/data/multisensorfusion/transforms3d.py:379-414
//       (per sweep: homogeneous 4x4 transform into the key frame in float64, time lag into the 5th column)
//   ObjectRangeFilter.mask_points_by_range   transforms3d.py:247-256,271-276 (inclusive x/y bounds, float32 compare)
//   the points part of collate_fn   data/multisensorfusion/nuscenes_multimodal.py:418-470 (zero padding to the longest
//       cloud; here the batch stays RAGGED: concatenated rows + per-sample offsets, which is exactly what
//       ud_hard_voxelize takes -- the padded origin points of the reference, quirk 1, never exist).
// Segments: one per (sample, sweep) in sample-major order; segment 0 of a sample is its key frame (identity transform,
// time column forced to 0 as transforms3d.py:388-389 does).  The 3x4 transform is applied in float64 with the operation order
// of a row-by-column product and rounded to float32 once, like the numpy code it replaces.
// Kept points keep their input order (flag -> exclusive scan -> scatter): deterministic.
#include "common.cuh"
#include <cub/cub.cuh>

namespace {

constexpr int IS_MAX_SEGMENTS = 1024;

__device__ __forceinline__ int seg_of(const int* __restrict__ seg_offsets, int n_seg, int i) {
    int lo = 0, hi = n_seg;                  // largest s with seg_offsets[s] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_offsets + mid) <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// pass 1: transform in place into `xyzt` (n, 4 doubles are not stored: results are recomputed in pass 2) -> keep flags
__global__ void __launch_bounds__(256)
k_sweep_flags(const float* __restrict__ pts, int n, int C, const int* __restrict__ seg_offsets, int n_seg,
              const double* __restrict__ mats, float x0, float y0, float x1, float y1, int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { flags[i] = 0; return; }                       // sentinel: scan[n] == kept total
    const int s = seg_of(seg_offsets, n_seg, i);
    const double* m = mats + (size_t)s * 12;
    const float* p = pts + (size_t)i * C;
    const double px = (double)p[0], py = (double)p[1], pz = (double)p[2];
    const float x = (float)(((m[0] * px + m[1] * py) + m[2] * pz) + m[3]);
    const float y = (float)(((m[4] * px + m[5] * py) + m[6] * pz) + m[7]);
    flags[i] = (x >= x0 && x <= x1 && y >= y0 && y <= y1) ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_sweep_emit(const float* __restrict__ pts, int n, int C, const int* __restrict__ seg_offsets, int n_seg,
             const double* __restrict__ mats, const float* __restrict__ time_lag, const int* __restrict__ flags,
             const int* __restrict__ scan, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const int s = seg_of(seg_offsets, n_seg, i);
    const double* m = mats + (size_t)s * 12;
    const float* p = pts + (size_t)i * C;
    const double px = (double)p[0], py = (double)p[1], pz = (double)p[2];
    float* o = out + (size_t)scan[i] * C;
    o[0] = (float)(((m[0] * px + m[1] * py) + m[2] * pz) + m[3]);
    o[1] = (float)(((m[4] * px + m[5] * py) + m[6] * pz) + m[7]);
    o[2] = (float)(((m[8] * px + m[9] * py) + m[10] * pz) + m[11]);
    for (int c = 3; c < C; ++c) o[c] = p[c];
    if (C == 5) o[4] = __ldg(time_lag + s);
}

// out_offsets[b] = kept rows before sample b (b = 0..B), from the scan at the sample boundaries
__global__ void k_sample_offsets(const int* __restrict__ scan, const int* __restrict__ seg_offsets,
                                 const int* __restrict__ sample_first_seg, int B, int* __restrict__ out_offsets) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > B) return;
    out_offsets[b] = scan[seg_offsets[sample_first_seg[b]]];
}

}  // namespace

extern "C" {

int ud_collect_sweeps_workspace_bytes(int n_points, size_t* bytes) {
    if (n_points < 0 || !bytes) return UD_ERR_INVALID_ARG;
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr, n_points + 1)));
    *bytes = 2 * ud_align_up(((size_t)n_points + 1) * 4) + ud_align_up(scan_bytes);
    return UD_OK;
}

// points (n, C) f32: all segments concatenated; seg_offsets (n_seg + 1) DEVICE int32 row offsets; sample_first_seg (B + 1)
// DEVICE int32: first segment of every sample (last entry = n_seg); mats (n_seg, 3, 4) DEVICE f64: sweep -> key-frame
// transform of every segment (identity for key frames); time_lag (n_seg) DEVICE f32: value of the 5th column (C == 5).
// range_xy HOST float[4] = (x_min, y_min, x_max, y_max), inclusive.  out (n, C) f32 (capacity n rows); out_offsets (B + 1)
// DEVICE int32: kept rows before every sample, out_offsets[B] = total.
int ud_collect_sweeps(const float* points, int n, int C, const int* seg_offsets, int n_seg, const int* sample_first_seg,
                      int batch_size, const double* mats, const float* time_lag, const float* range_xy, float* out,
                      int* out_offsets, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || C < 3 || C > 8 || n_seg < 1 || n_seg > IS_MAX_SEGMENTS || batch_size < 1 || !seg_offsets ||
        !sample_first_seg || !mats || !range_xy || !out_offsets || (C == 5 && !time_lag) || (n > 0 && (!points || !out)))
        return UD_ERR_INVALID_ARG;
    UdArena a(workspace, workspace_bytes);
    int* flags = a.take<int>((size_t)n + 1);
    int* scan = a.take<int>((size_t)n + 1);
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr, n + 1)));
    void* cub_temp = a.take<char>(scan_bytes);
    if (!a.ok()) return UD_ERR_WORKSPACE;
    k_sweep_flags<<<ud_blocks((long long)n + 1, 256), 256, 0, stream>>>(points, n, C, seg_offsets, n_seg, mats, range_xy[0],
                                                                        range_xy[1], range_xy[2], range_xy[3], flags);
    UD_LAUNCH_CHECK();
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(cub_temp, scan_bytes, flags, scan, n + 1, stream)));
    if (n > 0) {
        k_sweep_emit<<<ud_blocks(n, 256), 256, 0, stream>>>(points, n, C, seg_offsets, n_seg, mats, time_lag, flags, scan,
                                                            out);
        UD_LAUNCH_CHECK();
    }
    k_sample_offsets<<<ud_blocks(batch_size + 1, 64), 64, 0, stream>>>(scan, seg_offsets, sample_first_seg, batch_size,
                                                                       out_offsets);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
