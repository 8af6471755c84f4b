// Deterministic hard voxelization + fused mean-VFE for sm_100a.
//
// Replaces spconv.pytorch.utils.PointToVoxel as driven by
//   reference unidistill/data/det3d/preprocess/voxelization.py:44-73
// and folds in MeanVFE (reference layers/blocks_3d/det3d/vfe/mean_vfe.py:14-34).
//
// Semantics = the first-come CPU voxelizer (oracle/voxelize_ref.c):
//   c_j = floor((p_j - min_j) / vsize_j) in fp32 (IEEE division, no FMA),
//   a point is kept iff 0 <= c_j < grid_j for j in x,y,z;
//   voxels are numbered in order of first appearance inside each sample,
//   at most `max_voxels` voxels per sample (later voxels are skipped, points
//   of existing voxels still accumulate), first `max_points` points per voxel.
//
// GPU formulation (no atomics => bit-reproducible):
//   1. key_i = linear (b,z,y,x) cell id, or `invalid` for rejected points
//   2. stable radix sort of (key, i)     -> points of a voxel are contiguous and
//                                           stay in input order
//   3. segment heads mark flag[first point of voxel] = 1
//   4. exclusive scan of flag over the *point* axis = first-come rank of the voxel
//   5. one thread per head gathers <= max_points rows, writes the padded voxel,
//      its (b,z,y,x), its count and its mean.
// All B samples go through one launch sequence; the only host-visible result
// is counts[0] (total voxels), which the caller may read whenever it wants.
#include "common.cuh"
#include <cub/cub.cuh>

struct UdOffsets {
    int v[UD_MAX_BATCH + 1];
};

struct UdVoxelGeom {
    float vx, vy, vz;
    float minx, miny, minz;
    int gx, gy, gz;
};

__global__ void __launch_bounds__(256)
k_point_keys(const float* __restrict__ pts, int n, int C, UdOffsets offs, int B, UdVoxelGeom g,
             uint32_t invalid_key, uint32_t* __restrict__ keys, int* __restrict__ idx) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int b = 0;
    while (b + 1 < B && i >= offs.v[b + 1]) ++b;
    const float* p = pts + (size_t)i * C;
    // plain IEEE sub + div (division is never contracted into an FMA)
    float fx = floorf(__fdiv_rn(__fsub_rn(p[0], g.minx), g.vx));
    float fy = floorf(__fdiv_rn(__fsub_rn(p[1], g.miny), g.vy));
    float fz = floorf(__fdiv_rn(__fsub_rn(p[2], g.minz), g.vz));
    bool ok = (fx >= 0.f) && (fx < (float)g.gx) && (fy >= 0.f) && (fy < (float)g.gy) &&
              (fz >= 0.f) && (fz < (float)g.gz);
    uint32_t key = invalid_key;
    if (ok) {
        uint32_t cx = (uint32_t)fx, cy = (uint32_t)fy, cz = (uint32_t)fz;
        key = (((uint32_t)b * g.gz + cz) * g.gy + cy) * g.gx + cx;
    }
    keys[i] = key;
    idx[i] = i;
}

__global__ void __launch_bounds__(256)
k_segment_heads(const uint32_t* __restrict__ keys_sorted, const int* __restrict__ idx_sorted, int n,
                uint32_t invalid_key, int* __restrict__ first_flag) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t k = keys_sorted[j];
    if (k == invalid_key) return;
    if (j == 0 || keys_sorted[j - 1] != k) first_flag[idx_sorted[j]] = 1;
}

// counts[0] = total kept voxels, counts[1+b] = kept voxels of sample b.
// meta[b] = first-come rank base of sample b, meta[UD_MAX_BATCH + b] = output row base.
__global__ void k_sample_counts(const int* __restrict__ scan, UdOffsets offs, int B, int max_voxels,
                                int* __restrict__ counts, int* __restrict__ meta) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int out_base = 0;
    for (int b = 0; b < B; ++b) {
        int r0 = scan[offs.v[b]];
        int r1 = scan[offs.v[b + 1]];
        int kept = r1 - r0;
        if (kept > max_voxels) kept = max_voxels;
        meta[b] = r0;
        meta[UD_MAX_BATCH + b] = out_base;
        counts[1 + b] = kept;
        out_base += kept;
    }
    counts[0] = out_base;
}

__global__ void __launch_bounds__(128)
k_emit_voxels(const float* __restrict__ pts, int n, int C, const uint32_t* __restrict__ keys_sorted,
              const int* __restrict__ idx_sorted, const int* __restrict__ scan,
              const int* __restrict__ meta, UdVoxelGeom g, uint32_t invalid_key, int max_points,
              int max_voxels, int out_capacity, float* __restrict__ voxels, int* __restrict__ coords,
              int* __restrict__ num_points, float* __restrict__ mean) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    uint32_t key = keys_sorted[j];
    if (key == invalid_key) return;
    if (j > 0 && keys_sorted[j - 1] == key) return;  // not a segment head

    uint32_t t = key;
    int x = t % g.gx; t /= g.gx;
    int y = t % g.gy; t /= g.gy;
    int z = t % g.gz; t /= g.gz;
    int b = (int)t;

    int first_pt = idx_sorted[j];
    int local_rank = scan[first_pt] - meta[b];
    if (local_rank >= max_voxels) return;  // voxel beyond the per-sample cap: skipped
    int row = meta[UD_MAX_BATCH + b] + local_rank;
    if (row >= out_capacity) return;

    float sum[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) sum[c] = 0.f;
    int cnt = 0;
    float* vrow = voxels ? voxels + (size_t)row * max_points * C : nullptr;
    for (int s = 0; s < max_points; ++s) {
        bool have = (j + s < n) && (keys_sorted[j + s] == key);
        if (have) {
            const float* p = pts + (size_t)idx_sorted[j + s] * C;
            for (int c = 0; c < C; ++c) {
                float v = p[c];
                if (c < 8) sum[c] = __fadd_rn(sum[c], v);
                if (vrow) vrow[s * C + c] = v;
            }
            ++cnt;
        } else {
            if (vrow)
                for (int c = 0; c < C; ++c) vrow[s * C + c] = 0.f;
            // keep looping only to zero-fill the padded slots
            if (!vrow) break;
        }
    }
    coords[row * 4 + 0] = b;
    coords[row * 4 + 1] = z;
    coords[row * 4 + 2] = y;
    coords[row * 4 + 3] = x;
    num_points[row] = cnt;
    if (mean) {
        float denom = (float)(cnt > 0 ? cnt : 1);
        for (int c = 0; c < C && c < 8; ++c) mean[(size_t)row * C + c] = __fdiv_rn(sum[c], denom);
    }
}

static int ud_key_bits(uint64_t max_key_inclusive) {
    int bits = 1;
    while (bits < 32 && (max_key_inclusive >> bits) != 0) ++bits;
    return bits;
}

struct UdVoxelizeWs {
    uint32_t* keys;
    uint32_t* keys_sorted;
    int* idx;
    int* idx_sorted;
    int* flag;   // n + 1
    int* scan;   // n + 1
    int* meta;   // 2 * UD_MAX_BATCH
    void* cub_temp;
    size_t cub_bytes;
    size_t total;
};

static int ud_voxelize_layout(void* ws, size_t ws_bytes, int n, UdVoxelizeWs* L) {
    UdArena a(ws, ws_bytes);
    L->keys = a.take<uint32_t>(n);
    L->keys_sorted = a.take<uint32_t>(n);
    L->idx = a.take<int>(n);
    L->idx_sorted = a.take<int>(n);
    L->flag = a.take<int>((size_t)n + 1);
    L->scan = a.take<int>((size_t)n + 1);
    L->meta = a.take<int>(2 * UD_MAX_BATCH);
    size_t sort_bytes = 0, scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int>(nullptr, sort_bytes, nullptr, nullptr,
                                                                nullptr, nullptr, n, 0, 32)));
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 n + 1)));
    L->cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    L->cub_temp = a.take<char>(L->cub_bytes);
    L->total = a.off;
    return UD_OK;
}

extern "C" {

// Bytes of scratch `ud_hard_voxelize` needs for `num_points_total` points.
int ud_voxelize_workspace_bytes(int num_points_total, size_t* bytes) {
    if (num_points_total < 0 || !bytes) return UD_ERR_INVALID_ARG;
    UdVoxelizeWs L;
    int rc = ud_voxelize_layout(nullptr, 0, num_points_total > 0 ? num_points_total : 1, &L);
    if (rc != UD_OK) return rc;
    *bytes = L.total;
    return UD_OK;
}

// points:           (N_total, C) f32 device, samples concatenated
// sample_offsets:   HOST array of batch_size+1 row offsets into `points`
// voxel_size/pc_range/grid_size: HOST arrays (x,y,z) / (xmin,ymin,zmin,xmax,ymax,zmax) / (gx,gy,gz)
// voxels (nullable): (out_capacity, max_points, C) f32   coords: (out_capacity, 4) int32 [b,z,y,x]
// num_points: (out_capacity,) int32                       mean (nullable): (out_capacity, C) f32
// counts: (1 + batch_size,) int32 device: [total, per-sample...]
int ud_hard_voxelize(const float* points, const int* sample_offsets, int batch_size, int num_features,
                     const float* voxel_size, const float* pc_range, const int* grid_size,
                     int max_points, int max_voxels, int out_capacity, float* voxels, int* coords,
                     int* num_points, float* mean, int* counts, void* workspace,
                     size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (batch_size < 1 || batch_size > UD_MAX_BATCH || num_features < 3 || num_features > 8 ||
        max_points < 1 || max_voxels < 1 || !coords || !num_points || !counts)
        return UD_ERR_INVALID_ARG;
    UdOffsets offs;
    for (int b = 0; b <= batch_size; ++b) offs.v[b] = sample_offsets[b];
    for (int b = batch_size + 1; b <= UD_MAX_BATCH; ++b) offs.v[b] = sample_offsets[batch_size];
    int n = offs.v[batch_size];
    if (n < 0 || offs.v[0] != 0) return UD_ERR_INVALID_ARG;
    for (int b = 0; b < batch_size; ++b)
        if (offs.v[b + 1] < offs.v[b]) return UD_ERR_INVALID_ARG;

    UdVoxelGeom g;
    g.vx = voxel_size[0]; g.vy = voxel_size[1]; g.vz = voxel_size[2];
    g.minx = pc_range[0]; g.miny = pc_range[1]; g.minz = pc_range[2];
    g.gx = grid_size[0]; g.gy = grid_size[1]; g.gz = grid_size[2];
    uint64_t cells = (uint64_t)batch_size * g.gx * g.gy * g.gz;
    if (g.gx < 1 || g.gy < 1 || g.gz < 1 || cells >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    uint32_t invalid_key = (uint32_t)cells;  // one past the largest valid key
    int end_bit = ud_key_bits(cells);

    if (n == 0) {
        UD_CUDA_TRY(cudaMemsetAsync(counts, 0, sizeof(int) * (1 + batch_size), stream));
        return UD_OK;
    }
    UdVoxelizeWs L;
    int rc = ud_voxelize_layout(workspace, workspace_bytes, n, &L);
    if (rc != UD_OK) return rc;
    if (!workspace || L.total > workspace_bytes) return UD_ERR_WORKSPACE;

    k_point_keys<<<ud_blocks(n, 256), 256, 0, stream>>>(points, n, num_features, offs, batch_size, g,
                                                        invalid_key, L.keys, L.idx);
    UD_LAUNCH_CHECK();
    size_t cub_bytes = L.cub_bytes;
    UD_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int>(L.cub_temp, cub_bytes, L.keys,
                                                                L.keys_sorted, L.idx, L.idx_sorted, n, 0,
                                                                end_bit, stream)));
    UD_CUDA_TRY(cudaMemsetAsync(L.flag, 0, sizeof(int) * ((size_t)n + 1), stream));
    k_segment_heads<<<ud_blocks(n, 256), 256, 0, stream>>>(L.keys_sorted, L.idx_sorted, n, invalid_key,
                                                           L.flag);
    UD_LAUNCH_CHECK();
    cub_bytes = L.cub_bytes;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(L.cub_temp, cub_bytes, L.flag, L.scan,
                                                                 n + 1, stream)));
    k_sample_counts<<<1, 32, 0, stream>>>(L.scan, offs, batch_size, max_voxels, counts, L.meta);
    UD_LAUNCH_CHECK();
    k_emit_voxels<<<ud_blocks(n, 128), 128, 0, stream>>>(points, n, num_features, L.keys_sorted,
                                                         L.idx_sorted, L.scan, L.meta, g, invalid_key,
                                                         max_points, max_voxels, out_capacity, voxels,
                                                         coords, num_points, mean);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
