/*
 * oracle/lidar_ref.c -- TEST INFRASTRUCTURE ONLY (checker; never linked into the product path).
 *
 * Plain-C restatement of the LiDAR-branch algorithms of megvii-research/CVPR2023-UniDistill whose
 * arithmetic lives in the third-party `spconv` wheel (spconv-cu111 >= 2.1.12, requirements.txt:16; not
 * vendored under /root/reference, so PARITY IS UNPINNED BY THE REFERENCE for these functions: there
 * are no golden vectors to check against; the restatement follows spconv's published CPU algorithms
 * and is pinned by hand-computable known-answer tests + dense conv3d equivalence in tests/).
 *
 *   ud_ref_voxelize    <- PointToVoxel as called by unidistill/data/det3d/preprocess/voxelization.py:31-38,54
 *                         (first-come dense-table voxelizer, spconv Point2VoxelCPU / mmdet3d hard_voxelize
 *                         CPU kernel) + MeanVFE, layers/blocks_3d/det3d/vfe/mean_vfe.py:25-32
 *   ud_ref_rulebook_*  <- get_indice_pairs implied by spconv_backbone.py:259-340
 *                         (API shape: layers/blocks_3d/det3d/centernet_backbone.py:221-251)
 *   ud_ref_sparse_conv <- indice_conv / indice_subm_conv, ConvAlgo.Native (spconv_backbone.py:7):
 *                         per kernel offset: gather rows, multiply by W[k], scatter-add
 *   ud_ref_dense       <- SparseConvTensor.dense() + view, map_to_bev/height_compression.py:19-21
 *
 * Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/build_oracle.py).  -ffp-contract=off keeps
 * the float index math free of FMAs so that voxel indices are bit-comparable with the CUDA kernels.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * First-come hard voxelization of ONE sample.
 * points (n, C); outputs sized for max_voxels: voxels (max_voxels, max_points, C) zero-filled by caller,
 * coords (max_voxels, 3) as (z, y, x), num (max_voxels).  Returns the number of voxels.
 * Rule (spconv 2.x / mmcv hard_voxelize CPU): iterate points in order; reject if any coordinate
 * c = floor((p - min) / vsize) is outside [0, grid); a new voxel gets id = running count unless the cap
 * is reached, in which case the point is skipped (`continue`); at most max_points points per voxel.
 * ---------------------------------------------------------------------------------------------- */
int ud_ref_voxelize(const float* points, int n, int C, const float* vsize, const float* range,
                    const int* grid /* x,y,z */, int max_points, int max_voxels, float* voxels,
                    int* coords, int* num, float* mean /* nullable (max_voxels, C) */) {
    int64_t cells = (int64_t)grid[0] * grid[1] * grid[2];
    int* table = (int*)malloc(sizeof(int) * (size_t)cells);
    if (!table) return -1;
    memset(table, 0xFF, sizeof(int) * (size_t)cells);
    int voxel_num = 0;
    for (int i = 0; i < n; ++i) {
        const float* p = points + (size_t)i * C;
        int c[3];
        int failed = 0;
        for (int j = 0; j < 3; ++j) {
            float f = floorf((p[j] - range[j]) / vsize[j]);
            if (!(f >= 0.0f && f < (float)grid[j])) { failed = 1; break; }
            c[j] = (int)f;
        }
        if (failed) continue;
        int64_t lin = ((int64_t)c[2] * grid[1] + c[1]) * grid[0] + c[0];
        int vid = table[lin];
        if (vid == -1) {
            if (voxel_num >= max_voxels) continue;
            vid = voxel_num++;
            table[lin] = vid;
            coords[vid * 3 + 0] = c[2];
            coords[vid * 3 + 1] = c[1];
            coords[vid * 3 + 2] = c[0];
            num[vid] = 0;
        }
        if (num[vid] < max_points) {
            memcpy(voxels + ((size_t)vid * max_points + num[vid]) * C, p, sizeof(float) * C);
            num[vid] += 1;
        }
    }
    free(table);
    if (mean) {
        for (int v = 0; v < voxel_num; ++v) {
            float denom = (float)(num[v] > 0 ? num[v] : 1);
            for (int ch = 0; ch < C; ++ch) {
                float s = 0.0f;
                for (int t = 0; t < num[v]; ++t) s = s + voxels[((size_t)v * max_points + t) * C + ch];
                mean[(size_t)v * C + ch] = s / denom;
            }
        }
    }
    return voxel_num;
}

/* ------------------------------------------------------------------------------------------------
 * Coordinate hash (open addressing) used by the rulebook builders.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    int64_t* keys;
    int* vals;
    uint64_t mask;
} ud_map;

static uint64_t ud_mix(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

static int ud_map_init(ud_map* m, int n) {
    uint64_t cap = 1024;
    while (cap < (uint64_t)n * 2 + 2) cap <<= 1;
    m->keys = (int64_t*)malloc(sizeof(int64_t) * cap);
    m->vals = (int*)malloc(sizeof(int) * cap);
    if (!m->keys || !m->vals) return -1;
    for (uint64_t i = 0; i < cap; ++i) m->keys[i] = -1;
    m->mask = cap - 1;
    return 0;
}

static void ud_map_free(ud_map* m) { free(m->keys); free(m->vals); }

/* returns existing value or inserts val and returns it */
static int ud_map_get_or_put(ud_map* m, int64_t key, int val) {
    uint64_t s = ud_mix((uint64_t)key) & m->mask;
    while (m->keys[s] != -1) {
        if (m->keys[s] == key) return m->vals[s];
        s = (s + 1) & m->mask;
    }
    m->keys[s] = key;
    m->vals[s] = val;
    return val;
}

static int ud_map_find(const ud_map* m, int64_t key) {
    uint64_t s = ud_mix((uint64_t)key) & m->mask;
    while (m->keys[s] != -1) {
        if (m->keys[s] == key) return m->vals[s];
        s = (s + 1) & m->mask;
    }
    return -1;
}

static int64_t ud_lin(int b, int z, int y, int x, const int* shape) {
    return (((int64_t)b * shape[0] + z) * shape[1] + y) * shape[2] + x;
}

/* ------------------------------------------------------------------------------------------------
 * Submanifold rulebook.  coords (n,4) [b,z,y,x].  pairs (2,K,n) -1 padded, pair_num (K).
 * pair (i -> o, k) iff coord_i = coord_o + (k - center) * dilation; per k the pairs are appended while
 * iterating the INPUT rows in order (spconv CPU order).
 * ---------------------------------------------------------------------------------------------- */
int ud_ref_rulebook_subm(const int* coords, int n, const int* shape, const int* ksize, const int* dil,
                         int* pairs, int* pair_num) {
    int K = ksize[0] * ksize[1] * ksize[2];
    ud_map m;
    if (ud_map_init(&m, n)) return -1;
    for (int i = 0; i < n; ++i) {
        const int* c = coords + i * 4;
        ud_map_get_or_put(&m, ud_lin(c[0], c[1], c[2], c[3], shape), i);
    }
    for (int k = 0; k < K; ++k) pair_num[k] = 0;
    for (size_t t = 0; t < (size_t)2 * K * n; ++t) pairs[t] = -1;
    for (int i = 0; i < n; ++i) {
        const int* c = coords + i * 4;
        for (int k = 0; k < K; ++k) {
            int kx = k % ksize[2], ky = (k / ksize[2]) % ksize[1], kz = k / (ksize[2] * ksize[1]);
            /* output o with coord_o = coord_i - (k - center)*dil */
            int z = c[1] - (kz - ksize[0] / 2) * dil[0];
            int y = c[2] - (ky - ksize[1] / 2) * dil[1];
            int x = c[3] - (kx - ksize[2] / 2) * dil[2];
            if (z < 0 || z >= shape[0] || y < 0 || y >= shape[1] || x < 0 || x >= shape[2]) continue;
            int o = ud_map_find(&m, ud_lin(c[0], z, y, x, shape));
            if (o < 0) continue;
            int pos = pair_num[k]++;
            pairs[(size_t)k * n + pos] = i;
            pairs[(size_t)(K + k) * n + pos] = o;
        }
    }
    ud_map_free(&m);
    return 0;
}

static int ud_cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}

/* ------------------------------------------------------------------------------------------------
 * Strided sparse conv rulebook.  out = (in + pad - k*dil) / stride when divisible and in bounds.
 * Output rows are ordered ascending by linear (b,z,y,x) index (canonical order).
 * out_coords must hold n * K rows in the worst case.  Returns n_out.
 * ---------------------------------------------------------------------------------------------- */
int ud_ref_rulebook_conv(const int* coords, int n, const int* in_shape, const int* out_shape,
                         const int* ksize, const int* stride, const int* pad, const int* dil,
                         int* out_coords, int* pairs, int* pair_num) {
    int K = ksize[0] * ksize[1] * ksize[2];
    (void)in_shape;
    int64_t* cand = (int64_t*)malloc(sizeof(int64_t) * (size_t)n * K + 8);
    size_t nc = 0;
    for (int i = 0; i < n; ++i) {
        const int* c = coords + i * 4;
        for (int k = 0; k < K; ++k) {
            int kk[3] = {k / (ksize[2] * ksize[1]), (k / ksize[2]) % ksize[1], k % ksize[2]};
            int o[3], ok = 1;
            for (int d = 0; d < 3; ++d) {
                int num = c[1 + d] + pad[d] - kk[d] * dil[d];
                if (num < 0 || num % stride[d] != 0) { ok = 0; break; }
                o[d] = num / stride[d];
                if (o[d] >= out_shape[d]) { ok = 0; break; }
            }
            if (ok) cand[nc++] = ud_lin(c[0], o[0], o[1], o[2], out_shape);
        }
    }
    qsort(cand, nc, sizeof(int64_t), ud_cmp_i64);
    size_t n_out = 0;
    for (size_t t = 0; t < nc; ++t)
        if (t == 0 || cand[t] != cand[t - 1]) cand[n_out++] = cand[t];
    ud_map m;
    if (ud_map_init(&m, (int)n_out)) return -1;
    for (size_t o = 0; o < n_out; ++o) {
        int64_t key = cand[o];
        ud_map_get_or_put(&m, key, (int)o);
        int x = (int)(key % out_shape[2]); key /= out_shape[2];
        int y = (int)(key % out_shape[1]); key /= out_shape[1];
        int z = (int)(key % out_shape[0]); key /= out_shape[0];
        out_coords[o * 4 + 0] = (int)key;
        out_coords[o * 4 + 1] = z;
        out_coords[o * 4 + 2] = y;
        out_coords[o * 4 + 3] = x;
    }
    for (int k = 0; k < K; ++k) pair_num[k] = 0;
    for (size_t t = 0; t < (size_t)2 * K * n; ++t) pairs[t] = -1;
    for (int i = 0; i < n; ++i) {
        const int* c = coords + i * 4;
        for (int k = 0; k < K; ++k) {
            int kk[3] = {k / (ksize[2] * ksize[1]), (k / ksize[2]) % ksize[1], k % ksize[2]};
            int o[3], ok = 1;
            for (int d = 0; d < 3; ++d) {
                int num = c[1 + d] + pad[d] - kk[d] * dil[d];
                if (num < 0 || num % stride[d] != 0) { ok = 0; break; }
                o[d] = num / stride[d];
                if (o[d] >= out_shape[d]) { ok = 0; break; }
            }
            if (!ok) continue;
            int oi = ud_map_find(&m, ud_lin(c[0], o[0], o[1], o[2], out_shape));
            int pos = pair_num[k]++;
            pairs[(size_t)k * n + pos] = i;
            pairs[(size_t)(K + k) * n + pos] = oi;
        }
    }
    ud_map_free(&m);
    free(cand);
    return (int)n_out;
}

/* ------------------------------------------------------------------------------------------------
 * Native sparse conv on a pair list: for k, for pair (i,o): out[o,:] += in[i,:] @ w[k]   (+ bias first)
 * in (n_in,cin), w (K,cin,cout), out (n_out,cout).
 * ---------------------------------------------------------------------------------------------- */
void ud_ref_sparse_conv(const float* in, int n_in, const float* w, const float* bias, const int* pairs,
                        const int* pair_num, int K, int cin, int cout, float* out, int n_out) {
    for (int o = 0; o < n_out; ++o)
        for (int c = 0; c < cout; ++c) out[(size_t)o * cout + c] = bias ? bias[c] : 0.0f;
    for (int k = 0; k < K; ++k) {
        const float* wk = w + (size_t)k * cin * cout;
        for (int p = 0; p < pair_num[k]; ++p) {
            int i = pairs[(size_t)k * n_in + p];
            int o = pairs[(size_t)(K + k) * n_in + p];
            const float* x = in + (size_t)i * cin;
            float* y = out + (size_t)o * cout;
            for (int ci = 0; ci < cin; ++ci) {
                float xv = x[ci];
                const float* wr = wk + (size_t)ci * cout;
                for (int co = 0; co < cout; ++co) y[co] += xv * wr[co];
            }
        }
    }
}

/* dense(): out (B, C, D, H, W) zero-filled by the caller; feat (n, C); coords (n,4) [b,z,y,x] */
void ud_ref_dense(const float* feat, const int* coords, int n, int C, int D, int H, int W, float* out) {
    for (int o = 0; o < n; ++o) {
        const int* c = coords + o * 4;
        for (int ch = 0; ch < C; ++ch)
            out[((((size_t)c[0] * C + ch) * D + c[1]) * H + c[2]) * W + c[3]] = feat[(size_t)o * C + ch];
    }
}
