// Rulebook (kernel map) construction for submanifold and strided sparse 3-D convolutions.
//
// Replaces spconv's `get_indice_pairs` that every spconv.SubMConv3d / SparseConv3d of
//   reference unidistill/layers/blocks_3d/det3d/spconv_backbone.py:259-340
// builds implicitly (the operator API is visible at
//   reference unidistill/layers/blocks_3d/det3d/centernet_backbone.py:221-251).
//
// Conventions (cross-correlation, identical to a dense conv3d on the occupancy grid):
//   kernel offset index k = (kz*KY + ky)*KX + kx
//   pair (i -> o, k) exists iff  in_coord = out_coord*stride - pad + k*dilation   (per axis)
//   SubM: out set == in set (same row order).
//   Strided: out set = all in-bounds out_coords with >= 1 pair, rows sorted ascending by the
//            linear (b,z,y,x) index (canonical order; obtained from a rank/select bitmap,
//            so no sort and no atomics on the ordering => deterministic).
//
// Two table forms are produced, both int32 with -1 == absent:
//   nbr_out[o][k] = i   (output-major gather table; forward + wgrad)
//   nbr_in [i][k] = o   (input-major table; dgrad)   -- SubM: nbr_in[i][k] = nbr_out[i][K-1-k]
// plus, on request, the spconv-shaped pair list (2,K,N_in) + indice_pair_num (K).
#include "common.cuh"
#include <cub/cub.cuh>

#define UD_EMPTY64 0xFFFFFFFFFFFFFFFFull

struct UdConvGeom {
    int B;
    int iD, iH, iW;  // input spatial shape (z, y, x)
    int oD, oH, oW;  // output spatial shape
    int kD, kH, kW;
    int sD, sH, sW;
    int pD, pH, pW;
    int dD, dH, dW;
};

__device__ __forceinline__ void ud_hash_insert(unsigned long long* table, uint32_t mask, uint32_t key,
                                               int val) {
    uint32_t slot = ud_hash32(key) & mask;
    unsigned long long packed = ((unsigned long long)key << 32) | (uint32_t)val;
    while (true) {
        unsigned long long old = atomicCAS(&table[slot], UD_EMPTY64, packed);
        if (old == UD_EMPTY64) return;
        if ((uint32_t)(old >> 32) == key) return;  // duplicate coordinate: first writer wins
        slot = (slot + 1) & mask;
    }
}

__device__ __forceinline__ int ud_hash_lookup(const unsigned long long* __restrict__ table, uint32_t mask,
                                              uint32_t key) {
    uint32_t slot = ud_hash32(key) & mask;
    while (true) {
        unsigned long long v = table[slot];
        if (v == UD_EMPTY64) return -1;
        if ((uint32_t)(v >> 32) == key) return (int)(uint32_t)(v & 0xFFFFFFFFull);
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(256)
k_hash_build(const int4* __restrict__ coords, int n, int D, int H, int W, unsigned long long* table,
             uint32_t mask) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int4 c = coords[i];  // (b, z, y, x)
    uint32_t key = (((uint32_t)c.x * D + c.y) * H + c.z) * W + c.w;
    ud_hash_insert(table, mask, key, i);
}

__global__ void __launch_bounds__(256)
k_subm_table(const int4* __restrict__ coords, int n, UdConvGeom g, const unsigned long long* __restrict__ table,
             uint32_t mask, int* __restrict__ nbr) {
    int K = g.kD * g.kH * g.kW;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * K) return;
    int o = (int)(t / K);
    int k = (int)(t - (long long)o * K);
    int kx = k % g.kW;
    int ky = (k / g.kW) % g.kH;
    int kz = k / (g.kW * g.kH);
    int4 c = coords[o];
    int z = c.y + (kz - g.kD / 2) * g.dD;
    int y = c.z + (ky - g.kH / 2) * g.dH;
    int x = c.w + (kx - g.kW / 2) * g.dW;
    int r = -1;
    if (z >= 0 && z < g.iD && y >= 0 && y < g.iH && x >= 0 && x < g.iW) {
        uint32_t key = (((uint32_t)c.x * g.iD + z) * g.iH + y) * g.iW + x;
        r = ud_hash_lookup(table, mask, key);
    }
    nbr[t] = r;
}

// out coordinate along one axis for input coordinate `in` and kernel tap `k`; -1 if none.
__device__ __forceinline__ int ud_out_coord(int in, int k, int stride, int pad, int dil, int out_dim) {
    int num = in + pad - k * dil;
    if (num < 0) return -1;
    if (num % stride != 0) return -1;
    int o = num / stride;
    return o < out_dim ? o : -1;
}

__device__ __forceinline__ bool ud_pair_key(const int4 c, int k, const UdConvGeom& g, uint32_t* key) {
    int kx = k % g.kW;
    int ky = (k / g.kW) % g.kH;
    int kz = k / (g.kW * g.kH);
    int oz = ud_out_coord(c.y, kz, g.sD, g.pD, g.dD, g.oD);
    int oy = ud_out_coord(c.z, ky, g.sH, g.pH, g.dH, g.oH);
    int ox = ud_out_coord(c.w, kx, g.sW, g.pW, g.dW, g.oW);
    if ((oz | oy | ox) < 0) return false;
    *key = (((uint32_t)c.x * g.oD + oz) * g.oH + oy) * g.oW + ox;
    return true;
}

__global__ void __launch_bounds__(256)
k_conv_mark(const int4* __restrict__ coords, int n, UdConvGeom g, unsigned int* __restrict__ bitmap) {
    int K = g.kD * g.kH * g.kW;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * K) return;
    int i = (int)(t / K);
    int k = (int)(t - (long long)i * K);
    uint32_t key;
    if (ud_pair_key(coords[i], k, g, &key)) atomicOr(&bitmap[key >> 5], 1u << (key & 31));
}

__global__ void __launch_bounds__(256)
k_popc(const unsigned int* __restrict__ bitmap, int words, int* __restrict__ pc) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < words) pc[w] = __popc(bitmap[w]);
    if (w == words) pc[w] = 0;  // sentinel so that scan[words] == total
}

__global__ void k_copy_count(const int* __restrict__ prefix, int words, int* __restrict__ n_out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) *n_out = prefix[words];
}

__global__ void __launch_bounds__(256)
k_conv_out_coords(const unsigned int* __restrict__ bitmap, const int* __restrict__ prefix, int words,
                  UdConvGeom g, int n_out, int4* __restrict__ out_coords) {
    int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= words) return;
    unsigned int bits = bitmap[w];
    int o = prefix[w];
    while (bits) {
        int bpos = __ffs(bits) - 1;
        bits &= bits - 1;
        uint32_t key = ((uint32_t)w << 5) | (uint32_t)bpos;
        uint32_t t = key;
        int x = t % g.oW; t /= g.oW;
        int y = t % g.oH; t /= g.oH;
        int z = t % g.oD; t /= g.oD;
        if (o < n_out) out_coords[o] = make_int4((int)t, z, y, x);
        ++o;
    }
}

__global__ void __launch_bounds__(256)
k_conv_tables(const int4* __restrict__ coords, int n, UdConvGeom g, const unsigned int* __restrict__ bitmap,
              const int* __restrict__ prefix, int n_out, int* __restrict__ nbr_in, int* __restrict__ nbr_out) {
    int K = g.kD * g.kH * g.kW;
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n * K) return;
    int i = (int)(t / K);
    int k = (int)(t - (long long)i * K);
    uint32_t key;
    int o = -1;
    if (ud_pair_key(coords[i], k, g, &key)) {
        unsigned int word = bitmap[key >> 5];
        unsigned int below = word & ((1u << (key & 31)) - 1u);
        o = prefix[key >> 5] + __popc(below);
        if (o < n_out) nbr_out[(size_t)o * K + k] = i; else o = -1;
    }
    nbr_in[t] = o;
}

// ---- spconv-shaped pair list export ---------------------------------------------------------
__global__ void __launch_bounds__(256)
k_pair_flags(const int* __restrict__ nbr_in, int n_in, int K, int mirror, int* __restrict__ flags) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_in * K;
    if (t > total) return;
    if (t == total) { flags[t] = 0; return; }
    int k = (int)(t / n_in);
    int i = (int)(t - (long long)k * n_in);
    int kk = mirror ? (K - 1 - k) : k;
    flags[t] = nbr_in[(size_t)i * K + kk] >= 0 ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_pair_emit(const int* __restrict__ nbr_in, int n_in, int K, int mirror, const int* __restrict__ scan,
            int* __restrict__ pairs, int* __restrict__ pair_num) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long total = (long long)n_in * K;
    if (t >= total) return;
    int k = (int)(t / n_in);
    int i = (int)(t - (long long)k * n_in);
    int kk = mirror ? (K - 1 - k) : k;
    int o = nbr_in[(size_t)i * K + kk];
    int base = scan[(long long)k * n_in];
    if (o >= 0) {
        int pos = scan[t] - base;
        pairs[(size_t)k * n_in + pos] = i;
        pairs[(size_t)(K + k) * n_in + pos] = o;
    }
    if (i == 0) pair_num[k] = scan[(long long)(k + 1) * n_in] - base;
}

static uint32_t ud_pow2_at_least(uint64_t v) {
    uint64_t p = 1024;
    while (p < v) p <<= 1;
    return (uint32_t)p;
}

static int ud_geom_check(const UdConvGeom& g, bool in_side) {
    uint64_t cells = (uint64_t)g.B * (in_side ? (uint64_t)g.iD * g.iH * g.iW : (uint64_t)g.oD * g.oH * g.oW);
    if (g.B < 1 || cells == 0 || cells >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    int K = g.kD * g.kH * g.kW;
    if (K < 1 || K > 32) return UD_ERR_UNSUPPORTED;
    return UD_OK;
}

static UdConvGeom ud_make_geom(int B, const int* in_shape, const int* out_shape, const int* ksize,
                               const int* stride, const int* padding, const int* dilation) {
    UdConvGeom g;
    g.B = B;
    g.iD = in_shape[0]; g.iH = in_shape[1]; g.iW = in_shape[2];
    g.oD = out_shape[0]; g.oH = out_shape[1]; g.oW = out_shape[2];
    g.kD = ksize[0]; g.kH = ksize[1]; g.kW = ksize[2];
    g.sD = stride[0]; g.sH = stride[1]; g.sW = stride[2];
    g.pD = padding[0]; g.pH = padding[1]; g.pW = padding[2];
    g.dD = dilation[0]; g.dH = dilation[1]; g.dW = dilation[2];
    return g;
}

extern "C" {

// ---- submanifold ------------------------------------------------------------------------------
int ud_rulebook_subm_workspace_bytes(int n, size_t* bytes) {
    if (n < 0 || !bytes) return UD_ERR_INVALID_ARG;
    *bytes = ud_align_up((size_t)ud_pow2_at_least(2ull * (uint64_t)(n > 0 ? n : 1)) * sizeof(unsigned long long));
    return UD_OK;
}

// coords: (n,4) int32 [b,z,y,x] device; spatial_shape/ksize/dilation: HOST int[3] (z,y,x)
// nbr_out: (n, K) int32 device, K = kz*ky*kx (odd sizes only).
int ud_rulebook_subm(const int* coords, int n, int batch_size, const int* spatial_shape, const int* ksize,
                     const int* dilation, int* nbr_out, void* workspace, size_t workspace_bytes,
                     void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || (n > 0 && (!coords || !nbr_out))) return UD_ERR_INVALID_ARG;
    int ones[3] = {1, 1, 1}, zeros[3] = {0, 0, 0};
    UdConvGeom g = ud_make_geom(batch_size, spatial_shape, spatial_shape, ksize, ones, zeros, dilation);
    if ((g.kD & 1) == 0 || (g.kH & 1) == 0 || (g.kW & 1) == 0) return UD_ERR_UNSUPPORTED;
    int rc = ud_geom_check(g, true);
    if (rc != UD_OK) return rc;
    if (n == 0) return UD_OK;
    uint32_t cap = ud_pow2_at_least(2ull * (uint64_t)n);
    if (!workspace || workspace_bytes < (size_t)cap * sizeof(unsigned long long)) return UD_ERR_WORKSPACE;
    unsigned long long* table = (unsigned long long*)workspace;
    UD_CUDA_TRY(cudaMemsetAsync(table, 0xFF, (size_t)cap * sizeof(unsigned long long), stream));
    k_hash_build<<<ud_blocks(n, 256), 256, 0, stream>>>((const int4*)coords, n, g.iD, g.iH, g.iW, table,
                                                        cap - 1);
    UD_LAUNCH_CHECK();
    int K = g.kD * g.kH * g.kW;
    k_subm_table<<<ud_blocks((long long)n * K, 256), 256, 0, stream>>>((const int4*)coords, n, g, table,
                                                                       cap - 1, nbr_out);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// ---- strided / regular sparse conv --------------------------------------------------------------
static size_t ud_conv_words(const UdConvGeom& g) {
    uint64_t cells = (uint64_t)g.B * g.oD * g.oH * g.oW;
    return (size_t)((cells + 31) / 32);
}

int ud_rulebook_conv_workspace_bytes(int batch_size, const int* out_shape, size_t* bytes) {
    if (!out_shape || !bytes || batch_size < 1) return UD_ERR_INVALID_ARG;
    uint64_t cells = (uint64_t)batch_size * out_shape[0] * out_shape[1] * out_shape[2];
    if (cells == 0 || cells >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    size_t words = (size_t)((cells + 31) / 32);
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 (int)words + 1)));
    *bytes = ud_align_up(words * 4) + 2 * ud_align_up((words + 1) * 4) + ud_align_up(scan_bytes);
    return UD_OK;
}

struct UdConvWs {
    unsigned int* bitmap;
    int* pc;
    int* prefix;
    void* cub_temp;
    size_t cub_bytes;
    size_t words;
};

static int ud_conv_layout(void* ws, size_t ws_bytes, const UdConvGeom& g, UdConvWs* L) {
    L->words = ud_conv_words(g);
    UdArena a(ws, ws_bytes);
    L->bitmap = a.take<unsigned int>(L->words);
    L->pc = a.take<int>(L->words + 1);
    L->prefix = a.take<int>(L->words + 1);
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 (int)L->words + 1)));
    L->cub_bytes = scan_bytes;
    L->cub_temp = a.take<char>(scan_bytes);
    if (!a.ok()) return UD_ERR_WORKSPACE;
    return UD_OK;
}

// Phase 1: mark the output occupancy bitmap, rank it, write the number of output rows to
// n_out_dev (device int).  The workspace must be kept untouched until phase 2 ran.
int ud_rulebook_conv_count(const int* coords, int n_in, int batch_size, const int* in_shape,
                           const int* out_shape, const int* ksize, const int* stride, const int* padding,
                           const int* dilation, int* n_out_dev, void* workspace, size_t workspace_bytes,
                           void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!n_out_dev || n_in < 0 || (n_in > 0 && !coords)) return UD_ERR_INVALID_ARG;
    UdConvGeom g = ud_make_geom(batch_size, in_shape, out_shape, ksize, stride, padding, dilation);
    int rc = ud_geom_check(g, false);
    if (rc != UD_OK) return rc;
    UdConvWs L;
    rc = ud_conv_layout(workspace, workspace_bytes, g, &L);
    if (rc != UD_OK) return rc;
    int K = g.kD * g.kH * g.kW;
    int words = (int)L.words;
    UD_CUDA_TRY(cudaMemsetAsync(L.bitmap, 0, L.words * 4, stream));
    if (n_in > 0) {
        k_conv_mark<<<ud_blocks((long long)n_in * K, 256), 256, 0, stream>>>((const int4*)coords, n_in, g,
                                                                             L.bitmap);
        UD_LAUNCH_CHECK();
    }
    k_popc<<<ud_blocks(words + 1, 256), 256, 0, stream>>>(L.bitmap, words, L.pc);
    UD_LAUNCH_CHECK();
    size_t cub_bytes = L.cub_bytes;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(L.cub_temp, cub_bytes, L.pc, L.prefix,
                                                                 words + 1, stream)));
    k_copy_count<<<1, 32, 0, stream>>>(L.prefix, words, n_out_dev);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Phase 2: with n_out known on the host, write out_coords (n_out,4), nbr_in (n_in,K), nbr_out (n_out,K).
int ud_rulebook_conv_fill(const int* coords, int n_in, int batch_size, const int* in_shape,
                          const int* out_shape, const int* ksize, const int* stride, const int* padding,
                          const int* dilation, int n_out, int* out_coords, int* nbr_in, int* nbr_out,
                          void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_in < 0 || n_out < 0) return UD_ERR_INVALID_ARG;
    if (n_in > 0 && (!coords || !nbr_in)) return UD_ERR_INVALID_ARG;
    if (n_out > 0 && (!out_coords || !nbr_out)) return UD_ERR_INVALID_ARG;
    UdConvGeom g = ud_make_geom(batch_size, in_shape, out_shape, ksize, stride, padding, dilation);
    int rc = ud_geom_check(g, false);
    if (rc != UD_OK) return rc;
    UdConvWs L;
    rc = ud_conv_layout(workspace, workspace_bytes, g, &L);
    if (rc != UD_OK) return rc;
    int K = g.kD * g.kH * g.kW;
    int words = (int)L.words;
    if (n_out > 0) {
        UD_CUDA_TRY(cudaMemsetAsync(nbr_out, 0xFF, (size_t)n_out * K * sizeof(int), stream));
        k_conv_out_coords<<<ud_blocks(words, 256), 256, 0, stream>>>(L.bitmap, L.prefix, words, g, n_out,
                                                                     (int4*)out_coords);
        UD_LAUNCH_CHECK();
    }
    if (n_in > 0) {
        k_conv_tables<<<ud_blocks((long long)n_in * K, 256), 256, 0, stream>>>(
            (const int4*)coords, n_in, g, L.bitmap, L.prefix, n_out, nbr_in, nbr_out);
        UD_LAUNCH_CHECK();
    }
    return UD_OK;
}

// ---- spconv-shaped export ---------------------------------------------------------------------
int ud_rulebook_pairs_workspace_bytes(int n_in, int K, size_t* bytes) {
    if (n_in < 0 || K < 1 || !bytes) return UD_ERR_INVALID_ARG;
    size_t items = (size_t)n_in * K + 1;
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 (int)items)));
    *bytes = 2 * ud_align_up(items * 4) + ud_align_up(scan_bytes);
    return UD_OK;
}

// nbr_in: (n_in,K) input-major table (or, with mirror=1, a SubM output-major table).
// pairs: (2,K,n_in) int32, -1 padded; pair_num: (K,) int32.  Pairs of one k are ordered by
// ascending input row (the order spconv's CPU rulebook builder produces).
int ud_rulebook_pairs(const int* nbr_in, int n_in, int K, int mirror, int* pairs, int* pair_num,
                      void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!pair_num || n_in < 0 || K < 1 || (n_in > 0 && (!nbr_in || !pairs))) return UD_ERR_INVALID_ARG;
    if ((long long)n_in * K + 1 > 0x7FFFFFFFll) return UD_ERR_UNSUPPORTED;
    UD_CUDA_TRY(cudaMemsetAsync(pair_num, 0, sizeof(int) * K, stream));
    if (n_in == 0) return UD_OK;
    size_t items = (size_t)n_in * K + 1;
    UdArena a(workspace, workspace_bytes);
    int* flags = a.take<int>(items);
    int* scan = a.take<int>(items);
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 (int)items)));
    void* cub_temp = a.take<char>(scan_bytes);
    if (!a.ok()) return UD_ERR_WORKSPACE;
    UD_CUDA_TRY(cudaMemsetAsync(pairs, 0xFF, (size_t)2 * K * n_in * sizeof(int), stream));
    k_pair_flags<<<ud_blocks((long long)items, 256), 256, 0, stream>>>(nbr_in, n_in, K, mirror, flags);
    UD_LAUNCH_CHECK();
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(cub_temp, scan_bytes, flags, scan,
                                                                 (int)items, stream)));
    k_pair_emit<<<ud_blocks((long long)n_in * K, 256), 256, 0, stream>>>(nbr_in, n_in, K, mirror, scan,
                                                                         pairs, pair_num);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
