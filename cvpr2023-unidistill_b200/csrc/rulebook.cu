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

// Output-row order of a strided convolution = ascending key over the output cells:
//   order 0 (canonical, what the rulebook parity tests pin): key = linear (b, z, y, x) index;
//   order 1 (Morton):  key = b << bits | interleave(z, y, x) -- a 128-row tile of the next layer is then a compact 3-D
//   block whose 3x3x3 neighbourhood touches ~290 distinct rows instead of ~770 (profiles/r01_tile_locality_probe.txt), which
//   the gather of the convolution kernel turns into L1 / L2 hits.  The row order of a sparse tensor between the
//   voxelizer and the BEV scatter is not observable (the scatter is order independent).
// The interleave handles unequal axis sizes: bit positions are dealt out z, y, x round-robin from the least significant
// bit while an axis still has bits left (per-axis spread tables, built once per call by k_spread_tables).
struct UdKeyLayout {
    int order;          // 0 linear, 1 Morton
    int bits_total;     // Morton: bits of one sample's key
    int bz, by, bx;     // Morton: bits per axis
    unsigned char pos[3][12];   // Morton: pos[axis][i] = key bit of coordinate bit i
};

struct UdConvGeom {
    UdKeyLayout kl;
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

// One block = 32 consecutive rows x all K taps, staged in shared memory so that BOTH table forms are written coalesced:
//   nbr   (n, K)   row-major  (API form: pair export, fp32 SIMT kernels)            -- optional
//   nbr_t (K4, ld) tap-major  (what the tensor-core kernels read: the 128 indices of a (tile, tap) are one 512-byte
//                  segment; K4 = K rounded up to 4, ld = n rounded up to 128, padding -1)  -- optional
// plus tile_masks[row / 128] |= 1 << k for every tap that occurs in the 128-row tile (zeroed by the launcher).  This is the
// round-1 builder and its separate transpose kernel (0.8 ms per step) in one pass.
// ncu (profiles/r03_rulebook_ncu.md) showed the round-2 form of this kernel at 78 % SM throughput with 277 warp instructions per
// row: the per-(row, tap) integer divisions by the run-time kernel size and the three multiplies of the key, not memory, bound
// it.  Now a block decodes the K taps ONCE into (dz, dy, dx) and the key delta dz*H*W + dy*W + dx (shared memory), and a
// (row, tap) thread does three range checks, one add and the hash probe.
__global__ void __launch_bounds__(256)
k_subm_table(const int4* __restrict__ coords, int n, UdConvGeom g, const unsigned long long* __restrict__ table,
             uint32_t mask, int ld, int* __restrict__ nbr, int* __restrict__ nbr_t, unsigned int* __restrict__ tile_masks) {
    __shared__ int tile[32][33];
    __shared__ int4 tap_s[32];                         // (dz, dy, dx, key delta) of tap k
    const int K = g.kD * g.kH * g.kW;
    const int r0 = blockIdx.x * 32;
    if (threadIdx.x < 32) {
        const int k = threadIdx.x;
        int4 t = make_int4(0, 0, 0, 0);
        if (k < K) {
            t.x = (k / (g.kW * g.kH) - g.kD / 2) * g.dD;
            t.y = ((k / g.kW) % g.kH - g.kH / 2) * g.dH;
            t.z = (k % g.kW - g.kW / 2) * g.dW;
            t.w = (t.x * g.iH + t.y) * g.iW + t.z;
        }
        tap_s[k] = t;
    }
    __syncthreads();
    const int k = threadIdx.x & 31;                    // a warp = one row, lanes = taps: the 27 probes of a row together
    const int4 tp = tap_s[k];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int rl = (threadIdx.x >> 5) + 8 * it;
        const int o = r0 + rl;
        int r = -1;
        if (o < n && k < K) {
            const int4 c = __ldg(coords + o);          // same address for the whole warp: one broadcast load
            const int z = c.y + tp.x, y = c.z + tp.y, x = c.w + tp.z;
            if ((unsigned)z < (unsigned)g.iD && (unsigned)y < (unsigned)g.iH && (unsigned)x < (unsigned)g.iW) {
                const uint32_t key = (((uint32_t)c.x * g.iD + c.y) * g.iH + c.z) * g.iW + c.w + (uint32_t)tp.w;
                r = ud_hash_lookup(table, mask, key);
            }
        }
        tile[rl][k] = r;
    }
    __syncthreads();
    if (nbr != nullptr) {                              // 32 rows x K ints are contiguous in the row-major table
        const int rows = min(32, n - r0);
        for (int e = threadIdx.x; e < rows * K; e += 256) nbr[(size_t)r0 * K + e] = tile[e / K][e % K];
    }
    if (nbr_t != nullptr) {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // lanes = rows, warps = taps
        const int k_pad = (K + 3) / 4 * 4;
        unsigned int m = 0u;
        for (int kk = ty; kk < k_pad; kk += 8) {
            const int v = kk < K ? tile[tx][kk] : -1;
            if (r0 + tx < ld) nbr_t[(size_t)kk * ld + r0 + tx] = v;
            if (__ballot_sync(0xffffffffu, r0 + tx < n && v >= 0)) m |= 1u << kk;
        }
        if (tile_masks != nullptr && tx == 0 && m) atomicOr(tile_masks + (r0 >> 7), m);
    }
}

// out coordinate along one axis for input coordinate `in` and kernel tap `k`; -1 if none.
__device__ __forceinline__ int ud_out_coord(int in, int k, int stride, int pad, int dil, int out_dim) {
    int num = in + pad - k * dil;
    if (num < 0) return -1;
    int o;
    if (stride == 1) o = num;
    else if (stride == 2) { if (num & 1) return -1; o = num >> 1; }      // the backbone's strides: no integer division
    else { if (num % stride != 0) return -1; o = num / stride; }
    return o < out_dim ? o : -1;
}

__device__ __forceinline__ uint32_t ud_spread(const UdKeyLayout& kl, int axis, int v) {
    uint32_t r = 0u;
    const int nb = axis == 0 ? kl.bz : (axis == 1 ? kl.by : kl.bx);
    for (int i = 0; i < nb; ++i) r |= (uint32_t)((v >> i) & 1) << kl.pos[axis][i];
    return r;
}

// spread tables: tab[0 .. oD) z, tab[oD .. oD+oH) y, then x
__global__ void k_spread_tables(UdKeyLayout kl, int oD, int oH, int oW, uint32_t* __restrict__ tab) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < oD) tab[t] = ud_spread(kl, 0, t);
    else if (t < oD + oH) tab[t] = ud_spread(kl, 1, t - oD);
    else if (t < oD + oH + oW) tab[t] = ud_spread(kl, 2, t - oD - oH);
}

__device__ __forceinline__ uint32_t ud_out_key(const UdConvGeom& g, const uint32_t* __restrict__ tab, int b, int oz,
                                               int oy, int ox) {
    if (g.kl.order == 0) return (((uint32_t)b * g.oD + oz) * g.oH + oy) * g.oW + ox;
    return ((uint32_t)b << g.kl.bits_total) | __ldg(tab + oz) | __ldg(tab + g.oD + oy) | __ldg(tab + g.oD + g.oH + ox);
}

// One thread per input row walks the taps axis by axis and only descends where the axis has an output coordinate (a
// stride-2 3x3x3 kernel has 1 or 2 of 3 per axis: <= 8 of 27 taps).  The round-2 form -- one thread per (row, tap), three
// divisions each -- ran at 81 % SM throughput, i.e. was bound by its integer arithmetic (profiles/r03_rulebook_ncu.md).
__global__ void __launch_bounds__(256)
k_conv_mark(const int4* __restrict__ coords, int n, UdConvGeom g, const uint32_t* __restrict__ tab,
            unsigned int* __restrict__ bitmap) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coords[i];
    for (int kz = 0; kz < g.kD; ++kz) {
        const int oz = ud_out_coord(c.y, kz, g.sD, g.pD, g.dD, g.oD);
        if (oz < 0) continue;
        for (int ky = 0; ky < g.kH; ++ky) {
            const int oy = ud_out_coord(c.z, ky, g.sH, g.pH, g.dH, g.oH);
            if (oy < 0) continue;
            for (int kx = 0; kx < g.kW; ++kx) {
                const int ox = ud_out_coord(c.w, kx, g.sW, g.pW, g.dW, g.oW);
                if (ox < 0) continue;
                const uint32_t key = ud_out_key(g, tab, c.x, oz, oy, ox);
                atomicOr(&bitmap[key >> 5], 1u << (key & 31));
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_morton_keys(const int4* __restrict__ coords, int n, UdConvGeom g, uint32_t* __restrict__ keys, int* __restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = coords[i];
    keys[i] = ((uint32_t)c.x << g.kl.bits_total) | ud_spread(g.kl, 0, c.y) | ud_spread(g.kl, 1, c.z) | ud_spread(g.kl, 2, c.w);
    idx[i] = i;
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
        int x, y, z;
        if (g.kl.order == 0) {
            x = t % g.oW; t /= g.oW;
            y = t % g.oH; t /= g.oH;
            z = t % g.oD; t /= g.oD;
        } else {
            x = y = z = 0;
            for (int i = 0; i < g.kl.bz; ++i) z |= (int)((t >> g.kl.pos[0][i]) & 1u) << i;
            for (int i = 0; i < g.kl.by; ++i) y |= (int)((t >> g.kl.pos[1][i]) & 1u) << i;
            for (int i = 0; i < g.kl.bx; ++i) x |= (int)((t >> g.kl.pos[2][i]) & 1u) << i;
            t >>= g.kl.bits_total;
        }
        if (o < n_out) out_coords[o] = make_int4((int)t, z, y, x);
        ++o;
    }
}

// Same 32-row block scheme for the input-major table of a strided convolution; the output-major table is scattered
// (nbr_out[o][k] = i, both forms, pre-filled with -1 by the launcher) with its tile masks set by a test-then-atomicOr.
__global__ void __launch_bounds__(256)
k_conv_tables(const int4* __restrict__ coords, int n, UdConvGeom g, const uint32_t* __restrict__ tab,
              const unsigned int* __restrict__ bitmap, const int* __restrict__ prefix, int n_out, int ld_in, int ld_out,
              int* __restrict__ nbr_in, int* __restrict__ nbr_out, int* __restrict__ nbr_in_t,
              unsigned int* __restrict__ masks_in, int* __restrict__ nbr_out_t, unsigned int* __restrict__ masks_out) {
    __shared__ int tile[32][33];
    const int K = g.kD * g.kH * g.kW;
    const int r0 = blockIdx.x * 32;
    for (int e = threadIdx.x; e < 32 * 33; e += 256) (&tile[0][0])[e] = -1;
    __syncthreads();
    {
        // 8 threads per row: each walks the row's taps axis by axis (only where an output coordinate exists, see
        // k_conv_mark) and takes every 8th existing pair -- the rank lookups (two random reads) of a row run in parallel
        const int rl = threadIdx.x >> 3, j = threadIdx.x & 7;
        const int i = r0 + rl;
        if (i < n) {
            const int4 c = __ldg(coords + i);
            int cnt = 0;
            for (int kz = 0; kz < g.kD; ++kz) {
                const int oz = ud_out_coord(c.y, kz, g.sD, g.pD, g.dD, g.oD);
                if (oz < 0) continue;
                for (int ky = 0; ky < g.kH; ++ky) {
                    const int oy = ud_out_coord(c.z, ky, g.sH, g.pH, g.dH, g.oH);
                    if (oy < 0) continue;
                    for (int kx = 0; kx < g.kW; ++kx) {
                        const int ox = ud_out_coord(c.w, kx, g.sW, g.pW, g.dW, g.oW);
                        if (ox < 0) continue;
                        if (((cnt++) & 7) != j) continue;
                        const int k = (kz * g.kH + ky) * g.kW + kx;
                        const uint32_t key = ud_out_key(g, tab, c.x, oz, oy, ox);
                        const unsigned int word = __ldg(bitmap + (key >> 5));
                        const unsigned int below = word & ((1u << (key & 31)) - 1u);
                        const int o = __ldg(prefix + (key >> 5)) + __popc(below);
                        if (o < n_out) {
                            if (nbr_out) nbr_out[(size_t)o * K + k] = i;
                            if (nbr_out_t) {
                                nbr_out_t[(size_t)k * ld_out + o] = i;
                                unsigned int* mw = masks_out + (o >> 7);
                                if (!((*reinterpret_cast<volatile unsigned int*>(mw) >> k) & 1u)) atomicOr(mw, 1u << k);
                            }
                            tile[rl][k] = o;
                        }
                    }
                }
            }
        }
    }
    __syncthreads();
    if (nbr_in != nullptr) {
        const int rows = min(32, n - r0);
        for (int e = threadIdx.x; e < rows * K; e += 256) nbr_in[(size_t)r0 * K + e] = tile[e / K][e % K];
    }
    if (nbr_in_t != nullptr) {
        const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
        const int k_pad = (K + 3) / 4 * 4;
        unsigned int m = 0u;
        for (int k = ty; k < k_pad; k += 8) {
            const int v = k < K ? tile[tx][k] : -1;
            if (r0 + tx < ld_in) nbr_in_t[(size_t)k * ld_in + r0 + tx] = v;
            if (__ballot_sync(0xffffffffu, r0 + tx < n && v >= 0)) m |= 1u << k;
        }
        if (tx == 0 && m) atomicOr(masks_in + (r0 >> 7), m);
    }
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

static int ud_bits_for(int n) {
    int b = 0;
    while ((1 << b) < n) ++b;
    return b;
}

// order 1: deal key bits out z, y, x round-robin from bit 0 while an axis has bits left
static int ud_make_layout(int order, int oD, int oH, int oW, int B, UdKeyLayout* kl) {
    kl->order = order;
    kl->bz = ud_bits_for(oD); kl->by = ud_bits_for(oH); kl->bx = ud_bits_for(oW);
    kl->bits_total = kl->bz + kl->by + kl->bx;
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < 12; ++i) kl->pos[a][i] = 0;
    if (order == 0) return UD_OK;
    if (order != 1) return UD_ERR_INVALID_ARG;
    if (kl->bz > 12 || kl->by > 12 || kl->bx > 12) return UD_ERR_UNSUPPORTED;
    if (((uint64_t)B << kl->bits_total) >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    int used[3] = {0, 0, 0};
    const int nb[3] = {kl->bz, kl->by, kl->bx};
    int p = 0;
    while (p < kl->bits_total)
        for (int a = 2; a >= 0; --a)          // x takes the lowest bit of each round, then y, then z
            if (used[a] < nb[a]) kl->pos[a][used[a]++] = (unsigned char)p++;
    return UD_OK;
}

static uint64_t ud_out_cells(const UdConvGeom& g) {
    if (g.kl.order == 0) return (uint64_t)g.B * g.oD * g.oH * g.oW;
    return (uint64_t)g.B << g.kl.bits_total;
}

static int ud_geom_check(const UdConvGeom& g, bool in_side) {
    uint64_t cells = in_side ? (uint64_t)g.B * g.iD * g.iH * g.iW : ud_out_cells(g);
    if (g.B < 1 || cells == 0 || cells >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    int K = g.kD * g.kH * g.kW;
    if (K < 1 || K > 32) return UD_ERR_UNSUPPORTED;
    return UD_OK;
}

static UdConvGeom ud_make_geom(int B, const int* in_shape, const int* out_shape, const int* ksize,
                               const int* stride, const int* padding, const int* dilation, int order = 0,
                               int* rc = nullptr) {
    UdConvGeom g;
    int r = ud_make_layout(order, out_shape[0], out_shape[1], out_shape[2], B, &g.kl);
    if (rc) *rc = r;
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
                     const int* dilation, int* nbr_out, int* nbr_out_t, unsigned int* tile_masks, void* workspace,
                     size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || (n > 0 && (!coords || (!nbr_out && !nbr_out_t)))) return UD_ERR_INVALID_ARG;
    if (nbr_out_t && !tile_masks) return UD_ERR_INVALID_ARG;
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
    const int ld = (n + 127) / 128 * 128;
    if (tile_masks) UD_CUDA_TRY(cudaMemsetAsync(tile_masks, 0, sizeof(unsigned int) * (size_t)(ld / 128), stream));
    k_subm_table<<<ud_blocks(nbr_out_t ? ld : n, 32), 256, 0, stream>>>((const int4*)coords, n, g, table, cap - 1, ld,
                                                                        nbr_out, nbr_out_t, tile_masks);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// ---- strided / regular sparse conv --------------------------------------------------------------
static size_t ud_conv_words(const UdConvGeom& g) { return (size_t)((ud_out_cells(g) + 31) / 32); }

int ud_rulebook_conv_workspace_bytes(int batch_size, const int* out_shape, int order, size_t* bytes) {
    if (!out_shape || !bytes || batch_size < 1) return UD_ERR_INVALID_ARG;
    UdKeyLayout kl;
    int rc = ud_make_layout(order, out_shape[0], out_shape[1], out_shape[2], batch_size, &kl);
    if (rc != UD_OK) return rc;
    uint64_t cells = order == 0 ? (uint64_t)batch_size * out_shape[0] * out_shape[1] * out_shape[2]
                                : ((uint64_t)batch_size << kl.bits_total);
    if (cells == 0 || cells >= 0xFFFFFFFFull) return UD_ERR_UNSUPPORTED;
    size_t words = (size_t)((cells + 31) / 32);
    size_t scan_bytes = 0;
    UD_CUDA_TRY((cub::DeviceScan::ExclusiveSum<const int*, int*>(nullptr, scan_bytes, nullptr, nullptr,
                                                                 (int)words + 1)));
    *bytes = ud_align_up(words * 4) + 2 * ud_align_up((words + 1) * 4) + ud_align_up(scan_bytes) +
             ud_align_up((size_t)(out_shape[0] + out_shape[1] + out_shape[2]) * 4);
    return UD_OK;
}

struct UdConvWs {
    uint32_t* tab;
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
    L->tab = a.take<uint32_t>((size_t)(g.oD + g.oH + g.oW));
    if (!a.ok()) return UD_ERR_WORKSPACE;
    return UD_OK;
}

// Phase 1: mark the output occupancy bitmap, rank it, write the number of output rows to
// n_out_dev (device int).  The workspace must be kept untouched until phase 2 ran.
int ud_rulebook_conv_count(const int* coords, int n_in, int batch_size, const int* in_shape,
                           const int* out_shape, const int* ksize, const int* stride, const int* padding,
                           const int* dilation, int order, int* n_out_dev, void* workspace, size_t workspace_bytes,
                           void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!n_out_dev || n_in < 0 || (n_in > 0 && !coords)) return UD_ERR_INVALID_ARG;
    int rc = UD_OK;
    UdConvGeom g = ud_make_geom(batch_size, in_shape, out_shape, ksize, stride, padding, dilation, order, &rc);
    if (rc != UD_OK) return rc;
    rc = ud_geom_check(g, false);
    if (rc != UD_OK) return rc;
    UdConvWs L;
    rc = ud_conv_layout(workspace, workspace_bytes, g, &L);
    if (rc != UD_OK) return rc;
    int words = (int)L.words;
    UD_CUDA_TRY(cudaMemsetAsync(L.bitmap, 0, L.words * 4, stream));
    if (g.kl.order != 0) {
        k_spread_tables<<<ud_blocks(g.oD + g.oH + g.oW, 256), 256, 0, stream>>>(g.kl, g.oD, g.oH, g.oW, L.tab);
        UD_LAUNCH_CHECK();
    }
    if (n_in > 0) {
        k_conv_mark<<<ud_blocks(n_in, 256), 256, 0, stream>>>((const int4*)coords, n_in, g, L.tab, L.bitmap);
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
                          const int* dilation, int order, int n_out, int* out_coords, int* nbr_in, int* nbr_out,
                          int* nbr_in_t, unsigned int* masks_in, int* nbr_out_t, unsigned int* masks_out,
                          void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_in < 0 || n_out < 0) return UD_ERR_INVALID_ARG;
    if (n_in > 0 && (!coords || (!nbr_in && !nbr_in_t))) return UD_ERR_INVALID_ARG;
    if (n_out > 0 && (!out_coords || (!nbr_out && !nbr_out_t))) return UD_ERR_INVALID_ARG;
    if ((nbr_in_t && !masks_in) || (nbr_out_t && !masks_out)) return UD_ERR_INVALID_ARG;
    int rc = UD_OK;
    UdConvGeom g = ud_make_geom(batch_size, in_shape, out_shape, ksize, stride, padding, dilation, order, &rc);
    if (rc != UD_OK) return rc;
    rc = ud_geom_check(g, false);
    if (rc != UD_OK) return rc;
    UdConvWs L;
    rc = ud_conv_layout(workspace, workspace_bytes, g, &L);
    if (rc != UD_OK) return rc;
    int K = g.kD * g.kH * g.kW;
    int words = (int)L.words;
    const int ld_in = (n_in + 127) / 128 * 128, ld_out = (n_out + 127) / 128 * 128;
    const int k_pad = (K + 3) / 4 * 4;
    if (n_out > 0) {
        if (nbr_out) UD_CUDA_TRY(cudaMemsetAsync(nbr_out, 0xFF, (size_t)n_out * K * sizeof(int), stream));
        if (nbr_out_t) {
            UD_CUDA_TRY(cudaMemsetAsync(nbr_out_t, 0xFF, (size_t)k_pad * ld_out * sizeof(int), stream));
            UD_CUDA_TRY(cudaMemsetAsync(masks_out, 0, sizeof(unsigned int) * (size_t)(ld_out / 128), stream));
        }
        k_conv_out_coords<<<ud_blocks(words, 256), 256, 0, stream>>>(L.bitmap, L.prefix, words, g, n_out,
                                                                     (int4*)out_coords);
        UD_LAUNCH_CHECK();
    }
    if (n_in > 0) {
        if (masks_in) UD_CUDA_TRY(cudaMemsetAsync(masks_in, 0, sizeof(unsigned int) * (size_t)(ld_in / 128), stream));
        k_conv_tables<<<ud_blocks(nbr_in_t ? ld_in : n_in, 32), 256, 0, stream>>>(
            (const int4*)coords, n_in, g, L.tab, L.bitmap, L.prefix, n_out, ld_in, ld_out, nbr_in, nbr_out, nbr_in_t,
            masks_in, nbr_out_t, masks_out);
        UD_LAUNCH_CHECK();
    }
    return UD_OK;
}

// ---- Morton order of a coordinate set ---------------------------------------------------------------
int ud_morton_order_workspace_bytes(int n, size_t* bytes) {
    if (n < 0 || !bytes) return UD_ERR_INVALID_ARG;
    size_t sort_bytes = 0;
    int m = n > 0 ? n : 1;
    UD_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int>(nullptr, sort_bytes, nullptr, nullptr, nullptr, nullptr,
                                                                m, 0, 32)));
    *bytes = 3 * ud_align_up((size_t)m * 4) + ud_align_up(sort_bytes);
    return UD_OK;
}

// coords (n,4) int32 [b,z,y,x] on a (z,y,x) grid `spatial_shape` (HOST int[3]) -> order (n) int32: row indices sorted by
// (b, Morton(z,y,x)) (stable: equal keys keep their input order).  Gathering features / coordinates with `order` gives
// the row order in which 128 consecutive rows form a compact 3-D block.
int ud_morton_order(const int* coords, int n, int batch_size, const int* spatial_shape, int* order, void* workspace,
                    size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || batch_size < 1 || !spatial_shape || (n > 0 && (!coords || !order))) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    UdConvGeom g;
    int rc = ud_make_layout(1, spatial_shape[0], spatial_shape[1], spatial_shape[2], batch_size, &g.kl);
    if (rc != UD_OK) return rc;
    g.B = batch_size;
    g.oD = spatial_shape[0]; g.oH = spatial_shape[1]; g.oW = spatial_shape[2];
    UdArena a(workspace, workspace_bytes);
    uint32_t* keys = a.take<uint32_t>(n);
    uint32_t* keys_sorted = a.take<uint32_t>(n);
    int* idx = a.take<int>(n);
    size_t sort_bytes = 0;
    UD_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int>(nullptr, sort_bytes, nullptr, nullptr, nullptr, nullptr,
                                                                n, 0, 32)));
    void* cub_temp = a.take<char>(sort_bytes);
    if (!a.ok()) return UD_ERR_WORKSPACE;
    k_morton_keys<<<ud_blocks(n, 256), 256, 0, stream>>>((const int4*)coords, n, g, keys, idx);
    UD_LAUNCH_CHECK();
    int end_bit = g.kl.bits_total + ud_bits_for(batch_size);
    if (end_bit > 32) end_bit = 32;
    UD_CUDA_TRY((cub::DeviceRadixSort::SortPairs<uint32_t, int>(cub_temp, sort_bytes, keys, keys_sorted, idx, order, n, 0,
                                                                end_bit, stream)));
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
