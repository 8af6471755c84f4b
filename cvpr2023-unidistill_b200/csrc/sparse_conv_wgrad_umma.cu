// Weight gradient of the sparse 3-D convolution on the 5th-generation tensor cores (tcgen05, TMEM), sm_100a only.
//
// Same contract as ud_sparse_conv_wgrad (sparse_conv.cu):
//     dw[k][ci][co] += sum_o in[nbr[o][k]][ci] * dout[o][co]          nbr[o][k] == -1 -> no contribution
// i.e. the wgrad half of spconv's indice_conv backward behind
//   reference unidistill/layers/blocks_3d/det3d/spconv_backbone.py:21-48,70-92 (ConvAlgo.Native under autograd).
//
// Formulation: for every kernel tap a GEMM whose reduction dimension is the row (voxel) dimension,
//     dW[k] (CIN x COUT) = A_k^T (CIN x rows) * G (rows x COUT),   A_k[o, :] = in[nbr[o][k], :]  (zero row if absent)
// Both operands lie in memory with the channel dimension contiguous, i.e. "MN-major" in UMMA terms (the reduction
// dimension is the strided one), which tcgen05.mma kind::tf32 reads directly from shared memory: a pipeline stage holds
//     A  = 64 gathered rows x 128 columns  (128/CIN consecutive taps side by side: a "super-tap", UMMA M = 128)
//     G  = the 64 dout rows x COUT         (UMMA N = COUT), loaded once per 64 rows and shared by all taps of the CTA
// as [64 rows x 32 fp32] sub-tiles in the canonical MN-major SWIZZLE_128B_BASE32B layout (32-byte chunk ^= row & 3,
// 4-row groups 512 B apart, sub-tiles 8 KB apart = the descriptor's leading byte offset).  One tcgen05.mma consumes 8
// rows (K = 8) of both tiles.  The fp32 accumulators of up to 512/COUT super-taps stay in TMEM for the whole CTA and are
// added to dw once at the end; the grid is (row chunks) x (super-tap groups), about one CTA per SM.
// Precision: operands are rounded to TF32 to-nearest on their way into shared memory (truncation would bias every
// product by ~-2^-11), the accumulation is fp32: measured ~1e-4 relative against the fp32 SIMT kernel, inside the
// 1e-3 BASELINE.json allows for fp32 losses/gradients.
//
// Warp roles (544 threads): warps 0-15  four producer groups, one per gathered-stage ring slot: row indices from the
//                                       tap-major table -> 16-byte gathers -> cvt.rna.tf32 -> swizzled st.shared ->
//                                       fence.proxy.async -> mbarrier
//                           warp 16     TMEM owner + MMA issuer (one elected lane), tcgen05.commit frees the stages
//                           warps 0-3   epilogue afterwards: tcgen05.ld -> red.global.add into dw
#include "common.cuh"
#include "umma_common.cuh"

namespace {

using namespace ud_umma;

constexpr int WG_TILE = 128;                 // rows per tile of the tap-major table / per tile-mask word
constexpr int WG_RS = 64;                    // rows per pipeline stage
// Template parameter NG = producer groups = gathered-stage ring slots of a CTA:
//   NG 4: one CTA per SM, 512 TMEM columns of accumulators (the round-1 form);
//   NG 2: TWO CTAs per SM with 256 TMEM columns each -- as in k_spconv_umma_dg (profiles/r03_pipeline_probes.md) the single
//         MMA-issuing warp is a serial chain of long-latency instructions per stage, and two resident CTAs give the SM two
//         of them.  Costs: half the accumulator slots per CTA, so twice as many super-tap groups re-read dout.
constexpr int WG_GROUP_THREADS = 128;
constexpr int WG_G_STAGES = 2;
constexpr int WG_ATOM_BYTES = WG_RS * 128;   // one [64 rows x 32 fp32] swizzled sub-tile
constexpr int WG_A_BYTES = 4 * WG_ATOM_BYTES;   // 64 rows x 128 columns
constexpr int WG_MAX_TAPS = 32;
__host__ __device__ constexpr int wg_threads(int ng) { return ng * WG_GROUP_THREADS + 32; }
__host__ __device__ constexpr int wg_tmem_cols(int ng) { return ng == 2 ? 256 : 512; }

__host__ __device__ constexpr size_t wg_smem_bytes(int cout, int ng = 4) {
    return (size_t)ng * WG_A_BYTES + (size_t)WG_G_STAGES * WG_RS * cout * 4 + 1024;
}

// MN-major TF32 operand.  The only swizzled layout tcgen05 accepts for 32-bit MN-major data is SWIZZLE_128B_BASE32B
// (CuTe's Layout_MN_SW128_32B_Atom): 32 fp32 of the M/N dimension contiguous (one 128-byte row per reduction index),
// the 32-byte chunk of a row XOR-ed with (row & 3), 4-row groups 512 B apart (stride byte offset), the next 32 columns
// WG_ATOM_BYTES further (leading byte offset).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128_32b(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address      bits [0,14)
    d |= (uint64_t)(WG_ATOM_BYTES >> 4) << 16;           // leading byte offset bits [16,30)
    d |= (uint64_t)(512 >> 4) << 32;                     // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                              // SWIZZLE_128B_BASE32B
    return d;
}
// byte offset of 16-byte chunk `chunk16` of row `row` inside a [rows x 32 fp32] sub-tile in that layout
__device__ __forceinline__ uint32_t swz32_off(int row, int chunk16) {
    return (uint32_t)(row * 128 + ((((chunk16 >> 1) ^ (row & 3)) << 5) | ((chunk16 & 1) << 4)));
}

__device__ __forceinline__ uint32_t rna_tf32(uint32_t x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(__uint_as_float(x)));
    return r;
}
__device__ __forceinline__ uint4 rna_tf32(uint4 v) {
    return make_uint4(rna_tf32(v.x), rna_tf32(v.y), rna_tf32(v.z), rna_tf32(v.w));
}

template <int CIN, int COUT, int NG>
__global__ void __launch_bounds__(wg_threads(NG), NG == 2 ? 2 : 1)
k_spconv_wgrad_umma(const float* __restrict__ in, const float* __restrict__ dout, const int* __restrict__ nbr_t,
                    const unsigned int* __restrict__ tile_masks, int n_rows, int K, int tiles_per_cta, int st_per_cta,
                    float* __restrict__ dw) {
    constexpr int TAPG = 128 / CIN;                   // kernel taps per super-tap (UMMA M = TAPG * CIN = 128)
    constexpr int NCOL = COUT < 32 ? 32 : COUT;       // UMMA N: a 16-wide dout tile is zero-padded to one 32-column sub-tile
    constexpr int NGA = NCOL / 32;                    // 32-column sub-tiles of the dout stage
    constexpr int G_BYTES = WG_RS * NCOL * 4;
    constexpr int G_STEP = WG_GROUP_THREADS / (8 * NGA);     // dout rows covered by one pass of a producer group
    constexpr int G_NJ = WG_RS / G_STEP;
    // kind::tf32, fp32 accumulate, A and B MN-major, N = COUT, M = 128
    constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) |
                               ((uint32_t)(NCOL >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr int WG_GROUPS = NG;                     // producer groups, one per gathered-stage ring slot
    constexpr int WG_A_STAGES = NG;
    constexpr int WG_MMA_WARP = NG * WG_GROUP_THREADS / 32;
    constexpr int WG_THREADS = wg_threads(NG);
    constexpr int TMEM_COLS = wg_tmem_cols(NG);
    static_assert((NG == 4 || NG == 2) && WG_G_STAGES <= WG_GROUPS, "one producer group per ring slot");
    static_assert(CIN == 16 || CIN == 32 || CIN == 64 || CIN == 128, "CIN must divide 128 in whole 16-byte chunks");
    static_assert(COUT == 16 || COUT == 32 || COUT == 64 || COUT == 128, "unsupported COUT");

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[2 * WG_A_STAGES + 2 * WG_G_STAGES + 1];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int chunk_mask_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int n_tiles = (n_rows + WG_TILE - 1) / WG_TILE;
    const int tile0 = blockIdx.x * tiles_per_cta;
    const int tile1 = min(n_tiles, tile0 + tiles_per_cta);
    const int K_eff = (K + TAPG - 1) / TAPG;
    const int st0 = blockIdx.y * st_per_cta;              // first super-tap of this CTA
    const int nst = min(st_per_cta, K_eff - st0);          // its super-taps (TMEM accumulator slots)
    const size_t ld_t = (size_t)n_tiles * WG_TILE;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t smem_g = smem_base + WG_A_STAGES * WG_A_BYTES;
    unsigned char* const smem_g_al = smem_al + WG_A_STAGES * WG_A_BYTES;
    const uint32_t bar_a_full = smem_u32(&bars[0]);
    const uint32_t bar_a_empty = bar_a_full + 8 * WG_A_STAGES;
    const uint32_t bar_g_full = bar_a_empty + 8 * WG_A_STAGES;
    const uint32_t bar_g_empty = bar_g_full + 8 * WG_G_STAGES;
    const uint32_t bar_acc = bar_g_empty + 8 * WG_G_STAGES;

    if (tid == 0) {
        for (int s = 0; s < WG_A_STAGES; ++s) {
            mbar_init(bar_a_full + 8 * s, WG_GROUP_THREADS / 32);   // lane 0 of each producer warp
            mbar_init(bar_a_empty + 8 * s, 1);
        }
        for (int s = 0; s < WG_G_STAGES; ++s) {
            mbar_init(bar_g_full + 8 * s, WG_GROUP_THREADS / 32);
            mbar_init(bar_g_empty + 8 * s, 1);
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
        chunk_mask_s = 0u;
    }
    if (warp == WG_MMA_WARP) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
    __syncthreads();

    // bit s <- super-tap st0 + s has at least one of its taps in the 128-row tile whose tap mask is m
    auto group_mask = [&](unsigned int m) {
        unsigned int r = 0u;
        for (int s = 0; s < nst; ++s)
            if ((m >> ((st0 + s) * TAPG)) & ((1u << TAPG) - 1u)) r |= 1u << s;
        return r;
    };
    {
        unsigned int m = 0u;
        for (int t = tile0 + tid; t < tile1; t += WG_THREADS) m |= group_mask(__ldg(tile_masks + t));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
        if (lane == 0 && m) atomicOr(&chunk_mask_s, m);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned int chunk_mask = chunk_mask_s;     // accumulator slots that receive at least one MMA
    const uint32_t tmem_acc = tmem_base_s;

    if (warp < WG_MMA_WARP) {
        // ============ producers: group g takes pipeline items g, g + 4, ... ============
        const int g = warp >> 2;
        const int tg = tid & (WG_GROUP_THREADS - 1);
        const int c16 = tg & 7;                       // 16-byte chunk of a 128-byte sub-tile row
        // Item -> group by the ring slot it fills (dout stage ig -> slot ig % 2 -> group ig % 2; gathered stage ia ->
        // slot ia % 4 -> group ia % 4), so every mbarrier is waited on by the same threads phase after phase: a parity
        // wait can only tell the current phase from the previous one, and groups drift apart by more than that.
        int ia = 0, ig = 0;
        for (int t = tile0; t < tile1; ++t) {
#if !defined(UD_WG_NO_PREFETCH)
            // row indices are on the critical path of every gathered stage (index load -> gather -> store are serial inside a
            // group): one group per tile prefetches the NEXT tile's index segments (this CTA's taps x 128 rows = 4 lines per tap).
            // Measured on B200 (bs-4 step): 64->64 0.677 -> 0.632 ms with the index prefetch, -> 0.621 with dout as well.
            if (t + 1 < tile1 && ((t - tile0) % WG_GROUPS) == g) {
                const int n_lines = nst * TAPG * 4;
                if (tg < n_lines) {
                    const int tap = st0 * TAPG + tg / 4;
                    if (tap < K)
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(nbr_t + (size_t)tap * ld_t + (size_t)(t + 1) * WG_TILE + (tg & 3) * 32));
                }
                // and the next tile's dout rows (128 rows x COUT floats, contiguous) towards L2
                for (int l = tg; l < WG_TILE * COUT / 32; l += WG_GROUP_THREADS) {
                    const size_t off = (size_t)(t + 1) * WG_TILE * COUT + (size_t)l * 32;
                    if (off < (size_t)n_rows * COUT) asm volatile("prefetch.global.L2 [%0];" ::"l"(dout + off));
                }
            }
#endif
            const unsigned int m = group_mask(__ldg(tile_masks + t));
            if (!m) continue;
            for (int h = 0; h < WG_TILE / WG_RS; ++h) {
                const int row0 = t * WG_TILE + h * WG_RS;
                if (row0 >= n_rows) break;
                if ((ig % WG_G_STAGES) == g) {
                    // ---- dout stage: rows row0 .. row0+63, all COUT columns ----
                    const int slot = ig % WG_G_STAGES;
                    const int atom = (tg >> 3) % NGA;
                    const int rb = tg / (8 * NGA);
                    uint4 v[G_NJ];
#pragma unroll
                    for (int j = 0; j < G_NJ; ++j) {
                        const int grow = row0 + rb + G_STEP * j;
                        v[j] = make_uint4(0u, 0u, 0u, 0u);
                        if (grow < n_rows && atom * 32 + c16 * 4 < COUT)
                            v[j] = __ldg(reinterpret_cast<const uint4*>(dout + (size_t)grow * COUT + atom * 32 + c16 * 4));
                    }
                    mbar_wait(bar_g_empty + 8 * slot, (((uint32_t)(ig / WG_G_STAGES)) & 1u) ^ 1u);
                    unsigned char* base = smem_g_al + slot * G_BYTES + atom * WG_ATOM_BYTES;
#pragma unroll
                    for (int j = 0; j < G_NJ; ++j)
                        *reinterpret_cast<uint4*>(base + swz32_off(rb + G_STEP * j, c16)) = rna_tf32(v[j]);
                    fence_proxy_async();
                    __syncwarp();                     // one arrive per warp (32 same-address arrivals serialise)
                    if (lane == 0) mbar_arrive(bar_g_full + 8 * slot);
                }
                ++ig;
                for (unsigned int rem = m; rem; rem &= rem - 1u) {
                    if ((ia % WG_A_STAGES) == g) {
                        // ---- gathered stage: 64 rows x (TAPG taps x CIN channels) ----
                        const int s = __ffs(rem) - 1;
                        const int slot = ia % WG_A_STAGES;
                        const int atom = (tg >> 3) & 3;
                        const int rb = tg >> 5;           // rows rb + 4 j
                        // this thread's 16-byte chunk = column atom * 32 + c16 * 4 of the stage = (tap, channel offset)
                        const int tap = (st0 + s) * TAPG + (atom * 32 + c16 * 4) / CIN;
                        const int coff = (atom * 32 + c16 * 4) % CIN;
                        const bool tap_ok = tap < K;          // the last super-tap may reach past the kernel's taps
                        const int* col = nbr_t + (size_t)(tap_ok ? tap : 0) * ld_t + row0 + rb;
                        int idx[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) idx[j] = tap_ok ? __ldg(col + 4 * j) : -1;
                        uint4 v[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            v[j] = make_uint4(0u, 0u, 0u, 0u);
#if defined(UD_WG_EXP) && (UD_WG_EXP == 1 || UD_WG_EXP == 3)
                            if (idx[j] == -12345)
#else
                            if (idx[j] >= 0)
#endif
                                v[j] = __ldg(reinterpret_cast<const uint4*>(in + (size_t)idx[j] * CIN + coff));
                        }
                        mbar_wait(bar_a_empty + 8 * slot, (((uint32_t)(ia / WG_A_STAGES)) & 1u) ^ 1u);
                        unsigned char* base = smem_al + slot * WG_A_BYTES + atom * WG_ATOM_BYTES;
#if defined(UD_WG_EXP) && UD_WG_EXP == 3
                        if (idx[0] == -12345)
#endif
#pragma unroll
                        for (int j = 0; j < 16; ++j)
                            *reinterpret_cast<uint4*>(base + swz32_off(rb + 4 * j, c16)) = rna_tf32(v[j]);
                        fence_proxy_async();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(bar_a_full + 8 * slot);
                    }
                    ++ia;
                }
            }
        }
    } else {
        // ============ MMA issuer ============
        unsigned int started = 0u;                    // accumulator slots written at least once
        int ia = 0, ig = 0;
        for (int t = tile0; t < tile1; ++t) {
            const unsigned int m = group_mask(__ldg(tile_masks + t));
            if (!m) continue;
            for (int h = 0; h < WG_TILE / WG_RS; ++h) {
                const int row0 = t * WG_TILE + h * WG_RS;
                if (row0 >= n_rows) break;
                const int gslot = ig % WG_G_STAGES;
                mbar_wait(bar_g_full + 8 * gslot, ((uint32_t)(ig / WG_G_STAGES)) & 1u);
                const uint32_t g0 = smem_g + gslot * G_BYTES;
                for (unsigned int rem = m; rem; rem &= rem - 1u) {
                    const int s = __ffs(rem) - 1;
                    const int aslot = ia % WG_A_STAGES;
                    mbar_wait(bar_a_full + 8 * aslot, ((uint32_t)(ia / WG_A_STAGES)) & 1u);
                    tc_fence_after();
                    if (elect_one_sync()) {
                        // one descriptor per operand and stage; 8 rows further = +1024 B = +64 in the 16-byte address field
                        // (a stage never crosses the 256 KB the field spans)
                        const uint64_t da = umma_desc_mn_sw128_32b(smem_base + aslot * WG_A_BYTES);
                        const uint64_t dg = umma_desc_mn_sw128_32b(g0);
                        const uint32_t d = tmem_acc + (uint32_t)(s * NCOL);
                        const uint32_t first = (started >> s) & 1u;
#pragma unroll
                        for (int ks = 0; ks < WG_RS / 8; ++ks)      // 8 rows per MMA = two 512-byte swizzle groups
#if defined(UD_WG_EXP) && UD_WG_EXP == 2
                            if (n_rows == -7)
#endif
                            umma_tf32(d, da + (uint64_t)(ks * 64), dg + (uint64_t)(ks * 64), IDESC, (first || ks > 0) ? 1u : 0u);
                        umma_commit(bar_a_empty + 8 * aslot);
                    }
                    started |= 1u << s;
                    __syncwarp();
                    ++ia;
                }
                if (elect_one_sync()) umma_commit(bar_g_empty + 8 * gslot);
                __syncwarp();
                ++ig;
            }
        }
        if (elect_one_sync()) umma_commit(bar_acc);
        __syncwarp();
    }

    // ============ epilogue: accumulator row m = (tap within the super-tap) * CIN + ci ============
    if (warp < 4 && chunk_mask != 0u) {
        mbar_wait(bar_acc, 0u);
        tc_fence_after();
        const int mrow = warp * 32 + lane;
        const int gtap = mrow / CIN, ci = mrow % CIN;
        for (unsigned int rem = chunk_mask; rem; rem &= rem - 1u) {
            const int s = __ffs(rem) - 1;
            const int tap = (st0 + s) * TAPG + gtap;
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(tmem_acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)(s * NCOL + c0), v);
                tmem_ld_wait();
                if (tap < K) {
                    float* dst = dw + ((size_t)tap * CIN + ci) * COUT + c0;
#pragma unroll
                    for (int i = 0; i < (COUT < 32 ? COUT : 32); ++i) {
                        const float x = __uint_as_float(v[i]);
                        if (x != 0.f) atomicAdd(dst + i, x);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == WG_MMA_WARP) tmem_dealloc(tmem_acc, TMEM_COLS);
}

template <int CIN, int COUT, int NG>
int launch_wgrad_umma_n(const float* in, const float* dout, const int* nbr_t, const unsigned int* tile_masks, int n_rows,
                        int K, float* dw, cudaStream_t stream) {
    constexpr int TAPG = 128 / CIN;
    const int k_eff = (K + TAPG - 1) / TAPG;
    const int st_max = wg_tmem_cols(NG) / (COUT < 32 ? 32 : COUT);     // accumulator slots TMEM holds
    const int ngroups = (k_eff + st_max - 1) / st_max;
    const int st_per_cta = (k_eff + ngroups - 1) / ngroups;
    const int n_tiles = (n_rows + WG_TILE - 1) / WG_TILE;
    int chunks = UD_NUM_SMS * (NG == 2 ? 2 : 1) / ngroups;             // about one CTA per SM (NG 4, 5) / two (NG 2)
    if (chunks < 1) chunks = 1;
    if (chunks > n_tiles) chunks = n_tiles;
    const int tiles_per_cta = (n_tiles + chunks - 1) / chunks;
    chunks = (n_tiles + tiles_per_cta - 1) / tiles_per_cta;
    const size_t smem = wg_smem_bytes(COUT < 32 ? 32 : COUT, NG);
    auto kern = k_spconv_wgrad_umma<CIN, COUT, NG>;
    UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NG == 2)
        UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid(chunks, (k_eff + st_per_cta - 1) / st_per_cta);
    kern<<<grid, wg_threads(NG), smem, stream>>>(in, dout, nbr_t, tile_masks, n_rows, K, tiles_per_cta, st_per_cta, dw);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Producer groups per CTA for a layer shape (UD_WG_NG=2|4 overrides, diagnostic).  NG 2 needs two CTAs' shared memory on
// one SM: 2 x (2 x 32 KB gathered stages + 2 dout stages), i.e. COUT <= 64.
int wg_ng(int cin, int cout) {
    static const int forced = [] {
        const char* e = getenv("UD_WG_NG");
        return e ? atoi(e) : 0;
    }();
    (void)cin;
    if (cout > 64) return 4;
    if (forced == 2 || forced == 4) return forced;
    // measured on B200 (bs-4 distillation step): NG 2 is slower where it matters (64 -> 64: 0.671 -> 0.810 ms: twice the
    // super-tap groups re-read dout and split 14 super-taps 4/4/4/2) and within 2 % elsewhere -- unlike the forward kernel,
    // this one is not bound by its issuing warp.  Kept as an option.  (Five producer groups instead of four were measured too:
    // 0.68 -> 0.75 ms, the fifth group's registers spill.)
    return 4;
}

template <int CIN, int COUT>
int launch_wgrad_umma(const float* in, const float* dout, const int* nbr_t, const unsigned int* tile_masks, int n_rows,
                      int K, float* dw, cudaStream_t stream) {
    if constexpr (COUT <= 64) {
        if (wg_ng(CIN, COUT) == 2)
            return launch_wgrad_umma_n<CIN, COUT, 2>(in, dout, nbr_t, tile_masks, n_rows, K, dw, stream);
    }
    return launch_wgrad_umma_n<CIN, COUT, 4>(in, dout, nbr_t, tile_masks, n_rows, K, dw, stream);
}

bool wgrad_shape_ok(int K, int cin, int cout) {
    if (K < 1 || K > WG_MAX_TAPS) return false;
    const bool ci = cin == 16 || cin == 32 || cin == 64 || cin == 128;
    const bool co = cout == 16 || cout == 32 || cout == 64 || cout == 128;
    if ((cin == 16 || cout == 16) && (cin > 32 || cout > 32)) return false;     // not instantiated
    return ci && co;
}

}  // namespace

#define UD_DISPATCH_WGRAD_UMMA(CI, CO)         \
    if (cin == CI && cout == CO)               \
        return launch_wgrad_umma<CI, CO>(in, dout, nbr_t, tile_masks, n_rows, K, dw, stream);

extern "C" {

// 1 when (K, cin, cout) has a tensor-core wgrad instantiation, else 0 (use ud_sparse_conv_wgrad).
int ud_sparse_conv_tc_wgrad_supported(int K, int cin, int cout) { return wgrad_shape_ok(K, cin, cout) ? 1 : 0; }

// in (n_in, cin) f32 and dout (n_rows, cout) f32 with 16-byte aligned rows; nbr_t (K4, ld) tap-major gather table and
// tile_masks (ceil(n_rows/128)) from ud_rulebook_transpose of the (n_rows, K) table; dw (K, cin, cout) f32 accumulated
// into (the caller zeroes it).
int ud_sparse_conv_tc_wgrad(const float* in, const float* dout, const int* nbr_t, const unsigned int* tile_masks,
                            int n_rows, int K, int cin, int cout, float* dw, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_rows < 0 || !dw) return UD_ERR_INVALID_ARG;
    if (!wgrad_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if (n_rows == 0) return UD_OK;
    if (!in || !dout || !nbr_t || !tile_masks) return UD_ERR_INVALID_ARG;
    if ((((uintptr_t)in | (uintptr_t)dout) & 15u) != 0) return UD_ERR_INVALID_ARG;
    UD_DISPATCH_WGRAD_UMMA(16, 16)
    UD_DISPATCH_WGRAD_UMMA(16, 32)
    UD_DISPATCH_WGRAD_UMMA(32, 16)
    UD_DISPATCH_WGRAD_UMMA(32, 32)
    UD_DISPATCH_WGRAD_UMMA(32, 64)
    UD_DISPATCH_WGRAD_UMMA(32, 128)
    UD_DISPATCH_WGRAD_UMMA(64, 32)
    UD_DISPATCH_WGRAD_UMMA(64, 64)
    UD_DISPATCH_WGRAD_UMMA(64, 128)
    UD_DISPATCH_WGRAD_UMMA(128, 32)
    UD_DISPATCH_WGRAD_UMMA(128, 64)
    UD_DISPATCH_WGRAD_UMMA(128, 128)
    return UD_ERR_UNSUPPORTED;
}

}  // extern "C"
