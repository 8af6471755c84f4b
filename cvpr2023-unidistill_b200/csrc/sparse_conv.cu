// Sparse 3-D convolution (SubM and strided) as an output-stationary gather-GEMM, fp32 SIMT.
//
// Replaces spconv's indice_conv / indice_subm_conv (ConvAlgo.Native: 27 x gather + GEMM +
// scatter-add launches per layer) used by every spconv.SubMConv3d / SparseConv3d in
//   reference unidistill/layers/blocks_3d/det3d/spconv_backbone.py:10-113,259-340.
//
//   out[o, :] = bias + sum_k in[nbr[o][k], :] @ W[k]        (nbr[o][k] == -1 -> no contribution)
//
// One launch per layer, no atomics, deterministic.  The same kernel computes dgrad when given the
// input-major table and W^T (see ud_sparse_conv_forward docs).  Weight gradients are a separate
// split-rows reduction (k_spconv_wgrad).
//
// CTA = 256 threads, tile = 128 output rows x all COUT columns, fp32 accumulators in registers.
// For every kernel tap k that has at least one present neighbour in the tile (tile-level skip;
// LiDAR rows are spatially coherent so whole taps vanish on the ground plane) the CTA gathers
// A_k[128 x KC] (transposed in smem so the FMA loop reads float4 rows) and W[k][KC x COUT].
#include "common.cuh"

namespace {

constexpr int TM = 128;
constexpr int NT = 256;
constexpr int APAD = 4;

template <int COUT> struct TileCfg;
template <> struct TileCfg<16>  { static constexpr int RM = 2, RN = 4; };
template <> struct TileCfg<32>  { static constexpr int RM = 4, RN = 4; };
template <> struct TileCfg<64>  { static constexpr int RM = 8, RN = 4; };
template <> struct TileCfg<128> { static constexpr int RM = 8, RN = 8; };

template <int CIN, int COUT>
__global__ void __launch_bounds__(NT)
k_spconv_gather_gemm(const float* __restrict__ in, const float* __restrict__ w,
                     const float* __restrict__ bias, const int* __restrict__ nbr, int n_out, int K,
                     float* __restrict__ out) {
    constexpr int RM = TileCfg<COUT>::RM;
    constexpr int RN = TileCfg<COUT>::RN;
    constexpr int TX = COUT / RN;
    constexpr int TY = TM / RM;
    static_assert(TX * TY == NT, "thread layout");
    constexpr int KC = CIN >= 32 ? 32 : (CIN >= 16 ? 16 : 8);   // channel chunk staged in smem
    constexpr int NCHUNK = (CIN + KC - 1) / KC;
    constexpr int WARP_ROWS = (32 / TX) * RM;                    // rows covered by one warp

    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* nbr_s = reinterpret_cast<int*>(smem_raw);                          // [TM][K]
    float* As = reinterpret_cast<float*>(nbr_s + TM * 32);                  // [KC][TM+APAD]
    float* Ws = As + KC * (TM + APAD);                                      // [KC][COUT]
    __shared__ unsigned int tile_mask;

    const int tid = threadIdx.x;
    const int tx = tid % TX;
    const int ty = tid / TX;
    const int lane = tid & 31;
    const int row0 = blockIdx.x * TM;

    if (tid == 0) tile_mask = 0u;
    __syncthreads();
    // ---- preload the neighbour table of this tile (coalesced) ----
    {
        const int rows_here = min(TM, n_out - row0);
        const int total = rows_here * K;
        const int* src = nbr + (size_t)row0 * K;
        unsigned int m = 0u;
        for (int e = tid; e < TM * K; e += NT) {
            int v = e < total ? src[e] : -1;
            nbr_s[e] = v;
            if (v >= 0) m |= 1u << (e % K);
        }
        m = __reduce_or_sync(0xffffffffu, m);
        if (lane == 0 && m) atomicOr(&tile_mask, m);
    }
    __syncthreads();
    const unsigned int kmask = tile_mask;
    // per-warp tap mask: which taps touch the rows this warp accumulates
    unsigned int wmask = 0u;
    {
        const int wrow0 = (tid / 32) * WARP_ROWS;
        for (int e = lane; e < WARP_ROWS * K; e += 32) {
            int r = e / K, k = e - r * K;
            if (nbr_s[(wrow0 + r) * K + k] >= 0) wmask |= 1u << k;
        }
        wmask = __reduce_or_sync(0xffffffffu, wmask);
    }

    float acc[RM][RN];
#pragma unroll
    for (int r = 0; r < RM; ++r)
#pragma unroll
        for (int c = 0; c < RN; ++c) acc[r][c] = 0.f;

    // gather mapping: 2 threads per row, each moves half of the KC chunk
    const int grow = tid % TM;
    const int ghalf = tid / TM;            // 0 or 1
    constexpr int GC = KC / 2;             // channels per gather thread

    for (int k = 0; k < K; ++k) {
        if (!((kmask >> k) & 1u)) continue;   // block-uniform
        const int src_row = nbr_s[grow * K + k];
        for (int ch = 0; ch < NCHUNK; ++ch) {
            const int c0 = ch * KC;
            __syncthreads();   // previous chunk fully consumed
            // ---- gather A chunk (transposed store) ----
            {
                const int cbeg = c0 + ghalf * GC;
                if constexpr (CIN % 4 == 0) {
#pragma unroll
                    for (int v = 0; v < GC / 4; ++v) {
                        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (src_row >= 0)
                            val = __ldg(reinterpret_cast<const float4*>(in + (size_t)src_row * CIN + cbeg) + v);
                        float* dst = As + (ghalf * GC + v * 4) * (TM + APAD) + grow;
                        dst[0 * (TM + APAD)] = val.x;
                        dst[1 * (TM + APAD)] = val.y;
                        dst[2 * (TM + APAD)] = val.z;
                        dst[3 * (TM + APAD)] = val.w;
                    }
                } else {
#pragma unroll
                    for (int v = 0; v < GC; ++v) {
                        int c = cbeg + v;
                        float val = 0.f;
                        if (src_row >= 0 && c < CIN) val = __ldg(in + (size_t)src_row * CIN + c);
                        As[(ghalf * GC + v) * (TM + APAD) + grow] = val;
                    }
                }
            }
            // ---- stage W[k][c0:c0+KC][:] ----
            {
                const float* wk = w + ((size_t)k * CIN + c0) * COUT;
                constexpr int NV = KC * COUT / 4;
                for (int e = tid; e < NV; e += NT) {
                    int c = (e * 4) / COUT;
                    float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 + c < CIN) val = __ldg(reinterpret_cast<const float4*>(wk) + e);
                    reinterpret_cast<float4*>(Ws)[e] = val;
                }
            }
            __syncthreads();
            if (!((wmask >> k) & 1u)) continue;   // warp-uniform: none of this warp's rows has tap k
#pragma unroll 8
            for (int kk = 0; kk < KC; ++kk) {
                float a[RM], b[RN];
                const float* ap = As + kk * (TM + APAD) + ty * RM;
                if constexpr (RM == 2) {
                    float2 t = *reinterpret_cast<const float2*>(ap);
                    a[0] = t.x; a[1] = t.y;
                } else {
#pragma unroll
                    for (int r4 = 0; r4 < RM / 4; ++r4) {
                        float4 t = *reinterpret_cast<const float4*>(ap + r4 * 4);
                        a[r4 * 4 + 0] = t.x; a[r4 * 4 + 1] = t.y; a[r4 * 4 + 2] = t.z; a[r4 * 4 + 3] = t.w;
                    }
                }
                const float* bp = Ws + kk * COUT;
#pragma unroll
                for (int c4 = 0; c4 < RN / 4; ++c4) {
                    float4 t = *reinterpret_cast<const float4*>(bp + c4 * (COUT / (RN / 4)) + tx * 4);
                    b[c4 * 4 + 0] = t.x; b[c4 * 4 + 1] = t.y; b[c4 * 4 + 2] = t.z; b[c4 * 4 + 3] = t.w;
                }
#pragma unroll
                for (int r = 0; r < RM; ++r)
#pragma unroll
                    for (int c = 0; c < RN; ++c) acc[r][c] = fmaf(a[r], b[c], acc[r][c]);
            }
        }
    }

    // ---- epilogue: bias + store ----
#pragma unroll
    for (int c4 = 0; c4 < RN / 4; ++c4) {
        const int col = c4 * (COUT / (RN / 4)) + tx * 4;
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (bias) bv = __ldg(reinterpret_cast<const float4*>(bias + col));
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            const int row = row0 + ty * RM + r;
            if (row < n_out) {
                float4 o;
                o.x = acc[r][c4 * 4 + 0] + bv.x;
                o.y = acc[r][c4 * 4 + 1] + bv.y;
                o.z = acc[r][c4 * 4 + 2] + bv.z;
                o.w = acc[r][c4 * 4 + 3] + bv.w;
                *reinterpret_cast<float4*>(out + (size_t)row * COUT + col) = o;
            }
        }
    }
}

template <int CIN, int COUT>
int launch_fwd(const float* in, const float* w, const float* bias, const int* nbr, int n_out, int K,
               float* out, cudaStream_t stream) {
    constexpr int KC = CIN >= 32 ? 32 : (CIN >= 16 ? 16 : 8);
    size_t smem = (size_t)TM * 32 * sizeof(int) + (size_t)KC * (TM + APAD) * sizeof(float) +
                  (size_t)KC * COUT * sizeof(float);
    auto kern = k_spconv_gather_gemm<CIN, COUT>;
    UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<ud_blocks(n_out, TM), NT, smem, stream>>>(in, w, bias, nbr, n_out, K, out);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Small-COUT fallback (COUT not in {16,32,64,128}, e.g. dgrad into 5 input channels): one thread per
// (row, cout) pair, plain loops.  Never on the hot path.
__global__ void __launch_bounds__(256)
k_spconv_generic(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                 const int* __restrict__ nbr, int n_out, int K, int CIN, int COUT, float* __restrict__ out) {
    long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)n_out * COUT) return;
    int o = (int)(t / COUT);
    int co = (int)(t - (long long)o * COUT);
    float acc = bias ? bias[co] : 0.f;
    for (int k = 0; k < K; ++k) {
        int i = nbr[(size_t)o * K + k];
        if (i < 0) continue;
        const float* x = in + (size_t)i * CIN;
        const float* wk = w + (size_t)k * CIN * COUT + co;
        for (int c = 0; c < CIN; ++c) acc = fmaf(x[c], wk[(size_t)c * COUT], acc);
    }
    out[t] = acc;
}

// ------------------------------------------------------------------------------------------------
// wgrad:  dW[k][ci][co] = sum_o in[nbr[o][k]][ci] * dout[o][co]
// grid = (row chunks, K).  Each CTA compacts the rows of its chunk that own tap k, stages them in
// smem 32 at a time and accumulates a CIN x COUT register tile; one atomicAdd per element at the end.
template <int CIN, int COUT>
__global__ void __launch_bounds__(NT)
k_spconv_wgrad(const float* __restrict__ in, const float* __restrict__ dout, const int* __restrict__ nbr,
               int n_out, int K, int rows_per_cta, float* __restrict__ dw) {
    constexpr int ELEMS = CIN * COUT;
    constexpr int PER_T = (ELEMS + NT - 1) / NT;               // accumulators per thread
    // thread tile RI x RO with RI*RO == PER_T
    constexpr int RO = PER_T >= 64 ? 8 : (PER_T >= 16 ? 4 : (PER_T >= 4 ? 2 : (PER_T >= 2 ? 2 : 1)));
    constexpr int RI = PER_T / RO > 0 ? PER_T / RO : 1;
    constexpr int TXO = COUT / RO;                             // threads along COUT
    constexpr int TYI = (CIN + RI - 1) / RI;                   // threads along CIN
    static_assert(TXO * TYI <= NT, "wgrad thread layout");
    constexpr int R = 32;                                      // staged rows per step
    __shared__ float As[R][CIN + 1];
    __shared__ float Ds[R][COUT + 4];
    __shared__ int pair_i[NT];
    __shared__ int pair_o[NT];
    __shared__ int warp_cnt[NT / 32];
    __shared__ int n_pairs_s;

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wid = tid >> 5;
    const int k = blockIdx.y;
    const int tx = tid % TXO;
    const int ty = tid / TXO;
    const bool active = ty < TYI;

    float acc[RI][RO];
#pragma unroll
    for (int a = 0; a < RI; ++a)
#pragma unroll
        for (int b = 0; b < RO; ++b) acc[a][b] = 0.f;

    const int rbeg = blockIdx.x * rows_per_cta;
    const int rend = min(n_out, rbeg + rows_per_cta);
    for (int base = rbeg; base < rend; base += NT) {
        // ---- ordered compaction of rows [base, base+NT) that have tap k ----
        int o = base + tid;
        int i = (o < rend) ? nbr[(size_t)o * K + k] : -1;
        unsigned int bal = __ballot_sync(0xffffffffu, i >= 0);
        if (lane == 0) warp_cnt[wid] = __popc(bal);
        __syncthreads();
        int off = 0;
        for (int w2 = 0; w2 < wid; ++w2) off += warp_cnt[w2];
        if (i >= 0) {
            int pos = off + __popc(bal & ((1u << lane) - 1u));
            pair_i[pos] = i;
            pair_o[pos] = o;
        }
        if (tid == NT - 1) n_pairs_s = off + __popc(bal);
        __syncthreads();
        const int n_pairs = n_pairs_s;
        for (int p0 = 0; p0 < n_pairs; p0 += R) {
            const int rows = min(R, n_pairs - p0);
            // stage A rows and dOut rows
            for (int e = tid; e < R * CIN; e += NT) {
                int r = e / CIN, c = e - r * CIN;
                As[r][c] = r < rows ? __ldg(in + (size_t)pair_i[p0 + r] * CIN + c) : 0.f;
            }
            for (int e = tid; e < R * COUT; e += NT) {
                int r = e / COUT, c = e - r * COUT;
                Ds[r][c] = r < rows ? __ldg(dout + (size_t)pair_o[p0 + r] * COUT + c) : 0.f;
            }
            __syncthreads();
            if (active) {
#pragma unroll 4
                for (int r = 0; r < R; ++r) {
                    float a[RI], d[RO];
#pragma unroll
                    for (int x = 0; x < RI; ++x) {
                        int ci = ty * RI + x;
                        a[x] = ci < CIN ? As[r][ci] : 0.f;
                    }
#pragma unroll
                    for (int y = 0; y < RO; ++y) d[y] = Ds[r][tx * RO + y];
#pragma unroll
                    for (int x = 0; x < RI; ++x)
#pragma unroll
                        for (int y = 0; y < RO; ++y) acc[x][y] = fmaf(a[x], d[y], acc[x][y]);
                }
            }
            __syncthreads();
        }
    }
    if (active) {
        float* dwk = dw + (size_t)k * CIN * COUT;
#pragma unroll
        for (int x = 0; x < RI; ++x) {
            int ci = ty * RI + x;
            if (ci >= CIN) continue;
#pragma unroll
            for (int y = 0; y < RO; ++y) {
                float v = acc[x][y];
                if (v != 0.f) atomicAdd(dwk + (size_t)ci * COUT + tx * RO + y, v);
            }
        }
    }
}

template <int CIN, int COUT>
int launch_wgrad(const float* in, const float* dout, const int* nbr, int n_out, int K, float* dw,
                 cudaStream_t stream) {
    // aim for ~4 CTAs per SM overall
    int chunks = (UD_NUM_SMS * 4 + K - 1) / K;
    int rows_per_cta = (n_out + chunks - 1) / chunks;
    rows_per_cta = (rows_per_cta + NT - 1) / NT * NT;
    if (rows_per_cta < NT) rows_per_cta = NT;
    chunks = (n_out + rows_per_cta - 1) / rows_per_cta;
    if (chunks < 1) chunks = 1;
    dim3 grid(chunks, K);
    k_spconv_wgrad<CIN, COUT><<<grid, NT, 0, stream>>>(in, dout, nbr, n_out, K, rows_per_cta, dw);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

__global__ void __launch_bounds__(256)
k_spconv_wgrad_generic(const float* __restrict__ in, const float* __restrict__ dout,
                       const int* __restrict__ nbr, int n_out, int K, int CIN, int COUT,
                       int rows_per_cta, float* __restrict__ dw) {
    // grid = (chunks, K); thread e owns element(s) e of the CIN x COUT tile
    const int k = blockIdx.y;
    const int rbeg = blockIdx.x * rows_per_cta;
    const int rend = min(n_out, rbeg + rows_per_cta);
    for (int e = threadIdx.x; e < CIN * COUT; e += blockDim.x) {
        int ci = e / COUT, co = e - ci * COUT;
        float acc = 0.f;
        for (int o = rbeg; o < rend; ++o) {
            int i = nbr[(size_t)o * K + k];
            if (i >= 0) acc = fmaf(in[(size_t)i * CIN + ci], dout[(size_t)o * COUT + co], acc);
        }
        if (acc != 0.f) atomicAdd(dw + ((size_t)k * CIN + ci) * COUT + co, acc);
    }
}

}  // namespace

#define UD_DISPATCH_CONV(CI, CO)                                                              \
    if (cin == CI && cout == CO)                                                              \
        return launch_fwd<CI, CO>(in, w, bias, nbr, n_out, K, out, stream);

#define UD_DISPATCH_WGRAD(CI, CO)                                                             \
    if (cin == CI && cout == CO)                                                              \
        return launch_wgrad<CI, CO>(in, dout, nbr, n_out, K, dw, stream);

extern "C" {

// in: (n_in, cin) f32; w: (K, cin, cout) f32; bias: (cout) or NULL; nbr: (n_out, K) int32 gather table;
// out: (n_out, cout) f32.
// dgrad is the same call with  in := dout, nbr := input-major table, w := W^T laid out (K, cout, cin)
// (for SubM additionally flipped in k), out := din.
int ud_sparse_conv_forward(const float* in, const float* w, const float* bias, const int* nbr, int n_out,
                           int K, int cin, int cout, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!in || !w || !nbr || !out || n_out < 0 || K < 1 || K > 32 || cin < 1 || cout < 1)
        return UD_ERR_INVALID_ARG;
    if (n_out == 0) return UD_OK;
    UD_DISPATCH_CONV(5, 16)
    UD_DISPATCH_CONV(16, 16)
    UD_DISPATCH_CONV(16, 32)
    UD_DISPATCH_CONV(32, 16)
    UD_DISPATCH_CONV(32, 32)
    UD_DISPATCH_CONV(32, 64)
    UD_DISPATCH_CONV(64, 32)
    UD_DISPATCH_CONV(64, 64)
    UD_DISPATCH_CONV(64, 128)
    UD_DISPATCH_CONV(128, 64)
    UD_DISPATCH_CONV(128, 128)
    k_spconv_generic<<<ud_blocks((long long)n_out * cout, 256), 256, 0, stream>>>(in, w, bias, nbr, n_out, K,
                                                                                  cin, cout, out);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// dw: (K, cin, cout) f32, accumulated into (caller zeroes it).
int ud_sparse_conv_wgrad(const float* in, const float* dout, const int* nbr, int n_out, int K, int cin,
                         int cout, float* dw, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!in || !dout || !nbr || !dw || n_out < 0 || K < 1 || K > 32 || cin < 1 || cout < 1)
        return UD_ERR_INVALID_ARG;
    if (n_out == 0) return UD_OK;
    UD_DISPATCH_WGRAD(5, 16)
    UD_DISPATCH_WGRAD(16, 16)
    UD_DISPATCH_WGRAD(16, 32)
    UD_DISPATCH_WGRAD(32, 32)
    UD_DISPATCH_WGRAD(32, 64)
    UD_DISPATCH_WGRAD(64, 64)
    UD_DISPATCH_WGRAD(64, 128)
    UD_DISPATCH_WGRAD(128, 128)
    {
        int chunks = (UD_NUM_SMS * 4 + K - 1) / K;
        int rows_per_cta = (n_out + chunks - 1) / chunks;
        if (rows_per_cta < 1) rows_per_cta = 1;
        chunks = (n_out + rows_per_cta - 1) / rows_per_cta;
        dim3 grid(chunks, K);
        k_spconv_wgrad_generic<<<grid, 256, 0, stream>>>(in, dout, nbr, n_out, K, cin, cout, rows_per_cta, dw);
        UD_LAUNCH_CHECK();
    }
    return UD_OK;
}

}  // extern "C"
