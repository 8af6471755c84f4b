// Sparse 3-D convolution on the 5th-generation tensor cores (tcgen05, TMEM accumulators), sm_100a only.
//
// Same contract as ud_sparse_conv_forward (sparse_conv.cu):
//     out[o, :] = bias + sum_k in[nbr[o][k], :] @ W[k]            nbr[o][k] == -1 -> no contribution
// replacing spconv's ConvAlgo.Native gather/GEMM/scatter-add triplets behind
//   reference unidistill/layers/blocks_3d/det3d/spconv_backbone.py:21-48,70-92 (+ their dgrad, which is
//   this same kernel run on the transposed table and weights).
//
// Formulation: implicit GEMM, output stationary.  A tile = 128 consecutive output rows (UMMA M = 128) x all COUT columns
// (UMMA N = COUT) with its fp32 accumulator in TMEM; one CTA owns up to 256/COUT (<= 8) tiles so that a weight slice is
// fetched once per CTA.  A pipeline stage = one kernel (super-)tap x 64 input channels x one tile (narrow layers put
// 2 (CIN 32) / 4 (CIN 16) taps side by side); the per-tile tap masks of the rulebook decide which stages exist.
// Precision: the reference is fp32 and BASELINE.json asks for 1e-3 through 21 layers.  Both operands are split into two
// bf16 terms (x = hi + lo to 2^-17) and the product is accumulated as lo*W_hi + hi*W_lo + hi*W_hi, three
// tcgen05.mma kind::f16 per 16 channels, fp32 accumulation: <= 2e-5 of the largest output against the C oracle.
// The gathered operand never touches shared memory (global -> registers in the tcgen05.st.16x256b fragment layout ->
// bf16 hi/lo -> TMEM); shared memory holds only the weight-slice ring and the row-index ring.  DESIGN.md section 3.1.
// The kernel's predecessors (A through shared memory, 3xTF32, TF32 + bf16 corrections: "variants 0-3" of round 1) are kept
// for reference under tools/conv_variants/ and are not part of the library.
#include "common.cuh"
#include "umma_common.cuh"
#include "det_reduce.cuh"
#include "tma.cuh"
#include <cuda_bf16.h>

namespace {

using namespace ud_umma;

constexpr int TILE_M = 128;
constexpr int MAX_TAPS = 32;
constexpr int MAX_TILES = 8;           // output tiles (TMEM accumulators) per CTA
constexpr int MAX_SCHED = MAX_TAPS * 4 * 8;   // >= taps * stages per tap * tiles for every instantiation
constexpr int TS_GROUP = 128;          // threads of one gather set
constexpr int TS_ACC_COLS = 256;       // TMEM columns holding accumulators; the other 256 hold the A-operand ring


// taps that share one 64-channel stage (narrow layers)
__host__ __device__ constexpr int ts_tapg(int cin) { return cin < 64 ? 64 / cin : 1; }
// bytes of one weight slice: [COUT x 64] bf16 W_hi + [COUT x 64] bf16 W_lo
__host__ __device__ constexpr int dg_b_bytes(int cout) { return 4 * cout * 64; }

// K order inside a 64-channel stage.  For CIN >= 32 the gather warps read 32 bytes (8 channels) per lane: lane l of a quad
// holds physical channels 8 l .. 8 l + 7 of a 32-channel block and feeds them to K positions 4 l .. 4 l + 3 of the block's two
// 16-channel MMA units, i.e. physical position pe = 32 b + 8 l + 4 u + i  <->  logical K = 16 (2 b + u) + 4 l + i.  A and B only
// have to agree on the order, so the permutation lives in the weight image.  CIN 16 keeps the natural order.
__host__ __device__ __forceinline__ int dg_logical_k(int pe, int cin) {
    if (cin < 32) return pe;
    const int b = pe >> 5, l = (pe >> 3) & 3, u = (pe >> 2) & 1, i = pe & 3;
    return 16 * (2 * b + u) + 4 * l + i;
}

// Packed weight image: per (super-tap, stage of 64 channels) one block of
//   [COUT x 64 bf16 W_hi] [COUT x 64 bf16 W_lo],  W_hi = bf16(W), W_lo = bf16(W - W_hi), 128-byte swizzle.
// Block size = 4 * COUT * 64 bytes.
__global__ void __launch_bounds__(256)
k_pack_weights_bf3(const float* __restrict__ w, int K, int cin, int cout, int tapg, int mode,
                   unsigned char* __restrict__ packed) {
    constexpr int ks = 64;
    const int cin_eff = cin * tapg;
    const int k_eff = (K + tapg - 1) / tapg;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)k_eff * cin_eff * cout;
    if (t >= total) return;
    const int n = (int)(t % cout);
    const int cv = (int)((t / cout) % cin_eff);             // virtual channel = g * cin + c
    const int kj = (int)(t / ((long long)cout * cin_eff));
    const int tap = kj * tapg + cv / cin, c = cv % cin;
    // source layout: mode -1: w[k][c][n] (K, cin, cout); modes 0-2 read the spconv-2 PARAMETER (Cout_p, K, Cin_p) in place:
    //   0 forward        (cin = Cin_p, cout = Cout_p):  w[k][c][n] = P[n][k][c]
    //   1 dgrad, SubM    (cin = Cout_p, cout = Cin_p):  w[k][c][n] = P[c][K-1-k][n]   (transposed, taps mirrored)
    //   2 dgrad, strided (cin = Cout_p, cout = Cin_p):  w[k][c][n] = P[c][k][n]
    float v = 0.f;
    if (tap < K) {
        if (mode < 0) v = w[((size_t)tap * cin + c) * cout + n];
        else if (mode == 0) v = w[((size_t)n * K + tap) * cin + c];
        else v = w[((size_t)c * K + (mode == 1 ? K - 1 - tap : tap)) * cout + n];
    }
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
    const int nstage = cin_eff / ks;
    const int stage = cv / ks;
    const int e = dg_logical_k(cv % ks, cin);          // K position inside the 64-wide stage (permuted for cin >= 32)
    __nv_bfloat16* blk = reinterpret_cast<__nv_bfloat16*>(packed + (size_t)(kj * nstage + stage) * ((size_t)4 * cout * ks));
    const int chunk = e >> 3, within = e & 7;               // 16-byte chunk = 8 bf16
    const size_t off = (size_t)n * ks + (size_t)((chunk ^ (n & 7)) * 8 + within);
    blk[off] = hi;
    blk[(size_t)cout * ks + off] = lo;
}

// fp32 rows -> "split" rows of the same size: every 4 consecutive channels c .. c+3 become one 16-byte group
//   { bf16x2(hi c, hi c+1), bf16x2(hi c+2, hi c+3), bf16x2(lo c, lo c+1), bf16x2(lo c+2, lo c+3) },   hi = bf16_rn(x), lo = bf16_rn(x - hi)
// i.e. exactly the registers a gather thread of k_spconv_umma_dg hands to tcgen05.st for its 4 channels.  The convolution
// gathers every input row once per kernel tap that uses it (~16x on average), and splitting after the gather costs 64
// conversion + 64 subtract / shift instructions per thread and stage -- the XU and ALU pipes, not the tensor or memory
// pipes, bound the kernel (profiles/r02_ncu_full_dg.md).  Splitting ONCE per element moves that work out of the gather.
__global__ void __launch_bounds__(256)
k_split_bf16(const float4* __restrict__ x, long long n4, uint4* __restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4; t += (long long)gridDim.x * blockDim.x) {
        const float4 v = x[t];
        const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
        const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
        const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
        uint4 o;
        o.x = *reinterpret_cast<const uint32_t*>(&h01);
        o.y = *reinterpret_cast<const uint32_t*>(&h23);
        o.z = *reinterpret_cast<const uint32_t*>(&l01);
        o.w = *reinterpret_cast<const uint32_t*>(&l23);
        out[t] = o;
    }
}

// [rows][K] -> [K][rows] (tap-major), so that the 128 indices of one (tile, tap) are one 512-byte segment; also
// tile_masks[row / 128] |= 1 << k for every tap k that occurs in the 128-row tile (zeroed by the launcher).
__global__ void __launch_bounds__(256)
k_transpose_table(const int* __restrict__ src, int n, int K, int ld, int* __restrict__ dst,
                  unsigned int* __restrict__ tile_masks) {
    __shared__ int tile[32][33];
    const int r0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i;
        tile[i][tx] = (r < n && tx < K) ? src[(size_t)r * K + tx] : -1;
    }
    __syncthreads();
    unsigned int m = 0u;
    const int k_pad = (K + 3) / 4 * 4;                 // rows [K, k_pad) = -1: multi-tap stages read whole tap groups
    for (int k = ty; k < k_pad; k += 8) {
        const int r = r0 + tx;
        const int v = k < K ? tile[tx][k] : -1;
        if (r < ld) dst[(size_t)k * ld + r] = v;
        const unsigned any = __ballot_sync(0xffffffffu, r < n && v >= 0);
        if (any) m |= 1u << k;
    }
    if (tile_masks && tx == 0 && m) atomicOr(tile_masks + (r0 >> 7), m);
}

// Builds the stage schedule of a CTA from the tap masks of its tiles (warp 0, lane == tap): for every tap present,
// for every channel chunk, one weight slice (b_sched) followed by one A stage per tile that has the tap (a_sched).
template <int NCHUNK>
__device__ __forceinline__ void build_schedule(const unsigned int* tile_mask_s, int tiles, int K, int lane,
                                               unsigned short* a_sched, unsigned short* b_sched,
                                               unsigned short* b_count, int* n_a_s, int* n_b_s) {
    unsigned int tm = 0u;                                 // bit t: tile t has tap `lane`
    if (lane < K)
        for (int t = 0; t < tiles; ++t) tm |= ((tile_mask_s[t] >> lane) & 1u) << t;
    const int cnt = __popc(tm);
    const unsigned active = __ballot_sync(0xffffffffu, cnt > 0);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int a_base = (incl - cnt) * NCHUNK;
    const int b_base = __popc(active & ((1u << lane) - 1u)) * NCHUNK;
    if (cnt > 0) {
        for (int ch = 0; ch < NCHUNK; ++ch) {
            b_sched[b_base + ch] = (unsigned short)(lane | (ch << 5));
            b_count[b_base + ch] = (unsigned short)cnt;
            unsigned int rem = tm;
            int j = 0;
            while (rem) {
                const int t = __ffs(rem) - 1;
                rem &= rem - 1;
                a_sched[a_base + ch * cnt + j] = (unsigned short)(lane | (ch << 5) | (t << 8));
                ++j;
            }
        }
    }
    if (lane == 31) *n_a_s = incl * NCHUNK;
    if (lane == 0) *n_b_s = __popc(active) * NCHUNK;
}

// ------------------------------------------------------------------------------------------------
// One CTA = `tiles` (<= 512/COUT, <= 8) consecutive 128-row output tiles, one TMEM accumulator each, so that a
// weight slice W[tap][32-channel chunk] is fetched once per CTA and reused by every tile that has the tap.
// Schedule (identical in all roles, built once in shared memory):

__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float even, float odd) {   // element 2j in the low half
    const __nv_bfloat162 t = __floats2bfloat162_rn(even, odd);
    return *reinterpret_cast<const uint32_t*>(&t);
}

// (f0, f1) -> hi = {bf16_rn(f1) : bf16_rn(f0)}, lo = {bf16_rn(f1 - hi1) : bf16_rn(f0 - hi0)} in six instructions (cvt pack, shift,
// and, two subtractions, cvt pack: three on the ALU pipe, three on the FMA pipe); the bfloat16 intrinsics round-trip through
// PRMT pairs and cost eight.  (sub.f32x2 for the two subtractions was measured too: the register-pair moves eat the gain.)
__device__ __forceinline__ void split_bf16_pair(float f0, float f1, uint32_t& hi, uint32_t& lo) {
    uint32_t h, l;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(f1), "f"(f0));
    const float r0 = f0 - __uint_as_float(h << 16);
    const float r1 = f1 - __uint_as_float(h & 0xffff0000u);
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(r1), "f"(r0));
    hi = h;
    lo = l;
}


// memory.  ncu showed the L1/shared-memory data pipe (not the tensor pipe) as the busiest unit of the TS kernels: per
// 64-channel stage 32 KB of gathered rows were written to shared memory by cp.async, read back by the transform warps and
// 24 KB of weight slices re-read by the MMAs.  Here a gather warp loads its rows straight into registers in the fragment
// layout of tcgen05.st.16x256b (measured with tools/tmem_layout_probe.cu: thread t holds TMEM lanes t/4 and t/4+8,
// columns 2(t%4), 2(t%4)+1 of every 8-column unit), so 4 lanes read 64 contiguous bytes of a row (whole 32-byte sectors),
// splits them into bf16 hi/lo in registers and stores them to the stage's TMEM columns.
//   warps 0-15 : four gather sets (set = stage % 4 = TMEM A slot; warp & 3 = TMEM lane quadrant); epilogue afterwards
//   warp 16    : MMA issuer          warp 17 : weight slices (bulk copies)          warp 18 : row-index segments (bulk copies)
// -DUD_DG_PROBE: cycle accounting of the roles of k_spconv_umma_dg (tools/gpu_dg_probe.py builds such a library; the shipped
// one has none of it).  Slots: 0 stages seen by gather warp 0 of a set, 1 index wait, 2 index read + load issue, 3 data wait +
// conversion, 4 TMEM slot wait, 5 TMEM stores + arrive | 8 MMA stages, 9 wait weights, 10 wait A, 11 issue + commit |
// 12 CTAs, 13 CTA cycles from start to the end of the main loop (MMA warp), 14 epilogue cycles (warp 0) | 6 the MMA warp's
// tcgen05.fence::after_thread_sync after its A wait
#ifdef UD_DG_TRACE
// event trace of ONE CTA (blockIdx.x == UD_DG_TRACE): [event][stage] = clock64().  Events: 0 set: indices landed, 1 set: rows
// converted (ready to store), 2 set: TMEM slot free, 3 set: arrived | 4 MMA warp: A full, 5 MMA warp: stage issued + committed,
// 6 MMA warp: weights full
#define DG_TRACE_STAGES 128
__device__ long long ud_dg_trace[8][DG_TRACE_STAGES];
#define DG_TRACE(ev, st, cond) if (blockIdx.x == UD_DG_TRACE && (st) < DG_TRACE_STAGES && (cond)) ud_dg_trace[ev][st] = clock64();
#else
#define DG_TRACE(ev, st, cond)
#endif
#ifdef UD_DG_PROBE
__device__ unsigned long long ud_dg_probe_acc[16];
#define DG_PROBE_DECL(n) long long pr_[n] = {}; long long pt_ = clock64();
#define DG_PROBE_LAP(i) { const long long now_ = clock64(); pr_[i] += now_ - pt_; pt_ = now_; }
#define DG_PROBE_FLUSH(base, n, cond) if (cond) { for (int i_ = 0; i_ < (n); ++i_) atomicAdd(&ud_dg_probe_acc[(base) + i_], (unsigned long long)pr_[i_]); }
#else
#define DG_PROBE_DECL(n)
#define DG_PROBE_LAP(i)
#define DG_PROBE_FLUSH(base, n, cond)
#endif

// A gather warp hands its part of an A stage to the MMA warp: ONE mbarrier arrive per warp.  (Every thread arriving on its
// own -- 128 same-address arrivals per stage, 32 of them serialised inside each warp instruction -- delayed the phase flip
// by ~150 cycles per stage, all of it on the MMA warp's critical path: profiles/r03_pipeline_probes.md.)  tcgen05.wait::st
// is warp-collective, so after it every lane's TMEM stores are complete; the fence orders them before the arrive.
__device__ __forceinline__ void dg_set_arrive(uint32_t bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar);
}
#if defined(UD_MBAR_SPIN_MMA)
#define DG_MMA_WAIT mbar_wait_spin
#else
#define DG_MMA_WAIT mbar_wait
#endif
constexpr int DG_SETS = 4;
// Per-CTA configuration of k_spconv_umma_dg, template parameter NSETS = gather sets = TMEM A slots:
//   NSETS 4: one CTA per SM -- 512 TMEM columns (256 accumulators | 4 x 64 A ring), 16 gather warps + MMA + weights + indices warp
//   NSETS 2: TWO CTAs per SM -- 256 TMEM columns each (128 accumulators | 2 x 64 A ring), 8 gather warps + MMA + one loader warp.
// Why two: the event trace of a CTA (profiles/r03_pipeline_probes.md) shows the single MMA-issuing warp as the serial
// bottleneck of the wide layers -- ~660-730 cycles of dependent instruction latency per stage (barrier tests, schedule reads,
// R2UR moves, 12 UTCHMMA + commit) for 390 cycles of tensor work -- while gather sets and tensor pipe both have slack.  Two
// resident CTAs give the SM two independent issuers, and one CTA's epilogue / prologue overlaps the other's main loop.
__host__ __device__ constexpr int dg_threads(int nsets) { return nsets == 4 ? 608 : (4 * nsets + 2) * 32; }
__host__ __device__ constexpr int dg_acc_cols(int nsets) { return 64 * nsets; }
__host__ __device__ constexpr int dg_tiles_max(int nsets, int cout) { return dg_acc_cols(nsets) / cout < 8 ? dg_acc_cols(nsets) / cout : 8; }
__host__ __device__ constexpr int dg_b_stages(int nsets, int cout) { return (nsets == 2 && cout >= 128) ? 2 : 4; }
constexpr int DG_IDX_STAGES = 16;
// statistics buffer: [0, 4096) final sums (2 * COUT doubles), then the workspace of det_reduce.cuh
constexpr int DG_STAT_FINAL_BYTES = 4096;
// Optional output transform of the forward kernels (inference: the eval-mode BatchNorm1d, residual add and ReLU that follow a
// convolution in spconv_backbone.py:50-56,97-113 folded into its epilogue): out = relu?((acc + bias[c]) * scale[c] + shift[c] (+ residual)).
// Same arithmetic as k_bn_apply (fmaf, then the residual, then the clamp), so fused and unfused results are bit-identical.
struct DgPost {
    const float* scale;        // (cout) or nullptr = no transform (bias path)
    const float* shift;        // (cout)
    const float* residual;     // (n_out, cout) or nullptr
    int relu;
};

// staging buffer of the epilogue (dg_epilogue): one [32 rows x 36 floats] tile per epilogue warp
constexpr int DG_EP_PITCH = 36;                          // floats per staged row: 16-byte aligned, conflict-free LDS/STS.128
constexpr int DG_EP_STAGE_FLOATS = 32 * DG_EP_PITCH;
constexpr size_t DG_EP_STAGE_BYTES = (size_t)16 * DG_EP_STAGE_FLOATS * 4;
__host__ __device__ constexpr size_t dg_ep_stage_bytes(int nsets) { return (size_t)4 * nsets * DG_EP_STAGE_FLOATS * 4; }
__host__ __device__ constexpr size_t dg_smem_bytes(int cin, int cout, int nsets = 4) {
    // the epilogue's staging tiles alias the weight ring when that is large enough (it is idle by then), else follow the index ring
    return (size_t)dg_b_stages(nsets, cout) * dg_b_bytes(cout) + (size_t)DG_IDX_STAGES * ts_tapg(cin) * TILE_M * 4 +
           ((size_t)dg_b_stages(nsets, cout) * dg_b_bytes(cout) >= dg_ep_stage_bytes(nsets) ? 0 : dg_ep_stage_bytes(nsets)) + 1024;
}

__device__ __forceinline__ uint4 ldg_stream(const float* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
// 256-bit variant: four lanes cover a whole 128-byte line of a row, so L1 / L2 see one full-line request per row instead of
// two half-line ones (the gather is limited by the L2 -> SM request rate, profiles/r02_unit_major_schedule.md)
__device__ __forceinline__ void ldg_stream256(const float* p, uint4& a, uint4& b) {
    asm volatile("ld.global.nc.L1::evict_last.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p));
}
__device__ __forceinline__ void tmem_st_16x256b_x4(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// Epilogue shared by the tensor-core forward kernels (called by ALL threads; warps 0-15 do the work: warp w -> TMEM lane
// quadrant w & 3, tiles (w >> 2), (w >> 2) + 4, ...): tcgen05.ld -> shared-memory transpose -> + bias -> st.global.v4, plus the
// per-channel sum / sum of squares of the output for the BatchNorm that follows, reduced in a fixed order (det_reduce.cuh).
// The accumulator chunk a warp has read (`stage`: DG_EP_STAGE_FLOATS floats per warp) -- thread = row,
// 32 consecutive columns -- is transposed through shared memory so that a store instruction writes 4 rows x 128 contiguous
// bytes instead of 32 rows x 16 bytes (the role-cycle probe showed the epilogue at ~20 % of a CTA's life, bound by the 32
// partial lines per store), and the per-column sums need 2 shuffle rounds instead of 5.

template <int COUT, int NTHREADS, int NSETS = 4>
__device__ __forceinline__ void dg_epilogue(int tid, int n_a, uint32_t bar_acc, int tiles, int rows_cta, int row0,
                                            const unsigned int* tile_mask_s, uint32_t tmem_acc,
                                            const float* __restrict__ bias, float* __restrict__ out,
                                            unsigned char* __restrict__ stats, float* wpart_s, int stat_group,
                                            int* stat_last_s, float* stage, const DgPost post = DgPost{nullptr, nullptr, nullptr, 0}) {
    const int warp = tid >> 5;
    const int lane = tid & 31;
    // ============ epilogue (warps 0 .. 4 NSETS - 1: warp w -> quadrant w&3, tiles (w>>2), (w>>2)+NSETS, ...) ============
    if (warp < 4 * NSETS) {
        if (n_a > 0) {
            mbar_wait(bar_acc, 0u);
            tc_fence_after();
        }
        const int q = warp & 3;
        constexpr int CW = COUT >= 32 ? 32 : 16;         // columns per chunk
        constexpr int LPR = CW / 4;                      // lanes per row when reading back (one float4 each)
        constexpr int RPI = 32 / LPR;                    // rows per store instruction
        float* const st = stage + warp * DG_EP_STAGE_FLOATS;
        const int rsub = lane / LPR;
        const int c4 = (lane % LPR) * 4;
#pragma unroll 1
        for (int t = warp >> 2; t < tiles; t += NSETS) {
            if (t * TILE_M >= rows_cta) break;
            const bool has_acc = tile_mask_s[t] != 0u;
            const int r_base = t * TILE_M + q * 32;       // first row (CTA-local) of this warp's 32
#pragma unroll 1
            for (int c0 = 0; c0 < COUT; c0 += CW) {
                uint32_t v[32];
                if (has_acc) {
                    const uint32_t taddr = tmem_acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * COUT + c0);
                    if constexpr (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int i = 0; i < CW; ++i) v[i] = 0u;
                }
                __syncwarp();                             // the previous chunk has been read back
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                    *reinterpret_cast<uint4*>(st + lane * DG_EP_PITCH + i) = make_uint4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                float4 b = make_float4(0.f, 0.f, 0.f, 0.f), sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = b;
                if (bias) b = __ldg(reinterpret_cast<const float4*>(bias + c0 + c4));
                if (post.scale) {
                    sc = __ldg(reinterpret_cast<const float4*>(post.scale + c0 + c4));
                    sh = __ldg(reinterpret_cast<const float4*>(post.shift + c0 + c4));
                }
                float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < 32 / RPI; ++j) {
                    const int r = j * RPI + rsub;
                    float4 o = *reinterpret_cast<const float4*>(st + r * DG_EP_PITCH + c4);
                    o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                    if (post.scale) {
                        o.x = fmaf(o.x, sc.x, sh.x); o.y = fmaf(o.y, sc.y, sh.y);
                        o.z = fmaf(o.z, sc.z, sh.z); o.w = fmaf(o.w, sc.w, sh.w);
                    }
                    if (r_base + r < rows_cta) {
                        if (post.residual) {
                            const float4 rs = __ldg(reinterpret_cast<const float4*>(
                                post.residual + (size_t)(row0 + r_base + r) * COUT + c0 + c4));
                            o.x += rs.x; o.y += rs.y; o.z += rs.z; o.w += rs.w;
                        }
                        if (post.relu) {
                            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
                        }
                        *reinterpret_cast<float4*>(out + (size_t)(row0 + r_base + r) * COUT + c0 + c4) = o;
                        sa.x += o.x; sa.y += o.y; sa.z += o.z; sa.w += o.w;
                        sq.x += o.x * o.x; sq.y += o.y * o.y; sq.z += o.z * o.z; sq.w += o.w * o.w;
                    }
                }
                if (stats != nullptr) {
#pragma unroll
                    for (int sft = LPR; sft < 32; sft <<= 1) {      // fold the RPI row groups: fixed order
                        sa.x += __shfl_xor_sync(0xffffffffu, sa.x, sft); sa.y += __shfl_xor_sync(0xffffffffu, sa.y, sft);
                        sa.z += __shfl_xor_sync(0xffffffffu, sa.z, sft); sa.w += __shfl_xor_sync(0xffffffffu, sa.w, sft);
                        sq.x += __shfl_xor_sync(0xffffffffu, sq.x, sft); sq.y += __shfl_xor_sync(0xffffffffu, sq.y, sft);
                        sq.z += __shfl_xor_sync(0xffffffffu, sq.z, sft); sq.w += __shfl_xor_sync(0xffffffffu, sq.w, sft);
                    }
                    if (lane < LPR) {                     // the warp's own slot, tiles in ascending order: no atomics
                        float* pa = wpart_s + (warp * 2 + 0) * COUT + c0 + c4;
                        float* pq = wpart_s + (warp * 2 + 1) * COUT + c0 + c4;
                        pa[0] += sa.x; pa[1] += sa.y; pa[2] += sa.z; pa[3] += sa.w;
                        pq[0] += sq.x; pq[1] += sq.y; pq[2] += sq.z; pq[3] += sq.w;
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (stats != nullptr) {
        // 16 warp slots -> this CTA's partial (fixed order), then the fixed-order cross-CTA reduction of det_reduce.cuh:
        // final[2 * COUT] doubles = per-channel sum | sum of squares of `out`, bit-reproducible from run to run
        float v = 0.f;
        if (tid < 2 * COUT) {
#pragma unroll
            for (int w = 0; w < 4 * NSETS; ++w) v += wpart_s[(w * 2 + tid / COUT) * COUT + tid % COUT];
        }
        ud_det_reduce((double)v, tid, 2 * COUT, stats + DG_STAT_FINAL_BYTES, stat_group,
                      reinterpret_cast<double*>(stats), stat_last_s);
        __syncthreads();
    }
}

template <int CIN, int COUT, bool SPLIT, int NSETS>
__global__ void __launch_bounds__(dg_threads(NSETS), NSETS == 4 ? 1 : 2)
k_spconv_umma_dg(const __grid_constant__ CUtensorMap idx_map, const float* __restrict__ in,
                 const float* __restrict__ wpacked, const float* __restrict__ bias,
                 const int* __restrict__ nbr_t, const unsigned int* __restrict__ tile_masks, int n_out, int K,
                 int tiles, float* __restrict__ out, unsigned char* __restrict__ stats, int stat_group, const DgPost post) {
    constexpr int TAPG = ts_tapg(CIN);               // kernel taps per 64-channel stage
    constexpr int KS = 64;
    constexpr int NCHUNK = CIN * TAPG / KS;          // stages per (super-)tap
    constexpr int IDX_INTS = TAPG * TILE_M;
    constexpr int A_COLS = KS;                       // TMEM columns of one A stage: 32 hi | 32 lo (bf16 pairs)
    constexpr int A_STAGES = NSETS;                  // one gather set per TMEM A slot
    constexpr int ACC_COLS = dg_acc_cols(NSETS);     // accumulators [0, ACC_COLS), A ring [ACC_COLS, 2 ACC_COLS)
    constexpr int NTHREADS = dg_threads(NSETS);
    constexpr int GW = 4 * NSETS;                    // gather (and epilogue) warps; warp GW issues the MMAs
    constexpr int DG_B_STAGES = dg_b_stages(NSETS, COUT);
    constexpr int B_BYTES = dg_b_bytes(COUT);
    constexpr int BBF_BYTES = COUT * KS * 2;
    constexpr uint32_t IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                  ((uint32_t)(TILE_M >> 4) << 24);
    static_assert(NSETS == 4 || NSETS == 2, "4 gather sets (one CTA per SM) or 2 (two CTAs per SM)");
    static_assert(A_STAGES * A_COLS == ACC_COLS, "A ring and accumulators split the allocation in halves");
    static_assert((CIN * TAPG) % KS == 0 && COUT % 16 == 0 && COUT >= 16 && COUT <= 256, "unsupported channel count");
    static_assert(MAX_TAPS * NCHUNK * dg_tiles_max(NSETS, COUT) <= MAX_SCHED, "schedule list too small");
    constexpr int NBAR = 2 * A_STAGES + 2 * DG_B_STAGES + 2 * DG_IDX_STAGES + 1;

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[NBAR];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int tile_mask_s[MAX_TILES];
    __shared__ unsigned short a_sched[MAX_SCHED];          // tap | chunk << 5 | tile << 8
    __shared__ unsigned short b_sched[MAX_TAPS * 4];
    __shared__ unsigned short b_count[MAX_TAPS * 4];
    __shared__ int n_a_s, n_b_s;
    __shared__ float wpart_s[GW * 2 * COUT];               // fused BatchNorm statistics: per epilogue warp, sum | sum of squares
    __shared__ int stat_last_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int row0 = blockIdx.x * tiles * TILE_M;
    const int rows_cta = min(tiles * TILE_M, n_out - row0);
    if (stats != nullptr)
        for (int i = tid; i < GW * 2 * COUT; i += NTHREADS) wpart_s[i] = 0.f;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t smem_b = smem_base;
    const uint32_t smem_idx = smem_b + DG_B_STAGES * B_BYTES;
    const int* const idx_ring = reinterpret_cast<const int*>(smem_al + DG_B_STAGES * B_BYTES);
    const uint32_t bar_a_full = smem_u32(&bars[0]);
    const uint32_t bar_a_empty = bar_a_full + 8 * A_STAGES;
    const uint32_t bar_b_full = bar_a_empty + 8 * A_STAGES;
    const uint32_t bar_b_empty = bar_b_full + 8 * DG_B_STAGES;
    const uint32_t bar_idx_full = bar_b_empty + 8 * DG_B_STAGES;
    const uint32_t bar_idx_empty = bar_idx_full + 8 * DG_IDX_STAGES;
    const uint32_t bar_acc = bar_idx_empty + 8 * DG_IDX_STAGES;

    if (tid == 0) {
        for (int s = 0; s < A_STAGES; ++s) {
            mbar_init(bar_a_full + 8 * s, TS_GROUP / 32);    // lane 0 of each warp of the owning gather set (see dg_set_arrive)
            mbar_init(bar_a_empty + 8 * s, 1);
        }
        for (int s = 0; s < DG_B_STAGES; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, 1);
        }
        for (int s = 0; s < DG_IDX_STAGES; ++s) {
            mbar_init(bar_idx_full + 8 * s, 1);
            mbar_init(bar_idx_empty + 8 * s, TS_GROUP / 32);   // lane 0 of each warp of the consuming gather set
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    const int K_eff = (K + TAPG - 1) / TAPG;
    if (tid < MAX_TILES) {
        const int tile_g = blockIdx.x * tiles + tid;
        unsigned int m = (tid < tiles && tile_g * TILE_M < n_out) ? __ldg(tile_masks + tile_g) : 0u;
        if (TAPG > 1) {
            unsigned int me = 0u;
            for (int j = 0; j < K_eff; ++j)
                if ((m >> (j * TAPG)) & ((1u << TAPG) - 1u)) me |= 1u << j;
            m = me;
        }
        tile_mask_s[tid] = m;
    }
    if (warp == GW) tmem_alloc(smem_u32(&tmem_base_s), 2 * ACC_COLS);
    __syncthreads();
    if (warp == 0) build_schedule<NCHUNK>(tile_mask_s, tiles, K_eff, lane, a_sched, b_sched, b_count, &n_a_s, &n_b_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_a = n_a_s;
    const int n_b = n_b_s;
    const uint32_t tmem_acc = tmem_base_s;
    const uint32_t tmem_a = tmem_acc + ACC_COLS;

    if (warp < GW) {
        // ============ gather sets: global -> registers (st.16x256b fragment layout) -> bf16 hi/lo -> TMEM ============
        const int set = warp >> 2;
        const int q = warp & 3;
        const int rsub = lane >> 2;                      // row within an 8-row group
        const int csub = (lane & 3) * 4;                 // first of this thread's 4 channels within a 16-channel unit
        DG_PROBE_DECL(6)
        for (int it = set; it < n_a; it += NSETS) {
            const int islot = it % DG_IDX_STAGES;
#ifdef UD_DG_PROBE
            pr_[0] += 1;
            pt_ = clock64();
#endif
            mbar_wait(bar_idx_full + 8 * islot, ((uint32_t)(it / DG_IDX_STAGES)) & 1u);
            DG_PROBE_LAP(1)
            DG_TRACE(0, it, q == 0 && lane == 0)
            // rows 32 q + 16 h + rsub + 8 rr of the tile, for every tap of the stage
            int idx[2][2][TAPG];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                    for (int g = 0; g < TAPG; ++g)
                        idx[h][rr][g] = idx_ring[islot * IDX_INTS + g * TILE_M + q * 32 + h * 16 + rr * 8 + rsub];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_idx_empty + 8 * islot);
            const int c0 = TAPG == 1 ? ((a_sched[it] >> 5) & 7) * KS : 0;
            uint4 x[2][2][4];
            if constexpr (TAPG <= 2) {
                // CIN >= 32: one 32-byte load per lane and 32-channel block.  Lane l of a quad reads PHYSICAL channels
                // [8 l, 8 l + 8) of the block and uses them as LOGICAL K positions 4 l .. 4 l + 3 of the block's two
                // 16-channel units; the weight image is packed in the same permuted K order (dg_logical_k).
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j2 = 0; j2 < 2; ++j2) {
                            const int g = TAPG == 1 ? 0 : j2;                     // TAPG 2 (CIN 32): block j2 = tap j2
                            const int ch = (TAPG == 1 ? c0 + 32 * j2 : 0) + 2 * csub;
                            const int r = idx[h][rr][g];
                            x[h][rr][2 * j2] = make_uint4(0u, 0u, 0u, 0u);
                            x[h][rr][2 * j2 + 1] = make_uint4(0u, 0u, 0u, 0u);
#if !defined(UD_DG_EXP) || (UD_DG_EXP != 1 && UD_DG_EXP != 3 && UD_DG_EXP != 4)
                            if (r >= 0) ldg_stream256(in + (size_t)r * CIN + ch, x[h][rr][2 * j2], x[h][rr][2 * j2 + 1]);
#else
                            if (r == -12345) ldg_stream256(in + (size_t)r * CIN + ch, x[h][rr][2 * j2], x[h][rr][2 * j2 + 1]);
#endif
                        }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            // TAPG 4 (CIN 16): unit j = tap j, 64-byte rows: 16-byte loads, natural K order
                            const int r = idx[h][rr][j];
                            x[h][rr][j] = make_uint4(0u, 0u, 0u, 0u);
#if !defined(UD_DG_EXP) || (UD_DG_EXP != 1 && UD_DG_EXP != 3 && UD_DG_EXP != 4)
                            if (r >= 0) x[h][rr][j] = ldg_stream(in + (size_t)r * CIN + csub);
#else
                            if (r == -12345) x[h][rr][j] = ldg_stream(in + (size_t)r * CIN + csub);
#endif
                        }
            }
            const int aslot = it % A_STAGES;             // == set
            DG_PROBE_LAP(2)
            uint32_t hi[2][16], lo[2][16];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const uint4 v = x[h][rr][j];
                        if constexpr (SPLIT) {           // `in` holds pre-split rows (k_split_bf16): nothing to convert
                            hi[h][4 * j + 2 * rr + 0] = v.x;
                            hi[h][4 * j + 2 * rr + 1] = v.y;
                            lo[h][4 * j + 2 * rr + 0] = v.z;
                            lo[h][4 * j + 2 * rr + 1] = v.w;
                        } else {
                            split_bf16_pair(__uint_as_float(v.x), __uint_as_float(v.y), hi[h][4 * j + 2 * rr + 0],
                                            lo[h][4 * j + 2 * rr + 0]);
                            split_bf16_pair(__uint_as_float(v.z), __uint_as_float(v.w), hi[h][4 * j + 2 * rr + 1],
                                            lo[h][4 * j + 2 * rr + 1]);
                        }
                    }
#ifdef UD_DG_PROBE
            {   // the lap must not start before the converted values exist
                uint32_t dep = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int i = 0; i < 16; ++i) dep ^= hi[h][i] ^ lo[h][i];
                asm volatile("" :: "r"(dep));
            }
#endif
            DG_PROBE_LAP(3)
            DG_TRACE(1, it, q == 0 && lane == 0)
            // the conversions do not need the TMEM slot: only the stores wait for the MMAs that read it last
            mbar_wait(bar_a_empty + 8 * aslot, (((uint32_t)(it / A_STAGES)) & 1u) ^ 1u);
            tc_fence_after();
            DG_PROBE_LAP(4)
            DG_TRACE(2, it, q == 0 && lane == 0)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint32_t taddr = tmem_a + (uint32_t)(aslot * A_COLS) + ((uint32_t)(q * 32 + h * 16) << 16);
#if defined(UD_DG_EXP) && (UD_DG_EXP == 3 || UD_DG_EXP == 4)
                if (hi[h][0] == 0x12345u) {
#endif
                tmem_st_16x256b_x4(taddr, hi[h]);
                tmem_st_16x256b_x4(taddr + KS / 2, lo[h]);
#if defined(UD_DG_EXP) && (UD_DG_EXP == 3 || UD_DG_EXP == 4)
                }
#endif
            }
            tmem_st_wait();
            dg_set_arrive(bar_a_full + 8 * aslot, lane);
            DG_PROBE_LAP(5)
            DG_TRACE(3, it, q == 0 && lane == 0)
        }
        DG_PROBE_FLUSH(0, 6, q == 0 && lane == 0)
    } else if (warp == GW) {
        // ============ MMA issuer ============
        unsigned int started = 0u;
        int it = 0;
        DG_PROBE_DECL(6)
#ifdef UD_DG_PROBE
        const long long cta_t0 = pt_;
#endif
        for (int ib = 0; ib < n_b; ++ib) {
            const int bslot = ib % DG_B_STAGES;
            DG_MMA_WAIT(bar_b_full + 8 * bslot, ((uint32_t)(ib / DG_B_STAGES)) & 1u);
            DG_PROBE_LAP(1)
            DG_TRACE(6, it, lane == 0)
            const uint32_t b_base = smem_b + bslot * B_BYTES;
            const uint64_t db_hi = umma_desc_sw128(b_base);
            const uint64_t db_lo = umma_desc_sw128(b_base + BBF_BYTES);
            const int cnt = b_count[ib];
            for (int c = 0; c < cnt; ++c, ++it) {
                const int aslot = it % A_STAGES;
                const int tile = a_sched[it] >> 8;
                DG_MMA_WAIT(bar_a_full + 8 * aslot, ((uint32_t)(it / A_STAGES)) & 1u);
                DG_PROBE_LAP(2)
                DG_TRACE(4, it, lane == 0)
                tc_fence_after();
                DG_PROBE_LAP(4)
                if (elect_one_sync()) {
                    const uint32_t a_stage = tmem_a + (uint32_t)(aslot * A_COLS);
                    const uint32_t d = tmem_acc + (uint32_t)(tile * COUT);
                    const uint32_t first = (started >> tile) & 1u;
#pragma unroll
                    for (int k16 = 0; k16 < KS / 16; ++k16) {
                        const uint32_t a_hi = a_stage + (uint32_t)(k16 * 8);
                        const uint32_t a_lo = a_hi + KS / 2;
                        const uint64_t adv = (uint64_t)(k16 * 2);
#if defined(UD_DG_EXP) && UD_DG_EXP == 2
                        if (n_out == -7)
#endif
                        {
                        umma_bf16_ts(d, a_lo, db_hi + adv, IDESC_BF, (first || k16 > 0) ? 1u : 0u);
                        umma_bf16_ts(d, a_hi, db_lo + adv, IDESC_BF, 1u);
                        umma_bf16_ts(d, a_hi, db_hi + adv, IDESC_BF, 1u);
                        }
                    }
                    umma_commit(bar_a_empty + 8 * aslot);
                }
                started |= 1u << tile;
                __syncwarp();
#ifdef UD_DG_PROBE
                pr_[0] += 1;
#endif
                DG_PROBE_LAP(3)
                DG_TRACE(5, it, lane == 0)
#if defined(UD_DG_EXP) && UD_DG_EXP == 4
                // serialised: wait for this stage's own MMAs (slot 5 = cycles from the end of the issue to their completion)
                mbar_wait(bar_a_empty + 8 * aslot, ((uint32_t)(it / A_STAGES)) & 1u);
                DG_PROBE_LAP(5)
#endif
            }
            if (elect_one_sync()) umma_commit(bar_b_empty + 8 * bslot);
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(bar_acc);
        __syncwarp();
#ifdef UD_DG_PROBE
        if (lane == 0) {
            atomicAdd(&ud_dg_probe_acc[6], (unsigned long long)pr_[4]);      // tcgen05.fence::after_thread_sync after the A wait
            atomicAdd(&ud_dg_probe_acc[15], (unsigned long long)pr_[5]);     // UD_DG_EXP 4: end of issue -> completion of the stage's MMAs
            atomicAdd(&ud_dg_probe_acc[12], 1ull);
            atomicAdd(&ud_dg_probe_acc[13], (unsigned long long)(clock64() - cta_t0));
        }
#endif
        DG_PROBE_FLUSH(8, 4, lane == 0)
    } else {
        // ============ weight slices and per-stage row-index segments (bulk async copies): lane 0 of warps GW + 1 and GW + 2
        // (NSETS 4), lanes 0 and 1 of warp GW + 1 (NSETS 2: one warp less so that two CTAs fit the register file) ============
        const bool w_role = warp == GW + 1 && lane == 0;
        const bool i_role = NSETS == 4 ? (warp == GW + 2 && lane == 0) : (warp == GW + 1 && lane == 1);
        if (w_role) {
            for (int ib = 0; ib < n_b; ++ib) {
                const int bslot = ib % DG_B_STAGES;
                mbar_wait(bar_b_empty + 8 * bslot, (((uint32_t)(ib / DG_B_STAGES)) & 1u) ^ 1u);
                const unsigned e = b_sched[ib];
                const int tap = e & 31, ch = e >> 5;
                const float* src = wpacked + (size_t)(tap * NCHUNK + ch) * (B_BYTES / 4);
                mbar_arrive_expect_tx(bar_b_full + 8 * bslot, B_BYTES);
                bulk_g2s(smem_b + bslot * B_BYTES, src, B_BYTES, bar_b_full + 8 * bslot);
            }
        } else if (i_role) {
            // row-index segments of every stage: ONE TMA tensor load per stage -- a [TAPG taps x 128 rows] box of the tap-major
            // table (2-D tensor map over nbr_t: K4 x ld int32), landing as TAPG consecutive 512-byte segments
            tma_prefetch_desc(&idx_map);
            for (int it = 0; it < n_a; ++it) {
                const int islot = it % DG_IDX_STAGES;
                mbar_wait(bar_idx_empty + 8 * islot, (((uint32_t)(it / DG_IDX_STAGES)) & 1u) ^ 1u);
                const unsigned e = a_sched[it];
                mbar_arrive_expect_tx(bar_idx_full + 8 * islot, IDX_INTS * 4);
                tma_load_2d(smem_idx + islot * IDX_INTS * 4, &idx_map, row0 + (int)(e >> 8) * TILE_M, (int)(e & 31) * TAPG,
                            bar_idx_full + 8 * islot);
            }
        }
        __syncwarp();
    }

#ifdef UD_DG_PROBE
    const long long ep_t0 = clock64();
#endif
    dg_epilogue<COUT, NTHREADS, NSETS>(tid, n_a, bar_acc, tiles, rows_cta, row0, tile_mask_s, tmem_acc, bias, out, stats, wpart_s,
                                       stat_group, &stat_last_s,
                                       reinterpret_cast<float*>(smem_al + ((size_t)DG_B_STAGES * B_BYTES >= dg_ep_stage_bytes(NSETS)
                                                                               ? 0 : DG_B_STAGES * B_BYTES + DG_IDX_STAGES * IDX_INTS * 4)),
                                       post);
#ifdef UD_DG_PROBE
    if (tid == 0) atomicAdd(&ud_dg_probe_acc[14], (unsigned long long)(clock64() - ep_t0));
#endif
    if (warp == GW) tmem_dealloc(tmem_acc, 2 * ACC_COLS);
}

// =====================================================================================================================
// Window-staged variant for 3x3x3 kernels with 64 input channels (k_spconv_umma_win).
//
// Under the raster row order (rows sorted by (b, z, y, x)) the rows that the three kx taps of one (kz, ky) pair gather for
// a 128-row tile are x-adjacent voxels of neighbouring lines, i.e. they lie in a narrow index window: measured on the
// bench clouds, 97 % of them inside [lo, lo + 160) with lo = the smallest gathered row (tools/window_coverage_probe.py,
// profiles/r02_window_probe.txt).  So instead of 3 x 128 per-thread gathers from global memory, TMA stages that window
// ONCE per (tile, kz, ky) -- two tensor loads of a [160 rows x 32 channels] box, 128-byte swizzle -- and the gather warps
// build the three taps' operands from shared memory (conflict-free LDS.128 thanks to the swizzle); rows outside the
// window (3 %) still come from global memory.  Global gather traffic drops ~3x and, more to the point, stops depending on
// how many loads an SM can keep in flight (the limit of k_spconv_umma_dg, profiles/r02_ncu_full_dg.md).
// Schedule = unit major: unit = (kz, ky) group; inside a unit tile by tile, the tile's three kx stages back to back; the
// unit's three weight slices stay resident in a 4-slot ring (a slice is released after its last tile).
//   warps 0-15 gather sets | 16 MMA issuer | 17 weight slices | 18 row-index segments (TMA) | 19 windows (TMA)
constexpr int WIN_ROWS = 160;
constexpr int WIN_SLOTS = 3;
constexpr int WIN_THREADS = 640;
constexpr int WIN_B_RING = 4;
constexpr int WIN_HALF_BYTES = WIN_ROWS * 128;          // one 32-channel half of a window, 128-byte rows
constexpr int WIN_BYTES = 2 * WIN_HALF_BYTES;
constexpr int WIN_GROUPS = 9;
constexpr int WIN_CIN = 64;
__host__ __device__ constexpr size_t win_smem_bytes(int cout) {
    return (size_t)WIN_B_RING * dg_b_bytes(cout) + (size_t)DG_IDX_STAGES * TILE_M * 4 + (size_t)WIN_SLOTS * WIN_BYTES + 1024;
}

// lane g < 9 owns group g (taps 3g .. 3g+2).  A group that occurs in any tile gets all three weight slices, a tile that has
// any tap of the group gets all three stages (an absent tap is an all -1 index segment = a zero operand), so every
// (group, tile) window is read by exactly 3 stages x 4 warps.
//   b_sched[ib] = tap;  w_sched[u] = group | tile << 4;  a_sched[it] = tap | tile << 8 | release << 11;  a_slice[it] = ib
__device__ __forceinline__ void build_schedule_win(const unsigned int* tile_mask_s, int tiles, int lane,
                                                   unsigned short* a_sched, unsigned short* a_slice,
                                                   unsigned short* b_sched, unsigned short* w_sched, int* n_a_s,
                                                   int* n_b_s) {
    unsigned int gm = 0u;                                 // bit t: tile t has some tap of group `lane`
    if (lane < WIN_GROUPS)
        for (int t = 0; t < tiles; ++t) gm |= (((tile_mask_s[t] >> (3 * lane)) & 7u) != 0u ? 1u : 0u) << t;
    const int n_ut = __popc(gm);
    const int n_gr = gm != 0u;
    int incl_u = n_ut, incl_g = n_gr;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int vu = __shfl_up_sync(0xffffffffu, incl_u, o);
        const int vg = __shfl_up_sync(0xffffffffu, incl_g, o);
        if (lane >= o) { incl_u += vu; incl_g += vg; }
    }
    if (gm) {
        const int b_base = 3 * (incl_g - 1);
        int u = incl_u - n_ut;
        const unsigned int last = 31u - (unsigned int)__clz(gm);
        for (int x = 0; x < 3; ++x) b_sched[b_base + x] = (unsigned short)(3 * lane + x);
        for (int t = 0; t < tiles; ++t) {
            if (!((gm >> t) & 1u)) continue;
            w_sched[u] = (unsigned short)(lane | (t << 4));
            for (int x = 0; x < 3; ++x) {
                a_sched[3 * u + x] = (unsigned short)((3 * lane + x) | (t << 8) | ((last == (unsigned)t) << 11));
                a_slice[3 * u + x] = (unsigned short)(b_base + x);
            }
            ++u;
        }
    }
    if (lane == 31) {
        *n_a_s = 3 * incl_u;
        *n_b_s = 3 * incl_g;
    }
}

template <int COUT>
__global__ void __launch_bounds__(WIN_THREADS, 1)
k_spconv_umma_win(const __grid_constant__ CUtensorMap idx_map, const __grid_constant__ CUtensorMap in_map,
                  const float* __restrict__ in, const float* __restrict__ wpacked, const float* __restrict__ bias,
                  const unsigned int* __restrict__ tile_masks, const int* __restrict__ win_lo, int n_out, int tiles,
                  float* __restrict__ out, unsigned char* __restrict__ stats, int stat_group) {
    constexpr int KS = 64;
    constexpr int A_COLS = KS;
    constexpr int A_STAGES = (512 - TS_ACC_COLS) / A_COLS;
    constexpr int B_BYTES = dg_b_bytes(COUT);
    constexpr int BBF_BYTES = COUT * KS * 2;
    constexpr uint32_t IDESC_BF = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(COUT >> 3) << 17) |
                                  ((uint32_t)(TILE_M >> 4) << 24);
    static_assert(A_STAGES == DG_SETS, "one gather set per TMEM A slot");
    constexpr int NBAR = 2 * A_STAGES + 2 * WIN_B_RING + 2 * DG_IDX_STAGES + 2 * WIN_SLOTS + 1;

    extern __shared__ unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long bars[NBAR];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int tile_mask_s[MAX_TILES];
    __shared__ unsigned short a_sched[WIN_GROUPS * 3 * MAX_TILES];
    __shared__ unsigned short a_slice[WIN_GROUPS * 3 * MAX_TILES];
    __shared__ unsigned short b_sched[WIN_GROUPS * 3];
    __shared__ unsigned short w_sched[WIN_GROUPS * MAX_TILES];
    __shared__ int n_a_s, n_b_s;
    __shared__ float wpart_s[16 * 2 * COUT];
    __shared__ int stat_last_s;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int row0 = blockIdx.x * tiles * TILE_M;
    const int rows_cta = min(tiles * TILE_M, n_out - row0);
    if (stats != nullptr)
        for (int i = tid; i < 16 * 2 * COUT; i += WIN_THREADS) wpart_s[i] = 0.f;

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    unsigned char* const smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t smem_b = smem_base;
    const uint32_t smem_idx = smem_b + WIN_B_RING * B_BYTES;
    const int* const idx_ring = reinterpret_cast<const int*>(smem_al + WIN_B_RING * B_BYTES);
    constexpr int W_OFF = WIN_B_RING * B_BYTES + DG_IDX_STAGES * TILE_M * 4;        // multiple of 1024
    const uint32_t smem_w = smem_base + W_OFF;
    const unsigned char* const win_ring = smem_al + W_OFF;
    const uint32_t bar_a_full = smem_u32(&bars[0]);
    const uint32_t bar_a_empty = bar_a_full + 8 * A_STAGES;
    const uint32_t bar_b_full = bar_a_empty + 8 * A_STAGES;
    const uint32_t bar_b_empty = bar_b_full + 8 * WIN_B_RING;
    const uint32_t bar_idx_full = bar_b_empty + 8 * WIN_B_RING;
    const uint32_t bar_idx_empty = bar_idx_full + 8 * DG_IDX_STAGES;
    const uint32_t bar_w_full = bar_idx_empty + 8 * DG_IDX_STAGES;
    const uint32_t bar_w_empty = bar_w_full + 8 * WIN_SLOTS;
    const uint32_t bar_acc = bar_w_empty + 8 * WIN_SLOTS;

    if (tid == 0) {
        for (int s = 0; s < A_STAGES; ++s) {
            mbar_init(bar_a_full + 8 * s, TS_GROUP / 32);
            mbar_init(bar_a_empty + 8 * s, 1);
        }
        for (int s = 0; s < WIN_B_RING; ++s) {
            mbar_init(bar_b_full + 8 * s, 1);
            mbar_init(bar_b_empty + 8 * s, 1);
        }
        for (int s = 0; s < DG_IDX_STAGES; ++s) {
            mbar_init(bar_idx_full + 8 * s, 1);
            mbar_init(bar_idx_empty + 8 * s, TS_GROUP / 32);
        }
        for (int s = 0; s < WIN_SLOTS; ++s) {
            mbar_init(bar_w_full + 8 * s, 1);
            mbar_init(bar_w_empty + 8 * s, 3 * (TS_GROUP / 32));      // 3 stages x 4 warps read every window
        }
        mbar_init(bar_acc, 1);
        fence_barrier_init();
    }
    if (tid < MAX_TILES) {
        const int tile_g = blockIdx.x * tiles + tid;
        tile_mask_s[tid] = (tid < tiles && tile_g * TILE_M < n_out) ? __ldg(tile_masks + tile_g) : 0u;
    }
    if (warp == 16) tmem_alloc(smem_u32(&tmem_base_s), 512);
    __syncthreads();
    if (warp == 0) build_schedule_win(tile_mask_s, tiles, lane, a_sched, a_slice, b_sched, w_sched, &n_a_s, &n_b_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int n_a = n_a_s;
    const int n_b = n_b_s;
    const uint32_t tmem_acc = tmem_base_s;
    const uint32_t tmem_a = tmem_acc + TS_ACC_COLS;

    if (warp < 16) {
        // ============ gather sets: window in shared memory (or global) -> registers -> bf16 hi/lo -> TMEM ============
        const int set = warp >> 2;
        const int q = warp & 3;
        const int rsub = lane >> 2;
        const int l4 = lane & 3;                         // this lane reads channels [8 l4, 8 l4 + 8) of each 32-channel half
        for (int it = set; it < n_a; it += DG_SETS) {
            const int islot = it % DG_IDX_STAGES;
            mbar_wait(bar_idx_full + 8 * islot, ((uint32_t)(it / DG_IDX_STAGES)) & 1u);
            int idx[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) idx[h][rr] = idx_ring[islot * TILE_M + q * 32 + h * 16 + rr * 8 + rsub];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_idx_empty + 8 * islot);
            const int u = it / 3;
            const unsigned we = w_sched[u];
            const int lo = __ldg(win_lo + (size_t)(blockIdx.x * tiles + (int)(we >> 4)) * WIN_GROUPS + (we & 15u));
            const int wslot = u % WIN_SLOTS;
            mbar_wait(bar_w_full + 8 * wslot, ((uint32_t)(u / WIN_SLOTS)) & 1u);
            const unsigned char* const wbase = win_ring + wslot * WIN_BYTES;
            uint4 x[2][2][4];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = idx[h][rr];
                    const unsigned local = (unsigned)(r - lo);
#pragma unroll
                    for (int j2 = 0; j2 < 2; ++j2) {
                        x[h][rr][2 * j2] = make_uint4(0u, 0u, 0u, 0u);
                        x[h][rr][2 * j2 + 1] = make_uint4(0u, 0u, 0u, 0u);
                        if (r >= 0) {
                            if (local < (unsigned)WIN_ROWS) {
                                const unsigned char* rowp = wbase + j2 * WIN_HALF_BYTES + local * 128u;
                                const unsigned sw = local & 7u;
                                x[h][rr][2 * j2] = *reinterpret_cast<const uint4*>(rowp + (((2u * l4) ^ sw) << 4));
                                x[h][rr][2 * j2 + 1] = *reinterpret_cast<const uint4*>(rowp + (((2u * l4 + 1u) ^ sw) << 4));
                            } else {
                                ldg_stream256(in + (size_t)r * WIN_CIN + 32 * j2 + 8 * l4, x[h][rr][2 * j2],
                                              x[h][rr][2 * j2 + 1]);
                            }
                        }
                    }
                }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_w_empty + 8 * wslot);
            const int aslot = it % A_STAGES;
            mbar_wait(bar_a_empty + 8 * aslot, (((uint32_t)(it / A_STAGES)) & 1u) ^ 1u);
            tc_fence_after();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t hi[16], lo_[16];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const uint4 v = x[h][rr][j];
                        const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y);
                        const float f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
                        const __nv_bfloat162 h01 = __floats2bfloat162_rn(f0, f1), h23 = __floats2bfloat162_rn(f2, f3);
                        hi[4 * j + 2 * rr + 0] = *reinterpret_cast<const uint32_t*>(&h01);
                        hi[4 * j + 2 * rr + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                        lo_[4 * j + 2 * rr + 0] = pack_bf16x2(f0 - __low2float(h01), f1 - __high2float(h01));
                        lo_[4 * j + 2 * rr + 1] = pack_bf16x2(f2 - __low2float(h23), f3 - __high2float(h23));
                    }
                const uint32_t taddr = tmem_a + (uint32_t)(aslot * A_COLS) + ((uint32_t)(q * 32 + h * 16) << 16);
                tmem_st_16x256b_x4(taddr, hi);
                tmem_st_16x256b_x4(taddr + KS / 2, lo_);
            }
            tmem_st_wait();
            dg_set_arrive(bar_a_full + 8 * aslot, lane);
        }
    } else if (warp == 16) {
        // ============ MMA issuer ============
        unsigned int started = 0u;
        for (int it = 0; it < n_a; ++it) {
            const unsigned e = a_sched[it];
            const int ib = a_slice[it];
            const int bslot = ib % WIN_B_RING;
            // (re-)waiting on a completed phase is immediate: the slot is not re-armed before its release below
            mbar_wait(bar_b_full + 8 * bslot, ((uint32_t)(ib / WIN_B_RING)) & 1u);
            const uint32_t b_base = smem_b + bslot * B_BYTES;
            const uint64_t db_hi = umma_desc_sw128(b_base);
            const uint64_t db_lo = umma_desc_sw128(b_base + BBF_BYTES);
            const int aslot = it % A_STAGES;
            const int tile = (e >> 8) & 7;
            mbar_wait(bar_a_full + 8 * aslot, ((uint32_t)(it / A_STAGES)) & 1u);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint32_t a_stage = tmem_a + (uint32_t)(aslot * A_COLS);
                const uint32_t d = tmem_acc + (uint32_t)(tile * COUT);
                const uint32_t first = (started >> tile) & 1u;
#pragma unroll
                for (int k16 = 0; k16 < KS / 16; ++k16) {
                    const uint32_t a_hi = a_stage + (uint32_t)(k16 * 8);
                    const uint32_t a_lo = a_hi + KS / 2;
                    const uint64_t adv = (uint64_t)(k16 * 2);
                    umma_bf16_ts(d, a_lo, db_hi + adv, IDESC_BF, (first || k16 > 0) ? 1u : 0u);
                    umma_bf16_ts(d, a_hi, db_lo + adv, IDESC_BF, 1u);
                    umma_bf16_ts(d, a_hi, db_hi + adv, IDESC_BF, 1u);
                }
                umma_commit(bar_a_empty + 8 * aslot);
                if ((e >> 11) & 1u) umma_commit(bar_b_empty + 8 * bslot);
            }
            started |= 1u << tile;
            __syncwarp();
        }
        if (elect_one_sync()) umma_commit(bar_acc);
        __syncwarp();
    } else {
        if (warp == 17 && lane == 0) {
            // ---- weight slices (bulk copies of the pre-swizzled image) ----
            for (int ib = 0; ib < n_b; ++ib) {
                const int bslot = ib % WIN_B_RING;
                mbar_wait(bar_b_empty + 8 * bslot, (((uint32_t)(ib / WIN_B_RING)) & 1u) ^ 1u);
                const float* src = wpacked + (size_t)b_sched[ib] * (B_BYTES / 4);
                mbar_arrive_expect_tx(bar_b_full + 8 * bslot, B_BYTES);
                bulk_g2s(smem_b + bslot * B_BYTES, src, B_BYTES, bar_b_full + 8 * bslot);
            }
        } else if (warp == 18 && lane == 0) {
            // ---- row-index segments: one [1 tap x 128 rows] TMA box per stage ----
            tma_prefetch_desc(&idx_map);
            for (int it = 0; it < n_a; ++it) {
                const int islot = it % DG_IDX_STAGES;
                mbar_wait(bar_idx_empty + 8 * islot, (((uint32_t)(it / DG_IDX_STAGES)) & 1u) ^ 1u);
                const unsigned e = a_sched[it];
                mbar_arrive_expect_tx(bar_idx_full + 8 * islot, TILE_M * 4);
                tma_load_2d(smem_idx + islot * TILE_M * 4, &idx_map, row0 + (int)((e >> 8) & 7u) * TILE_M, (int)(e & 31u),
                            bar_idx_full + 8 * islot);
            }
        } else if (warp == 19 && lane == 0) {
            // ---- windows: rows [lo, lo + WIN_ROWS) of the feature matrix, two 32-channel halves, 128-byte swizzle ----
            tma_prefetch_desc(&in_map);
            const int n_u = n_a / 3;
            for (int u = 0; u < n_u; ++u) {
                const int wslot = u % WIN_SLOTS;
                mbar_wait(bar_w_empty + 8 * wslot, (((uint32_t)(u / WIN_SLOTS)) & 1u) ^ 1u);
                const unsigned we = w_sched[u];
                int lo = __ldg(win_lo + (size_t)(blockIdx.x * tiles + (int)(we >> 4)) * WIN_GROUPS + (we & 15u));
                if (lo == 0x7fffffff) lo = 0;                 // no valid row in this (tile, group): nothing will be read
                mbar_arrive_expect_tx(bar_w_full + 8 * wslot, WIN_BYTES);
                tma_load_2d(smem_w + wslot * WIN_BYTES, &in_map, 0, lo, bar_w_full + 8 * wslot);
                tma_load_2d(smem_w + wslot * WIN_BYTES + WIN_HALF_BYTES, &in_map, 32, lo, bar_w_full + 8 * wslot);
            }
        }
        __syncwarp();
    }

    static_assert((size_t)WIN_SLOTS * WIN_BYTES >= DG_EP_STAGE_BYTES, "the epilogue stages its tiles in the (idle) window ring");
    dg_epilogue<COUT, WIN_THREADS>(tid, n_a, bar_acc, tiles, rows_cta, row0, tile_mask_s, tmem_acc, bias, out, stats, wpart_s,
                                   stat_group, &stat_last_s, reinterpret_cast<float*>(smem_al + W_OFF));
    if (warp == 16) tmem_dealloc(tmem_acc, 512);
}

// win_lo[tile * 9 + g] = smallest valid row index among the three taps 3g .. 3g+2 of the tile's 128 rows (INT_MAX: none)
__global__ void __launch_bounds__(288)
k_window_lo(const int* __restrict__ nbr_t, int ld, int* __restrict__ win_lo) {
    const int tile = blockIdx.x, g = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int m = 0x7fffffff;
    for (int x = 0; x < 3; ++x) {
        const int* seg = nbr_t + (size_t)(3 * g + x) * ld + (size_t)tile * TILE_M;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int v = __ldg(seg + lane + 32 * j);
            if (v >= 0) m = min(m, v);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) win_lo[(size_t)tile * WIN_GROUPS + g] = m;
}

// tiles per CTA and grid size for n_out rows (one wave-balanced choice shared by the launcher and the size queries);
// nsets 4: one CTA per SM with up to 256 / cout tiles, nsets 2: two CTAs per SM with up to 128 / cout tiles each
void dg_grid(int n_out, int cout, int* tiles_out, int* ctas_out, int nsets = 4) {
    const int n_tiles = (n_out + TILE_M - 1) / TILE_M;
    const int tmax = dg_tiles_max(nsets, cout);
    const int slots = UD_NUM_SMS * (nsets == 4 ? 1 : 2);
    const int waves = (n_tiles + slots * tmax - 1) / (slots * tmax);
    int tiles = (n_tiles + slots * waves - 1) / (slots * (waves > 0 ? waves : 1));
    if (tiles < 1) tiles = 1;
    if (tiles > tmax) tiles = tmax;
    *tiles_out = tiles;
    *ctas_out = (n_tiles + tiles - 1) / tiles;
}

// Gather sets per CTA for a layer shape: 2 (two CTAs per SM) where it measured faster on B200 (profiles/r03_pipeline_probes.md).
// UD_DG_NSETS=2|4 overrides the table (diagnostic).
int dg_nsets(int cin, int cout) {
    static const int forced = [] {
        const char* e = getenv("UD_DG_NSETS");
        return e ? atoi(e) : 0;
    }();
    if (forced == 2 || forced == 4) return forced;
    (void)cin;
    (void)cout;
    return 2;       // B200, bs-4 distillation step: 64->64 0.527 -> 0.432 ms, 128->128 0.405 -> 0.369, 32->32 0.301 -> 0.253, 16->16 0.100 -> 0.088
}

size_t dg_stat_bytes(int n_out, int cout) {           // the larger of the two configurations (the query does not know cin)
    size_t best = 0;
    for (int nsets = 2; nsets <= 4; nsets += 2) {
        int tiles, ctas;
        dg_grid(n_out, cout, &tiles, &ctas, nsets);
        const size_t b = (size_t)DG_STAT_FINAL_BYTES + ud_det_reduce_bytes(ctas, 2 * cout);
        best = b > best ? b : best;
    }
    return best;
}

template <int CIN, int COUT, bool SPLIT, int NSETS>
int launch_umma_dg_n(const float* in, const float* wpacked, const float* bias, const int* nbr_t,
                     const unsigned int* tile_masks, int n_out, int K, float* out, unsigned char* stats,
                     cudaStream_t stream, const DgPost post) {
    const size_t smem = dg_smem_bytes(CIN, COUT, NSETS);
    auto kern = k_spconv_umma_dg<CIN, COUT, SPLIT, NSETS>;
    UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (NSETS == 2)         // two CTAs per SM need the large shared-memory carve-out
        UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int tiles, ctas;
    dg_grid(n_out, COUT, &tiles, &ctas, NSETS);
    if (stats != nullptr)       // arrival counters of the statistics reduction
        UD_CUDA_TRY(cudaMemsetAsync(stats + DG_STAT_FINAL_BYTES, 0, UD_DET_COUNTER_BYTES, stream));
    CUtensorMap idx_map;        // tap-major table as a 2-D tensor: (K rounded up to 4) rows x (n_out rounded up to 128) int32
    const uint64_t ld = (uint64_t)((n_out + TILE_M - 1) / TILE_M) * TILE_M;
    int rc = ud_make_tmap_2d_b32(&idx_map, nbr_t, CU_TENSOR_MAP_DATA_TYPE_INT32, (uint64_t)((K + 3) / 4 * 4), ld, ld * 4,
                                 (uint32_t)ts_tapg(CIN), TILE_M);
    if (rc != UD_OK) return rc;
    kern<<<ctas, dg_threads(NSETS), smem, stream>>>(idx_map, in, wpacked, bias, nbr_t, tile_masks, n_out, K, tiles, out, stats,
                                                    ud_det_group(ctas), post);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

template <int CIN, int COUT, bool SPLIT>
int launch_umma_dg(const float* in, const float* wpacked, const float* bias, const int* nbr_t,
                   const unsigned int* tile_masks, int n_out, int K, float* out, unsigned char* stats,
                   cudaStream_t stream, const DgPost post = DgPost{nullptr, nullptr, nullptr, 0}) {
    if (dg_nsets(CIN, COUT) == 2)
        return launch_umma_dg_n<CIN, COUT, SPLIT, 2>(in, wpacked, bias, nbr_t, tile_masks, n_out, K, out, stats, stream, post);
    return launch_umma_dg_n<CIN, COUT, SPLIT, 4>(in, wpacked, bias, nbr_t, tile_masks, n_out, K, out, stats, stream, post);
}

template <int COUT>
int launch_umma_win(const float* in, int n_in, const float* wpacked, const float* bias, const int* nbr_t,
                    const unsigned int* tile_masks, const int* win_lo, int n_out, float* out, unsigned char* stats,
                    cudaStream_t stream) {
    const size_t smem = win_smem_bytes(COUT);
    auto kern = k_spconv_umma_win<COUT>;
    UD_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int tiles, ctas;
    dg_grid(n_out, COUT, &tiles, &ctas);
    if (stats != nullptr)
        UD_CUDA_TRY(cudaMemsetAsync(stats + DG_STAT_FINAL_BYTES, 0, UD_DET_COUNTER_BYTES, stream));
    CUtensorMap idx_map, in_map;
    const uint64_t ld = (uint64_t)((n_out + TILE_M - 1) / TILE_M) * TILE_M;
    int rc = ud_make_tmap_2d_b32(&idx_map, nbr_t, CU_TENSOR_MAP_DATA_TYPE_INT32, 28, ld, ld * 4, 1, TILE_M);
    if (rc != UD_OK) return rc;
    rc = ud_make_tmap_2d_b32(&in_map, in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)n_in, WIN_CIN, WIN_CIN * 4, WIN_ROWS, 32,
                             true);
    if (rc != UD_OK) return rc;
    kern<<<ctas, WIN_THREADS, smem, stream>>>(idx_map, in_map, in, wpacked, bias, tile_masks, win_lo, n_out, tiles, out,
                                              stats, ud_det_group(ctas));
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// forward shapes of VoxelResBackBone8x and the dgrad (cin/cout swapped) of every layer but the first
bool tc_shape_ok(int K, int cin, int cout) {
    static const int shapes[][2] = {{16, 16}, {16, 32}, {32, 16}, {32, 32}, {32, 64}, {64, 32}, {64, 64}, {64, 128},
                                    {128, 64}, {128, 128}};
    if (K < 1 || K > MAX_TAPS) return false;
    for (const auto& s : shapes)
        if (s[0] == cin && s[1] == cout) return true;
    return false;
}

}  // namespace

#define UD_DISPATCH_DG(CI, CO)                                                                                               \
    if (cin == CI && cout == CO)                                                                                            \
        return in_is_split ? launch_umma_dg<CI, CO, true>(in, wpacked, bias, nbr_t, tile_masks, n_out, K, out, stats, stream)  \
                           : launch_umma_dg<CI, CO, false>(in, wpacked, bias, nbr_t, tile_masks, n_out, K, out, stats, stream);
#define UD_DISPATCH_DG_POST(CI, CO)                                                                                          \
    if (cin == CI && cout == CO)                                                                                            \
        return launch_umma_dg<CI, CO, false>(in, wpacked, bias, nbr_t, tile_masks, n_out, K, out, nullptr, stream, post);

extern "C" {

// 1 when (K, cin, cout) has a tensor-core instantiation, else 0.
int ud_sparse_conv_tc_supported(int K, int cin, int cout) { return tc_shape_ok(K, cin, cout) ? 1 : 0; }

int ud_sparse_conv_tc_weight_bytes(int K, int cin, int cout, size_t* bytes) {
    if (!bytes) return UD_ERR_INVALID_ARG;
    if (!tc_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    const int tapg = ts_tapg(cin);
    *bytes = (size_t)((K + tapg - 1) / tapg) * (cin * tapg) * cout * 2 * sizeof(__nv_bfloat16);
    return UD_OK;
}

// w: (K, cin, cout) f32 (the layout ud_sparse_conv_forward takes) -> packed bf16 hi/lo image, 128-byte swizzle applied
// (16-byte aligned): one bulk copy per (super-tap, 64-channel stage) in the kernel.
int ud_sparse_conv_tc_pack_weights(const float* w, int K, int cin, int cout, void* packed, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!w || !packed) return UD_ERR_INVALID_ARG;
    if (!tc_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if (((uintptr_t)packed & 15u) != 0) return UD_ERR_INVALID_ARG;
    const int tapg = ts_tapg(cin);
    const long long total = (long long)((K + tapg - 1) / tapg) * cin * tapg * cout;
    k_pack_weights_bf3<<<ud_blocks(total, 256), 256, 0, stream>>>(w, K, cin, cout, tapg, -1, (unsigned char*)packed);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// The same packed image straight from the spconv-2 parameter layout (Cout_p, kz, ky, kx, Cin_p) -- no (K, cin, cout) copy, no
// flip / transpose passes in front of the data-gradient convolution.  cin / cout are those of the convolution that will RUN:
//   mode 0: the forward convolution (cin = Cin_p, cout = Cout_p);
//   mode 1: the data gradient of a submanifold convolution = the same kernel on the transposed weights with mirrored taps
//           (cin = Cout_p, cout = Cin_p);   mode 2: the data gradient of a strided convolution (transposed, not mirrored).
int ud_sparse_conv_tc_pack_weights_param(const float* w_param, int K, int cin, int cout, int mode, void* packed, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!w_param || !packed || mode < 0 || mode > 2) return UD_ERR_INVALID_ARG;
    if (!tc_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if (((uintptr_t)packed & 15u) != 0) return UD_ERR_INVALID_ARG;
    const int tapg = ts_tapg(cin);
    const long long total = (long long)((K + tapg - 1) / tapg) * cin * tapg * cout;
    k_pack_weights_bf3<<<ud_blocks(total, 256), 256, 0, stream>>>(w_param, K, cin, cout, tapg, mode, (unsigned char*)packed);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// nbr (n, K) int32 row-major gather table -> nbr_t (K4, ld) tap-major, K4 = K rounded up to 4 and ld = n rounded up to 128
// (all padding = -1, base 16-byte aligned so every (tile, tap) index segment can be bulk-copied) + tile_masks
// (ceil(n/128)) uint32 (bit k set when tap k occurs in the 128-row tile): the two inputs ud_sparse_conv_tc_forward reads.
int ud_rulebook_transpose(const int* nbr, int n, int K, int* nbr_t, unsigned int* tile_masks, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || K < 1 || K > MAX_TAPS || (n > 0 && (!nbr || !nbr_t || !tile_masks))) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    UD_CUDA_TRY(cudaMemsetAsync(tile_masks, 0, sizeof(unsigned int) * (size_t)((n + TILE_M - 1) / TILE_M), stream));
    const int ld = (n + TILE_M - 1) / TILE_M * TILE_M;
    k_transpose_table<<<ud_blocks(ld, 32), 256, 0, stream>>>(nbr, n, K, ld, nbr_t, tile_masks);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Window-staged kernel (k_spconv_umma_win): 1 when a (K, cin, cout) has an instantiation.
int ud_sparse_conv_tc_window_supported(int K, int cin, int cout) { return (K == 27 && cin == WIN_CIN && cout == 64) ? 1 : 0; }

// win_lo (ceil(n/128), 9) int32 <- per 128-row tile and (kz, ky) tap group, the smallest valid row index of the tap-major
// table nbr_t (28, ld) of a 3x3x3 kernel (INT_MAX when the group is absent): where the window of the group starts.
int ud_rulebook_windows(const int* nbr_t, int n, int K, int* win_lo, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || K != 27 || (n > 0 && (!nbr_t || !win_lo))) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    const int n_tiles = (n + TILE_M - 1) / TILE_M;
    k_window_lo<<<n_tiles, 288, 0, stream>>>(nbr_t, n_tiles * TILE_M, win_lo);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Same contract as ud_sparse_conv_tc_forward for the shapes ud_sparse_conv_tc_window_supported() accepts, plus n_in (rows
// of `in`: bounds of the TMA tensor map) and win_lo from ud_rulebook_windows.  Meant for tables whose rows are in raster
// order (any order is correct: rows outside a window are gathered from global memory).
int ud_sparse_conv_tc_forward_window(const float* in, int n_in, const void* wpacked_, const float* bias, const int* nbr_t,
                                     const unsigned int* tile_masks, const int* win_lo, int n_out, int K, int cin,
                                     int cout, float* out, void* bn_stats, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const float* wpacked = (const float*)wpacked_;
    unsigned char* stats = (unsigned char*)bn_stats;
    if (!in || !wpacked || !nbr_t || !tile_masks || !win_lo || !out || n_out < 0 || n_in < 1) return UD_ERR_INVALID_ARG;
    if (!ud_sparse_conv_tc_window_supported(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if ((((uintptr_t)in | (uintptr_t)wpacked | (uintptr_t)out | (uintptr_t)stats) & 15u) != 0) return UD_ERR_INVALID_ARG;
    if (n_out == 0) {
        if (stats) UD_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * cout * sizeof(double), stream));
        return UD_OK;
    }
    return launch_umma_win<64>(in, n_in, wpacked, bias, nbr_t, tile_masks, win_lo, n_out, out, stats, stream);
}

// x (n, C) f32 -> out (n, C) 4-byte words in the split layout (see k_split_bf16): the form of `in` that
// ud_sparse_conv_tc_forward takes with in_is_split = 1.  C % 4 == 0; 16-byte aligned.
int ud_split_bf16(const float* x, int n, int C, void* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || C < 4 || C % 4 != 0 || (n > 0 && (!x || !out))) return UD_ERR_INVALID_ARG;
    if ((((uintptr_t)x | (uintptr_t)out) & 15u) != 0) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    const long long n4 = (long long)n * C / 4;
    const long long want = (n4 + 255) / 256;
    const int grid = (int)(want < UD_NUM_SMS * 16 ? want : UD_NUM_SMS * 16);
    k_split_bf16<<<grid, 256, 0, stream>>>((const float4*)x, n4, (uint4*)out);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// Size of the statistics buffer ud_sparse_conv_tc_forward needs for n_out rows (0 rows -> header only).
int ud_sparse_conv_tc_stats_bytes(int n_out, int cout, size_t* bytes) {
    if (!bytes || n_out < 0 || cout < 1 || cout > 256) return UD_ERR_INVALID_ARG;
    *bytes = dg_stat_bytes(n_out > 0 ? n_out : 1, cout);
    return UD_OK;
}

// in (n_in, cin) f32, 16-byte aligned rows; wpacked from ud_sparse_conv_tc_pack_weights; bias (cout)|NULL;
// nbr_t (K, n_out) int32 tap-major gather table and tile_masks (ceil(n_out/128)) from ud_rulebook_transpose;
// out (n_out, cout) f32.
// bn_stats (nullable): buffer of ud_sparse_conv_tc_stats_bytes(n_out, cout) bytes, 16-byte aligned, no initialisation
// needed.  On return its first 2 * cout doubles hold the per-channel sum and sum of squares of `out` -- the statistics
// pass of the BatchNorm that follows (spconv_backbone.py:50-56,97-113), fused into the epilogue and reduced in a FIXED
// order (bit-reproducible from run to run); the rest of the buffer is scratch of that reduction.
int ud_sparse_conv_tc_forward(const float* in, int in_is_split, const void* wpacked_, const float* bias, const int* nbr_t,
                              const unsigned int* tile_masks, int n_out, int K, int cin, int cout, float* out,
                              void* bn_stats, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const float* wpacked = (const float*)wpacked_;
    unsigned char* stats = (unsigned char*)bn_stats;
    if (!in || !wpacked || !nbr_t || !tile_masks || !out || n_out < 0) return UD_ERR_INVALID_ARG;
    if (!tc_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if ((((uintptr_t)in | (uintptr_t)wpacked | (uintptr_t)out | (uintptr_t)stats) & 15u) != 0) return UD_ERR_INVALID_ARG;
    if (n_out == 0) {
        if (stats) UD_CUDA_TRY(cudaMemsetAsync(stats, 0, 2 * cout * sizeof(double), stream));
        return UD_OK;
    }
    UD_DISPATCH_DG(16, 16)
    UD_DISPATCH_DG(16, 32)
    UD_DISPATCH_DG(32, 16)
    UD_DISPATCH_DG(32, 32)
    UD_DISPATCH_DG(32, 64)
    UD_DISPATCH_DG(64, 32)
    UD_DISPATCH_DG(64, 64)
    UD_DISPATCH_DG(64, 128)
    UD_DISPATCH_DG(128, 64)
    UD_DISPATCH_DG(128, 128)
    return UD_ERR_UNSUPPORTED;
}

// Inference form: out = relu?((conv(in) + bias) * scale[c] + shift[c] (+ residual)) -- the eval-mode BatchNorm1d (+ residual)
// + ReLU that follows every convolution of the backbone, folded into the convolution's epilogue.
int ud_sparse_conv_tc_forward_affine(const float* in, const void* wpacked_, const float* bias, const float* scale, const float* shift,
                                     const float* residual, int relu, const int* nbr_t, const unsigned int* tile_masks,
                                     int n_out, int K, int cin, int cout, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const float* wpacked = (const float*)wpacked_;
    if (!in || !wpacked || !scale || !shift || !nbr_t || !tile_masks || !out || n_out < 0) return UD_ERR_INVALID_ARG;
    if (!tc_shape_ok(K, cin, cout)) return UD_ERR_UNSUPPORTED;
    if ((((uintptr_t)in | (uintptr_t)wpacked | (uintptr_t)out | (uintptr_t)scale | (uintptr_t)shift | (uintptr_t)residual) & 15u) != 0)
        return UD_ERR_INVALID_ARG;
    if (n_out == 0) return UD_OK;
    const DgPost post{scale, shift, residual, relu};
    UD_DISPATCH_DG_POST(16, 16)
    UD_DISPATCH_DG_POST(16, 32)
    UD_DISPATCH_DG_POST(32, 16)
    UD_DISPATCH_DG_POST(32, 32)
    UD_DISPATCH_DG_POST(32, 64)
    UD_DISPATCH_DG_POST(64, 32)
    UD_DISPATCH_DG_POST(64, 64)
    UD_DISPATCH_DG_POST(64, 128)
    UD_DISPATCH_DG_POST(128, 64)
    UD_DISPATCH_DG_POST(128, 128)
    return UD_ERR_UNSUPPORTED;
}

}  // extern "C"

#ifdef UD_DG_TRACE
extern "C" int ud_debug_dg_trace(long long* host_out) {
    UD_CUDA_TRY(cudaMemcpyFromSymbol(host_out, ud_dg_trace, sizeof(long long) * 8 * DG_TRACE_STAGES));
    return UD_OK;
}
#endif
#ifdef UD_DG_PROBE
extern "C" int ud_debug_dg_probe(unsigned long long* host_out16, int reset) {
    if (host_out16 != nullptr) UD_CUDA_TRY(cudaMemcpyFromSymbol(host_out16, ud_dg_probe_acc, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long zero[16] = {};
        UD_CUDA_TRY(cudaMemcpyToSymbol(ud_dg_probe_acc, zero, sizeof(zero)));
    }
    return UD_OK;
}
#endif
