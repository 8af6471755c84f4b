// FCOS-style target assignment of the CenterPoint head, one kernel for the whole batch and all tasks.
//
// Replaces  reference unidistill/layers/head/det3d/target_assigner/fcos_assigner.py:73-285  (B x n_tasks Python
// iterations with host syncs and ~40 tiny kernels each, run for the student AND the teacher every step) and the
// ~200-kernel tensor program of this repo's torch mirror (same file name under unidistill/), whose results it
// reproduces bit for bit:
//   * positives of a task = union over the task's GT boxes of their `topk` nearest anchor points (squared distance
//     in feature-map units, computed as (ax-px)^2 + (ay-py)^2 in fp32 without FMA contraction, like torch);
//   * each positive anchor regresses to its nearest GT of the task (first one in task order on ties);
//   * outputs listed in ascending anchor order, truncated / zero padded to max_objs; the heat-map is one-hot at ALL
//     positives;  box code = [(x-ax)/osf, (y-ay)/osf, z, log(dx*vs0), log(dy*vs1), log(dz), sin(yaw), cos(yaw), rest...].
// Anchors form a regular fw x fh grid (anchor a = i*fw + j); the `topk` (<= 9) nearest anchors of a point lie in the
// (2R+1)^2 window around the (clamped) nearest grid node: inside the map they are within 2 cells; for a point off a map
// corner they are up to 9 consecutive anchors of one border row, i.e. at most 8 columns from the corner.  With R = 8 the
// window is exact for any point, with R = 4 up to ~7 cells outside the map (tests/test_fcos_window.py checks both on the
// CPU).  (2R+1)^2 candidates per GT are ranked instead of all fw*fh.
// grid (n_tasks, B), 256 threads; one warp per GT for the top-k, bitmap + popcount scan for the ordered compaction.
#include "common.cuh"

namespace {

constexpr int FA_THREADS = 256;
constexpr int FA_MAX_GT = 128;          // GT rows per sample
constexpr int FA_MAX_WORDS = 1024;      // anchors / 32  (<= 32768 anchors)
constexpr int FA_MAX_TASKS = 16;
// Radius 8 is exact for ANY centre position when topk <= 9 (proof + lane-level emulation in tests/test_fcos_window.py;
// tests/test_gpu_assigner.py draws GT centres up to 12 cells outside the map).  Radius 4, the round-1 build, was exact
// only up to ~7 cells outside.
constexpr int FA_RADIUS = 8;
constexpr int FA_MAX_TOPK = FA_RADIUS + 1;
constexpr int FA_WIN = 2 * FA_RADIUS + 1;
constexpr int FA_SLOTS = (FA_WIN * FA_WIN + 31) / 32;   // window candidates per lane

struct FaParams {
    int cls_prefix[FA_MAX_TASKS + 1];   // classes before task t
};

__device__ __forceinline__ float fa_d2(float ax, float ay, float px, float py) {
    const float dx = __fsub_rn(ax, px), dy = __fsub_rn(ay, py);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__global__ void __launch_bounds__(FA_THREADS)
k_fcos_assign(const float* __restrict__ boxes, const int* __restrict__ task_id, const int* __restrict__ cls_off,
              const float* __restrict__ anchors, int B, int M, int NB, int fw, int fh, float stride, FaParams prm,
              int topk, int max_objs, float vs0, float vs1, float osf, float* __restrict__ heat,
              long long* __restrict__ ind, unsigned char* __restrict__ mask, long long* __restrict__ cat,
              float* __restrict__ enc, int* __restrict__ n_pos_out) {
    __shared__ float gx[FA_MAX_GT], gy[FA_MAX_GT];
    __shared__ int gidx[FA_MAX_GT];
    __shared__ int n_members_s;
    __shared__ unsigned int bitmap[FA_MAX_WORDS];
    __shared__ int prefix[FA_MAX_WORDS];
    __shared__ int warp_tot[FA_THREADS / 32];

    const int t = blockIdx.x, b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int A = fw * fh;
    const int words = (A + 31) / 32;
    const int ED = NB + 1;
    const int n_cls = prm.cls_prefix[t + 1] - prm.cls_prefix[t];
    const float* bx = boxes + (size_t)b * M * NB;

    // ---- members of this task, in row order ----
    if (tid == 0) {
        int n = 0;
        for (int m = 0; m < M; ++m)
            if (task_id[(size_t)b * M + m] == t) {
                gx[n] = bx[(size_t)m * NB + 0];
                gy[n] = bx[(size_t)m * NB + 1];
                gidx[n] = m;
                ++n;
            }
        n_members_s = n;
    }
    for (int w = tid; w < words; w += FA_THREADS) bitmap[w] = 0u;
    __syncthreads();
    const int n_members = n_members_s;

    // ---- top-k nearest anchors of every member (one warp per GT) ----
    for (int g = warp; g < n_members; g += FA_THREADS / 32) {
        const float px = gx[g], py = gy[g];
        int cj = (int)rintf(px / stride), ci = (int)rintf(py / stride);
        if (!(px == px)) cj = 0;                            // NaN guards
        if (!(py == py)) ci = 0;
        cj = min(max(cj, 0), fw - 1);
        ci = min(max(ci, 0), fh - 1);
        float cd[FA_SLOTS];
        int ca[FA_SLOTS];
#pragma unroll
        for (int s = 0; s < FA_SLOTS; ++s) {
            const int c = lane + 32 * s;
            cd[s] = INFINITY;
            ca[s] = 0x7fffffff;
            if (c < FA_WIN * FA_WIN) {
                const int i = ci + c / FA_WIN - FA_RADIUS, j = cj + c % FA_WIN - FA_RADIUS;
                if (i >= 0 && i < fh && j >= 0 && j < fw) {
                    const int a = i * fw + j;
                    cd[s] = fa_d2(__ldg(anchors + 2 * a), __ldg(anchors + 2 * a + 1), px, py);
                    ca[s] = a;
                    if (!(cd[s] == cd[s])) cd[s] = INFINITY;
                }
            }
        }
        for (int r = 0; r < topk; ++r) {
            // lane-local best, then warp argmin on (distance, anchor index)
            int bs = 0;
#pragma unroll
            for (int s = 1; s < FA_SLOTS; ++s)
                if (cd[s] < cd[bs] || (cd[s] == cd[bs] && ca[s] < ca[bs])) bs = s;
            float d = cd[bs];
            int a = ca[bs];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float od = __shfl_xor_sync(0xffffffffu, d, o);
                const int oa = __shfl_xor_sync(0xffffffffu, a, o);
                if (od < d || (od == d && oa < a)) { d = od; a = oa; }
            }
            if (a == 0x7fffffff) break;                     // fewer candidates than topk
            if (ca[bs] == a) {                              // the owner retires its candidate and marks the anchor
                cd[bs] = INFINITY;
                ca[bs] = 0x7fffffff;
                atomicOr(&bitmap[a >> 5], 1u << (a & 31));
            }
        }
    }
    __syncthreads();

    // ---- exclusive popcount scan over the bitmap words: rank of every positive in ascending anchor order ----
    constexpr int WPT = FA_MAX_WORDS / FA_THREADS;          // words per thread (4)
    int cnt[WPT];
    int local = 0;
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid * WPT + k;
        cnt[k] = w < words ? __popc(bitmap[w]) : 0;
        local += cnt[k];
    }
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    int n_pos = 0;
    for (int w = 0; w < FA_THREADS / 32; ++w) n_pos += warp_tot[w];
    if (n_pos_out != nullptr && tid == 0) n_pos_out[t * B + b] = n_pos;     // > max_objs: the reference would assert
    int run = base + incl - local;
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
        const int w = tid * WPT + k;
        if (w < words) prefix[w] = run;
        run += cnt[k];
    }
    __syncthreads();

    // ---- outputs ----
    const size_t slot0 = ((size_t)t * B + b) * max_objs;
    float* heat_t = heat + (size_t)B * A * prm.cls_prefix[t] + (size_t)b * n_cls * A;     // (B, n_cls, A) block of task t
    for (int w = tid; w < words; w += FA_THREADS) {
        unsigned int bits = bitmap[w];
        int r = prefix[w];
        while (bits) {
            const int a = (w << 5) + __ffs(bits) - 1;
            bits &= bits - 1u;
            const float ax = __ldg(anchors + 2 * a), ay = __ldg(anchors + 2 * a + 1);
            // nearest member (first minimum in task order; NaN distances never win)
            int best = 0;
            float bd = fa_d2(ax, ay, gx[0], gy[0]);
            if (!(bd == bd)) bd = INFINITY;
            for (int g = 1; g < n_members; ++g) {
                const float d = fa_d2(ax, ay, gx[g], gy[g]);
                if (d < bd) { bd = d; best = g; }
            }
            const int m = gidx[best];
            const int cls = cls_off[(size_t)b * M + m];
            if (cls >= 0 && cls < n_cls) heat_t[(size_t)cls * A + a] = 1.0f;
            if (r < max_objs) {
                const float* gb = bx + (size_t)m * NB;
                ind[slot0 + r] = a;
                mask[slot0 + r] = 1;
                cat[slot0 + r] = cls;
                float* e = enc + (slot0 + r) * ED;
                e[0] = __fdiv_rn(__fsub_rn(gb[0], ax), osf);
                e[1] = __fdiv_rn(__fsub_rn(gb[1], ay), osf);
                e[2] = gb[2];
                e[3] = logf(__fmul_rn(gb[3], vs0));
                e[4] = logf(__fmul_rn(gb[4], vs1));
                e[5] = logf(gb[5]);
                e[6] = sinf(gb[6]);
                e[7] = cosf(gb[6]);
                for (int c = 7; c < NB; ++c) e[c + 1] = gb[c];
            }
            ++r;
        }
    }
    for (int r = min(n_pos, max_objs) + tid; r < max_objs; r += FA_THREADS) {
        ind[slot0 + r] = 0;
        mask[slot0 + r] = 0;
        cat[slot0 + r] = 0;
        float* e = enc + (slot0 + r) * ED;
        for (int c = 0; c < ED; ++c) e[c] = 0.f;
    }
}

}  // namespace

extern "C" {

// boxes (B,M,NB) f32 in feature-map units [x,y,z,dx,dy,dz,yaw,(vx,vy)], rows already in the reference's task order;
// task_id (B,M) int32 (-1: row takes no part), cls_off (B,M) int32 (class index inside its task); anchors (fw*fh,2) f32
// (x,y), anchor a = i*fw + j on a regular grid of pitch `stride`; cls_prefix (n_tasks+1) HOST ints (classes before each
// task).  Outputs: heat (concatenation over tasks of (B,n_cls_t,fh*fw) blocks; the caller zeroes it), ind / cat
// (n_tasks,B,max_objs) int64, mask (n_tasks,B,max_objs) u8, enc (n_tasks,B,max_objs,NB+1) f32 -- every slot is written.
// n_pos (n_tasks,B) int32, optional (NULL to skip): positives BEFORE truncation to max_objs (the reference asserts
// n_pos < max_objs, fcos_assigner.py `padding_to_maxobjs`; here the overflow is reported instead of silently truncated).
int ud_fcos_assign(const float* boxes, const int* task_id, const int* cls_off, const float* anchors, int B, int M, int NB,
                   int fw, int fh, float stride, int n_tasks, const int* cls_prefix, int topk, int max_objs, float vs0,
                   float vs1, float osf, float* heat, long long* ind, unsigned char* mask, long long* cat, float* enc,
                   int* n_pos, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || M < 1 || NB < 7 || fw < 1 || fh < 1 || n_tasks < 1 || !cls_prefix || topk < 1 || max_objs < 1)
        return UD_ERR_INVALID_ARG;
    if (B == 0) return UD_OK;
    if (!boxes || !task_id || !cls_off || !anchors || !heat || !ind || !mask || !cat || !enc) return UD_ERR_INVALID_ARG;
    if (M > FA_MAX_GT || n_tasks > FA_MAX_TASKS || topk > FA_MAX_TOPK || (long long)fw * fh > 32LL * FA_MAX_WORDS ||
        B > 65535 || !(stride > 0.f))
        return UD_ERR_UNSUPPORTED;
    FaParams prm;
    for (int t = 0; t <= n_tasks; ++t) prm.cls_prefix[t] = cls_prefix[t];
    dim3 grid(n_tasks, B);
    k_fcos_assign<<<grid, FA_THREADS, 0, stream>>>(boxes, task_id, cls_off, anchors, B, M, NB, fw, fh, stride, prm, topk,
                                                   max_objs, vs0, vs1, osf, heat, ind, mask, cat, enc, n_pos);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
