// Rotated-BEV IoU and greedy NMS for the detection head's proposal layer (inference post-processing).
//
// Replaces the reference's binary-only extension
//   unidistill/layers/head/det3d/generate_proposals/iou3d_nms_cuda*.so  (nms_gpu, boxes_iou_bev_gpu)
// called from generate_proposals/centerpoint_gen_proposals.py:85-106.  Boxes are (n,7) f32 [x,y,z,dx,dy,dz,heading].
// The overlap follows the published OpenPCDet iou3d_nms algorithm (edge-pair intersections, corner-in-box tests with
// a 1e-2 margin, polar-angle sort about the mean point, fan area) so that keep decisions match the reference's;
// oracle/nms_ref.c restates the same arithmetic on the CPU.
//
// What is different operationally: the reference builds the 64x64-block suppression mask on the GPU, copies it to the
// host and sweeps it there (one device->host sync per (task, sample)).  Here every (task, sample) set of a batch goes
// through ONE mask launch (grid.z = set) and ONE sweep launch (one CTA per set, suppression words in shared memory,
// early exit once `max_keep` boxes are kept); counts stay on the device, nothing synchronises.
#include "common.cuh"

namespace {

constexpr float NMS_EPS = 1e-8f;
constexpr float NMS_MARGIN = 1e-2f;
constexpr int NMS_BLOCK = 64;          // boxes per mask word

struct Pt { float x, y; };

// no FMA contraction in the predicates: the CPU oracle is compiled with -ffp-contract=off
__device__ __forceinline__ float cross2(Pt a, Pt b) { return __fsub_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)); }
__device__ __forceinline__ float cross3(Pt p1, Pt p2, Pt p0) {
    return __fsub_rn(__fmul_rn(__fsub_rn(p1.x, p0.x), __fsub_rn(p2.y, p0.y)),
                     __fmul_rn(__fsub_rn(p2.x, p0.x), __fsub_rn(p1.y, p0.y)));
}
__device__ __forceinline__ bool rect_cross(Pt p1, Pt p2, Pt q1, Pt q2) {
    return fminf(p1.x, p2.x) <= fmaxf(q1.x, q2.x) && fminf(q1.x, q2.x) <= fmaxf(p1.x, p2.x) &&
           fminf(p1.y, p2.y) <= fmaxf(q1.y, q2.y) && fminf(q1.y, q2.y) <= fmaxf(p1.y, p2.y);
}
__device__ __forceinline__ bool in_box(const float* box, Pt p) {
    float s, c;
    sincosf(-box[6], &s, &c);
    const float ox = __fsub_rn(p.x, box[0]), oy = __fsub_rn(p.y, box[1]);
    const float rx = __fadd_rn(__fmul_rn(ox, c), __fmul_rn(oy, -s));
    const float ry = __fadd_rn(__fmul_rn(ox, s), __fmul_rn(oy, c));
    return fabsf(rx) < __fadd_rn(box[3] / 2, NMS_MARGIN) && fabsf(ry) < __fadd_rn(box[4] / 2, NMS_MARGIN);
}
__device__ __forceinline__ bool seg_intersection(Pt p1, Pt p0, Pt q1, Pt q0, Pt* ans) {
    if (!rect_cross(p0, p1, q0, q1)) return false;
    const float s1 = cross3(q0, p1, p0), s2 = cross3(p1, q1, p0), s3 = cross3(p0, q1, q0), s4 = cross3(q1, p1, q0);
    if (!(__fmul_rn(s1, s2) > 0.f && __fmul_rn(s3, s4) > 0.f)) return false;
    const float s5 = cross3(q1, p1, p0);
    const float den = __fsub_rn(s5, s1);
    if (fabsf(den) > NMS_EPS) {
        ans->x = __fdiv_rn(__fsub_rn(__fmul_rn(s5, q0.x), __fmul_rn(s1, q1.x)), den);
        ans->y = __fdiv_rn(__fsub_rn(__fmul_rn(s5, q0.y), __fmul_rn(s1, q1.y)), den);
    } else {
        const float a0 = __fsub_rn(p0.y, p1.y), b0 = __fsub_rn(p1.x, p0.x);
        const float c0 = __fsub_rn(__fmul_rn(p0.x, p1.y), __fmul_rn(p1.x, p0.y));
        const float a1 = __fsub_rn(q0.y, q1.y), b1 = __fsub_rn(q1.x, q0.x);
        const float c1 = __fsub_rn(__fmul_rn(q0.x, q1.y), __fmul_rn(q1.x, q0.y));
        const float D = __fsub_rn(__fmul_rn(a0, b1), __fmul_rn(a1, b0));
        ans->x = __fdiv_rn(__fsub_rn(__fmul_rn(b0, c1), __fmul_rn(b1, c0)), D);
        ans->y = __fdiv_rn(__fsub_rn(__fmul_rn(a1, c0), __fmul_rn(a0, c1)), D);
    }
    return true;
}
__device__ __forceinline__ void box_corners(const float* box, Pt* out /* 5 */) {
    const float hx = box[3] / 2, hy = box[4] / 2;
    float s, c;
    sincosf(box[6], &s, &c);
    const float x1 = __fsub_rn(box[0], hx), x2 = __fadd_rn(box[0], hx);
    const float y1 = __fsub_rn(box[1], hy), y2 = __fadd_rn(box[1], hy);
    const float xs[4] = {x1, x2, x2, x1}, ys[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float px = __fsub_rn(xs[k], box[0]), py = __fsub_rn(ys[k], box[1]);
        out[k].x = __fadd_rn(__fadd_rn(__fmul_rn(px, c), __fmul_rn(py, -s)), box[0]);
        out[k].y = __fadd_rn(__fadd_rn(__fmul_rn(px, s), __fmul_rn(py, c)), box[1]);
    }
    out[4] = out[0];
}

__device__ float box_overlap(const float* a, const float* b) {
    // disjoint bounding circles (with room for the 1e-2 in-box margin): no intersection point, no corner inside
    {
        const float dx = a[0] - b[0], dy = a[1] - b[1];
        const float ra = 0.5f * sqrtf(a[3] * a[3] + a[4] * a[4]), rb = 0.5f * sqrtf(b[3] * b[3] + b[4] * b[4]);
        const float r = ra + rb + 0.05f;
        if (dx * dx + dy * dy > r * r) return 0.f;
    }
    Pt ca[5], cb[5], pts[16];
    box_corners(a, ca);
    box_corners(b, cb);
    Pt ctr = {0.f, 0.f};
    int cnt = 0;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (seg_intersection(ca[i + 1], ca[i], cb[j + 1], cb[j], &pts[cnt])) {
                ctr.x = __fadd_rn(ctr.x, pts[cnt].x);
                ctr.y = __fadd_rn(ctr.y, pts[cnt].y);
                ++cnt;
            }
    for (int k = 0; k < 4; ++k) {
        if (in_box(a, cb[k])) { ctr.x = __fadd_rn(ctr.x, cb[k].x); ctr.y = __fadd_rn(ctr.y, cb[k].y); pts[cnt++] = cb[k]; }
        if (in_box(b, ca[k])) { ctr.x = __fadd_rn(ctr.x, ca[k].x); ctr.y = __fadd_rn(ctr.y, ca[k].y); pts[cnt++] = ca[k]; }
    }
    if (cnt == 0) return 0.f;
    ctr.x = __fdiv_rn(ctr.x, (float)cnt);
    ctr.y = __fdiv_rn(ctr.y, (float)cnt);
    float ang[16];
    for (int i = 0; i < cnt; ++i) ang[i] = atan2f(__fsub_rn(pts[i].y, ctr.y), __fsub_rn(pts[i].x, ctr.x));
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (ang[i] > ang[i + 1]) {
                const Pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
                const float ta = ang[i]; ang[i] = ang[i + 1]; ang[i + 1] = ta;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        const Pt u = {__fsub_rn(pts[k].x, pts[0].x), __fsub_rn(pts[k].y, pts[0].y)};
        const Pt v = {__fsub_rn(pts[k + 1].x, pts[0].x), __fsub_rn(pts[k + 1].y, pts[0].y)};
        area = __fadd_rn(area, cross2(u, v));
    }
    return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_bev(const float* a, const float* b) {
    const float sa = __fmul_rn(a[3], a[4]), sb = __fmul_rn(b[3], b[4]);
    const float so = box_overlap(a, b);
    return __fdiv_rn(so, fmaxf(__fsub_rn(__fadd_rn(sa, sb), so), NMS_EPS));
}

__global__ void __launch_bounds__(256)
k_boxes_iou_bev(const float* __restrict__ a, int na, const float* __restrict__ b, int nb, float* __restrict__ ans) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (long long)na * nb) return;
    const int i = (int)(t / nb), j = (int)(t % nb);
    float ba[7], bb[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) { ba[k] = a[(size_t)i * 7 + k]; bb[k] = b[(size_t)j * 7 + k]; }
    ans[t] = iou_bev(ba, bb);
}

// mask[set][i][w] bit j  <-  IoU(box i, box 64 w + j) > thresh, for j > i only.  grid (col blocks, row blocks, sets).
__global__ void __launch_bounds__(NMS_BLOCK)
k_nms_mask(const float* __restrict__ boxes, const int* __restrict__ n_valid, int n, int col_blocks, float thresh,
           unsigned long long* __restrict__ mask) {
    const int set = blockIdx.z;
    const int row_start = blockIdx.y, col_start = blockIdx.x;
    if (row_start > col_start) return;
    const int nv = n_valid ? min(n, max(0, n_valid[set])) : n;
    if (row_start * NMS_BLOCK >= nv || col_start * NMS_BLOCK >= nv) return;
    const float* bx = boxes + (size_t)set * n * 7;
    const int row_size = min(nv - row_start * NMS_BLOCK, NMS_BLOCK);
    const int col_size = min(nv - col_start * NMS_BLOCK, NMS_BLOCK);
    __shared__ float cols[NMS_BLOCK * 7];
    if (threadIdx.x < col_size)
#pragma unroll
        for (int k = 0; k < 7; ++k) cols[threadIdx.x * 7 + k] = bx[(size_t)(col_start * NMS_BLOCK + threadIdx.x) * 7 + k];
    __syncthreads();
    if (threadIdx.x < row_size) {
        const int i = row_start * NMS_BLOCK + threadIdx.x;
        float cur[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) cur[k] = bx[(size_t)i * 7 + k];
        unsigned long long bits = 0ull;
        const int start = row_start == col_start ? threadIdx.x + 1 : 0;
        for (int j = start; j < col_size; ++j)
            if (iou_bev(cur, cols + j * 7) > thresh) bits |= 1ull << j;
        mask[((size_t)set * n + i) * col_blocks + col_start] = bits;
    }
}

// One CTA per set.  Row blocks in order: the 64 decisions of a block depend only on the block's own suppression word
// and the diagonal mask words (thread 0, 64 cheap steps); the kept rows' other words are then OR-ed into the
// suppression words by all threads.  Stops once max_keep boxes are kept.
__global__ void __launch_bounds__(NMS_BLOCK)
k_nms_sweep(const unsigned long long* __restrict__ mask, const int* __restrict__ n_valid, int n, int col_blocks,
            int max_keep, int* __restrict__ keep, int* __restrict__ num_keep) {
    extern __shared__ unsigned long long sm[];
    unsigned long long* remv = sm;                        // col_blocks words
    __shared__ unsigned long long kept_bits_s;
    __shared__ unsigned long long diag_s[NMS_BLOCK];
    __shared__ int num_s;
    const int set = blockIdx.x;
    const int nv = n_valid ? min(n, max(0, n_valid[set])) : n;
    const unsigned long long* mk = mask + (size_t)set * n * col_blocks;
    int* kp = keep + (size_t)set * max_keep;
    for (int w = threadIdx.x; w < col_blocks; w += NMS_BLOCK) remv[w] = 0ull;
    for (int k = threadIdx.x; k < max_keep; k += NMS_BLOCK) kp[k] = -1;
    if (threadIdx.x == 0) num_s = 0;
    __syncthreads();
    const int row_blocks = (nv + NMS_BLOCK - 1) / NMS_BLOCK;
    for (int rb = 0; rb < row_blocks; ++rb) {
        const int rows = min(NMS_BLOCK, nv - rb * NMS_BLOCK);
        // diagonal word of this thread's row (k_nms_mask writes only words (row block <= column block) below n_valid)
        diag_s[threadIdx.x] = threadIdx.x < rows ? mk[(size_t)(rb * NMS_BLOCK + threadIdx.x) * col_blocks + rb] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[rb], kept = 0ull;
            int num = num_s;
            for (int i = 0; i < rows && num < max_keep; ++i)
                if (!((cur >> i) & 1ull)) {
                    kp[num++] = rb * NMS_BLOCK + i;
                    kept |= 1ull << i;
                    cur |= diag_s[i];
                }
            kept_bits_s = kept;
            num_s = num;
        }
        __syncthreads();
        if (num_s >= max_keep) break;
        const unsigned long long kept = kept_bits_s;
        for (int w = rb + 1 + threadIdx.x; w < row_blocks; w += NMS_BLOCK) {
            unsigned long long acc = remv[w];
            for (unsigned long long rem = kept; rem; rem &= rem - 1ull) {
                const int i = __ffsll((long long)rem) - 1;
                acc |= mk[(size_t)(rb * NMS_BLOCK + i) * col_blocks + w];
            }
            remv[w] = acc;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) num_keep[set] = num_s;
}

}  // namespace

extern "C" {

// boxes_iou_bev_gpu: a (na,7), b (nb,7) -> ans (na,nb) f32
int ud_boxes_iou_bev(const float* a, int na, const float* b, int nb, float* ans, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (na < 0 || nb < 0) return UD_ERR_INVALID_ARG;
    if (na == 0 || nb == 0) return UD_OK;
    if (!a || !b || !ans) return UD_ERR_INVALID_ARG;
    k_boxes_iou_bev<<<ud_blocks((long long)na * nb, 256), 256, 0, stream>>>(a, na, b, nb, ans);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

int ud_nms_bev_workspace_bytes(int n_sets, int n, size_t* bytes) {
    if (!bytes || n_sets < 0 || n < 0) return UD_ERR_INVALID_ARG;
    const size_t col_blocks = (size_t)(n + NMS_BLOCK - 1) / NMS_BLOCK;
    *bytes = ud_align_up((size_t)n_sets * n * col_blocks * sizeof(unsigned long long)) + 256;
    return UD_OK;
}

// Batched greedy NMS.  boxes (n_sets, n, 7) f32, each set sorted by descending score; n_valid (n_sets) device int32 or
// NULL (= n): only the first n_valid[s] boxes of a set take part.  keep (n_sets, max_keep) int32 receives the kept
// positions in ascending order, padded with -1; num_keep (n_sets) int32 their count (<= max_keep: the sweep stops
// there, like the reference's selected[:post_max_size]).
int ud_nms_bev(const float* boxes, const int* n_valid, int n_sets, int n, float thresh, int max_keep, int* keep,
               int* num_keep, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n_sets < 0 || n < 0 || max_keep < 1 || !keep || !num_keep) return UD_ERR_INVALID_ARG;
    if (n_sets == 0) return UD_OK;
    if (n == 0) {
        UD_CUDA_TRY(cudaMemsetAsync(keep, 0xff, sizeof(int) * (size_t)n_sets * max_keep, stream));
        UD_CUDA_TRY(cudaMemsetAsync(num_keep, 0, sizeof(int) * (size_t)n_sets, stream));
        return UD_OK;
    }
    if (!boxes) return UD_ERR_INVALID_ARG;
    const int col_blocks = (n + NMS_BLOCK - 1) / NMS_BLOCK;
    if (n_sets > 65535 || col_blocks > 65535) return UD_ERR_INVALID_ARG;
    size_t need = 0;
    ud_nms_bev_workspace_bytes(n_sets, n, &need);
    if (!workspace || workspace_bytes < need) return UD_ERR_WORKSPACE;
    unsigned long long* mask = (unsigned long long*)ud_align_up((size_t)(uintptr_t)workspace);
    dim3 grid(col_blocks, col_blocks, n_sets);
    k_nms_mask<<<grid, NMS_BLOCK, 0, stream>>>(boxes, n_valid, n, col_blocks, thresh, mask);
    UD_LAUNCH_CHECK();
    const size_t smem = sizeof(unsigned long long) * (size_t)col_blocks;
    if (smem > 40 * 1024) return UD_ERR_UNSUPPORTED;
    k_nms_sweep<<<n_sets, NMS_BLOCK, smem, stream>>>(mask, n_valid, n, col_blocks, max_keep, keep, num_keep);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
