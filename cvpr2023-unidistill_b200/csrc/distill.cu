// Fused cross-modality distillation losses for sm_100a.
//
// Replaces the torch-op chains (2 grid_samples + ~15 tiny kernels + 2 bmm per loss, a CPU/numpy hop for
// the box corners and the Gaussian mask) of the reference's distill scripts
//   exps/multisensor_fusion/nuscenes/BEVFusion/BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar.py
//     :73-97   center_to_corner_box2d          -> k_box_corners
//     :196-245 FeatureDistillLoss              -> k_box_distill<mode 0>   (one CTA per GT box)
//     :248-323 BEVDistillLoss (relation/Gram)  -> k_box_distill<mode 1>
//     :100-178,326-385 ResponseDistillLoss + calculate_box_mask_gaussian -> k_response (one thread/pixel)
// Teacher and student maps are read once, inside the same kernel.  Each kernel has a backward twin that
// recomputes the samples and emits gradients w.r.t. the student tensors only (the teacher is frozen).
//
// Quirks reproduced on purpose: grid_sample with align_corners=False and zero padding, the
// swapped coordinate columns (row <- x cell, col <- y cell), the Gaussian mask being drawn for padded GT
// rows too (radius 0 at the map centre), int() truncation of the Gaussian centres.
#include "common.cuh"

namespace {

constexpr int NKP = 9;   // 4 corners + centre + 4 edge mid-points

struct Taps {
    int x0, y0;
    float w00, w01, w10, w11;   // (y0,x0) (y0,x1) (y1,x0) (y1,x1)
};

// ATen grid_sampler_2d, bilinear, padding zeros, align_corners = false
__device__ __forceinline__ Taps make_taps(float gx, float gy, int W, int H) {
    float ix = ((gx + 1.f) * W - 1.f) / 2.f;
    float iy = ((gy + 1.f) * H - 1.f) / 2.f;
    float fx = floorf(ix), fy = floorf(iy);
    Taps t;
    t.x0 = (int)fx;
    t.y0 = (int)fy;
    float ax = ix - fx, ay = iy - fy;          // distance to the west / north tap
    t.w00 = (1.f - ax) * (1.f - ay);
    t.w01 = ax * (1.f - ay);
    t.w10 = (1.f - ax) * ay;
    t.w11 = ax * ay;
    return t;
}

// Memory layout of a (B, C, H, W) map: element (c, y, x) of a sample = base + c * lc + y * ly + x * lx.
//   NCHW: lc = H W, ly = W, lx = 1      NHWC (torch channels_last, what the dense trunk runs in): lc = 1, ly = W C, lx = C
// With NHWC the C channels of a tap are contiguous, so the threads of a block (one channel each) read whole lines, and the
// four layout copies per step that fed these kernels NCHW maps (0.15 ms each at 512 channels) are gone.
struct MapLayout { int lc, ly, lx; };
__host__ __device__ __forceinline__ MapLayout map_layout(int nhwc, int C, int H, int W) {
    MapLayout l;
    if (nhwc) { l.lc = 1; l.ly = W * C; l.lx = C; } else { l.lc = H * W; l.ly = W; l.lx = 1; }
    return l;
}

__device__ __forceinline__ float sample(const float* __restrict__ plane, const Taps& t, int W, int H, const MapLayout l) {
    float v = 0.f;
    const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    if (yin0 && xin0) v += __ldg(plane + t.y0 * l.ly + t.x0 * l.lx) * t.w00;
    if (yin0 && xin1) v += __ldg(plane + t.y0 * l.ly + (t.x0 + 1) * l.lx) * t.w01;
    if (yin1 && xin0) v += __ldg(plane + (t.y0 + 1) * l.ly + t.x0 * l.lx) * t.w10;
    if (yin1 && xin1) v += __ldg(plane + (t.y0 + 1) * l.ly + (t.x0 + 1) * l.lx) * t.w11;
    return v;
}

__device__ __forceinline__ void scatter(float* plane, const Taps& t, int W, int H, float g, const MapLayout l) {
    const bool xin0 = t.x0 >= 0 && t.x0 < W, xin1 = t.x0 + 1 >= 0 && t.x0 + 1 < W;
    const bool yin0 = t.y0 >= 0 && t.y0 < H, yin1 = t.y0 + 1 >= 0 && t.y0 + 1 < H;
    if (yin0 && xin0) atomicAdd(plane + t.y0 * l.ly + t.x0 * l.lx, g * t.w00);
    if (yin0 && xin1) atomicAdd(plane + t.y0 * l.ly + (t.x0 + 1) * l.lx, g * t.w01);
    if (yin1 && xin0) atomicAdd(plane + (t.y0 + 1) * l.ly + t.x0 * l.lx, g * t.w10);
    if (yin1 && xin1) atomicAdd(plane + (t.y0 + 1) * l.ly + (t.x0 + 1) * l.lx, g * t.w11);
}

// key-points of one box from its 4 BEV corners (cell units), normalised and column-swapped like the
// reference (:200-226).  Thread p (< 9) fills taps[p].
__device__ __forceinline__ void keypoint_taps(const float* __restrict__ c8, int p, int W, int H, Taps* out) {
    float x, y;
    if (p < 4) {
        x = c8[p * 2];
        y = c8[p * 2 + 1];
    } else if (p == 4) {
        x = (c8[0] + c8[2] + c8[4] + c8[6]) / 4.f;
        y = (c8[1] + c8[3] + c8[5] + c8[7]) / 4.f;
    } else {
        const int a = (p == 8) ? 0 : p - 5;        // edges (0,1) (1,2) (2,3) (0,3)
        const int b = (p == 8) ? 3 : p - 4;
        x = (c8[a * 2] + c8[b * 2]) / 2.f;
        y = (c8[a * 2 + 1] + c8[b * 2 + 1]) / 2.f;
    }
    const float w = (float)W, h = (float)H;
    const float nx = (x - w / 2.f) / (w / 2.f);
    const float ny = (y - h / 2.f) / (h / 2.f);
    // swapped: grid[...,0] (width axis) <- ny, grid[...,1] (height axis) <- nx
    *out = make_taps(ny, nx, W, H);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    v = ud_warp_sum(v);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) red[wid] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) r += red[i];
    return r;
}

// ------------------------------------------------------------------------------------------------
// mode 0: feature L1     out[box] = mean_{p,c} |S_s - S_t|
// mode 1: relation       out[box] = mean_{i,j} |G_s[i,j] - G_t[i,j]|,  G = N N^T, N[i] = S[i] / (|S[i]| + 1e-4)
template <int MODE>
__global__ void __launch_bounds__(256)
k_box_distill_fwd(const float* __restrict__ s_map, const float* __restrict__ t_map, int C, int H, int W,
                  const MapLayout ls, const MapLayout lt,
                  const float* __restrict__ corners, const unsigned char* __restrict__ valid, int M,
                  float* __restrict__ out) {
    extern __shared__ float sm[];            // MODE 1: S_s[9][C], S_t[9][C]
    __shared__ Taps taps[NKP];
    __shared__ float red[8];
    __shared__ float inv_s[NKP], inv_t[NKP];
    __shared__ float gram_s[NKP * NKP], gram_t[NKP * NKP];
    const int box = blockIdx.x;
    const int b = box / M;
    if (!valid[box]) {
        if (threadIdx.x == 0) out[box] = 0.f;
        return;
    }
    if (threadIdx.x < NKP) keypoint_taps(corners + (size_t)box * 8, threadIdx.x, W, H, &taps[threadIdx.x]);
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const float* sb = s_map + (size_t)b * C * plane;
    const float* tb = t_map + (size_t)b * C * plane;
    if (MODE == 0) {
        float acc = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x)
#pragma unroll
            for (int p = 0; p < NKP; ++p)
                acc += fabsf(sample(sb + c * ls.lc, taps[p], W, H, ls) - sample(tb + c * lt.lc, taps[p], W, H, lt));
        acc = block_sum(acc, red);
        if (threadIdx.x == 0) out[box] = acc / (float)(NKP * C);
        return;
    }
    float* Ss = sm;
    float* St = sm + NKP * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
#pragma unroll
        for (int p = 0; p < NKP; ++p) {
            Ss[p * C + c] = sample(sb + c * ls.lc, taps[p], W, H, ls);
            St[p * C + c] = sample(tb + c * lt.lc, taps[p], W, H, lt);
        }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = wid; r < 2 * NKP; r += nw) {
        const float* row = (r < NKP) ? Ss + r * C : St + (r - NKP) * C;
        float ss = 0.f;
        for (int c = lane; c < C; c += 32) ss = fmaf(row[c], row[c], ss);
        ss = ud_warp_sum(ss);
        if (lane == 0) {
            float inv = 1.f / (sqrtf(ss) + 1e-4f);
            if (r < NKP) inv_s[r] = inv; else inv_t[r - NKP] = inv;
        }
    }
    __syncthreads();
    for (int e = wid; e < 2 * NKP * NKP; e += nw) {
        const int which = e / (NKP * NKP);
        const int ij = e - which * NKP * NKP;
        const int i = ij / NKP, j = ij - i * NKP;
        const float* S = which ? St : Ss;
        float d = 0.f;
        for (int c = lane; c < C; c += 32) d = fmaf(S[i * C + c], S[j * C + c], d);
        d = ud_warp_sum(d);
        if (lane == 0) {
            if (which) gram_t[ij] = d * inv_t[i] * inv_t[j];
            else gram_s[ij] = d * inv_s[i] * inv_s[j];
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float acc = 0.f;
        for (int e = 0; e < NKP * NKP; ++e) acc += fabsf(gram_s[e] - gram_t[e]);
        out[box] = acc / (float)(NKP * NKP);
    }
}

// grad_s_map (zero-initialised by the caller) += d out[box] / d s_map * upstream[box]
template <int MODE>
__global__ void __launch_bounds__(256)
k_box_distill_bwd(const float* __restrict__ s_map, const float* __restrict__ t_map, int C, int H, int W,
                  const MapLayout ls, const MapLayout lt,
                  const float* __restrict__ corners, const unsigned char* __restrict__ valid, int M,
                  const float* __restrict__ upstream, float* __restrict__ grad_s) {
    extern __shared__ float sm[];            // MODE 1: S_s[9][C], S_t[9][C] (S_t later reused for dN)
    __shared__ Taps taps[NKP];
    __shared__ float red[8];
    __shared__ float nrm_s[NKP], inv_s[NKP], inv_t[NKP], dots[NKP];
    __shared__ float gram_s[NKP * NKP], gram_t[NKP * NKP], Dm[NKP * NKP];
    const int box = blockIdx.x;
    const int b = box / M;
    if (!valid[box]) return;
    const float up = upstream[box];
    if (threadIdx.x < NKP) keypoint_taps(corners + (size_t)box * 8, threadIdx.x, W, H, &taps[threadIdx.x]);
    __syncthreads();
    const size_t plane = (size_t)H * W;
    const float* sb = s_map + (size_t)b * C * plane;
    const float* tb = t_map + (size_t)b * C * plane;
    float* gb = grad_s + (size_t)b * C * plane;
    if (MODE == 0) {
        const float scale = up / (float)(NKP * C);
        for (int c = threadIdx.x; c < C; c += blockDim.x)
#pragma unroll
            for (int p = 0; p < NKP; ++p) {
                float d = sample(sb + c * ls.lc, taps[p], W, H, ls) - sample(tb + c * lt.lc, taps[p], W, H, lt);
                float g = d > 0.f ? scale : (d < 0.f ? -scale : 0.f);
                if (g != 0.f) scatter(gb + c * ls.lc, taps[p], W, H, g, ls);
            }
        return;
    }
    float* Ss = sm;
    float* St = sm + NKP * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x)
#pragma unroll
        for (int p = 0; p < NKP; ++p) {
            Ss[p * C + c] = sample(sb + c * ls.lc, taps[p], W, H, ls);
            St[p * C + c] = sample(tb + c * lt.lc, taps[p], W, H, lt);
        }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int r = wid; r < 2 * NKP; r += nw) {
        const float* row = (r < NKP) ? Ss + r * C : St + (r - NKP) * C;
        float ss = 0.f;
        for (int c = lane; c < C; c += 32) ss = fmaf(row[c], row[c], ss);
        ss = ud_warp_sum(ss);
        if (lane == 0) {
            float n = sqrtf(ss);
            if (r < NKP) { nrm_s[r] = n; inv_s[r] = 1.f / (n + 1e-4f); }
            else inv_t[r - NKP] = 1.f / (n + 1e-4f);
        }
    }
    __syncthreads();
    for (int e = wid; e < 2 * NKP * NKP; e += nw) {
        const int which = e / (NKP * NKP);
        const int ij = e - which * NKP * NKP;
        const int i = ij / NKP, j = ij - i * NKP;
        const float* S = which ? St : Ss;
        float d = 0.f;
        for (int c = lane; c < C; c += 32) d = fmaf(S[i * C + c], S[j * C + c], d);
        d = ud_warp_sum(d);
        if (lane == 0) {
            if (which) gram_t[ij] = d * inv_t[i] * inv_t[j];
            else gram_s[ij] = d * inv_s[i] * inv_s[j];
        }
    }
    __syncthreads();
    if (threadIdx.x < NKP * NKP) {
        float d = gram_s[threadIdx.x] - gram_t[threadIdx.x];
        Dm[threadIdx.x] = (d > 0.f ? up : (d < 0.f ? -up : 0.f)) / (float)(NKP * NKP);
    }
    __syncthreads();
    // dN[i][c] = sum_j (D[i][j] + D[j][i]) * N_s[j][c]   -> stored over S_t
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float n[NKP];
#pragma unroll
        for (int j = 0; j < NKP; ++j) n[j] = Ss[j * C + c] * inv_s[j];
#pragma unroll
        for (int i = 0; i < NKP; ++i) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < NKP; ++j) a = fmaf(Dm[i * NKP + j] + Dm[j * NKP + i], n[j], a);
            St[i * C + c] = a;
        }
    }
    __syncthreads();
    for (int i = 0; i < NKP; ++i) {
        float part = 0.f;
        for (int c = threadIdx.x; c < C; c += blockDim.x) part = fmaf(St[i * C + c], Ss[i * C + c], part);
        part = block_sum(part, red);
        if (threadIdx.x == 0) dots[i] = part;
    }
    __syncthreads();
    // dS[i] = dN[i] * inv - S[i] * <dN[i], S[i]> * inv^2 / |S[i]|        (n = s / (r + eps))
    for (int c = threadIdx.x; c < C; c += blockDim.x)
#pragma unroll
        for (int i = 0; i < NKP; ++i) {
            float inv = inv_s[i];
            float g = St[i * C + c] * inv;
            if (nrm_s[i] > 0.f) g -= Ss[i * C + c] * dots[i] * inv * inv / nrm_s[i];
            if (g != 0.f) scatter(gb + c * ls.lc, taps[i], W, H, g, ls);
        }
}

// ------------------------------------------------------------------------------------------------
// center_to_corner_box2d (reference :73-97) + conversion to BEV cell units (:478-483).
// boxes (B*M, stride) [x,y,z,dx,dy,dz,yaw,...]; corners (B*M, 4, 2): x0y0, x0y1, x1y1, x1y0 rotated by yaw.
__global__ void k_box_corners(const float* __restrict__ boxes, int n, int stride, float x0, float y0, float cell_x,
                              float cell_y, float* __restrict__ corners) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* bx = boxes + (size_t)i * stride;
    const double cx = bx[0], cy = bx[1], dx = bx[3], dy = bx[4];
    const double c = (double)cosf(bx[6]), s = (double)sinf(bx[6]);   // numpy: float32 sin/cos, float64 products
    const double ox[4] = {-0.5, -0.5, 0.5, 0.5};
    const double oy[4] = {-0.5, 0.5, 0.5, -0.5};
    for (int k = 0; k < 4; ++k) {
        double lx = dx * ox[k], ly = dy * oy[k];
        float gx = (float)(lx * c - ly * s + cx);
        float gy = (float)(lx * s + ly * c + cy);
        corners[(size_t)i * 8 + k * 2 + 0] = (gx - x0) / cell_x;
        corners[(size_t)i * 8 + k * 2 + 1] = (gy - y0) / cell_y;
    }
}

// ------------------------------------------------------------------------------------------------
struct GaussBox {
    int cx, cy, r;
    double inv2s2;   // 1 / (2 sigma^2)
};

// gaussian_radius (reference :131-151), evaluated in fp32 like numpy does on float32 inputs
__device__ __forceinline__ float gaussian_radius_f(float height, float width, float min_overlap) {
    float a1 = 1.f;
    float b1 = height + width;
    float c1 = width * height * (1.f - min_overlap) / (1.f + min_overlap);
    float sq1 = sqrtf(b1 * b1 - 4.f * a1 * c1);
    float r1 = (b1 + sq1) / 2.f;
    float a2 = 4.f;
    float b2 = 2.f * (height + width);
    float c2 = (1.f - min_overlap) * width * height;
    float sq2 = sqrtf(b2 * b2 - 4.f * a2 * c2);
    float r2 = (b2 + sq2) / 2.f;
    float a3 = 4.f * min_overlap;
    float b3 = -2.f * min_overlap * (height + width);
    float c3 = (min_overlap - 1.f) * width * height;
    float sq3 = sqrtf(b3 * b3 - 4.f * a3 * c3);
    float r3 = (b3 + sq3) / 2.f;
    return fminf(r1, fminf(r2, r3));
}

// one thread per (b, pixel); boxes of the sample are staged in smem
// MODE 0: forward partial sums (per block: sum_cls, sum_reg, sum_mask)
// MODE 1: backward: dense grads w.r.t. cls_s / reg_s
template <int MODE>
__global__ void __launch_bounds__(256)
k_response(const float* __restrict__ cls_s, const float* __restrict__ cls_t, const float* __restrict__ reg_s,
           const float* __restrict__ reg_t, int n_cls, int n_reg, int H, int W, const float* __restrict__ boxes,
           int M, int stride, const int* __restrict__ n_rows, float x0, float y0, float cell_x, float cell_y,
           float clamp_lo, float* __restrict__ partial, float* __restrict__ mask_out,
           const float* __restrict__ upstream, float* __restrict__ d_cls, float* __restrict__ d_reg) {
    extern __shared__ __align__(16) unsigned char raw[];
    GaussBox* gb = reinterpret_cast<GaussBox*>(raw);
    __shared__ float red[8];
    const int blocks_per_sample = (H * W + blockDim.x - 1) / blockDim.x;
    const int b = blockIdx.x / blocks_per_sample;
    const int pix = (blockIdx.x - b * blocks_per_sample) * blockDim.x + threadIdx.x;
    const int rows = n_rows[b];
    for (int j = threadIdx.x; j < rows; j += blockDim.x) {
        const float* bx = boxes + ((size_t)b * M + j) * stride;
        float w = bx[3] / cell_x, h = bx[4] / cell_y;
        float rad = gaussian_radius_f(w, h, 0.7f);     // gaussian_radius((w, h)): height <- w, width <- h
        int r = (int)rad;
        if (r < 0) r = 0;
        GaussBox g;
        g.cx = (int)((bx[0] - x0) / cell_x);           // int(): truncation toward zero
        g.cy = (int)((bx[1] - y0) / cell_y);
        g.r = r;
        double sigma = (double)(2 * r + 1) / 6.0;
        g.inv2s2 = 1.0 / (2.0 * sigma * sigma);
        gb[j] = g;
    }
    __syncthreads();
    float m = 0.f;
    const int px = pix % W, py = pix / W;
    const bool live = pix < H * W;
    if (live) {
        for (int j = 0; j < rows; ++j) {
            const GaussBox g = gb[j];
            // the reference skips a box whose (clipped) window is empty or inconsistent: centre further than
            // r outside the map on the low side (negative python slice) -- see draw_umich_gaussian :163-178
            if (g.cx + g.r + 1 <= 0 || g.cy + g.r + 1 <= 0 || g.cx - g.r >= W || g.cy - g.r >= H) continue;
            int dx = px - g.cx, dy = py - g.cy;
            if (dx < -g.r || dx > g.r || dy < -g.r || dy > g.r) continue;
            float v = (float)exp(-(double)(dx * dx + dy * dy) * g.inv2s2);
            m = fmaxf(m, v);
        }
    }
    const size_t plane = (size_t)H * W;
    float dcls = 0.f, dreg = 0.f;
    if (live) {
        const float* cs = cls_s + (size_t)b * n_cls * plane + pix;
        const float* ct = cls_t + (size_t)b * n_cls * plane + pix;
        float ms = -INFINITY, mt = -INFINITY;
        int arg = 0;
        for (int c = 0; c < n_cls; ++c) {
            float a = cs[c * plane];
            if (a > ms) { ms = a; arg = c; }
            float t = 1.f / (1.f + expf(-ct[c * plane] * 0.5f));       // sigmoid(x / 2)
            t = fminf(fmaxf(t, clamp_lo), 1.f - clamp_lo);
            mt = fmaxf(mt, t);
        }
        const float* rs = reg_s + (size_t)b * n_reg * plane + pix;
        const float* rt = reg_t + (size_t)b * n_reg * plane + pix;
        if (MODE == 0) {
            dcls = fabsf(ms - mt) * m;
            float acc = 0.f;
            for (int c = 0; c < n_reg; ++c) acc += fabsf(rs[c * plane] - rt[c * plane]);
            dreg = acc / (float)n_reg * m;
            if (mask_out) mask_out[(size_t)b * plane + pix] = m;
        } else {
            const float g_cls = __ldg(upstream), g_reg = __ldg(upstream + 1);
            float* dc = d_cls + (size_t)b * n_cls * plane + pix;
            float d = ms - mt;
            float g = (d > 0.f ? g_cls : (d < 0.f ? -g_cls : 0.f)) * m;
            for (int c = 0; c < n_cls; ++c) dc[c * plane] = (c == arg) ? g : 0.f;
            float* dr = d_reg + (size_t)b * n_reg * plane + pix;
            const float gr = g_reg * m / (float)n_reg;
            for (int c = 0; c < n_reg; ++c) {
                float e = rs[c * plane] - rt[c * plane];
                dr[c * plane] = e > 0.f ? gr : (e < 0.f ? -gr : 0.f);
            }
        }
    }
    if (MODE == 0) {
        float a = block_sum(dcls, red);
        float bsum = block_sum(dreg, red);
        float c = block_sum(m, red);
        if (threadIdx.x == 0) {
            partial[(size_t)blockIdx.x * 3 + 0] = a;
            partial[(size_t)blockIdx.x * 3 + 1] = bsum;
            partial[(size_t)blockIdx.x * 3 + 2] = c;
        }
    }
}

}  // namespace

extern "C" {

// boxes (n, stride) f32 rows [x,y,z,dx,dy,dz,yaw,...] -> corners (n,4,2) f32 in BEV cell units
int ud_box_corners_bev(const float* boxes, int n, int stride, float x_min, float y_min, float cell_x,
                       float cell_y, float* corners, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (n < 0 || stride < 7 || (n > 0 && (!boxes || !corners))) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    k_box_corners<<<ud_blocks(n, 128), 128, 0, stream>>>(boxes, n, stride, x_min, y_min, cell_x, cell_y, corners);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// mode 0: FeatureDistillLoss per-box terms, mode 1: BEVDistillLoss (relation) per-box terms.
// s_map/t_map (B,C,H,W) f32; corners (B,M,4,2) f32 cell units; valid (B,M) u8; out (B*M) f32 (0 for invalid).
// Same with the memory layout of each map stated: *_nhwc = 0 for (B,C,H,W) contiguous, 1 for torch channels_last (NHWC).
int ud_box_distill_forward_layout(int mode, const float* s_map, int s_nhwc, const float* t_map, int t_nhwc, int B, int C,
                                  int H, int W, const float* corners, const unsigned char* valid, int M, float* out,
                                  void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!s_map || !t_map || !corners || !valid || !out || B < 1 || C < 1 || H < 1 || W < 1 || M < 1)
        return UD_ERR_INVALID_ARG;
    const MapLayout ls = map_layout(s_nhwc, C, H, W), lt = map_layout(t_nhwc, C, H, W);
    if (mode == 0) {
        k_box_distill_fwd<0><<<B * M, 256, 0, stream>>>(s_map, t_map, C, H, W, ls, lt, corners, valid, M, out);
    } else if (mode == 1) {
        size_t smem = (size_t)2 * NKP * C * sizeof(float);
        if (smem > 200 * 1024) return UD_ERR_UNSUPPORTED;
        UD_CUDA_TRY(cudaFuncSetAttribute(k_box_distill_fwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_distill_fwd<1><<<B * M, 256, smem, stream>>>(s_map, t_map, C, H, W, ls, lt, corners, valid, M, out);
    } else {
        return UD_ERR_INVALID_ARG;
    }
    UD_LAUNCH_CHECK();
    return UD_OK;
}

int ud_box_distill_forward(int mode, const float* s_map, const float* t_map, int B, int C, int H, int W,
                           const float* corners, const unsigned char* valid, int M, float* out, void* stream_) {
    return ud_box_distill_forward_layout(mode, s_map, 0, t_map, 0, B, C, H, W, corners, valid, M, out, stream_);
}

// grad_s (B,C,H,W) f32 is fully written (zeroed here, then accumulated); upstream (B*M) f32 = dL/d out[box].
// grad_s has the layout of s_map.
int ud_box_distill_backward_layout(int mode, const float* s_map, int s_nhwc, const float* t_map, int t_nhwc, int B, int C,
                                   int H, int W, const float* corners, const unsigned char* valid, int M,
                                   const float* upstream, float* grad_s, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    const MapLayout ls = map_layout(s_nhwc, C, H, W), lt = map_layout(t_nhwc, C, H, W);
    if (!s_map || !t_map || !corners || !valid || !upstream || !grad_s || B < 1 || C < 1 || H < 1 || W < 1 || M < 1)
        return UD_ERR_INVALID_ARG;
    UD_CUDA_TRY(cudaMemsetAsync(grad_s, 0, (size_t)B * C * H * W * sizeof(float), stream));
    if (mode == 0) {
        k_box_distill_bwd<0><<<B * M, 256, 0, stream>>>(s_map, t_map, C, H, W, ls, lt, corners, valid, M, upstream, grad_s);
    } else if (mode == 1) {
        size_t smem = (size_t)2 * NKP * C * sizeof(float);
        if (smem > 200 * 1024) return UD_ERR_UNSUPPORTED;
        UD_CUDA_TRY(cudaFuncSetAttribute(k_box_distill_bwd<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_distill_bwd<1><<<B * M, 256, smem, stream>>>(s_map, t_map, C, H, W, ls, lt, corners, valid, M, upstream,
                                                           grad_s);
    } else {
        return UD_ERR_INVALID_ARG;
    }
    UD_LAUNCH_CHECK();
    return UD_OK;
}

int ud_box_distill_backward(int mode, const float* s_map, const float* t_map, int B, int C, int H, int W,
                            const float* corners, const unsigned char* valid, int M, const float* upstream,
                            float* grad_s, void* stream_) {
    return ud_box_distill_backward_layout(mode, s_map, 0, t_map, 0, B, C, H, W, corners, valid, M, upstream, grad_s, stream_);
}

// Response distillation, forward.  cls_s (B,n_cls,H,W) student heat-map (already clamp-sigmoid), cls_t teacher
// logits (sigmoid(x/2) + clamp applied here), reg_s/reg_t (B,n_reg,H,W); boxes (B,M,stride) GT rows WITH the
// label column; n_rows (B) int32 = rows before the first all-zero row (reference :109-111).
// partial (n_blocks,3) f32 = per-block [sum |dcls|*mask, sum mean|dreg|*mask, sum mask], n_blocks =
// B * ceil(H*W/256); mask_out (B,H,W) f32 or NULL.
int ud_response_distill_forward(const float* cls_s, const float* cls_t, const float* reg_s, const float* reg_t,
                                int B, int n_cls, int n_reg, int H, int W, const float* boxes, int M, int stride,
                                const int* n_rows, float x_min, float y_min, float cell_x, float cell_y,
                                float clamp_lo, float* partial, float* mask_out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cls_s || !cls_t || !reg_s || !reg_t || !boxes || !n_rows || !partial || B < 1 || M < 1 || stride < 5)
        return UD_ERR_INVALID_ARG;
    int bps = (H * W + 255) / 256;
    size_t smem = (size_t)M * sizeof(GaussBox);
    k_response<0><<<B * bps, 256, smem, stream>>>(cls_s, cls_t, reg_s, reg_t, n_cls, n_reg, H, W, boxes, M, stride,
                                                  n_rows, x_min, y_min, cell_x, cell_y, clamp_lo, partial, mask_out,
                                                  nullptr, nullptr, nullptr);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// d_cls (B,n_cls,H,W), d_reg (B,n_reg,H,W) fully written; upstream: DEVICE pointer to [dL/d(sum_cls), dL/d(sum_reg)]
// (the gradient of the (3,) forward result: no host read of the two scalars in the middle of a backward pass)
int ud_response_distill_backward(const float* cls_s, const float* cls_t, const float* reg_s, const float* reg_t,
                                 int B, int n_cls, int n_reg, int H, int W, const float* boxes, int M, int stride,
                                 const int* n_rows, float x_min, float y_min, float cell_x, float cell_y,
                                 float clamp_lo, const float* upstream, float* d_cls, float* d_reg,
                                 void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!cls_s || !cls_t || !reg_s || !reg_t || !boxes || !n_rows || !upstream || !d_cls || !d_reg || B < 1 || M < 1 || stride < 5)
        return UD_ERR_INVALID_ARG;
    int bps = (H * W + 255) / 256;
    size_t smem = (size_t)M * sizeof(GaussBox);
    k_response<1><<<B * bps, 256, smem, stream>>>(cls_s, cls_t, reg_s, reg_t, n_cls, n_reg, H, W, boxes, M, stride,
                                                  n_rows, x_min, y_min, cell_x, cell_y, clamp_lo, nullptr, nullptr,
                                                  upstream, d_cls, d_reg);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
