// Row-matrix BatchNorm1d (+residual, +ReLU) and the sparse->dense BEV scatter.
//
// BatchNorm: reference spconv_backbone.py:255 builds nn.BatchNorm1d(eps=1e-3, momentum=0.01) after every
// sparse conv and applies it to the (N_active, C) feature matrix (post_act_block :10-58,
// SparseBasicBlock.forward :97-113).  Here one statistics pass + one fused apply pass
// (scale/shift, optional residual add, optional ReLU) replaces 3-4 elementwise passes.
//
// Scatter: reference layers/blocks_2d/det3d/map_to_bev/height_compression.py:11-22
// (`sp_tensor.dense()` -> (B, C, D, H, W) -> view(B, C*D, H, W)); channel index = c*D + z.
#include "common.cuh"
#include "det_reduce.cuh"
#include "tma.cuh"
#include "umma_common.cuh"

namespace {

// ---- statistics: per-channel sum and sum of squares over the active rows ----------------------
// block = 256 threads = (256/C) row lanes x C channels, C in {16,32,64,128}
__global__ void __launch_bounds__(256)
k_bn_stats(const float* __restrict__ x, int n, int C, double* __restrict__ sums, unsigned char* __restrict__ ws,
           int group) {
    __shared__ float s_sum[256];
    __shared__ float s_sq[256];
    __shared__ int flag_s;
    const int c = threadIdx.x % C;
    const int rl = threadIdx.x / C;
    const int RL = blockDim.x / C;
    float a = 0.f, b = 0.f;
    for (long long r = (long long)blockIdx.x * RL + rl; r < n; r += (long long)gridDim.x * RL) {
        float v = x[r * C + c];
        a += v;
        b = fmaf(v, v, b);
    }
    s_sum[threadIdx.x] = a;
    s_sq[threadIdx.x] = b;
    __syncthreads();
    // this block's partial of column tid (tid < C: sum, tid in [C, 2C): sum of squares), then the fixed-order reduction
    double v = 0.0;
    if (threadIdx.x < 2 * C) {
        const float* src = threadIdx.x < C ? s_sum : s_sq;
        const int cc = threadIdx.x % C;
        for (int j = 0; j < RL; ++j) v += (double)src[j * C + cc];
    }
    ud_det_reduce(v, threadIdx.x, 2 * C, ws, group, sums, &flag_s);
}

// scale = gamma * invstd, shift = beta - mean * scale; optional running-stat update (training).
__global__ void k_bn_finalize(const double* __restrict__ sum, const double* __restrict__ sqsum, int n, int C,
                              const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                              float momentum, float* running_mean, float* running_var,
                              float* __restrict__ scale, float* __restrict__ shift,
                              float* __restrict__ save_mean, float* __restrict__ save_invstd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double cnt = n > 0 ? (double)n : 1.0;
    double mean = sum[c] / cnt;
    double var = sqsum[c] / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f;
    float bt = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = bt - (float)mean * sc;
    if (save_mean) save_mean[c] = (float)mean;
    if (save_invstd) save_invstd[c] = invstd;
    if (running_mean && running_var) {
        double unbiased = n > 1 ? var * cnt / (cnt - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// eval-mode scale/shift from running statistics
__global__ void k_bn_eval_coeffs(const float* __restrict__ running_mean, const float* __restrict__ running_var,
                                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                 int C, float* __restrict__ scale, float* __restrict__ shift) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float invstd = 1.f / sqrtf(running_var[c] + eps);
    float sc = (gamma ? gamma[c] : 1.f) * invstd;
    scale[c] = sc;
    shift[c] = (beta ? beta[c] : 0.f) - running_mean[c] * sc;
}

// y = relu?( x*scale[c] + shift[c] (+ residual) ), float4 vectorised; C % 4 == 0
__global__ void __launch_bounds__(256)
k_bn_apply(const float4* x, long long n4, int C4, const float4* __restrict__ scale,
           const float4* __restrict__ shift, const float4* residual, int relu, float4* y) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n4;
         t += (long long)gridDim.x * blockDim.x) {
        int c4 = (int)(t % C4);
        float4 v = x[t];
        float4 sc = __ldg(scale + c4), sh = __ldg(shift + c4);
        v.x = fmaf(v.x, sc.x, sh.x);
        v.y = fmaf(v.y, sc.y, sh.y);
        v.z = fmaf(v.z, sc.z, sh.z);
        v.w = fmaf(v.w, sc.w, sh.w);
        if (residual) {
            float4 r = residual[t];
            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
        }
        if (relu) {
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        y[t] = v;
    }
}

// ---- BN backward (training) ---------------------------------------------------------------------
// g = dy * (y > 0) if relu else dy.   sum_g[c] = sum g, sum_gx[c] = sum g * xhat, xhat = (x-mean)*invstd
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x, int n,
                int C, const float* __restrict__ mean, const float* __restrict__ invstd, int relu,
                double* __restrict__ sums, unsigned char* __restrict__ ws, int group) {
    __shared__ float s_a[256];
    __shared__ float s_b[256];
    __shared__ int flag_s;
    const int c = threadIdx.x % C;
    const int rl = threadIdx.x / C;
    const int RL = blockDim.x / C;
    const float m = mean[c], is = invstd[c];
    float a = 0.f, b = 0.f;
    for (long long r = (long long)blockIdx.x * RL + rl; r < n; r += (long long)gridDim.x * RL) {
        float g = dy[r * C + c];
        if (relu && !(y[r * C + c] > 0.f)) g = 0.f;
        a += g;
        b = fmaf(g, (x[r * C + c] - m) * is, b);
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    double v = 0.0;
    if (threadIdx.x < 2 * C) {
        const float* src = threadIdx.x < C ? s_a : s_b;
        const int cc = threadIdx.x % C;
        for (int j = 0; j < RL; ++j) v += (double)src[j * C + cc];
    }
    ud_det_reduce(v, threadIdx.x, 2 * C, ws, group, sums, &flag_s);      // sums[0..C) = sum g, [C..2C) = sum g * xhat
}

// dx = gamma*invstd * (g - sum_g/n - xhat * sum_gx/n);  dres = g (optional)
__global__ void __launch_bounds__(256)
k_bn_bwd_apply(const float* __restrict__ dy, const float* __restrict__ y, const float* __restrict__ x,
               long long total, int C, int n, const float* __restrict__ mean, const float* __restrict__ invstd,
               const float* __restrict__ gamma, const double* __restrict__ sum_g,
               const double* __restrict__ sum_gx, int relu, float* __restrict__ dx, float* __restrict__ dres) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        int c = (int)(t % C);
        float g = dy[t];
        if (relu && !(y[t] > 0.f)) g = 0.f;
        float is = invstd[c];
        float xhat = (x[t] - mean[c]) * is;
        float inv_n = 1.f / (float)n;
        float mg = (float)sum_g[c] * inv_n;
        float mgx = (float)sum_gx[c] * inv_n;
        float gm = gamma ? gamma[c] : 1.f;
        dx[t] = gm * is * (g - mg - xhat * mgx);
        if (dres) dres[t] = g;
    }
}

// float4 forms of the two passes (C % 4 == 0, 16-byte aligned tensors).  A thread owns 4 fixed channels (its per-channel
// constants live in registers) and walks rows with a stride that is a multiple of C/4 vectors.
//   reduce: block = 256 threads = (256 / (C/4)) rows x C/4 vectors; per-thread fp32 partials over its rows, then the rows of
//           the block are folded in index order (fixed order => bit-reproducible, like the scalar form).
//   apply:  walks the tensors FROM THE END: the reduce pass has just streamed dy / y / x front to back, so their tails are
//           what the 126 MB L2 still holds.
__global__ void __launch_bounds__(256)
k_bn_bwd_reduce4(const float4* __restrict__ dy, const float4* __restrict__ y, const float4* __restrict__ x, int n, int C4,
                 const float4* __restrict__ mean, const float4* __restrict__ invstd, int relu,
                 double* __restrict__ sums, unsigned char* __restrict__ ws, int group) {
    __shared__ float4 s_a[256];
    __shared__ float4 s_b[256];
    __shared__ int flag_s;
    const int c4 = threadIdx.x % C4;
    const int rl = threadIdx.x / C4;
    const int RL = 256 / C4;
    const float4 m = mean[c4], is = invstd[c4];
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    for (long long r = (long long)blockIdx.x * RL + rl; r < n; r += (long long)gridDim.x * RL) {
        const long long t = r * C4 + c4;
        float4 g = __ldg(dy + t);              // default caching: the apply pass re-reads the tail of all three from L2
        const float4 xv = __ldg(x + t);
        if (relu) {
            const float4 yv = __ldg(y + t);
            if (!(yv.x > 0.f)) g.x = 0.f;
            if (!(yv.y > 0.f)) g.y = 0.f;
            if (!(yv.z > 0.f)) g.z = 0.f;
            if (!(yv.w > 0.f)) g.w = 0.f;
        }
        a.x += g.x; a.y += g.y; a.z += g.z; a.w += g.w;
        b.x = fmaf(g.x, (xv.x - m.x) * is.x, b.x);
        b.y = fmaf(g.y, (xv.y - m.y) * is.y, b.y);
        b.z = fmaf(g.z, (xv.z - m.z) * is.z, b.z);
        b.w = fmaf(g.w, (xv.w - m.w) * is.w, b.w);
    }
    s_a[threadIdx.x] = a;
    s_b[threadIdx.x] = b;
    __syncthreads();
    const int C = 4 * C4;
    double v = 0.0;
    if (threadIdx.x < 2 * C) {
        const float* src = reinterpret_cast<const float*>(threadIdx.x < C ? s_a : s_b);
        const int cc = threadIdx.x % C;
        for (int j = 0; j < RL; ++j) v += (double)src[j * C + cc];
    }
    ud_det_reduce(v, threadIdx.x, 2 * C, ws, group, sums, &flag_s);      // sums[0..C) = sum g, [C..2C) = sum g * xhat
}

__global__ void __launch_bounds__(256)
k_bn_bwd_apply4(const float4* __restrict__ dy, const float4* __restrict__ y, const float4* __restrict__ x,
                long long total4, int C4, int n, const float4* __restrict__ mean, const float4* __restrict__ invstd,
                const float* __restrict__ gamma, const double* __restrict__ sum_g, const double* __restrict__ sum_gx,
                int relu, float4* __restrict__ dx, float4* __restrict__ dres) {
    const long long stride = (long long)gridDim.x * blockDim.x;          // multiple of C4 (256 % C4 == 0)
    long long t = total4 - 1 - ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (t < 0) return;
    const int c4 = (int)(t % C4);
    const float inv_n = 1.f / (float)n;
    const float4 m = mean[c4], is = invstd[c4];
    float4 mg, mgx, k;
    mg.x = (float)sum_g[4 * c4 + 0] * inv_n; mg.y = (float)sum_g[4 * c4 + 1] * inv_n;
    mg.z = (float)sum_g[4 * c4 + 2] * inv_n; mg.w = (float)sum_g[4 * c4 + 3] * inv_n;
    mgx.x = (float)sum_gx[4 * c4 + 0] * inv_n; mgx.y = (float)sum_gx[4 * c4 + 1] * inv_n;
    mgx.z = (float)sum_gx[4 * c4 + 2] * inv_n; mgx.w = (float)sum_gx[4 * c4 + 3] * inv_n;
    k.x = (gamma ? gamma[4 * c4 + 0] : 1.f) * is.x; k.y = (gamma ? gamma[4 * c4 + 1] : 1.f) * is.y;
    k.z = (gamma ? gamma[4 * c4 + 2] : 1.f) * is.z; k.w = (gamma ? gamma[4 * c4 + 3] : 1.f) * is.w;
    for (; t >= 0; t -= stride) {
        float4 g = __ldcs(dy + t);
        const float4 xv = __ldcs(x + t);
        if (relu) {
            const float4 yv = __ldcs(y + t);
            if (!(yv.x > 0.f)) g.x = 0.f;
            if (!(yv.y > 0.f)) g.y = 0.f;
            if (!(yv.z > 0.f)) g.z = 0.f;
            if (!(yv.w > 0.f)) g.w = 0.f;
        }
        float4 o;       // same expression as the scalar kernel: gamma * invstd * (g - mean_g - xhat * mean_gx)
        o.x = k.x * (g.x - mg.x - ((xv.x - m.x) * is.x) * mgx.x);
        o.y = k.y * (g.y - mg.y - ((xv.y - m.y) * is.y) * mgx.y);
        o.z = k.z * (g.z - mg.z - ((xv.z - m.z) * is.z) * mgx.z);
        o.w = k.w * (g.w - mg.w - ((xv.w - m.w) * is.w) * mgx.w);
        dx[t] = o;
        if (dres) dres[t] = g;
    }
}

// ---- scatter to BEV ---------------------------------------------------------------------------
// CTA = 32 consecutive rows (rows of one conv output are sorted by (b,z,y,x), so they are mostly x-adjacent) x all C
// channels.  The 32 x C voxel tile is staged in shared memory by TMA: C/32 tensor loads (cp.async.bulk.tensor.2d) of a
// [32 rows x 32 channels] box each, 128-byte swizzle, rows past the end read as zero.  The NCHW writes then walk the tile
// channel by channel with lanes = voxels, so x-adjacent voxels write adjacent floats; the swizzle (16-byte chunk index XOR
// row & 7) spreads a column read over 8 bank groups (4-way conflict instead of 32-way for a dense 128-byte pitch).
__global__ void __launch_bounds__(256)
k_scatter_bev(const __grid_constant__ CUtensorMap feat_map, const int4* __restrict__ coords, int n, int C, int D, int H,
              int W, float* __restrict__ out) {
    extern __shared__ __align__(1024) unsigned char tile_raw[];   // C/32 boxes of [32 rows][32 floats], swizzled
    __shared__ long long base_off[32];
    __shared__ __align__(8) unsigned long long bar;
    const int row0 = blockIdx.x * 32;
    const int rows = min(32, n - row0);
    const uint32_t bar_a = ud_umma::smem_u32(&bar);
    if (threadIdx.x == 0) {
        ud_umma::mbar_init(bar_a, 1);
        ud_umma::fence_barrier_init();
    }
    if (threadIdx.x < 32) {
        long long off = -1;
        if (threadIdx.x < rows) {
            int4 q = coords[row0 + threadIdx.x];   // b, z, y, x
            off = (((long long)q.x * C * D + q.y) * H + q.z) * W + q.w;
        }
        base_off[threadIdx.x] = off;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ud_umma::mbar_arrive_expect_tx(bar_a, 32u * (uint32_t)C * 4u);
        for (int b = 0; b < C / 32; ++b) tma_load_2d(ud_umma::smem_u32(tile_raw) + b * 4096, &feat_map, b * 32, row0, bar_a);
    }
    ud_umma::mbar_wait(bar_a, 0u);
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const long long off = base_off[lane];
    const long long cstride = (long long)D * H * W;
    if (off >= 0)
        for (int c = wid; c < C; c += nw) {
            const int j = c & 31;
            const float v = *reinterpret_cast<const float*>(tile_raw + (c >> 5) * 4096 + lane * 128 +
                                                            ((((j >> 2) ^ (lane & 7)) << 4) | ((j & 3) << 2)));
            out[off + c * cstride] = v;
        }
}

// Channel counts that are not a multiple of 32 (e.g. the 5-channel pillar means of BASELINE configs[0]): the tile is staged
// by the threads themselves with a padded pitch.
__global__ void __launch_bounds__(256)
k_scatter_bev_generic(const float* __restrict__ feat, const int4* __restrict__ coords, int n, int C, int D, int H,
                      int W, float* __restrict__ out) {
    extern __shared__ float tile_g[];   // [32][C+1]
    __shared__ long long base_off[32];
    const int row0 = blockIdx.x * 32;
    const int rows = min(32, n - row0);
    for (int e = threadIdx.x; e < 32 * C; e += blockDim.x) {
        int r = e / C, c = e - r * C;
        tile_g[r * (C + 1) + c] = r < rows ? feat[(size_t)(row0 + r) * C + c] : 0.f;
    }
    if (threadIdx.x < 32) {
        long long off = -1;
        if (threadIdx.x < rows) {
            int4 q = coords[row0 + threadIdx.x];   // b, z, y, x
            off = (((long long)q.x * C * D + q.y) * H + q.z) * W + q.w;
        }
        base_off[threadIdx.x] = off;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const long long off = base_off[lane];
    const long long cstride = (long long)D * H * W;
    if (off >= 0)
        for (int c = wid; c < C; c += nw) out[off + c * cstride] = tile_g[lane * (C + 1) + c];
}

__global__ void __launch_bounds__(256)
k_gather_bev(const float* __restrict__ dense, const int4* __restrict__ coords, int n, int C, int D, int H,
             int W, float* __restrict__ feat) {
    extern __shared__ float tile[];   // [32][C+1]
    __shared__ long long base_off[32];
    const int row0 = blockIdx.x * 32;
    const int rows = min(32, n - row0);
    if (threadIdx.x < 32) {
        long long off = -1;
        if (threadIdx.x < rows) {
            int4 q = coords[row0 + threadIdx.x];
            off = (((long long)q.x * C * D + q.y) * H + q.z) * W + q.w;
        }
        base_off[threadIdx.x] = off;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int wid = threadIdx.x >> 5;
    const int nw = blockDim.x >> 5;
    const long long off = base_off[lane];
    const long long cstride = (long long)D * H * W;
    for (int c = wid; c < C; c += nw) tile[lane * (C + 1) + c] = off >= 0 ? dense[off + c * cstride] : 0.f;
    __syncthreads();
    for (int e = threadIdx.x; e < rows * C; e += blockDim.x) {
        int r = e / C, c = e - r * C;
        feat[(size_t)(row0 + r) * C + c] = tile[r * (C + 1) + c];
    }
}

}  // namespace

extern "C" {

static int ud_bn_reduce_grid(int n, int C) {
    const int RL = 256 / C;
    const long long want = ((long long)n + RL - 1) / RL;
    return (int)(want < UD_NUM_SMS * 8 ? (want > 0 ? want : 1) : UD_NUM_SMS * 8);
}

// Scratch of the fixed-order reductions behind ud_bn_stats and ud_bn_backward for (n, C) inputs.
int ud_bn_workspace_bytes(int n, int C, size_t* bytes) {
    if (!bytes || n < 0 || C < 1 || C > 128 || 256 % C != 0) return UD_ERR_INVALID_ARG;
    *bytes = ud_det_reduce_bytes(ud_bn_reduce_grid(n, C), 2 * C);
    return UD_OK;
}

// sums: (2, C) f64 = per-channel sum | sum of squares, fully written (fixed summation order: bit-reproducible).
// C must divide 256, C <= 128.
int ud_bn_stats(const float* x, int n, int C, double* sums, void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !sums || n < 0 || C < 1 || C > 128 || 256 % C != 0) return UD_ERR_INVALID_ARG;
    if (n == 0) {
        UD_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
        return UD_OK;
    }
    const int grid = ud_bn_reduce_grid(n, C);
    if (!workspace || workspace_bytes < ud_det_reduce_bytes(grid, 2 * C)) return UD_ERR_WORKSPACE;
    UD_CUDA_TRY(cudaMemsetAsync(workspace, 0, UD_DET_COUNTER_BYTES, stream));
    k_bn_stats<<<grid, 256, 0, stream>>>(x, n, C, sums, (unsigned char*)workspace, ud_det_group(grid));
    UD_LAUNCH_CHECK();
    return UD_OK;
}

int ud_bn_finalize(const double* sum, const double* sqsum, int n, int C, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* scale, float* shift, float* save_mean, float* save_invstd, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!sum || !sqsum || !scale || !shift || C < 1) return UD_ERR_INVALID_ARG;
    k_bn_finalize<<<ud_blocks(C, 128), 128, 0, stream>>>(sum, sqsum, n, C, gamma, beta, eps, momentum,
                                                         running_mean, running_var, scale, shift, save_mean,
                                                         save_invstd);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

int ud_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma,
                      const float* beta, float eps, int C, float* scale, float* shift, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!running_mean || !running_var || !scale || !shift || C < 1) return UD_ERR_INVALID_ARG;
    k_bn_eval_coeffs<<<ud_blocks(C, 128), 128, 0, stream>>>(running_mean, running_var, gamma, beta, eps, C,
                                                            scale, shift);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// y may alias x.  residual may be NULL.  C % 4 == 0.
int ud_bn_apply(const float* x, int n, int C, const float* scale, const float* shift, const float* residual,
                int relu, float* y, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!x || !y || !scale || !shift || n < 0 || C < 4 || C % 4 != 0) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    long long n4 = (long long)n * C / 4;
    long long want = (n4 + 255) / 256;
    int grid = (int)(want < UD_NUM_SMS * 16 ? want : UD_NUM_SMS * 16);
    k_bn_apply<<<grid, 256, 0, stream>>>((const float4*)x, n4, C / 4, (const float4*)scale,
                                         (const float4*)shift, (const float4*)residual, relu, (float4*)y);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// sums: (2, C) f64, fully written: [0..C) = sum_g (d beta), [C..2C) = sum_gx (d gamma), fixed summation order;
// dx: (n,C); dres nullable (n,C); workspace: ud_bn_workspace_bytes(n, C).
int ud_bn_backward(const float* dy, const float* y, const float* x, int n, int C, const float* mean,
                   const float* invstd, const float* gamma, int relu, double* sums, float* dx, float* dres,
                   void* workspace, size_t workspace_bytes, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!dy || !x || !mean || !invstd || !sums || !dx || n < 0 || C < 1 || C > 128 || 256 % C != 0)
        return UD_ERR_INVALID_ARG;
    if (relu && !y) return UD_ERR_INVALID_ARG;
    if (n == 0) {
        UD_CUDA_TRY(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * C, stream));
        return UD_OK;
    }
    const int grid = ud_bn_reduce_grid(n, C);
    if (!workspace || workspace_bytes < ud_det_reduce_bytes(grid, 2 * C)) return UD_ERR_WORKSPACE;
    UD_CUDA_TRY(cudaMemsetAsync(workspace, 0, UD_DET_COUNTER_BYTES, stream));
    double* sum_g = sums;
    double* sum_gx = sums + C;
    long long total = (long long)n * C;
    const bool vec = C % 4 == 0 && C >= 4 &&
                     ((((uintptr_t)dy | (uintptr_t)y | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)mean |
                        (uintptr_t)invstd) & 15u) == 0);
    if (vec) {
        // same grid as the scalar form (the workspace was sized for it); a block covers 4x the rows per iteration
        k_bn_bwd_reduce4<<<grid, 256, 0, stream>>>((const float4*)dy, (const float4*)y, (const float4*)x, n, C / 4,
                                                   (const float4*)mean, (const float4*)invstd, relu, sums,
                                                   (unsigned char*)workspace, ud_det_group(grid));
        UD_LAUNCH_CHECK();
        const long long total4 = total / 4;
        const long long want4 = (total4 + 255) / 256;
        const int grid4 = (int)(want4 < UD_NUM_SMS * 8 ? want4 : UD_NUM_SMS * 8);
        k_bn_bwd_apply4<<<grid4, 256, 0, stream>>>((const float4*)dy, (const float4*)y, (const float4*)x, total4, C / 4, n,
                                                   (const float4*)mean, (const float4*)invstd, gamma, sum_g, sum_gx, relu,
                                                   (float4*)dx, (float4*)dres);
        UD_LAUNCH_CHECK();
        return UD_OK;
    }
    k_bn_bwd_reduce<<<grid, 256, 0, stream>>>(dy, y, x, n, C, mean, invstd, relu, sums, (unsigned char*)workspace,
                                              ud_det_group(grid));
    UD_LAUNCH_CHECK();
    long long want2 = (total + 255) / 256;
    int grid2 = (int)(want2 < UD_NUM_SMS * 16 ? want2 : UD_NUM_SMS * 16);
    k_bn_bwd_apply<<<grid2, 256, 0, stream>>>(dy, y, x, total, C, n, mean, invstd, gamma, sum_g, sum_gx, relu,
                                              dx, dres);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// feat: (n, C); coords: (n,4) [b,z,y,x]; out: (B, C*D, H, W) f32.  zero_fill != 0 clears `out` first.
int ud_scatter_to_bev(const float* feat, const int* coords, int n, int batch_size, int C, int D, int H,
                      int W, int zero_fill, float* out, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (!feat || !coords || !out || n < 0 || C < 1 || C > 1024) return UD_ERR_INVALID_ARG;
    if (zero_fill)
        UD_CUDA_TRY(cudaMemsetAsync(out, 0, (size_t)batch_size * C * D * H * W * sizeof(float), stream));
    if (n == 0) return UD_OK;
    if (C % 32 != 0 || (((uintptr_t)feat) & 15u) != 0) {           // no [32 x 32] TMA boxes: thread-staged tile
        size_t smem_g = (size_t)32 * (C + 1) * sizeof(float);
        k_scatter_bev_generic<<<ud_blocks(n, 32), 256, smem_g, stream>>>(feat, (const int4*)coords, n, C, D, H, W, out);
        UD_LAUNCH_CHECK();
        return UD_OK;
    }
    CUtensorMap feat_map;
    int rc = ud_make_tmap_2d_b32(&feat_map, feat, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (uint64_t)n, (uint64_t)C,
                                 (uint64_t)C * sizeof(float), 32, 32, true);
    if (rc != UD_OK) return rc;
    size_t smem = (size_t)32 * C * sizeof(float) + 1024;
    k_scatter_bev<<<ud_blocks(n, 32), 256, smem, stream>>>(feat_map, (const int4*)coords, n, C, D, H, W, out);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

// backward of the scatter: feat_grad[o, c] = dense_grad[b, c*D + z, y, x]
int ud_gather_from_bev(const float* dense, const int* coords, int n, int batch_size, int C, int D, int H,
                       int W, float* feat, void* stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)batch_size;
    if (!dense || !coords || !feat || n < 0 || C < 1 || C > 1024) return UD_ERR_INVALID_ARG;
    if (n == 0) return UD_OK;
    size_t smem = (size_t)32 * (C + 1) * sizeof(float);
    k_gather_bev<<<ud_blocks(n, 32), 256, smem, stream>>>(dense, (const int4*)coords, n, C, D, H, W, feat);
    UD_LAUNCH_CHECK();
    return UD_OK;
}

}  // extern "C"
