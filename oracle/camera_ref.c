/*
 * oracle/camera_ref.c -- TEST INFRASTRUCTURE ONLY (checker / reported CPU baseline).
 *
 * Plain-C restatement of the reference's camera-branch pooling kernel.  The reference ships it only as
 * a binary (unidistill/layers/blocks_3d/mmdet3d/voxel_pooling_ext.cpython-36m-x86_64-linux-gnu.so,
 * cubins sm_61..86, torch-1.9 ABI: cannot be loaded here), so its semantics were recovered from the SASS
 * and the Python call site (lss_fpn.py:12-62): PARITY IS UNPINNED BY THE REFERENCE (no tests,
 * no golden vectors); pinned by known-answer tests in tests/test_oracle_camera.py.
 *
 *   for every point p of sample b: (x,y,z) = geom[b,p]; if 0<=x<nx, 0<=y<ny, 0<=z<nz:
 *       pos_memo[b,p] = (b,y,x);  out[b,y,x,:] += feats[b,p,:]          (out pre-zeroed, pos_memo pre -1)
 * The GPU reference accumulates with fp32 atomics in arbitrary order; this restatement adds in point order.
 */
#include <stddef.h>

void ud_ref_voxel_pooling(int B, int N, int C, int nx, int ny, int nz, const int* geom, const float* feats,
                          float* out /* (B,ny,nx,C) zeroed */, int* pos_memo /* (B,N,3) filled with -1 */) {
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < N; ++p) {
            const int* g = geom + ((size_t)b * N + p) * 3;
            int x = g[0], y = g[1], z = g[2];
            if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz) continue;
            int* pm = pos_memo + ((size_t)b * N + p) * 3;
            pm[0] = b; pm[1] = y; pm[2] = x;
            const float* f = feats + ((size_t)b * N + p) * C;
            float* o = out + (((size_t)b * ny + y) * nx + x) * C;
            for (int c = 0; c < C; ++c) o[c] += f[c];
        }
}

/* VoxelPooling.backward, lss_fpn.py:64-79: grad_in[b,p,:] = grad_out[b,:,y,x] for kept points, else 0.
 * grad_out is (B,C,ny,nx). */
void ud_ref_voxel_pooling_backward(int B, int N, int C, int nx, int ny, const float* grad_out,
                                   const int* pos_memo, float* grad_in /* (B,N,C) zeroed */) {
    for (int b = 0; b < B; ++b)
        for (int p = 0; p < N; ++p) {
            const int* pm = pos_memo + ((size_t)b * N + p) * 3;
            if (pm[0] < 0) continue;
            float* gi = grad_in + ((size_t)b * N + p) * C;
            for (int c = 0; c < C; ++c)
                gi[c] = grad_out[(((size_t)pm[0] * C + c) * ny + pm[1]) * nx + pm[2]];
        }
}

/* Fused form used for the CPU baseline at full size (the lifted volume is never stored):
 * out (B,ny,nx,C) zeroed; depth (B*ncam,D,H,W); ctx_nhwc (B*ncam,H,W,C); geom (B, ncam*D*H*W, 3). */
void ud_ref_lift_splat(int B, int ncam, int D, int H, int W, int C, int nx, int ny, int nz, const int* geom,
                       const float* depth, const float* ctx_nhwc, float* out) {
    size_t HW = (size_t)H * W;
    size_t N = (size_t)ncam * D * HW;
    for (int b = 0; b < B; ++b)
        for (int cam = 0; cam < ncam; ++cam)
            for (int d = 0; d < D; ++d)
                for (size_t hw = 0; hw < HW; ++hw) {
                    size_t p = ((size_t)cam * D + d) * HW + hw;
                    const int* g = geom + ((size_t)b * N + p) * 3;
                    int x = g[0], y = g[1], z = g[2];
                    if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz) continue;
                    float dp = depth[(size_t)b * N + p];
                    const float* row = ctx_nhwc + (((size_t)b * ncam + cam) * HW + hw) * C;
                    float* o = out + (((size_t)b * ny + y) * nx + x) * C;
                    for (int c = 0; c < C; ++c) o[c] += dp * row[c];
                }
}
