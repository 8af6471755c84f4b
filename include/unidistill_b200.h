/*
 * unidistill_b200.h -- C ABI of the Blackwell (sm_100a) BEV feature-extraction + distillation path.
 *
 * Every entry point is `extern "C"`, takes plain device pointers, sizes and a CUDA stream
 * (passed as void* == cudaStream_t) and returns 0 on success, a cudaError_t value when a CUDA call
 * failed, or one of the UD_ERR_* codes below.  Nothing synchronises the stream, nothing allocates:
 * scratch comes from the caller (`*_workspace_bytes` tells how much).  "HOST" marks the few small
 * arrays that are read on the host at call time.
 *
 * The reference (megvii-research/CVPR2023-UniDistill) exposes no C ABI of its own for this path; its
 * native arithmetic lives in the spconv wheel and in the prebuilt voxel_pooling_ext .so.  Each
 * declaration cites the reference call site (file:line under /root/reference) whose native callee it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference side.
 */
#ifndef UNIDISTILL_B200_H
#define UNIDISTILL_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UD_OK 0
#define UD_ERR_INVALID_ARG 1001
#define UD_ERR_WORKSPACE 1002
#define UD_ERR_UNSUPPORTED 1003

/* library / build identification ("unidistill_b200 sm_100a <date>") */
const char* ud_version(void);
/* human-readable text for a return code of this library (static storage) */
const char* ud_error_string(int code);
/* kernels launched by this library since the last call with reset != 0 (launch accounting for bench.py) */
int ud_launch_count(int reset);

/* ---------------------------------------------------------------------------------------------
 * Input side: sweep merge + x/y range filter + ragged collate for the whole batch.  Replaces
 *   CollectLidarSweeps.forward (unidistill/data/multisensorfusion/transforms3d.py:379-414),
 *   ObjectRangeFilter.mask_points_by_range (transforms3d.py:247-256,271-276) and the zero-padding points part of
 *   collate_fn (data/multisensorfusion/nuscenes_multimodal.py:418-470).
 * points (n,C) f32: every (sample, sweep) segment concatenated, sample-major, a sample's key frame first; seg_offsets
 * (n_seg+1) / sample_first_seg (B+1) DEVICE int32; mats (n_seg,3,4) DEVICE f64 sweep->key-frame transforms (identity for
 * key frames), applied in float64 and rounded to float32 once like the numpy code; time_lag (n_seg) DEVICE f32 -> 5th
 * column; range_xy HOST float[4] (x_min,y_min,x_max,y_max), inclusive.  out (n,C) f32 holds the kept points in input order
 * (ragged batch: out_offsets (B+1) DEVICE int32, out_offsets[B] = total) -- the layout ud_hard_voxelize takes.
 * ------------------------------------------------------------------------------------------- */
int ud_collect_sweeps_workspace_bytes(int n_points, size_t* bytes);
int ud_collect_sweeps(const float* points, int n, int C, const int* seg_offsets, int n_seg, const int* sample_first_seg,
                      int batch_size, const double* mats, const float* time_lag, const float* range_xy, float* out,
                      int* out_offsets, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hard voxelization + mean VFE.
 * Replaces spconv.pytorch.utils.PointToVoxel.__call__ as driven by
 *   unidistill/data/det3d/preprocess/voxelization.py:31-38,53-68   (per-sample loop + clones + cat)
 * and the reduction of unidistill/layers/blocks_3d/det3d/vfe/mean_vfe.py:25-32.
 * points (N_total,C) f32: the batch's clouds concatenated; sample_offsets HOST int[batch_size+1].
 * voxel_size HOST float[3] (x,y,z); pc_range HOST float[6]; grid_size HOST int[3] (x,y,z).
 * Outputs (row capacity out_capacity): voxels (cap,max_points,C) f32 or NULL; coords (cap,4) int32
 * [b,z,y,x]; num_points (cap) int32; mean (cap,C) f32 or NULL; counts (1+batch_size) int32 device =
 * [total rows, rows of sample 0, ...].  Deterministic first-come order (see oracle/voxelize_ref.c).
 * ------------------------------------------------------------------------------------------- */
int ud_voxelize_workspace_bytes(int num_points_total, size_t* bytes);
int ud_hard_voxelize(const float* points, const int* sample_offsets, int batch_size, int num_features,
                     const float* voxel_size, const float* pc_range, const int* grid_size,
                     int max_points, int max_voxels, int out_capacity, float* voxels, int* coords,
                     int* num_points, float* mean, int* counts, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Rulebooks.  Replace spconv's get_indice_pairs, implicit in every SubMConv3d/SparseConv3d of
 *   unidistill/layers/blocks_3d/det3d/spconv_backbone.py:259-340
 * (operator API visible at layers/blocks_3d/det3d/centernet_backbone.py:221-251).
 * coords (n,4) int32 [b,z,y,x]; shapes/ksize/stride/padding/dilation are HOST int[3] in (z,y,x).
 * Tables are int32 with -1 == absent, K = kz*ky*kx (<= 32), tap index k = (kz*KY+ky)*KX+kx:
 *   nbr_out (n_out,K): input row feeding output row o through tap k
 *   nbr_in  (n_in, K): output row fed by input row i through tap k
 * Strided conv outputs are ordered ascending by a key over the output cells, chosen by `order`:
 *   0 = linear (b,z,y,x) index (canonical; the order the rulebook parity tests pin),
 *   1 = (b, Morton(z,y,x)): 128 consecutive output rows form a compact 3-D block, which is what the convolution kernels'
 *       gathers want.  The row order of a sparse tensor is not observable after SparseConvTensor.dense().
 * ------------------------------------------------------------------------------------------- */
int ud_rulebook_subm_workspace_bytes(int n, size_t* bytes);
/* Every builder can emit, besides / instead of the row-major (n,K) table, the tap-major form the tensor-core kernels read:
 * nbr_*_t (K4, ld) with K4 = K rounded up to 4 and ld = n rounded up to 128 (padding -1), plus tile masks (ld/128) uint32
 * with bit k set when tap k occurs in the 128-row tile (same layout as ud_rulebook_transpose produces).  Any of the table
 * outputs may be NULL (at least one per side); a tap-major table needs its mask array. */
int ud_rulebook_subm(const int* coords, int n, int batch_size, const int* spatial_shape, const int* ksize,
                     const int* dilation, int* nbr_out, int* nbr_out_t, unsigned int* tile_masks, void* workspace,
                     size_t workspace_bytes, void* stream);
int ud_rulebook_conv_workspace_bytes(int batch_size, const int* out_shape, int order, size_t* bytes);
/* phase 1: n_out_dev (device int) <- number of output rows; keep the workspace for phase 2 */
int ud_rulebook_conv_count(const int* coords, int n_in, int batch_size, const int* in_shape,
                           const int* out_shape, const int* ksize, const int* stride, const int* padding,
                           const int* dilation, int order, int* n_out_dev, void* workspace, size_t workspace_bytes,
                           void* stream);
/* phase 2: out_coords (n_out,4), nbr_in (n_in,K), nbr_out (n_out,K) */
int ud_rulebook_conv_fill(const int* coords, int n_in, int batch_size, const int* in_shape,
                          const int* out_shape, const int* ksize, const int* stride, const int* padding,
                          const int* dilation, int order, int n_out, int* out_coords, int* nbr_in, int* nbr_out,
                          int* nbr_in_t, unsigned int* masks_in, int* nbr_out_t, unsigned int* masks_out,
                          void* workspace, size_t workspace_bytes, void* stream);
/* order (n) int32 <- row indices of coords (n,4) sorted by (b, Morton(z,y,x)) on the (z,y,x) grid spatial_shape (HOST
 * int[3]); stable.  Used once at the entry of the sparse encoder (the voxelizer emits first-come order). */
int ud_morton_order_workspace_bytes(int n, size_t* bytes);
int ud_morton_order(const int* coords, int n, int batch_size, const int* spatial_shape, int* order, void* workspace,
                    size_t workspace_bytes, void* stream);
/* spconv-shaped export: pairs (2,K,n_in) int32 (-1 padded), pair_num (K); mirror=1 reads a SubM nbr_out */
int ud_rulebook_pairs_workspace_bytes(int n_in, int K, size_t* bytes);
int ud_rulebook_pairs(const int* nbr_in, int n_in, int K, int mirror, int* pairs, int* pair_num,
                      void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Sparse convolution.  Replaces spconv.pytorch.functional.indice_conv / indice_subm_conv
 * (ConvAlgo.Native) behind spconv_backbone.py:21-48,70-92 and their backward.
 *   forward: out[o,:] = bias + sum_k in[nbr[o][k],:] @ w[k]     w (K,cin,cout) f32, bias (cout)|NULL
 *   dgrad  : the same call with in:=dout, nbr:=nbr_in, w:=W^T as (K,cout,cin) (SubM: nbr_out and the
 *            taps reversed), cin/cout swapped
 *   wgrad  : dw[k] += in[nbr[o][k],:]^T dout[o,:]                dw (K,cin,cout), caller zeroes it
 * ------------------------------------------------------------------------------------------- */
int ud_sparse_conv_forward(const float* in, const float* w, const float* bias, const int* nbr, int n_out,
                           int K, int cin, int cout, float* out, void* stream);
int ud_sparse_conv_wgrad(const float* in, const float* dout, const int* nbr, int n_out, int K, int cin,
                         int cout, float* dw, void* stream);

/* Tensor-core kernel (tcgen05.mma kind::f16 on a bf16 hi/lo split of both operands == fp32-accurate to ~2e-5, fp32
 * accumulators in TMEM, gathered rows global -> registers -> TMEM; csrc/sparse_conv_umma.cu).  Same arithmetic contract as
 * ud_sparse_conv_forward for the (cin, cout) pairs ud_sparse_conv_tc_supported() accepts; the weights are first packed
 * into the swizzled bf16 hi/lo image the kernel's bulk copies read (ud_sparse_conv_tc_weight_bytes). */
int ud_sparse_conv_tc_supported(int K, int cin, int cout);
int ud_sparse_conv_tc_weight_bytes(int K, int cin, int cout, size_t* bytes);
int ud_sparse_conv_tc_pack_weights(const float* w, int K, int cin, int cout, void* packed, void* stream);
/* The same image straight from the spconv-2 parameter layout (Cout_p, kz, ky, kx, Cin_p) of spconv.SubMConv3d /
 * SparseConv3d.weight (spconv_backbone.py:259-340), for the convolution that will run with (cin, cout):
 * mode 0 forward (cin = Cin_p, cout = Cout_p); mode 1 data gradient of a submanifold convolution (cin = Cout_p,
 * cout = Cin_p: transposed weights, mirrored taps); mode 2 data gradient of a strided convolution (transposed). */
int ud_sparse_conv_tc_pack_weights_param(const float* w_param, int K, int cin, int cout, int mode, void* packed, void* stream);
/* nbr (n,K) row-major gather table -> nbr_t (K4,ld) tap-major with K4 = K rounded up to 4, ld = n rounded up to 128
 * (padding -1; 16-byte aligned base) + tile_masks (ceil(n/128)) uint32 with bit k set when tap k occurs in that 128-row
 * tile: the two table forms the tensor-core kernels read */
int ud_rulebook_transpose(const int* nbr, int n, int K, int* nbr_t, unsigned int* tile_masks, void* stream);
/* bn_stats (nullable): scratch of ud_sparse_conv_tc_stats_bytes(n_out, cout) bytes, 16-byte aligned, uninitialised.  On
 * return its first 2*cout doubles are the per-channel sum and sum of squares of `out`, i.e. the statistics pass of the
 * BatchNorm1d that follows the convolution (spconv_backbone.py:50-56,97-113), reduced in a fixed order (bit-reproducible) */
int ud_sparse_conv_tc_stats_bytes(int n_out, int cout, size_t* bytes);
/* in_is_split: 0 = `in` is the fp32 feature matrix; 1 = `in` is its split image from ud_split_bf16 (same size: every 4
 * channels stored as {bf16x2 hi01, hi23, lo01, lo23}), which takes the fp32 -> bf16 hi/lo conversion out of the gather loop:
 * a row is gathered ~16 times (once per tap that uses it) but has to be split only once. */
int ud_split_bf16(const float* x, int n, int C, void* out, void* stream);
int ud_sparse_conv_tc_forward(const float* in, int in_is_split, const void* packed, const float* bias, const int* nbr_t,
                              const unsigned int* tile_masks, int n_out, int K, int cin, int cout, float* out,
                              void* bn_stats, void* stream);
/* Inference form of ud_sparse_conv_tc_forward: out = relu?((conv(in) + bias[c]) * scale[c] + shift[c] (+ residual[o][c])),
 * i.e. the eval-mode BatchNorm1d (scale / shift from ud_bn_eval_coeffs), the residual add of SparseBasicBlock and the ReLU
 * that follow the convolution (spconv_backbone.py:50-56,97-113), applied in its epilogue with the arithmetic of
 * ud_bn_apply (bit-identical to ud_sparse_conv_tc_forward -> ud_bn_apply).  bias and residual nullable. */
int ud_sparse_conv_tc_forward_affine(const float* in, const void* packed, const float* bias, const float* scale, const float* shift,
                                     const float* residual, int relu, const int* nbr_t, const unsigned int* tile_masks,
                                     int n_out, int K, int cin, int cout, float* out, void* stream);
/* Window-staged variant for 3x3x3 kernels with 64 input channels (raster row order): TMA stages, once per (tile, kz, ky),
 * the 160-row window of the feature matrix that holds the rows the three kx taps gather (97 % of them on the bench
 * clouds) and the operands are built from shared memory; rows outside the window come from global memory, so any table is
 * handled correctly.  win_lo (ceil(n/128), 9) int32 from ud_rulebook_windows (first row of every window); n_in = rows of `in`. */
int ud_sparse_conv_tc_window_supported(int K, int cin, int cout);
int ud_rulebook_windows(const int* nbr_t, int n, int K, int* win_lo, void* stream);
int ud_sparse_conv_tc_forward_window(const float* in, int n_in, const void* packed, const float* bias, const int* nbr_t,
                                     const unsigned int* tile_masks, const int* win_lo, int n_out, int K, int cin,
                                     int cout, float* out, void* bn_stats, void* stream);

/* Tensor-core weight gradient (csrc/sparse_conv_wgrad_umma.cu): dw[k] += in[nbr[o][k],:]^T dout[o,:] as tcgen05 MMAs
 * over the row dimension (both operands MN-major in shared memory, rounded to TF32 to-nearest on the way in, fp32
 * accumulators for up to 512/cout kernel taps in TMEM).  nbr_t / tile_masks: ud_rulebook_transpose of the same (n_rows,K)
 * table ud_sparse_conv_wgrad takes; dw (K,cin,cout) is accumulated into (caller zeroes it).  Replaces the wgrad half of
 * spconv's indice_conv backward (spconv_backbone.py:21-48,70-92 under autograd). */
int ud_sparse_conv_tc_wgrad_supported(int K, int cin, int cout);
int ud_sparse_conv_tc_wgrad(const float* in, const float* dout, const int* nbr_t, const unsigned int* tile_masks,
                            int n_rows, int K, int cin, int cout, float* dw, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm1d over active rows (+residual, +ReLU).  Replaces the nn.BatchNorm1d / nn.ReLU /
 * replace_feature passes of spconv_backbone.py:50-56,97-113,255.
 * ------------------------------------------------------------------------------------------- */
/* ud_bn_stats / ud_bn_backward reduce over the rows in a FIXED order (no floating-point atomics): bit-reproducible.
 * sums (2,C) f64 is fully written: statistics pass [sum | sum of squares]; backward [sum g | sum g*xhat] = [d beta | d gamma].
 * workspace: ud_bn_workspace_bytes(n, C).  C divides 256, C <= 128. */
int ud_bn_workspace_bytes(int n, int C, size_t* bytes);
int ud_bn_stats(const float* x, int n, int C, double* sums, void* workspace, size_t workspace_bytes, void* stream);
int ud_bn_finalize(const double* sum, const double* sqsum, int n, int C, const float* gamma,
                   const float* beta, float eps, float momentum, float* running_mean, float* running_var,
                   float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);
int ud_bn_eval_coeffs(const float* running_mean, const float* running_var, const float* gamma,
                      const float* beta, float eps, int C, float* scale, float* shift, void* stream);
int ud_bn_apply(const float* x, int n, int C, const float* scale, const float* shift, const float* residual,
                int relu, float* y, void* stream);
int ud_bn_backward(const float* dy, const float* y, const float* x, int n, int C, const float* mean,
                   const float* invstd, const float* gamma, int relu, double* sums, float* dx, float* dres,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Scatter to BEV.  Replaces SparseConvTensor.dense() + view of
 *   unidistill/layers/blocks_2d/det3d/map_to_bev/height_compression.py:19-21
 * out (B, C*D, H, W) f32, channel = c*D + z.  ud_gather_from_bev is its backward.
 * ------------------------------------------------------------------------------------------- */
int ud_scatter_to_bev(const float* feat, const int* coords, int n, int batch_size, int C, int D, int H,
                      int W, int zero_fill, float* out, void* stream);
int ud_gather_from_bev(const float* dense, const int* coords, int n, int batch_size, int C, int D, int H,
                       int W, float* feat, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Camera branch: frustum -> BEV pooling.  Replaces the prebuilt voxel_pooling_ext extension
 * (voxel_pooling_forward_wrapper -> voxel_pooling_forward_kernel_launcher, binary only) called from
 *   unidistill/layers/blocks_3d/mmdet3d/lss_fpn.py:48-59 (forward) and restates :64-79 (backward).
 * Argument order of ud_voxel_pooling_forward == the reference launcher's
 *   (batch_size, num_points, num_channels, num_voxel_x, num_voxel_y, num_voxel_z, geom_xyz,
 *    input_features, output_features, pos_memo, stream) + workspace.
 * geom_xyz (B,N,3) int32; input_features (B,N,C) f32; output_features (B,ny,nx,C) f32 and pos_memo
 * (B,N,3) int32 are fully written (no pre-zero / pre-fill needed); pos_memo may be NULL.
 * Deterministic (sorted segmented reduction, no atomics).
 * ------------------------------------------------------------------------------------------- */
int ud_voxel_pooling_workspace_bytes(int batch_size, int num_points, int num_channels, int num_voxel_x,
                                     int num_voxel_y, size_t* bytes);
int ud_voxel_pooling_forward(int batch_size, int num_points, int num_channels, int num_voxel_x, int num_voxel_y,
                             int num_voxel_z, const int* geom_xyz, const float* input_features,
                             float* output_features, int* pos_memo, void* workspace, size_t workspace_bytes,
                             void* stream);
/* grad_output (B,C,ny,nx) contiguous -> grad_input (B,N,C), fully written */
int ud_voxel_pooling_backward(int batch_size, int num_points, int num_channels, int num_voxel_x, int num_voxel_y,
                              const float* grad_output, const int* pos_memo, float* grad_input, void* workspace,
                              size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LSS frustum geometry -> BEV voxel index in one kernel.  Replaces LSSFPN.get_geometry (three batched 4x4 mat-vec
 * products over the (B,ncam,D,H,W) frustum, lss_fpn.py:200-240) and the index computation
 * `((geom - (voxel_coord - voxel_size / 2)) / voxel_size).int()` (lss_fpn.py:311-313; truncation toward zero).
 * frustum (D*H*W,4) f32 [u,v,depth,1]; ida_inv = ida_mat^-1 and cam2ego = sensor2ego @ intrin^-1, both (B*ncam,4,4) f32
 * row-major (computed by the caller: B*ncam tiny matrices); bda (B,4,4) f32 or NULL; voxel_lo / voxel_size HOST float[3].
 * idx (B,ncam,D*H*W,3) int32 and/or geom (same shape, f32); either output may be NULL.
 * ------------------------------------------------------------------------------------------- */
int ud_lss_geometry(const float* frustum, int num_frustum_points, int batch_size, int num_cams, const float* ida_inv,
                    const float* cam2ego, const float* bda, const float* voxel_lo, const float* voxel_size, int* idx,
                    float* geom, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused LSS lift-splat: out[b,:,y,x] = sum over frustum points p in cell (y,x) of depth[p] * ctx[pix(p),:].
 * Replaces the materialised outer product + permute + pooling of lss_fpn.py:289-316 (484 MB/sample x2).
 * geom_xyz (B, ncam*D*H*W, 3) int32 (lss_fpn.py:311-313); depth (B*ncam,D,H,W) f32 (already softmaxed);
 * ctx (B*ncam,C,H,W) f32; out (B,C,ny,nx) f32.  C in {128,256}.  `state` (state_bytes from
 * ud_lift_splat_sizes) carries what the backward needs.
 * ------------------------------------------------------------------------------------------- */
int ud_lift_splat_sizes(int batch_size, int num_cams, int D, int H, int W, int C, int nx, int ny,
                        size_t* workspace_bytes, size_t* state_bytes);
int ud_lift_splat_forward(int batch_size, int num_cams, int D, int H, int W, int C, int nx, int ny, int nz,
                          const int* geom_xyz, const float* depth, const float* ctx, float* out, void* state,
                          size_t state_bytes, void* workspace, size_t workspace_bytes, void* stream);
int ud_lift_splat_backward(int batch_size, int num_cams, int D, int H, int W, int C, int nx, int ny,
                           const float* grad_out, const float* depth, const void* state, size_t state_bytes,
                           float* d_depth, float* d_ctx, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused distillation losses.  Replace the module-level functions of the reference's distill scripts
 *   exps/multisensor_fusion/nuscenes/BEVFusion/BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar.py
 *     :73-97 center_to_corner_box2d (+ :478-483 cell units)   -> ud_box_corners_bev
 *     :196-245 FeatureDistillLoss (mode 0), :248-323 BEVDistillLoss (mode 1) -> ud_box_distill_*
 *     :326-385 ResponseDistillLoss + :100-128 calculate_box_mask_gaussian    -> ud_response_distill_*
 * (identical in the three sibling scripts; camera_exp_distill_fusion uses clamp_lo = 1e-3).
 * Maps are (B,C,H,W) f32; corners (B,M,4,2) f32 in BEV cells; valid (B,M) u8.  Per-box / per-block partial
 * results are returned so the caller applies the (all-reduced) normalisers; backward entry points give
 * gradients w.r.t. the student tensors only; `upstream` of both backward entry points is a DEVICE pointer (per-box
 * upstream gradients / [dL/dsum_cls, dL/dsum_reg]) so that a backward pass never reads scalars back to the host.
 * ------------------------------------------------------------------------------------------- */
int ud_box_corners_bev(const float* boxes, int n, int stride, float x_min, float y_min, float cell_x,
                       float cell_y, float* corners, void* stream);
int ud_box_distill_forward(int mode, const float* s_map, const float* t_map, int B, int C, int H, int W,
                           const float* corners, const unsigned char* valid, int M, float* out, void* stream);
int ud_box_distill_backward(int mode, const float* s_map, const float* t_map, int B, int C, int H, int W,
                            const float* corners, const unsigned char* valid, int M, const float* upstream,
                            float* grad_s, void* stream);
/* The same two with the memory layout of each map stated (s_nhwc / t_nhwc: 0 = (B,C,H,W) contiguous, 1 = torch
 * channels_last, i.e. NHWC -- the layout the dense trunk computes in; grad_s has the layout of s_map).  The reference
 * samples these maps with F.grid_sample on NCHW tensors (..._distill_lidar.py:200-226); reading NHWC in place saves the
 * four layout copies per step that fed the NCHW entry points. */
int ud_box_distill_forward_layout(int mode, const float* s_map, int s_nhwc, const float* t_map, int t_nhwc, int B, int C,
                                  int H, int W, const float* corners, const unsigned char* valid, int M, float* out,
                                  void* stream);
int ud_box_distill_backward_layout(int mode, const float* s_map, int s_nhwc, const float* t_map, int t_nhwc, int B, int C,
                                   int H, int W, const float* corners, const unsigned char* valid, int M,
                                   const float* upstream, float* grad_s, void* stream);
int ud_response_distill_forward(const float* cls_s, const float* cls_t, const float* reg_s, const float* reg_t,
                                int B, int n_cls, int n_reg, int H, int W, const float* boxes, int M, int stride,
                                const int* n_rows, float x_min, float y_min, float cell_x, float cell_y,
                                float clamp_lo, float* partial, float* mask_out, void* stream);
int ud_response_distill_backward(const float* cls_s, const float* cls_t, const float* reg_s, const float* reg_t,
                                 int B, int n_cls, int n_reg, int H, int W, const float* boxes, int M, int stride,
                                 const int* n_rows, float x_min, float y_min, float cell_x, float cell_y,
                                 float clamp_lo, const float* upstream, float* d_cls, float* d_reg,
                                 void* stream);

/* ---------------------------------------------------------------------------------------------
 * Proposal post-processing (inference): rotated-BEV IoU and greedy NMS.  Replaces the binary-only
 * extension unidistill/layers/head/det3d/generate_proposals/iou3d_nms_cuda*.so:
 *     boxes_iou_bev_gpu(boxes_a (N,7), boxes_b (M,7), ans_iou (N,M))          -> ud_boxes_iou_bev
 *     nms_gpu(boxes (N,7) cuda f32 sorted by score, keep int64 CPU, thresh)    -> ud_nms_bev (batched, device-side sweep)
 * called from generate_proposals/centerpoint_gen_proposals.py:85-106.  Boxes are [x,y,z,dx,dy,dz,heading] f32.
 * ud_nms_bev processes n_sets independent box lists of n rows each (one (task, sample) pair per set); only the first
 * n_valid[s] rows (device int32, NULL = n) of set s take part; keep (n_sets,max_keep) int32 = kept row positions in
 * ascending order padded with -1, num_keep (n_sets) int32 (<= max_keep, == selected[:post_max_size]).
 * ------------------------------------------------------------------------------------------- */
int ud_boxes_iou_bev(const float* boxes_a, int na, const float* boxes_b, int nb, float* ans_iou, void* stream);
int ud_nms_bev_workspace_bytes(int n_sets, int n, size_t* bytes);
int ud_nms_bev(const float* boxes, const int* n_valid, int n_sets, int n, float thresh, int max_keep, int* keep,
               int* num_keep, void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FCOS-style target assignment of the CenterPoint head: replaces FCOSAssigner.assign_targets,
 * layers/head/det3d/target_assigner/fcos_assigner.py:73-285, for the whole batch and all tasks in one launch.
 * boxes (B,M,NB) f32 in feature-map units [x,y,z,dx,dy,dz,yaw,(vx,vy)] in the reference's task order; task_id / cls_off
 * (B,M) int32 (task of the row or -1, class index inside the task); anchors (fw*fh,2) f32 (x,y), a = i*fw + j, regular
 * grid of pitch `stride`; cls_prefix: n_tasks+1 HOST ints.  heat: concatenation over tasks of (B,n_cls_t,fh*fw) f32 blocks,
 * zeroed by the caller; ind / cat (n_tasks,B,max_objs) int64, mask (n_tasks,B,max_objs) u8, enc (n_tasks,B,max_objs,NB+1).
 * n_pos (n_tasks,B) int32 or NULL: positives before truncation to max_objs (the reference asserts n_pos < max_objs in
 * `padding_to_maxobjs`; the overflow is reported here instead).  topk <= 9.
 * ------------------------------------------------------------------------------------------- */
int ud_fcos_assign(const float* boxes, const int* task_id, const int* cls_off, const float* anchors, int B, int M, int NB,
                   int fw, int fh, float stride, int n_tasks, const int* cls_prefix, int topk, int max_objs, float vs0,
                   float vs1, float osf, float* heat, long long* ind, unsigned char* mask, long long* cat, float* enc,
                   int* n_pos, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UNIDISTILL_B200_H */
