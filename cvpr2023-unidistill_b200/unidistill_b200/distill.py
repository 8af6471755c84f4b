"""Fused distillation-loss operators (host side of csrc/distill.cu)."""
import torch
from torch.autograd import Function

from . import native


def valid_box_prefix(gt_boxes):
    """(B, M) bool: rows up to the last non-zero row of every sample; row 0 always True.
    Sync-free form of the reference's Python loop (..._distill_lidar.py:449-455, quirk 8)."""
    B, M = gt_boxes.shape[0], gt_boxes.shape[1]
    ar = torch.arange(M, device=gt_boxes.device)
    nonzero = gt_boxes.sum(dim=2) != 0
    last = (nonzero.long() * ar).max(dim=1).values
    return ar.unsqueeze(0) <= last.unsqueeze(1)


def rows_before_first_zero(gt_boxes):
    """(B,) int32: index of the first row whose sum is exactly 0 (else M) -- the `break` of
    calculate_box_mask_gaussian (..._distill_lidar.py:109-111)."""
    B, M = gt_boxes.shape[0], gt_boxes.shape[1]
    zero = gt_boxes.sum(dim=2) == 0
    ar = torch.arange(M, device=gt_boxes.device).unsqueeze(0).expand(B, M)
    first = torch.where(zero, ar, torch.full_like(ar, M)).min(dim=1).values
    return first.int().contiguous()


def box_corners_bev(gt_boxes, pc_range, voxel_size, out_size_factor):
    """GT boxes (B, M, >=7) -> BEV corners (B, M, 4, 2) in feature-map cells, on the device
    (replaces the per-sample CPU numpy hop of ..._distill_lidar.py:466-483)."""
    native.require_cuda(gt_boxes)
    lib = native.load()
    g = gt_boxes.contiguous().float()
    B, M, S = g.shape
    out = torch.empty((B, M, 4, 2), dtype=torch.float32, device=g.device)
    rc = lib.ud_box_corners_bev(native.ptr(g), B * M, S, float(pc_range[0]), float(pc_range[1]),
                                float(voxel_size[0] * out_size_factor), float(voxel_size[1] * out_size_factor),
                                native.ptr(out), native.stream_ptr(g.device))
    native.check(rc, "ud_box_corners_bev")
    return out


def _map_layout(x):
    """(tensor the kernels can read in place, nhwc flag): a dense channels_last (NHWC) map is taken as it is -- the dense
    trunk computes in that layout -- anything else is made (B, C, H, W) contiguous."""
    x = x.float()
    if x.dim() == 4 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        return x, 1
    return x.contiguous(), 0


class BoxDistill(Function):
    """Per-box feature-L1 (mode 0) or relation/Gram-L1 (mode 1) terms: (B*M,) with 0 for invalid boxes."""

    @staticmethod
    def forward(ctx, student, teacher, corners, valid, mode):
        native.require_cuda(student, teacher, corners)
        lib = native.load()
        s, s_nhwc = _map_layout(student)
        t, t_nhwc = _map_layout(teacher.detach())
        c = corners.detach().contiguous().float()
        v = valid.to(torch.uint8).contiguous()
        B, C, H, W = [int(x) for x in s.shape]
        M = int(c.shape[1])
        out = torch.empty((B * M,), dtype=torch.float32, device=s.device)
        rc = lib.ud_box_distill_forward_layout(int(mode), native.ptr(s), s_nhwc, native.ptr(t), t_nhwc, B, C, H, W,
                                               native.ptr(c), native.ptr(v), M, native.ptr(out),
                                               native.stream_ptr(s.device))
        native.check(rc, "ud_box_distill_forward_layout")
        ctx.save_for_backward(s, t, c, v)
        ctx.mode = int(mode)
        ctx.layout = (s_nhwc, t_nhwc)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        s, t, c, v = ctx.saved_tensors
        lib = native.load()
        B, C, H, W = [int(x) for x in s.shape]
        M = int(c.shape[1])
        g = grad_out.contiguous().float()
        grad_s = torch.empty_like(s)          # preserve_format: the gradient has the student's layout
        rc = lib.ud_box_distill_backward_layout(ctx.mode, native.ptr(s), ctx.layout[0], native.ptr(t), ctx.layout[1],
                                                B, C, H, W, native.ptr(c), native.ptr(v), M, native.ptr(g),
                                                native.ptr(grad_s), native.stream_ptr(s.device))
        native.check(rc, "ud_box_distill_backward_layout")
        return grad_s, None, None, None, None


class ResponseDistill(Function):
    """-> (sum |max_c cls_s - max_c cls_t'| * mask, sum mean_c |reg_s - reg_t| * mask, sum mask) as a (3,) tensor,
    cls_t' = clamp(sigmoid(cls_t / 2)); Gaussian mask rasterised on the device."""

    @staticmethod
    def forward(ctx, cls_s, cls_t, reg_s, reg_t, gt_boxes, n_rows, geom):
        native.require_cuda(cls_s, cls_t, reg_s, reg_t, gt_boxes)
        lib = native.load()
        cs, ct = cls_s.contiguous().float(), cls_t.detach().contiguous().float()
        rs, rt = reg_s.contiguous().float(), reg_t.detach().contiguous().float()
        gt = gt_boxes.detach().contiguous().float()
        B, n_cls, H, W = [int(x) for x in cs.shape]
        n_reg = int(rs.shape[1])
        M, S = int(gt.shape[1]), int(gt.shape[2])
        x0, y0, cell_x, cell_y, clamp_lo = geom
        n_blocks = B * ((H * W + 255) // 256)
        partial = torch.empty((n_blocks, 3), dtype=torch.float32, device=cs.device)
        rc = lib.ud_response_distill_forward(native.ptr(cs), native.ptr(ct), native.ptr(rs), native.ptr(rt), B, n_cls,
                                             n_reg, H, W, native.ptr(gt), M, S, native.ptr(n_rows), float(x0),
                                             float(y0), float(cell_x), float(cell_y), float(clamp_lo),
                                             native.ptr(partial), None, native.stream_ptr(cs.device))
        native.check(rc, "ud_response_distill_forward")
        ctx.save_for_backward(cs, ct, rs, rt, gt, n_rows)
        ctx.geom = geom
        return partial.sum(dim=0)

    @staticmethod
    def backward(ctx, grad):
        cs, ct, rs, rt, gt, n_rows = ctx.saved_tensors
        lib = native.load()
        B, n_cls, H, W = [int(x) for x in cs.shape]
        n_reg = int(rs.shape[1])
        M, S = int(gt.shape[1]), int(gt.shape[2])
        x0, y0, cell_x, cell_y, clamp_lo = ctx.geom
        g = grad.detach().float().contiguous()      # (3,) on the device: [dL/dsum_cls, dL/dsum_reg, dL/dsum_mask]
        d_cls = torch.empty_like(cs)
        d_reg = torch.empty_like(rs)
        rc = lib.ud_response_distill_backward(native.ptr(cs), native.ptr(ct), native.ptr(rs), native.ptr(rt), B,
                                              n_cls, n_reg, H, W, native.ptr(gt), M, S, native.ptr(n_rows),
                                              float(x0), float(y0), float(cell_x), float(cell_y), float(clamp_lo),
                                              native.ptr(g), native.ptr(d_cls), native.ptr(d_reg),
                                              native.stream_ptr(cs.device))
        native.check(rc, "ud_response_distill_backward")
        return d_cls, None, d_reg, None, None, None, None


def response_mask(gt_boxes, shape_hw, pc_range, voxel_size, out_size_scale):
    """Gaussian box mask (B, H, W) as calculate_box_mask_gaussian would build it (diagnostic / tests)."""
    native.require_cuda(gt_boxes)
    lib = native.load()
    gt = gt_boxes.detach().contiguous().float()
    B, M, S = gt.shape
    H, W = shape_hw
    dev = gt.device
    z1 = torch.zeros((B, 1, H, W), device=dev)
    n_rows = rows_before_first_zero(gt)
    mask = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    partial = torch.empty((B * ((H * W + 255) // 256), 3), dtype=torch.float32, device=dev)
    rc = lib.ud_response_distill_forward(native.ptr(z1), native.ptr(z1), native.ptr(z1), native.ptr(z1), B, 1, 1, H,
                                         W, native.ptr(gt), M, S, native.ptr(n_rows), float(pc_range[0]),
                                         float(pc_range[1]), float(voxel_size[0] * out_size_scale),
                                         float(voxel_size[1] * out_size_scale), 1e-4, native.ptr(partial),
                                         native.ptr(mask), native.stream_ptr(dev))
    native.check(rc, "ud_response_distill_forward")
    return mask
