"""Camera-branch operators: frustum->BEV voxel pooling and the fused LSS lift-splat
(host side of csrc/voxel_pool.cu; reference layers/blocks_3d/mmdet3d/lss_fpn.py:12-82,279-319)."""
import ctypes

import torch
from torch.autograd import Function

from . import native


class VoxelPooling(Function):
    """Same contract as the reference's VoxelPooling Function (lss_fpn.py:12-79):
    forward(geom_xyz int32 [B,...,3], input_features f32 [B,...,C], voxel_num int64[3]) -> (B, C, ny, nx)
    (a permuted view of the NHWC buffer, exactly like the reference); backward returns a gradient for
    input_features only."""

    @staticmethod
    def forward(ctx, geom_xyz, input_features, voxel_num):
        native.require_cuda(geom_xyz, input_features)
        assert geom_xyz.is_contiguous()
        assert input_features.is_contiguous()
        lib = native.load()
        ctx.mark_non_differentiable(geom_xyz)
        in_shape = input_features.shape
        geom = geom_xyz.reshape(geom_xyz.shape[0], -1, geom_xyz.shape[-1])
        feats = input_features.reshape(geom.shape[0], -1, in_shape[-1])
        assert geom.shape[1] == feats.shape[1]
        if geom.dtype != torch.int32:
            raise TypeError("voxel_pooling: geom_xyz must be int32 (the reference calls .int() first)")
        B, N, C = int(feats.shape[0]), int(feats.shape[1]), int(feats.shape[2])
        vn = voxel_num.tolist() if torch.is_tensor(voxel_num) else list(voxel_num)
        nx, ny, nz = int(vn[0]), int(vn[1]), int(vn[2])
        dev = feats.device
        out = torch.empty((B, ny, nx, C), dtype=torch.float32, device=dev)
        pos_memo = torch.empty((B, N, 3), dtype=torch.int32, device=dev)
        nbytes = native.query_bytes("ud_voxel_pooling_workspace_bytes", B, N, C, nx, ny)
        ws = native.WORKSPACE.get(nbytes, dev)
        rc = lib.ud_voxel_pooling_forward(B, N, C, nx, ny, nz, native.ptr(geom), native.ptr(feats.float()),
                                          native.ptr(out), native.ptr(pos_memo), native.ptr(ws), ws.numel(),
                                          native.stream_ptr(dev))
        native.check(rc, "ud_voxel_pooling_forward")
        ctx.save_for_backward(pos_memo)
        ctx.meta = (B, N, C, nx, ny, tuple(in_shape))
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, grad_output):
        (pos_memo,) = ctx.saved_tensors
        B, N, C, nx, ny, in_shape = ctx.meta
        lib = native.load()
        g = grad_output.contiguous().float()
        dev = g.device
        grad_in = torch.empty((B, N, C), dtype=torch.float32, device=dev)
        nbytes = native.query_bytes("ud_voxel_pooling_workspace_bytes", B, N, C, nx, ny)
        ws = native.WORKSPACE.get(nbytes, dev)
        rc = lib.ud_voxel_pooling_backward(B, N, C, nx, ny, native.ptr(g), native.ptr(pos_memo),
                                           native.ptr(grad_in), native.ptr(ws), ws.numel(),
                                           native.stream_ptr(dev))
        native.check(rc, "ud_voxel_pooling_backward")
        return None, grad_in.reshape(in_shape), None


voxel_pooling = VoxelPooling.apply


def _lift_splat_sizes(B, ncam, D, H, W, C, nx, ny):
    lib = native.load()
    ws, st = ctypes.c_size_t(0), ctypes.c_size_t(0)
    rc = lib.ud_lift_splat_sizes(B, ncam, D, H, W, C, nx, ny, ctypes.byref(ws), ctypes.byref(st))
    native.check(rc, "ud_lift_splat_sizes")
    return int(ws.value), int(st.value)


class LiftSplat(Function):
    """out (B,C,ny,nx) = splat(depth (x) ctx) without materialising the lifted volume.

    depth (B*ncam, D, H, W) softmax probabilities, ctx (B*ncam, C, H, W), geom_xyz (B, ncam, D, H, W, 3)
    int32 voxel indices.  Gradients flow to depth and ctx."""

    @staticmethod
    def forward(ctx_, depth, ctx, geom_xyz, voxel_num):
        native.require_cuda(depth, ctx, geom_xyz)
        lib = native.load()
        depth = depth.contiguous().float()
        ctx = ctx.contiguous().float()
        geom_xyz = geom_xyz.contiguous()
        if geom_xyz.dtype != torch.int32:
            raise TypeError("lift_splat: geom_xyz must be int32")
        B = int(geom_xyz.shape[0])
        ncam = int(geom_xyz.shape[1])
        BN, D, H, W = [int(s) for s in depth.shape]
        C = int(ctx.shape[1])
        assert BN == B * ncam and tuple(ctx.shape) == (BN, C, H, W)
        assert tuple(geom_xyz.shape) == (B, ncam, D, H, W, 3)
        vn = voxel_num.tolist() if torch.is_tensor(voxel_num) else list(voxel_num)
        nx, ny, nz = int(vn[0]), int(vn[1]), int(vn[2])
        dev = depth.device
        ws_bytes, st_bytes = _lift_splat_sizes(B, ncam, D, H, W, C, nx, ny)
        ws = native.WORKSPACE.get(ws_bytes, dev)
        state = torch.empty((st_bytes,), dtype=torch.uint8, device=dev)
        out = torch.empty((B, C, ny, nx), dtype=torch.float32, device=dev)
        rc = lib.ud_lift_splat_forward(B, ncam, D, H, W, C, nx, ny, nz, native.ptr(geom_xyz), native.ptr(depth),
                                       native.ptr(ctx), native.ptr(out), native.ptr(state), st_bytes,
                                       native.ptr(ws), ws.numel(), native.stream_ptr(dev))
        native.check(rc, "ud_lift_splat_forward")
        ctx_.save_for_backward(depth, state)
        ctx_.meta = (B, ncam, D, H, W, C, nx, ny, st_bytes)
        ctx_.mark_non_differentiable(geom_xyz)
        return out

    @staticmethod
    def backward(ctx_, grad_out):
        depth, state = ctx_.saved_tensors
        B, ncam, D, H, W, C, nx, ny, st_bytes = ctx_.meta
        lib = native.load()
        g = grad_out.contiguous().float()
        dev = g.device
        ws_bytes, _ = _lift_splat_sizes(B, ncam, D, H, W, C, nx, ny)
        ws = native.WORKSPACE.get(ws_bytes, dev)
        d_depth = torch.empty_like(depth)
        d_ctx = torch.empty((B * ncam, C, H, W), dtype=torch.float32, device=dev)
        rc = lib.ud_lift_splat_backward(B, ncam, D, H, W, C, nx, ny, native.ptr(g), native.ptr(depth),
                                        native.ptr(state), st_bytes, native.ptr(d_depth), native.ptr(d_ctx),
                                        native.ptr(ws), ws.numel(), native.stream_ptr(dev))
        native.check(rc, "ud_lift_splat_backward")
        return d_depth, d_ctx, None, None


lift_splat = LiftSplat.apply


def lss_geometry_index(frustum, sensor2ego_mat, intrin_mat, ida_mat, bda_mat, voxel_coord, voxel_size, want_geom=False):
    """Frustum -> ego coordinates -> BEV voxel index in one kernel (csrc/lss_geometry.cu): the reference's
    LSSFPN.get_geometry + `.int()` index (lss_fpn.py:200-240, 311-313) without the three 1.9 M-product batched matmuls.
    frustum (D,H,W,4); sensor2ego_mat / intrin_mat / ida_mat (B,ncam,4,4); bda_mat (B,4,4) | None.
    -> geom_xyz int32 (B,ncam,D,H,W,3) [, geom f32 of the same shape]."""
    native.require_cuda(frustum, sensor2ego_mat, intrin_mat, ida_mat)
    lib = native.load()
    B, ncam = int(sensor2ego_mat.shape[0]), int(sensor2ego_mat.shape[1])
    D, H, W = [int(s) for s in frustum.shape[:3]]
    dev = frustum.device
    # linalg.inv_ex == inverse() without the host-side singularity check (a device -> host read per call)
    ida_inv = torch.linalg.inv_ex(ida_mat.reshape(B * ncam, 4, 4).float())[0].contiguous()
    cam2ego = sensor2ego_mat.float().matmul(torch.linalg.inv_ex(intrin_mat.float())[0]).reshape(B * ncam, 4, 4).contiguous()
    bda = bda_mat.reshape(B, 4, 4).float().contiguous() if bda_mat is not None else None
    fr = frustum.float().contiguous()
    key = (voxel_coord.data_ptr(), voxel_size.data_ptr(), voxel_coord._version, voxel_size._version)
    host = _VOXEL_HOST.get(key)
    if host is None:                                   # two tiny buffers, read to the host once per module
        lo = (voxel_coord - voxel_size / 2.0).float().cpu().tolist()
        host = (native.host_floats(lo), native.host_floats(voxel_size.float().cpu().tolist()))
        if len(_VOXEL_HOST) > 64:
            _VOXEL_HOST.clear()
        _VOXEL_HOST[key] = host
    idx = torch.empty((B, ncam, D, H, W, 3), dtype=torch.int32, device=dev)
    geom = torch.empty((B, ncam, D, H, W, 3), dtype=torch.float32, device=dev) if want_geom else None
    rc = lib.ud_lss_geometry(native.ptr(fr), D * H * W, B, ncam, native.ptr(ida_inv), native.ptr(cam2ego), native.ptr(bda),
                             native.host_ptr(host[0]), native.host_ptr(host[1]), native.ptr(idx), native.ptr(geom),
                             native.stream_ptr(dev))
    native.check(rc, "ud_lss_geometry")
    return (idx, geom) if want_geom else idx


_VOXEL_HOST = {}
