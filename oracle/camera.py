"""oracle/camera.py -- TEST INFRASTRUCTURE (checker / reported CPU baseline only).

CPU restatement of the reference's camera branch between depth_net and the BEV map:
  frustum / geometry      <- LSSFPN.create_frustum / get_geometry, lss_fpn.py:173-240 (torch CPU, same ops)
  voxel index             <- lss_fpn.py:311-313  ((geom - (coord - size/2)) / size).int()  -- truncation
  lift (depth (x) ctx)    <- lss_fpn.py:289-310
  voxel_pooling fwd/bwd   <- lss_fpn.py:12-82 + SASS-derived kernel semantics (oracle/camera_ref.c)
"""
import ctypes

import numpy as np
import torch

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build())
        for name in ("ud_ref_voxel_pooling", "ud_ref_voxel_pooling_backward", "ud_ref_lift_splat"):
            getattr(_lib, name).restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def create_frustum(final_dim, downsample, d_bound):
    """lss_fpn.py:173-198 -> (D, fH, fW, 4) [u, v, d, 1]"""
    ogfH, ogfW = final_dim
    fH, fW = ogfH // downsample, ogfW // downsample
    d = torch.arange(*d_bound, dtype=torch.float).view(-1, 1, 1).expand(-1, fH, fW)
    D = d.shape[0]
    x = torch.linspace(0, ogfW - 1, fW, dtype=torch.float).view(1, 1, fW).expand(D, fH, fW)
    y = torch.linspace(0, ogfH - 1, fH, dtype=torch.float).view(1, fH, 1).expand(D, fH, fW)
    return torch.stack((x, y, d, torch.ones_like(d)), -1)


def get_geometry(frustum, sensor2ego, intrin, ida, bda):
    """lss_fpn.py:200-240 -> (B, ncam, D, H, W, 3) float32 ego coordinates"""
    B, ncam = sensor2ego.shape[:2]
    pts = ida.view(B, ncam, 1, 1, 1, 4, 4).inverse().matmul(frustum.unsqueeze(-1))
    pts = torch.cat((pts[:, :, :, :, :, :2] * pts[:, :, :, :, :, 2:3], pts[:, :, :, :, :, 2:]), 5)
    combine = sensor2ego.matmul(torch.inverse(intrin))
    pts = combine.view(B, ncam, 1, 1, 1, 4, 4).matmul(pts)
    if bda is not None:
        bda = bda.unsqueeze(1).repeat(1, ncam, 1, 1).view(B, ncam, 1, 1, 1, 4, 4)
        pts = (bda @ pts).squeeze(-1)
    else:
        pts = pts.squeeze(-1)
    return pts[..., :3]


def voxel_indices(geom, voxel_coord, voxel_size):
    """lss_fpn.py:311-313 (note .int() truncates toward zero: (-1,0) lands in cell 0, quirk 4)"""
    return ((geom - (voxel_coord - voxel_size / 2.0)) / voxel_size).int()


def voxel_pooling(geom_xyz, feats, voxel_num):
    """-> (out (B,C,ny,nx) float32, pos_memo (B,N,3) int32); geom (B,...,3) int32, feats (B,...,C)."""
    geom = np.ascontiguousarray(np.asarray(geom_xyz, np.int32).reshape(geom_xyz.shape[0], -1, 3))
    f = np.ascontiguousarray(np.asarray(feats, np.float32).reshape(geom.shape[0], -1, feats.shape[-1]))
    B, N, C = f.shape
    nx, ny, nz = [int(v) for v in voxel_num]
    out = np.zeros((B, ny, nx, C), np.float32)
    pos = np.full((B, N, 3), -1, np.int32)
    lib().ud_ref_voxel_pooling(B, N, C, nx, ny, nz, _p(geom), _p(f), _p(out), _p(pos))
    return out.transpose(0, 3, 1, 2), pos


def voxel_pooling_backward(grad_out, pos_memo, C):
    g = np.ascontiguousarray(grad_out, np.float32)
    B, _, ny, nx = g.shape
    N = pos_memo.shape[1]
    gi = np.zeros((B, N, C), np.float32)
    lib().ud_ref_voxel_pooling_backward(B, N, C, nx, ny, _p(g), _p(np.ascontiguousarray(pos_memo)), _p(gi))
    return gi


def lift(depth, ctx):
    """lss_fpn.py:289-310: (B*ncam, D, H, W) x (B*ncam, C, H, W) -> (B*ncam, D, H, W, C) (per-camera layout of
    the permuted volume the reference feeds to voxel_pooling)."""
    d = torch.as_tensor(depth)
    c = torch.as_tensor(ctx)
    vol = d.unsqueeze(1) * c.unsqueeze(2)            # (BN, C, D, H, W)
    return vol.permute(0, 2, 3, 4, 1).contiguous()


def lift_splat_reference(depth, ctx, geom_xyz, voxel_num):
    """Reference pipeline with the materialised volume: lift -> reshape (B, ncam*D*H*W, C) -> voxel_pooling."""
    B = geom_xyz.shape[0]
    vol = lift(depth, ctx).numpy()
    C = vol.shape[-1]
    out, pos = voxel_pooling(np.asarray(geom_xyz), vol.reshape(B, -1, C), voxel_num)
    return np.ascontiguousarray(out)


def lift_splat_fused(depth, ctx, geom_xyz, voxel_num):
    """Same result without storing the volume (CPU baseline at full size)."""
    depth = np.ascontiguousarray(depth, np.float32)
    ctx = np.asarray(ctx, np.float32)
    BN, D, H, W = depth.shape
    C = ctx.shape[1]
    geom = np.ascontiguousarray(np.asarray(geom_xyz, np.int32).reshape(geom_xyz.shape[0], -1, 3))
    B = geom.shape[0]
    ncam = BN // B
    ctx_nhwc = np.ascontiguousarray(ctx.transpose(0, 2, 3, 1))
    nx, ny, nz = [int(v) for v in voxel_num]
    out = np.zeros((B, ny, nx, C), np.float32)
    lib().ud_ref_lift_splat(B, ncam, D, H, W, C, nx, ny, nz, _p(geom), _p(depth), _p(ctx_nhwc), _p(out))
    return np.ascontiguousarray(out.transpose(0, 3, 1, 2))
