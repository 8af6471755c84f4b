"""oracle/lidar.py -- TEST INFRASTRUCTURE (checker / reported CPU baseline only).

numpy / torch-CPU face of oracle/lidar_ref.c plus a CPU restatement of the reference's LiDAR encoder:
  voxelize            <- unidistill/data/det3d/preprocess/voxelization.py:44-73 (+ spconv PointToVoxel, [EXT])
  mean_vfe            <- unidistill/layers/blocks_3d/det3d/vfe/mean_vfe.py:25-32
  rulebook_* / sparse_conv_* <- spconv get_indice_pairs / indice_conv Native ([EXT], spconv>=2.1.12)
  backbone_forward    <- unidistill/layers/blocks_3d/det3d/spconv_backbone.py:252-384 (VoxelResBackBone8x)
  dense / height_compression <- unidistill/layers/blocks_2d/det3d/map_to_bev/height_compression.py:11-22
Parity of the spconv pieces is UNPINNED by the reference (no tests, wheel not vendored); they are pinned
by known-answer tests and dense-conv3d equivalence in tests/test_oracle_lidar.py.
"""
import ctypes
import os

import numpy as np
import torch
import torch.nn.functional as F

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_oracle.build()
        _lib = ctypes.CDLL(path)
        _lib.ud_ref_voxelize.restype = ctypes.c_int
        _lib.ud_ref_rulebook_subm.restype = ctypes.c_int
        _lib.ud_ref_rulebook_conv.restype = ctypes.c_int
        _lib.ud_ref_sparse_conv.restype = None
        _lib.ud_ref_dense.restype = None
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ---------------------------------------------------------------------------------------------------
def voxelize_one(points, voxel_size, pc_range, grid_size, max_points, max_voxels):
    """First-come hard voxelization of one cloud -> (voxels (M,P,C), coords (M,3) zyx, num (M,), mean (M,C))."""
    pts = _f32(points)
    n, C = pts.shape
    max_voxels = max(1, min(int(max_voxels), n))     # n points cannot open more than n voxels: same result, less zeroing
    voxels = np.zeros((max_voxels, max_points, C), np.float32)
    coords = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    mean = np.zeros((max_voxels, C), np.float32)
    vs, pr, gs = _f32(voxel_size), _f32(pc_range), _i32(grid_size)
    m = lib().ud_ref_voxelize(_p(pts), n, C, _p(vs), _p(pr), _p(gs), int(max_points), int(max_voxels),
                              _p(voxels), _p(coords), _p(num), _p(mean))
    if m < 0:
        raise MemoryError("ud_ref_voxelize")
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy(), mean[:m].copy()


def voxelize(points_list, voxel_size, pc_range, grid_size, max_points, max_voxels):
    """Batch form == Voxelization.forward (voxelization.py:53-73): per-sample loop, batch index prepended."""
    vox, crd, num, mean = [], [], [], []
    for b, pts in enumerate(points_list):
        v, c, k, m = voxelize_one(pts, voxel_size, pc_range, grid_size, max_points, max_voxels)
        vox.append(v)
        crd.append(np.concatenate([np.full((c.shape[0], 1), b, np.int32), c], axis=1))
        num.append(k)
        mean.append(m)
    return (np.concatenate(vox, 0), np.concatenate(crd, 0).astype(np.int32), np.concatenate(num, 0),
            np.concatenate(mean, 0))


def mean_vfe(voxels, num):
    """mean_vfe.py:25-32 with numpy semantics (sum over the slot axis / clamp(num,1))."""
    s = voxels.sum(axis=1, dtype=np.float32)
    return (s / np.maximum(num.reshape(-1, 1), 1).astype(np.float32)).astype(np.float32)


# ---------------------------------------------------------------------------------------------------
def _t3(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple, np.ndarray)) else [int(v)] * 3


def conv_out_shape(in_shape, ksize, stride, padding, dilation):
    return [(in_shape[d] + 2 * padding[d] - dilation[d] * (ksize[d] - 1) - 1) // stride[d] + 1 for d in range(3)]


def rulebook_subm(coords, spatial_shape, ksize=3, dilation=1):
    coords = _i32(coords)
    n = coords.shape[0]
    ks, dl, shp = _t3(ksize), _t3(dilation), _i32(spatial_shape)
    K = ks[0] * ks[1] * ks[2]
    pairs = np.empty((2, K, max(n, 1)), np.int32)
    pair_num = np.zeros((K,), np.int32)
    rc = lib().ud_ref_rulebook_subm(_p(coords), n, _p(shp), _p(_i32(ks)), _p(_i32(dl)), _p(pairs), _p(pair_num))
    assert rc == 0
    return pairs[:, :, :n] if n else pairs[:, :, :0], pair_num


def rulebook_conv(coords, spatial_shape, ksize, stride, padding, dilation=1):
    coords = _i32(coords)
    n = coords.shape[0]
    ks, st, pd, dl = _t3(ksize), _t3(stride), _t3(padding), _t3(dilation)
    K = ks[0] * ks[1] * ks[2]
    out_shape = conv_out_shape(list(spatial_shape), ks, st, pd, dl)
    out_coords = np.zeros((max(n * K, 1), 4), np.int32)
    pairs = np.empty((2, K, max(n, 1)), np.int32)
    pair_num = np.zeros((K,), np.int32)
    n_out = lib().ud_ref_rulebook_conv(_p(coords), n, _p(_i32(spatial_shape)), _p(_i32(out_shape)), _p(_i32(ks)),
                                       _p(_i32(st)), _p(_i32(pd)), _p(_i32(dl)), _p(out_coords), _p(pairs),
                                       _p(pair_num))
    assert n_out >= 0
    return out_coords[:n_out].copy(), (pairs[:, :, :n] if n else pairs[:, :, :0]), pair_num, out_shape


def pair_set(pairs, pair_num):
    """Canonical, order-free view of a rulebook: sorted (k, i, o) triples."""
    out = []
    for k in range(pairs.shape[1]):
        m = int(pair_num[k])
        if m:
            out.append(np.stack([np.full((m,), k, np.int64), pairs[0, k, :m].astype(np.int64),
                                 pairs[1, k, :m].astype(np.int64)], 1))
    if not out:
        return np.zeros((0, 3), np.int64)
    a = np.concatenate(out, 0)
    return a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]


def weight_to_taps(weight):
    """(Cout,kz,ky,kx,Cin) -> (K,Cin,Cout)"""
    w = np.asarray(weight, np.float32)
    cout, cin = w.shape[0], w.shape[-1]
    return np.ascontiguousarray(w.reshape(cout, -1, cin).transpose(1, 2, 0))


def sparse_conv_c(feat, weight, bias, pairs, pair_num, n_out):
    """Plain-C sequential Native algorithm (exact fp32 order: k, then pair order)."""
    feat = _f32(feat)
    w = weight_to_taps(weight)
    K, cin, cout = w.shape
    out = np.zeros((n_out, cout), np.float32)
    b = _f32(bias) if bias is not None else None
    pairs = _i32(pairs)
    # the C routine indexes pairs with stride n_in == pairs.shape[2]
    assert pairs.shape[2] == feat.shape[0] or feat.shape[0] == 0
    lib().ud_ref_sparse_conv(_p(feat), feat.shape[0], _p(w), _p(b), _p(pairs), _p(_i32(pair_num)), K, cin, cout,
                             _p(out), n_out)
    return out


def sparse_conv_torch(feat, weight, bias, pairs, pair_num, n_out):
    """spconv ConvAlgo.Native restated with torch CPU ops: per offset gather -> mm -> index_add_.
    This is the timed CPU baseline (uses all host threads through torch's BLAS)."""
    feat = torch.as_tensor(feat, dtype=torch.float32)
    w = torch.as_tensor(weight, dtype=torch.float32)
    cout, cin = w.shape[0], w.shape[-1]
    w_k = w.reshape(cout, -1, cin).permute(1, 2, 0).contiguous()
    out = torch.zeros((n_out, cout), dtype=torch.float32)
    if bias is not None:
        out += torch.as_tensor(bias, dtype=torch.float32)
    pairs_t = torch.as_tensor(np.asarray(pairs)).long()
    for k in range(w_k.shape[0]):
        m = int(pair_num[k])
        if m == 0:
            continue
        i_idx, o_idx = pairs_t[0, k, :m], pairs_t[1, k, :m]
        out.index_add_(0, o_idx, feat.index_select(0, i_idx) @ w_k[k])
    return out


def dense(feat, coords, batch_size, spatial_shape):
    feat, coords = _f32(feat), _i32(coords)
    n, C = feat.shape
    D, H, W = [int(s) for s in spatial_shape]
    out = np.zeros((batch_size, C, D, H, W), np.float32)
    lib().ud_ref_dense(_p(feat), _p(coords), n, C, D, H, W, _p(out))
    return out


def height_compression(feat, coords, batch_size, spatial_shape):
    d = dense(feat, coords, batch_size, spatial_shape)
    B, C, D, H, W = d.shape
    return d.reshape(B, C * D, H, W)


# ---------------------------------------------------------------------------------------------------
# VoxelResBackBone8x restated over the oracle pieces (spconv_backbone.py:252-384).
def _bn(x, sd, prefix, training, eps=1e-3, momentum=0.01):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"].clone(), sd[prefix + ".running_var"].clone()
    return F.batch_norm(x, rm, rv, w, b, training, momentum, eps)


class _SpT:
    def __init__(self, feat, coords, shape, batch_size):
        self.feat, self.coords, self.shape, self.batch_size = feat, coords, list(shape), batch_size


def backbone_forward(state_dict, voxel_features, voxel_coords, batch_size, sparse_shape, training=True,
                     conv_fn=sparse_conv_torch, return_stages=False):
    """CPU forward of VoxelResBackBone8x.  state_dict: the module's own state_dict (torch CPU tensors).
    Returns (features (N_out,128), coords (N_out,4), spatial_shape [2,H/8,W/8])."""
    sd = {k: torch.as_tensor(v).float().cpu() if torch.is_tensor(v) and v.is_floating_point() else v
          for k, v in state_dict.items()}
    rb_cache = {}

    def subm(x, wkey, bkey, rb_key):
        if rb_key not in rb_cache:
            rb_cache[rb_key] = rulebook_subm(x.coords, x.shape, 3, 1)
        pairs, pn = rb_cache[rb_key]
        bias = sd[bkey] if bkey in sd else None
        return conv_fn(x.feat, sd[wkey], bias, pairs, pn, x.feat.shape[0])

    def sbb(x, prefix, rb_key):
        # SparseBasicBlock.forward, spconv_backbone.py:97-113
        out = subm(x, prefix + ".conv1.weight", prefix + ".conv1.bias", rb_key)
        out = torch.relu(_bn(out, sd, prefix + ".bn1", training))
        xo = _SpT(out, x.coords, x.shape, x.batch_size)
        out = subm(xo, prefix + ".conv2.weight", prefix + ".conv2.bias", rb_key)
        out = _bn(out, sd, prefix + ".bn2", training)
        out = torch.relu(out + x.feat)
        return _SpT(out, x.coords, x.shape, x.batch_size)

    def down(x, prefix, ksize, stride, padding):
        oc, pairs, pn, oshape = rulebook_conv(x.coords, x.shape, ksize, stride, padding, 1)
        out = conv_fn(x.feat, sd[prefix + ".0.weight"], None, pairs, pn, oc.shape[0])
        out = torch.relu(_bn(out, sd, prefix + ".1", training))
        return _SpT(out, oc, oshape, x.batch_size)

    x = _SpT(torch.as_tensor(voxel_features, dtype=torch.float32), _i32(voxel_coords), sparse_shape, batch_size)
    f = subm(x, "conv_input.0.weight", "conv_input.0.bias", "subm1")
    x = _SpT(torch.relu(_bn(f, sd, "conv_input.1", training)), x.coords, x.shape, batch_size)
    stages = {}
    x = sbb(x, "conv1.0", "subm1")
    x = sbb(x, "conv1.1", "subm1")
    stages["x_conv1"] = x
    x = down(x, "conv2.0", 3, 2, 1)
    x = sbb(x, "conv2.1", "res2")
    x = sbb(x, "conv2.2", "res2")
    stages["x_conv2"] = x
    x = down(x, "conv3.0", 3, 2, 1)
    x = sbb(x, "conv3.1", "res3")
    x = sbb(x, "conv3.2", "res3")
    stages["x_conv3"] = x
    x = down(x, "conv4.0", 3, 2, (0, 1, 1))
    x = sbb(x, "conv4.1", "res4")
    x = sbb(x, "conv4.2", "res4")
    stages["x_conv4"] = x
    x = down(x, "conv_out", (3, 1, 1), (2, 1, 1), 0)
    if return_stages:
        return x.feat, x.coords, x.shape, stages
    return x.feat, x.coords, x.shape
