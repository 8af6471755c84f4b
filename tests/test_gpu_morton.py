"""Morton (Z-order) row ordering of the sparse tensors (csrc/rulebook.cu `order` = 1, ud_morton_order): the same
coordinate sets, pair sets and convolution results as under the canonical order, only permuted -- and 128-row tiles that
touch fewer distinct input rows.  (An option, not the default: under the raster order the rows a tap gathers are consecutive
in memory, which measured faster on B200 -- unidistill_b200/sparse.py.)"""
import numpy as np
import pytest
import torch

from oracle import lidar as O
from test_gpu_lidar import _clustered_sparse

pytestmark = pytest.mark.gpu


def _bits(n):
    b = 0
    while (1 << b) < n:
        b += 1
    return b


def _morton_keys(coords, shape):
    """numpy restatement of the key layout of csrc/rulebook.cu (bits dealt x, y, z round-robin from bit 0)."""
    nb = [_bits(shape[0]), _bits(shape[1]), _bits(shape[2])]
    pos = [[0] * 12 for _ in range(3)]
    used = [0, 0, 0]
    p, total = 0, sum(nb)
    while p < total:
        for a in (2, 1, 0):
            if used[a] < nb[a]:
                pos[a][used[a]] = p
                used[a] += 1
                p += 1
    key = coords[:, 0].astype(np.int64) << total
    for a in range(3):
        v = coords[:, 1 + a].astype(np.int64)
        for i in range(nb[a]):
            key |= ((v >> i) & 1) << pos[a][i]
    return key


@pytest.mark.parametrize("ksize,stride,pad", [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)])
def test_morton_rulebook_is_a_permutation_of_the_canonical_one(cuda, native_lib, ksize, stride, pad):
    from unidistill_b200 import sparse

    rng = np.random.default_rng(3)
    B, shape, n = 3, [11, 30, 28], 5000
    coords, feat = _clustered_sparse(rng, B, shape, n, 16)
    ct = torch.from_numpy(coords).to(cuda)
    lin = sparse.build_conv_rulebook(ct, B, shape, ksize, stride, pad, order="linear")
    mor = sparse.build_conv_rulebook(ct, B, shape, ksize, stride, pad, order="morton")
    assert lin.n_out == mor.n_out and lin.out_shape == mor.out_shape
    oc_l, oc_m = lin.out_coords.cpu().numpy(), mor.out_coords.cpu().numpy()
    keys = _morton_keys(oc_m, mor.out_shape)
    assert np.all(np.diff(keys) > 0), "rows are not in ascending Morton-key order"
    # same coordinate set; perm[m] = canonical row of Morton row m
    lin_key = {tuple(c): i for i, c in enumerate(oc_l.tolist())}
    perm = np.array([lin_key[tuple(c)] for c in oc_m.tolist()])
    assert sorted(perm.tolist()) == list(range(lin.n_out))
    # identical pair sets under the permutation
    assert np.array_equal(mor.nbr_out.cpu().numpy(), lin.nbr_out.cpu().numpy()[perm])
    inv = np.empty_like(perm)
    inv[perm] = np.arange(perm.size)
    t_in_l = lin.nbr_in.cpu().numpy()
    assert np.array_equal(mor.nbr_in.cpu().numpy(), np.where(t_in_l >= 0, inv[np.maximum(t_in_l, 0)], -1))
    # convolution results: the same rows (bit for bit: per-row accumulation order does not depend on the row order)
    w = torch.from_numpy((rng.standard_normal((32,) + tuple(O._t3(ksize)) + (16,)) * 0.1).astype(np.float32)).to(cuda)
    f = torch.from_numpy(feat).to(cuda)
    out_l = sparse.sparse_conv(f, w, None, lin)
    out_m = sparse.sparse_conv(f, w, None, mor)
    assert torch.equal(out_m, out_l[torch.from_numpy(perm).to(cuda)])


def test_morton_order_of_a_coordinate_set(cuda, native_lib):
    from unidistill_b200 import sparse

    rng = np.random.default_rng(4)
    B, shape = 4, [41, 128, 96]
    coords, _ = _clustered_sparse(rng, B, shape, 30000, 1)
    coords = coords[rng.permutation(coords.shape[0])]                 # arbitrary (first-come like) input order
    order = sparse.morton_order(torch.from_numpy(coords).to(cuda), B, shape).cpu().numpy()
    assert sorted(order.tolist()) == list(range(coords.shape[0]))
    keys = _morton_keys(coords[order], shape)
    assert np.all(np.diff(keys) > 0)
    assert sparse.morton_order(torch.zeros((0, 4), dtype=torch.int32, device=cuda), 1, shape).numel() == 0


def test_tiles_touch_fewer_distinct_rows(cuda, native_lib):
    """The point of the ordering: distinct input rows gathered by a 128-row tile of a SubM rulebook."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(6)
    B, shape = 1, [21, 256, 256]
    # a ground-plane like sheet + clutter: what a LiDAR voxel set looks like at stage 2
    yy, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    z = (6 + 2 * np.sin(xx / 40.0) + 2 * np.cos(yy / 33.0)).astype(np.int32)
    sheet = np.stack([np.zeros_like(z), z, yy, xx], -1).reshape(-1, 4)
    sheet = sheet[rng.random(sheet.shape[0]) < 0.6]
    coords = np.unique(sheet, axis=0).astype(np.int32)

    def distinct_per_tile(c):
        rb = sparse.build_subm_rulebook(torch.from_numpy(np.ascontiguousarray(c)).to(cuda), B, shape, 3)
        t = rb.nbr_out.cpu().numpy()
        vals = []
        for r0 in range(0, t.shape[0] - 127, 128 * 8):
            blk = t[r0:r0 + 128]
            vals.append(np.unique(blk[blk >= 0]).size)
        return float(np.mean(vals))

    lin = coords[np.lexsort((coords[:, 3], coords[:, 2], coords[:, 1], coords[:, 0]))]
    order = sparse.morton_order(torch.from_numpy(lin).to(cuda), B, shape).cpu().numpy()
    a, b = distinct_per_tile(lin), distinct_per_tile(lin[order])
    assert b < 0.9 * a, (a, b)              # measured 208 -> 175 distinct rows per tile on this sheet-like set


def test_backbone_same_bev_under_both_orders(cuda, native_lib):
    """VoxelResBackBone8x + scatter: Morton row order (default) vs canonical order give the same BEV map (BatchNorm sums
    are taken in a different row order: 1e-5)."""
    from unidistill.layers.blocks_2d.det3d import HeightCompression
    from unidistill.layers.blocks_3d.det3d import VoxelResBackBone8x
    from unidistill_b200 import sparse

    rng = np.random.default_rng(8)
    grid = [64, 64, 40]
    coords, feat = _clustered_sparse(rng, 2, [41, 64, 64], 6000, 5)
    torch.manual_seed(0)
    net = VoxelResBackBone8x(5, np.array(grid)).to(cuda).train()
    f, c = torch.from_numpy(feat).to(cuda), torch.from_numpy(coords).to(cuda)
    outs = []
    for mode in ("morton", "linear"):
        old = sparse.CONV_OUTPUT_ORDER
        sparse.CONV_OUTPUT_ORDER = mode
        net.morton_order = mode == "morton"
        try:
            with torch.no_grad():
                sp, stride, _ = net(f, c, 2)
                outs.append(HeightCompression(256)(sp, stride)[0])
        finally:
            sparse.CONV_OUTPUT_ORDER = old
    scale = float(outs[1].abs().max())
    assert float((outs[0] - outs[1]).abs().max()) <= 1e-4 * scale
