"""Pins the CPU oracle of the LiDAR branch (oracle/lidar_ref.c, oracle/lidar.py).

The reference has no tests or golden vectors for these spconv-backed functions, so the
oracle is pinned by hand-computable known answers and by equivalence with dense torch convolutions.
"""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import lidar as O

VS = [0.5, 0.5, 1.0]
RANGE = [0.0, 0.0, 0.0, 4.0, 4.0, 2.0]
GRID = [8, 8, 2]


def _pts(rows):
    return np.array(rows, np.float32)


# ---- voxelizer known answers -------------------------------------------------------
def test_single_point_at_half_voxel_lands_in_voxel_zero():
    v, c, n, m = O.voxelize_one(_pts([[0.25, 0.25, 0.5, 7.0, 0.1]]), VS, RANGE, GRID, 10, 100)
    assert c.tolist() == [[0, 0, 0]] and n.tolist() == [1]
    assert np.array_equal(v[0, 0], np.array([0.25, 0.25, 0.5, 7.0, 0.1], np.float32))
    assert np.array_equal(v[0, 1:], np.zeros((9, 5), np.float32))
    assert np.array_equal(m[0], v[0, 0])


def test_upper_bound_is_exclusive_and_lower_inclusive():
    pts = _pts([[4.0, 1.0, 1.0, 0, 0], [1.0, 4.0, 1.0, 0, 0], [1.0, 1.0, 2.0, 0, 0], [-1e-6, 1.0, 1.0, 0, 0],
                [0.0, 0.0, 0.0, 0, 0], [3.999, 3.999, 1.999, 0, 0]])
    v, c, n, m = O.voxelize_one(pts, VS, RANGE, GRID, 10, 100)
    assert c.tolist() == [[0, 0, 0], [1, 7, 7]]   # zyx


def test_eleven_points_keep_first_ten_in_order():
    pts = np.zeros((11, 5), np.float32)
    pts[:, :3] = [1.1, 1.1, 0.5]
    pts[:, 3] = np.arange(11)
    v, c, n, m = O.voxelize_one(pts, VS, RANGE, GRID, 10, 100)
    assert n.tolist() == [10]
    assert v[0, :, 3].tolist() == list(range(10))
    assert m[0, 3] == np.float32(4.5)


def test_first_come_order_and_cap_skips_new_voxels_only():
    # voxels appear in order A, B, C ; cap = 2 -> C never created, later points of A/B still counted
    A, B, C = [0.1, 0.1, 0.1], [1.1, 0.1, 0.1], [2.1, 0.1, 0.1]
    pts = _pts([A + [0, 0], B + [1, 0], C + [2, 0], A + [3, 0], C + [4, 0], B + [5, 0]])
    v, c, n, m = O.voxelize_one(pts, VS, RANGE, GRID, 10, 2)
    assert c.tolist() == [[0, 0, 0], [0, 0, 2]]
    assert n.tolist() == [2, 2]
    assert v[0, :2, 3].tolist() == [0, 3] and v[1, :2, 3].tolist() == [1, 5]


def test_zero_padded_points_fill_origin_voxel():
    # quirk 1: collate_fn zero padding yields real points at the origin
    rng_pts = _pts([[10.0, 10.0, 0.0, 1, 0]])
    pad = np.zeros((5, 5), np.float32)
    r = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
    v, c, n, m = O.voxelize_one(np.concatenate([rng_pts, pad]), [0.1, 0.1, 0.2], r, [1024, 1024, 40], 10, 100)
    assert c.shape[0] == 2 and n.tolist() == [1, 5]
    assert c[1].tolist() == [25, 512, 512]


def test_batch_form_prepends_batch_index():
    a = _pts([[0.25, 0.25, 0.5, 0, 0]])
    b = _pts([[1.25, 0.25, 0.5, 0, 0], [0.25, 0.25, 0.5, 0, 0]])
    v, c, n, m = O.voxelize([a, b], VS, RANGE, GRID, 10, 100)
    assert c.tolist() == [[0, 0, 0, 0], [1, 0, 0, 2], [1, 0, 0, 0]]
    assert np.allclose(O.mean_vfe(v, n), m)


# ---- rulebooks (SURVEY 8c (ii),(iii)) ------------------------------------------------------------
def test_subm_isolated_voxel_only_centre_tap():
    pairs, pn = O.rulebook_subm(np.array([[0, 3, 3, 3]], np.int32), [8, 8, 8], 3)
    assert pn.tolist() == [0] * 13 + [1] + [0] * 13
    assert pairs[0, 13, 0] == 0 and pairs[1, 13, 0] == 0


def test_subm_two_neighbours_x():
    # rows: 0 at x=3, 1 at x=4.  tap k (kx=2 -> +1 in x) reads input at out+1: out 0 <- in 1 at k=14
    coords = np.array([[0, 3, 3, 3], [0, 3, 3, 4]], np.int32)
    pairs, pn = O.rulebook_subm(coords, [8, 8, 8], 3)
    s = O.pair_set(pairs, pn).tolist()
    assert s == [[12, 0, 1], [13, 0, 0], [13, 1, 1], [14, 1, 0]]


def test_subm_does_not_cross_batches_or_borders():
    coords = np.array([[0, 0, 0, 0], [1, 0, 0, 1]], np.int32)
    pairs, pn = O.rulebook_subm(coords, [2, 2, 2], 3)
    assert pn.sum() == 2


def test_strided_conv_single_voxel_even_vs_odd():
    # k=3, s=2, p=1: out = (in + 1 - k) / 2.  Even coordinates feed 1 output per axis (k=1), odd ones 2
    # (k=0 and k=2) -> 1 vs 8 outputs for an isolated voxel.
    oc, pairs, pn, osh = O.rulebook_conv(np.array([[0, 2, 2, 2]], np.int32), [8, 8, 8], 3, 2, 1)
    assert osh == [4, 4, 4] and oc.tolist() == [[0, 1, 1, 1]] and pn.sum() == 1 and pn[13] == 1
    oc, pairs, pn, osh = O.rulebook_conv(np.array([[0, 3, 3, 3]], np.int32), [8, 8, 8], 3, 2, 1)
    assert oc.shape[0] == 8 and pn.sum() == 8
    assert oc.tolist() == sorted(oc.tolist())          # ascending linear order
    assert oc[0].tolist() == [0, 1, 1, 1] and oc[-1].tolist() == [0, 2, 2, 2]
    # tap (0,0,0) reaches the far output (in = 2*out - 1 + 0), tap (2,2,2) the near one
    assert pairs[1, 0, 0] == 7 and pairs[1, 26, 0] == 0


def test_strided_conv_shapes_of_the_backbone():
    # spconv_backbone.py:279-340 shape chain for the 128^2 benchmark grid
    shape = [41, 1024, 1024]
    shape = O.conv_out_shape(shape, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1]); assert shape == [21, 512, 512]
    shape = O.conv_out_shape(shape, [3, 3, 3], [2, 2, 2], [1, 1, 1], [1, 1, 1]); assert shape == [11, 256, 256]
    shape = O.conv_out_shape(shape, [3, 3, 3], [2, 2, 2], [0, 1, 1], [1, 1, 1]); assert shape == [5, 128, 128]
    shape = O.conv_out_shape(shape, [3, 1, 1], [2, 1, 1], [0, 0, 0], [1, 1, 1]); assert shape == [2, 128, 128]


def _random_sparse(rng, B, shape, n, C):
    lin = rng.choice(B * shape[0] * shape[1] * shape[2], n, replace=False)
    x = lin % shape[2]; lin //= shape[2]
    y = lin % shape[1]; lin //= shape[1]
    z = lin % shape[0]; b = lin // shape[0]
    coords = np.stack([b, z, y, x], 1).astype(np.int32)
    feat = rng.standard_normal((n, C)).astype(np.float32)
    return coords, feat


def _densify(coords, feat, B, shape):
    d = torch.zeros((B, feat.shape[1]) + tuple(shape))
    c = torch.from_numpy(coords).long()
    d[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = torch.from_numpy(feat)
    return d


@pytest.mark.parametrize("conv_fn", [O.sparse_conv_c, O.sparse_conv_torch])
def test_subm_conv_equals_dense_conv3d_on_active_sites(conv_fn):
    rng = np.random.default_rng(0)
    B, shape, n, cin, cout = 2, [5, 9, 8], 120, 6, 7
    coords, feat = _random_sparse(rng, B, shape, n, cin)
    w = rng.standard_normal((cout, 3, 3, 3, cin)).astype(np.float32)
    bias = rng.standard_normal((cout,)).astype(np.float32)
    pairs, pn = O.rulebook_subm(coords, shape, 3)
    out = np.asarray(conv_fn(feat, w, bias, pairs, pn, n))
    dense = F.conv3d(_densify(coords, feat, B, shape), torch.from_numpy(w).permute(0, 4, 1, 2, 3),
                     torch.from_numpy(bias), padding=1)
    c = torch.from_numpy(coords).long()
    ref = dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]].numpy()
    assert np.allclose(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("ksize,stride,pad", [(3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)])
def test_strided_conv_equals_dense_conv3d(ksize, stride, pad):
    rng = np.random.default_rng(1)
    B, shape, n, cin, cout = 2, [7, 10, 12], 150, 4, 5
    coords, feat = _random_sparse(rng, B, shape, n, cin)
    ks = O._t3(ksize)
    w = rng.standard_normal((cout, ks[0], ks[1], ks[2], cin)).astype(np.float32)
    oc, pairs, pn, osh = O.rulebook_conv(coords, shape, ksize, stride, pad)
    out = O.sparse_conv_c(feat, w, None, pairs, pn, oc.shape[0])
    dense = F.conv3d(_densify(coords, feat, B, shape), torch.from_numpy(w).permute(0, 4, 1, 2, 3), None,
                     stride=O._t3(stride), padding=O._t3(pad))
    assert list(dense.shape[2:]) == osh
    # active outputs = sites whose receptive field holds at least one active input
    occ = F.conv3d(_densify(coords, np.ones((n, 1), np.float32), B, shape), torch.ones((1, 1) + tuple(ks)),
                   None, stride=O._t3(stride), padding=O._t3(pad))
    active = (occ[:, 0] > 0).nonzero().numpy().astype(np.int32)
    assert np.array_equal(active, oc)          # same set, same (ascending) order
    c = torch.from_numpy(oc).long()
    ref = dense[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]].numpy()
    assert np.allclose(out, ref, rtol=1e-4, atol=1e-4)
    assert np.allclose(np.asarray(O.sparse_conv_torch(feat, w, None, pairs, pn, oc.shape[0])), ref, rtol=1e-4,
                       atol=1e-4)


def test_centre_only_weights_is_per_voxel_linear():
    rng = np.random.default_rng(2)
    coords, feat = _random_sparse(rng, 1, [6, 6, 6], 40, 5)
    w = np.zeros((3, 3, 3, 3, 5), np.float32)
    lin = rng.standard_normal((3, 5)).astype(np.float32)
    w[:, 1, 1, 1, :] = lin
    pairs, pn = O.rulebook_subm(coords, [6, 6, 6], 3)
    out = O.sparse_conv_c(feat, w, None, pairs, pn, 40)
    assert np.allclose(out, feat @ lin.T, rtol=1e-5, atol=1e-5)


def test_dense_and_height_compression_layout():
    feat = np.arange(2 * 3, dtype=np.float32).reshape(2, 3) + 1
    coords = np.array([[0, 1, 2, 3], [1, 0, 0, 1]], np.int32)
    d = O.dense(feat, coords, 2, [2, 4, 5])
    assert d.shape == (2, 3, 2, 4, 5) and d.sum() == feat.sum()
    assert d[0, 2, 1, 2, 3] == feat[0, 2] and d[1, 0, 0, 0, 1] == feat[1, 0]
    hc = O.height_compression(feat, coords, 2, [2, 4, 5])
    assert hc.shape == (2, 6, 4, 5)
    assert hc[0, 2 * 2 + 1, 2, 3] == feat[0, 2]          # channel = c*D + z (height_compression.py:21)


def test_backbone_oracle_runs_and_matches_shapes():
    """CPU restatement of VoxelResBackBone8x on a tiny grid (module only supplies the weights)."""
    from unidistill.layers.blocks_3d.det3d import VoxelResBackBone8x

    torch.manual_seed(0)
    rng = np.random.default_rng(3)
    grid = np.array([32, 32, 40])
    net = VoxelResBackBone8x(5, grid)
    coords, feat = _random_sparse(rng, 2, [41, 32, 32], 1500, 5)
    f, c, shp = O.backbone_forward(net.state_dict(), feat, coords, 2, [41, 32, 32], training=True)
    assert shp == [2, 4, 4]
    assert f.shape[1] == 128 and f.shape[0] == c.shape[0] > 0 and torch.isfinite(f).all()
