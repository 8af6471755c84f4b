"""Pins the CPU oracle of the camera branch (oracle/camera_ref.c, oracle/camera.py) with known answers
(SURVEY 8c (iv), quirk 4); the reference ships the pooling kernel as a binary only."""
import numpy as np
import torch

from oracle import camera as OC


def test_all_points_into_one_cell_is_column_sum():
    rng = np.random.default_rng(0)
    feats = rng.standard_normal((1, 50, 8)).astype(np.float32)
    geom = np.tile(np.array([[3, 2, 0]], np.int32), (1, 50, 1))
    out, pos = OC.voxel_pooling(geom, feats, [5, 4, 1])
    assert out.shape == (1, 8, 4, 5)
    ref = np.zeros(8, np.float32)
    for p in range(50):
        ref = ref + feats[0, p]
    assert np.array_equal(out[0, :, 2, 3], ref)
    assert np.count_nonzero(out) == 8
    assert (pos == np.array([0, 2, 3])).all()


def test_out_of_range_points_dropped_and_memo_minus_one():
    feats = np.ones((2, 4, 3), np.float32)
    geom = np.array([[[0, 0, 0], [-1, 0, 0], [0, 4, 0], [0, 0, 1]],
                     [[4, 3, 0], [5, 3, 0], [2, 2, 0], [2, 2, -1]]], np.int32)
    out, pos = OC.voxel_pooling(geom, feats, [5, 4, 1])
    assert out.sum() == 3 * 3
    assert pos[0].tolist() == [[0, 0, 0], [-1, -1, -1], [-1, -1, -1], [-1, -1, -1]]
    assert pos[1].tolist() == [[1, 3, 4], [-1, -1, -1], [1, 2, 2], [-1, -1, -1]]
    g = np.arange(2 * 3 * 4 * 5, dtype=np.float32).reshape(2, 3, 4, 5)
    gi = OC.voxel_pooling_backward(g, pos, 3)
    assert np.array_equal(gi[1, 0], g[1, :, 3, 4]) and np.array_equal(gi[1, 2], g[1, :, 2, 2])
    assert not gi[0, 1:].any() and not gi[1, 1].any()


def test_int_truncation_quirk():
    # values in (-1, 0) truncate to cell 0 (reference lss_fpn.py:311-313 uses .int(), not floor)
    coord = torch.tensor([-50.8, -50.8, -1.0])
    size = torch.tensor([0.8, 0.8, 8.0])
    geom = torch.tensor([[-51.5, -51.3, -5.5], [-52.1, 0.0, 0.0]])
    idx = OC.voxel_indices(geom, coord, size)
    assert idx[0].tolist() == [0, 0, 0]
    assert idx[1, 0].item() == -1 and idx[1, 1].item() == 64


def test_fused_equals_materialised_pipeline():
    torch.manual_seed(0)
    B, ncam, D, H, W, C = 2, 3, 5, 4, 6, 8
    depth = torch.randn(B * ncam, D, H, W).softmax(1)
    ctx = torch.randn(B * ncam, C, H, W)
    geom = torch.randint(-1, 8, (B, ncam, D, H, W, 3), dtype=torch.int32)
    geom[..., 2] = torch.randint(-1, 2, geom[..., 2].shape, dtype=torch.int32)
    a = OC.lift_splat_reference(depth, ctx, geom.numpy(), [7, 6, 1])
    b = OC.lift_splat_fused(depth.numpy(), ctx.numpy(), geom.numpy(), [7, 6, 1])
    assert a.shape == (B, C, 6, 7)
    assert np.allclose(a, b, rtol=1e-6, atol=1e-6)


def test_geometry_matches_module_and_hits_the_map():
    """oracle geometry == mirrored LSSFPN.get_geometry (same torch ops); synthetic rig covers the BEV map."""
    from unidistill_b200 import synthetic as S

    mats = S.camera_rig(1, seed=0)
    fr = OC.create_frustum((256, 704), 16, [2.0, 58.0, 0.5])
    assert fr.shape == (112, 16, 44, 4)
    g = OC.get_geometry(fr, mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                        mats["bda_mat"])
    assert g.shape == (1, 6, 112, 16, 44, 3)
    idx = OC.voxel_indices(g, torch.tensor([-50.8, -50.8, -1.0]), torch.tensor([0.8, 0.8, 8.0]))
    ok = ((idx[..., 0] >= 0) & (idx[..., 0] < 128) & (idx[..., 1] >= 0) & (idx[..., 1] < 128)
          & (idx[..., 2] >= 0) & (idx[..., 2] < 1))
    frac = ok.float().mean().item()
    assert 0.5 < frac < 0.99, frac
