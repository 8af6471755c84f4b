"""GPU parity of the glue that decides indices and losses: LSS geometry -> `.int()` BEV cell index
(reference layers/blocks_3d/mmdet3d/lss_fpn.py:200-240,311-313) and CenterHeadIouAware.get_loss
(layers/head/det3d/center_head_iou_aware.py:55-298), both against fixtures produced by the REAL reference Python
(tests/golden/make_golden.py) -- computed here on the B200 through the mirrored modules the product uses."""
import numpy as np
import pytest
import torch

from test_golden_mirror import _head, _load

pytestmark = pytest.mark.gpu

# Borderline rule for truncated indices.  The reference computes geometry with fp32 inverses / batched matmuls whose
# rounding depends on the BLAS behind them (its own CPU and GPU runs differ in the last bits), then truncates.  An index
# may therefore differ from the fixture ONLY where the fixture's continuous cell coordinate lies within CELL_EPS of an
# integer (the point sits on a cell boundary to within 0.4 mm at the 0.8 m pitch); everywhere else it must be bit-exact.
CELL_EPS = 5e-4


def test_lss_geometry_and_cell_index_on_gpu(cuda):
    from unidistill.layers.blocks_3d.mmdet3d.lss_fpn import LSSFPN

    g = _load("lss_geometry.npz")
    lss = LSSFPN([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8], [2.0, 58.0, 0.5], (64, 176), 16, 256,
                 dict(type="ResNet", depth=18), dict(type="SECONDFPN", in_channels=[64, 128, 256, 512],
                                                     upsample_strides=[0.25, 0.5, 1, 2], out_channels=[128] * 4),
                 dict(in_channels=512, mid_channels=512)).to(cuda)
    mats = {k: torch.from_numpy(g[k]).to(cuda) for k in ("sensor2ego_mats", "intrin_mats", "ida_mats", "bda_mat")}
    geom = lss.get_geometry(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"])
    ref_geom = g["geom"]
    err = np.abs(geom.cpu().numpy() - ref_geom).max() / np.abs(ref_geom).max()
    assert err < 1e-5, "geometry rel err %g" % err
    idx = ((geom - (lss.voxel_coord - lss.voxel_size / 2.0)) / lss.voxel_size).int().cpu().numpy()
    lo = g["voxel_coord"] - g["voxel_size"] / 2.0
    cont = (ref_geom - lo) / g["voxel_size"]                               # the fixture's continuous cell coordinate
    borderline = np.abs(cont - np.rint(cont)) <= CELL_EPS
    diff = idx != g["idx"]
    assert not (diff & ~borderline).any(), "%d non-borderline indices differ" % int((diff & ~borderline).sum())
    assert diff.mean() < 1e-4, "fraction of differing (borderline) indices %g" % diff.mean()
    # a differing index is off by exactly one cell
    assert np.abs(idx.astype(np.int64) - g["idx"])[diff].max(initial=1) == 1


@pytest.mark.parametrize("with_bda", [True, False])
def test_fused_geometry_kernel_matches_reference(cuda, native_lib, with_bda):
    """ud_lss_geometry (frustum -> ego -> truncated voxel index in one kernel) against the reference fixture: float
    geometry to 1e-5, indices bit-exact away from cell boundaries (same borderline rule), and against the torch op chain
    of the mirrored get_geometry on the same device."""
    from unidistill.layers.blocks_3d.mmdet3d.lss_fpn import LSSFPN
    from unidistill_b200.camera import lss_geometry_index

    g = _load("lss_geometry.npz")
    lss = LSSFPN([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8], [2.0, 58.0, 0.5], (64, 176), 16, 256,
                 dict(type="ResNet", depth=18), dict(type="SECONDFPN", in_channels=[64, 128, 256, 512],
                                                     upsample_strides=[0.25, 0.5, 1, 2], out_channels=[128] * 4),
                 dict(in_channels=512, mid_channels=512)).to(cuda)
    mats = {k: torch.from_numpy(g[k]).to(cuda) for k in ("sensor2ego_mats", "intrin_mats", "ida_mats", "bda_mat")}
    args = (mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
            mats["bda_mat"] if with_bda else None)
    idx, geom = lss_geometry_index(lss.frustum, *args, lss.voxel_coord, lss.voxel_size, want_geom=True)
    ref_geom_t = lss.get_geometry(*args)
    ref_idx_t = ((ref_geom_t - (lss.voxel_coord - lss.voxel_size / 2.0)) / lss.voxel_size).int()
    assert idx.shape == ref_idx_t.shape and idx.dtype == torch.int32
    scale = float(ref_geom_t.abs().max())
    assert float((geom - ref_geom_t).abs().max()) / scale < 1e-5
    lo = (lss.voxel_coord - lss.voxel_size / 2.0)
    cont = (ref_geom_t - lo) / lss.voxel_size
    borderline = (cont - cont.round()).abs() <= CELL_EPS
    diff = idx != ref_idx_t
    assert not bool((diff & ~borderline).any()) and float(diff.float().mean()) < 1e-4
    if with_bda:                                                           # the committed reference fixture
        err = np.abs(geom.cpu().numpy() - g["geom"]).max() / np.abs(g["geom"]).max()
        assert err < 1e-5, err
        lo_np = g["voxel_coord"] - g["voxel_size"] / 2.0
        cont_np = (g["geom"] - lo_np) / g["voxel_size"]
        bl = np.abs(cont_np - np.rint(cont_np)) <= CELL_EPS
        d = idx.cpu().numpy() != g["idx"]
        assert not (d & ~bl).any() and d.mean() < 1e-4
    # the module uses the kernel: same BEV map as with the torch geometry (up to the borderline points)
    assert lss.fuse_geometry


def test_get_loss_on_gpu_matches_reference(cuda, native_lib):
    """Head forward + target assignment (fused kernel) + get_loss on the GPU against the reference's values."""
    g = _load("head_loss.npz")
    head = _head(g).to(cuda).train()
    x, gt = torch.from_numpy(g["x"]).to(cuda), torch.from_numpy(g["gt_boxes"]).to(cuda)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                                # fp32 convs: compare at the fixture's 1e-4
    try:
        ret = head(x, gt)
        for _, enc in ret["box_encoding"].items():
            enc.masked_fill_(torch.isinf(enc), 0)
        for t, d in enumerate(ret["multi_head_features"]):
            for k, v in d.items():
                # cuDNN's fp32 algorithms (Winograd / split-k) reorder the 3x3 sums: 1e-4 absolute on O(1..10) maps
                assert np.allclose(v.detach().cpu().numpy(), g["%s_%d" % (k, t)], rtol=1e-4, atol=1e-4), (k, t)
        loss, tb = head.get_loss(ret)
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    TOL = 1e-3                                                             # north_star: fp32 losses within 1e-3 relative
    assert abs(float(loss) - float(g["loss"])) <= TOL * abs(float(g["loss"])), (float(loss), float(g["loss"]))
    for t, d in enumerate(ret["multi_head_features"]):
        assert np.allclose(d["hm"].detach().cpu().numpy(), g["hm_sig_%d" % t], rtol=1e-4, atol=1e-5)   # quirk 6
        for name in ("loss", "hm_loss", "loc_loss", "num_positive"):
            ref = float(g["tb.task_%d.%s" % (t, name)])
            assert abs(float(tb["task_%d/%s" % (t, name)]) - ref) <= TOL * max(1.0, abs(ref)), (t, name)
    # pre-reduced normalisers (the training step's single fused all-reduce) give the same loss
    ret2 = head(x, gt)
    for _, enc in ret2["box_encoding"].items():
        enc.masked_fill_(torch.isinf(enc), 0)
    ret2["num_pos_reduced"] = torch.stack([ret2["mask"][t].float().sum() for t in range(6)])
    ret2["num_hm_reduced"] = torch.stack([ret2["heatmap"][t].eq(1).float().sum() for t in range(6)])
    loss2, _ = head.get_loss(ret2)
    assert abs(float(loss2) - float(g["loss"])) <= TOL * abs(float(g["loss"]))
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in head.parameters() if p.grad is not None)
