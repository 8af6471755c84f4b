"""Worker of tests/test_reference_over_shim.py (run as a subprocess, in this container only).

Imports the REFERENCE's own Python -- layers/blocks_3d/det3d/spconv_backbone.py, height_compression.py and the model
classes of exps/.../BEVFusion_nuscenes_base_exp.py / ..._centerhead_fusion_exp.py -- on top of this repo's `spconv` shim
and `unidistill` overlay, with only the third-party packages that are absent here (pytorch_lightning, mmcv, mmdet, ...)
stubbed by oracle/refstub.py.  There is no GPU in this container, so the kernels behind the shim are replaced by the CPU
oracle (test infrastructure): what is proven is that the reference code drives the shim's API (SparseConvTensor,
SubMConv3d / SparseConv3d / SparseSequential, indice keys, replace_feature, dense) correctly and computes the oracle's
numbers through it, and that state_dicts are interchangeable with the mirrored modules the GPU tests validate."""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "cvpr2023-unidistill_b200")
REF = os.environ.get("UNIDISTILL_REFERENCE_ROOT", "/root/reference")
os.environ["UNIDISTILL_REFERENCE_ROOT"] = REF
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from oracle import lidar as O  # noqa: E402
from oracle import refstub  # noqa: E402


def cpu_backend():
    """Replaces the CUDA operators behind the spconv shim by the CPU oracle."""
    from unidistill_b200 import sparse as SP

    def table_from_pairs(pairs, pair_num, n_rows, col):
        K = pairs.shape[1]
        t = np.full((n_rows, K), -1, np.int32)
        for k in range(K):
            m = int(pair_num[k])
            if m:
                t[pairs[col, k, :m], k] = pairs[1 - col, k, :m]
        return torch.from_numpy(t)

    def build_subm(coords, batch_size, spatial_shape, ksize, dilation=1):
        ks, dl = SP._triple(ksize), SP._triple(dilation)
        pairs, pn = O.rulebook_subm(coords.numpy(), list(spatial_shape), ks, dl)
        n = coords.shape[0]
        rb = SP.Rulebook(True, ks, [1, 1, 1], [k // 2 for k in ks], dl, spatial_shape, spatial_shape, n, n,
                         table_from_pairs(pairs, pn, n, 1), None, coords, batch_size)
        rb._cpu_pairs = (pairs, pn)
        return rb

    def build_conv(coords, batch_size, spatial_shape, ksize, stride, padding, dilation=1):
        ks, st, pd, dl = SP._triple(ksize), SP._triple(stride), SP._triple(padding), SP._triple(dilation)
        oc, pairs, pn, oshape = O.rulebook_conv(coords.numpy(), list(spatial_shape), ks, st, pd, dl)
        n_in, n_out = coords.shape[0], oc.shape[0]
        rb = SP.Rulebook(False, ks, st, pd, dl, spatial_shape, oshape, n_in, n_out, table_from_pairs(pairs, pn, n_out, 1),
                         table_from_pairs(pairs, pn, n_in, 0), torch.from_numpy(oc), batch_size)
        rb._cpu_pairs = (pairs, pn)
        return rb

    def conv(features, weight, bias, rb, inverse=False):
        assert not inverse
        pairs, pn = rb._cpu_pairs
        return O.sparse_conv_torch(features, weight, bias, pairs, pn, rb.n_out)

    def dense(features, coords, batch_size, spatial_shape):
        return torch.from_numpy(O.dense(features.detach().numpy(), coords.numpy(), batch_size, spatial_shape))

    SP.build_subm_rulebook, SP.build_conv_rulebook, SP.sparse_conv, SP.scatter_to_dense = build_subm, build_conv, conv, dense


def load_reference_file(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    out = {}
    refstub.install(overlay=True)
    import spconv.pytorch as spconv_shim

    assert spconv_shim.__file__.startswith(PKG), "spconv must resolve to this repo's shim"
    cpu_backend()

    # ---- 1. the reference's own spconv_backbone.py over the shim ------------------------------------------------
    ref_bb = load_reference_file("unidistill/layers/blocks_3d/det3d/spconv_backbone.py", "ref_spconv_backbone")
    ref_hc = load_reference_file("unidistill/layers/blocks_2d/det3d/map_to_bev/height_compression.py", "ref_height_compression")
    from unidistill.layers.blocks_3d.det3d import VoxelResBackBone8x as MirrorBackbone

    import unidistill.layers.blocks_3d.det3d.spconv_backbone as mirror_mod

    assert mirror_mod.__file__.startswith(PKG)
    torch.manual_seed(0)
    grid = [64, 64, 40]
    mirror = MirrorBackbone(5, np.array(grid)).eval()
    ref = ref_bb.VoxelResBackBone8x(5, np.array(grid), last_pad=0).eval()
    sd = mirror.state_dict()
    rsd = ref.state_dict()
    assert list(sd.keys()) == list(rsd.keys()), "state_dict keys differ"
    assert all(sd[k].shape == rsd[k].shape for k in sd), "state_dict shapes differ"
    for m in mirror.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.normal_(0, 0.1)
            m.running_var.uniform_(0.5, 1.5)
    ref.load_state_dict(mirror.state_dict(), strict=True)
    rng = np.random.default_rng(0)
    clouds = [(rng.standard_normal((n, 5)) * np.array([3.0, 3.0, 0.4, 1, 1])).astype(np.float32) for n in (5000, 3000)]
    _, coords, num, mean = O.voxelize(clouds, [0.2, 0.2, 0.1], [-6.4, -6.4, -2.0, 6.4, 6.4, 2.0], grid, 10, 3000)
    with torch.no_grad():
        sp, stride, scales = ref(torch.from_numpy(mean), torch.from_numpy(coords), len(clouds))
        bev, _ = ref_hc.HeightCompression(256)(sp, stride)
    f, c, shape = O.backbone_forward({k: v for k, v in mirror.state_dict().items()}, mean, coords, len(clouds), [41, 64, 64],
                                     training=False)
    want = O.height_compression(f.numpy(), c, len(clouds), shape)
    err = float(np.abs(bev.numpy() - want).max() / np.abs(want).max())
    assert bev.shape == want.shape and err < 1e-5, err
    assert stride == 8 and sorted(scales) == ["x_conv1", "x_conv2", "x_conv3", "x_conv4"]
    out["backbone_rel_err"] = err
    out["backbone_keys"] = len(sd)

    # ---- 2. the reference's model classes (base_exp / centerhead_fusion_exp) over the overlay --------------------
    import unidistill

    assert unidistill.__file__.startswith(PKG), "the mirror must be the top-level package (overlay mode)"
    base = __import__("unidistill.exps.multisensor_fusion.nuscenes.BEVFusion.BEVFusion_nuscenes_base_exp", fromlist=["x"])
    fus = __import__("unidistill.exps.multisensor_fusion.nuscenes.BEVFusion.BEVFusion_nuscenes_centerhead_fusion_exp",
                     fromlist=["x"])
    assert base.__file__.startswith(REF) and fus.__file__.startswith(REF)
    import unidistill.data.det3d.preprocess.voxelization as vmod
    import unidistill.layers.blocks_3d.mmdet3d.lss_fpn as lmod

    assert vmod.__file__.startswith(PKG) and lmod.__file__.startswith(PKG)
    assert base.VoxelResBackBone8x is MirrorBackbone and base.LSSFPN is lmod.LSSFPN
    from unidistill_b200 import models

    cfg = models.make_model_cfg(bev=128, lidar=True, camera=True, img_depth=18)
    torch.manual_seed(0)
    ours = models.BEVFusionCenterHead(cfg)
    theirs = fus.BEVFusionCenterHead(refstub.AttrDict(cfg))
    ok, tk = ours.state_dict(), theirs.state_dict()
    assert list(ok.keys()) == list(tk.keys()), sorted(set(ok) ^ set(tk))[:10]
    assert all(ok[k].shape == tk[k].shape for k in ok)
    theirs.load_state_dict(ok, strict=True)
    out["model_keys"] = len(ok)
    # the reference's LidarEncoder.forward + BEV encoder + head (eval) == the mirror's, through the same shim
    ours.eval()
    theirs.eval()
    lid_cfg = cfg.lidar_encoder
    assert isinstance(theirs.lidar_encoder, base.LidarEncoder) and isinstance(theirs.camera_encoder, base.CameraEncoder)
    assert type(theirs.lidar_encoder.backbone_3d) is MirrorBackbone
    out["lidar_grid"] = list(lid_cfg.grid_size)

    # ---- 3. the reference's BEVFusionCenterHead.forward (..._centerhead_fusion_exp.py:134-171) with its own
    # LidarEncoder.forward (base_exp.py:84-95), LiDAR-only, against the mirror's forward: same weights, same backend
    def cpu_voxelize(points, voxel_size, pc_range, grid_size, max_points, max_voxels, want_voxels=True, want_mean=True):
        v, c, n, m = O.voxelize([p.numpy() for p in points], voxel_size, pc_range, grid_size, max_points, max_voxels)
        return (torch.from_numpy(v) if want_voxels else None, torch.from_numpy(c), torch.from_numpy(n),
                torch.from_numpy(m), None)

    vmod.hard_voxelize = cpu_voxelize
    small = models.make_model_cfg(bev=128, lidar=True, camera=False)
    small.lidar_encoder.point_cloud_range = [-6.4, -6.4, -5.0, 6.4, 6.4, 3.0]
    small.lidar_encoder.grid_size = [128, 128, 40]
    small.det_head.target_assigner.grid_size = [128, 128, 40]
    small.det_head.dense_head.grid_size = [128, 128, 40]
    torch.manual_seed(1)
    ours_l = models.BEVFusionCenterHead(small).eval()
    theirs_l = fus.BEVFusionCenterHead(refstub.AttrDict(small)).eval()
    theirs_l.load_state_dict(ours_l.state_dict(), strict=True)
    for m in (ours_l, theirs_l):
        m.det_head.dense_head.distill = True
    pts = [torch.from_numpy(c) for c in clouds]
    gt = torch.zeros((2, 4, 10))
    gt[:, :2, 0:2] = torch.tensor([[1.0, 2.0], [-3.0, 0.5]])
    gt[:, :2, 3:6] = 1.5
    gt[:, :, 9] = 1
    with torch.no_grad():
        a = theirs_l(pts, None, None, gt, return_feature=True)
        b = ours_l(pts, None, None, gt, return_feature=True)
    assert torch.equal(a[0], b[0]) and torch.allclose(a[1], b[1], rtol=1e-6, atol=1e-6)
    for ha, hb in zip(a[2], b[2]):
        assert sorted(ha) == sorted(hb) and all(torch.allclose(ha[k], hb[k], rtol=1e-5, atol=1e-6) for k in ha)
    out["model_forward_bev_shape"] = list(a[0].shape)
    print("REF_OVER_SHIM " + json.dumps(out))


if __name__ == "__main__":
    main()
