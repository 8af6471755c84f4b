"""The mirrored torch-level modules (assigner, head + loss, BEV backbone, MeanVFE, LSS geometry) reproduce
golden vectors produced by the REAL reference Python (tests/golden/make_golden.py, run where
/root/reference exists).  CPU-only: these modules are cuDNN/ATen consumers, not hand kernels."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CLASS_NAMES = ["car", "truck", "construction_vehicle", "bus", "trailer", "barrier", "motorcycle", "bicycle",
               "pedestrian", "traffic_cone"]
TASKS = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["truck", "construction_vehicle"]),
         dict(num_class=2, class_names=["bus", "trailer"]), dict(num_class=1, class_names=["barrier"]),
         dict(num_class=2, class_names=["motorcycle", "bicycle"]),
         dict(num_class=2, class_names=["pedestrian", "traffic_cone"])]
PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
VOXEL = [0.4, 0.4, 0.2]
GRID = [256, 256, 40]


def _load(name):
    return np.load(os.path.join(GOLD, name), allow_pickle=False)


def _assigner(max_objs=500):
    from unidistill.layers.head.det3d import FCOSAssigner

    return FCOSAssigner(out_size_factor=8, tasks=TASKS, dense_reg=1, gaussian_overlap=0.1, max_objs=max_objs,
                        min_radius=2, mapping={n: i + 1 for i, n in enumerate(CLASS_NAMES)}, grid_size=GRID,
                        pc_range=PC_RANGE[0:2], voxel_size=VOXEL[0:2], assign_topk=9, no_log=False,
                        with_velocity=True)


def test_fcos_assigner_matches_reference():
    g = _load("assigner.npz")
    tgt = _assigner().assign_targets(torch.from_numpy(g["gt_boxes"]))
    for t in range(6):
        assert np.array_equal(tgt["heatmap"][t].numpy(), g["heatmap_%d" % t]), "heatmap task %d" % t
        assert np.array_equal(tgt["mask"][t].numpy(), g["mask_%d" % t])
        assert np.array_equal(tgt["ind"][t].numpy(), g["ind_%d" % t])
        assert np.array_equal(tgt["cat"][t].numpy(), g["cat_%d" % t])
        a, b = tgt["box_encoding"][t].numpy(), g["box_encoding_%d" % t]
        assert a.shape == b.shape
        finite = np.isfinite(b)
        assert np.array_equal(np.isfinite(a), finite)            # same -inf pattern (log 0 of degenerate boxes)
        assert np.allclose(a[finite], b[finite], rtol=1e-6, atol=1e-6)


def test_assigner_input_not_mutated_and_empty_batch():
    g = _load("assigner.npz")
    gt = torch.from_numpy(g["gt_boxes"]).clone()
    keep = gt.clone()
    _assigner().assign_targets(gt)
    assert torch.equal(gt, keep)
    empty = torch.zeros((2, 5, 10))
    empty[:, :, 9] = 1
    tgt = _assigner().assign_targets(empty)
    # quirk 8: row 0 of an empty sample still counts as one (degenerate) car box
    assert int(tgt["mask"][0].sum()) == 2 * 9 and int(tgt["mask"][1].sum()) == 0


def _head(g):
    from unidistill.layers.head.det3d import CenterHeadIouAware, IouAwareGenProposals
    from unidistill.layers.losses.det3d import CenterNetRegLoss, FocalLoss

    proposal = IouAwareGenProposals(
        dataset_name="nuscenes", class_names=[t["class_names"] for t in TASKS],
        post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
        pc_range=PC_RANGE[0:2], out_size_factor=8, voxel_size=VOXEL[0:2], no_log=False,
        iou_aware_list=[0.65] * 10, nms_iou_threshold_train=0.8, nms_pre_max_size_train=1500,
        nms_post_max_size_train=80, nms_iou_threshold_test=0.1, nms_pre_max_size_test=1500,
        nms_post_max_size_test=100)
    head = CenterHeadIouAware(
        dataset_name="nuscenes", tasks=TASKS, target_assigner=_assigner(), proposal_layer=proposal,
        out_size_factor=8, input_channels=24, grid_size=GRID, point_cloud_range=PC_RANGE,
        code_weights=[1.0] * 8 + [0.2, 0.2], loc_weight=0.25, iou_weight=5.0, share_conv_channel=16,
        common_heads={"iou": [1, 2], "reg": [2, 2], "height": [1, 2], "dim": [3, 2], "rot": [2, 2], "vel": [2, 2]},
        upsample_for_pedestrian=False, mode="3d", init_bias=-2.19, predict_boxes_when_training=False)
    head.add_module("crit", FocalLoss(0.25, 2))
    head.add_module("crit_reg", CenterNetRegLoss())
    head.add_module("crit_iou_aware", CenterNetRegLoss())
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}
    missing, unexpected = head.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    return head


def test_center_head_forward_and_loss_match_reference():
    g = _load("head_loss.npz")
    head = _head(g).train()
    ret = head(torch.from_numpy(g["x"]), torch.from_numpy(g["gt_boxes"]))
    for _, enc in ret["box_encoding"].items():
        enc[torch.isinf(enc)] = 0
    for t, d in enumerate(ret["multi_head_features"]):
        for k, v in d.items():
            assert np.allclose(v.detach().numpy(), g["%s_%d" % (k, t)], rtol=1e-4, atol=1e-5), (k, t)
    loss, tb = head.get_loss(ret)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    for t, d in enumerate(ret["multi_head_features"]):
        # quirk 6: get_loss leaves clamp(sigmoid(hm)) in the dict
        assert np.allclose(d["hm"].detach().numpy(), g["hm_sig_%d" % t], rtol=1e-5, atol=1e-6)
        for name in ("loss", "hm_loss", "loc_loss", "num_positive"):
            ref = float(g["tb.task_%d.%s" % (t, name)])
            assert abs(float(tb["task_%d/%s" % (t, name)]) - ref) <= 1e-4 * max(1.0, abs(ref)), (t, name)
    loss.backward()
    assert all(torch.isfinite(p.grad).all() for p in head.parameters() if p.grad is not None)


def test_head_loss_with_prereduced_normalisers_is_identical():
    g = _load("head_loss.npz")
    head = _head(g).train()
    ret = head(torch.from_numpy(g["x"]), torch.from_numpy(g["gt_boxes"]))
    for _, enc in ret["box_encoding"].items():
        enc[torch.isinf(enc)] = 0
    ret["num_pos_reduced"] = torch.stack([ret["mask"][t].float().sum() for t in range(6)])
    loss, _ = head.get_loss(ret)
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))


def test_batched_head_loss_matches_reference_and_the_per_task_loop(monkeypatch):
    """With pre-reduced normalisers get_loss runs once over a task dimension (_get_loss_batched): same loss as the
    reference fixture, same tb_dict, same side effects and the same gradients as the per-task loop."""
    from unidistill.layers.head.det3d import center_head_iou_aware as M

    g = _load("head_loss.npz")
    out = []
    for batched in (True, False):
        monkeypatch.setattr(M, "BATCH_LOSS", batched)
        head = _head(g).train()
        x = torch.from_numpy(g["x"]).clone().requires_grad_(True)
        ret = head(x, torch.from_numpy(g["gt_boxes"]))
        for _, enc in ret["box_encoding"].items():
            enc[torch.isinf(enc)] = 0
        ret["num_pos_reduced"] = torch.stack([ret["mask"][t].float().sum() for t in range(6)])
        ret["num_hm_reduced"] = torch.stack([ret["heatmap"][t].eq(1).float().sum() for t in range(6)])
        assert head._can_batch_loss(ret) == batched
        loss, tb = head.get_loss(ret)
        loss.backward()
        out.append((loss, tb, ret, x.grad, {n: p.grad for n, p in head.named_parameters()}))
    (lb, tbb, rb, gxb, gpb), (ll, tbl, rl, gxl, gpl) = out
    assert abs(float(lb) - float(g["loss"])) <= 1e-4 * abs(float(g["loss"]))
    assert abs(float(lb) - float(ll)) <= 1e-5 * abs(float(ll))
    assert set(tbb) == set(tbl)
    for k in tbl:
        assert tbb[k].shape == tbl[k].shape and torch.allclose(tbb[k], tbl[k], rtol=1e-5, atol=1e-6), k
    for t in range(6):
        assert np.allclose(rb["multi_head_features"][t]["hm"].detach().numpy(), g["hm_sig_%d" % t], rtol=1e-5, atol=1e-6)
        assert torch.allclose(rb["pred_box_encoding"][t], rl["pred_box_encoding"][t])
    assert torch.allclose(gxb, gxl, rtol=1e-4, atol=1e-7)
    for n in gpl:
        if gpl[n] is None:
            assert gpb[n] is None, n
        else:
            assert torch.allclose(gpb[n], gpl[n], rtol=1e-4, atol=1e-6), n


def test_bev_backbone_and_mean_vfe_match_reference():
    from unidistill.layers.blocks_2d.det3d import BaseBEVBackbone
    from unidistill.layers.blocks_3d.det3d import MeanVFE

    g = _load("dense_blocks.npz")
    net = BaseBEVBackbone([1, 2], [1, 2], [8, 16], [1, 2], [8, 8], 6).eval()
    net.load_state_dict({k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd.")}, strict=True)
    with torch.no_grad():
        y, pyr = net(torch.from_numpy(g["x"]))
    assert np.allclose(y.numpy(), g["y"], rtol=1e-5, atol=1e-6)
    assert sorted(pyr.keys()) == list(g["pyr_keys"])
    mean = MeanVFE(5)(torch.from_numpy(g["vox"]), torch.from_numpy(g["num"]))
    assert np.allclose(mean.numpy(), g["mean"], rtol=1e-6, atol=1e-7)
    # oracle restatement agrees too
    from oracle import lidar as O

    assert np.allclose(O.mean_vfe(g["vox"], g["num"]), g["mean"], rtol=1e-6, atol=1e-7)


def test_lss_geometry_matches_reference():
    from oracle import camera as OC
    from unidistill.layers.blocks_3d.mmdet3d.lss_fpn import LSSFPN

    g = _load("lss_geometry.npz")
    lss = LSSFPN([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8], [2.0, 58.0, 0.5], (64, 176), 16, 256,
                 dict(type="ResNet", depth=18), dict(type="SECONDFPN", in_channels=[64, 128, 256, 512],
                                                     upsample_strides=[0.25, 0.5, 1, 2], out_channels=[128] * 4),
                 dict(in_channels=512, mid_channels=512))
    assert np.array_equal(lss.frustum.numpy(), g["frustum"])
    assert np.array_equal(lss.voxel_num.numpy(), g["voxel_num"])
    assert np.array_equal(lss.voxel_size.numpy(), g["voxel_size"]) and np.array_equal(lss.voxel_coord.numpy(),
                                                                                      g["voxel_coord"])
    mats = {k: torch.from_numpy(g[k]) for k in ("sensor2ego_mats", "intrin_mats", "ida_mats", "bda_mat")}
    geom = lss.get_geometry(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"])
    assert np.allclose(geom.numpy(), g["geom"], rtol=1e-5, atol=1e-4)
    idx = ((geom - (lss.voxel_coord - lss.voxel_size / 2.0)) / lss.voxel_size).int()
    # same ops in the same order on the same host BLAS: indices agree except for points that sit within
    # float noise of a cell boundary
    assert (idx.numpy() != g["idx"]).mean() < 1e-4
    fr = OC.create_frustum((64, 176), 16, [2.0, 58.0, 0.5])
    og = OC.get_geometry(fr, mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                         mats["bda_mat"])
    assert np.allclose(og.numpy(), g["geom"], rtol=1e-5, atol=1e-4)
    oi = OC.voxel_indices(og, torch.from_numpy(g["voxel_coord"]), torch.from_numpy(g["voxel_size"]))
    assert (oi.numpy() != g["idx"]).mean() < 1e-4


def test_default_centernet_focal_loss_matches_reference():
    """CenterHead.build_losses() installs CenterNetFocalLoss (reference layers/losses/det3d.py:325-360): Gaussian
    targets, no alpha, (1 - gt)^4 negative weights -- a different function from FocalLoss(0.25, 2)."""
    from unidistill.layers.losses.det3d import CenterNetFocalLoss, FocalLoss

    g = _load("focal_default.npz")
    pred, gt = torch.from_numpy(g["pred"]), torch.from_numpy(g["gt"])
    crit = CenterNetFocalLoss()
    assert abs(float(crit(pred, gt)) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    assert abs(float(crit(pred[1:, 1:2], gt[1:, 1:2])) - float(g["loss_nopos"])) <= 1e-6 * abs(float(g["loss_nopos"]))
    # pre-reduced normaliser path gives the same value
    assert abs(float(crit(pred, gt, torch.tensor(2.0))) - float(g["loss"])) <= 1e-6 * abs(float(g["loss"]))
    f = FocalLoss(0.25, 2)(pred, (gt == 1).float())
    assert abs(float(f) - float(g["loss_focal_025_2"])) <= 1e-6 * abs(float(g["loss_focal_025_2"]))
    assert abs(float(f) - float(g["loss"])) > 1e-3            # the two criteria are not interchangeable


def test_iou_branch_cannot_poison_the_gradient():
    """Random-init heads can predict log-sizes whose exp() overflows.  The reference keeps the IoU loss out of the graph
    unless loc_loss < 1 (Python `if loc_loss.item() < 1`, center_head_iou_aware.py:131-133); the sync-free torch.where of the
    mirror always builds the branch, so its exp() must not overflow (0 * inf = NaN in ExpBackward): gradients stay finite
    and equal those of the loss without the IoU term."""
    g = _load("head_loss.npz")
    head = _head(g).train()
    with torch.no_grad():
        for task in head.tasks:
            task.dim[3].bias.fill_(120.0)                     # exp(120) = inf in fp32
    x, gt = torch.from_numpy(g["x"]), torch.from_numpy(g["gt_boxes"])
    ret = head(x, gt)
    for _, enc in ret["box_encoding"].items():
        enc[torch.isinf(enc)] = 0
    loss, tb = head.get_loss(ret)
    assert torch.isfinite(loss) and any(float(tb["task_%d/loc_loss" % t]) >= 1 for t in range(6))
    loss.backward()
    grads = [p.grad for p in head.parameters() if p.grad is not None]
    assert grads and all(bool(torch.isfinite(gr).all()) for gr in grads)
    # tasks whose localisation loss is >= 1 contribute `loss` WITHOUT the IoU term, the others with it (reference :131-133)
    want = 0.0
    for t in range(6):
        part = head.auto_loss(tb["task_%d/hm_loss" % t], tb["task_%d/loc_loss" % t], tb["task_%d/iou_aware_loss" % t])
        if float(tb["task_%d/loc_loss" % t]) < 1:
            part = part + tb["task_%d/iou_loss" % t] * head.iou_weight
        want = want + float(part)
    assert abs(float(loss) - want) <= 1e-5 * abs(want)
