"""Pins oracle/distill.py against golden vectors produced by the REAL reference functions
(tests/golden/distill.npz <- tests/golden/make_golden.py)."""
import os

import numpy as np
import torch

from oracle import distill as OD

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "distill.npz")
OSF = 8
_HEADS = ("reg", "height", "dim", "rot", "vel", "iou", "hm")


def _g():
    return np.load(GOLD)


def _resp(g, prefix):
    return [{k: torch.from_numpy(g["%s_%s_%d" % (prefix, k, t)]) for k in _HEADS} for t in range(6)]


def test_glue_valid_prefix_and_corners():
    g = _g()
    gt = torch.from_numpy(g["gt_boxes"])
    assert np.array_equal(OD.valid_prefix(gt[:, :, :9]).numpy(), g["indice"])
    c = OD.bev_corners(gt, g["pc_range"].tolist(), g["voxel"].tolist(), OSF)
    assert np.allclose(c.numpy(), g["coords"], rtol=1e-6, atol=1e-5)


def test_feature_and_relation_losses_and_grads():
    g = _g()
    coords, ind = torch.from_numpy(g["coords"]), torch.from_numpy(g["indice"])
    fs = torch.from_numpy(g["fs"]).requires_grad_(True)
    loss = OD.feature_distill_loss(fs, torch.from_numpy(g["ft"]), coords, ind)
    assert abs(float(loss) - float(g["loss_feat"])) < 1e-6 * max(1, abs(float(g["loss_feat"])))
    loss.backward()
    assert np.allclose(fs.grad.numpy(), g["g_fs"], rtol=1e-5, atol=1e-8)
    bs = torch.from_numpy(g["bs"]).requires_grad_(True)
    loss = OD.bev_distill_loss(bs, torch.from_numpy(g["bt"]), coords, ind)
    assert abs(float(loss) - float(g["loss_rel"])) < 1e-6 * max(1, abs(float(g["loss_rel"])))
    loss.backward()
    assert np.allclose(bs.grad.numpy(), g["g_bs"], rtol=1e-4, atol=1e-8)


def test_gaussian_mask_and_radius():
    g = _g()
    mask = OD.box_mask_gaussian((3, 32, 32), g["gt_boxes"], g["pc_range"].tolist(), g["voxel"].tolist(), OSF)
    assert np.array_equal(mask.numpy(), g["gaussian_mask"])
    # quirk 7: padded rows (label column = 1) draw a radius-0 Gaussian at the map centre
    assert mask[2, 16, 16] == 1.0
    radii = [OD.gaussian_radius(s) for s in [(1.0, 1.0), (5.6, 2.4), (0.3, 9.0), (14.0, 3.5)]]
    assert np.allclose(radii, g["radii"], rtol=1e-12)


def test_response_loss():
    g = _g()
    rs, rt = _resp(g, "rs"), _resp(g, "rt")
    lc, lr = OD.response_distill_loss(rs, rt, torch.from_numpy(g["gt_boxes"]), g["pc_range"].tolist(),
                                      g["voxel"].tolist(), OSF)
    assert abs(float(lc) - float(g["loss_cls"])) < 1e-6 and abs(float(lr) - float(g["loss_reg"])) < 1e-6
