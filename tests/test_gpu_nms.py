"""Rotated-BEV IoU / batched NMS kernels (csrc/nms.cu) and the mirrored proposal layer against the CPU oracle
(oracle/nms_ref.c) and the committed fixture produced by the REAL reference's IouAwareGenProposals."""
import os

import numpy as np
import pytest
import torch

from oracle import nms as ON

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proposals.npz")
TASKS = [["car"], ["truck", "construction_vehicle"], ["bus", "trailer"], ["barrier"], ["motorcycle", "bicycle"],
         ["pedestrian", "traffic_cone"]]


def _boxes(rng, n, spread=15.0):
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-spread, spread, (n, 2))
    b[:, 2] = rng.normal(-1, 0.3, n)
    b[:, 3] = rng.uniform(0.4, 6, n)
    b[:, 4] = rng.uniform(0.4, 3, n)
    b[:, 5] = 1.5
    b[:, 6] = rng.uniform(-3.2, 3.2, n)
    return b


def test_iou_bev_matches_oracle(cuda, native_lib):
    from unidistill_b200 import nms

    rng = np.random.default_rng(0)
    a, b = _boxes(rng, 257), _boxes(rng, 300)
    b[:40] = a[:40]                                          # identical pairs
    b[40:80, :2] = a[40:80, :2] + 0.01                       # near-identical pairs
    got = nms.boxes_iou_bev(torch.from_numpy(a).to(cuda), torch.from_numpy(b).to(cuda)).cpu().numpy()
    ref = ON.boxes_iou_bev(a, b)
    # same single-precision algorithm; sin/cos/atan2 of the two math libraries differ in the last ulp
    assert np.abs(got - ref).max() < 2e-5
    assert (got[ref == 0] == 0).all()
    assert nms.boxes_iou_bev(torch.from_numpy(a[:0]).to(cuda), torch.from_numpy(b).to(cuda)).shape == (0, 300)


def _ref_keep(boxes, n_valid, thresh, max_keep):
    return ON.nms_bev(boxes[:n_valid], thresh)[:max_keep]


@pytest.mark.parametrize("n,thresh,max_keep", [(1, 0.1, 8), (63, 0.1, 100), (64, 0.5, 64), (65, 0.1, 10), (1500, 0.1, 100),
                                               (1500, 0.8, 80), (700, 0.3, 700)])
def test_nms_matches_oracle(cuda, native_lib, n, thresh, max_keep):
    from unidistill_b200 import nms

    rng = np.random.default_rng(n)
    S = 5
    sets = np.stack([_boxes(rng, n, spread=4.0 + 6.0 * s) for s in range(S)])
    n_valid = np.array([n, max(0, n - 7), n // 2, 0, min(n, 3)], np.int32)
    keep, num = nms.nms_bev_batched(torch.from_numpy(sets).to(cuda), torch.from_numpy(n_valid).to(cuda), thresh, max_keep)
    keep, num = keep.cpu().numpy(), num.cpu().numpy()
    iou_gpu = [nms.boxes_iou_bev(torch.from_numpy(s).to(cuda), torch.from_numpy(s).to(cuda)).cpu().numpy() for s in sets]
    for s in range(S):
        ref = _ref_keep(sets[s], int(n_valid[s]), thresh, max_keep)
        got = keep[s, : num[s]]
        assert (keep[s, num[s]:] == -1).all()
        if not np.array_equal(got, ref):
            # the only legitimate difference: a pair whose IoU sits within float round-off of the threshold
            iou_ref = ON.boxes_iou_bev(sets[s], sets[s])
            border = np.abs(iou_ref - thresh) < 1e-4
            assert border.any(), "set %d: keep lists differ without a borderline pair" % s
            # then the kernel must at least be the exact greedy sweep of the GPU's own IoU matrix
            removed = np.zeros(n, bool)
            mine = []
            for i in range(int(n_valid[s])):
                if removed[i] or len(mine) >= max_keep:
                    continue
                mine.append(i)
                removed |= (iou_gpu[s][i] > thresh) & (np.arange(n) > i)
            assert got.tolist() == mine
    # without n_valid every row takes part
    keep2, num2 = nms.nms_bev_batched(torch.from_numpy(sets[:1]).to(cuda), None, thresh, max_keep)
    assert np.array_equal(keep2.cpu().numpy()[0, : int(num2[0])], keep[0, : num[0]])


def test_iou3d_nms_cuda_drop_in(cuda, native_lib):
    """The module the reference imports: nms_gpu(boxes, keep_cpu_long, thresh) -> num_out, boxes_iou_bev_gpu."""
    from unidistill.layers.head.det3d.generate_proposals import iou3d_nms_cuda

    rng = np.random.default_rng(5)
    b = _boxes(rng, 400, spread=8.0)
    keep = torch.LongTensor(400)
    num = iou3d_nms_cuda.nms_gpu(torch.from_numpy(b).to(cuda), keep, 0.1)
    assert keep[:num].tolist() == ON.nms_bev(b, 0.1).tolist()
    ans = torch.zeros((400, 400), device=cuda)
    iou3d_nms_cuda.boxes_iou_bev_gpu(torch.from_numpy(b).to(cuda), torch.from_numpy(b).to(cuda), ans)
    assert np.abs(ans.cpu().numpy() - ON.boxes_iou_bev(b, b)).max() < 2e-5
    with pytest.raises(RuntimeError):
        iou3d_nms_cuda.nms_gpu(torch.from_numpy(b), keep, 0.1)


def _layer():
    from unidistill.layers.head.det3d import IouAwareGenProposals

    return IouAwareGenProposals(
        dataset_name="nuscenes", class_names=TASKS, post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0],
        score_threshold=0.1, pc_range=[-51.2, -51.2], out_size_factor=8, voxel_size=[0.4, 0.4], no_log=False,
        iou_aware_list=[0.65, 0.5, 0.65, 0.3, 0.65, 0.65], nms_iou_threshold_train=0.8, nms_pre_max_size_train=300,
        nms_post_max_size_train=40, nms_iou_threshold_test=0.1, nms_pre_max_size_test=300, nms_post_max_size_test=25)


@pytest.mark.parametrize("mode", ["test", "train"])
def test_generate_predicted_boxes_matches_reference_golden(cuda, native_lib, mode):
    g = np.load(GOLDEN)
    feats = [{k: torch.from_numpy(g["in_%d_%s" % (t, k)]).to(cuda) for k in ("hm", "reg", "height", "dim", "rot", "vel", "iou")}
             for t in range(6)]
    layer = _layer()
    layer.training = mode == "train"
    res = layer.generate_predicted_boxes({"multi_head_features": feats}, {})
    assert res["has_class_labels"] is True
    for b in range(2):
        ref_boxes, ref_scores, ref_labels = (g["%s_pred_%s_%d" % (mode, k, b)] for k in ("boxes", "scores", "labels"))
        pd = res["pred_dicts"][b]
        assert pd["pred_boxes"].shape == ref_boxes.shape, (pd["pred_boxes"].shape, ref_boxes.shape)
        assert np.array_equal(pd["pred_labels"].cpu().numpy(), ref_labels)
        assert np.allclose(pd["pred_scores"].cpu().numpy(), ref_scores, rtol=1e-5, atol=1e-6)
        assert np.allclose(pd["pred_boxes"].cpu().numpy(), ref_boxes, rtol=1e-4, atol=1e-4)
    for k in ("rois", "roi_scores", "roi_labels"):
        assert np.allclose(res[k].cpu().numpy(), g["%s_%s" % (mode, k)], rtol=1e-4, atol=1e-4), k


def test_per_sample_proposal_layer_equals_batched_path(cuda, native_lib):
    """The reference-shaped per-sample path (proposal_layer -> nms_options -> iou3d_nms_cuda.nms_gpu) and the batched
    device-side path select the same boxes."""
    g = np.load(GOLDEN)
    feats = [{k: torch.from_numpy(g["in_%d_%s" % (t, k)]).to(cuda) for k in ("hm", "reg", "height", "dim", "rot", "vel", "iou")}
             for t in range(6)]
    layer = _layer()
    layer.training = False
    res = layer.generate_predicted_boxes({"multi_head_features": feats}, {})
    layer._use_sizes()
    per_task = []
    for t, pd in enumerate(feats):
        hm, rots, rotc, hei, dim, vel, reg, iou = layer._task_inputs(pd)
        per_task.append(layer.proposal_layer(hm, rots, rotc, hei, dim, vel, reg=reg, task_id=t, iouhm=iou))
    for b in range(2):
        boxes = torch.cat([per_task[t][b]["boxes"] for t in range(6)])
        scores = torch.cat([per_task[t][b]["scores"] for t in range(6)])
        assert torch.allclose(boxes, res["pred_dicts"][b]["pred_boxes"])
        assert torch.allclose(scores, res["pred_dicts"][b]["pred_scores"])


def test_eval_mode_model_forward_returns_predictions(cuda, native_lib):
    """model.eval(): forward ends in the proposal layer (reference CenterHead.forward eval branch, center_head.py:137-146)."""
    from unidistill_b200 import models, synthetic as S

    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    model = models.BEVFusionCenterHead(cfg).to(cuda).eval()
    batch = S.lidar_batch(2, seed=3, points_per_sweep=6000, sweeps=2)
    pts = [batch["points"][b].to(cuda) for b in range(2)]
    with torch.no_grad():
        out = model(pts, None, None, None)
    assert len(out["pred_dicts"]) == 2 and out["rois"].shape == (2, 600, 9)
    for pd in out["pred_dicts"]:
        n = pd["pred_boxes"].shape[0]
        assert pd["pred_boxes"].shape == (n, 9) and pd["pred_scores"].shape == (n,) and pd["pred_labels"].shape == (n,)
        assert n <= 600 and bool((pd["pred_scores"] > 0.1).all()) if n else True
