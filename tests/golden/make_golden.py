"""Generates tests/golden/*.npz by running the REAL reference Python (from /root/reference, imported under
the third-party stubs of oracle/refstub.py) on small seeded inputs.  Run in this container only:

    python tests/golden/make_golden.py

The fixtures pin (a) the oracle restatements in oracle/*.py and (b) the mirrored modules, on the GPU box
where /root/reference does not exist.  Inputs are stored next to the outputs, so tests never need the
generator's RNG.
"""
import math
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import refstub  # noqa: E402

AttrDict = refstub.install()

CLASS_NAMES = ["car", "truck", "construction_vehicle", "bus", "trailer", "barrier", "motorcycle", "bicycle",
               "pedestrian", "traffic_cone"]
TASKS = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["truck", "construction_vehicle"]),
         dict(num_class=2, class_names=["bus", "trailer"]), dict(num_class=1, class_names=["barrier"]),
         dict(num_class=2, class_names=["motorcycle", "bicycle"]),
         dict(num_class=2, class_names=["pedestrian", "traffic_cone"])]
PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]
VOXEL = [0.4, 0.4, 0.2]          # coarse grid so that the BEV map is 32 x 32 at stride 8
GRID = [256, 256, 40]
OSF = 8


def np_(t):
    return t.detach().cpu().numpy()


def make_gt(seed, B=3, M=12):
    """(B, M, 10) GT rows [x,y,z,dx,dy,dz,yaw,vx,vy,class(1-based)], zero padded; sample 2 is empty."""
    rng = np.random.default_rng(seed)
    gt = np.zeros((B, M, 10), np.float32)
    for b in range(B - 1):
        n = 9 - 2 * b
        gt[b, :n, 0:2] = (rng.random((n, 2)) * 2 - 1) * 48
        gt[b, :n, 2] = rng.normal(-1, 0.3, n)
        gt[b, :n, 3:6] = rng.random((n, 3)) * np.array([6, 3, 3]) + 0.4
        gt[b, :n, 6] = (rng.random(n) * 2 - 1) * 3.2
        gt[b, :n, 7:9] = rng.normal(0, 2, (n, 2))
        gt[b, :n, 9] = rng.integers(1, 11, n)
        if b == 0:
            gt[b, n - 1, 0:2] = [52.0, 3.0]      # centre outside the map (quirk 13)
            gt[b, 3] = 0                         # an all-zero row inside the valid prefix
            gt[b, 3, 9] = 1
    # padded rows carry label 0 + 1 = 1 in the reference training step (labels += 1, quirk 7)
    for b in range(B):
        pad = gt[b, :, :9].sum(1) == 0
        gt[b, pad, 9] = 1
    return gt


def golden_assigner_and_head():
    from unidistill.layers.head.det3d import CenterHeadIouAware, FCOSAssigner, IouAwareGenProposals
    from unidistill.layers.losses.det3d import CenterNetRegLoss, FocalLoss

    tasks = [AttrDict(t) for t in TASKS]
    assigner = FCOSAssigner(out_size_factor=OSF, tasks=tasks, dense_reg=1, gaussian_overlap=0.1, max_objs=500,
                            min_radius=2, mapping={n: i + 1 for i, n in enumerate(CLASS_NAMES)}, grid_size=GRID,
                            pc_range=PC_RANGE[0:2], voxel_size=VOXEL[0:2], assign_topk=9, no_log=False,
                            with_velocity=True)
    gt = torch.from_numpy(make_gt(0))
    tgt = assigner.assign_targets(gt.clone())
    out = {"gt_boxes": np_(gt)}
    for key in ("heatmap", "ind", "mask", "cat", "box_encoding"):
        for t in range(6):
            out["%s_%d" % (key, t)] = np_(tgt[key][t])
    np.savez_compressed(os.path.join(HERE, "assigner.npz"), **out)

    proposal = IouAwareGenProposals(
        dataset_name="nuscenes", class_names=[t["class_names"] for t in TASKS],
        post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
        pc_range=PC_RANGE[0:2], out_size_factor=OSF, voxel_size=VOXEL[0:2], no_log=False,
        iou_aware_list=[0.65] * 10, nms_iou_threshold_train=0.8, nms_pre_max_size_train=1500,
        nms_post_max_size_train=80, nms_iou_threshold_test=0.1, nms_pre_max_size_test=1500,
        nms_post_max_size_test=100)
    torch.manual_seed(0)
    head = CenterHeadIouAware(
        dataset_name="nuscenes", tasks=tasks, target_assigner=assigner, proposal_layer=proposal,
        out_size_factor=OSF, input_channels=24, grid_size=GRID, point_cloud_range=PC_RANGE,
        code_weights=[1.0] * 8 + [0.2, 0.2], loc_weight=0.25, iou_weight=5.0, share_conv_channel=16,
        common_heads={"iou": [1, 2], "reg": [2, 2], "height": [1, 2], "dim": [3, 2], "rot": [2, 2], "vel": [2, 2]},
        upsample_for_pedestrian=False, mode="3d", init_bias=-2.19, predict_boxes_when_training=False)
    head.add_module("crit", FocalLoss(0.25, 2))
    head.add_module("crit_reg", CenterNetRegLoss())
    head.add_module("crit_iou_aware", CenterNetRegLoss())
    with torch.no_grad():
        head.auto_loss.params.copy_(torch.linspace(0.8, 1.3, 12))
    head.train()
    x = torch.randn(3, 24, 32, 32)
    ret = head(x, gt.clone())
    for _, enc in ret["box_encoding"].items():
        enc[torch.isinf(enc)] = 0
    raw = {"%s_%d" % (k, t): np_(v) for t, d in enumerate(ret["multi_head_features"]) for k, v in d.items()}
    loss, tb = head.get_loss(ret)
    hm_after = {"hm_sig_%d" % t: np_(d["hm"]) for t, d in enumerate(ret["multi_head_features"])}
    sd = {"sd." + k: np_(v) for k, v in head.state_dict().items()}
    scal = {"tb." + k.replace("/", "."): np.float32(float(v)) for k, v in tb.items()}
    np.savez_compressed(os.path.join(HERE, "head_loss.npz"), x=np_(x), gt_boxes=np_(gt), loss=np_(loss), **raw,
                        **hm_after, **sd, **scal)


def golden_distill():
    m = refstub.load_distill_module("BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar")
    torch.manual_seed(1)
    gt = torch.from_numpy(make_gt(1))
    B, M = gt.shape[0], gt.shape[1]
    # --- the glue of training_step (..._distill_lidar.py:447-483), executed verbatim on the reference fns
    indice = torch.zeros((B, M))
    for i in range(B):
        cnt = M - 1
        while cnt > 0 and gt[i][cnt, :9].sum() == 0:
            cnt -= 1
        indice[i][: cnt + 1] = 1
    indice = indice.bool()
    coords = torch.zeros((B, M, 4, 2))
    for i in range(B):
        coords[i] = m.center_to_corner_box2d(gt[i][:, :2].numpy(), gt[i][:, 3:5].numpy(), gt[i][:, 6].numpy(),
                                             origin=(0.5, 0.5))
    corners_metric = coords.clone()
    coords[:, :, :, 0] = (coords[:, :, :, 0] - PC_RANGE[0]) / (VOXEL[0] * OSF)
    coords[:, :, :, 1] = (coords[:, :, :, 1] - PC_RANGE[1]) / (VOXEL[1] * OSF)

    fs = torch.randn(B, 16, 32, 32, requires_grad=True)
    ft = torch.randn(B, 16, 32, 32)
    bs_ = torch.randn(B, 24, 32, 32, requires_grad=True)
    bt = torch.randn(B, 24, 32, 32)
    loss_feat = m.FeatureDistillLoss(fs, ft, coords.clone(), indice)
    loss_rel = m.BEVDistillLoss(bs_, bt, coords.clone(), indice)
    g_fs, = torch.autograd.grad(loss_feat, fs)
    g_bs, = torch.autograd.grad(loss_rel, bs_)

    ncls = [1, 2, 2, 1, 2, 2]
    heads = {"reg": 2, "height": 1, "dim": 3, "rot": 2, "vel": 2, "iou": 1}
    resp_s, resp_t, store = [], [], {}
    for t in range(6):
        ds, dt = {}, {}
        for k, c in list(heads.items()) + [("hm", ncls[t])]:
            a = torch.randn(B, c, 32, 32)
            b = torch.randn(B, c, 32, 32)
            if k == "hm":   # the student's hm is already clamp(sigmoid) when the distill loss reads it (quirk 6)
                a = torch.clamp(a.sigmoid(), min=1e-4, max=1 - 1e-4)
            a.requires_grad_(True)
            ds[k], dt[k] = a, b
            store["rs_%s_%d" % (k, t)] = np_(a)
            store["rt_%s_%d" % (k, t)] = np_(b)
        resp_s.append(ds)
        resp_t.append(dt)
    loss_cls, loss_reg = m.ResponseDistillLoss(resp_s, resp_t, gt, PC_RANGE, VOXEL, OSF)
    leaves = [resp_s[t][k] for t in range(6) for k in list(heads) + ["hm"]]
    grads = torch.autograd.grad(loss_cls + loss_reg, leaves, allow_unused=True)
    for (t, k), g in zip([(t, k) for t in range(6) for k in list(heads) + ["hm"]], grads):
        store["grs_%s_%d" % (k, t)] = np_(g) if g is not None else np.zeros(1, np.float32)
    mask = m.calculate_box_mask_gaussian((B, 66, 32, 32), gt.numpy(), PC_RANGE, VOXEL, OSF)
    radii = np.array([m.gaussian_radius((w, h)) for w, h in [(1.0, 1.0), (5.6, 2.4), (0.3, 9.0), (14.0, 3.5)]],
                     np.float64)
    np.savez_compressed(
        os.path.join(HERE, "distill.npz"), gt_boxes=np_(gt), indice=np_(indice), coords=np_(coords),
        corners_metric=np_(corners_metric), fs=np_(fs), ft=np_(ft), bs=np_(bs_), bt=np_(bt),
        loss_feat=np_(loss_feat), loss_rel=np_(loss_rel), g_fs=np_(g_fs), g_bs=np_(g_bs),
        loss_cls=np_(loss_cls), loss_reg=np_(loss_reg), gaussian_mask=np_(mask), radii=radii,
        pc_range=np.array(PC_RANGE, np.float32), voxel=np.array(VOXEL, np.float32), **store)


def golden_dense_blocks():
    from unidistill.layers.blocks_2d.det3d import BaseBEVBackbone
    from unidistill.layers.blocks_3d.det3d import MeanVFE
    from unidistill.layers.blocks_3d.mmdet3d.lss_fpn import LSSFPN

    torch.manual_seed(2)
    net = BaseBEVBackbone([1, 2], [1, 2], [8, 16], [1, 2], [8, 8], 6)
    net.eval()
    x = torch.randn(2, 6, 16, 16)
    with torch.no_grad():
        y, pyr = net(x)
    sd = {"sd." + k: np_(v) for k, v in net.state_dict().items()}
    vox = torch.randn(7, 10, 5)
    num = torch.tensor([1, 10, 3, 0, 5, 2, 7])
    for i, n in enumerate(num.tolist()):
        vox[i, n:] = 0
    mean = MeanVFE(5)(vox, num)
    np.savez_compressed(os.path.join(HERE, "dense_blocks.npz"), x=np_(x), y=np_(y), vox=np_(vox), num=np_(num),
                        mean=np_(mean), pyr_keys=np.array(sorted(pyr.keys())), **sd)

    # LSS geometry on a reduced image (64 x 176 -> 4 x 11 feature map), full depth range
    sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "ud_synth", os.path.join(ROOT, "cvpr2023-unidistill_b200", "unidistill_b200", "synthetic.py"))
    synth = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(synth)
    lss = LSSFPN([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8], [2.0, 58.0, 0.5], (64, 176), 16, 256,
                 dict(type="ResNet"), dict(type="SECONDFPN"), dict(in_channels=512, mid_channels=512))
    mats = synth.camera_rig(2, seed=5, final_dim=(64, 176), random_bda=True)
    geom = lss.get_geometry(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats["bda_mat"])
    idx = ((geom - (lss.voxel_coord - lss.voxel_size / 2.0)) / lss.voxel_size).int()
    np.savez_compressed(os.path.join(HERE, "lss_geometry.npz"), frustum=np_(lss.frustum), geom=np_(geom),
                        idx=np_(idx), voxel_size=np_(lss.voxel_size), voxel_coord=np_(lss.voxel_coord),
                        voxel_num=np_(lss.voxel_num), **{k: np_(v) for k, v in mats.items()})


def golden_proposals():
    """IouAwareGenProposals.generate_predicted_boxes of the REAL reference (top-k, decoding, IoU-aware ranking, range /
    score masks, per-sample NMS calls, roi assembly) on seeded head outputs.  The binary-only `iou3d_nms_cuda.nms_gpu`
    it calls is replaced by the CPU restatement oracle/nms_ref.c (so this fixture pins everything around the NMS against
    the reference, and the NMS itself against the oracle)."""
    import importlib

    from oracle import nms as ON

    nms_mod = importlib.import_module("unidistill.layers.head.det3d.generate_proposals.iou3d_nms_cuda")

    def nms_gpu(boxes, keep, thresh):
        k = ON.nms_bev(np_(boxes), float(thresh))
        keep[: len(k)] = torch.from_numpy(k)
        return len(k)

    nms_mod.nms_gpu = nms_gpu
    from unidistill.layers.head.det3d.generate_proposals import IouAwareGenProposals

    cfg = dict(dataset_name="nuscenes", class_names=[t["class_names"] for t in TASKS],
               post_center_limit_range=[-61.2, -61.2, -10.0, 61.2, 61.2, 10.0], score_threshold=0.1,
               pc_range=PC_RANGE[0:2], out_size_factor=OSF, voxel_size=VOXEL[0:2], no_log=False,
               iou_aware_list=[0.65, 0.5, 0.65, 0.3, 0.65, 0.65], nms_iou_threshold_train=0.8,
               nms_pre_max_size_train=300, nms_post_max_size_train=40, nms_iou_threshold_test=0.1,
               nms_pre_max_size_test=300, nms_post_max_size_test=25)
    layer = IouAwareGenProposals(**cfg)
    g = torch.Generator().manual_seed(11)
    B, H, W = 2, 32, 32
    feats, out = [], {}
    for t, task in enumerate(TASKS):
        nc = task["num_class"]
        d = {"hm": torch.randn(B, nc, H, W, generator=g) * 1.5 - 1.5, "reg": torch.rand(B, 2, H, W, generator=g),
             "height": torch.randn(B, 1, H, W, generator=g) * 0.5 - 1.0,
             "dim": torch.randn(B, 3, H, W, generator=g) * 0.3 + torch.tensor([1.4, 0.6, 0.5]).view(1, 3, 1, 1),
             "rot": torch.randn(B, 2, H, W, generator=g), "vel": torch.randn(B, 2, H, W, generator=g),
             "iou": torch.randn(B, 1, H, W, generator=g)}
        if t == 3:
            d["hm"][1] = -9.0                     # a (task, sample) pair with no candidate above the score threshold
        feats.append(d)
        for k, v in d.items():
            out["in_%d_%s" % (t, k)] = np_(v)
    for mode in ("test", "train"):
        layer.training = mode == "train"
        res = layer.generate_predicted_boxes({"multi_head_features": [dict(d) for d in feats]}, {})
        for b, pd in enumerate(res["pred_dicts"]):
            for k, v in pd.items():
                out["%s_%s_%d" % (mode, k, b)] = np_(v)
        for k in ("rois", "roi_scores", "roi_labels"):
            out["%s_%s" % (mode, k)] = np_(res[k])
    out["cfg_iou_aware_list"] = np.asarray(cfg["iou_aware_list"], np.float32)
    np.savez_compressed(os.path.join(HERE, "proposals.npz"), **out)
    print("proposals: test-mode boxes per sample", [out["test_pred_boxes_%d" % b].shape[0] for b in range(B)],
          "train-mode", [out["train_pred_boxes_%d" % b].shape[0] for b in range(B)])


def _input_side_samples(seed=31, B=3):
    """Synthetic per-sample dicts in the layout the reference's dataset hands to its transforms."""
    rng = np.random.default_rng(seed)

    def pose(yaw, t):
        m = np.eye(4)
        c, s_ = np.cos(yaw), np.sin(yaw)
        m[:2, :2] = [[c, -s_], [s_, c]]
        m[:3, 3] = t
        return m

    samples = []
    for b in range(B):
        n_sw = [3, 0, 2][b % 3]
        key_e2g = pose(rng.uniform(-3, 3), rng.uniform(-200, 200, 3) * [1, 1, 0.01])
        key_l2e = pose(rng.uniform(-0.05, 0.05), [0.94, 0.0, 1.84])
        ts = 1.5e15 + 5e5 * b
        pts = (rng.standard_normal((900 + 100 * b, 5)) * [30, 30, 1.5, 40, 1]).astype(np.float32)
        sweeps, s_e2g, s_ts = [], [], []
        for k in range(n_sw):
            sweeps.append((rng.standard_normal((700 + 50 * k, 5)) * [30, 30, 1.5, 40, 1]).astype(np.float32))
            s_e2g.append(key_e2g @ pose(rng.uniform(-0.02, 0.02), [0.4 * (k + 1), 0.05 * k, 0.0]))
            s_ts.append(ts - 5e4 * (k + 1))
        samples.append(dict(points=pts, sweep_points=sweeps, key_e2g=key_e2g, key_l2e=key_l2e, s_e2g=s_e2g, ts=ts, s_ts=s_ts))
    return samples


def golden_input_side():
    """CollectLidarSweeps + ObjectRangeFilter (points part) of the REAL reference on synthetic sweeps."""
    from unidistill.data.multisensorfusion.transforms3d import CollectLidarSweeps, ObjectRangeFilter

    out = {}
    rng_filter = ObjectRangeFilter(PC_RANGE).eval()
    for b, s_ in enumerate(_input_side_samples()):
        d = {"points": s_["points"].copy(), "sweep_points": [x.copy() for x in s_["sweep_points"]],
             "info": {"ego_to_global": s_["key_e2g"], "lidar_to_ego": s_["key_l2e"], "timestamp": s_["ts"],
                      "sweep_lidar_infos": [dict(sweep_lidar_to_ego=m, sweep_lidar_timestamp=t)
                                            for m, t in zip(s_["s_e2g"], s_["s_ts"])]}}
        d = CollectLidarSweeps()(d)
        out["merged_%d" % b] = d["points"].astype(np.float32)
        d = rng_filter(d)
        out["filtered_%d" % b] = d["points"].astype(np.float32)
        out["in_points_%d" % b] = s_["points"]
        out["key_e2g_%d" % b], out["key_l2e_%d" % b], out["ts_%d" % b] = s_["key_e2g"], s_["key_l2e"], np.float64(s_["ts"])
        out["n_sweeps_%d" % b] = np.int64(len(s_["sweep_points"]))
        for k, (x, m, t) in enumerate(zip(s_["sweep_points"], s_["s_e2g"], s_["s_ts"])):
            out["sweep_%d_%d" % (b, k)], out["s_e2g_%d_%d" % (b, k)], out["s_ts_%d_%d" % (b, k)] = x, m, np.float64(t)
    np.savez_compressed(os.path.join(HERE, "input_side.npz"), **out)


def golden_focal_default():
    """The head's DEFAULT heat-map criterion, CenterNetFocalLoss (reference layers/losses/det3d.py:325-360), on a
    Gaussian (non-binary) target: no alpha, negatives = gt < 1 weighted by (1 - gt)^4, no epsilon in log(1 - pred)."""
    from unidistill.layers.losses.det3d import CenterNetFocalLoss, FocalLoss

    g = torch.Generator().manual_seed(21)
    pred = torch.clamp(torch.randn(2, 3, 16, 16, generator=g).sigmoid(), 1e-4, 1 - 1e-4)
    gt = torch.rand(2, 3, 16, 16, generator=g) ** 3
    gt[0, 0, 3, 4] = 1.0
    gt[1, 2, 9, 9] = 1.0
    gt[1, 1] = 0.0
    out = {"pred": np_(pred), "gt": np_(gt), "loss": np_(CenterNetFocalLoss()(pred, gt)),
           "loss_nopos": np_(CenterNetFocalLoss()(pred[1:, 1:2], gt[1:, 1:2])),
           "loss_focal_025_2": np_(FocalLoss(0.25, 2)(pred, (gt == 1).float()))}
    np.savez_compressed(os.path.join(HERE, "focal_default.npz"), **out)


def golden_iou3d():
    """Rotated-BEV IoU matrices computed by the REFERENCE BINARY's own CPU routine boxes_iou_bev_cpu (executed through
    oracle/refshim): pins oracle/nms_ref.c and, through it, the CUDA kernels."""
    from oracle import refshim

    rng = np.random.default_rng(2023)
    n = 260
    b = np.zeros((n, 7), np.float32)
    b[:, 0:2] = rng.uniform(-12, 12, (n, 2))
    b[:, 2] = rng.normal(-1, 0.5, n)
    b[:, 3] = rng.uniform(0.3, 7, n)
    b[:, 4] = rng.uniform(0.3, 3, n)
    b[:, 5] = rng.uniform(0.5, 3, n)
    b[:, 6] = rng.uniform(-3.3, 3.3, n)
    # special cases: identical boxes, the same rectangle described with a quarter / half turn, nested, edge-touching,
    # axis-aligned grids (many collinear edges), tiny and huge boxes, shared centres
    b[1] = b[0]
    b[2] = b[0]; b[2, 3], b[2, 4], b[2, 6] = b[0, 4], b[0, 3], b[0, 6] + np.float32(np.pi / 2)
    b[3] = b[0]; b[3, 6] = b[0, 6] + np.float32(np.pi)
    b[4] = b[0]; b[4, 3:5] *= 0.5
    b[5] = [0, 0, 0, 4, 2, 1, 0]; b[6] = [4, 0, 0, 4, 2, 1, 0]; b[7] = [2, 0, 0, 4, 2, 1, 0]; b[8] = [0, 2, 0, 4, 2, 1, 0]
    b[9] = [0, 0, 0, 2, 2, 1, np.pi / 4]; b[10] = [0, 0, 0, 4, 4, 1, 0]
    b[11] = [1, 1, 0, 0.05, 0.05, 1, 0.3]; b[12] = [1, 1, 0, 40, 30, 1, 0.3]
    for i in range(13, 33):
        b[i] = [float((i - 13) % 5) * 1.5, float((i - 13) // 5) * 1.0, 0, 2.0, 1.2, 1, 0]
    b[33:60, 0:2] = b[60:87, 0:2]
    iou = refshim.reference_boxes_iou_bev_cpu(b, b)
    a2 = b[:40].copy()
    a2[:, 0:2] += rng.normal(0, 0.3, (40, 2)).astype(np.float32)
    iou2 = refshim.reference_boxes_iou_bev_cpu(a2, b)
    np.savez_compressed(os.path.join(HERE, "iou3d_ref.npz"), boxes=b, iou=iou, boxes_a2=a2, iou_a2=iou2)
    print("iou3d_ref: %d x %d and %d x %d matrices from the reference binary; nonzero pairs %d" % (
        n, n, 40, n, int((iou > 0).sum())))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "iou3d":
        golden_iou3d()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "input_side":
        golden_input_side()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "focal":
        golden_focal_default()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "proposals":
        golden_proposals()
        sys.exit(0)
    golden_assigner_and_head()
    golden_distill()
    golden_dense_blocks()
    golden_proposals()
    golden_iou3d()
    golden_focal_default()
    golden_input_side()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
