"""Fused distillation losses with the signatures of the functions the reference defines inside each
distill experiment script (BEVFusion_nuscenes_centerhead_*_exp_distill_*.py:196-385):

    FeatureDistillLoss(feature_lidar, feature_fuse, gt_boxes_bev_coords, gt_boxes_indices)
    BEVDistillLoss(bev_lidar, bev_fuse, gt_boxes_bev_coords, gt_boxes_indices)
    ResponseDistillLoss(resp_lidar, resp_fuse, gt_boxes, pc_range, voxel_size, out_size_scale)

(first argument = student, second = teacher, as in the scripts).  Opt-in from a script is one line:
`from unidistill.layers.losses.distill import FeatureDistillLoss, BEVDistillLoss, ResponseDistillLoss`.
Each loss is one fused kernel reading teacher and student once, plus its backward twin.
`weight` / `mask_weight`: optional pre-reduced normalisers (one fused all-reduce per step instead of three).
"""
import torch

from unidistill.utils.torch_dist import reduce_mean
from unidistill_b200 import distill as _d


def _box_loss(student, teacher, coords, indices, mode, weight=None):
    terms = _d.BoxDistill.apply(student, teacher, coords, indices, mode)
    if weight is None:
        weight = reduce_mean(indices.float().sum())
    return terms.sum() / (weight + 1e-4)


def FeatureDistillLoss(feature_lidar, feature_fuse, gt_boxes_bev_coords, gt_boxes_indices, weight=None):
    """L1 between student / teacher low-level BEV features at 9 key-points per GT box (reference :196-245)."""
    return _box_loss(feature_lidar, feature_fuse, gt_boxes_bev_coords, gt_boxes_indices, 0, weight)


def BEVDistillLoss(bev_lidar, bev_fuse, gt_boxes_bev_coords, gt_boxes_indices, weight=None):
    """L1 between the 9x9 Gram matrices of the L2-normalised key-point features (reference :248-323)."""
    return _box_loss(bev_lidar, bev_fuse, gt_boxes_bev_coords, gt_boxes_indices, 1, weight)


_REG_ORDER = ("reg", "height", "dim", "rot", "vel", "iou")


def ResponseDistillLoss(resp_lidar, resp_fuse, gt_boxes, pc_range, voxel_size, out_size_scale, clamp_min=1e-4,
                        mask_weight=None):
    """Gaussian-masked L1 between student / teacher head responses (reference :326-385).
    resp_lidar[t]["hm"] is the student's already clamp-sigmoid'ed heat-map (quirk 6); the teacher's logits get
    clamp(sigmoid(x / 2)) inside the kernel (clamp_min = 1e-3 in camera_exp_distill_fusion, quirk 11)."""
    cls_s = torch.cat([t["hm"] for t in resp_lidar], dim=1)
    cls_t = torch.cat([t["hm"] for t in resp_fuse], dim=1)
    reg_s = torch.cat([t[k] for t in resp_lidar for k in _REG_ORDER], dim=1)
    reg_t = torch.cat([t[k] for t in resp_fuse for k in _REG_ORDER], dim=1)
    n_rows = _d.rows_before_first_zero(gt_boxes)
    geom = (pc_range[0], pc_range[1], voxel_size[0] * out_size_scale, voxel_size[1] * out_size_scale, clamp_min)
    sums = _d.ResponseDistill.apply(cls_s, cls_t, reg_s, reg_t, gt_boxes, n_rows, geom)
    weight = reduce_mean(sums[2].detach()) if mask_weight is None else mask_weight
    return sums[0] / (weight + 1e-4), sums[1] / (weight + 1e-4)


def gt_boxes_bev_corners(gt_boxes, pc_range, voxel_size, out_size_factor):
    """Device replacement of the per-sample numpy `center_to_corner_box2d` hop + cell-unit conversion of the
    reference training_step (:466-483) -> (B, M, 4, 2)."""
    return _d.box_corners_bev(gt_boxes, pc_range, voxel_size, out_size_factor)


def gt_boxes_valid_prefix(gt_boxes):
    """(B, M) bool `gt_boxes_indice` of the reference training_step (:447-455), without host syncs."""
    return _d.valid_box_prefix(gt_boxes)
