"""Axis-aligned BEV IoU helpers used by the IoU-aware head loss
(reference utils/det3d_utils/box_utils.py:318-373)."""
import numpy as np
import torch

from . import common_utils


def boxes_iou_normal(boxes_a, boxes_b):
    """(N,4),(M,4) [x1,y1,x2,y2] -> (N,M) IoU."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 4
    x_min = torch.max(boxes_a[:, 0, None], boxes_b[None, :, 0])
    x_max = torch.min(boxes_a[:, 2, None], boxes_b[None, :, 2])
    y_min = torch.max(boxes_a[:, 1, None], boxes_b[None, :, 1])
    y_max = torch.min(boxes_a[:, 3, None], boxes_b[None, :, 3])
    inter = torch.clamp_min(x_max - x_min, min=0) * torch.clamp_min(y_max - y_min, min=0)
    area_a = (boxes_a[:, 2] - boxes_a[:, 0]) * (boxes_a[:, 3] - boxes_a[:, 1])
    area_b = (boxes_b[:, 2] - boxes_b[:, 0]) * (boxes_b[:, 3] - boxes_b[:, 1])
    return inter / torch.clamp_min(area_a[:, None] + area_b[None, :] - inter, min=1e-6)


def boxes_iou_normal_paired(boxes_a, boxes_b):
    """Row-wise IoU of two (N,4) sets == diagonal of boxes_iou_normal without the N^2 matrix."""
    x_min = torch.max(boxes_a[:, 0], boxes_b[:, 0])
    x_max = torch.min(boxes_a[:, 2], boxes_b[:, 2])
    y_min = torch.max(boxes_a[:, 1], boxes_b[:, 1])
    y_max = torch.min(boxes_a[:, 3], boxes_b[:, 3])
    inter = torch.clamp_min(x_max - x_min, min=0) * torch.clamp_min(y_max - y_min, min=0)
    area_a = (boxes_a[:, 2] - boxes_a[:, 0]) * (boxes_a[:, 3] - boxes_a[:, 1])
    area_b = (boxes_b[:, 2] - boxes_b[:, 0]) * (boxes_b[:, 3] - boxes_b[:, 1])
    return inter / torch.clamp_min(area_a + area_b - inter, min=1e-6)


def boxes3d_lidar_to_aligned_bev_boxes(boxes3d):
    """(N,7+) [x,y,z,dx,dy,dz,heading] -> (N,4) nearest axis-aligned BEV box."""
    rot = common_utils.limit_period(boxes3d[:, 6], offset=0.5, period=np.pi).abs()
    dxy = boxes3d[:, 3:5]        # slices, not list indices: no host-built index tensor (CUDA-graph capturable)
    dims = torch.where(rot[:, None] < np.pi / 4, dxy, dxy.flip(1))
    return torch.cat((boxes3d[:, 0:2] - dims / 2, boxes3d[:, 0:2] + dims / 2), dim=1)


def boxes3d_nearest_bev_iou(boxes_a, boxes_b):
    return boxes_iou_normal(boxes3d_lidar_to_aligned_bev_boxes(boxes_a),
                            boxes3d_lidar_to_aligned_bev_boxes(boxes_b))


def boxes3d_nearest_bev_iou_paired(boxes_a, boxes_b):
    return boxes_iou_normal_paired(boxes3d_lidar_to_aligned_bev_boxes(boxes_a),
                                   boxes3d_lidar_to_aligned_bev_boxes(boxes_b))
