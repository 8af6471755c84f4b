"""oracle/distill.py -- TEST INFRASTRUCTURE (checker / reported CPU baseline only).

torch-CPU restatement of the three distillation losses and their glue, following
exps/multisensor_fusion/nuscenes/BEVFusion/BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar.py:
  corners_nd / rotation_2d_reverse / center_to_corner_box2d   :25-97
  calculate_box_mask_gaussian, gaussian_radius, gaussian2D, draw_umich_gaussian   :100-178
  FeatureDistillLoss :196-245, BEVDistillLoss :248-323, ResponseDistillLoss :326-385
  training_step glue (valid prefix, corners in cell units)    :447-483
Pinned against golden vectors produced by the real reference functions (tests/golden/distill.npz,
tests/test_oracle_distill.py).  world_size == 1 here, so reduce_mean is the identity.
"""
import numpy as np
import torch
import torch.nn.functional as F


def center_to_corner_box2d(centers, dims, angles):
    """numpy, same dtype promotion as the reference: float32 inputs, float64 corner math -> (N,4,2) float64."""
    norm = np.stack(np.unravel_index(np.arange(4), [2, 2]), axis=1)[[0, 1, 3, 2]] - np.array((0.5, 0.5))
    corners = dims.reshape([-1, 1, 2]) * norm.reshape([1, 4, 2])
    s, c = np.sin(angles), np.cos(angles)
    rot_t = np.stack([[c, s], [-s, c]])
    corners = np.einsum("aij,jka->aik", corners, rot_t)
    return corners + centers.reshape([-1, 1, 2])


def valid_prefix(gt_boxes):
    B, M = gt_boxes.shape[:2]
    ind = torch.zeros((B, M))
    for i in range(B):
        cnt = M - 1
        while cnt > 0 and gt_boxes[i][cnt].sum() == 0:
            cnt -= 1
        ind[i][: cnt + 1] = 1
    return ind.bool()


def bev_corners(gt_boxes, pc_range, voxel_size, osf):
    """(B,M,>=7) -> (B,M,4,2) float32 corners in BEV cells (:466-483)."""
    B, M = gt_boxes.shape[:2]
    out = torch.zeros((B, M, 4, 2))
    g = gt_boxes.detach().cpu().numpy()
    for i in range(B):
        out[i] = torch.from_numpy(center_to_corner_box2d(g[i][:, :2], g[i][:, 3:5], g[i][:, 6]))
    out[:, :, :, 0] = (out[:, :, :, 0] - pc_range[0]) / (voxel_size[0] * osf)
    out[:, :, :, 1] = (out[:, :, :, 1] - pc_range[1]) / (voxel_size[1] * osf)
    return out


def _keypoint_grid(coords, h, w):
    c = coords
    pts = torch.cat((c, c.mean(dim=2, keepdim=True), c[:, :, [0, 1]].mean(dim=2, keepdim=True),
                     c[:, :, [1, 2]].mean(dim=2, keepdim=True), c[:, :, [2, 3]].mean(dim=2, keepdim=True),
                     c[:, :, [0, 3]].mean(dim=2, keepdim=True)), dim=2).clone()
    pts[..., 0] = (pts[..., 0] - w / 2) / (w / 2)
    pts[..., 1] = (pts[..., 1] - h / 2) / (h / 2)
    return pts[..., [1, 0]]          # the column swap of :226 (quirk 5)


def _sample(fmap, grid):
    return F.grid_sample(fmap, grid, mode="bilinear", padding_mode="zeros", align_corners=False).permute(0, 2, 3, 1)


def feature_distill_loss(student, teacher, coords, indices):
    h, w = student.shape[-2:]
    grid = _keypoint_grid(coords, h, w)
    diff = (_sample(student, grid)[indices] - _sample(teacher, grid)[indices]).abs()
    loss = diff.mean(2).mean(1).sum()
    return loss / (indices.float().sum() + 1e-4)


def gram_pair(student, teacher, coords):
    """(B, M, 9, 9) Gram matrices of the normalised key-point features of student and teacher."""
    h, w = student.shape[-2:]
    grid = _keypoint_grid(coords, h, w)

    def gram(fmap):
        s = _sample(fmap, grid)
        s = s / (torch.norm(s, dim=-1, keepdim=True) + 1e-4)
        return s @ s.transpose(-1, -2)

    return gram(student), gram(teacher)


def well_conditioned_boxes(student, teacher, coords, indices, tol=1e-5):
    """Boxes whose relation-loss gradient is well defined: the L1 between Gram entries has a sign() in its
    derivative, so an off-diagonal |G_s - G_t| below fp32 round-off makes the gradient a coin flip (any two
    fp32 implementations, and the reference on two devices, disagree there).  Evaluated in float64."""
    gs, gt = gram_pair(student.double(), teacher.double(), coords.double())
    d = (gs - gt).abs()
    off = ~torch.eye(d.shape[-1], dtype=torch.bool)
    risky = ((d < tol) & (d > 0) & off).flatten(2).any(-1)
    return indices & ~risky


def bev_distill_loss(student, teacher, coords, indices):
    h, w = student.shape[-2:]
    grid = _keypoint_grid(coords, h, w)

    def gram(fmap):
        s = _sample(fmap, grid)                                   # (B, M, 9, C)
        s = s / (torch.norm(s, dim=-1, keepdim=True) + 1e-4)
        return s @ s.transpose(-1, -2)                            # (B, M, 9, 9)

    diff = (gram(student)[indices] - gram(teacher)[indices]).abs()
    return diff.mean(2).mean(1).sum() / (indices.float().sum() + 1e-4)


def gaussian_radius(size, min_overlap=0.7):
    height, width = size
    b1 = height + width
    c1 = width * height * (1 - min_overlap) / (1 + min_overlap)
    r1 = (b1 + np.sqrt(b1 ** 2 - 4 * c1)) / 2
    b2 = 2 * (height + width)
    c2 = (1 - min_overlap) * width * height
    r2 = (b2 + np.sqrt(b2 ** 2 - 16 * c2)) / 2
    a3 = 4 * min_overlap
    b3 = -2 * min_overlap * (height + width)
    c3 = (min_overlap - 1) * width * height
    r3 = (b3 + np.sqrt(b3 ** 2 - 4 * a3 * c3)) / 2
    return min(r1, r2, r3)


def box_mask_gaussian(shape_bhw, gt, pc_range, voxel_size, osf):
    """calculate_box_mask_gaussian (:100-128) incl. the python slicing semantics of draw_umich_gaussian."""
    B, H, W = shape_bhw
    mask = np.zeros((B, H, W), np.float32)
    gt = np.asarray(gt, np.float32)
    for i in range(B):
        for j in range(gt.shape[1]):
            if gt[i][j].sum() == 0:
                break
            w = gt[i][j][3] / (voxel_size[0] * osf)
            h = gt[i][j][4] / (voxel_size[1] * osf)
            r = max(0, int(gaussian_radius((w, h))))
            x = int((gt[i][j][0] - pc_range[0]) / (voxel_size[0] * osf))
            y = int((gt[i][j][1] - pc_range[1]) / (voxel_size[1] * osf))
            d = 2 * r + 1
            yy, xx = np.ogrid[-r: r + 1, -r: r + 1]
            g = np.exp(-(xx * xx + yy * yy) / (2 * (d / 6) * (d / 6)))
            g[g < np.finfo(g.dtype).eps * g.max()] = 0
            left, right = min(x, r), min(W - x, r + 1)
            top, bottom = min(y, r), min(H - y, r + 1)
            mh = mask[i][y - top: y + bottom, x - left: x + right]
            mg = g[r - top: r + bottom, r - left: r + right]
            if min(mg.shape) > 0 and min(mh.shape) > 0:
                np.maximum(mh, mg, out=mh)
    return torch.from_numpy(mask)


_REG_ORDER = ("reg", "height", "dim", "rot", "vel", "iou")


def response_distill_loss(resp_s, resp_t, gt_boxes, pc_range, voxel_size, osf, clamp_min=1e-4):
    cls_s = torch.cat([t["hm"] for t in resp_s], dim=1)
    cls_t = torch.cat([torch.clamp((t["hm"] / 2).sigmoid(), min=clamp_min, max=1 - clamp_min) for t in resp_t], dim=1)
    reg_s = torch.cat([t[k] for t in resp_s for k in _REG_ORDER], dim=1)
    reg_t = torch.cat([t[k] for t in resp_t for k in _REG_ORDER], dim=1)
    mask = box_mask_gaussian((reg_s.shape[0], reg_s.shape[2], reg_s.shape[3]), gt_boxes.detach().cpu().numpy(),
                             pc_range, voxel_size, osf)
    d_reg = (reg_s - reg_t).abs().mean(dim=1) * mask
    d_cls = (cls_s.max(dim=1)[0] - cls_t.max(dim=1)[0]).abs() * mask
    weight = mask.sum()
    return d_cls.sum() / (weight + 1e-4), d_reg.sum() / (weight + 1e-4)
