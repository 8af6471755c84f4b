"""mmdet3d / mmcv-style aliases of the voxelizer (BASELINE.json north_star: "keeping the unidistill.layers / mmdet3d
operator signatures").  The reference never calls these (it goes through spconv's PointToVoxel,
unidistill/data/det3d/preprocess/voxelization.py:31-38); they exist so that code written against
`mmdet3d.ops.Voxelization` / `mmcv.ops.voxelize.hard_voxelize` can use the same sm_100a kernel.

    hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points, max_voxels,
                  NDim=3, deterministic=True) -> voxel_num
        writes caller-allocated voxels (max_voxels, max_points, C), coors (max_voxels, 3) int32 [z, y, x],
        num_points_per_voxel (max_voxels,) int32; returns the number of voxels produced.
    Voxelization(voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True)
        forward(points (N, C)) -> (voxels (M, P, C), coors (M, 3) int32 [z, y, x], num_points (M,))
Our kernel is always deterministic (first-come order), so `deterministic=False` is accepted and ignored.
"""
import torch
from torch import nn

from .voxelize import hard_voxelize as _hard_voxelize


def _grid(voxel_size, coors_range):
    return [int(round((float(coors_range[3 + i]) - float(coors_range[i])) / float(voxel_size[i]))) for i in range(3)]


def hard_voxelize(points, voxels, coors, num_points_per_voxel, voxel_size, coors_range, max_points=35,
                  max_voxels=20000, NDim=3, deterministic=True):
    if NDim != 3:
        raise NotImplementedError("hard_voxelize: NDim must be 3")
    vs = [float(v) for v in voxel_size]
    cr = [float(v) for v in coors_range]
    v, c, n, _, _ = _hard_voxelize([points.contiguous().float()], vs, cr, _grid(vs, cr), int(max_points),
                                   int(max_voxels), want_voxels=True, want_mean=False)
    m = int(c.shape[0])
    voxels[:m].copy_(v)
    coors[:m].copy_(c[:, 1:])
    num_points_per_voxel[:m].copy_(n)
    return m


class Voxelization(nn.Module):
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000, deterministic=True):
        super().__init__()
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.max_num_points = int(max_num_points)
        self.max_voxels = max_voxels if isinstance(max_voxels, (tuple, list)) else (max_voxels, max_voxels)
        self.deterministic = deterministic
        self.grid_size = _grid(self.voxel_size, self.point_cloud_range)

    def forward(self, points):
        max_voxels = self.max_voxels[0] if self.training else self.max_voxels[1]
        v, c, n, _, _ = _hard_voxelize([points.contiguous().float()], self.voxel_size, self.point_cloud_range,
                                       self.grid_size, self.max_num_points, int(max_voxels), want_voxels=True,
                                       want_mean=False)
        return v, c[:, 1:].contiguous(), n
