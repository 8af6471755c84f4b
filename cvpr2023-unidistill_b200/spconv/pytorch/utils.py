"""spconv.pytorch.utils.PointToVoxel -- signature kept (reference voxelization.py:31-38,54)."""
import torch

from unidistill_b200.voxelize import hard_voxelize


class PointToVoxel(object):
    """Hard voxelizer for one point cloud.

    __call__(pc (N,C) f32) -> (voxels (M,P,C), indices (M,3) int32 zyx, num_per_voxel (M,) int32).
    Deterministic first-come ordering (spconv's CPU semantics); see csrc/voxelize.cu.
    """

    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels,
                 max_num_points_per_voxel, device=torch.device("cuda")):
        self.vsize = [float(v) for v in vsize_xyz]
        self.coors_range = [float(v) for v in coors_range_xyz]
        self.num_point_features = int(num_point_features)
        self.max_num_voxels = int(max_num_voxels)
        self.max_num_points_per_voxel = int(max_num_points_per_voxel)
        self.device = device
        self.grid_size = [
            int(round((self.coors_range[3 + i] - self.coors_range[i]) / self.vsize[i])) for i in range(3)
        ]

    def __call__(self, pc, clear_voxels=True, empty_mean=False):
        if pc.shape[1] != self.num_point_features:
            raise RuntimeError("PointToVoxel: expected %d point features, got %d"
                               % (self.num_point_features, pc.shape[1]))
        voxels, coords, num, mean, _ = hard_voxelize(
            [pc], self.vsize, self.coors_range, self.grid_size, self.max_num_points_per_voxel,
            self.max_num_voxels, want_voxels=True, want_mean=True)
        voxels._ud_mean = mean
        return voxels, coords[:, 1:].contiguous(), num

    def generate_voxel_with_id(self, pc, clear_voxels=True, empty_mean=False):
        raise NotImplementedError("generate_voxel_with_id is not used by UniDistill")
