"""MeanVFE (reference layers/blocks_3d/det3d/vfe/mean_vfe.py:6-34): mean of the points of each voxel.

When the voxels come from our Voxelization the mean was already produced inside the voxelize kernel
(attribute `_ud_mean`) and is returned as is; any other tensor takes the generic reduction."""
import torch

from .vfe_template import VFETemplate


class MeanVFE(VFETemplate):
    def __init__(self, num_point_features):
        super().__init__()
        self.num_point_features = num_point_features

    def get_output_feature_dim(self):
        return self.num_point_features

    def forward(self, voxel_features, voxel_num_points, **kwargs):
        fused = getattr(voxel_features, "_ud_mean", None)
        if fused is not None and fused.shape[1] == self.num_point_features \
                and fused.shape[0] == voxel_features.shape[0]:
            return fused
        total = voxel_features[:, :, : self.num_point_features].sum(dim=1)
        count = torch.clamp_min(voxel_num_points.view(-1, 1), min=1.0).type_as(voxel_features)
        return (total / count).contiguous()
