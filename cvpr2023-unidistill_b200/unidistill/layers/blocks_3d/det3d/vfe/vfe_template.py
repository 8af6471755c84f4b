from torch import nn


class VFETemplate(nn.Module):
    """Base class of voxel feature encoders (reference vfe/vfe_template.py)."""

    def get_output_feature_dim(self):
        raise NotImplementedError

    def forward(self, **kwargs):
        raise NotImplementedError
