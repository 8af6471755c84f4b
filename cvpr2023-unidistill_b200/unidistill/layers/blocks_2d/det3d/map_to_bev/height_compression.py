"""HeightCompression (reference map_to_bev/height_compression.py:4-22): sparse -> (N, C*D, H, W).

`dense()` of our SparseConvTensor is one scatter kernel that writes the (N, C, D, H, W) buffer
directly, so the view below is free (channel index = c*D + z, exactly the reference's layout)."""
import torch.nn as nn


class HeightCompression(nn.Module):
    def __init__(self, num_bev_features):
        super().__init__()
        self.num_bev_features = num_bev_features

    def forward(self, encoded_spconv_tensor, encoded_spconv_tensor_stride):
        volume = encoded_spconv_tensor.dense()
        n, c, d, h, w = volume.shape
        return volume.view(n, c * d, h, w), encoded_spconv_tensor_stride
