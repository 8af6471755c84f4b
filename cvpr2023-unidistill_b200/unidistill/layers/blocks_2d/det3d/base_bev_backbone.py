"""BaseBEVBackbone -- SECOND-style 2-D BEV encoder (reference layers/blocks_2d/det3d/base_bev_backbone.py).

Dense convolutions stay on cuDNN (not a hand-kernel target); the module tree is the
reference's (`blocks.<i>.<j>`, `deblocks.<i>.<j>`) so checkpoints load by key.  The self-calibrated
conv variant (use_scconv) is out of scope: every BEVFusion config sets it False.
"""
import numpy as np
import torch
import torch.nn as nn


def _bn(channels):
    return nn.BatchNorm2d(channels, eps=1e-3, momentum=0.01)


class BaseBEVBackbone(nn.Module):
    def __init__(self, layer_nums, layer_strides, num_filters, upsample_strides, num_upsample_filters,
                 input_channels, use_scconv=False, upsample_output=False):
        super().__init__()
        if use_scconv:
            raise NotImplementedError("SCBottleneck blocks are not used by the BEVFusion experiments")
        layer_nums = list(layer_nums or [])
        layer_strides = list(layer_strides or [])
        num_filters = list(num_filters or [])
        upsample_strides = list(upsample_strides or [])
        num_upsample_filters = list(num_upsample_filters or [])
        assert len(layer_nums) == len(layer_strides) == len(num_filters)
        assert len(upsample_strides) == len(num_upsample_filters)

        levels = len(layer_nums)
        in_ch = [input_channels] + num_filters[:-1]
        self.blocks = nn.ModuleList()
        self.deblocks = nn.ModuleList()
        for lvl in range(levels):
            width = num_filters[lvl]
            layers = [nn.ZeroPad2d(1),
                      nn.Conv2d(in_ch[lvl], width, 3, stride=layer_strides[lvl], padding=0, bias=False),
                      _bn(width), nn.ReLU()]
            for _ in range(layer_nums[lvl]):
                layers += [nn.Conv2d(width, width, 3, padding=1, bias=False), _bn(width), nn.ReLU()]
            self.blocks.append(nn.Sequential(*layers))
            if upsample_strides:
                s = upsample_strides[lvl]
                up = num_upsample_filters[lvl]
                if s >= 1:
                    conv = nn.ConvTranspose2d(width, up, s, stride=s, bias=False)
                else:
                    k = int(np.round(1 / s))
                    conv = nn.Conv2d(width, up, k, stride=k, bias=False)
                self.deblocks.append(nn.Sequential(conv, _bn(up), nn.ReLU()))

        c_in = sum(num_upsample_filters)
        if len(upsample_strides) > levels:
            s = upsample_strides[-1]
            self.deblocks.append(nn.Sequential(nn.ConvTranspose2d(c_in, c_in, s, stride=s, bias=False),
                                               _bn(c_in), nn.ReLU()))
        self.num_bev_features = c_in
        self.upsample_featuremap = upsample_output
        if upsample_output:
            self.upsample_conv = nn.Sequential(nn.ConvTranspose2d(c_in, c_in, 2, stride=2, bias=False),
                                               _bn(c_in), nn.ReLU())

    def forward(self, spatial_features):
        """(B, C, H, W) -> (fused (B, sum(up), H, W), {"spatial_features_<s>x": level output})"""
        ups, pyramid = [], {}
        x = spatial_features
        for lvl, block in enumerate(self.blocks):
            x = block(x)
            pyramid["spatial_features_%dx" % int(spatial_features.shape[2] / x.shape[2])] = x
            ups.append(self.deblocks[lvl](x) if len(self.deblocks) > 0 else x)
        if len(ups) > 1:
            x = torch.cat(ups, dim=1)
        elif len(ups) == 1:
            x = ups[0]
        if len(self.deblocks) > len(self.blocks):
            x = self.deblocks[-1](x)
        if self.upsample_featuremap:
            x = self.upsample_conv(x)
        return x, pyramid
