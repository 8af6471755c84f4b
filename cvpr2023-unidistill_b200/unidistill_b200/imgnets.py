"""Image backbone / neck stand-ins for the mmdet / mmdet3d registries (absent on this image).

Only what the BEVFusion camera experiments build (reference BEVFusion_nuscenes_centerhead_fusion_exp.py:24-41):
ResNet-50 with out_indices [0..3] and SECONDFPN(in=[256,512,1024,2048], strides=[.25,.5,1,2], out=128x4).
Dense convolutions -> cuDNN; these are library consumers, not hand-kernel targets.
Module / parameter names follow mmdet's ResNet and mmdet3d's SECONDFPN so checkpoints load by key.
"""
import numpy as np
import torch
from torch import nn


class ResNet(nn.Module):
    def __init__(self, depth=50, out_indices=(0, 1, 2, 3), frozen_stages=-1, norm_eval=False, init_cfg=None,
                 **kwargs):
        super().__init__()
        import torchvision

        ctor = {18: torchvision.models.resnet18, 34: torchvision.models.resnet34,
                50: torchvision.models.resnet50, 101: torchvision.models.resnet101}[depth]
        net = ctor(weights=None)   # no network: random init instead of torchvision://resnet50
        self.conv1, self.bn1, self.relu, self.maxpool = net.conv1, net.bn1, net.relu, net.maxpool
        self.layer1, self.layer2, self.layer3, self.layer4 = net.layer1, net.layer2, net.layer3, net.layer4
        self.out_indices = tuple(out_indices)
        self.frozen_stages = frozen_stages
        self.norm_eval = norm_eval

    def init_weights(self):
        return None

    def forward(self, x):
        if isinstance(self.bn1, nn.Identity):      # frozen + folded: bias and ReLU in the convolution's epilogue
            from .fused_dense import conv_bias_relu

            x = self.maxpool(conv_bias_relu(self.conv1, x))
        else:
            x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        outs = []
        for i, layer in enumerate((self.layer1, self.layer2, self.layer3, self.layer4)):
            x = layer(x)
            if i in self.out_indices:
                outs.append(x)
        return tuple(outs)

    def train(self, mode=True):
        super().train(mode)
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm):
                    m.eval()
        return self


class SECONDFPN(nn.Module):
    def __init__(self, in_channels=(128, 128, 256), out_channels=(256, 256, 256), upsample_strides=(1, 2, 4),
                 use_conv_for_no_stride=False, **kwargs):
        super().__init__()
        assert len(in_channels) == len(out_channels) == len(upsample_strides)
        blocks = []
        for cin, cout, s in zip(in_channels, out_channels, upsample_strides):
            if s > 1 or (s == 1 and not use_conv_for_no_stride):
                s = int(s)
                conv = nn.ConvTranspose2d(cin, cout, kernel_size=s, stride=s, bias=False)
            else:
                k = int(np.round(1 / s))
                conv = nn.Conv2d(cin, cout, kernel_size=k, stride=k, bias=False)
            blocks.append(nn.Sequential(conv, nn.BatchNorm2d(cout, eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
        self.deblocks = nn.ModuleList(blocks)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def forward(self, xs):
        ups = [blk(x) for blk, x in zip(self.deblocks, xs)]
        return [torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]]


_BACKBONES = {"ResNet": ResNet}
_NECKS = {"SECONDFPN": SECONDFPN}


def build_backbone(conf):
    conf = dict(conf)
    kind = conf.pop("type")
    if kind not in _BACKBONES:
        raise NotImplementedError("image backbone %r needs mmdet; only %s are built in" % (kind, list(_BACKBONES)))
    return _BACKBONES[kind](**conf)


def build_neck(conf):
    conf = dict(conf)
    kind = conf.pop("type")
    if kind not in _NECKS:
        raise NotImplementedError("image neck %r needs mmdet3d; only %s are built in" % (kind, list(_NECKS)))
    return _NECKS[kind](**conf)
