"""Base classes of the multi-sensor models (reference models/multisensor_fusion/base.py:12-67)."""
from abc import ABCMeta

import torch.nn as nn
from torch.nn.modules.batchnorm import _BatchNorm


class BaseMultiSensorFusion(nn.Module):
    def _has(self, name):
        return hasattr(self, name) and getattr(self, name)

    @property
    def with_lidar_encoder(self):
        return self._has("lidar_encoder")

    @property
    def with_camera_encoder(self):
        return self._has("camera_encoder")

    @property
    def with_radar_encoder(self):
        return self._has("radar_encoder")

    @property
    def with_fusion_encoder(self):
        return self._has("fusion_encoder")

    def extract_feat(self, **kwargs):
        raise NotImplementedError("Must be implemented by yourself!")

    def onnx_export(self, **kwargs):
        raise NotImplementedError(f"{self.__class__.__name__} does not support ONNX EXPORT")


class BaseEncoder(nn.Module, metaclass=ABCMeta):
    def __init__(self, frozen_encoder: bool = False, frozen_bn: bool = False):
        super().__init__()
        self.frozen_encoder = frozen_encoder
        self.frozen_bn = frozen_bn

    def _freeze_stages(self):
        if self.frozen_encoder:
            for p in self.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        self._freeze_stages()
        if mode and self.frozen_bn:
            for m in self.modules():
                if isinstance(m, _BatchNorm):
                    m.eval()
        return self
