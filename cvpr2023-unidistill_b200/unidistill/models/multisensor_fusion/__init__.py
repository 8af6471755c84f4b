from .base import BaseEncoder, BaseMultiSensorFusion

__all__ = ["BaseMultiSensorFusion", "BaseEncoder"]
