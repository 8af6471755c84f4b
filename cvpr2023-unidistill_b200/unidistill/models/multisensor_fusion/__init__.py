from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .base import BaseEncoder, BaseMultiSensorFusion  # noqa: E402

__all__ = ["BaseMultiSensorFusion", "BaseEncoder"]
