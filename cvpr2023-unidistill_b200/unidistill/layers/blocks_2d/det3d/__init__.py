from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .base_bev_backbone import BaseBEVBackbone  # noqa: E402
from .map_to_bev import HeightCompression, PointPillarScatter  # noqa: E402

__all__ = ["BaseBEVBackbone", "HeightCompression", "PointPillarScatter"]
