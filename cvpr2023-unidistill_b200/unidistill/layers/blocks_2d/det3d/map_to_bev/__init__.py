from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .height_compression import HeightCompression  # noqa: E402
from .pointpillar_scatter import PointPillarScatter  # noqa: E402

__all__ = ["HeightCompression", "PointPillarScatter"]
