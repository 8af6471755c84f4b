from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .mean_vfe import MeanVFE  # noqa: E402
from .vfe_template import VFETemplate  # noqa: E402

__all__ = ["MeanVFE", "VFETemplate"]
