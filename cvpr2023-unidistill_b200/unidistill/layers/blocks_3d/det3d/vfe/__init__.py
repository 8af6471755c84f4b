from .mean_vfe import MeanVFE
from .vfe_template import VFETemplate

__all__ = ["MeanVFE", "VFETemplate"]
