from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .base_assigner import BaseAssigner  # noqa: E402
from .fcos_assigner import FCOSAssigner  # noqa: E402

__all__ = ["BaseAssigner", "FCOSAssigner"]
