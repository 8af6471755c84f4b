from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .base_gen_proposals import BaseGenProposals  # noqa: E402
from .centerpoint_gen_proposals import CenterPointGenProposals  # noqa: E402
from .iou_aware_gen_proposals import IouAwareGenProposals  # noqa: E402

__all__ = ["BaseGenProposals", "CenterPointGenProposals", "IouAwareGenProposals"]
