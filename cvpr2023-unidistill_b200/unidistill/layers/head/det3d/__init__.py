from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .center_head import CenterHead, SepHead  # noqa: E402
from .center_head_iou_aware import CenterHeadIouAware  # noqa: E402
from .generate_proposals import BaseGenProposals, CenterPointGenProposals, IouAwareGenProposals  # noqa: E402
from .target_assigner import BaseAssigner, FCOSAssigner  # noqa: E402

__all__ = ["CenterHead", "SepHead", "CenterHeadIouAware", "BaseGenProposals", "CenterPointGenProposals",
           "IouAwareGenProposals", "BaseAssigner", "FCOSAssigner"]
