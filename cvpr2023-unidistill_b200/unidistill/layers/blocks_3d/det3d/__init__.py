from unidistill._overlay import extend as _extend

_extend(__path__, __name__)

from .spconv_backbone import SparseBasicBlock, VoxelResBackBone8x, post_act_block  # noqa: E402
from .vfe import MeanVFE  # noqa: E402

__all__ = ["MeanVFE", "VoxelResBackBone8x", "SparseBasicBlock", "post_act_block"]
