from unidistill._overlay import extend as _extend

_extend(__path__, __name__, globals())

from .spconv_backbone import SparseBasicBlock, VoxelBackBone8x, VoxelResBackBone8x, post_act_block  # noqa: E402
from .vfe import MeanVFE  # noqa: E402

__all__ = ["MeanVFE", "VoxelBackBone8x", "VoxelResBackBone8x", "SparseBasicBlock", "post_act_block"]
