"""spconv.core: only ConvAlgo is referenced by UniDistill (spconv_backbone.py:5-7)."""
import enum


class ConvAlgo(enum.Enum):
    Native = 0
    MaskImplicitGemm = 1
    MaskSplitImplicitGemm = 2
