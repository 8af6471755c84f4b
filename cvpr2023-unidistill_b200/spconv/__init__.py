"""Drop-in shim for the slice of the `spconv` 2.x API that UniDistill touches.

The reference imports `spconv.pytorch as spconv`, `spconv.core.ConvAlgo` and
`spconv.pytorch.utils.PointToVoxel` (reference spconv_backbone.py:3-7, voxelization.py:4).  spconv's own
CUDA kernels do not exist for sm_100; this package keeps the names and call signatures and routes every
operator to the hand-written sm_100a kernels in libunidistill_b200.so.  No CPU path.
"""
__version__ = "2.1.21+unidistill_b200"
