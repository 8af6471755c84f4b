"""spconv.pytorch.ops.get_indice_pairs -- the explicit rulebook API that the reference's hand-rolled
SparseConvolution uses (reference layers/blocks_3d/det3d/centernet_backbone.py:232-251)."""
import torch

from unidistill_b200 import sparse as _sp


def get_conv_output_size(input_size, kernel_size, stride, padding, dilation):
    return _sp.conv_out_shape(list(input_size), list(kernel_size), list(stride), list(padding), list(dilation))


def get_indice_pairs(indices, batch_size, spatial_shape, ksize=3, stride=1, padding=0, dilation=1,
                     out_padding=0, subm=False, transpose=False, grid=None, use_hash=False):
    """-> (outids (N_out,4) int32, indice_pairs (2,K,N_in) int32 -1 padded, indice_pair_num (K,) int32)."""
    if transpose:
        raise NotImplementedError("unidistill_b200 spconv shim: transposed rulebooks are not provided")
    indices = indices.int().contiguous()
    if subm:
        rb = _sp.build_subm_rulebook(indices, batch_size, list(spatial_shape), ksize, dilation)
    else:
        rb = _sp.build_conv_rulebook(indices, batch_size, list(spatial_shape), ksize, stride, padding, dilation)
    pairs, pair_num = rb.spconv_pairs()
    return rb.out_coords, pairs, pair_num
