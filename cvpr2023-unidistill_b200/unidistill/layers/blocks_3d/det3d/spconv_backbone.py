"""VoxelResBackBone8x -- the sparse 3-D encoder of the LiDAR branch.

Same module tree (hence the same state_dict keys) and the same forward contract as the reference
(unidistill/layers/blocks_3d/det3d/spconv_backbone.py:61-113,252-384): conv_input, conv1..conv4,
conv_out; channels 16/32/64/128; BatchNorm1d(eps=1e-3, momentum=0.01).  Every conv runs as one
gather-GEMM launch on a cached rulebook and every BN(+residual)+ReLU as one fused pass.
"""
import os
from functools import partial

import numpy as np
import spconv.pytorch as spconv
import torch.nn as nn
from spconv.core import ConvAlgo

from unidistill_b200.sparse import batch_norm_act, morton_order

# Build every rulebook of the encoder before its feature pass (spconv.prebuild_rulebooks of this repo's shim): the host reads
# of the strided builders then sit in one short geometry phase instead of between the convolutions.  UD_PREBUILD_RULEBOOKS=0
# restores the lazy, layer-by-layer construction of the reference.
PREBUILD_RULEBOOKS = os.environ.get("UD_PREBUILD_RULEBOOKS", "1") != "0"

SPCONV_ALGO_TYPE = ConvAlgo.Native


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0,
                   conv_type="subm", norm_fn=None):
    common = dict(bias=False, indice_key=indice_key, algo=SPCONV_ALGO_TYPE)
    if conv_type == "subm":
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, **common)
    elif conv_type == "spconv":
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding,
                                   **common)
    elif conv_type == "inverseconv":
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, **common)
    else:
        raise NotImplementedError(conv_type)
    return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


def _fusable(bn, feats):
    c = feats.shape[1]
    return type(bn) is nn.BatchNorm1d and feats.is_cuda and c % 4 == 0 and 256 % c == 0 and c <= 128


class SparseBasicBlock(spconv.SparseModule):
    """conv-bn-relu-conv-bn-(+identity)-relu on a fixed coordinate set (reference :61-113)."""

    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_fn=None, downsample=None, indice_key=None):
        super().__init__()
        assert norm_fn is not None
        conv_kw = dict(kernel_size=3, stride=stride, padding=1, bias=True, indice_key=indice_key,
                       algo=SPCONV_ALGO_TYPE)
        self.conv1 = spconv.SubMConv3d(inplanes, planes, **conv_kw)
        self.bn1 = norm_fn(planes)
        self.relu = nn.ReLU()
        self.conv2 = spconv.SubMConv3d(planes, planes, **conv_kw)
        self.bn2 = norm_fn(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x if self.downsample is None else self.downsample(x)
        if not self.bn1.training and not self.bn2.training and x.features.is_cuda and x.indices.shape[0] > 0:
            # frozen / inference: BatchNorm, residual and ReLU ride in the convolutions' epilogues (None = not applicable)
            out = self.conv1(x, bn=self.bn1, relu=True)
            if out is not None:
                out2 = self.conv2(out, bn=self.bn2, relu=True, residual=identity.features)
                if out2 is not None:
                    return out2
        out = self.conv1(x)
        if _fusable(self.bn1, out.features):
            out = out.replace_feature(batch_norm_act(out.features, self.bn1, relu=True))
            out = self.conv2(out)
            out = out.replace_feature(
                batch_norm_act(out.features, self.bn2, relu=True, residual=identity.features))
            return out
        out = out.replace_feature(self.relu(self.bn1(out.features)))
        out = self.conv2(out)
        out = out.replace_feature(self.bn2(out.features))
        out = out.replace_feature(self.relu(out.features + identity.features))
        return out


class VoxelBackBone8x(nn.Module):
    """Plain (non-residual) variant, reference :116-250: conv-bn-relu blocks only, batch_dict in / batch_dict out.  Not used
    by the BEVFusion experiments; kept so that `unidistill.layers.blocks_3d.det3d` exports what the reference's does."""

    # (attribute, cin, cout, padding of the strided conv | None for the SubM-only stage, strided key, SubM key, #SubM)
    _PLAN = (("conv1", 16, 16, None, None, "subm1", 1), ("conv2", 16, 32, 1, "spconv2", "subm2", 2),
             ("conv3", 32, 64, 1, "spconv3", "subm3", 2), ("conv4", 64, 64, (0, 1, 1), "spconv4", "subm4", 2))

    def __init__(self, input_channels, grid_size, last_pad=0):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = np.asarray(grid_size)[::-1] + [1, 0, 0]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1"), norm_fn(16), nn.ReLU())
        for name, cin, cout, pad, down_key, subm_key, n_subm in self._PLAN:
            blocks = []
            if down_key is not None:
                blocks.append(post_act_block(cin, cout, 3, norm_fn=norm_fn, stride=2, padding=pad, indice_key=down_key,
                                             conv_type="spconv"))
            blocks += [post_act_block(cout, cout, 3, norm_fn=norm_fn, padding=1, indice_key=subm_key)
                       for _ in range(n_subm)]
            setattr(self, name, spconv.SparseSequential(*blocks))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key="spconv_down2", algo=SPCONV_ALGO_TYPE),
            norm_fn(128), nn.ReLU())
        self.num_point_features = 128

    def forward(self, batch_dict):
        x = spconv.SparseConvTensor(features=batch_dict["voxel_features"], indices=batch_dict["voxel_coords"].int(),
                                    spatial_shape=self.sparse_shape, batch_size=batch_dict["batch_size"])
        x = self.conv_input(x)
        scales = {}
        for name in ("conv1", "conv2", "conv3", "conv4"):
            x = getattr(self, name)(x)
            scales["x_" + name] = x
        batch_dict.update({"encoded_spconv_tensor": self.conv_out(x), "encoded_spconv_tensor_stride": 8,
                           "multi_scale_3d_features": scales})
        return batch_dict


class VoxelResBackBone8x(nn.Module):
    # (attribute name, in, out, padding of the strided conv, indice keys)
    _STAGES = (
        ("conv2", 16, 32, 1, "spconv2", "res2"),
        ("conv3", 32, 64, 1, "spconv3", "res3"),
        ("conv4", 64, 128, (0, 1, 1), "spconv4", "res4"),
    )

    def __init__(self, input_channels, grid_size, last_pad=0):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        # (z, y, x) with one extra z slice, as in the reference (:257)
        self.sparse_shape = np.asarray(grid_size)[::-1] + [1, 0, 0]
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1",
                              algo=SPCONV_ALGO_TYPE),
            norm_fn(16),
            nn.ReLU(),
        )
        self.conv1 = spconv.SparseSequential(
            SparseBasicBlock(16, 16, norm_fn=norm_fn, indice_key="res1"),
            SparseBasicBlock(16, 16, norm_fn=norm_fn, indice_key="res1"),
        )
        for name, cin, cout, pad, down_key, res_key in self._STAGES:
            setattr(self, name, spconv.SparseSequential(
                post_act_block(cin, cout, 3, norm_fn=norm_fn, stride=2, padding=pad, indice_key=down_key,
                               conv_type="spconv"),
                SparseBasicBlock(cout, cout, norm_fn=norm_fn, indice_key=res_key),
                SparseBasicBlock(cout, cout, norm_fn=norm_fn, indice_key=res_key),
            ))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(128, 128, (3, 1, 1), stride=(2, 1, 1), padding=last_pad, bias=False,
                                indice_key="spconv_down2"),
            norm_fn(128),
            nn.ReLU(),
        )
        self.num_point_features = 128
        # True: rows are put into Morton (Z-order) order on entry (pair with unidistill_b200.sparse.CONV_OUTPUT_ORDER =
        # "morton" for the strided layers).  The row order of a sparse tensor is not observable after dense().  Off by
        # default: measured slower than the raster order on B200 (see unidistill_b200/sparse.py).
        self.morton_order = os.environ.get("UD_CONV_ORDER", "linear") == "morton"

    def forward(self, voxel_features, voxel_coords, batch_size, indice_dict=None):
        """voxel_features (M,C), voxel_coords (M,4) [b,z,y,x] -> (SparseConvTensor, 8, {x_conv1..4}).
        `indice_dict` (extension, default = the reference's behaviour): rulebook cache to start from / fill -- a second
        encoder run over the SAME voxel_coords tensor reuses the tables instead of rebuilding them."""
        voxel_coords = voxel_coords.int()
        if self.morton_order and voxel_coords.is_cuda and voxel_coords.shape[0] > 0:
            cached = indice_dict.get("__morton__") if indice_dict is not None else None
            if cached is None or cached[0] is not voxel_coords:
                order = morton_order(voxel_coords.contiguous(), batch_size, self.sparse_shape)
                cached = (voxel_coords, order, voxel_coords.index_select(0, order))
                if indice_dict is not None:
                    indice_dict["__morton__"] = cached
            voxel_features = voxel_features.index_select(0, cached[1])
            voxel_coords = cached[2]
        x = spconv.SparseConvTensor(features=voxel_features, indices=voxel_coords,
                                    spatial_shape=self.sparse_shape, batch_size=batch_size, indice_dict=indice_dict)
        if PREBUILD_RULEBOOKS and voxel_coords.is_cuda and hasattr(spconv, "prebuild_rulebooks"):
            # geometry first: all rulebooks (and their host reads) before the first feature kernel
            spconv.prebuild_rulebooks(self, x.indices, x.spatial_shape, batch_size, x.indice_dict)
        x = self.conv_input(x)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        out = self.conv_out(x_conv4)
        scales = {"x_conv1": x_conv1, "x_conv2": x_conv2, "x_conv3": x_conv3, "x_conv4": x_conv4}
        return out, 8, scales
