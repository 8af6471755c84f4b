"""spconv.pytorch shim: SparseConvTensor, SubMConv3d, SparseConv3d, SparseInverseConv3d,
SparseSequential, SparseModule -- the symbols reference spconv_backbone.py:21-113,259-359 and
height_compression.py:19 use, backed by unidistill_b200.sparse (sm_100a kernels).

Parameter layout follows spconv 2.x: weight (Cout, kz, ky, kx, Cin), optional bias (Cout,), so state
dicts trained with the reference load by key and shape.
"""
import math
from collections import OrderedDict

import torch
from torch import nn

from unidistill_b200 import sparse as _sp

from . import functional, ops  # noqa: F401  (re-exported sub-modules)
from ..core import ConvAlgo  # noqa: F401


class SparseModule(nn.Module):
    """Marker base class: modules that consume and produce a SparseConvTensor."""


class SparseConvTensor(object):
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None,
                 indice_dict=None, benchmark=False):
        if indices.dtype != torch.int32:
            raise TypeError("SparseConvTensor: indices must be int32 [b, z, y, x]")
        self._features = features
        self.indices = indices.contiguous()
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark
        self.benchmark_record = {}

    @property
    def features(self):
        return self._features

    @features.setter
    def features(self, value):
        self._features = value

    def replace_feature(self, feature):
        new = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid,
                               self.voxel_num, self.indice_dict, self.benchmark)
        new.benchmark_record = self.benchmark_record
        return new

    @property
    def spatial_size(self):
        out = 1
        for s in self.spatial_shape:
            out *= s
        return out

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key, None)

    def dense(self, channels_first=True):
        out = _sp.scatter_to_dense(self.features, self.indices, self.batch_size, self.spatial_shape)
        if not channels_first:
            return out.permute(0, 2, 3, 4, 1).contiguous()
        return out

    @property
    def sparity(self):
        return self.indices.shape[0] / max(1, self.spatial_size) / max(1, self.batch_size)


def _ntuple3(v):
    return [int(x) for x in v] if isinstance(v, (list, tuple)) else [int(v)] * 3


class SparseConvolution(SparseModule):
    """Common body of SubMConv3d / SparseConv3d / SparseInverseConv3d."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, output_padding=0, transposed=False, inverse=False,
                 indice_key=None, algo=None, fp32_accum=None, record_voxel_count=False, act_type=None,
                 act_alpha=0, act_beta=0, name=None):
        super().__init__()
        if ndim != 3:
            raise NotImplementedError("unidistill_b200 spconv shim: only 3-D convolutions")
        if groups != 1:
            raise NotImplementedError("unidistill_b200 spconv shim: groups must be 1")
        if transposed:
            raise NotImplementedError("unidistill_b200 spconv shim: SparseConvTranspose3d is not provided")
        self.ndim = ndim
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        self.kernel_size = _ntuple3(kernel_size)
        self.stride = _ntuple3(stride)
        self.padding = _ntuple3(padding)
        self.dilation = _ntuple3(dilation)
        self.output_padding = _ntuple3(output_padding)
        self.conv1x1 = all(k == 1 for k in self.kernel_size)
        self.groups = groups
        self.subm, self.inverse, self.transposed = subm, inverse, transposed
        self.indice_key = indice_key
        self.algo = algo
        self.weight = nn.Parameter(torch.empty(self.out_channels, *self.kernel_size, self.in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(self.out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def extra_repr(self):
        return "%d, %d, kernel_size=%s, stride=%s, padding=%s, subm=%s, indice_key=%s" % (
            self.in_channels, self.out_channels, self.kernel_size, self.stride, self.padding, self.subm,
            self.indice_key)

    def reset_parameters(self):
        fan_in = self.in_channels * self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        bound = 1.0 / math.sqrt(fan_in)  # kaiming_uniform_(a=sqrt(5)) on fan_in, as nn.Conv3d
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def _rulebook(self, x):
        rb = x.find_indice_pair(self.indice_key)
        if self.inverse:
            if rb is None or rb.subm:
                raise RuntimeError("SparseInverseConv3d needs the rulebook of its paired SparseConv3d "
                                   "(indice_key=%r)" % self.indice_key)
            return rb
        if rb is not None:
            if self.subm and rb.subm and rb.ksize == self.kernel_size and rb.n_in == x.indices.shape[0]:
                return rb
            if (not self.subm) and (not rb.subm) and rb.ksize == self.kernel_size and \
                    rb.stride == self.stride and rb.n_in == x.indices.shape[0]:
                return rb
        if self.subm:
            # identical coordinate sets share one table even under different keys (subm1 / res1)
            cache_key = ("__subm__", x.indices.data_ptr(), int(x.indices.shape[0]), tuple(self.kernel_size),
                         tuple(self.dilation))
            rb = x.indice_dict.get(cache_key)
            if rb is None:
                rb = _sp.build_subm_rulebook(x.indices, x.batch_size, x.spatial_shape, self.kernel_size,
                                             self.dilation)
                x.indice_dict[cache_key] = rb
        else:
            rb = _sp.build_conv_rulebook(x.indices, x.batch_size, x.spatial_shape, self.kernel_size,
                                         self.stride, self.padding, self.dilation)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = rb
        return rb

    def forward(self, x, bn=None, relu=False, residual=None):
        """`bn` (extension used by SparseSequential / SparseBasicBlock): an eval-mode BatchNorm1d whose application, the
        optional residual add and ReLU are folded into this convolution's epilogue when no gradient is needed; returns
        None in that case if the fused form does not apply (the caller then runs the modules one by one)."""
        assert isinstance(x, SparseConvTensor)
        rb = self._rulebook(x)
        if bn is not None:
            if self.inverse:
                return None
            feats = _sp.sparse_conv_eval_bn_act(x.features, self.weight, self.bias, rb, bn, relu=relu, residual=residual)
            if feats is None:
                return None
        else:
            feats = _sp.sparse_conv(x.features, self.weight, self.bias, rb, self.inverse)
        if self.subm:
            out = SparseConvTensor(feats, x.indices, x.spatial_shape, x.batch_size, x.grid, x.voxel_num,
                                   x.indice_dict, x.benchmark)
        elif self.inverse:
            # back to the coordinate set the paired conv consumed
            out = SparseConvTensor(feats, rb.in_coords_for_inverse, rb.in_shape, x.batch_size, x.grid,
                                   x.voxel_num, x.indice_dict, x.benchmark)
        else:
            out = SparseConvTensor(feats, rb.out_coords, rb.out_shape, x.batch_size, x.grid, x.voxel_num,
                                   x.indice_dict, x.benchmark)
            rb.in_coords_for_inverse = x.indices
        return out


def prebuild_rulebooks(module, indices, spatial_shape, batch_size, indice_dict):
    """Extension: builds the rulebooks of every sparse convolution under `module` from the voxel coordinates alone, in
    module-registration order (== execution order for sequential encoders such as VoxelResBackBone8x), into `indice_dict`.
    Rulebooks depend on coordinates only, and the strided builders read the output row count back to the host; doing all of
    that BEFORE the feature pass keeps the five host reads of an encoder in one short geometry phase, after which the
    convolution / BatchNorm launches run ahead of the GPU without a single synchronisation.  A rulebook that does not fit the
    tensor a layer later sees is simply rebuilt by that layer (`_rulebook` checks key, kernel, stride and row count)."""
    x = SparseConvTensor(None, indices, spatial_shape, batch_size, indice_dict=indice_dict)
    for m in module.modules():
        if not isinstance(m, SparseConvolution) or m.inverse:
            continue
        rb = m._rulebook(x)
        if not m.subm:
            x = SparseConvTensor(None, rb.out_coords, rb.out_shape, batch_size, indice_dict=indice_dict)
    return indice_dict


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None, fp32_accum=None, name=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         subm=True, indice_key=indice_key, algo=algo, name=name)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None, fp32_accum=None, name=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         subm=False, indice_key=indice_key, algo=algo, name=name)


class SparseInverseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, indice_key, bias=True, algo=None,
                 fp32_accum=None, name=None, **kwargs):
        super().__init__(3, in_channels, out_channels, kernel_size, bias=bias, inverse=True,
                         indice_key=indice_key, algo=algo, name=name)


def is_spconv_module(module):
    return isinstance(module, SparseModule)


class SparseSequential(SparseModule):
    """nn.Sequential over a SparseConvTensor: sparse modules get the tensor, dense modules its features.

    The pattern [sparse conv, BatchNorm1d, ReLU] (reference post_act_block, spconv_backbone.py:50-56)
    is executed as conv + one fused BN/ReLU pass.
    """

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            module = mods[i]
            if is_spconv_module(module):
                assert isinstance(input, SparseConvTensor)
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if (isinstance(module, SparseConvolution) and type(nxt) is nn.BatchNorm1d and not nxt.training
                        and input.features.is_cuda and input.indices.shape[0] > 0):
                    # [conv, eval-mode BatchNorm1d(, ReLU)] without gradients: one kernel
                    relu = i + 2 < len(mods) and type(mods[i + 2]) is nn.ReLU
                    fused = module(input, bn=nxt, relu=relu)
                    if fused is not None:
                        input = fused
                        i += 3 if relu else 2
                        continue
                input = module(input)
                i += 1
                continue
            if isinstance(input, SparseConvTensor):
                if input.indices.shape[0] == 0:
                    i += 1
                    continue
                nxt = mods[i + 1] if i + 1 < len(mods) else None
                if (type(module) is nn.BatchNorm1d and input.features.is_cuda
                        and input.features.shape[1] % 4 == 0 and 256 % input.features.shape[1] == 0
                        and input.features.shape[1] <= 128):
                    relu = type(nxt) is nn.ReLU
                    feats = _sp.batch_norm_act(input.features, module, relu=relu)
                    input = input.replace_feature(feats)
                    i += 2 if relu else 1
                    continue
                input = input.replace_feature(module(input.features))
            else:
                input = module(input)
            i += 1
        return input
