"""cuDNN epilogue fusion for FROZEN dense networks (the distillation teacher: eval mode, no gradients, BatchNorm already
folded into the convolutions by step.fold_frozen_batchnorm).

After folding, every layer of the teacher's ResNet / BEV backbone / head is `conv + bias` followed by a ReLU (and, at the end
of a residual block, by `+ identity` and a ReLU).  PyTorch runs that as three kernels -- the convolution, a broadcast
`add_` of the bias over the whole activation and a clamp -- i.e. two extra read+write passes over tensors of up to 277 MB
(24 images x 256 ch x 64 x 176).  cuDNN can apply bias, residual and ReLU in the convolution's epilogue
(`torch.cudnn_convolution_relu`, `torch.cudnn_convolution_add_relu`): same arithmetic, no extra passes.  These are library
calls on the dense part SURVEY 8a #14-#15 leaves to cuDNN; the hand-written kernels are elsewhere.
Only used under `torch.no_grad()` on CUDA tensors; otherwise the modules behave exactly as before."""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


def _fusable_conv(conv):
    return (type(conv) is nn.Conv2d and conv.padding_mode == "zeros" and isinstance(conv.padding, tuple)
            and not conv.weight.requires_grad)


def _pad_input_channels(conv, x):
    """(x, weight) with the input-channel dimension zero-padded to a multiple of 4 and both in NHWC.  The ResNet stem reads
    3-channel images: with C = 3 cuDNN falls back to an indexed sm80 kernel without shared-memory staging (0.42 ms for 24
    images on B200); with a fourth, all-zero channel (and zero weights for it: exact zeros in every sum, same result) the
    vectorised NHWC implicit GEMM applies.  The padded weight is cached on the module until the weight changes."""
    cin = int(conv.weight.shape[1])
    if cin % 4 == 0 or conv.groups != 1:
        return x, conv.weight
    pad = 4 - cin % 4
    key = (conv.weight.data_ptr(), conv.weight._version)
    hit = getattr(conv, "_ud_padded_weight", None)
    if hit is None or hit[0] != key:
        w = F.pad(conv.weight.detach(), (0, 0, 0, 0, 0, pad)).contiguous(memory_format=torch.channels_last)
        hit = (key, w)
        conv._ud_padded_weight = hit
    return F.pad(x, (0, 0, 0, 0, 0, pad)).contiguous(memory_format=torch.channels_last), hit[1]


def conv_bias_relu(conv, x, residual=None):
    """relu(conv(x) (+ residual)) with bias / residual / ReLU in the cuDNN epilogue when that is possible."""
    if x.is_cuda and not torch.is_grad_enabled() and _fusable_conv(conv):
        if residual is None:
            x, w = _pad_input_channels(conv, x)
            return torch.cudnn_convolution_relu(x, w, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)
        return torch.cudnn_convolution_add_relu(x, conv.weight, residual, 1.0, conv.bias, conv.stride, conv.padding,
                                                conv.dilation, conv.groups)
    y = conv(x)
    if residual is not None:
        y = y + residual
    return F.relu(y)


class FusedConvReLU(nn.Module):
    """Drop-in for a `Conv2d -> ReLU` pair inside an nn.Sequential (the ReLU's slot becomes nn.Identity).  Keeps the
    convolution as sub-module `conv`... the state_dict keys of a frozen teacher are not used after construction."""

    def __init__(self, conv):
        super().__init__()
        self.conv = conv

    def forward(self, x):
        return conv_bias_relu(self.conv, x)


def _shortcut(block, last_conv, x):
    """(identity tensor, bias for the block's last convolution).  A projection shortcut is `conv1x1 (+ folded BN)`: PyTorch
    runs its bias as a broadcast add over the whole activation (0.19 / 0.10 / 0.04 ms for the three projections of the
    teacher's ResNet-50 on B200); here the projection runs WITHOUT its bias, which is added to the last convolution's bias
    instead -- relu(conv3(out) + (b3 + b_ds) + conv_ds(x)) -- inside that convolution's cuDNN epilogue."""
    ds = block.downsample
    if ds is None:
        return x, last_conv.bias
    if (x.is_cuda and not torch.is_grad_enabled() and isinstance(ds, nn.Sequential) and len(ds) >= 1 and _fusable_conv(ds[0])
            and all(isinstance(m, nn.Identity) for m in list(ds)[1:]) and ds[0].bias is not None
            and last_conv.bias is not None and _fusable_conv(last_conv)):
        conv = ds[0]
        key = (conv.bias.data_ptr(), conv.bias._version, last_conv.bias.data_ptr(), last_conv.bias._version)
        hit = getattr(block, "_ud_shortcut_bias", None)
        if hit is None or hit[0] != key:
            hit = (key, (last_conv.bias.detach() + conv.bias.detach()))
            block._ud_shortcut_bias = hit
        return F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups), hit[1]
    return ds(x), last_conv.bias


def _conv_add_relu(conv, x, residual, bias):
    if x.is_cuda and not torch.is_grad_enabled() and _fusable_conv(conv):
        return torch.cudnn_convolution_add_relu(x, conv.weight, residual, 1.0, bias, conv.stride, conv.padding, conv.dilation,
                                                conv.groups)
    return F.relu(F.conv2d(x, conv.weight, bias, conv.stride, conv.padding, conv.dilation, conv.groups) + residual)


def _bottleneck_forward(self, x):
    identity, bias = _shortcut(self, self.conv3, x)
    out = conv_bias_relu(self.conv1, x)
    out = conv_bias_relu(self.conv2, out)
    return _conv_add_relu(self.conv3, out, identity, bias)


def _basicblock_forward(self, x):
    identity, bias = _shortcut(self, self.conv2, x)
    out = conv_bias_relu(self.conv1, x)
    return _conv_add_relu(self.conv2, out, identity, bias)


def fuse_frozen_conv_relu(module, skip=()):
    """Rewrites, in place, the conv -> (Identity) -> ReLU chains of a frozen, BN-folded network: nn.Sequential runs, and the
    residual blocks of torchvision's ResNet (conv / folded-bn attribute pairs).  Modules whose type name is in `skip` are
    left alone (the CenterHead towers run batched elsewhere).  Returns the number of fused convolutions."""
    n = 0
    for m in list(module.modules()):
        if type(m).__name__ in skip:
            continue
        cls = type(m).__name__
        if cls in ("Bottleneck", "BasicBlock") and all(isinstance(getattr(m, "bn%d" % i, None), nn.Identity)
                                                       for i in range(1, 4 if cls == "Bottleneck" else 3)):
            convs = [getattr(m, "conv%d" % i) for i in range(1, 4 if cls == "Bottleneck" else 3)]
            if all(_fusable_conv(c) for c in convs):
                m.forward = types.MethodType(_bottleneck_forward if cls == "Bottleneck" else _basicblock_forward, m)
                n += len(convs)
            continue
        if isinstance(m, nn.Sequential):
            names = list(m._modules.keys())
            i = 0
            while i < len(names):
                conv = m._modules[names[i]]
                if _fusable_conv(conv):
                    j = i + 1
                    while j < len(names) and isinstance(m._modules[names[j]], nn.Identity):
                        j += 1
                    if j < len(names) and type(m._modules[names[j]]) is nn.ReLU:
                        m._modules[names[i]] = FusedConvReLU(conv)
                        m._modules[names[j]] = nn.Identity()
                        n += 1
                        i = j
                i += 1
    return n
