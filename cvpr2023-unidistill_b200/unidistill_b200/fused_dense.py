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


def _bottleneck_forward(self, x):
    identity = x if self.downsample is None else self.downsample(x)
    out = conv_bias_relu(self.conv1, x)
    out = conv_bias_relu(self.conv2, out)
    return conv_bias_relu(self.conv3, out, identity)


def _basicblock_forward(self, x):
    identity = x if self.downsample is None else self.downsample(x)
    out = conv_bias_relu(self.conv1, x)
    return conv_bias_relu(self.conv2, out, identity)


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
