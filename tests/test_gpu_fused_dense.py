"""Frozen, BatchNorm-folded image backbone with cuDNN epilogue fusion (unidistill_b200/fused_dense.py) against the untouched
eval-mode network on the GPU: the 4-channel stem, bias / residual / ReLU in the convolution epilogues and the projection
shortcuts whose bias moves into the block's last convolution."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("depth", [18, 50])
def test_fused_frozen_resnet_equals_the_plain_eval_network(cuda, native_lib, depth):
    from unidistill_b200 import fused_dense, imgnets
    from unidistill_b200.step import fold_frozen_batchnorm

    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False          # compare fp32 with fp32
    try:
        torch.manual_seed(depth)
        plain = imgnets.ResNet(depth=depth).to(cuda).eval()
        with torch.no_grad():
            for m in plain.modules():                # non-trivial folded scales and biases
                if isinstance(m, torch.nn.BatchNorm2d):
                    m.running_mean.normal_(0, 0.2)
                    m.running_var.uniform_(0.5, 1.5)
                    m.weight.uniform_(0.5, 1.5)
                    m.bias.normal_(0, 0.2)
        fused = copy.deepcopy(plain)
        for p in fused.parameters():
            p.requires_grad = False
        assert fold_frozen_batchnorm(fused) > 10
        fused = fused.to(memory_format=torch.channels_last)
        assert fused_dense.fuse_frozen_conv_relu(fused) > 10
        x = torch.randn(4, 3, 64, 96, device=cuda)
        with torch.no_grad():
            ref = plain(x)
            got = fused(x.contiguous(memory_format=torch.channels_last))
        assert len(ref) == len(got) == 4
        for r, g in zip(ref, got):
            scale = float(r.abs().max())
            assert float((r - g).abs().max()) <= 2e-4 * scale, (depth, float((r - g).abs().max()), scale)
        # every projection shortcut took the bias-free path (its bias lives in the block's last convolution)
        n_proj = sum(1 for m in fused.modules() if type(m).__name__ in ("Bottleneck", "BasicBlock") and m.downsample is not None)
        n_moved = sum(1 for m in fused.modules() if hasattr(m, "_ud_shortcut_bias"))
        assert n_proj > 0 and n_moved == n_proj
        assert hasattr(fused.conv1, "_ud_padded_weight") and fused.conv1._ud_padded_weight[1].shape[1] == 4
    finally:
        torch.backends.cudnn.allow_tf32 = old
