"""fold_frozen_batchnorm: the frozen teacher with BatchNorm folded into its dense convolutions computes the same
function (CPU, fp32) -- BEV backbone (Sequential runs, ConvTranspose deblocks) and the ResNet stand-in (convN/bnN
attribute pairs, downsample Sequentials)."""
import torch
from torch import nn


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            m.weight.data = torch.rand(m.weight.shape, generator=g) + 0.5
            m.bias.data = torch.randn(m.bias.shape, generator=g) * 0.1
            m.running_mean.data = torch.randn(m.running_mean.shape, generator=g) * 0.2
            m.running_var.data = torch.rand(m.running_var.shape, generator=g) + 0.5


def test_fold_bev_backbone_and_resnet():
    from unidistill.layers.blocks_2d.det3d import BaseBEVBackbone
    from unidistill_b200.imgnets import ResNet, SECONDFPN
    from unidistill_b200.step import fold_frozen_batchnorm

    torch.manual_seed(0)
    bev = BaseBEVBackbone(layer_nums=[2, 2], layer_strides=[1, 2], num_filters=[16, 32], upsample_strides=[1, 2],
                          num_upsample_filters=[32, 32], input_channels=8).eval()
    res = ResNet(depth=18).eval()
    neck = SECONDFPN(in_channels=[64, 128, 256, 512], upsample_strides=[0.25, 0.5, 1, 2], out_channels=[8] * 4).eval()
    for i, m in enumerate((bev, res, neck)):
        _randomise_bn(m, i)
    x_bev, x_img = torch.randn(2, 8, 16, 16), torch.randn(2, 3, 64, 96)
    with torch.no_grad():
        ref_bev = bev(x_bev)[0]
        ref_img = neck(res(x_img))[0]
    n_bn = sum(isinstance(m, nn.BatchNorm2d) for mod in (bev, res, neck) for m in mod.modules())
    folded = sum(fold_frozen_batchnorm(m) for m in (bev, res, neck))
    assert folded == n_bn and folded > 0
    assert not any(isinstance(m, nn.BatchNorm2d) for mod in (bev, res, neck) for m in mod.modules())
    with torch.no_grad():
        got_bev = bev(x_bev)[0]
        got_img = neck(res(x_img))[0]
    assert torch.allclose(got_bev, ref_bev, rtol=1e-4, atol=1e-5)
    assert torch.allclose(got_img, ref_img, rtol=1e-4, atol=1e-5)


def test_fold_leaves_train_mode_bn_alone():
    from unidistill_b200.step import fold_frozen_batchnorm

    m = nn.Sequential(nn.Conv2d(3, 4, 3), nn.BatchNorm2d(4), nn.ReLU()).train()
    assert fold_frozen_batchnorm(m) == 0
    assert isinstance(m[1], nn.BatchNorm2d)
