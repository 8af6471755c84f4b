"""GPU parity of the camera-branch kernels (voxel pooling, fused lift-splat) against the CPU oracle."""
import numpy as np
import pytest
import torch

from oracle import camera as OC

pytestmark = pytest.mark.gpu


def _rel_err(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


def _rand_geom(rng, B, N, nx, ny, nz, frac_out=0.2):
    g = np.stack([rng.integers(0, nx, (B, N)), rng.integers(0, ny, (B, N)), rng.integers(0, nz, (B, N))], -1)
    bad = rng.random((B, N)) < frac_out
    g[bad] = np.stack([rng.integers(-3, nx + 3, bad.sum()), rng.integers(-3, ny + 3, bad.sum()),
                       rng.integers(-1, nz + 1, bad.sum())], -1)
    return g.astype(np.int32)


@pytest.mark.parametrize("B,N,C,nx,ny", [(1, 50, 8, 5, 4), (2, 3000, 256, 16, 12), (3, 1777, 80, 9, 7),
                                         (2, 999, 7, 6, 6)])
def test_voxel_pooling_forward_backward(cuda, native_lib, B, N, C, nx, ny):
    from unidistill_b200.camera import voxel_pooling

    rng = np.random.default_rng(N)
    geom = _rand_geom(rng, B, N, nx, ny, 1)
    feats = rng.standard_normal((B, N, C)).astype(np.float32)
    f = torch.from_numpy(feats).to(cuda).requires_grad_(True)
    out = voxel_pooling(torch.from_numpy(geom).to(cuda), f, torch.tensor([nx, ny, 1]))
    ref, pos = OC.voxel_pooling(geom, feats, [nx, ny, 1])
    assert out.shape == (B, C, ny, nx) and not out.is_contiguous()     # permuted NHWC view, like the reference
    # same summation order (ascending point index) as the C restatement => bit exact
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    g = rng.standard_normal((B, C, ny, nx)).astype(np.float32)
    out.backward(torch.from_numpy(g).to(cuda))
    assert np.array_equal(f.grad.cpu().numpy(), OC.voxel_pooling_backward(g, pos, C))


def test_voxel_pooling_all_points_one_cell_and_none(cuda, native_lib):
    from unidistill_b200.camera import voxel_pooling

    rng = np.random.default_rng(1)
    feats = rng.standard_normal((1, 4097, 256)).astype(np.float32)
    geom = np.tile(np.array([[3, 2, 0]], np.int32), (1, 4097, 1))
    out = voxel_pooling(torch.from_numpy(geom).to(cuda), torch.from_numpy(feats).to(cuda), torch.tensor([5, 4, 1]))
    ref, _ = OC.voxel_pooling(geom, feats, [5, 4, 1])
    assert np.array_equal(out.cpu().numpy(), ref)
    geom[:] = -7
    out = voxel_pooling(torch.from_numpy(geom).to(cuda), torch.from_numpy(feats).to(cuda), torch.tensor([5, 4, 1]))
    assert not out.cpu().numpy().any()


def _torch_lift_splat(depth, ctx, geom, nx, ny, nz):
    """Differentiable torch-CPU restatement of lss_fpn.py:289-316 + index_add pooling."""
    B = geom.shape[0]
    C = ctx.shape[1]
    vol = (depth.unsqueeze(1) * ctx.unsqueeze(2)).permute(0, 2, 3, 4, 1)          # (BN, D, H, W, C)
    vol = vol.reshape(B, -1, C)
    g = geom.reshape(B, -1, 3).long()
    ok = (g[..., 0] >= 0) & (g[..., 0] < nx) & (g[..., 1] >= 0) & (g[..., 1] < ny) & (g[..., 2] >= 0) & (g[..., 2] < nz)
    bidx = torch.arange(B).view(B, 1).expand_as(ok)
    cell = (bidx * ny + g[..., 1]) * nx + g[..., 0]
    out = torch.zeros((B * ny * nx, C), dtype=vol.dtype)
    out = out.index_add(0, cell[ok], vol[ok])
    return out.view(B, ny, nx, C).permute(0, 3, 1, 2)


@pytest.mark.parametrize("B,ncam,D,H,W,C,nx,ny", [(2, 3, 7, 4, 6, 128, 40, 33), (1, 6, 16, 8, 11, 256, 64, 64)])
def test_lift_splat_forward_backward(cuda, native_lib, B, ncam, D, H, W, C, nx, ny):
    from unidistill_b200.camera import lift_splat

    torch.manual_seed(D)
    rng = np.random.default_rng(D)
    depth = torch.randn(B * ncam, D, H, W).softmax(1)
    ctx = torch.randn(B * ncam, C, H, W)
    geom = torch.from_numpy(_rand_geom(rng, B, ncam * D * H * W, nx, ny, 1).reshape(B, ncam, D, H, W, 3))
    dg = depth.to(cuda).requires_grad_(True)
    cg = ctx.to(cuda).requires_grad_(True)
    out = lift_splat(dg, cg, geom.to(cuda), [nx, ny, 1])
    assert out.shape == (B, C, ny, nx) and out.is_contiguous()
    ref = OC.lift_splat_reference(depth, ctx, geom.numpy(), [nx, ny, 1])
    assert _rel_err(out.detach().cpu().numpy(), ref) < 1e-5
    dc = depth.clone().requires_grad_(True)
    cc = ctx.clone().requires_grad_(True)
    ref_t = _torch_lift_splat(dc, cc, geom, nx, ny, 1)
    gout = torch.randn(B, C, ny, nx)
    ref_t.backward(gout)
    out.backward(gout.to(cuda))
    assert _rel_err(dg.grad.cpu().numpy(), dc.grad.numpy()) < 1e-4
    assert _rel_err(cg.grad.cpu().numpy(), cc.grad.numpy()) < 1e-4


def test_lssfpn_fused_equals_unfused_full_size(cuda, native_lib):
    """BASELINE camera config geometry (6 x 256x704, D=112, 128^2 BEV): fused lift-splat == materialised
    outer product + voxel_pooling (the reference's formulation) run through the same module."""
    from unidistill.layers.blocks_3d.mmdet3d.lss_fpn import LSSFPN
    from unidistill_b200 import synthetic as S

    torch.manual_seed(0)
    m = LSSFPN([-51.2, 51.2, 0.8], [-51.2, 51.2, 0.8], [-5, 3, 8], [2.0, 58.0, 0.5], (256, 704), 16, 256,
               dict(type="ResNet", depth=18, out_indices=[0, 1, 2, 3]),
               dict(type="SECONDFPN", in_channels=[64, 128, 256, 512], upsample_strides=[0.25, 0.5, 1, 2],
                    out_channels=[128] * 4),
               dict(in_channels=512, mid_channels=512)).to(cuda).eval()
    batch = S.camera_batch(1, seed=0, random_bda=True)
    imgs = batch["imgs"].to(cuda)
    mats = {k: v.to(cuda) for k, v in batch["mats_dict"].items()}
    with torch.no_grad():
        m.fuse_lift_splat = True
        a = m(imgs, mats)
        m.fuse_lift_splat = False
        b = m(imgs, mats)
    assert a.shape == (1, 256, 128, 128)
    assert _rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
    assert float((a.abs().sum(1) > 0).float().mean()) > 0.2
