"""Tensor-core (tcgen05 / TMEM, bf16 hi/lo operand split) sparse convolution vs the C oracle and vs the fp32 SIMT kernel.

Tolerance: BASELINE.json's north_star allows 1e-3 relative for fp32 features; the hi/lo operand split keeps
the tensor-core path at ~1e-5 (the MMA accumulates with truncation), asserted here as 5e-5 relative to the
largest output.
"""
import os

import numpy as np
import pytest
import torch

from oracle import lidar as O

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _canonical_conv_order():
    """These tests compare rows with the oracle's canonical (ascending linear index) output order; the Morton order the
    model path uses by default is covered by tests/test_gpu_morton.py."""
    from unidistill_b200 import sparse

    old = sparse.CONV_OUTPUT_ORDER
    sparse.CONV_OUTPUT_ORDER = "linear"
    yield
    sparse.CONV_OUTPUT_ORDER = old
TC_TOL = 5e-5


def _cloud(rng, B, shape, n, C):
    D, H, W = shape
    lin = rng.choice(B * D * H * W, size=n, replace=False)
    lin.sort()
    b, r = np.divmod(lin, D * H * W)
    z, r = np.divmod(r, H * W)
    y, x = np.divmod(r, W)
    coords = np.stack([b, z, y, x], 1).astype(np.int32)
    feat = rng.standard_normal((n, C)).astype(np.float32)
    return coords, feat


def _rel(a, b):
    return float(np.abs(a - b).max() / max(1e-12, np.abs(b).max()))


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 16), (32, 32), (32, 64), (64, 32), (64, 64), (64, 128),
                                      (128, 64), (128, 128)])
def test_tc_subm_matches_oracle(cuda, native_lib, cin, cout):
    from unidistill_b200 import sparse

    assert native_lib.ud_sparse_conv_tc_supported(27, cin, cout) == 1
    rng = np.random.default_rng(cin * 7 + cout)
    B, shape, n = 2, [7, 20, 20], 1931                     # 1931 rows: 15 full tiles + a ragged tail
    coords, feat = _cloud(rng, B, shape, n, cin)
    feat *= np.exp(rng.uniform(-6, 6, size=(n, 1))).astype(np.float32)   # wide dynamic range across rows
    w = (rng.standard_normal((cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)
    bias = rng.standard_normal((cout,)).astype(np.float32)
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    f, wt, bt = torch.from_numpy(feat).to(cuda), torch.from_numpy(w).to(cuda), torch.from_numpy(bias).to(cuda)
    assert sparse.USE_TENSOR_CORES
    out_tc = sparse.sparse_conv(f, wt, bt, rb).cpu().numpy()
    sparse.USE_TENSOR_CORES = False
    try:
        out_simt = sparse.sparse_conv(f, wt, bt, rb).cpu().numpy()
    finally:
        sparse.USE_TENSOR_CORES = True
    rp, rpn = O.rulebook_subm(coords, shape, 3)
    ref = O.sparse_conv_c(feat, w, bias, rp, rpn, n)
    assert _rel(out_simt, ref) < 1e-5
    assert _rel(out_tc, ref) < TC_TOL, "tensor-core path rel err %g" % _rel(out_tc, ref)
    # row-wise check: every row individually accurate relative to its own magnitude (no cross-row leakage)
    row_scale = np.abs(ref).max(axis=1) + 1e-30
    assert float((np.abs(out_tc - ref).max(axis=1) / row_scale).max()) < 1e-3


@pytest.mark.parametrize("cin,cout,ksize,stride,pad", [(16, 32, 3, 2, 1), (32, 64, 3, 2, 1), (64, 128, 3, 2, (0, 1, 1)),
                                                       (128, 128, (3, 1, 1), (2, 1, 1), 0)])
def test_tc_strided_matches_oracle(cuda, native_lib, cin, cout, ksize, stride, pad):
    from unidistill_b200 import sparse

    rng = np.random.default_rng(17)
    B, shape, n = 2, [11, 26, 22], 3000
    coords, feat = _cloud(rng, B, shape, n, cin)
    ks = O._t3(ksize)
    w = (rng.standard_normal((cout, ks[0], ks[1], ks[2], cin)) / np.sqrt(27 * cin)).astype(np.float32)
    rb = sparse.build_conv_rulebook(torch.from_numpy(coords).to(cuda), B, shape, ksize, stride, pad)
    out = sparse.sparse_conv(torch.from_numpy(feat).to(cuda), torch.from_numpy(w).to(cuda), None, rb)
    oc, rp, rpn, _ = O.rulebook_conv(coords, shape, ksize, stride, pad)
    ref = O.sparse_conv_c(feat, w, None, rp, rpn, oc.shape[0])
    assert _rel(out.cpu().numpy(), ref) < TC_TOL


@pytest.mark.parametrize("cin,cout", [(16, 16), (32, 64), (64, 64), (128, 128)])
def test_tc_fused_batchnorm_statistics(cuda, native_lib, cin, cout):
    """The epilogue's per-channel sum / sum of squares == a separate pass over the conv output (ragged tail rows
    and the bias included), and BatchNorm driven by them == nn.BatchNorm1d."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(cin + cout)
    B, shape, n = 2, [7, 24, 24], 2777
    coords, feat = _cloud(rng, B, shape, n, cin)
    w = torch.from_numpy((rng.standard_normal((cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)).to(cuda)
    bias = torch.from_numpy(rng.standard_normal((cout,)).astype(np.float32)).to(cuda)
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    out = sparse.sparse_conv(torch.from_numpy(feat).to(cuda), w, bias, rb)
    sums = getattr(out, "_ud_bn_sums", None)
    assert sums is not None and tuple(sums.shape) == (2, cout)
    ref0, ref1 = out.double().sum(0), (out.double() ** 2).sum(0)
    assert torch.allclose(sums[0], ref0, rtol=1e-5, atol=1e-4 * float(ref1.max().sqrt()))
    assert torch.allclose(sums[1], ref1, rtol=1e-5)
    # fixed-order reduction: a second launch on the same input reproduces the statistics bit for bit
    for _ in range(3):
        again = sparse.sparse_conv(torch.from_numpy(feat).to(cuda), w, bias, rb)
        assert torch.equal(again, out) and torch.equal(again._ud_bn_sums, sums)
    bn = torch.nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01).to(cuda).train()
    bn_ref = torch.nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01).to(cuda).train()
    y = sparse.batch_norm_act(out, bn, relu=True)
    y_ref = torch.relu(bn_ref(out.clone()))
    assert float((y - y_ref).abs().max()) < 1e-4
    assert torch.allclose(bn.running_var, bn_ref.running_var, rtol=1e-5, atol=1e-7)


def test_tc_single_row_and_isolated_voxels(cuda, native_lib):
    """One active voxel (only the centre tap exists) and voxels with no neighbours: out = in @ W[centre]."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(3)
    for n in (1, 5):
        coords = np.array([[0, 2 + 3 * i, 2 + 3 * i, 2 + 3 * i] for i in range(n)], np.int32)
        feat = rng.standard_normal((n, 64)).astype(np.float32)
        w = rng.standard_normal((64, 3, 3, 3, 64)).astype(np.float32)
        rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), 1, [20, 20, 20], 3)
        out = sparse.sparse_conv(torch.from_numpy(feat).to(cuda), torch.from_numpy(w).to(cuda), None, rb)
        ref = feat.astype(np.float64) @ w[:, 1, 1, 1, :].T.astype(np.float64)
        assert _rel(out.cpu().numpy(), ref.astype(np.float32)) < TC_TOL


def test_tc_linearity_full_size(cuda, native_lib):
    """Benchmark-sized layer (64->64 on ~180k voxels): conv(a*x + y) == a*conv(x) + conv(y), and the
    tensor-core result equals the SIMT kernel's -- size-independent checks where the oracle is too slow."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(23)
    B, shape, n = 4, [11, 256, 256], 180000
    coords, _ = _cloud(rng, B, shape, n, 1)
    c = torch.from_numpy(coords).to(cuda)
    rb = sparse.build_subm_rulebook(c, B, shape, 3)
    g = torch.Generator(device=cuda).manual_seed(0)
    x = torch.randn((n, 64), device=cuda, generator=g)
    y = torch.randn((n, 64), device=cuda, generator=g)
    w = torch.randn((64, 3, 3, 3, 64), device=cuda, generator=g) / 40.0
    fx, fy = sparse.sparse_conv(x, w, None, rb), sparse.sparse_conv(y, w, None, rb)
    fxy = sparse.sparse_conv(2.5 * x + y, w, None, rb)
    lin = 2.5 * fx + fy
    assert float((fxy - lin).abs().max() / lin.abs().max()) < TC_TOL
    sparse.USE_TENSOR_CORES = False
    try:
        ref = sparse.sparse_conv(x, w, None, rb)
    finally:
        sparse.USE_TENSOR_CORES = True
    assert float((fx - ref).abs().max() / ref.abs().max()) < TC_TOL


# ---------------------------------------------------------------------------------------------------
# tensor-core weight gradient (csrc/sparse_conv_wgrad_umma.cu).  Operands are rounded to TF32 (to nearest, 2^-11 each),
# products are accumulated in fp32: measured 2.5e-4 .. 4.3e-4 relative to the largest |dW| (independent of the row
# count: error and value both grow as sqrt(rows)); asserted at north_star's 1e-3 for fp32 losses/gradients.
WGRAD_TOL = 1e-3


def _wgrad_both(sparse, f, g, rb, w_shape):
    """dW through the autograd function with the tensor-core kernel and with the fp32 SIMT kernel."""
    res = []
    for tc in (True, False):
        old = sparse.USE_TC_WGRAD
        sparse.USE_TC_WGRAD = tc
        try:
            w = torch.zeros(w_shape, device=f.device, requires_grad=True)
            out = sparse.sparse_conv(f, w, None, rb)
            (dw,) = torch.autograd.grad(out, [w], g)
        finally:
            sparse.USE_TC_WGRAD = old
        res.append(dw)
    return res


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 16), (32, 32), (32, 64), (64, 32), (64, 64), (64, 128),
                                      (128, 64), (128, 128), (32, 128), (128, 32)])
def test_tc_wgrad_subm_matches_oracle(cuda, native_lib, cin, cout):
    from unidistill_b200 import sparse

    assert native_lib.ud_sparse_conv_tc_wgrad_supported(27, cin, cout) == 1
    rng = np.random.default_rng(cin * 5 + cout)
    B, shape, n = 2, [7, 20, 20], 1931                     # 15 full 128-row tiles + a ragged tail (< 64 rows)
    coords, feat = _cloud(rng, B, shape, n, cin)
    gout = rng.standard_normal((n, cout)).astype(np.float32)
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    f, g = torch.from_numpy(feat).to(cuda), torch.from_numpy(gout).to(cuda)
    dw_tc, dw_simt = _wgrad_both(sparse, f, g, rb, (cout, 3, 3, 3, cin))
    # oracle: dW[k] = sum over pairs of in[i]^T dout[o]  (float64)
    rp, rpn = O.rulebook_subm(coords, shape, 3)
    ref = np.zeros((27, cin, cout))
    for k in range(27):
        m = int(rpn[k])
        if m:
            ref[k] = feat[rp[0, k, :m]].astype(np.float64).T @ gout[rp[1, k, :m]].astype(np.float64)
    ref = ref.transpose(2, 0, 1).reshape(cout, 3, 3, 3, cin)
    assert _rel(dw_simt.cpu().numpy(), ref) < 2e-5
    err = _rel(dw_tc.cpu().numpy(), ref)
    assert err < WGRAD_TOL, "tensor-core wgrad rel err %g" % err
    # no systematic shrinkage (what truncating instead of rounding the TF32 operands would produce)
    num = float((dw_tc.double().cpu() * torch.from_numpy(ref)).sum())
    den = float((torch.from_numpy(ref) ** 2).sum())
    assert abs(num / den - 1.0) < 1e-4


@pytest.mark.parametrize("cin,cout,ksize,stride,pad", [(16, 32, 3, 2, 1), (32, 64, 3, 2, 1), (64, 128, 3, 2, (0, 1, 1)),
                                                       (128, 128, (3, 1, 1), (2, 1, 1), 0)])
def test_tc_wgrad_strided_matches_simt(cuda, native_lib, cin, cout, ksize, stride, pad):
    from unidistill_b200 import sparse

    rng = np.random.default_rng(29)
    B, shape, n = 2, [11, 26, 22], 3000
    coords, feat = _cloud(rng, B, shape, n, cin)
    ks = O._t3(ksize)
    rb = sparse.build_conv_rulebook(torch.from_numpy(coords).to(cuda), B, shape, ksize, stride, pad)
    g = torch.from_numpy(rng.standard_normal((rb.n_out, cout)).astype(np.float32)).to(cuda)
    dw_tc, dw_simt = _wgrad_both(sparse, torch.from_numpy(feat).to(cuda), g, rb, (cout, ks[0], ks[1], ks[2], cin))
    assert float((dw_tc - dw_simt).abs().max() / dw_simt.abs().max()) < WGRAD_TOL


def test_tc_wgrad_isolated_voxels_and_single_row(cuda, native_lib):
    """Only the centre tap exists: dW[centre] = in^T dout, every other tap exactly zero."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(4)
    for n in (1, 5):
        coords = np.array([[0, 2 + 3 * i, 2 + 3 * i, 2 + 3 * i] for i in range(n)], np.int32)
        feat = rng.standard_normal((n, 64)).astype(np.float32)
        gout = rng.standard_normal((n, 64)).astype(np.float32)
        rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), 1, [20, 20, 20], 3)
        dw_tc, _ = _wgrad_both(sparse, torch.from_numpy(feat).to(cuda), torch.from_numpy(gout).to(cuda), rb,
                               (64, 3, 3, 3, 64))
        dw = dw_tc.cpu().numpy()                            # (cout, kz, ky, kx, cin)
        ref = gout.astype(np.float64).T @ feat.astype(np.float64)
        assert _rel(dw[:, 1, 1, 1, :], ref.astype(np.float32)) < WGRAD_TOL
        dw[:, 1, 1, 1, :] = 0
        assert not dw.any()


def test_tc_wgrad_linearity_full_size(cuda, native_lib):
    """Benchmark-sized layer (64->64 on ~180k voxels): dW is linear in dout and equals the SIMT kernel's."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(31)
    B, shape, n = 4, [11, 256, 256], 180000
    coords, _ = _cloud(rng, B, shape, n, 1)
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    gen = torch.Generator(device=cuda).manual_seed(0)
    x = torch.randn((n, 64), device=cuda, generator=gen)
    g1 = torch.randn((n, 64), device=cuda, generator=gen)
    g2 = torch.randn((n, 64), device=cuda, generator=gen)
    shp = (64, 3, 3, 3, 64)
    d1, ref1 = _wgrad_both(sparse, x, g1, rb, shp)
    d2, _ = _wgrad_both(sparse, x, g2, rb, shp)
    d12, _ = _wgrad_both(sparse, x, 2.0 * g1 + g2, rb, shp)
    lin = 2.0 * d1 + d2
    assert float((d12 - lin).abs().max() / lin.abs().max()) < WGRAD_TOL
    assert float((d1 - ref1).abs().max() / ref1.abs().max()) < WGRAD_TOL


def _window_vs_direct(cuda, coords, feat, B, shape, cout, bias=True):
    from unidistill_b200 import sparse

    rng = np.random.default_rng(5)
    cin = feat.shape[1]
    w = torch.from_numpy((rng.standard_normal((cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)).to(cuda)
    b = torch.from_numpy(rng.standard_normal((cout,)).astype(np.float32)).to(cuda) if bias else None
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    f = torch.from_numpy(feat).to(cuda)
    old = sparse.USE_WINDOW_KERNEL, sparse.PRE_SPLIT_INPUT
    try:
        sparse.USE_WINDOW_KERNEL, sparse.PRE_SPLIT_INPUT = True, False
        out_w = sparse.sparse_conv(f, w, b, rb)
        sparse.USE_WINDOW_KERNEL = False
        out_d = sparse.sparse_conv(f, w, b, rb)
        sparse.PRE_SPLIT_INPUT = True                # rows split once by ud_split_bf16: same hi/lo values, same products
        out_s = sparse.sparse_conv(f, w, b, rb)
        assert torch.equal(out_s, out_d), "pre-split input path differs from the in-kernel split"
        assert torch.equal(out_s._ud_bn_sums, out_d._ud_bn_sums)
    finally:
        sparse.USE_WINDOW_KERNEL, sparse.PRE_SPLIT_INPUT = old
    return out_w, out_d, rb, w, b


def test_window_kernel_equals_direct_gather_kernel(cuda, native_lib):
    """k_spconv_umma_win (TMA-staged windows, unit-major schedule) == k_spconv_umma_dg bit for bit: the same bf16 hi/lo
    products are accumulated per row in the same tap order; also the fused BatchNorm statistics and the C oracle."""
    assert native_lib.ud_sparse_conv_tc_window_supported(27, 64, 64) == 1
    rng = np.random.default_rng(12)
    for (B, shape, n, order) in [(2, [9, 40, 44], 9000, "raster"), (2, [9, 40, 44], 9000, "shuffled"), (1, [5, 20, 20], 130, "raster"),
                                 (3, [7, 30, 30], 4097, "raster")]:
        lin = rng.choice(B * shape[0] * shape[1] * shape[2], size=n, replace=False)
        if order == "raster":
            lin.sort()                                          # the layout the windows are meant for
        b_, r = np.divmod(lin, shape[0] * shape[1] * shape[2])
        z, r = np.divmod(r, shape[1] * shape[2])
        y, x = np.divmod(r, shape[2])
        coords = np.stack([b_, z, y, x], 1).astype(np.int32)
        feat = rng.standard_normal((n, 64)).astype(np.float32)
        out_w, out_d, rb, w, b = _window_vs_direct(cuda, coords, feat, B, shape, 64)
        assert torch.equal(out_w, out_d), (order, float((out_w - out_d).abs().max()))
        assert torch.equal(out_w._ud_bn_sums, out_d._ud_bn_sums)
        rp, rpn = O.rulebook_subm(coords, shape, 3)
        ref = O.sparse_conv_c(feat, w.cpu().numpy(), b.cpu().numpy(), rp, rpn, n)
        assert _rel(out_w.cpu().numpy(), ref) < TC_TOL


def test_window_kernel_full_size_sheet(cuda, native_lib):
    """A LiDAR-like sheet at the size of the bench's 64-channel stage (many tiles per CTA, several waves)."""
    rng = np.random.default_rng(13)
    B, shape = 2, [11, 256, 256]
    yy, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    pts = []
    for b_ in range(B):
        z = (4 + 2 * np.sin(xx / 37.0 + b_) + 2 * np.cos(yy / 29.0)).astype(np.int32)
        for dz in (0, 1):
            keep = rng.random(z.shape) < 0.8
            pts.append(np.stack([np.full_like(z, b_), z + dz, yy, xx], -1)[keep])
    coords = np.unique(np.concatenate(pts, 0), axis=0).astype(np.int32)       # sorted = raster order
    feat = rng.standard_normal((coords.shape[0], 64)).astype(np.float32)
    out_w, out_d, *_ = _window_vs_direct(cuda, coords, feat, B, shape, 64, bias=False)
    assert torch.equal(out_w, out_d)


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128), (128, 128)])
def test_pre_split_input_is_bit_identical(cuda, native_lib, cin, cout):
    """ud_split_bf16 + the SPLIT instantiation of the kernel == the in-kernel bf16 hi/lo split (same rounding, same products)."""
    from unidistill_b200 import sparse

    rng = np.random.default_rng(cin + 3 * cout)
    B, shape, n = 2, [7, 22, 22], 2300
    coords, feat = _cloud(rng, B, shape, n, cin)
    feat *= np.exp(rng.uniform(-8, 8, size=(n, 1))).astype(np.float32)
    w = torch.from_numpy((rng.standard_normal((cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)).to(cuda)
    rb = sparse.build_subm_rulebook(torch.from_numpy(coords).to(cuda), B, shape, 3)
    f = torch.from_numpy(feat).to(cuda)
    old = sparse.PRE_SPLIT_INPUT
    try:
        sparse.PRE_SPLIT_INPUT = True
        a = sparse.sparse_conv(f, w, None, rb)
        sparse.PRE_SPLIT_INPUT = False
        b = sparse.sparse_conv(f, w, None, rb)
    finally:
        sparse.PRE_SPLIT_INPUT = old
    assert torch.equal(a, b)


def _conv_digest_script():
    return r"""
import hashlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.getcwd(), "cvpr2023-unidistill_b200"))
from unidistill_b200 import sparse as SP
torch.manual_seed(3)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(4)
coords = torch.stack([torch.zeros(60000, dtype=torch.long), torch.randint(0, 21, (60000,), generator=g),
                      torch.randint(0, 96, (60000,), generator=g), torch.randint(0, 96, (60000,), generator=g)], 1)
coords = torch.unique(coords, dim=0).int().to(dev).contiguous()
rb = SP.build_subm_rulebook(coords, 1, [21, 96, 96], 3)
h = hashlib.sha256()
for cin, cout in ((16, 16), (32, 32), (64, 64), (64, 128), (128, 128)):
    x = torch.randn(coords.shape[0], cin, device=dev)
    w = torch.randn(cout, 3, 3, 3, cin, device=dev) * 0.05
    y = SP.sparse_conv(x, w, None, rb, False)
    h.update(y.cpu().numpy().tobytes())
print("DIGEST", h.hexdigest())
"""


@pytest.mark.gpu
def test_one_and_two_ctas_per_sm_give_identical_outputs(cuda, native_lib):
    """k_spconv_umma_dg with NSETS = 4 (one CTA per SM, 512 TMEM columns) and NSETS = 2 (two CTAs per SM, the default):
    the accumulation order per output tile is the same, so every output bit is.  The configuration is read once per process
    (UD_DG_NSETS), hence two subprocesses."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    digests = {}
    for nsets in ("2", "4"):
        env = dict(os.environ, UD_DG_NSETS=nsets)
        r = subprocess.run([sys.executable, "-c", _conv_digest_script()], cwd=root, env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        digests[nsets] = [ln for ln in r.stdout.splitlines() if ln.startswith("DIGEST")][0]
    assert digests["2"] == digests["4"]


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 64), (64, 64), (128, 64), (128, 128)])
def test_packing_from_the_parameter_layout_is_bit_identical(cuda, native_lib, cin, cout):
    """ud_sparse_conv_tc_pack_weights_param (reads the spconv-2 parameter (Cout, kz, ky, kx, Cin) in place; mode 0 forward,
    1 dgrad of a SubM convolution = transposed + mirrored taps, 2 dgrad of a strided one = transposed) against
    ud_sparse_conv_tc_pack_weights on the explicitly permuted / flipped copies."""
    from unidistill_b200 import native, sparse as SP

    lib = native.load()
    torch.manual_seed(cin * 7 + cout)
    K = 27
    w = torch.randn(cout, 3, 3, 3, cin, device=cuda)
    w_k = SP.weight_to_taps(w)                                             # (K, cin, cout)
    sp = native.stream_ptr(cuda)

    def pack_ref(wk, ci, co):
        n = native.query_bytes("ud_sparse_conv_tc_weight_bytes", K, ci, co)
        out = torch.zeros((n,), dtype=torch.uint8, device=cuda)
        native.check(lib.ud_sparse_conv_tc_pack_weights(native.ptr(wk.contiguous()), K, ci, co, native.ptr(out), sp), "pack")
        return out

    def pack_param(mode, ci, co):
        n = native.query_bytes("ud_sparse_conv_tc_weight_bytes", K, ci, co)
        out = torch.zeros((n,), dtype=torch.uint8, device=cuda)
        native.check(lib.ud_sparse_conv_tc_pack_weights_param(native.ptr(w), K, ci, co, mode, native.ptr(out), sp), "pack_param")
        return out

    assert torch.equal(pack_param(0, cin, cout), pack_ref(w_k, cin, cout))
    if lib.ud_sparse_conv_tc_supported(K, cout, cin):
        assert torch.equal(pack_param(1, cout, cin), pack_ref(w_k.flip(0).transpose(1, 2), cout, cin))
        assert torch.equal(pack_param(2, cout, cin), pack_ref(w_k.transpose(1, 2), cout, cin))
