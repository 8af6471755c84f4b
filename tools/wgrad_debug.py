"""Development aid: tensor-core wgrad vs the fp32 SIMT wgrad, per-tap errors and timings (B200)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))

from unidistill_b200 import native, sparse  # noqa: E402


def cloud(rng, B, shape, n):
    D, H, W = shape
    lin = np.sort(rng.choice(B * D * H * W, size=n, replace=False))
    b, r = np.divmod(lin, D * H * W)
    z, r = np.divmod(r, H * W)
    y, x = np.divmod(r, W)
    return np.stack([b, z, y, x], 1).astype(np.int32)


def wgrad(f, g, rb, cin, cout, tc):
    lib = native.load()
    K = rb.K
    dw = torch.zeros((K, cin, cout), device=f.device)
    sp = native.stream_ptr(f.device)
    if tc:
        tt, masks = rb.tap_major(rb.nbr_out)
        rc = lib.ud_sparse_conv_tc_wgrad(native.ptr(f), native.ptr(g), native.ptr(tt), native.ptr(masks), rb.n_out, K,
                                         cin, cout, native.ptr(dw), sp)
    else:
        rc = lib.ud_sparse_conv_wgrad(native.ptr(f), native.ptr(g), native.ptr(rb.nbr_out), rb.n_out, K, cin, cout,
                                      native.ptr(dw), sp)
    native.check(rc, "wgrad")
    return dw


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    t00 = time.time()
    small = [] if os.environ.get("UD_WG_SKIP_SMALL") else [(64, 64), (32, 32), (128, 128), (64, 128), (32, 64)]
    for (cin, cout) in small:
        for (B, shape, n) in [(1, [5, 12, 12], 300), (2, [7, 20, 20], 1931), (2, [9, 40, 40], 9000), (4, [11, 128, 128], 100000)]:
            coords = torch.from_numpy(cloud(rng, B, shape, n)).to(dev)
            rb = sparse.build_subm_rulebook(coords, B, shape, 3)
            f = torch.randn((n, cin), device=dev)
            g = torch.randn((n, cout), device=dev)
            a = wgrad(f, g, rb, cin, cout, True)
            b = wgrad(f, g, rb, cin, cout, False)
            torch.cuda.synchronize()
            err = (a - b).abs().amax(dim=(1, 2)) / b.abs().max()
            print("%3d->%3d n=%5d  max|tc| %.4g  max|simt| %.4g  rel err %.3e  worst taps %s" % (
                cin, cout, n, float(a.abs().max()), float(b.abs().max()), float(err.max()),
                [(int(i), "%.1e" % float(err[i])) for i in err.argsort(descending=True)[:3]]), flush=True)
    # timing at benchmark size
    B, shape, n = 4, [21, 512, 512], int(os.environ.get("UD_WG_N", "700000"))
    coords = torch.from_numpy(cloud(rng, B, shape, n)).to(dev)
    rb = sparse.build_subm_rulebook(coords, B, shape, 3)
    P = rb.num_pairs()
    print("full-size rulebook ready, P=%d, t=%.1fs" % (P, time.time() - t00), flush=True)
    for (cin, cout) in [(32, 32), (64, 64), (128, 128)]:
        f = torch.randn((n, cin), device=dev)
        g = torch.randn((n, cout), device=dev)
        for tc in (True, False):
            wgrad(f, g, rb, cin, cout, tc)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                dw = wgrad(f, g, rb, cin, cout, tc)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print("wgrad %3d->%3d N=%d P=%d %s: %.3f ms  %.1f TFLOP/s" % (cin, cout, n, P, "tcgen05" if tc else "simt   ",
                                                                          ms, 2.0 * P * cin * cout / ms / 1e9), flush=True)


if __name__ == "__main__":
    main()
