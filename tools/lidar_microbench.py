"""Per-stage CUDA-event timings of the LiDAR branch at the BASELINE config (bs=4, 40k x 10 sweeps, 128^2 BEV).

Development aid (not the bench contract): prints one line per stage with ms and the algorithmic
bytes / flops of SURVEY 8d so that achieved GB/s / TFLOP/s can be read off directly.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))

from unidistill_b200 import native, sparse, synthetic as S  # noqa: E402
from unidistill_b200.voxelize import hard_voxelize  # noqa: E402


def timed(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    t = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return float(np.median(t)), r


def main():
    dev = torch.device("cuda:0")
    bs = int(os.environ.get("UD_BS", "4"))
    batch = S.lidar_batch(bs, seed=0)
    pts = [batch["points"][b].to(dev) for b in range(bs)]
    n_pts = sum(p.shape[0] for p in pts)
    rows = []

    ms, (vox, coords, num, mean, counts) = timed(
        lambda: hard_voxelize(pts, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000))
    M = coords.shape[0]
    rows.append(("voxelize+meanvfe(+voxels)", ms, n_pts * 20 + M * (200 + 20 + 16 + 4), 0))
    ms, _ = timed(lambda: hard_voxelize(pts, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000,
                                        want_voxels=False))
    rows.append(("voxelize+meanvfe(no voxels)", ms, n_pts * 20 + M * (20 + 16 + 4), 0))

    shape = [41, 1024, 1024]
    torch.manual_seed(0)
    feats = mean
    plan = [("subm", 5, 16, None), ("subm", 16, 16, None), ("down", 16, 32, (3, 2, 1)), ("subm", 32, 32, None),
            ("down", 32, 64, (3, 2, 1)), ("subm", 64, 64, None), ("down", 64, 128, (3, 2, (0, 1, 1))),
            ("subm", 128, 128, None), ("down", 128, 128, ((3, 1, 1), (2, 1, 1), 0))]
    cur_coords, cur_shape = coords, shape
    rb = None
    only = os.environ.get("UD_MB_ONLY")          # e.g. "64,64": time just that layer, forward only (for ncu)
    only = tuple(int(v) for v in only.split(",")) if only else None
    for kind, cin, cout, geo in plan:
        if kind == "subm":
            if rb is None or not rb.subm or rb.n_in != cur_coords.shape[0]:
                ms, rb = timed(lambda: sparse.build_subm_rulebook(cur_coords, bs, cur_shape, 3))
                rows.append(("rulebook subm N=%d" % cur_coords.shape[0], ms,
                             cur_coords.shape[0] * 16 + rb.n_out * 27 * 4, 0))
        else:
            ks, st, pd = geo
            ms, rb = timed(lambda: sparse.build_conv_rulebook(cur_coords, bs, cur_shape, ks, st, pd))
            rows.append(("rulebook conv N=%d->%d" % (rb.n_in, rb.n_out), ms,
                         rb.n_in * 16 + (rb.n_in + rb.n_out) * rb.K * 4 + rb.n_out * 16, 0))
        if only is not None and (cin, cout) != only:
            if kind == "down":
                cur_coords, cur_shape = rb.out_coords, rb.out_shape
                rb = None
            continue
        P = rb.num_pairs()
        x = torch.randn((rb.n_in, cin), device=dev)
        ks3 = rb.ksize
        w = torch.randn((cout, ks3[0], ks3[1], ks3[2], cin), device=dev) * 0.05
        with torch.no_grad():
            ms, out = timed(lambda: sparse.sparse_conv(x, w, None, rb))
        byt = rb.n_in * cin * 4 + rb.n_out * cout * 4 + rb.n_out * rb.K * 4 + rb.K * cin * cout * 4
        rows.append(("conv %s %d->%d N=%d P=%d" % (kind, cin, cout, rb.n_out, P), ms, byt, 2.0 * P * cin * cout))
        if sparse.USE_TENSOR_CORES and native.load().ud_sparse_conv_tc_supported(rb.K, cin, cout):
            sparse.USE_TENSOR_CORES = False
            with torch.no_grad():
                ms2, out2 = timed(lambda: sparse.sparse_conv(x, w, None, rb))
            sparse.USE_TENSOR_CORES = True
            rows.append(("   (same, fp32 SIMT kernel; max |d| %.2e)" % float((out - out2).abs().max()), ms2, byt,
                         2.0 * P * cin * cout))
        if cout >= 16:
            bn = torch.nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01).to(dev)
            with torch.no_grad():
                ms, _ = timed(lambda: sparse.batch_norm_act(out, bn, relu=True))
            rows.append(("bn+relu(train) N=%d C=%d" % (out.shape[0], cout), ms, out.numel() * 4 * 3, 0))
        if only is not None:
            if kind == "down":
                cur_coords, cur_shape = rb.out_coords, rb.out_shape
                rb = None
            continue
        # backward kernels
        xg = x.clone().requires_grad_(True)
        wg = w.clone().requires_grad_(True)
        o = sparse.sparse_conv(xg, wg, None, rb)
        g = torch.randn_like(o)
        ms, _ = timed(lambda: torch.autograd.grad(o, [xg, wg], g, retain_graph=True))
        rows.append(("conv bwd (dgrad+wgrad) %d->%d" % (cin, cout), ms, 2 * byt, 4.0 * P * cin * cout))
        if kind == "down":
            cur_coords, cur_shape = rb.out_coords, rb.out_shape
            rb_prev = rb
            rb = None
    feat_out = torch.randn((cur_coords.shape[0], 128), device=dev)
    ms, bev = timed(lambda: sparse.scatter_to_dense(feat_out, cur_coords, bs, cur_shape))
    rows.append(("scatter-to-BEV N=%d" % cur_coords.shape[0], ms, feat_out.numel() * 4 + bev.numel() * 4, 0))

    report = []
    print("%-46s %9s %10s %10s" % ("stage", "ms", "GB/s", "TFLOP/s"))
    for name, ms, byt, fl in rows:
        gbs = byt / ms / 1e6
        tf = fl / ms / 1e9
        print("%-46s %9.3f %10.1f %10.2f" % (name, ms, gbs, tf))
        report.append({"stage": name, "ms": ms, "bytes": byt, "flops": fl, "GBps": gbs, "TFLOPs": tf})
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "lidar_microbench.json"), "w") as f:
        json.dump({"bs": bs, "n_points": n_pts, "voxels": int(M), "rows": report}, f, indent=1)


if __name__ == "__main__":
    main()
