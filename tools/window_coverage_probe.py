"""Probe for the window-staged sparse convolution: for every 128-row tile and every (kz, ky) tap group of the SubM rulebooks
of the bench clouds, the rows the three kx taps gather lie in a narrow index window under the raster row order.  Prints, per
stage, the share of gathered rows inside [lo, lo + W) with lo = the smallest gathered row of the (tile, group)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
from unidistill_b200 import sparse, synthetic as S  # noqa: E402
from unidistill_b200.voxelize import hard_voxelize  # noqa: E402

dev = torch.device("cuda:0")
batch = S.lidar_batch(4, seed=0)
pts = [batch["points"][b].to(dev) for b in range(4)]
_, coords, num, mean, _ = hard_voxelize(pts, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000, want_voxels=False)
shape = [41, 1024, 1024]
cur = coords
for stage, (ks, st, pd) in enumerate([(None, None, None), (3, 2, 1), (3, 2, 1), (3, 2, (0, 1, 1))]):
    if ks is not None:
        rb = sparse.build_conv_rulebook(cur, 4, shape, ks, st, pd)
        cur, shape = rb.out_coords, rb.out_shape
    rb = sparse.build_subm_rulebook(cur, 4, shape, 3)
    t = rb.nbr_out                                             # (n, 27)
    n = t.shape[0]
    nt = n // 128
    g = t[: nt * 128].view(nt, 128, 9, 3).permute(0, 2, 1, 3).reshape(nt, 9, 384).long()
    valid = g >= 0
    big = torch.where(valid, g, torch.full_like(g, 1 << 40))
    lo = big.min(dim=2, keepdim=True).values
    line = "stage %d: N=%d  pairs/row %.1f |" % (stage + 1, n, float(valid.sum()) / (nt * 128))
    for W in (128, 144, 160, 192, 256):
        inside = (valid & (g - lo < W)).sum().item() / max(1, valid.sum().item())
        line += "  W=%d: %.4f" % (W, inside)
    span = torch.where(valid, g, torch.full_like(g, -1)).max(dim=2).values - lo.squeeze(2) + 1
    has = valid.any(dim=2)
    line += " | span p50 %d p90 %d p99 %d" % tuple(int(np.percentile(span[has].cpu().numpy(), q)) for q in (50, 90, 99))
    print(line, flush=True)
