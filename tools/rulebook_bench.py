"""Rulebook sequence of VoxelResBackBone8x on the bench batch (bs = 4): voxelize, then SubM / strided builders stage by
stage, each timed with CUDA events (median of `reps`).  Development aid; also the command profiled with ncu."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
from unidistill_b200 import native, sparse as SP, synthetic as S, voxelize as V  # noqa: E402

reps = int(os.environ.get("REPS", "7"))
dev = torch.device("cuda:0")
lid = S.lidar_batch(4, seed=0)
pts = [p.to(dev) for p in lid["points"]]
out = V.hard_voxelize(pts, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000, want_voxels=False)
coords = out[1].int().contiguous()
grid = [41, 1024, 1024]          # (z, y, x) of the sparse tensor (VoxelResBackBone8x: grid_size[::-1] + [1, 0, 0])
print("voxels", coords.shape[0])


def timed(fn):
    ts = []
    r = None
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return r, statistics.median(ts)


total = 0.0
shape = grid
stages = [("subm1", None), ("down2", ([3, 3, 3], [2, 2, 2], [1, 1, 1])), ("down3", ([3, 3, 3], [2, 2, 2], [1, 1, 1])),
          ("down4", ([3, 3, 3], [2, 2, 2], [0, 1, 1])), ("out", ([3, 1, 1], [2, 1, 1], [0, 0, 0]))]
for name, conv in stages:
    if conv is not None:
        rb, ms = timed(lambda: SP.build_conv_rulebook(coords, 4, shape, conv[0], conv[1], conv[2]))
        rb.tap_major(rb.nbr_out)
        _, ms_t = timed(lambda: (rb._tap_major.pop(rb.nbr_out.data_ptr(), None), rb.tap_major(rb.nbr_out)))
        print("%-6s strided %s: n_in %7d -> n_out %7d  %.3f ms (+ transpose of nbr_out %.3f ms, incl. one host sync)" % (
            name, conv[0], rb.n_in, rb.n_out, ms, ms_t))
        total += ms + ms_t
        coords, shape = rb.out_coords, rb.out_shape
    if name != "out":
        rb, ms = timed(lambda: SP.build_subm_rulebook(coords, 4, shape, 3))
        print("%-6s SubM 3x3x3 on %s: n %7d  %.3f ms" % (name, shape, rb.n_in, ms))
        total += ms
print("rulebooks of one encoder pass: %.3f ms" % total)
