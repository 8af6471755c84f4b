"""CPU probe (numpy) for DESIGN.md section 9: how many DISTINCT input rows does a 128-row output tile of a SubM 3x3x3
convolution reference, under the current row order (first-come in point order) and under a Morton (z-order) sort of the
voxels?  pairs/tile is what the gather kernels fetch today (one row per (output row, tap) pair); distinct/tile is what a
kernel that stages a tile's neighbourhood once would fetch."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))

from oracle import lidar as O  # noqa: E402
from unidistill_b200 import synthetic as S  # noqa: E402


def part1by2(x):
    x = x.astype(np.uint64) & 0x1FFFFF
    x = (x | (x << 32)) & 0x1F00000000FFFF
    x = (x | (x << 16)) & 0x1F0000FF0000FF
    x = (x | (x << 8)) & 0x100F00F00F00F00F
    x = (x | (x << 4)) & 0x10C30C30C30C30C3
    x = (x | (x << 2)) & 0x1249249249249249
    return x


def morton(coords):
    return (part1by2(coords[:, 3]) | (part1by2(coords[:, 2]) << 1) | (part1by2(coords[:, 1]) << 2)) | \
        (coords[:, 0].astype(np.uint64) << 60)


def stats(coords, shape, name):
    """coords (N,4) [b,z,y,x] in the given row order."""
    D, H, W = shape
    lin = ((coords[:, 0].astype(np.int64) * D + coords[:, 1]) * H + coords[:, 2]) * W + coords[:, 3]
    order = np.argsort(lin)
    lin_sorted = lin[order]
    n = coords.shape[0]
    nbr = np.full((n, 27), -1, np.int64)
    k = 0
    for dz in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                z, y, x = coords[:, 1] + dz, coords[:, 2] + dy, coords[:, 3] + dx
                ok = (z >= 0) & (z < D) & (y >= 0) & (y < H) & (x >= 0) & (x < W)
                q = ((coords[:, 0].astype(np.int64) * D + z) * H + y) * W + x
                pos = np.searchsorted(lin_sorted, q)
                pos = np.minimum(pos, n - 1)
                hit = ok & (lin_sorted[pos] == q)
                nbr[hit, k] = order[pos[hit]]
                k += 1
    tiles = (n + 127) // 128
    pairs = distinct = 0
    for t in range(tiles):
        blk = nbr[t * 128: (t + 1) * 128]
        v = blk[blk >= 0]
        pairs += v.size
        distinct += np.unique(v).size
    print("%-46s N=%7d  pairs/row %5.2f  distinct rows per 128-row tile %6.1f  (gather reduction %.2fx)" % (
        name, n, pairs / n, distinct / tiles, pairs / max(1, distinct)))


def real_layer_sets():
    """The ACTUAL active sets of the sparse encoder (strided convolutions dilate them): bs=1, via the rulebook oracle."""
    batch = S.lidar_batch(1, seed=0)
    _, coords, _, _ = O.voxelize([batch["points"][0].numpy()], S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000)
    shape = [41, 1024, 1024]
    sets = [("stage 1 (subm1/res1, C=16)", coords, shape)]
    cur, cur_shape = coords, shape
    for name, pad in (("stage 2 (res2, C=32)", 1), ("stage 3 (res3, C=64)", 1), ("stage 4 (res4, C=128)", (0, 1, 1))):
        out_coords, _, _, out_shape = O.rulebook_conv(cur, cur_shape, 3, 2, pad)
        cur, cur_shape = np.ascontiguousarray(out_coords), [int(v) for v in out_shape]
        sets.append((name, cur, cur_shape))
    return sets


def main():
    if os.environ.get("UD_REAL_SETS", "1") != "0":
        print("active sets of the real encoder (bs=1, strided convs dilate the sets):")
        for name, c, sh in real_layer_sets():
            stats(c, sh, name + ", current order")
            stats(c[np.argsort(morton(c), kind="stable")], sh, name + ", Morton order")
        print()
    bs = int(os.environ.get("UD_BS", "2"))
    batch = S.lidar_batch(bs, seed=0)
    clouds = [batch["points"][b].numpy() for b in range(bs)]
    _, coords, _, _ = O.voxelize(clouds, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000)
    shape = [41, 1024, 1024]
    for level, name in ((0, "stage 1 (0.1 m voxels)"), (1, "stage 2 (stride 2)"), (2, "stage 3 (stride 4)"), (3, "stage 4 (stride 8)")):
        c = coords.copy()
        c[:, 1:] >>= level
        c = np.unique(c, axis=0)                    # occupancy of the coarser grid (a subset of the strided conv's outputs)
        sh = [max(1, (shape[0] + (1 << level) - 1) >> level), shape[1] >> level, shape[2] >> level]
        if level == 0:
            c = coords                              # keep the voxelizer's first-come order
            stats(c, sh, name + ", first-come order")
        else:
            stats(c, sh, name + ", ascending linear (b,z,y,x) order")
        m = c[np.argsort(morton(c), kind="stable")]
        stats(m, sh, name + ", Morton order")


if __name__ == "__main__":
    main()
