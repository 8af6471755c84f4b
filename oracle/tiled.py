"""oracle/tiled.py -- TEST INFRASTRUCTURE / design groundwork for DESIGN.md section 9 (lead 1).

CPU specification of the tile-local ("halo") rulebook a round-2 sparse-convolution kernel would consume: voxels sorted
along a Morton (z-order) curve, 128-row output tiles, for every tile the sorted list of DISTINCT input rows its 3x3x3
neighbourhood touches (the halo, staged once) and a (128, K) table of halo-local indices (0xFFFF = absent).  The
convolution evaluated through these tables must equal the plain rulebook convolution (oracle.lidar) -- that equality, the
halo statistics and the index-width bound are what tests/test_oracle_tiled.py pins.  Nothing here is imported by the product.
"""
import numpy as np

ABSENT = 0xFFFF


def _part1by2(x):
    x = x.astype(np.uint64) & np.uint64(0x1FFFFF)
    x = (x | (x << np.uint64(32))) & np.uint64(0x1F00000000FFFF)
    x = (x | (x << np.uint64(16))) & np.uint64(0x1F0000FF0000FF)
    x = (x | (x << np.uint64(8))) & np.uint64(0x100F00F00F00F00F)
    x = (x | (x << np.uint64(4))) & np.uint64(0x10C30C30C30C30C3)
    x = (x | (x << np.uint64(2))) & np.uint64(0x1249249249249249)
    return x


def morton_keys(coords):
    """coords (N,4) int [b,z,y,x] -> uint64 keys: batch index in the top bits, then interleaved z,y,x."""
    c = np.asarray(coords)
    key = _part1by2(c[:, 3]) | (_part1by2(c[:, 2]) << np.uint64(1)) | (_part1by2(c[:, 1]) << np.uint64(2))
    return key | (c[:, 0].astype(np.uint64) << np.uint64(58))


def morton_order(coords):
    """Permutation that sorts the rows along the curve (stable)."""
    return np.argsort(morton_keys(coords), kind="stable")


def subm_table(coords, shape, ksize=3):
    """(N, K) int64 gather table of a SubM convolution (same conventions as oracle.lidar.rulebook_subm: tap index
    k = (kz*KY + ky)*KX + kx, -1 = absent), vectorised."""
    c = np.asarray(coords).astype(np.int64)
    D, H, W = [int(s) for s in shape]
    n = c.shape[0]
    lin = ((c[:, 0] * D + c[:, 1]) * H + c[:, 2]) * W + c[:, 3]
    order = np.argsort(lin, kind="stable")
    lin_sorted = lin[order]
    r = ksize // 2
    K = ksize ** 3
    nbr = np.full((n, K), -1, np.int64)
    k = 0
    for dz in range(-r, r + 1):
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                z, y, x = c[:, 1] + dz, c[:, 2] + dy, c[:, 3] + dx
                ok = (z >= 0) & (z < D) & (y >= 0) & (y < H) & (x >= 0) & (x < W)
                q = ((c[:, 0] * D + z) * H + y) * W + x
                pos = np.minimum(np.searchsorted(lin_sorted, q), max(n - 1, 0))
                hit = ok & (lin_sorted[pos] == q) if n else ok
                nbr[hit, k] = order[pos[hit]]
                k += 1
    return nbr


def build_tile_halos(nbr, tile=128):
    """nbr (N, K) gather table (-1 absent) -> list over tiles of (halo_rows (h,) int64 ascending,
    local (rows_in_tile, K) uint16 with ABSENT for missing taps)."""
    n, K = nbr.shape
    out = []
    for t0 in range(0, n, tile):
        blk = nbr[t0: t0 + tile]
        halo = np.unique(blk[blk >= 0])
        if halo.size >= ABSENT:
            raise ValueError("halo of %d rows does not fit 16-bit local indices" % halo.size)
        local = np.full(blk.shape, ABSENT, np.uint16)
        m = blk >= 0
        local[m] = np.searchsorted(halo, blk[m]).astype(np.uint16)
        out.append((halo, local))
    return out


def tiled_conv(feat, w_taps, halos, bias=None):
    """out[o] = bias + sum_k halo_feat[local[o, k]] @ w_taps[k]; feat (N, Cin), w_taps (K, Cin, Cout) -> (N, Cout) float64."""
    K, cin, cout = w_taps.shape
    rows = sum(local.shape[0] for _, local in halos)
    out = np.zeros((rows, cout))
    r0 = 0
    for halo, local in halos:
        staged = feat[halo].astype(np.float64)              # the tile's neighbourhood, fetched once
        acc = np.zeros((local.shape[0], cout))
        for k in range(K):
            m = local[:, k] != ABSENT
            if m.any():
                acc[m] += staged[local[m, k]] @ w_taps[k].astype(np.float64)
        out[r0: r0 + local.shape[0]] = acc
        r0 += local.shape[0]
    if bias is not None:
        out += bias.astype(np.float64)
    return out


def halo_stats(halos, nbr, tile=128):
    """-> dict(mean_halo, max_halo, pairs_per_tile, reduction): rows a tile needs vs rows it gathers pair by pair."""
    sizes = np.array([h.size for h, _ in halos])
    pairs = int((nbr >= 0).sum())
    return {"tiles": len(halos), "mean_halo": float(sizes.mean()) if len(halos) else 0.0,
            "max_halo": int(sizes.max()) if len(halos) else 0, "pairs_per_tile": pairs / max(1, len(halos)),
            "reduction": pairs / max(1, int(sizes.sum()))}
