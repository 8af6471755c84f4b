"""Design groundwork for the round-2 tile-local sparse convolution: the halo formulation of
oracle/tiled.py computes the same convolution as the plain rulebook oracle, Morton ordering shrinks the halos, and
16-bit halo-local indices suffice."""
import numpy as np

from oracle import lidar as O
from oracle import tiled as T


def _cloud(rng, B, shape, n):
    D, H, W = shape
    # clustered occupancy (ground-plane-like slabs) so that neighbourhoods are populated
    z = rng.integers(0, 3, n)
    y = rng.integers(0, H, n)
    x = rng.integers(0, W, n)
    b = rng.integers(0, B, n)
    c = np.unique(np.stack([b, z, y, x], 1), axis=0).astype(np.int32)
    return c[rng.permutation(c.shape[0])]


def test_vectorised_table_equals_rulebook_oracle():
    rng = np.random.default_rng(0)
    shape = [5, 40, 40]
    coords = _cloud(rng, 2, shape, 2500)
    nbr = T.subm_table(coords, shape)
    rp, rpn = O.rulebook_subm(coords, shape, 3)
    ref = np.full(nbr.shape, -1, np.int64)
    for k in range(27):
        m = int(rpn[k])
        ref[rp[1, k, :m], k] = rp[0, k, :m]
    assert np.array_equal(nbr, ref)


def test_tiled_conv_equals_plain_conv_and_morton_shrinks_halos():
    rng = np.random.default_rng(1)
    shape = [5, 64, 64]
    coords = _cloud(rng, 2, shape, 6000)
    n = coords.shape[0]
    feat = rng.standard_normal((n, 16)).astype(np.float32)
    w = (rng.standard_normal((8, 3, 3, 3, 16)) / 20).astype(np.float32)
    bias = rng.standard_normal((8,)).astype(np.float32)
    w_taps = w.reshape(8, 27, 16).transpose(1, 2, 0)
    rp, rpn = O.rulebook_subm(coords, shape, 3)
    ref = O.sparse_conv_c(feat, w, bias, rp, rpn, n)

    stats = {}
    for name, perm in (("given", np.arange(n)), ("morton", T.morton_order(coords))):
        c, f = coords[perm], feat[perm]
        nbr = T.subm_table(c, shape)
        halos = T.build_tile_halos(nbr)
        out = T.tiled_conv(f, w_taps, halos, bias)
        assert np.abs(out - ref[perm]).max() < 1e-4 * np.abs(ref).max()
        stats[name] = T.halo_stats(halos, nbr)
        assert stats[name]["max_halo"] < T.ABSENT
    assert stats["morton"]["mean_halo"] < 0.6 * stats["given"]["mean_halo"]
    assert stats["morton"]["reduction"] > 2.0
    # every row of a tile is in its own halo (centre tap), so a halo is never smaller than its tile
    assert stats["morton"]["mean_halo"] >= 127.0


def test_morton_keys_keep_samples_apart_and_are_a_permutation():
    rng = np.random.default_rng(2)
    coords = _cloud(rng, 3, [41, 1024, 1024], 4000)
    perm = T.morton_order(coords)
    assert np.array_equal(np.sort(perm), np.arange(coords.shape[0]))
    b_sorted = coords[perm][:, 0]
    assert (np.diff(b_sorted) >= 0).all()
    assert np.unique(T.morton_keys(coords)).size == coords.shape[0]
