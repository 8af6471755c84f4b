"""SURVEY 8f-4 -- sweep merge + range filter + ragged collate: the numpy oracle against the fixture produced by the
reference's own CollectLidarSweeps / ObjectRangeFilter (CPU), and the GPU kernel ud_collect_sweeps against both (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import input_side as OI

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "input_side.npz")
PC_RANGE = [-51.2, -51.2, -5.0, 51.2, 51.2, 3.0]


def _samples(g):
    out = []
    b = 0
    while "in_points_%d" % b in g.files:
        n = int(g["n_sweeps_%d" % b])
        out.append(dict(points=g["in_points_%d" % b], sweep_points=[g["sweep_%d_%d" % (b, k)] for k in range(n)],
                        key_e2g=g["key_e2g_%d" % b], key_l2e=g["key_l2e_%d" % b], ts=float(g["ts_%d" % b]),
                        s_e2g=[g["s_e2g_%d_%d" % (b, k)] for k in range(n)], s_ts=[float(g["s_ts_%d_%d" % (b, k)]) for k in range(n)]))
        b += 1
    return out


def test_oracle_matches_reference_fixture():
    g = np.load(GOLD)
    for b, s in enumerate(_samples(g)):
        merged = OI.collect_lidar_sweeps(s["points"], s["sweep_points"], s["key_e2g"], s["key_l2e"], s["s_e2g"], s["ts"],
                                         s["s_ts"])
        assert np.array_equal(merged.astype(np.float32), g["merged_%d" % b])
        kept = merged[OI.mask_points_by_range(merged, PC_RANGE)]
        assert np.array_equal(kept.astype(np.float32), g["filtered_%d" % b])
        assert np.all(merged[: s["points"].shape[0], 4] == 0)            # key frame: time column zeroed


def _packed(samples):
    from unidistill_b200 import input_side as IS

    return IS.pack_sweeps([dict(points=s["points"], sweep_points=s["sweep_points"],
                                sweep_mats=[IS.sweep_to_key_matrix(s["key_l2e"], s["key_e2g"], m) for m in s["s_e2g"]],
                                sweep_time_lags=[(s["ts"] - t) / 1e6 for t in s["s_ts"]]) for s in samples], pin=False)


def test_pack_sweeps_layout():
    g = np.load(GOLD)
    samples = _samples(g)
    p = _packed(samples)
    n_seg = sum(1 + len(s["sweep_points"]) for s in samples)
    assert p["seg_offsets"].shape[0] == n_seg + 1 and p["mats"].shape == (n_seg, 3, 4) and p["mats"].dtype == torch.float64
    assert p["sample_first_seg"].tolist() == [0, 4, 5, 8] and int(p["seg_offsets"][-1]) == p["points"].shape[0]
    assert float(p["time_lag"][0]) == 0.0 and float(p["time_lag"][1]) == pytest.approx(0.05)


@pytest.mark.gpu
def test_kernel_matches_reference_fixture(cuda, native_lib):
    from unidistill_b200 import input_side as IS

    g = np.load(GOLD)
    samples = _samples(g)
    packed = {k: v.to(cuda) for k, v in _packed(samples).items()}
    flat, offsets = IS.collect_sweeps(packed, PC_RANGE)
    clouds = IS.split_ragged(flat, offsets)
    assert len(clouds) == len(samples)
    for b, c in enumerate(clouds):
        ref = g["filtered_%d" % b]
        got = c.cpu().numpy()
        assert got.shape == ref.shape, (b, got.shape, ref.shape)
        # xyz: float64 transform rounded once to float32, like numpy; a different summation order inside BLAS can move a
        # coordinate by one float32 ulp (never across the inclusive range bounds here: the row counts above are equal)
        assert np.array_equal(got[:, 3:], ref[:, 3:])
        ulp = np.abs(np.spacing(ref[:, :3]))
        assert np.all(np.abs(got[:, :3] - ref[:, :3]) <= ulp)
        assert (got[:, :3] != ref[:, :3]).mean() < 1e-3
    # ragged output feeds the voxelizer directly and equals voxelizing the reference's clouds
    from unidistill_b200 import synthetic as S
    from unidistill_b200.voxelize import hard_voxelize

    a = hard_voxelize(clouds, S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000)
    b_ = hard_voxelize([torch.from_numpy(g["filtered_%d" % i]).to(cuda) for i in range(len(samples))], S.VOXEL_SIZE_128,
                       S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000)
    assert torch.equal(a[1], b_[1]) and torch.equal(a[2], b_[2])


@pytest.mark.gpu
def test_kernel_edge_cases(cuda, native_lib):
    from unidistill_b200 import input_side as IS

    # a sample without sweeps, a sample whose points all fall outside, points exactly on the inclusive bounds
    key = np.array([[51.2, 0, 0, 1, 9], [-51.2, -51.2, 0, 2, 9], [51.200005, 0, 0, 3, 9], [0, 60, 0, 4, 9]], np.float32)
    far = np.full((5, 5), 500.0, np.float32)
    packed = {k: v.to(cuda) for k, v in IS.pack_sweeps([dict(points=key), dict(points=far)], pin=False).items()}
    flat, offsets = IS.collect_sweeps(packed, PC_RANGE)
    assert offsets.tolist() == [0, 2, 2]
    got = flat[:2].cpu().numpy()
    assert np.array_equal(got[:, 3], [1, 2]) and np.all(got[:, 4] == 0)
