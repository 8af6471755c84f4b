"""Fused FCOS target-assignment kernel (csrc/fcos_assign.cu) == the torch tensor program of the mirrored FCOSAssigner
(which tests/test_golden_mirror.py pins against the REAL reference) == the committed reference fixture, bit for bit."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "assigner.npz")
TASKS = [dict(num_class=1, class_names=["car"]), dict(num_class=2, class_names=["truck", "construction_vehicle"]),
         dict(num_class=2, class_names=["bus", "trailer"]), dict(num_class=1, class_names=["barrier"]),
         dict(num_class=2, class_names=["motorcycle", "bicycle"]),
         dict(num_class=2, class_names=["pedestrian", "traffic_cone"])]
CLASS_NAMES = [n for t in TASKS for n in t["class_names"]]


def _assigner(grid, voxel, max_objs):
    from unidistill.layers.head.det3d import FCOSAssigner

    return FCOSAssigner(out_size_factor=8, tasks=TASKS, dense_reg=1, gaussian_overlap=0.1, max_objs=max_objs, min_radius=2,
                        mapping={n: i + 1 for i, n in enumerate(CLASS_NAMES)}, grid_size=grid, pc_range=[-51.2, -51.2],
                        voxel_size=voxel, assign_topk=9, no_log=False, with_velocity=True)


def _both(assigner, gt):
    from unidistill.layers.head.det3d.target_assigner import fcos_assigner as F

    old = F.USE_ASSIGN_KERNEL
    try:
        F.USE_ASSIGN_KERNEL = True
        a = assigner.assign_targets(gt)
        F.USE_ASSIGN_KERNEL = False
        b = assigner.assign_targets(gt)
    finally:
        F.USE_ASSIGN_KERNEL = old
    return a, b


def _assert_same(a, b):
    for key in ("heatmap", "ind", "mask", "cat", "box_encoding"):
        for t in range(len(TASKS)):
            x, y = a[key][t], b[key][t]
            assert x.shape == y.shape and x.dtype == y.dtype, (key, t, x.shape, y.shape, x.dtype, y.dtype)
            if x.dtype.is_floating_point:
                same = (x == y) | (torch.isnan(x) & torch.isnan(y))
                assert bool(same.all()), "%s task %d: %d elements differ, max |d| %g" % (
                    key, t, int((~same).sum()), float((x - y)[~same].abs().max()))
            else:
                assert torch.equal(x, y), (key, t)


@pytest.mark.parametrize("seed,B,M,max_objs", [(0, 4, 40, 2500), (1, 2, 12, 500), (2, 3, 64, 60)])
def test_kernel_equals_tensor_program(cuda, native_lib, seed, B, M, max_objs):
    rng = np.random.default_rng(seed)
    gt = np.zeros((B, M, 10), np.float32)
    for b in range(B):
        n = int(rng.integers(1, M)) if b != B - 1 else 0          # last sample empty
        gt[b, :n, 0:2] = rng.uniform(-54, 54, (n, 2))               # a few centres outside the map (quirk 13)
        gt[b, :n, 2] = rng.normal(-1, 0.5, n)
        gt[b, :n, 3:6] = rng.uniform(0.3, 8, (n, 3))
        gt[b, :n, 6] = rng.uniform(-7, 7, n)
        gt[b, :n, 7:9] = rng.normal(0, 3, (n, 2))
        gt[b, :n, 9] = rng.integers(1, 11, n)
        if n > 4:
            gt[b, 2, :9] = 0                                        # an all-zero row inside the valid prefix
            gt[b, 3, 0:2] = gt[b, 1, 0:2]                           # two GTs at the same centre (nearest-GT tie)
            gt[b, 3, 9] = gt[b, 1, 9]
        gt[b, n:, 9] = 1                                            # padded rows carry label 0 + 1 (quirk 7)
    gt_t = torch.from_numpy(gt).to(cuda)
    a, b_ = _both(_assigner([1024, 1024, 40], [0.1, 0.1], max_objs), gt_t)
    _assert_same(a, b_)
    assert sum(int(a["mask"][t].sum()) for t in range(6)) > 0


@pytest.mark.parametrize("grid,voxel", [([1024, 1024, 40], [0.1, 0.1]), ([1440, 1440, 40], [0.075, 0.075])])
def test_centres_far_outside_the_map(cuda, native_lib, grid, voxel):
    """GT centres 8-12 (and up to 40) feature-map cells outside the map, along edges and off all four corners: the
    17x17 candidate window (FA_RADIUS = 8) must reproduce the full-grid top-9 of the tensor program (SURVEY quirk 13:
    the reference's range filter keeps boxes whose centre lies outside as long as a corner is inside)."""
    rng = np.random.default_rng(13)
    cell = voxel[0] * 8
    half = 51.2 if grid[0] == 1024 else 54.0
    B, M = 4, 48
    gt = np.zeros((B, M, 10), np.float32)
    for b in range(B):
        n = 40
        out = rng.uniform(8, 12 if b < 3 else 40, (n, 2)) * cell            # distance outside, per axis
        side = rng.integers(0, 3, (n, 2))                                   # 0: low side, 1: inside, 2: high side
        inside = rng.uniform(-half, half, (n, 2))
        xy = np.where(side == 0, -half - out, np.where(side == 2, half + out, inside))
        if grid[0] != 1024:
            xy = xy + (half - 51.2)                                         # _assigner's map starts at -51.2: centre +2.8
        gt[b, :n, 0:2] = xy
        gt[b, :n, 2] = rng.normal(-1, 0.5, n)
        gt[b, :n, 3:6] = rng.uniform(0.3, 12, (n, 3))
        gt[b, :n, 6] = rng.uniform(-3.2, 3.2, n)
        gt[b, :n, 9] = rng.integers(1, 11, n)
        gt[b, n:, 9] = 1
    a, b_ = _both(_assigner(grid, voxel, 2500), torch.from_numpy(gt).to(cuda))
    _assert_same(a, b_)
    for t in range(6):
        assert torch.equal(a["n_pos"][t].long(), a["mask"][t].sum(1)), "un-truncated count, task %d" % t


def test_overflow_is_reported_not_silent(cuda, native_lib):
    """More positives than max_objs: outputs are truncated like the tensor program, and the un-truncated count is
    returned (the reference asserts in padding_to_maxobjs)."""
    rng = np.random.default_rng(5)
    gt = np.zeros((1, 30, 10), np.float32)
    gt[0, :, 0:2] = rng.uniform(-45, 45, (30, 2))
    gt[0, :, 3:6] = 2.0
    gt[0, :, 9] = 1                                                          # all cars -> ~270 positives in task 0
    a, b_ = _both(_assigner([1024, 1024, 40], [0.1, 0.1], 60), torch.from_numpy(gt).to(cuda))
    _assert_same(a, b_)
    assert int(a["n_pos"][0][0]) > 60 and int(a["mask"][0].sum()) == 60
    assert int(a["n_pos"][0][0]) == int(b_["n_pos"][0][0]) == int(a["heatmap"][0].eq(1).sum())


def test_fewer_anchors_than_slots(cuda, native_lib):
    """A 16x16 map has 256 anchors but max_objs = 2500 slots: both paths zero-pad the tail."""
    gt = np.zeros((2, 6, 10), np.float32)
    gt[:, :3, 0:2] = np.array([[1.0, 2.0], [-3.0, 0.5], [4.0, -4.0]], np.float32)
    gt[:, :3, 3:6] = 1.5
    gt[:, :, 9] = 1
    gt[0, 1, 9] = 9
    from unidistill.layers.head.det3d import FCOSAssigner

    assigner = FCOSAssigner(out_size_factor=8, tasks=TASKS, dense_reg=1, gaussian_overlap=0.1, max_objs=2500, min_radius=2,
                            mapping={n: i + 1 for i, n in enumerate(CLASS_NAMES)}, grid_size=[128, 128, 40],
                            pc_range=[-6.4, -6.4], voxel_size=[0.1, 0.1], assign_topk=9, no_log=False, with_velocity=True)
    a, b_ = _both(assigner, torch.from_numpy(gt).to(cuda))
    _assert_same(a, b_)
    assert a["ind"][0].shape == (2, 2500)


def test_kernel_matches_reference_golden(cuda, native_lib):
    """The fixture was produced by the reference's own FCOSAssigner (tests/golden/make_golden.py)."""
    g = np.load(GOLDEN)
    gt = torch.from_numpy(g["gt_boxes"]).to(cuda)
    assigner = _assigner([256, 256, 40], [0.4, 0.4], 500)
    a, b = _both(assigner, gt)
    _assert_same(a, b)
    for t in range(6):
        for key, name in (("heatmap", "heatmap"), ("ind", "ind"), ("mask", "mask"), ("cat", "cat")):
            k = "%s_%d" % (name, t)
            if k in g.files:
                assert np.array_equal(a[key][t].cpu().numpy().astype(g[k].dtype), g[k]), k


def test_kernel_inside_cuda_graph(cuda, native_lib):
    """The assignment rides in the dense trunk's CUDA graph: capture + replay give the eager result."""
    rng = np.random.default_rng(7)
    gt = np.zeros((2, 20, 10), np.float32)
    gt[:, :9, 0:2] = rng.uniform(-50, 50, (2, 9, 2))
    gt[:, :9, 3:6] = rng.uniform(0.5, 5, (2, 9, 3))
    gt[:, :9, 6] = rng.uniform(-3, 3, (2, 9))
    gt[:, :, 9] = 1
    gt[:, :9, 9] = rng.integers(1, 11, (2, 9))
    gt_t = torch.from_numpy(gt).to(cuda)
    assigner = _assigner([1024, 1024, 40], [0.1, 0.1], 2500)
    eager = assigner.assign_targets(gt_t)
    static = gt_t.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        assigner.assign_targets(static)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = assigner.assign_targets(static)
    static.copy_(gt_t)
    graph.replay()
    _assert_same(out, eager)
