"""The fused assignment kernel (csrc/fcos_assign.cu) ranks only the anchors of a (2R+1)^2 window around the grid node nearest
to a GT centre instead of all fw*fh anchors.  Properties checked here on the CPU:
  * R = 8 (17x17): for points inside, on the border of and arbitrarily far outside the map the `topk` = 9 nearest anchors of
    the full grid (ties broken by anchor index) all lie in the window -- exact for any input;
  * R = 4 (9x9, the build validated on B200 in round 1 and shipped): exact inside the map and up to 7 cells outside it, not
    beyond (a bus or trailer whose corner is still inside the map -- which the reference's range filter keeps, quirk 13 -- can
    have its centre up to ~9 cells outside at the 0.6 m pitch of the 180^2 config);
  * a lane-for-lane emulation of the kernel's selection loop with R = 8 agrees with the brute-force ranking, so switching the
    constant is safe; it waits for the first GPU run of round 2 (validation debts).

Why R = 8 is exact for k = 9: d^2 = dx(j)^2 + dy(i)^2, and |x_j - px| does not decrease as j moves away from the clamped
nearest column cj.  If a node (i, j) with |j - cj| >= 9 were among the 9 nearest, the 9 nodes (i, cj) ... (i, j -+ 1) of the
same row would all be at least as near, so (i, j) could at best tie for the 10th place.  The same holds for rows."""
import numpy as np


def _topk_full(px, py, fw, fh, stride, k):
    j, i = np.meshgrid(np.arange(fw), np.arange(fh))
    ax, ay = (j * stride).astype(np.float32).ravel(), (i * stride).astype(np.float32).ravel()
    d2 = (ax - np.float32(px)) ** 2 + (ay - np.float32(py)) ** 2
    order = np.lexsort((np.arange(d2.size), d2))
    return set(order[:k].tolist()), d2


def _topk_window(px, py, fw, fh, stride, k, radius=8):
    cj = int(np.clip(np.rint(np.float32(px) / np.float32(stride)), 0, fw - 1))
    ci = int(np.clip(np.rint(np.float32(py) / np.float32(stride)), 0, fh - 1))
    cand = []
    for di in range(-radius, radius + 1):
        for dj in range(-radius, radius + 1):
            i, j = ci + di, cj + dj
            if 0 <= i < fh and 0 <= j < fw:
                d2 = (np.float32(j * stride) - np.float32(px)) ** 2 + (np.float32(i * stride) - np.float32(py)) ** 2
                cand.append((np.float32(d2), i * fw + j))
    cand.sort()
    return set(a for _, a in cand[:k])


def test_window_contains_the_nine_nearest_anchors():
    rng = np.random.default_rng(0)
    fw = fh = 128
    stride = 8.0
    pts = np.concatenate([
        rng.uniform(0, (fw - 1) * stride, (300, 2)),                       # inside the map
        rng.uniform(-12 * stride, (fw - 1 + 12) * stride, (400, 2)),       # up to 12 cells outside, corners included
        np.stack([rng.uniform(-3000, 4000, 150), rng.uniform(-3000, 4000, 150)], 1),   # far outside
        np.array([[512.0, 512.0], [0.0, 0.0], [1016.0, 1016.0], [-96.0, -96.0], [1112.0, -96.0], [1016.0, 1112.0],
                  [3.9, 4.1], [-96.0, 500.0], [-1e4, 500.0], [500.0, 2e4], [-5e3, -5e3], [1212.5, 2606.8],
                  [-4.773005, -84.656433]]),
    ])
    for px, py in pts:
        full, d2 = _topk_full(px, py, fw, fh, stride, 9)
        win = _topk_window(px, py, fw, fh, stride, 9)
        if full != win:
            # only an exact distance tie at the 9th place may differ (then both choices are equally near)
            kth = np.sort(d2)[8]
            assert all(d2[a] == kth for a in full ^ win), (px, py)


def test_window_on_small_and_rectangular_grids():
    rng = np.random.default_rng(1)
    for fw, fh in ((9, 9), (16, 40), (180, 180)):
        for px, py in np.stack([rng.uniform(-400, (fw - 1) * 8 + 400, 150), rng.uniform(-400, (fh - 1) * 8 + 400, 150)], 1):
            full, d2 = _topk_full(px, py, fw, fh, 8.0, 9)
            win = _topk_window(px, py, fw, fh, 8.0, 9)
            if full != win:
                kth = np.sort(d2)[8]
                assert all(d2[a] == kth for a in full ^ win), (fw, fh, px, py)


def test_a_9x9_window_would_not_be_enough():
    """10.6 cells off the map in y and 0.6 in x: the ninth nearest anchor is 5 columns from the corner."""
    full, _ = _topk_full(-4.773005, -84.656433, 128, 128, 8.0, 9)
    assert full != _topk_window(-4.773005, -84.656433, 128, 128, 8.0, 9, radius=4)
    assert full == _topk_window(-4.773005, -84.656433, 128, 128, 8.0, 9, radius=8)


def _kernel_selection(px, py, fw, fh, stride, topk, radius=8):
    """Lane-for-lane emulation of the selection loop of k_fcos_assign (csrc/fcos_assign.cu): 32 lanes, candidate c = lane + 32 s,
    per-lane best slot, butterfly arg-min on (distance, anchor), the owner retires its candidate."""
    win = 2 * radius + 1
    slots = (win * win + 31) // 32
    inf, big = np.float32(np.inf), 0x7FFFFFFF
    px, py, stride = np.float32(px), np.float32(py), np.float32(stride)
    cj = min(max(int(np.rint(px / stride)), 0), fw - 1)
    ci = min(max(int(np.rint(py / stride)), 0), fh - 1)
    cd = np.full((32, slots), inf, np.float32)
    ca = np.full((32, slots), big, np.int64)
    for lane in range(32):
        for s in range(slots):
            c = lane + 32 * s
            if c < win * win:
                i, j = ci + c // win - radius, cj + c % win - radius
                if 0 <= i < fh and 0 <= j < fw:
                    dx, dy = np.float32(np.float32(j * stride) - px), np.float32(np.float32(i * stride) - py)
                    cd[lane, s] = np.float32(np.float32(dx * dx) + np.float32(dy * dy))
                    ca[lane, s] = i * fw + j
    chosen = set()
    lanes = np.arange(32)
    for _ in range(topk):
        bs = np.zeros(32, int)
        for lane in range(32):
            for s in range(1, slots):
                b = bs[lane]
                if cd[lane, s] < cd[lane, b] or (cd[lane, s] == cd[lane, b] and ca[lane, s] < ca[lane, b]):
                    bs[lane] = s
        d, a = cd[lanes, bs].copy(), ca[lanes, bs].copy()
        for o in (16, 8, 4, 2, 1):
            od, oa = d[lanes ^ o], a[lanes ^ o]
            take = (od < d) | ((od == d) & (oa < a))
            d, a = np.where(take, od, d), np.where(take, oa, a)
        assert len(set(a.tolist())) == 1
        if a[0] == big:
            break
        for lane in range(32):
            if ca[lane, bs[lane]] == a[lane]:
                cd[lane, bs[lane]], ca[lane, bs[lane]] = inf, big
                chosen.add(int(a[lane]))
    return chosen


def test_lane_level_emulation_of_the_kernel_selection():
    rng = np.random.default_rng(3)
    pts = np.concatenate([rng.uniform(0, 1016, (12, 2)), rng.uniform(-200, 1216, (12, 2)),
                          np.stack([rng.uniform(-3000, 4000, 8), rng.uniform(-3000, 4000, 8)], 1),
                          np.array([[512.0, 512.0], [-4.773005, -84.656433], [1212.5, 2606.8], [0.0, 0.0], [1016.0, 1016.0]])])
    for px, py in pts:
        full, d2 = _topk_full(px, py, 128, 128, 8.0, 9)
        got = _kernel_selection(px, py, 128, 128, 8.0, 9)
        if got != full:
            kth = np.sort(d2)[8]
            assert all(d2[a] == kth for a in got ^ full), (px, py)


def test_shipped_9x9_window_is_exact_up_to_seven_cells_outside():
    rng = np.random.default_rng(5)
    fw = fh = 128
    stride = 8.0
    lo, hi = -7 * stride, (fw - 1 + 7) * stride
    pts = np.concatenate([rng.uniform(lo, hi, (500, 2)),
                          np.array([[lo, lo], [hi, hi], [lo, hi], [hi, lo], [lo, 500.3], [500.3, hi], [lo + 3.9, lo + 4.1]])])
    for px, py in pts:
        full, d2 = _topk_full(px, py, fw, fh, stride, 9)
        win = _topk_window(px, py, fw, fh, stride, 9, radius=4)
        if full != win:
            kth = np.sort(d2)[8]
            assert all(d2[a] == kth for a in full ^ win), (px, py)
