"""the 42 towers of a CenterHead run as one convolution + one BatchNorm + one ReLU + one grouped convolution
(`batched_towers`, reference layers/head/det3d/center_head.py:311-375 for what a tower is).  Same outputs, same gradients
on every original parameter, same running statistics as `[task(x) for task in tasks]`."""
import copy
from collections import OrderedDict

import pytest
import torch


def _tasks(cin=16, mid=16):
    from unidistill.layers.head.det3d.center_head import SepHead

    heads = OrderedDict([("iou", (1, 2)), ("reg", (2, 2)), ("height", (1, 2)), ("dim", (3, 2)), ("rot", (2, 2)),
                         ("vel", (2, 2))])
    tasks = torch.nn.ModuleList()
    for nc in (1, 2, 2, 1, 2, 2):
        h = OrderedDict(heads)
        h["hm"] = (nc, 2)
        tasks.append(SepHead(cin, h, head_conv=mid, bn=True, final_kernel=3))
    return tasks


def _compare(dev, cin=16, hw=12, tol=2e-3, out_tol=(1e-3, 1e-4)):
    from unidistill.layers.head.det3d.center_head import batched_towers

    torch.manual_seed(0)
    tasks = _tasks(cin, cin).to(dev)
    tasks2 = copy.deepcopy(tasks)
    x = torch.randn(2, cin, hw, hw, device=dev, requires_grad=True)
    x2 = x.detach().clone().requires_grad_(True)
    for mode in (True, False):
        tasks.train(mode)
        tasks2.train(mode)
        a = [t(x) for t in tasks]
        b = batched_towers(tasks2, x2)
        assert b is not None
        for da, db in zip(a, b):
            assert list(da.keys()) == list(db.keys())
            for k in da:
                assert da[k].shape == db[k].shape
                assert torch.allclose(da[k], db[k], rtol=out_tol[0], atol=out_tol[1]), (mode, k, float((da[k] - db[k]).abs().max()))
        la = sum(v.square().sum() for d in a for v in d.values())
        lb = sum(v.square().sum() for d in b for v in d.values())
        la.backward()
        lb.backward()
        assert float((x.grad - x2.grad).abs().max()) <= tol * float(x.grad.abs().max())
        for (n1, p1), (_, p2) in zip(tasks.named_parameters(), tasks2.named_parameters()):
            if mode and n1.endswith(".0.bias"):
                continue          # a bias in front of a train-mode BatchNorm has zero true gradient: both are round-off
            scale = max(float(p1.grad.abs().max()), 1e-3 * float(la.detach()) ** 0.5)
            assert float((p1.grad - p2.grad).abs().max()) <= tol * scale, (mode, n1)
        for (n1, b1), (_, b2) in zip(tasks.named_buffers(), tasks2.named_buffers()):
            assert torch.allclose(b1.float(), b2.float(), rtol=1e-4, atol=1e-5), n1
        x.grad = None
        x2.grad = None
        tasks.zero_grad()
        tasks2.zero_grad()


def test_batched_towers_equal_per_tower_cpu():
    _compare(torch.device("cpu"))


def test_irregular_towers_fall_back():
    from unidistill.layers.head.det3d.center_head import batched_towers

    tasks = _tasks()
    tasks[2].rot = torch.nn.Sequential(torch.nn.Conv2d(16, 2, 3, padding=1))
    assert batched_towers(tasks, torch.randn(1, 16, 8, 8)) is None


@pytest.mark.gpu
def test_batched_towers_equal_per_tower_gpu(cuda):
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        # cuDNN picks different fp32 algorithms for a 64->c and a 2688->76 convolution (Winograd variants carry ~1e-3)
        _compare(cuda, cin=64, hw=32, tol=1e-2, out_tol=(5e-3, 5e-3))
    finally:
        torch.backends.cudnn.allow_tf32 = old


@pytest.mark.gpu
def test_head_forward_uses_batched_towers_and_matches(cuda, native_lib):
    """CenterHead.forward on CUDA goes through the batched towers; the reference-golden GPU test (test_gpu_glue) pins its
    numbers, here: identical to the per-tower path on the same module, NHWC input included."""
    from test_golden_mirror import _head, _load
    from unidistill.layers.head.det3d import center_head as CH

    g = _load("head_loss.npz")
    head = _head(g).to(cuda).eval()
    head.distill = True
    x = torch.from_numpy(g["x"]).to(cuda).contiguous(memory_format=torch.channels_last)
    gt = torch.from_numpy(g["gt_boxes"]).to(cuda)
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        with torch.no_grad():
            a = head(x, gt)["multi_head_features"]
            CH.BATCH_TOWERS = False
            b = head(x, gt)["multi_head_features"]
    finally:
        CH.BATCH_TOWERS = True
        torch.backends.cudnn.allow_tf32 = old
    for da, db in zip(a, b):
        for k in db:
            # both sit within 1e-4 of the reference fixture (tests/test_gpu_glue.py): cuDNN picks different fp32 algorithms
            assert torch.allclose(da[k], db[k], rtol=5e-4, atol=5e-4), (k, float((da[k] - db[k]).abs().max()))
