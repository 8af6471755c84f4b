"""GPU micro-benchmark: the 42 CenterHead towers per tower vs batched, forward and forward+backward, at the
bench shape (4, 64, 128, 128) NHWC."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
from test_head_batching import _tasks  # noqa: E402
from unidistill.layers.head.det3d.center_head import batched_towers  # noqa: E402

torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
tasks = _tasks(64, 64).to(dev).to(memory_format=torch.channels_last).train()
x = torch.randn(4, 64, 128, 128, device=dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)


def run(fn, bwd):
    out = fn()
    if bwd:
        loss = sum(v.sum() for d in out for v in d.values())
        loss.backward()


for name, fn in (("per-tower", lambda: [t(x) for t in tasks]), ("batched", lambda: batched_towers(tasks, x))):
    for bwd in (False, True):
        for _ in range(5):
            run(fn, bwd)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(fn, bwd)
        e1.record()
        torch.cuda.synchronize()
        print("%-10s %-8s %.3f ms" % (name, "fwd+bwd" if bwd else "fwd", e0.elapsed_time(e1) / 20), flush=True)
