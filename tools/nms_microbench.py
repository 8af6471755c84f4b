"""Development aid: batched rotated NMS (csrc/nms.cu) timings at the reference's eval sizes
(6 tasks x bs=4 sets of 1500 candidates, IoU threshold 0.1, keep 100) and the whole proposal layer."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))

from unidistill_b200 import models, nms  # noqa: E402


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    S, n = 24, 1500
    b = np.zeros((S, n, 7), np.float32)
    b[..., 0:2] = rng.uniform(-50, 50, (S, n, 2))
    b[..., 3] = rng.uniform(1, 6, (S, n))
    b[..., 4] = rng.uniform(0.5, 3, (S, n))
    b[..., 5] = 1.5
    b[..., 6] = rng.uniform(-3.2, 3.2, (S, n))
    bt = torch.from_numpy(b).to(dev)
    for max_keep in (100, 1500):
        ms = timed(lambda: nms.nms_bev_batched(bt, None, 0.1, max_keep))
        keep, num = nms.nms_bev_batched(bt, None, 0.1, max_keep)
        print("nms_bev_batched  %d sets x %d boxes, thresh 0.1, max_keep %4d: %.3f ms  (kept/set mean %.1f)" % (
            S, n, max_keep, ms, float(num.float().mean())))
    ms = timed(lambda: nms.boxes_iou_bev(bt[0], bt[1]))
    print("boxes_iou_bev 1500 x 1500: %.3f ms (%.1f M pairs/s)" % (ms, 2.25 / ms * 1e3))
    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    head = models.DetHead(cfg.det_head).to(dev).eval()
    x = torch.randn(4, 512, 128, 128, device=dev)
    with torch.no_grad():
        ret = {"multi_head_features": [task(head.dense_head.shared_conv(x)) for task in head.dense_head.tasks]}
        ms = timed(lambda: head.dense_head.proposal_layer.generate_predicted_boxes(dict(ret), {}), iters=5)
    print("IouAwareGenProposals.generate_predicted_boxes bs=4, 6 tasks, K=1500, 128x128 maps: %.3f ms" % ms)


if __name__ == "__main__":
    main()
