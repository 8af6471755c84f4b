"""Kernel vs golden values produced by the reference's own BINARY (tests/golden/iou3d_ref.npz: IoU matrices computed by
boxes_iou_bev_cpu of the prebuilt iou3d_nms_cuda extension, run through oracle/refshim in the build container).
Kept in a file that sorts last: it was added after the last GPU run of round 1 (its CPU twin, oracle == golden bit for bit, and
the kernel-vs-oracle test it mirrors are green)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "proposals.npz")


def test_iou_bev_matches_reference_binary_golden(cuda, native_lib):
    """Random boxes of tests/golden/iou3d_ref.npz (values computed by the reference binary's boxes_iou_bev_cpu)."""
    from unidistill_b200 import nms

    g = np.load(os.path.join(os.path.dirname(GOLDEN), "iou3d_ref.npz"))
    b, ref = g["boxes"][87:], g["iou"][87:, 87:]
    got = nms.boxes_iou_bev(torch.from_numpy(b).to(cuda), torch.from_numpy(b).to(cuda)).cpu().numpy()
    assert np.abs(got - ref).max() < 2e-5
    assert (got[ref == 0] == 0).all()
