"""Drop-in module for the reference's prebuilt extension `iou3d_nms_cuda` (same module path, same function
signatures as the pybind exports `nm -DC` shows: nms_gpu, boxes_iou_bev_gpu), backed by csrc/nms.cu.

    nms_gpu(boxes (N,7) cuda f32 sorted by descending score, keep int64 CPU tensor (N,), thresh) -> num_out
        fills keep[:num_out] with the kept row indices (ascending); call site
        generate_proposals/centerpoint_gen_proposals.py:97-100.
    boxes_iou_bev_gpu(boxes_a (N,7), boxes_b (M,7), ans_iou (N,M) cuda f32) -> 1

The `keep` tensor living on the host is the reference's contract (it forces a device->host copy); the mirrored
proposal layer of this repo calls unidistill_b200.nms.nms_bev_batched instead and never leaves the device.
"""
import torch

from unidistill_b200 import nms as _nms


def nms_gpu(boxes, keep, thresh):
    if not boxes.is_cuda:
        raise RuntimeError("boxes must be a CUDA tensor ")
    n = int(boxes.shape[0])
    if n == 0:
        return 0
    k, num = _nms.nms_bev_batched(boxes.reshape(1, n, -1), None, float(thresh), n)
    num_out = int(num.item())
    keep[:num_out] = k[0, :num_out].to(device=keep.device, dtype=keep.dtype)
    return num_out


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    ans_iou.copy_(_nms.boxes_iou_bev(boxes_a, boxes_b))
    return 1
