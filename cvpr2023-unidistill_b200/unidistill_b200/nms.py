"""Rotated-BEV IoU and greedy NMS on top of the C ABI (csrc/nms.cu): the device-side replacement of the reference's
binary-only `iou3d_nms_cuda` extension (generate_proposals/centerpoint_gen_proposals.py:85-106)."""
import torch

from . import native

_WS_NMS = 2  # workspace slot of the suppression mask


def boxes_iou_bev(boxes_a, boxes_b):
    """(N,7+), (M,7+) cuda f32 [x,y,z,dx,dy,dz,heading,...] -> (N,M) f32 rotated IoU in the BEV plane."""
    native.require_cuda(boxes_a, boxes_b)
    a = boxes_a[:, :7].contiguous().float()
    b = boxes_b[:, :7].contiguous().float()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    rc = native.load().ud_boxes_iou_bev(native.ptr(a), int(a.shape[0]), native.ptr(b), int(b.shape[0]),
                                        native.ptr(out), native.stream_ptr(a.device))
    native.check(rc, "ud_boxes_iou_bev")
    return out


def nms_bev_batched(boxes_sorted, n_valid, thresh, max_keep):
    """boxes_sorted (S,n,7+) cuda f32, every set sorted by descending score; n_valid (S,) int32 cuda or None.
    -> keep (S,max_keep) int32 (positions into the sorted set, ascending, -1 padded), num_keep (S,) int32.
    Nothing synchronises: the counts stay on the device."""
    native.require_cuda(boxes_sorted)
    lib = native.load()
    S, n = int(boxes_sorted.shape[0]), int(boxes_sorted.shape[1])
    dev = boxes_sorted.device
    b = boxes_sorted[:, :, :7].contiguous().float()
    keep = torch.empty((S, int(max_keep)), dtype=torch.int32, device=dev)
    num = torch.empty((S,), dtype=torch.int32, device=dev)
    nv = None
    if n_valid is not None:
        nv = n_valid.to(device=dev, dtype=torch.int32).contiguous()
    nbytes = native.query_bytes("ud_nms_bev_workspace_bytes", S, n)
    ws = native.WORKSPACE.get(nbytes, dev, slot=_WS_NMS)
    rc = lib.ud_nms_bev(native.ptr(b), native.ptr(nv), S, n, float(thresh), int(max_keep), native.ptr(keep),
                        native.ptr(num), native.ptr(ws), ws.numel(), native.stream_ptr(dev))
    native.check(rc, "ud_nms_bev")
    return keep, num
