"""oracle/nms.py -- TEST INFRASTRUCTURE (checker only).

Rotated-BEV IoU / greedy NMS (oracle/nms_ref.c: the published OpenPCDet `iou3d_nms` algorithm the reference's
binary-only iou3d_nms_cuda extension implements) plus an independent float64 convex-polygon clipping used to pin that
restatement, and the reference's proposal decoding (generate_proposals/centerpoint_gen_proposals.py:41-106,
iou_aware_gen_proposals.py:42-247) restated per sample for small cases.
"""
import ctypes

import numpy as np

from . import build_oracle

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_oracle.build())
        _lib.ud_ref_iou_bev.restype = ctypes.c_float
        _lib.ud_ref_box_overlap.restype = ctypes.c_float
        _lib.ud_ref_boxes_iou_bev.restype = None
        _lib.ud_ref_nms_bev.restype = ctypes.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def boxes_iou_bev(a, b):
    """(na,7), (nb,7) float32 [x,y,z,dx,dy,dz,heading] -> (na,nb) float32 (== boxes_iou_bev_gpu)"""
    a = np.ascontiguousarray(a[:, :7], np.float32)
    b = np.ascontiguousarray(b[:, :7], np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    if out.size:
        lib().ud_ref_boxes_iou_bev(_p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


def nms_bev(boxes_sorted, thresh):
    """boxes (n,7) already sorted by descending score -> kept indices (== nms_gpu's keep[:num_out])"""
    b = np.ascontiguousarray(boxes_sorted[:, :7], np.float32)
    keep = np.zeros((max(1, b.shape[0]),), np.int64)
    n = lib().ud_ref_nms_bev(_p(b), b.shape[0], ctypes.c_float(thresh), _p(keep)) if b.shape[0] else 0
    return keep[:n]


def nms_gpu_like(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
    """centerpoint_gen_proposals.py:85-106 (_nms_gpu_3d): stable descending sort, NMS, truncation."""
    order = np.argsort(-scores, kind="stable")
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep = nms_bev(boxes[order], thresh)
    sel = order[keep]
    return sel[:post_max_size] if post_max_size is not None else sel


# ---------------------------------------------------------------------------------------------------
# independent check: exact (float64) intersection area of two rotated rectangles by Sutherland-Hodgman clipping
def _corners64(box):
    x, y, dx, dy, a = float(box[0]), float(box[1]), float(box[3]), float(box[4]), float(box[6])
    c, s = np.cos(a), np.sin(a)
    loc = np.array([[-dx / 2, -dy / 2], [dx / 2, -dy / 2], [dx / 2, dy / 2], [-dx / 2, dy / 2]])
    rot = np.array([[c, -s], [s, c]])
    return loc @ rot.T + np.array([x, y])


def _clip(poly, p0, p1):
    out = []
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        da = (p1[0] - p0[0]) * (a[1] - p0[1]) - (p1[1] - p0[1]) * (a[0] - p0[0])
        db = (p1[0] - p0[0]) * (b[1] - p0[1]) - (p1[1] - p0[1]) * (b[0] - p0[0])
        if da >= 0:
            out.append(a)
        if (da >= 0) != (db >= 0):
            t = da / (da - db)
            out.append(a + t * (b - a))
    return out


def exact_iou_bev(a, b):
    pa, pb = _corners64(a), _corners64(b)
    poly = [p for p in pa]
    for i in range(4):
        if not poly:
            break
        poly = _clip(poly, pb[i], pb[(i + 1) % 4])
    inter = 0.0
    if len(poly) >= 3:
        q = np.array(poly)
        inter = 0.5 * abs(np.sum(q[:, 0] * np.roll(q[:, 1], -1) - np.roll(q[:, 0], -1) * q[:, 1]))
    sa, sb = float(a[3]) * float(a[4]), float(b[3]) * float(b[4])
    return inter / max(sa + sb - inter, 1e-8)
