"""oracle/refshim -- TEST INFRASTRUCTURE: runs routines of the reference's prebuilt binaries in this container to
produce golden vectors (see iou3d_shim.cpp).  Builds into oracle/_ref/ (git-ignored)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF_SO = os.path.join(os.environ.get("UNIDISTILL_REFERENCE_ROOT", "/root/reference"),
                      "unidistill/layers/head/det3d/generate_proposals/iou3d_nms_cuda.cpython-36m-x86_64-linux-gnu.so")


def available():
    return os.path.exists(REF_SO)


def build():
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libud_refshim.so")
    src = os.path.join(HERE, "iou3d_shim.cpp")
    if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
        subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-o", lib, src, "-ldl"], check=True)
    return lib


def reference_boxes_iou_bev_cpu(boxes_a, boxes_b):
    """(na,7), (nb,7) float32 -> (na,nb) float32 computed by the reference binary's boxes_iou_bev_cpu."""
    import torch  # noqa: F401  (loads libc10 / libtorch*: the extension's DT_NEEDED sonames resolve against them)

    shim = ctypes.CDLL(build(), mode=ctypes.RTLD_GLOBAL)
    a = np.ascontiguousarray(boxes_a[:, :7], np.float32)
    b = np.ascontiguousarray(boxes_b[:, :7], np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    rc = shim.ud_refshim_boxes_iou_bev_cpu(REF_SO.encode(), a.ctypes.data_as(ctypes.c_void_p), a.shape[0],
                                           b.ctypes.data_as(ctypes.c_void_p), b.shape[0],
                                           out.ctypes.data_as(ctypes.c_void_p))
    if rc != 1:
        raise RuntimeError("reference boxes_iou_bev_cpu returned %d" % rc)
    return out
