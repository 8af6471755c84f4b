"""Role-cycle accounting of k_spconv_umma_dg (needs a library built with -DUD_DG_PROBE, passed through UD_B200_LIB).

Runs the LiDAR encoder forward of the bench batch (bs = 4) and prints, per tensor-core convolution call, the cycles per
pipeline stage each role spent waiting / working.  Diagnostic only."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
from unidistill_b200 import models, native, synthetic as S  # noqa: E402

dev = torch.device("cuda:0")
lib = native.load()
probe = lib._cdll.ud_debug_dg_probe
probe.argtypes = [ctypes.c_void_p, ctypes.c_int]
probe.restype = ctypes.c_int
buf = (ctypes.c_ulonglong * 16)()
rows = []


def wrap(orig, first):
    # `first` = position of n_out in the argument list (followed by K, cin, cout)
    def wrapped(*a):
        torch.cuda.synchronize()
        probe(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = orig(*a)
        e1.record()
        torch.cuda.synchronize()
        probe(ctypes.cast(buf, ctypes.c_void_p), 0)
        rows.append((int(a[first]), int(a[first + 1]), int(a[first + 2]), int(a[first + 3]), e0.elapsed_time(e1), list(buf)))
        return rc
    return wrapped


lib.ud_sparse_conv_tc_forward = wrap(lib.ud_sparse_conv_tc_forward, 6)
lib.ud_sparse_conv_tc_forward_affine = wrap(lib.ud_sparse_conv_tc_forward_affine, 9)
torch.manual_seed(0)
cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
enc = models.LidarEncoder(cfg.lidar_encoder).to(dev).eval()
lid = S.lidar_batch(4, seed=0)
pts = [p.to(dev) for p in lid["points"]]
with torch.no_grad():
    for it in range(2):
        rows.clear()
        enc(pts)
print("%8s %3s %4s %4s %7s | gather set/stage: %6s %6s %6s %6s %6s %6s | mma/stage: %6s %6s %6s | cta: %5s %8s %8s %6s | %6s %6s %6s" % (
    "n_out", "K", "cin", "cout", "ms", "stages", "idx_w", "issue", "dat+cv", "slot_w", "st+arr", "wait_b", "wait_a", "issue",
    "ctas", "cyc/cta", "cyc/stg", "epilog", "fence", "done", "-"))
for n_out, K, cin, cout, ms, b in rows:
    gs = max(1, b[0])
    ms_ = max(1, b[8])
    ctas = max(1, b[12])
    print("%8d %3d %4d %4d %7.3f | %18d %6.0f %6.0f %6.0f %6.0f %6.0f | %17.0f %6.0f %6.0f | %10d %8.0f %8.0f %6.0f | %6.0f %6.0f %6.0f" % (
        n_out, K, cin, cout, ms, b[0], b[1] / gs, b[2] / gs, b[3] / gs, b[4] / gs, b[5] / gs, b[9] / ms_, b[10] / ms_,
        b[11] / ms_, b[12], b[13] / ctas, b[13] / ms_, b[14] / ctas, b[6] / ms_, b[15] / ms_, 0))


# -DUD_DG_TRACE builds: event timeline of one CTA of the LAST traced call that matches UD_TRACE_SHAPE (default 64,64)
if hasattr(lib._cdll, "ud_debug_dg_trace"):
    pass
