"""Event timeline of one CTA of k_spconv_umma_dg (library built with -DUD_DG_PROBE -DUD_DG_TRACE=<block>, through
UD_B200_LIB): runs the LiDAR encoder forward of the bench batch and, after the first call with the requested (cin, cout),
dumps per stage the clock64() of: set indices landed / rows converted / slot free / arrived, MMA warp A full / issued."""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
from unidistill_b200 import models, native, synthetic as S  # noqa: E402

want = tuple(int(v) for v in os.environ.get("UD_TRACE_SHAPE", "64,64").split(","))
dev = torch.device("cuda:0")
lib = native.load()
trace = lib._cdll.ud_debug_dg_trace
trace.argtypes = [ctypes.c_void_p]
trace.restype = ctypes.c_int
NST = 128
buf = (ctypes.c_longlong * (8 * NST))()
got = []


def wrap(orig, first):
    def wrapped(*a):
        rc = orig(*a)
        if (int(a[first + 2]), int(a[first + 3])) == want and int(a[first + 1]) == 27 and not got:
            torch.cuda.synchronize()
            trace(ctypes.cast(buf, ctypes.c_void_p))
            got.append([list(buf[e * NST:(e + 1) * NST]) for e in range(8)])
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
    enc(pts)
    got.clear()
    enc(pts)
ev = got[0]
t0 = min(v for v in ev[4] if v > 0)
print("stage | set: idx  conv  slot  arrive | mma: b_full a_full issued | a_full-arrive  issued-a_full  slot-issued(it-4)")
for it in range(NST):
    if ev[4][it] == 0:
        break
    r = lambda e: ev[e][it] - t0 if ev[e][it] else -1
    prev_issued = ev[5][it - 4] - t0 if it >= 4 else -1
    print("%5d | %6d %6d %6d %6d | %6d %6d %6d | %6d %6d %6d" % (
        it, r(0), r(1), r(2), r(3), r(6), r(4), r(5), r(4) - r(3), r(5) - r(4), (r(2) - prev_issued) if it >= 4 else -1))
