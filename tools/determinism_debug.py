"""Development aid: run the LiDAR encoder several times on the same input and report the first operator whose
output differs between runs (run-to-run determinism of the sparse path)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))

from unidistill_b200 import models, sparse, synthetic as S  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
    model = models.BEVFusionCenterHead(cfg).to(dev).train()
    enc = model.lidar_encoder
    pps, sweeps, bs = int(os.environ.get("UD_DD_PPS", "8000")), int(os.environ.get("UD_DD_SWEEPS", "2")), 2
    batch = S.lidar_batch(bs, seed=9, points_per_sweep=pps, sweeps=sweeps)
    pts = [batch["points"][b].to(dev) for b in range(bs)]

    rec = []
    orig_conv = sparse._conv_forward_raw
    orig_bn = sparse.BatchNormActFunction.apply

    def conv(feats, w_k, bias, table, n_out, K, cin, cout, rb=None, weight=None, want_stats=False):
        out = orig_conv(feats, w_k, bias, table, n_out, K, cin, cout, rb, weight, want_stats)
        rec.append(("conv %d->%d K=%d n=%d in" % (cin, cout, K, n_out), feats.detach().clone()))
        rec.append(("conv %d->%d K=%d n=%d out" % (cin, cout, K, n_out), out.detach().clone()))
        s = getattr(out, "_ud_bn_sums", None)
        if s is not None:
            rec.append(("conv %d->%d fused sums" % (cin, cout), s.detach().clone()))
        return out

    sparse._conv_forward_raw = conv
    runs = []
    for r in range(int(os.environ.get("UD_DD_RUNS", "4"))):
        rec.clear()
        with torch.no_grad():
            voxels, coords, num = enc.voxelizer(pts)
            feats = enc.vfe(voxels, num)
            rec.append(("voxel coords", coords.clone()))
            rec.append(("vfe", feats.clone()))
            sp, stride, _ = enc.backbone_3d(feats, coords, bs)
            bev, _ = enc.map_to_bev(sp, stride)
            rec.append(("bev", bev.clone()))
        torch.cuda.synchronize()
        runs.append(list(rec))
    for r in range(1, len(runs)):
        print("run %d vs run 0" % r)
        for (name, a), (_, b) in zip(runs[r], runs[0]):
            if a.shape != b.shape:
                print("   %-34s SHAPE %s vs %s" % (name, tuple(a.shape), tuple(b.shape)))
                break
            d = (a.double() - b.double()).abs()
            mx = float(d.max()) if d.numel() else 0.0
            if mx > 0:
                rows = int((d.reshape(d.shape[0], -1).amax(1) > 1e-6 * float(b.abs().max())).sum())
                print("   %-34s max|d| %.3e (max|x| %.3e)  rows beyond 1e-6: %d of %d" % (
                    name, mx, float(b.abs().max()), rows, d.shape[0]))


if __name__ == "__main__":
    main()
