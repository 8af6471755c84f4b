"""GPU debugging aid: one distillation step under autograd anomaly detection + per-parameter gradient finiteness."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "distill_lidar"
dev = torch.device("cuda:0")
wl = bench.WORKLOADS[name](dev, 0)
st = wl.step_fn
st.bucket.zero()
with torch.autograd.detect_anomaly(check_nan=True):
    terms = st.losses(wl.dev)
print({k: float(v) for k, v in terms.items()}, flush=True)
# which loss term poisons the gradient?  backward each term separately (retain the graph)
L = st.last if st.keep_features else None
names = dict((id(p), n) for n, p in st.student.named_parameters())
for key in ("loss",):
    st.bucket.zero()
    try:
        with torch.autograd.detect_anomaly(check_nan=True):
            terms[key].backward(retain_graph=True)
        print(key, "backward: no anomaly", flush=True)
    except RuntimeError as e:
        print(key, "ANOMALY:", str(e)[:3000], flush=True)
    bad = [names[id(p)] for p in st.bucket.params if not bool(torch.isfinite(p.grad).all())]
    print("params with non-finite grad: %d of %d" % (len(bad), len(st.bucket.params)), bad[:40], flush=True)
