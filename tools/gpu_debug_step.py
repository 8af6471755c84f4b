"""GPU debugging aid: runs the distillation step a few times printing every loss term, the gradient norm and the number
of non-finite student parameters, then an aten-op table (torch.profiler, grouped by op + input shapes) of one step."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
    sys.path.insert(0, p)
import bench  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "distill_lidar"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 14
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
wl = bench.WORKLOADS[name](dev, 0)
st = wl.step_fn
names = dict((id(p), n) for n, p in st.student.named_parameters())
every = max(1, steps // 14)
for i in range(steps):
    t = wl.step()
    if i % every and i != steps - 1:
        continue
    bad = sum(int((~torch.isfinite(p)).sum()) for p in st.student.parameters())
    badb = sum(int((~torch.isfinite(b)).sum()) for b in st.student.buffers() if b.is_floating_point())
    print("step %2d  " % i + "  ".join("%s=%.5g" % (k, float(v)) for k, v in t.items()) + "  nonfinite params=%d buffers=%d"
          % (bad, badb), flush=True)
top = sorted(((float(p.grad.norm()), names[id(p)], float(p.detach().norm())) for p in st.bucket.params), reverse=True)[:12]
print("largest (clipped) gradient norms after the last step; parameter norm:")
for g, n, w in top:
    print("   %.4g  %s  |w|=%.4g" % (g, n, w))
if "--profile" in sys.argv:
    from torch.profiler import ProfilerActivity, profile

    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        wl.step()
        torch.cuda.synchronize()
    print(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=70,
                                                             max_name_column_width=50, max_shapes_column_width=70))
