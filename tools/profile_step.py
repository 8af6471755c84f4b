"""torch.profiler view of the LiDAR forward step: GPU busy time vs wall time per step, top kernels, idle gaps."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "cvpr2023-unidistill_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
import bench

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
wl = bench.LidarForward(dev, 0)
for _ in range(5):
    wl.step()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        wl.step()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.end - e.time_range.start for e in evs)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
print("steps %d: wall %.2f ms/step, GPU busy %.2f ms/step, idle %.2f ms/step, kernels/step %d" %
      (N, (t1 - t0) / 1e3 / N, busy / 1e3 / N, ((t1 - t0) - busy) / 1e3 / N, len(evs) // N))
# largest idle gaps and what follows them
gaps = []
end = evs[0].time_range.end
for a in evs[1:]:
    g = a.time_range.start - end
    if g > 20:
        gaps.append((g, a.name[:70]))
    end = max(end, a.time_range.end)
gaps.sort(reverse=True)
print("idle gaps > 20 us: %d, total %.2f ms/step" % (len(gaps), sum(g for g, _ in gaps) / 1e3 / N))
agg = {}
for g, n in gaps:
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += g
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print("  %7.1f us/step  x%-3d before %s" % (t / N, c // N if c >= N else c, n))
