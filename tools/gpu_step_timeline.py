"""GPU timeline of one distillation step: where the device idles.

Runs the bench workload under torch.profiler (kineto), writes the chrome trace (gzip) and prints, per phase of the step
(record_function ranges projected onto the GPU timeline by kineto), the GPU window, the kernel-busy time inside it and the
idle remainder, then the largest gaps with the kernels on either side.  Analysis only (numbers under a profiler are not
bench values); `python tools/gpu_step_timeline.py --analyse trace.json.gz` re-reads a trace without a GPU.
"""
import gzip
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def analyse(path, out=sys.stdout):
    opener = gzip.open if path.endswith(".gz") else open
    with opener(path, "rt") as f:
        tr = json.load(f)
    ev = tr["traceEvents"] if isinstance(tr, dict) else tr
    kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    kern.sort(key=lambda e: e["ts"])
    ann = [e for e in ev if e.get("cat") == "gpu_user_annotation" and "dur" in e]
    if not kern:
        print("no kernels in trace", file=out)
        return
    t0, t1 = kern[0]["ts"], max(e["ts"] + e["dur"] for e in kern)

    def busy_idle(lo, hi):
        """union of kernel intervals clipped to [lo, hi)"""
        busy, cur_end, n = 0.0, lo, 0
        for e in kern:
            s, t = e["ts"], e["ts"] + e["dur"]
            if t <= lo or s >= hi:
                continue
            n += 1
            s, t = max(s, cur_end), min(t, hi)
            if t > s:
                busy += t - s
                cur_end = t
        return busy, (hi - lo) - busy, n

    b, i, n = busy_idle(t0, t1)
    print("step window %.3f ms: %d kernels, busy %.3f ms, idle %.3f ms (%.1f %%)" % ((t1 - t0) / 1e3, n, b / 1e3, i / 1e3,
                                                                                    100 * i / (t1 - t0)), file=out)
    print("\nphases (gpu_user_annotation):", file=out)
    for a in sorted(ann, key=lambda e: e["ts"]):
        if not a["name"].startswith("ud:"):
            continue
        b, i, n = busy_idle(a["ts"], a["ts"] + a["dur"])
        print("  %-34s start %8.3f  window %8.3f ms  kernels %5d  busy %8.3f  idle %7.3f" % (
            a["name"], (a["ts"] - t0) / 1e3, a["dur"] / 1e3, n, b / 1e3, i / 1e3), file=out)
    # gaps
    gaps = []
    cur_end, prev = kern[0]["ts"] + kern[0]["dur"], kern[0]
    for e in kern[1:]:
        if e["ts"] > cur_end:
            gaps.append((e["ts"] - cur_end, prev["name"], e["name"], (cur_end - t0) / 1e3))
        if e["ts"] + e["dur"] > cur_end:
            cur_end, prev = e["ts"] + e["dur"], e
    hist = [(0, 2), (2, 5), (5, 10), (10, 20), (20, 50), (50, 100), (100, 1e9)]
    print("\ngap histogram (us): count, total ms", file=out)
    for lo, hi in hist:
        g = [x[0] for x in gaps if lo <= x[0] < hi]
        print("  [%4g, %4g)  %5d  %8.3f" % (lo, hi, len(g), sum(g) / 1e3), file=out)
    print("\nlargest gaps:", file=out)
    for g, p, nx, at in sorted(gaps, reverse=True)[:40]:
        print("  %8.1f us at %8.3f ms   %-60.60s -> %-60.60s" % (g, at, p, nx), file=out)
    # top kernels by total time
    tot = {}
    for e in kern:
        k = e["name"][:90]
        c = tot.setdefault(k, [0.0, 0])
        c[0] += e["dur"]
        c[1] += 1
    print("\ntop kernels:", file=out)
    for k, (d, c) in sorted(tot.items(), key=lambda kv: -kv[1][0])[:45]:
        print("  %8.3f ms  %5d x  %s" % (d / 1e3, c, k), file=out)


def main():
    if "--analyse" in sys.argv:
        analyse(sys.argv[sys.argv.index("--analyse") + 1])
        return
    for p in (ROOT, os.path.join(ROOT, "cvpr2023-unidistill_b200")):
        sys.path.insert(0, p)
    import torch
    from torch.profiler import ProfilerActivity, profile, record_function

    import bench

    name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "distill_lidar"
    outdir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(outdir, exist_ok=True)
    dev = torch.device("cuda:0")
    torch.backends.cudnn.benchmark = True
    wl = bench.WORKLOADS[name](dev, 0)
    st = wl.step_fn

    def wrap(obj, attr, label):
        fn = getattr(obj, attr)

        def inner(*a, **k):
            with record_function("ud:" + label):
                return fn(*a, **k)
        setattr(obj, attr, inner)

    wrap(st.student, "forward", "student.forward")
    wrap(st.teacher, "forward", "teacher.forward")
    for nm in ("lidar_encoder", "camera_encoder", "fusion_encoder", "bev_encoder", "det_head"):
        for model, tag in ((st.student, "student"), (st.teacher, "teacher")):
            m = getattr(model, nm, None)
            if m is not None:
                wrap(m, "forward", "%s.%s" % (tag, nm))
    wrap(st.bucket, "finish", "bucket.finish")
    if st.opt is not None:
        wrap(st.opt, "step", "optimizer.step")
    orig_losses = st.losses

    def losses(batch):
        with record_function("ud:losses(fwd)"):
            return orig_losses(batch)
    st.losses = losses
    import torch as _t
    orig_backward = _t.Tensor.backward

    def backward(self, *a, **k):
        with record_function("ud:backward"):
            return orig_backward(self, *a, **k)
    _t.Tensor.backward = backward

    for _ in range(4):
        wl.step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        with record_function("ud:step"):
            wl.step()
        torch.cuda.synchronize()
    raw = os.path.join(outdir, "timeline_%s.json" % name)
    prof.export_chrome_trace(raw)
    with open(raw, "rb") as f, gzip.open(raw + ".gz", "wb") as g:
        g.write(f.read())
    os.remove(raw)
    with open(os.path.join(outdir, "timeline_%s.txt" % name), "w") as f:
        analyse(raw + ".gz", f)
    analyse(raw + ".gz")


if __name__ == "__main__":
    main()
