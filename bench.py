#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 BEV feature-extraction / distillation path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference path

Workload at N=1 (BASELINE.json configs[1]): BEVFusion-LiDAR forward = hard voxelization -> mean VFE ->
VoxelResBackBone8x sparse encoder -> scatter to a 128x128 BEV -> BaseBEVBackbone -> CenterHead towers
(+ the target assignment CenterHead.forward performs in training mode), bs=4 per GPU, every sample =
10 sweeps x 40k synthetic ring-LiDAR points.  N>1: every rank runs its own batch (samples are independent;
a forward has no exchange step) -> weak scaling.  `--workload camera|distill|distill_lidar` select configs[2] / [3] / [4].

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "cvpr2023-unidistill_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "samples/sec BEVFusion fwd+distill (128x128 BEV)"
UNIT = "samples/s"
BS_PER_GPU = 4


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "tensor_tflops": float(p.get("bf16_tflops_sustained",
                                                                             p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json; sustained bf16 cuBLAS figure, kernel timed inside a step)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 100 ms while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nvml = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nvml = None

    def run(self):
        while not self._stop_evt.is_set():
            if self._nvml is not None:
                try:
                    n = self._nvml
                    self.samples.append(int(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM)))
                    r = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                    for name, bit in (("gpu_idle", 0x1), ("applications_clocks_setting", 0x2),
                                      ("sw_power_cap", 0x4), ("hw_slowdown", 0x8), ("sync_boost", 0x10),
                                      ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40),
                                      ("hw_power_brake_slowdown", 0x80)):
                        if r & bit and name != "gpu_idle":
                            self.reasons.add(name)
                except Exception:
                    pass
            else:
                self._smi_once()
            self._stop_evt.wait(float(os.environ.get("UD_BENCH_SAMPLE_S", "0.25")))

    def _smi_once(self):
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
            f = [x.strip() for x in out.strip().split(",")]
            self.samples.append(int(f[0]))
            self.max_mhz = int(f[1])
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        except Exception:
            pass

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------------
def _dist_setup(n_gpus):
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return world, rank, local


def _barrier(world):
    import torch

    if world > 1:
        import torch.distributed as dist

        dist.barrier()
    torch.cuda.synchronize()


def _max_over_ranks(x, world, device):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _sum_over_ranks(x, world, device):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class L2Flusher:
    """Writes a buffer larger than the 126 MB L2 between timed steps."""

    def __init__(self, device, mib=256):
        import torch

        self.buf = torch.empty(mib << 20, dtype=torch.uint8, device=device)
        self.v = 0

    def __call__(self):
        self.v = (self.v + 1) & 0x7F
        self.buf.fill_(self.v)


# ---------------------------------------------------------------------------------------------------
# workloads of this repo's arm
def _gt_with_labels(batch):
    import torch

    # reference training_step: gt_boxes (B,M,9) + (label+1) -> (B,M,10)  (..._fusion_exp.py:208-212)
    return torch.cat([batch["gt_boxes"], (batch["gt_labels"].float() + 1).unsqueeze(-1)], dim=-1)


class PipelinedE2E:
    """End-to-end stepping for the forward workloads: every step copies ITS inputs from pinned host memory and reads ITS
    result back to pinned host memory, but the copies run on a second stream so that the host->device copy of step i+1 and
    the device->host read of step i-1 overlap the kernels of step i (double-buffered device inputs / host outputs).
    All K copies in and K reads out happen between the start and the end event of the timed region."""

    def __init__(self, device, host_inputs, forward, result_of):
        import torch

        self.device, self.host_inputs, self.forward, self.result_of = device, host_inputs, forward, result_of
        self.copy_stream = torch.cuda.Stream(device=device)
        self.dev_in = [None, None]
        self.in_ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.in_free = [torch.cuda.Event(), torch.cuda.Event()]
        self.out_host = [None, None]
        self.out_done = [torch.cuda.Event(), torch.cuda.Event()]
        self.i = 0
        self.prefetched = False

    def _h2d(self, slot):
        import torch

        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.in_free[slot])            # the step that last read this buffer is done
            if self.dev_in[slot] is None:
                self.dev_in[slot] = [t.to(self.device, non_blocking=True) for t in self.host_inputs]
            else:
                for d, h in zip(self.dev_in[slot], self.host_inputs):
                    d.copy_(h, non_blocking=True)
            self.in_ready[slot].record(self.copy_stream)

    def begin(self):
        """Start of a timed region: nothing is in flight; the first step's copy is issued inside the region."""
        import torch

        self.i = 0
        self.prefetched = False
        cur = torch.cuda.current_stream(self.device)
        for e in self.in_free:
            e.record(cur)

    def step(self, last=False):
        import torch

        cur = torch.cuda.current_stream(self.device)
        slot = self.i % 2
        if not self.prefetched:
            self._h2d(slot)
        if not last:
            self._h2d(1 - slot)                                        # next step's inputs, overlapping this step
            self.prefetched = True
        cur.wait_event(self.in_ready[slot])
        res = self.result_of(self.forward(*self.dev_in[slot]))
        self.in_free[slot].record(cur)
        done = torch.cuda.Event()
        done.record(cur)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(done)
            if self.out_host[slot] is None:
                self.out_host[slot] = torch.empty(res.shape, dtype=res.dtype, pin_memory=True)
            self.out_host[slot].copy_(res, non_blocking=True)
            self.out_done[slot].record(self.copy_stream)
        res.record_stream(self.copy_stream)
        self.i += 1
        return self.out_host[slot]

    def end(self):
        """Joins the copy stream into the current stream (before the closing event of the timed region)."""
        import torch

        cur = torch.cuda.current_stream(self.device)
        for e in self.out_done:
            cur.wait_event(e)


class LidarForward:
    """configs[1]: BEVFusion-LiDAR SparseEncoder + CenterHead fwd, synthetic 40k pts x 10 sweeps, bs=4."""

    name = "BEVFusion-LiDAR SparseEncoder+CenterHead fwd, 40k pts x10 sweeps, bs=4/GPU, 128x128x5 BEV"

    def __init__(self, device, rank, seed=0):
        import torch

        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.device = device
        cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
        self.model = models.BEVFusionCenterHead(cfg).to(device).train()   # BN on batch statistics (student and
        self.model.use_channels_last(os.environ.get("UD_CHANNELS_LAST", "1") != "0")  # teacher both run so, quirk 9)
        self.model.use_cuda_graph(os.environ.get("UD_DENSE_GRAPH", "1") != "0")
        batch = S.lidar_batch(BS_PER_GPU, seed=seed + rank)
        self.points_host = batch["points"].pin_memory()                   # (B, Nmax, 5) zero padded (collate_fn)
        self.gt_host = _gt_with_labels(batch).pin_memory()
        self.points_dev = self.points_host.to(device)
        self.gt_dev = self.gt_host.to(device)
        self.h2d = self.points_host.numel() * 4 + self.gt_host.numel() * 4
        self.d2h = BS_PER_GPU * 10 * 128 * 128 * 4                          # class heat-maps (B, 10, 128, 128) f32
        self.samples = BS_PER_GPU
        import torch as _t

        self.pipe = PipelinedE2E(device, [self.points_host, self.gt_host], self._forward,
                                 lambda r: _t.cat([h["hm"] for h in r[2]], dim=1))

    def _forward(self, points, gt):
        import torch

        with torch.no_grad():
            pts = [points[b] for b in range(points.shape[0])]     # training_step splits the padded batch
            bev, feat, heads = self.model(pts, None, None, gt, return_feature=True)
        return bev, feat, heads

    def step(self):
        return self._forward(self.points_dev, self.gt_dev)


class CameraForward:
    """configs[2]: BEVFusion-camera LSS (fused lift-splat) + CenterHead fwd, 6 x 256x704, bs=4."""

    name = "BEVFusion-camera LSS lift-splat+CenterHead fwd, 6 cams 256x704, bs=4/GPU, 128x128 BEV"

    def __init__(self, device, rank, seed=0):
        import torch

        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.device = device
        cfg = models.make_model_cfg(bev=128, lidar=False, camera=True)
        self.model = models.BEVFusionCenterHead(cfg).to(device).train()
        self.model.use_channels_last(os.environ.get("UD_CHANNELS_LAST", "1") != "0")
        cam = S.camera_batch(BS_PER_GPU, seed=seed + rank)
        lid = S.lidar_batch(BS_PER_GPU, seed=seed + rank, points_per_sweep=64, sweeps=1)
        self.imgs_host = cam["imgs"].pin_memory()
        self.mats_host = {k: v.pin_memory() for k, v in cam["mats_dict"].items()}
        self.gt_host = _gt_with_labels(lid).pin_memory()
        self.imgs_dev = self.imgs_host.to(device)
        self.mats_dev = {k: v.to(device) for k, v in self.mats_host.items()}
        self.gt_dev = self.gt_host.to(device)
        self.h2d = (self.imgs_host.numel() + sum(v.numel() for v in self.mats_host.values())
                    + self.gt_host.numel()) * 4
        self.d2h = BS_PER_GPU * 10 * 128 * 128 * 4
        self.samples = BS_PER_GPU
        self._mat_keys = sorted(self.mats_host)
        import torch as _t

        self.pipe = PipelinedE2E(
            device, [self.imgs_host, self.gt_host] + [self.mats_host[k] for k in self._mat_keys],
            lambda imgs, gt, *mats: self._forward(imgs, dict(zip(self._mat_keys, mats)), gt),
            lambda r: _t.cat([h["hm"] for h in r[2]], dim=1))

    def _forward(self, imgs, mats, gt):
        import torch

        with torch.no_grad():
            return self.model(None, imgs, mats, gt, return_feature=True)

    def step(self):
        return self._forward(self.imgs_dev, self.mats_dev, self.gt_dev)


class DistillTrainStep:
    """configs[3]: camera-student <- LiDAR-teacher distillation step (det loss + feature/relation/response
    distillation, backward, one NCCL gradient all-reduce, AdamW), bs=4 per GPU."""

    name = ("camera-student <- LiDAR-teacher distill step (fwd teacher+student, 3 distill losses, bwd, grad "
            "all-reduce, AdamW), bs=4/GPU, 128x128 BEV")
    teacher_modalities = dict(lidar=True, camera=False)
    student_modalities = dict(lidar=False, camera=True)
    bn_note = ("student: train-mode batch statistics under autograd; teacher: frozen, eval mode, BatchNorm2d folded "
               "into its convolutions, dense trunk replayed from a CUDA graph")

    def __init__(self, device, rank, seed=0):
        import torch

        from unidistill_b200 import models, synthetic as S
        from unidistill_b200.step import DistillStep

        torch.manual_seed(0)
        self.device = device
        cl = os.environ.get("UD_CHANNELS_LAST", "1") != "0"
        t_cfg = models.make_model_cfg(bev=128, **self.teacher_modalities)
        s_cfg = models.make_model_cfg(bev=128, **self.student_modalities)
        teacher = models.BEVFusionCenterHead(t_cfg).to(device).use_channels_last(cl)
        student = models.BEVFusionCenterHead(s_cfg).to(device).use_channels_last(cl)
        self.step_fn = DistillStep(student, teacher, t_cfg["geometry"])
        lid = S.lidar_batch(BS_PER_GPU, seed=seed + rank)
        cam = S.camera_batch(BS_PER_GPU, seed=seed + rank)
        self.host = {"points": lid["points"].pin_memory(), "imgs": cam["imgs"].pin_memory(),
                     "gt_boxes": lid["gt_boxes"].pin_memory(), "gt_labels": lid["gt_labels"].pin_memory(),
                     "mats_dict": {k: v.pin_memory() for k, v in cam["mats_dict"].items()}}
        self.dev = self._to_device(self.host, False)
        self.h2d = (self.host["points"].numel() + self.host["imgs"].numel() + self.host["gt_boxes"].numel()
                    + sum(v.numel() for v in self.host["mats_dict"].values())) * 4 + self.host["gt_labels"].numel() * 8
        self.d2h = 0
        self.samples = BS_PER_GPU

    def _to_device(self, h, nb):
        d = {k: (v.to(self.device, non_blocking=nb) if not isinstance(v, dict) else
                 {kk: vv.to(self.device, non_blocking=nb) for kk, vv in v.items()}) for k, v in h.items()}
        d["points"] = [d["points"][b] for b in range(d["points"].shape[0])]
        return d

    def step(self):
        return self.step_fn(self.dev)

    def step_e2e(self):
        terms = self.step_fn(self._to_device(self.host, True))
        out = terms["loss"].cpu()
        self.d2h = 4
        return out


class DistillLidarTrainStep(DistillTrainStep):
    """configs[4]: LiDAR-student <- fusion-teacher distillation step (teacher = LiDAR + camera branches +
    FusionEncoder, frozen; student = LiDAR branch trained through the sparse-conv dgrad/wgrad kernels)."""

    name = ("LiDAR-student <- fusion-teacher distill step (fwd teacher+student, 3 distill losses, bwd, grad "
            "all-reduce, AdamW), 40k pts x10 sweeps, bs=4/GPU, 128x128 BEV")
    teacher_modalities = dict(lidar=True, camera=True)
    student_modalities = dict(lidar=True, camera=False)


WORKLOADS = {"lidar": LidarForward, "camera": CameraForward, "distill": DistillTrainStep,
             "distill_lidar": DistillLidarTrainStep}


# ---------------------------------------------------------------------------------------------------
def _roofline_from_records(records, peaks):
    """Groups the live per-launch timings of ud_sparse_conv_forward by (cin, cout, K) and reports the
    group with the largest total time (the dominant kernel of the step)."""
    groups = {}
    for key, meta, ms in records:
        if key != "sparse_conv_forward":
            continue
        g = groups.setdefault((meta["cin"], meta["cout"], meta["K"]), {"ms": 0.0, "launches": 0, "metas": []})
        g["ms"] += ms
        g["launches"] += 1
        g["metas"].append((meta, ms))
    if not groups:
        return None, 0.0
    total_conv_ms = sum(g["ms"] for g in groups.values())
    (cin, cout, K), g = max(groups.items(), key=lambda kv: kv[1]["ms"])
    pairs_cache = {}
    flops = byts = 0.0
    for meta, ms in g["metas"]:
        t = meta["table"]
        pk = (t.data_ptr(), tuple(t.shape))
        if pk not in pairs_cache:
            pairs_cache[pk] = int((t >= 0).sum().item())
        P = pairs_cache[pk]
        flops += 2.0 * P * cin * cout
        byts += 4.0 * (meta["n_in"] * cin + meta["n_out"] * cout + meta["n_out"] * K + K * cin * cout)
    avg_ms = g["ms"] / g["launches"]
    tflops = flops / g["launches"] / avg_ms / 1e9
    gbs = byts / g["launches"] / avg_ms / 1e6
    tensor_bound = cin >= 32           # C<=16 is gather-bandwidth bound, C>=32 tensor-pipe bound
    kname = "%s<%d,%d>" % (g["metas"][0][0].get("kernel", "k_spconv"), cin, cout)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")       # dram bytes/launch from the ncu --set full capture
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = (json.load(f).get(kname) or {}).get("dram_bytes_per_launch")
    roof = {
        "kernel": "%s (K=%d taps)" % (kname, K),
        "bound": "tensor" if tensor_bound else "hbm",
        "achieved": round(tflops if tensor_bound else gbs, 3),
        "peak": peaks["tensor_tflops"] if tensor_bound else peaks["hbm_gbs"],
        "unit": "TFLOP/s" if tensor_bound else "GB/s",
        "frac": round((tflops / peaks["tensor_tflops"]) if tensor_bound else (gbs / peaks["hbm_gbs"]), 5),
        "traffic": traffic,
        "avg_launch_ms": round(avg_ms, 4),
        "launches_per_step_group": g["launches"],
        "algorithmic_gflop_per_launch": round(flops / g["launches"] / 1e9, 3),
        "algorithmic_mb_per_launch": round(byts / g["launches"] / 1e6, 3),
        "achieved_gbs_algorithmic": round(gbs, 1),
        "peak_source": peaks["source"],
    }
    return roof, total_conv_ms


def run_ours(args):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference)")
    from unidistill_b200 import native

    native.load()
    torch.backends.cudnn.benchmark = os.environ.get("UD_CUDNN_BENCHMARK", "1") != "0"   # dense BEV/head convs: autotune
    world, rank, local = _dist_setup(args.gpus)
    device = torch.device("cuda", local if world > 1 else 0)
    wl = WORKLOADS[args.workload](device, rank)
    flush = L2Flusher(device)
    peaks = _peaks()

    for _ in range(max(args.warmup, 3)):
        wl.step()
    torch.cuda.synchronize()

    # ---- device-resident timing: K steps, per-step CUDA events, L2 flushed between steps ----------
    import gc

    gc.collect()
    gc.disable()          # keep generational collections out of the timed steps

    def timed_pass(with_kernel_events, with_sampler):
        """K steps, each bracketed by pre-created CUDA events, L2 flushed between steps."""
        native.TIMER.records = []
        if with_kernel_events:
            native.TIMER.prepare(32 * args.steps)
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.steps)]
        for a, b in events:                   # materialise the events before the region
            a.record()
            b.record()
        torch.cuda.synchronize()
        sampler = ClockSampler(local if world > 1 else 0) if with_sampler else None
        native.TIMER.enabled = with_kernel_events
        _barrier(world)
        if sampler is not None:
            sampler.start()
        native.launch_count(reset=True)
        wall0 = time.perf_counter()
        if args.profile:                      # ncu --profile-from-start off: capture exactly the timed steps
            torch.cuda.cudart().cudaProfilerStart()
        for i in range(args.steps):
            flush()
            e0, e1 = events[i]
            e0.record()
            wl.step()
            e1.record()
        _barrier(world)
        if args.profile:
            torch.cuda.cudart().cudaProfilerStop()
        wall = time.perf_counter() - wall0
        launches = native.launch_count(reset=True)
        clocks = sampler.stop() if sampler is not None else None
        native.TIMER.enabled = False
        return [a.elapsed_time(b) for a, b in events], native.TIMER.drain(), wall, launches, clocks

    # ---- pass 1 (the reported number): device-resident steps, clocks sampled during the region ----
    step_ms, _, wall, launches, clocks = timed_pass(False, True)
    total_ms = _max_over_ranks(sum(step_ms), world, device)
    samples_all = _sum_over_ranks(wl.samples * args.steps, world, device)
    value = samples_all / (total_ms / 1e3)
    # ---- pass 2: the same K steps with a CUDA-event pair around every sparse-conv launch (roofline).  Kept apart
    # from pass 1 because per-launch event records + NVML polling in one region produced 50-250 ms stalls on these
    # hosts (either alone is clean); its step times are reported as `roofline.pass_ms_per_step`.
    records, roof_step_ms = [], []
    if args.workload == "lidar" and not args.profile:
        roof_step_ms, records, _, _, _ = timed_pass(True, False)

    # ---- end-to-end: pinned host -> device copy + step + result read-back, every step -------------
    # One event pair around the K steps (copies of neighbouring steps overlap on a second stream for the forward
    # workloads, see PipelinedE2E; the distillation steps copy, run and read back in order on one stream).
    pipe = getattr(wl, "pipe", None)
    n_e2e = 1 if args.profile else args.steps

    def e2e_pass(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        _barrier(world)
        if pipe is not None:
            pipe.begin()
        e0.record()
        for i in range(n):
            flush()
            if pipe is not None:
                pipe.step(last=(i == n - 1))
            else:
                wl.step_e2e()
        if pipe is not None:
            pipe.end()
        e1.record()
        _barrier(world)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    if not args.profile:
        e2e_pass(2)
    e2e_ms = _max_over_ranks(e2e_pass(n_e2e), world, device) * (args.steps / n_e2e)
    gc.enable()
    e2e_value = samples_all / (e2e_ms / 1e3)

    roof, conv_ms = _roofline_from_records(records, peaks) if args.workload == "lidar" else (None, 0.0)

    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": round(total_ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "global_batch": BS_PER_GPU * world, "per_gpu_batch": BS_PER_GPU,
                   "bn": getattr(wl, "bn_note", "train-mode batch statistics, torch.no_grad"), "weights": "random init, seed 0",
                   "l2": "flushed (256 MiB fill) between timed steps; per-step CUDA events summed (e2e: one event pair around "
                         "the K steps, flushes included)",
                   "dense_convs": "cuDNN fp32 (torch default allow_tf32), NHWC memory format", "parallelism": "dp%d" % world},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(wl.h2d),
                "d2h_bytes_per_step": int(wl.d2h), "ms_per_step": round(e2e_ms / args.steps, 3)},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "wall_s_timed_region": round(wall, 3),
        "step_ms_min_med_max": [round(min(step_ms), 3), round(sorted(step_ms)[len(step_ms) // 2], 3),
                                round(max(step_ms), 3)],
        "step_ms": [round(x, 2) for x in step_ms],
    }
    if roof is not None:
        # all sparse-conv launches of a step (event-timed in pass 2) against the step time of pass 1: the per-launch event
        # records of pass 2 stretch the gaps between kernels, not the kernels
        roof["sparse_conv_ms_per_step"] = round(conv_ms / max(1, len(roof_step_ms)), 3)
        roof["share_of_step"] = round(conv_ms / max(1, len(roof_step_ms)) / max(1e-9, total_ms / args.steps), 4)
        roof["pass_ms_per_step"] = round(sum(roof_step_ms) / max(1, len(roof_step_ms)), 3)
        line["roofline"] = roof
    if args.profile:
        line["profiled"] = "run under a profiler range: numbers on this line are NOT bench values"
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/), timed on the host cores.
class CpuLidarForward:
    """One sample (one 400k-point cloud) per step through oracle/: C first-come voxelizer + mean VFE,
    spconv-Native restated with torch CPU ops (gather -> mm -> index_add_ per kernel offset), dense(),
    then the mirrored BaseBEVBackbone + CenterHead modules on torch CPU (the reference's own CPU path for the
    dense part)."""

    sample = "1 cloud (10 sweeps x 40k pts) per step, bs=1, same generator and weights as the GPU arm"

    def __init__(self, seed=0):
        import numpy as np
        import torch

        from oracle import lidar as O
        from unidistill.layers.blocks_3d.det3d import VoxelResBackBone8x
        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.O, self.S = O, S
        cfg = models.make_model_cfg(bev=128, lidar=True, camera=False)
        self.cfg = cfg
        bb = VoxelResBackBone8x(5, np.array(cfg.lidar_encoder.grid_size))
        self.sd = {k: v.detach().clone() for k, v in bb.state_dict().items()}
        self.bev_encoder = models.BevEncoder(cfg.bev_encoder).train()
        self.det_head = models.DetHead(cfg.det_head).train()
        batch = S.lidar_batch(BS_PER_GPU, seed=seed)
        self.clouds = [batch["points"][b].numpy() for b in range(BS_PER_GPU)]
        self.gt = _gt_with_labels(batch)
        self.i = 0

    def step(self):
        import torch

        O, S = self.O, self.S
        b = self.i % len(self.clouds)
        self.i += 1
        with torch.no_grad():
            _, coords, num, mean = O.voxelize([self.clouds[b]], S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128,
                                              10, 120000)
            f, c, shape = O.backbone_forward(self.sd, mean, coords, 1, [41, 1024, 1024], training=True)
            bev = torch.from_numpy(O.height_compression(f.numpy(), c, 1, shape))
            x = self.bev_encoder(bev)
            ret = self.det_head(x[0], self.gt[b:b + 1])
        return ret

    samples = 1


class CpuCameraForward:
    """Port of the reference's camera forward for one sample (bs=1, 6 cameras): image backbone / neck / depth net and
    LSS geometry on torch CPU (the reference's own Python, mirrored), then the reference's pooling path as written --
    the lifted (6*112*16*44, 256) volume is materialised (lss_fpn.py:289-310) and scattered by the C restatement of
    voxel_pooling_ext (oracle/camera_ref.c) -- then BEV backbone and head on torch CPU."""

    sample = "1 sample (6 cams 256x704) per step, bs=1, same generator and weights as the GPU arm"

    def __init__(self, seed=0):
        import torch

        from oracle import camera as OC
        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.OC = OC
        cfg = models.make_model_cfg(bev=128, lidar=False, camera=True)
        self.model = models.BEVFusionCenterHead(cfg).train()
        self.lss = self.model.camera_encoder.backbone
        cam = S.camera_batch(BS_PER_GPU, seed=seed)
        lid = S.lidar_batch(BS_PER_GPU, seed=seed, points_per_sweep=64, sweeps=1)
        self.imgs, self.mats = cam["imgs"], cam["mats_dict"]
        self.gt = _gt_with_labels(lid)
        self.i = 0

    def step(self):
        import torch

        OC, lss = self.OC, self.lss
        b = self.i % self.imgs.shape[0]
        self.i += 1
        with torch.no_grad():
            imgs = self.imgs[b:b + 1, 0:1]
            mats = {k: v[b:b + 1] for k, v in self.mats.items()}
            src = lss.get_cam_feats(imgs)[:, 0]
            ncam = src.shape[1]
            feat = lss._forward_depth_net(src.reshape(ncam, src.shape[2], src.shape[3], src.shape[4]), mats, 0)
            depth = feat[:, : lss.depth_channels].softmax(1)
            ctx = feat[:, lss.depth_channels: lss.depth_channels + lss.output_channels]
            geom = lss.get_geometry(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                                    mats.get("bda_mat", None))
            geom_xyz = OC.voxel_indices(geom, lss.voxel_coord, lss.voxel_size)
            bev = torch.from_numpy(OC.lift_splat_reference(depth, ctx, geom_xyz.numpy(), [int(v) for v in lss.voxel_num]))
            x = self.model.bev_encoder(bev)
            ret = self.model.det_head(x[0], self.gt[b:b + 1])
        return ret

    samples = 1


CPU_WORKLOADS = {"lidar": CpuLidarForward, "camera": CpuCameraForward}


def cpu_baseline(workload, budget_s=25.0):
    import torch

    if workload not in CPU_WORKLOADS:
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                "sample": "not implemented for the training-step workloads"}
    torch.set_num_threads(os.cpu_count() or 1)
    wl = CPU_WORKLOADS[workload]()
    wl.step()                                   # warm-up (builds the C oracle, primes torch's thread pool)
    t0 = time.perf_counter()
    n = 0
    while True:
        wl.step()
        n += 1
        el = time.perf_counter() - t0
        if el >= budget_s or n >= 8:
            break
    return {"value": round(n * wl.samples / el, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d x [%s], %.1f s" % (n, wl.sample, el)}


def run_reference(args):
    """The reference's CPU implementation of the path (spconv / voxel_pooling_ext are binary wheels that can
    neither be imported nor compiled here, so this is the oracle port), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    if args.workload not in CPU_WORKLOADS:
        print(json.dumps({"impl": "reference", "unavailable": "reference arm implements the forward workloads only"}))
        return
    wl = CPU_WORKLOADS[args.workload]()
    for _ in range(max(1, min(args.warmup, 2))):      # bounded: one CPU step is seconds long
        wl.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        wl.step()
    el = time.perf_counter() - t0
    value = args.steps * wl.samples / el
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": max(1, min(args.warmup, 2)),
        "ms_per_step": round(el / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload].name, "sample_per_step": wl.sample},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d x [%s]" % (args.steps, wl.sample)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lidar", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="bracket the timed steps with cudaProfilerStart/Stop (for ncu --profile-from-start off)")
    ap.add_argument("--cpu-budget", type=float, default=25.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
