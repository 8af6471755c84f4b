#!/usr/bin/env python
"""bench.py -- the headline benchmark of the B200 BEV feature-extraction / distillation path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference path

Headline workload (BASELINE.json metric "samples/sec BEVFusion fwd+distill (128x128 BEV)", configs[4]): ONE OPTIMISATION
STEP of the LiDAR-student <- fusion-teacher distillation experiment -- teacher forward (LiDAR + camera branches +
FusionEncoder, frozen) + student forward + detection loss + the three distillation losses + backward + gradient all-reduce
over NCCL (overlapped with the backward) + gradient-norm clip + AdamW -- bs = 4 per GPU, every sample = 10 sweeps x 40k
synthetic ring-LiDAR points + 6 cameras 256x704, 128x128 BEV.  N > 1: every rank runs its own batch, gradients are averaged
(weak scaling).  The same line carries, under "secondary", the camera-student <- LiDAR-teacher step (configs[3]) and, with
--all, the forward-only workloads (configs[1], [2]), the voxelize + scatter micro-workload (configs[0]) and the
configs[4] sweep (points x BEV size).  `--workload` picks another primary.

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
                "source": "measured (MEASURED_PEAKS.json; sustained bf16 cuBLAS figure, kernels timed inside a step)"}
    return {"hbm_gbs": 6650.0, "tensor_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU every 250 ms while the timed region runs."""

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


def _reduce(x, world, device, op):
    import torch

    if world == 1:
        return x
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op={"max": dist.ReduceOp.MAX, "sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN}[op])
    return float(t.item())


def _gather(x, world, device):
    import torch

    if world == 1:
        return [x]
    import torch.distributed as dist

    t = torch.tensor([x], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


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
    """End-to-end stepping: every step copies ITS inputs from pinned host memory and reads ITS result back to pinned host
    memory, but the copies run on a second stream so that the host->device copy of step i+1 and the device->host read of
    step i-1 overlap the kernels of step i (double-buffered device inputs / host outputs).  All K copies in and K reads
    out happen between the start and the end event of the timed region."""

    def __init__(self, device, host_inputs, forward, result_of, prefetch=None):
        import torch

        self.device, self.host_inputs, self.forward, self.result_of = device, host_inputs, forward, result_of
        self.prefetch = prefetch        # optional: prefetch(ready_event, *next step's device inputs) after a step's launch
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
        if self.prefetch is not None and not last:
            self.prefetch(self.in_ready[1 - slot], *self.dev_in[1 - slot])
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


def _geometry(bev):
    from unidistill_b200 import synthetic as S

    return (S.PC_RANGE_128, 128) if bev == 128 else (S.PC_RANGE_180, 180)


class RawLidarInput:
    """The LiDAR input of a batch as a dataset hands it over: raw sweeps + sweep -> key-frame matrices, packed into five host
    tensors (unidistill_b200.input_side.pack_sweeps).  `clouds(dev_tensors)` runs the input-side kernel (sweep merge + x/y
    range filter + ragged collate, ud_collect_sweeps) and returns the per-sample clouds the model takes: no zero padding is
    ever created, copied or voxelized."""

    KEYS = ("points", "seg_offsets", "sample_first_seg", "mats", "time_lag")

    def __init__(self, device, raw_batch, pc_range):
        from unidistill_b200 import input_side as IS

        self.IS, self.pc_range = IS, pc_range
        packed = IS.pack_sweeps(raw_batch["samples"], pin=True)
        self.host = [packed[k] for k in self.KEYS]
        self.bytes = sum(t.numel() * t.element_size() for t in self.host)
        self.resident = self.clouds([t.to(device) for t in self.host])

    def clouds(self, dev_tensors):
        flat, offsets = self.IS.collect_sweeps(dict(zip(self.KEYS, dev_tensors)), self.pc_range)
        return self.IS.split_ragged(flat, offsets)


class LidarForward:
    """configs[1]: BEVFusion-LiDAR SparseEncoder + CenterHead fwd, synthetic 40k pts x 10 sweeps, bs=4."""

    config_id = 1

    def __init__(self, device, rank, seed=0, bev=128, points_per_sweep=40000, **_):
        import torch

        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.device = device
        self.name = ("BEVFusion-LiDAR SparseEncoder+CenterHead fwd, %dk pts x10 sweeps, bs=4/GPU, %dx%dx5 BEV"
                     % (points_per_sweep // 1000, bev, bev))
        cfg = models.make_model_cfg(bev=bev, lidar=True, camera=False)
        self.model = models.BEVFusionCenterHead(cfg).to(device).train()   # BN on batch statistics (student and
        self.model.use_channels_last(os.environ.get("UD_CHANNELS_LAST", "1") != "0")  # teacher both run so, quirk 9)
        self.model.use_cuda_graph(os.environ.get("UD_DENSE_GRAPH", "1") != "0")
        batch = S.lidar_raw_batch(BS_PER_GPU, seed=seed + rank, points_per_sweep=points_per_sweep,
                                  pc_range=_geometry(bev)[0])
        self.raw = RawLidarInput(device, batch, _geometry(bev)[0])          # raw sweeps; merged + filtered on the GPU
        self.gt_host = _gt_with_labels(batch).pin_memory()
        self.points_dev = self.raw.resident                                 # ragged list of per-sample clouds
        self.gt_dev = self.gt_host.to(device)
        self.h2d = self.raw.bytes + self.gt_host.numel() * 4
        self.d2h = BS_PER_GPU * 10 * bev * bev * 4                          # class heat-maps (B, 10, H, W) f32
        self.samples = BS_PER_GPU
        self.bn_note = "train-mode batch statistics, torch.no_grad"
        self.pipe = PipelinedE2E(device, self.raw.host + [self.gt_host],
                                 lambda *t: self._forward(self.raw.clouds(t[:-1]), t[-1]),
                                 lambda r: torch.cat([h["hm"] for h in r[2]], dim=1))

    def _forward(self, pts, gt):
        import torch

        with torch.no_grad():
            bev, feat, heads = self.model(pts, None, None, gt, return_feature=True)
        return bev, feat, heads

    def step(self):
        return self._forward(self.points_dev, self.gt_dev)


class CameraForward:
    """configs[2]: BEVFusion-camera LSS (fused lift-splat) + CenterHead fwd, 6 x 256x704, bs=4."""

    config_id = 2

    def __init__(self, device, rank, seed=0, bev=128, **_):
        import torch

        from unidistill_b200 import models, synthetic as S

        torch.manual_seed(0)
        self.device = device
        self.name = "BEVFusion-camera LSS lift-splat+CenterHead fwd, 6 cams 256x704, bs=4/GPU, %dx%d BEV" % (bev, bev)
        cfg = models.make_model_cfg(bev=bev, lidar=False, camera=True)
        self.model = models.BEVFusionCenterHead(cfg).to(device).train()
        self.model.use_channels_last(os.environ.get("UD_CHANNELS_LAST", "1") != "0")
        cam = S.camera_batch(BS_PER_GPU, seed=seed + rank)
        lid = S.lidar_batch(BS_PER_GPU, seed=seed + rank, points_per_sweep=64, sweeps=1, pc_range=_geometry(bev)[0])
        self.imgs_host = cam["imgs"].pin_memory()
        self.mats_host = {k: v.pin_memory() for k, v in cam["mats_dict"].items()}
        self.gt_host = _gt_with_labels(lid).pin_memory()
        self.imgs_dev = self.imgs_host.to(device)
        self.mats_dev = {k: v.to(device) for k, v in self.mats_host.items()}
        self.gt_dev = self.gt_host.to(device)
        self.h2d = (self.imgs_host.numel() + sum(v.numel() for v in self.mats_host.values())
                    + self.gt_host.numel()) * 4
        self.d2h = BS_PER_GPU * 10 * bev * bev * 4
        self.samples = BS_PER_GPU
        self.bn_note = "train-mode batch statistics, torch.no_grad"
        self._mat_keys = sorted(self.mats_host)
        self.pipe = PipelinedE2E(
            device, [self.imgs_host, self.gt_host] + [self.mats_host[k] for k in self._mat_keys],
            lambda imgs, gt, *mats: self._forward(imgs, dict(zip(self._mat_keys, mats)), gt),
            lambda r: torch.cat([h["hm"] for h in r[2]], dim=1))

    def _forward(self, imgs, mats, gt):
        import torch

        with torch.no_grad():
            return self.model(None, imgs, mats, gt, return_feature=True)

    def step(self):
        return self._forward(self.imgs_dev, self.mats_dev, self.gt_dev)


class DistillTrainStep:
    """configs[3]: camera-student <- LiDAR-teacher distillation step (det loss + feature/relation/response
    distillation, backward, NCCL gradient all-reduce, clip, AdamW), bs=4 per GPU."""

    config_id = 3
    title = "camera-student <- LiDAR-teacher"
    teacher_modalities = dict(lidar=True, camera=False)
    student_modalities = dict(lidar=False, camera=True)
    preset = "camera_exp_distill_lidar"

    def __init__(self, device, rank, seed=0, bev=128, points_per_sweep=40000, teacher_mode="eval", **_):
        import torch

        from unidistill_b200 import models, synthetic as S
        from unidistill_b200.step import DistillStep

        torch.manual_seed(0)
        self.device = device
        self.name = ("%s distill step (fwd teacher+student, det + 3 distill losses, bwd, grad all-reduce, clip 0.1, AdamW), "
                     "%dk pts x10 sweeps + 6 cams 256x704, bs=4/GPU, %dx%d BEV"
                     % (self.title, points_per_sweep // 1000, bev, bev))
        cl = os.environ.get("UD_CHANNELS_LAST", "1") != "0"
        t_cfg = models.make_model_cfg(bev=bev, **self.teacher_modalities)
        s_cfg = models.make_model_cfg(bev=bev, **self.student_modalities)
        teacher = models.BEVFusionCenterHead(t_cfg).to(device).use_channels_last(cl)
        student = models.BEVFusionCenterHead(s_cfg).to(device).use_channels_last(cl)
        self.step_fn = DistillStep(student, teacher, t_cfg["geometry"], preset=self.preset, teacher_mode=teacher_mode,
                                   overlap_allreduce=os.environ.get("UD_OVERLAP_ALLREDUCE", "1") != "0",
                                   student_cuda_graph=os.environ.get("UD_STUDENT_GRAPH", "1") != "0",
                                   overlap_teacher_camera=os.environ.get("UD_OVERLAP_CAMERA", "0") != "0")
        self.bn_note = ("student: train-mode batch statistics under autograd%s; teacher: frozen, %s" % (
            ", dense trunk + loss + their backward replayed from CUDA graphs" if student._train_graph is not None else "",
            "eval mode, BatchNorm2d folded into its convolutions, dense trunk replayed from a CUDA graph"
            if teacher_mode == "eval" else "BatchNorm on batch statistics (pytorch-lightning train mode, quirk 9), "
            "buffers restored every step, dense trunk replayed from a CUDA graph"))
        lid = S.lidar_raw_batch(BS_PER_GPU, seed=seed + rank, points_per_sweep=points_per_sweep, pc_range=_geometry(bev)[0])
        cam = S.camera_batch(BS_PER_GPU, seed=seed + rank)
        self.raw = RawLidarInput(device, lid, _geometry(bev)[0])            # raw sweeps; merged + filtered on the GPU
        self._mat_keys = sorted(cam["mats_dict"])
        rest = [cam["imgs"].pin_memory(), lid["gt_boxes"].pin_memory(), lid["gt_labels"].pin_memory()] + \
               [cam["mats_dict"][k].pin_memory() for k in self._mat_keys]
        self.host = self.raw.host + rest
        self.dev = self._batch(self.raw.resident, *[t.to(device) for t in rest])
        self.h2d = sum(t.numel() * t.element_size() for t in self.host)
        self.d2h = 4
        self.samples = BS_PER_GPU
        n_raw = len(self.raw.host)
        # The voxelization + rulebooks of the NEXT step are built on a side stream right after this step's launch, under its
        # backward (DistillStep.prefetch_geometry): every timed step still does one full geometry phase -- the next one's.
        self.prefetch_geometry = (os.environ.get("UD_PREFETCH_GEOMETRY", "1") != "0"
                                  and (self.teacher_modalities["lidar"] or self.student_modalities["lidar"]))
        self.geometry_note = ("voxelization + rulebooks of step i+1 built on a side stream right after step i's launch, under its "
                              "backward (one complete geometry phase per timed step; UD_PREFETCH_GEOMETRY=0: in line)"
                              if self.prefetch_geometry else "voxelization and rulebooks built inside every step")
        self._n_raw = n_raw
        self._next_clouds = {}
        self.pipe = PipelinedE2E(device, self.host, self._e2e_forward, lambda terms: terms["loss"].reshape(1),
                                 prefetch=self._e2e_prefetch if self.prefetch_geometry else None)

    def _e2e_forward(self, *t):
        clouds = self._next_clouds.pop(t[0].data_ptr(), None)
        if clouds is None:
            clouds = self.raw.clouds(t[:self._n_raw])
        return self.step_fn(self._batch(clouds, *t[self._n_raw:]))

    def _e2e_prefetch(self, ready_event, *t):
        import torch

        side = self.step_fn.geometry_stream()
        side.wait_event(ready_event)                 # the next step's host -> device copies
        with torch.cuda.stream(side):
            clouds = self.raw.clouds(t[:self._n_raw])        # input-side kernel (sweep merge, range filter, ragged collate)
        self._next_clouds.clear()
        self._next_clouds[t[0].data_ptr()] = clouds
        self.step_fn.prefetch_geometry(clouds)

    def _batch(self, clouds, imgs, gt_boxes, gt_labels, *mats):
        return {"points": clouds, "imgs": imgs, "gt_boxes": gt_boxes, "gt_labels": gt_labels,
                "mats_dict": dict(zip(self._mat_keys, mats))}

    def step(self):
        r = self.step_fn(self.dev)
        if self.prefetch_geometry:
            self.step_fn.prefetch_geometry(self.dev["points"])     # the NEXT step's geometry (here: the same resident clouds,
        return r                                                   # recomputed every step like any other batch's would be)


class DistillLidarTrainStep(DistillTrainStep):
    """configs[4]: LiDAR-student <- fusion-teacher distillation step (teacher = LiDAR + camera branches +
    FusionEncoder, frozen; student = LiDAR branch trained through the sparse-conv dgrad/wgrad kernels)."""

    config_id = 4
    title = "LiDAR-student <- fusion-teacher"
    teacher_modalities = dict(lidar=True, camera=True)
    student_modalities = dict(lidar=True, camera=False)
    preset = "lidar_exp_distill_fusion"


class VoxelizeScatter:
    """configs[0]: hard_voxelize (+ mean VFE) on the fine grid and voxelize + scatter-to-BEV on the coarse pillar grid
    (0.8 m x 0.8 m x 1.6 m -> [128,128,5] -> dense (1, 5*5, 128, 128)), 20k synthetic points, bs=1."""

    config_id = 0

    def __init__(self, device, rank, seed=0, **_):
        import torch

        from unidistill_b200 import synthetic as S

        self.device = device
        self.name = "hard_voxelize + MeanVFE (fine grid) + voxelize + scatter-to-BEV (pillar grid 128x128x5), 20k pts, bs=1"
        rng_pts = S.lidar_sample(seed + rank, points_per_sweep=20000, sweeps=1)
        self.points_host = torch.from_numpy(rng_pts).pin_memory()
        self.points_dev = self.points_host.to(device)
        self.h2d = self.points_host.numel() * 4
        self.d2h = 25 * 128 * 128 * 4
        self.samples = 1
        self.bn_note = "no BatchNorm on this path"
        self.pipe = PipelinedE2E(device, [self.points_host], self._forward, lambda r: r)

    def _forward(self, pts):
        from unidistill_b200 import sparse, synthetic as S
        from unidistill_b200.voxelize import hard_voxelize

        _, coords, num, mean, _ = hard_voxelize([pts], S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000,
                                                want_voxels=False)
        _, pc, pn, pmean, _ = hard_voxelize([pts], [0.8, 0.8, 1.6], S.PC_RANGE_128, [128, 128, 5], 32, 120000,
                                            want_voxels=False)
        dense = sparse.scatter_to_dense(pmean, pc, 1, [5, 128, 128])
        return dense.view(1, 25, 128, 128)

    def step(self):
        return self._forward(self.points_dev)


WORKLOADS = {"lidar": LidarForward, "camera": CameraForward, "distill": DistillTrainStep,
             "distill_lidar": DistillLidarTrainStep, "voxelize": VoxelizeScatter}


# ---------------------------------------------------------------------------------------------------
# per-entry-point roofline: algorithmic bytes / flops of one C-ABI call from its scalar arguments
def _lazy_int(cache, t, fn):
    key = (t.data_ptr(), tuple(t.shape))
    if key not in cache:
        cache[key] = int(fn(t))
    return cache[key]


def _algorithmic(name, a, meta, cache):
    """-> (label, bound, bytes, flops) for one recorded call; `a` = its arguments with pointers replaced by None."""
    meta = meta or {}
    if name == "ud_collect_sweeps":
        return "input side: sweep merge + range filter + ragged collate", "hbm", 2 * a[1] * a[2] * 4, 0
    if name == "ud_hard_voxelize":
        n, C, maxp = meta.get("n_points", 0), a[3], a[7]
        m = _lazy_int(cache, meta["counts"], lambda t: t[0].item()) if "counts" in meta else 0
        return "voxelize+MeanVFE", "hbm", n * 4 * C + m * (4 * C + 16 + 4) + (m * maxp * C * 4 if meta.get("voxels") else 0), 0
    if name == "ud_rulebook_subm":
        n, K = a[1], meta.get("K", 27)
        return "rulebook SubM (K=%d)" % K, "hbm", n * 16 + n * K * 4, 0
    if name in ("ud_rulebook_conv_count", "ud_rulebook_conv_fill"):
        n_in, K = a[1], meta.get("K", 27)
        n_out = a[10] if name.endswith("fill") else 0
        return ("rulebook strided %s (K=%d)" % (name.rsplit("_", 1)[1], K), "hbm",
                n_in * 16 + ((n_in + n_out) * K * 4 + n_out * 16 if n_out else 0), 0)
    if name == "ud_rulebook_transpose":
        n, K = a[1], a[2]
        return "rulebook tap-major transpose + tile masks", "hbm", 2 * n * K * 4 + (n + 127) // 128 * 4, 0
    if name == "ud_rulebook_windows":
        return "rulebook window starts (per tile and tap group)", "hbm", a[1] * a[2] * 4, 0
    if name == "ud_rulebook_pairs":
        return "rulebook spconv-pairs export", "hbm", 3 * a[1] * a[2] * 4, 0
    if name in ("ud_sparse_conv_tc_forward", "ud_sparse_conv_forward", "ud_sparse_conv_tc_forward_window"):
        tc = name != "ud_sparse_conv_forward"
        n_out, K, cin, cout = ((a[7], a[8], a[9], a[10]) if name.endswith("window") else (a[6], a[7], a[8], a[9])) if tc \
            else (a[4], a[5], a[6], a[7])
        n_in = meta.get("n_in", n_out)
        P = _lazy_int(cache, meta["table"], lambda t: (t >= 0).sum().item()) if "table" in meta else n_out * K
        kern = meta.get("kernel", "k_spconv_umma_dg" if tc else "k_spconv_gather_gemm")
        return ("%s<%d,%d> K=%d (fwd/dgrad)" % (kern, cin, cout, K), "tensor" if (tc and cin >= 32) else "hbm",
                4 * (n_in * cin + n_out * cout + n_out * K + K * cin * cout), 2.0 * P * cin * cout)
    if name == "ud_sparse_conv_tc_forward_affine":
        n_out, K, cin, cout = a[9], a[10], a[11], a[12]
        n_in = meta.get("n_in", n_out)
        P = _lazy_int(cache, meta["table"], lambda t: (t >= 0).sum().item()) if "table" in meta else n_out * K
        # the same kernel instantiation as ud_sparse_conv_tc_forward (epilogue flags are run-time): one roofline group
        return ("k_spconv_umma_dg<%d,%d> K=%d (fwd/dgrad)" % (cin, cout, K),
                "tensor" if cin >= 32 else "hbm",
                4 * (n_in * cin + n_out * cout * (2 if meta.get("residual") else 1) + n_out * K + K * cin * cout),
                2.0 * P * cin * cout)
    if name in ("ud_sparse_conv_tc_wgrad", "ud_sparse_conv_wgrad"):
        tc = name == "ud_sparse_conv_tc_wgrad"
        n_rows, K, cin, cout = (a[4], a[5], a[6], a[7]) if tc else (a[3], a[4], a[5], a[6])
        n_in = meta.get("n_in", n_rows)
        P = _lazy_int(cache, meta["table"], lambda t: (t >= 0).sum().item()) if "table" in meta else n_rows * K
        return ("%s<%d,%d> K=%d" % ("k_spconv_wgrad_umma" if tc else "k_spconv_wgrad", cin, cout, K),
                "tensor" if (tc and cin >= 32) else "hbm",
                4 * (n_in * cin + n_rows * cout + n_rows * K + K * cin * cout), 2.0 * P * cin * cout)
    if name == "ud_split_bf16":
        return "bf16 hi/lo split of conv inputs", "hbm", 2 * a[1] * a[2] * 4, 0
    if name == "ud_sparse_conv_tc_pack_weights":
        return "conv weight pack (bf16 hi/lo, swizzled)", "hbm", 2 * a[1] * a[2] * a[3] * 4, 0
    if name == "ud_bn_stats":
        return "BN1d statistics", "hbm", a[1] * a[2] * 4, 0
    if name == "ud_bn_apply":
        return "BN1d apply (+res)+ReLU", "hbm", 2 * a[1] * a[2] * 4, 0
    if name == "ud_bn_backward":
        return "BN1d backward (reduce + apply)", "hbm", 4 * a[3] * a[4] * 4, 0
    if name in ("ud_bn_finalize", "ud_bn_eval_coeffs"):
        return "BN1d coefficients", "latency", 0, 0
    if name in ("ud_scatter_to_bev", "ud_gather_from_bev"):
        n, B, C, D, H, W = a[2:8]
        return ("scatter-to-BEV" if name == "ud_scatter_to_bev" else "gather-from-BEV (scatter bwd)", "hbm",
                n * C * 4 + n * 16 + (B * C * D * H * W * 4 if name == "ud_scatter_to_bev" else n * C * 4), 0)
    if name == "ud_lift_splat_forward":
        B, ncam, D, H, W, C, nx, ny = a[0:8]
        pts = B * ncam * D * H * W
        return "lift-splat fwd", "hbm", pts * 4 + B * ncam * C * H * W * 4 + pts * 12 + B * C * nx * ny * 4, 2.0 * pts * C
    if name == "ud_lift_splat_backward":
        B, ncam, D, H, W, C, nx, ny = a[0:8]
        pts = B * ncam * D * H * W
        return ("lift-splat bwd", "hbm", B * C * nx * ny * 4 + 2 * pts * 4 + 2 * B * ncam * C * H * W * 4, 4.0 * pts * C)
    if name == "ud_voxel_pooling_forward":
        B, N, C, nx, ny = a[0:5]
        return "voxel_pooling fwd", "hbm", B * N * (C * 4 + 12) + B * nx * ny * C * 4, 0
    if name == "ud_voxel_pooling_backward":
        B, N, C, nx, ny = a[0:5]
        return "voxel_pooling bwd", "hbm", B * N * (C * 4 + 12) + B * nx * ny * C * 4, 0
    if name in ("ud_box_distill_forward", "ud_box_distill_backward", "ud_box_distill_forward_layout",
                "ud_box_distill_backward_layout"):
        if name.endswith("_layout"):
            mode, B, C, H, W, M = a[0], a[5], a[6], a[7], a[8], a[11]
        else:
            mode, B, C, H, W, M = a[0], a[3], a[4], a[5], a[6], a[9]
        direction = "backward" if "backward" in name else "forward"
        byts = 2 * B * M * 9 * 4 * C * 4 + (B * C * H * W * 4 if direction == "backward" else 0)
        return "%s distill %s" % ("feature" if mode == 0 else "relation", direction), "hbm", byts, 0
    if name in ("ud_response_distill_forward", "ud_response_distill_backward"):
        B, n_cls, n_reg, H, W = a[4:9]
        ch = n_cls + n_reg
        return ("response distill %s" % name.rsplit("_", 1)[1], "hbm",
                (2 + (1 if name.endswith("backward") else 0)) * ch * B * H * W * 4, 0)
    if name == "ud_box_corners_bev":
        return "GT corners", "latency", 0, 0
    if name == "ud_fcos_assign":
        return "FCOS target assignment (all tasks, whole batch)", "latency", 0, 0
    if name in ("ud_nms_bev", "ud_boxes_iou_bev"):
        return "rotated NMS / IoU", "latency", 0, 0
    return name, "latency", 0, 0


def _roofline_all(records, peaks, n_steps, step_ms):
    groups, cache = {}, {}
    for name, meta, a, ms in records:
        label, bound, byts, flops = _algorithmic(name, a, meta, cache)
        g = groups.setdefault(label, {"bound": bound, "ms": 0.0, "launches": 0, "bytes": 0.0, "flops": 0.0, "entry": name})
        g["ms"] += ms
        g["launches"] += 1
        g["bytes"] += byts
        g["flops"] += flops
    out = []
    for label, g in groups.items():
        ms = g["ms"]
        e = {"kernel": label, "entry": g["entry"], "bound": g["bound"], "calls_per_step": round(g["launches"] / n_steps, 2),
             "ms_per_step": round(ms / n_steps, 4), "avg_call_ms": round(ms / g["launches"], 4),
             "share_of_step": round(ms / n_steps / max(step_ms, 1e-9), 4)}
        if g["bound"] == "tensor":
            tf = g["flops"] / max(ms, 1e-9) / 1e9
            e.update(achieved=round(tf, 2), peak=peaks["tensor_tflops"], unit="TFLOP/s", frac=round(tf / peaks["tensor_tflops"], 4),
                     algorithmic_gflop_per_call=round(g["flops"] / g["launches"] / 1e9, 3),
                     algorithmic_mb_per_call=round(g["bytes"] / g["launches"] / 1e6, 3),
                     achieved_gbs_algorithmic=round(g["bytes"] / max(ms, 1e-9) / 1e6, 1))
        elif g["bound"] == "hbm":
            gb = g["bytes"] / max(ms, 1e-9) / 1e6
            e.update(achieved=round(gb, 1), peak=peaks["hbm_gbs"], unit="GB/s", frac=round(gb / peaks["hbm_gbs"], 4),
                     algorithmic_mb_per_call=round(g["bytes"] / g["launches"] / 1e6, 3))
        out.append(e)
    out.sort(key=lambda e: -e["ms_per_step"])
    return out


def _dominant(roof_all):
    """The entry with the largest time per step among the hand-written compute kernels -> the `roofline` object."""
    cands = [e for e in roof_all if e["bound"] in ("tensor", "hbm")]
    if not cands:
        return None
    e = dict(max(cands, key=lambda x: x["ms_per_step"]))
    kname = e["kernel"].split(" ")[0]
    traffic = None
    for fn in ("r03_traffic.json", "r02_traffic.json", "r01_traffic.json"):
        tpath = os.path.join(ROOT, "profiles", fn)                      # dram bytes/launch from an ncu --set full capture
        if os.path.exists(tpath):
            with open(tpath) as f:
                t = {k.replace(" ", ""): v for k, v in json.load(f).items()}
            if kname in t:
                traffic = t[kname].get("dram_bytes_per_launch")
                e["traffic_source"] = "profiles/" + fn
                break
    e["traffic"] = traffic
    return e


def measure(wl, args, world, rank, local, device, peaks, with_roofline=True):
    """Warm-up, K device-timed steps (pass 1), the same K steps with per-call events (pass 2, roofline), K end-to-end steps."""
    import gc

    import torch

    from unidistill_b200 import native

    flush = L2Flusher(device)
    for _ in range(max(args.warmup, 3)):
        wl.step()
    torch.cuda.synchronize()
    gc.collect()
    gc.disable()          # keep generational collections out of the timed steps

    def timed_pass(with_kernel_events, with_sampler):
        """K steps, each bracketed by pre-created CUDA events, L2 flushed between steps."""
        native.TIMER.records = []
        if with_kernel_events:
            native.TIMER.prepare(400 * args.steps)
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
    my_total = sum(step_ms)
    total_ms = _reduce(my_total, world, device, "max")
    samples_all = _reduce(wl.samples * args.steps, world, device, "sum")
    value = samples_all / (total_ms / 1e3)
    rank_ms = sorted(x / args.steps for x in _gather(my_total, world, device))
    # ---- pass 2: the same K steps with a CUDA-event pair around every C-ABI call (per-kernel roofline).  Kept apart
    # from pass 1 because per-launch event records + NVML polling in one region produced 50-250 ms stalls on these
    # hosts (either alone is clean); its step times are reported as `roofline_pass_ms_per_step`.
    roof_all, roof_step_ms = [], []
    if with_roofline and not args.profile:
        roof_step_ms, records, _, _, _ = timed_pass(True, False)
        roof_all = _roofline_all(records, peaks, args.steps, total_ms / args.steps)

    # ---- end-to-end: pinned host -> device copy + step + result read-back, every step (copies of neighbouring steps
    # overlap on a second stream, see PipelinedE2E); one event pair around the K steps
    pipe = wl.pipe
    n_e2e = 1 if args.profile else args.steps

    def e2e_pass(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        _barrier(world)
        pipe.begin()
        e0.record()
        for i in range(n):
            flush()
            pipe.step(last=(i == n - 1))
        pipe.end()
        e1.record()
        _barrier(world)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    if not args.profile:
        e2e_pass(4)         # untimed: fills the copy pipeline and (geometry prefetch) the side stream's memory pool
    e2e_ms = _reduce(e2e_pass(n_e2e), world, device, "max") * (args.steps / n_e2e)
    gc.enable()
    e2e_value = samples_all / (e2e_ms / 1e3)

    res = {
        "value": round(value, 3), "ms_per_step": round(total_ms / args.steps, 3),
        "config": {"workload": wl.name, "baseline_config": "configs[%d]" % wl.config_id,
                   "global_batch": wl.samples * world, "per_gpu_batch": wl.samples,
                   "bn": wl.bn_note, "weights": "random init, seed 0",
                   "geometry": getattr(wl, "geometry_note", "voxelization and rulebooks built inside every step"),
                   "l2": "flushed (256 MiB fill) between timed steps; per-step CUDA events summed (e2e: one event pair "
                         "around the K steps, flushes included)",
                   "dense_convs": "cuDNN fp32 (torch default allow_tf32), NHWC memory format", "parallelism": "dp%d" % world},
        "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(wl.h2d),
                "d2h_bytes_per_step": int(wl.d2h), "ms_per_step": round(e2e_ms / args.steps, 3)},
        "gpu_launches": int(launches), "clocks": clocks, "wall_s_timed_region": round(wall, 3),
        "step_ms_min_med_max": [round(min(step_ms), 3), round(sorted(step_ms)[len(step_ms) // 2], 3),
                                round(max(step_ms), 3)],
        "rank_ms_per_step_min_med_max": [round(rank_ms[0], 3), round(rank_ms[len(rank_ms) // 2], 3), round(rank_ms[-1], 3)],
        "step_ms": [round(v, 3) for v in step_ms],
    }
    if roof_all:
        res["roofline"] = _dominant(roof_all)
        res["roofline"]["peak_source"] = peaks["source"]
        res["roofline_all"] = roof_all
        res["roofline_pass_ms_per_step"] = round(sum(roof_step_ms) / max(1, len(roof_step_ms)), 3)
        res["own_kernels_ms_per_step"] = round(sum(e["ms_per_step"] for e in roof_all), 3)
    step_fn = getattr(wl, "step_fn", None)
    if step_fn is not None:
        terms = wl.step()                   # one more step: the loss terms the optimisation is at (finite = training is alive)
        res["loss_terms_after_timed_steps"] = {k: round(float(v), 6) for k, v in terms.items()}
        res["allreduce"] = _allreduce_report(step_fn, wl, world, device, args)
    return res


def _allreduce_report(step_fn, wl, world, device, args):
    """Gradient exchange of the training step: size, bucket count, the stand-alone cost of the same all-reduces on an idle
    GPU, and how much of it stays exposed after the backward (step with overlap vs step with the exchange skipped)."""
    import torch

    b = step_fn.bucket
    rep = {"gradient_bytes": int(b.numel * 4), "buckets": len(b.buckets), "overlapped_with_backward": bool(b.overlap),
           "normalisers": "one 14-float all-reduce, issued before the student forward, asynchronous"}
    if world == 1:
        rep["standalone_ms"] = 0.0
        return rep
    import torch.distributed as dist

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for it in range(3):
        _barrier(world)
        e0.record()
        works = [dist.all_reduce(b.flat[lo:hi], async_op=True) for lo, hi, _ in b.buckets]
        for w in works:
            w.wait()
        e1.record()
        torch.cuda.synchronize()
    rep["standalone_ms"] = round(_reduce(e0.elapsed_time(e1), world, device, "max"), 3)
    b.flat.zero_()
    # exposed cost: K steps without the all-reduces themselves (NOT a valid training step at N > 1: timing probe only)
    # against K steps with it, timed like pass 1 (per-step events, L2 flushed between steps, collector off: a generational
    # collection on one rank stalls every rank through the exchange and would be booked as "exposed")
    import gc

    flush = L2Flusher(device)

    def k_steps(warm):
        for _ in range(warm):
            wl.step()
        events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        torch.cuda.synchronize()
        _barrier(world)
        for a, c in events:
            flush()
            a.record()
            wl.step()
            c.record()
        _barrier(world)
        torch.cuda.synchronize()
        return _reduce(sum(a.elapsed_time(c) for a, c in events), world, device, "max") / args.steps

    gc.collect()
    gc.disable()
    b.exchange = False
    no_xchg = k_steps(2)
    b.exchange = True
    with_xchg = k_steps(1)
    gc.enable()
    rep["step_ms_without_exchange"] = round(no_xchg, 3)
    rep["step_ms_with_exchange"] = round(with_xchg, 3)
    rep["exposed_ms"] = round(with_xchg - no_xchg, 3)
    return rep


def run_ours(args):
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device; this arm has no CPU fallback (use --impl reference)")
    from unidistill_b200 import native

    native.load()
    torch.backends.cudnn.benchmark = os.environ.get("UD_CUDNN_BENCHMARK", "1") != "0"   # dense BEV/head convs: autotune
    world, rank, local = _dist_setup(args.gpus)
    device = torch.device("cuda", local if world > 1 else 0)
    peaks = _peaks()

    def build(name, **kw):
        return WORKLOADS[name](device, rank, teacher_mode=args.teacher_mode, **kw)

    wl = build(args.workload)
    res = measure(wl, args, world, rank, local, device, peaks)
    line = {"metric": METRIC, "value": res.pop("value"), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": res.pop("ms_per_step"), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    line.update(res)
    if args.profile:
        line["profiled"] = "run under a profiler range: numbers on this line are NOT bench values"
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
        if args.workload in ("lidar", "distill_lidar"):
            line["parity"] = full_size_parity(wl, device)
    del wl
    torch.cuda.empty_cache()

    # ---- secondary workloads on the same line: the other distillation direction by default, everything with --all
    secondary = []
    if not args.profile and not args.no_secondary:
        secondary = [w for w in (["distill", "distill_lidar"] if not args.all else
                                 ["distill", "distill_lidar", "lidar", "camera", "voxelize"]) if w != args.workload]
    sec = {}
    for name in secondary:
        w2 = build(name)
        r2 = measure(w2, args, world, rank, local, device, peaks, with_roofline=True)
        r2["unit"] = UNIT
        r2.pop("wall_s_timed_region", None)
        if args.teacher_mode == "eval" and name in ("distill", "distill_lidar") and args.all:
            pass
        sec[name] = r2
        del w2
        torch.cuda.empty_cache()
    if args.all and not args.profile:
        # quirk 9: the same primary step with the teacher's BatchNorms on batch statistics (pytorch-lightning train mode)
        other = "train" if args.teacher_mode == "eval" else "eval"
        w3 = WORKLOADS[args.workload](device, rank, teacher_mode=other) if args.workload.startswith("distill") else None
        if w3 is not None:
            r3 = measure(w3, args, world, rank, local, device, peaks, with_roofline=False)
            sec["%s[teacher_mode=%s]" % (args.workload, other)] = {k: r3[k] for k in ("value", "ms_per_step", "e2e")}
            del w3
            torch.cuda.empty_cache()
        # configs[4] roofline sweep: points per sample x BEV size (the primary line is 400k x 128^2)
        sweep = []
        for bev in (128, 180):
            for pps in (10000, 20000, 40000):
                if bev == 128 and pps == 40000 and args.workload == "distill_lidar":
                    continue
                w4 = WORKLOADS["distill_lidar"](device, rank, bev=bev, points_per_sweep=pps, teacher_mode=args.teacher_mode)
                r4 = measure(w4, args, world, rank, local, device, peaks, with_roofline=True)
                sweep.append({"points_per_sample": pps * 10, "bev": bev, "value": r4["value"], "ms_per_step": r4["ms_per_step"],
                              "e2e": r4["e2e"]["value"], "roofline": r4.get("roofline")})
                del w4
                torch.cuda.empty_cache()
        sec["configs[4]_sweep"] = sweep
    if sec:
        line["secondary"] = sec
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# CPU restatement of the reference path (oracle/), timed on the host cores.
def _dense_torch(feat, coords, batch_size, shape):
    """SparseConvTensor.dense() + HeightCompression view with torch CPU ops (differentiable)."""
    import torch

    D, H, W = [int(s) for s in shape]
    c = torch.as_tensor(coords).long()
    out = feat.new_zeros((batch_size, feat.shape[1], D, H, W))
    out[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feat
    return out.view(batch_size, feat.shape[1] * D, H, W)


def _cpu_lidar_encoder(enc, clouds, training):
    """LidarEncoder.forward on the CPU through oracle/: C first-come voxelizer + mean VFE, spconv-Native restated with
    torch CPU ops (gather -> mm -> index_add_ per kernel offset; autograd-capable), dense()."""
    from oracle import lidar as O

    v = enc.voxelizer
    _, coords, num, mean = O.voxelize([c.numpy() for c in clouds], v.voxel_size.tolist(), v.point_cloud_range.tolist(),
                                      v.grid_size.tolist(), v.max_num_points, v.max_voxels)
    sd = dict(enc.backbone_3d.named_parameters())
    sd.update(dict(enc.backbone_3d.named_buffers()))
    shape = [int(s) for s in enc.backbone_3d.sparse_shape]
    f, c, oshape = O.backbone_forward(sd, mean, coords, len(clouds), shape, training=training)
    return _dense_torch(f, c, len(clouds), oshape)


def _cpu_camera_encoder(enc, imgs, mats):
    """CameraEncoder.forward on the CPU: image nets + LSS geometry on torch CPU, then the reference's pooling path as
    written -- the lifted (6*112*16*44, 256) volume is materialised (lss_fpn.py:289-310) and scattered per BEV cell
    (voxel_pooling restated as an index_add_, differentiable)."""
    import torch

    lss = enc.backbone
    B = imgs.shape[0]
    src = lss.get_cam_feats(imgs[:, 0:1])[:, 0]
    ncam = src.shape[1]
    feat = lss._forward_depth_net(src.reshape(B * ncam, src.shape[2], src.shape[3], src.shape[4]), mats, 0)
    depth = feat[:, : lss.depth_channels].softmax(1)
    ctx = feat[:, lss.depth_channels: lss.depth_channels + lss.output_channels]
    geom = lss.get_geometry(mats["sensor2ego_mats"][:, 0], mats["intrin_mats"][:, 0], mats["ida_mats"][:, 0],
                            mats.get("bda_mat", None))
    idx = ((geom - (lss.voxel_coord - lss.voxel_size / 2.0)) / lss.voxel_size).int()
    nx, ny, nz = [int(v) for v in lss.voxel_num]
    vol = (depth.unsqueeze(1) * ctx.unsqueeze(2)).permute(0, 2, 3, 4, 1)              # (B*ncam, D, H, W, C)
    C = vol.shape[-1]
    vol = vol.reshape(B, -1, C)
    idx = idx.reshape(B, -1, 3).long()
    ok = (idx[..., 0] >= 0) & (idx[..., 0] < nx) & (idx[..., 1] >= 0) & (idx[..., 1] < ny) & (idx[..., 2] >= 0) & \
         (idx[..., 2] < nz)
    cell = torch.where(ok, idx[..., 1] * nx + idx[..., 0], torch.full_like(idx[..., 0], nx * ny))
    out = vol.new_zeros((B, nx * ny + 1, C))
    for b in range(B):
        out[b].index_add_(0, cell[b], vol[b])
    return out[:, : nx * ny].reshape(B, ny, nx, C).permute(0, 3, 1, 2).contiguous()


class _CpuModel:
    """A mirrored BEVFusionCenterHead on torch CPU whose LiDAR / camera encoders run through the oracle."""

    def __init__(self, modalities, train):
        from unidistill_b200 import models

        cfg = models.make_model_cfg(bev=128, **modalities)
        self.cfg = cfg
        self.model = models.BEVFusionCenterHead(cfg)
        self.model.train(train)
        self.train = train

    def bev(self, clouds, imgs, mats):
        m = self.model
        lid = _cpu_lidar_encoder(m.lidar_encoder, clouds, self.train) if m.lidar_encoder is not None else None
        cam = _cpu_camera_encoder(m.camera_encoder, imgs, mats) if m.camera_encoder is not None else None
        if lid is not None and cam is not None:
            return m.fusion_encoder(lid, cam)
        return lid if lid is not None else cam


class CpuForward:
    """Port of a forward workload for ONE sample per step (bs=1): encoders through oracle/, dense part = the mirrored
    BaseBEVBackbone + CenterHead modules on torch CPU (the reference's own Python for that part)."""

    def __init__(self, modalities, seed=0):
        import torch

        from unidistill_b200 import synthetic as S

        torch.manual_seed(0)
        self.m = _CpuModel(modalities, True)
        self.lidar = modalities["lidar"]
        lid = S.lidar_batch(BS_PER_GPU, seed=seed) if self.lidar else S.lidar_batch(BS_PER_GPU, seed=seed,
                                                                                     points_per_sweep=64, sweeps=1)
        cam = S.camera_batch(BS_PER_GPU, seed=seed) if modalities["camera"] else None
        self.clouds = [lid["points"][b] for b in range(BS_PER_GPU)]
        self.gt = _gt_with_labels(lid)
        self.cam = cam
        self.i = 0
        self.sample = "1 sample per step (bs=1), same generator and weights as the GPU arm"

    def step(self):
        import torch

        b = self.i % BS_PER_GPU
        self.i += 1
        with torch.no_grad():
            imgs = mats = None
            if self.cam is not None:
                imgs = self.cam["imgs"][b:b + 1]
                mats = {k: v[b:b + 1] for k, v in self.cam["mats_dict"].items()}
            bev = self.m.bev([self.clouds[b]], imgs, mats)
            x = self.m.model.bev_encoder(bev)
            return self.m.model.det_head(x[0], self.gt[b:b + 1])

    samples = 1


class CpuDistillStep:
    """Port of one distillation optimisation step for ONE sample (bs=1) on the host cores: teacher forward (no grad) +
    student forward + detection loss + the three distillation losses (oracle/distill.py) + backward + clip + AdamW, the
    encoders through oracle/ (spconv-Native / voxel_pooling restated with torch CPU ops under autograd)."""

    def __init__(self, teacher_modalities, student_modalities, preset, seed=0):
        import torch

        from unidistill_b200 import synthetic as S
        from unidistill_b200.step import PRESETS

        torch.manual_seed(0)
        self.teacher = _CpuModel(teacher_modalities, False)
        self.student = _CpuModel(student_modalities, True)
        for p in self.teacher.model.parameters():
            p.requires_grad = False
        self.teacher.model.det_head.dense_head.distill = True
        self.cfg = PRESETS[preset]
        self.opt = torch.optim.AdamW(self.student.model.parameters(), lr=self.cfg["lr"], weight_decay=1e-7)
        lid = S.lidar_batch(BS_PER_GPU, seed=seed)
        cam = S.camera_batch(BS_PER_GPU, seed=seed)
        self.clouds = [lid["points"][b] for b in range(BS_PER_GPU)]
        self.gt_boxes = lid["gt_boxes"]
        self.gt = _gt_with_labels(lid)
        self.cam = cam
        self.i = 0
        self.sample = ("1 sample per step (bs=1): teacher fwd + student fwd + det + 3 distill losses + bwd + clip + AdamW, "
                       "same generator and weights as the GPU arm")

    def step(self):
        import torch

        from oracle import distill as OD

        b = self.i % BS_PER_GPU
        self.i += 1
        g = self.student.cfg["geometry"]
        imgs = self.cam["imgs"][b:b + 1]
        mats = {k: v[b:b + 1] for k, v in self.cam["mats_dict"].items()}
        gt = self.gt[b:b + 1]
        sm, tm = self.student.model, self.teacher.model
        self.opt.zero_grad(set_to_none=True)
        bev_s = self.student.bev([self.clouds[b]], imgs, mats)
        xs = sm.bev_encoder(bev_s)
        ret = sm.det_head(xs[0], gt)
        loss_det, _ = sm.det_head.dense_head.get_loss(ret)
        with torch.no_grad():
            bev_t = self.teacher.bev([self.clouds[b]], imgs, mats)
            xt = tm.bev_encoder(bev_t)
            ret_t = tm.det_head(xt[0], gt)
        valid = OD.valid_prefix(self.gt_boxes[b:b + 1])
        coords = OD.bev_corners(gt, g["pc_range"], g["voxel_size"], g["out_size_factor"])
        lf = OD.feature_distill_loss(bev_s, bev_t, coords, valid)
        lr = OD.bev_distill_loss(xs[0], xt[0], coords, valid)
        lc, lg = OD.response_distill_loss(ret["multi_head_features"], ret_t["multi_head_features"], gt, g["pc_range"],
                                          g["voxel_size"], g["out_size_factor"], clamp_min=self.cfg["clamp_min"])
        wf, wr, wp = self.cfg["weights"]
        loss = loss_det.mean() + wf * lf + wr * lr + wp * (lc + lg)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(sm.parameters(), 0.1)
        self.opt.step()
        return loss.detach()

    samples = 1


def _cpu_workload(name):
    if name == "lidar":
        return CpuForward(dict(lidar=True, camera=False))
    if name == "camera":
        return CpuForward(dict(lidar=False, camera=True))
    if name == "distill":
        return CpuDistillStep(DistillTrainStep.teacher_modalities, DistillTrainStep.student_modalities,
                              DistillTrainStep.preset)
    if name == "distill_lidar":
        return CpuDistillStep(DistillLidarTrainStep.teacher_modalities, DistillLidarTrainStep.student_modalities,
                              DistillLidarTrainStep.preset)
    if name == "voxelize":
        return CpuVoxelizeScatter()
    raise KeyError(name)


class CpuVoxelizeScatter:
    """configs[0] on the CPU: the C first-come voxelizer (that IS the algorithm of the reference's CPU path) on the fine and
    the pillar grid + dense()."""

    sample = "1 cloud of 20k points per step, bs=1"
    samples = 1

    def __init__(self, seed=0):
        from unidistill_b200 import synthetic as S

        self.pts = S.lidar_sample(seed, points_per_sweep=20000, sweeps=1)

    def step(self):
        from oracle import lidar as O
        from unidistill_b200 import synthetic as S

        O.voxelize([self.pts], S.VOXEL_SIZE_128, S.PC_RANGE_128, S.GRID_SIZE_128, 10, 120000)
        _, c, _, mean = O.voxelize([self.pts], [0.8, 0.8, 1.6], S.PC_RANGE_128, [128, 128, 5], 32, 120000)
        return O.height_compression(mean, c, 1, [5, 128, 128])


def cpu_baseline(workload, budget_s=25.0):
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    wl = _cpu_workload(workload)
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


def full_size_parity(wl, device):
    """Full-size GPU-vs-CPU check on one of the bench's own clouds (bs=1): a FRESH LiDAR encoder (seed 0, the weights both
    arms start from; eval-mode BatchNorm with non-trivial running statistics, so that batch statistics do not amplify 1-ulp
    differences): voxel coordinates and counts bit-exact, BEV map of voxelizer + 21-layer sparse encoder + scatter within
    1e-3 relative of the CPU port's."""
    import numpy as np
    import torch

    from oracle import lidar as O
    from unidistill_b200 import models

    model = getattr(wl, "model", None)
    if model is None:
        model = wl.step_fn.student
    torch.manual_seed(0)
    enc = models.LidarEncoder(model.cfg.lidar_encoder).to(device).eval()
    g = torch.Generator(device="cpu").manual_seed(1)
    for m in enc.modules():
        if isinstance(m, torch.nn.BatchNorm1d):
            m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.running_var.shape, generator=g) + 0.5)
    cloud = (wl.points_dev if hasattr(wl, "points_dev") else wl.dev["points"])[0].contiguous()
    with torch.no_grad():
        v = enc.voxelizer
        _, coords, num = v([cloud])
        bev = enc([cloud])
    torch.cuda.synchronize()
    _, rc, rn, rmean = O.voxelize([cloud.cpu().numpy()], v.voxel_size.tolist(), v.point_cloud_range.tolist(),
                                  v.grid_size.tolist(), v.max_num_points, v.max_voxels)
    sd = {k: t.detach().cpu() for k, t in enc.backbone_3d.state_dict().items()}
    f, c, shape = O.backbone_forward(sd, rmean, rc, 1, [int(s) for s in enc.backbone_3d.sparse_shape], training=False)
    ref = O.height_compression(f.numpy(), c, 1, shape)
    got = bev.float().cpu().numpy()
    rel = float(np.abs(got - ref).max() / max(1e-12, np.abs(ref).max()))
    return {"what": "bench cloud 0 (bs=1, fresh seed-0 encoder, eval-mode BN): GPU voxelizer + 21-layer sparse encoder + "
                    "scatter vs the CPU port", "voxels": int(rc.shape[0]),
            "voxel_coords_bit_exact": bool(np.array_equal(coords.cpu().numpy(), rc)),
            "voxel_counts_bit_exact": bool(np.array_equal(num.cpu().numpy(), rn)),
            "bev_rel_err": rel, "tol": 1e-3, "ok": bool(rel <= 1e-3)}


def run_reference(args):
    """The reference's CPU implementation of the path (spconv / voxel_pooling_ext are binary wheels that can
    neither be imported nor compiled here, so this is the oracle port), all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    wl = _cpu_workload(args.workload)
    warm = max(1, min(args.warmup, 1 if args.workload.startswith("distill") else 2))   # bounded: one CPU step is seconds long
    for _ in range(warm):
        wl.step()
    budget = float(os.environ.get("UD_REFERENCE_BUDGET_S", "150"))
    t0 = time.perf_counter()
    done = 0
    for _ in range(args.steps):
        wl.step()
        done += 1
        if time.perf_counter() - t0 > budget and done >= 2:
            break                                                        # bounded sample: stop after the time budget
    el = time.perf_counter() - t0
    value = done * wl.samples / el
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": done, "warmup": warm,
        "ms_per_step": round(el / done * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": getattr(WORKLOADS[args.workload], "title", args.workload) + " / configs[%d]" %
                   WORKLOADS[args.workload].config_id, "sample_per_step": wl.sample,
                   "steps_requested": args.steps, "time_budget_s": budget},
        "cpu_baseline": {"value": round(value, 4), "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": "%d x [%s]" % (done, wl.sample)},
        "e2e": {"value": round(value, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="distill_lidar", choices=sorted(WORKLOADS))
    ap.add_argument("--teacher-mode", default="eval", choices=["eval", "train"])
    ap.add_argument("--all", action="store_true", help="also run configs[0..2], the other teacher mode and the configs[4] sweep")
    ap.add_argument("--no-secondary", action="store_true")
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
