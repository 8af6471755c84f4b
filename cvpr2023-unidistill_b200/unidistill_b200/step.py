"""One optimisation step of a cross-modality distillation experiment.

Restates `Exp.training_step` of the reference's distill scripts
(exps/multisensor_fusion/nuscenes/BEVFusion/BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar.py:438-513 and
its three siblings) plus what pytorch-lightning's DDP trainer does around it (exps/base_cli.py:43-44,
BEVFusion_nuscenes_base_exp.py:436-441):

    target assignment (GT only; ONE for student and teacher)  -> one fused, asynchronous normaliser all-reduce
    student forward (train) -> detection loss                        (:463-465)
    teacher forward (frozen, `distill=True`, return_feature=True)    (:466-469)
    GT boxes -> BEV corners in cell units, valid-box prefix          (:447-455, :470-483)
    loss = det + w_f * feature + w_r * relation + w_p * (resp_cls + resp_reg)                     (:484-510)
    backward with the gradient all-reduce overlapped bucket by bucket (NCCL over NVLink, DDP's job in the reference)
    gradient-norm clip 0.1 (base_cli.py:43, `gradient_clip_val`), AdamW(weight_decay=1e-7), MultiStepLR([10, 15])

What changes operationally (same numbers): no `.cpu()` / numpy hop for the corners, no Python loops over the batch,
all GT-only normalisers (6 x num_pos, 6 x num_hm, #valid boxes, Gaussian-mask sum) travel in ONE 14-float all-reduce
issued BEFORE the student forward (they depend on the ground truth only) instead of ~27 blocking scalar
`reduce_mean`s inside it, and the teacher's buffers are restored from a device-side snapshot instead of
`load_state_dict` from host memory every step (:465).
"""
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from unidistill.layers.losses import distill as L
from unidistill.utils import torch_dist

from . import distill as D

# (feature, relation, response) loss weights, clamp of the teacher's heat-map sigmoid and learning rate of the reference's
# four distillation experiments (same directory, `training_step` / `__init__` of each script)
PRESETS = {
    "camera_exp_distill_lidar": dict(weights=(100.0, 40.0, 10.0), clamp_min=1e-4, lr=2e-4),
    "camera_exp_distill_fusion": dict(weights=(10.0, 5.0, 10.0), clamp_min=1e-3, lr=2e-4),
    "lidar_exp_distill_camera": dict(weights=(10.0, 5.0, 1.0), clamp_min=1e-4, lr=1e-3),
    "lidar_exp_distill_fusion": dict(weights=(10.0, 1.0, 10.0), clamp_min=1e-4, lr=1e-3),
}
DISTILL_WEIGHTS = PRESETS["camera_exp_distill_lidar"]["weights"]      # kept for callers of round 1


class FlatGradBucket:
    """All student gradients live in one flat fp32 buffer (param.grad are views), cut into contiguous buckets in
    REVERSE registration order -- the order backward produces them.  With `overlap`, a bucket's all-reduce is issued
    (asynchronously, on the process group's own stream) from a post-accumulate hook as soon as its last gradient has
    landed, so the exchange runs under the rest of the backward pass; `finish()` waits for the stragglers.  Works with
    any backend (gloo in the CPU tests).

    `steal` (default): during the backward `param.grad` is None, so autograd's AccumulateGrad keeps the incoming gradient
    tensor instead of launching `grad += incoming` per parameter (~250 small kernels per step here); when a bucket is
    complete its gradients are copied into the flat buffer with ONE `_foreach_copy_`, and `finish()` points every
    `param.grad` at its view of the (reduced) buffer again for clipping and the optimizer."""

    def __init__(self, params, bucket_bytes=16 << 20, overlap=True, steal=True):
        self.steal = bool(steal)
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.numel = n
        self.overlap = overlap
        self.buckets = []            # (lo, hi) element ranges of self.flat, in expected completion order
        self._bucket_of = {}
        self._view = {}              # id(param) -> its view of the flat buffer
        self._members = []           # per bucket: its parameters
        self._pending = []
        self._works = []
        self._hooks = []
        self.enabled = True
        self.exchange = True         # False: everything but the all-reduce itself (bench.py's exposed-cost probe)
        # lay the parameters out back to front so that bucket 0 (completed first) holds the LAST layers
        off = 0
        lo, cnt = 0, 0
        for p in reversed(self.params):
            # same strides as the parameter (NHWC-format conv weights stay NHWC): dense, so the segment is exact
            view = self.flat[off: off + p.numel()].as_strided(p.size(), p.stride())
            self._view[id(p)] = view
            p.grad = None if self.steal else view
            off += p.numel()
            self._bucket_of[id(p)] = len(self.buckets)
            if len(self._members) == len(self.buckets):
                self._members.append([])
            self._members[-1].append(p)
            cnt += 1
            if (off - lo) * 4 >= bucket_bytes:
                self.buckets.append((lo, off, cnt))
                lo, cnt = off, 0
        if off > lo or not self.buckets:
            self.buckets.append((lo, off, cnt))
        while len(self._members) < len(self.buckets):
            self._members.append([])
        if overlap:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self._reset()

    def _reset(self):
        self._pending = [c for _, _, c in self.buckets]
        self._works = []
        self._fired = [False] * len(self.buckets)

    def zero(self):
        self.flat.zero_()
        if self.steal:
            for p in self.params:
                p.grad = None
        self._reset()

    def _collect(self, b):
        """steal mode: the bucket's gradients (wherever autograd left them) -> their views of the flat buffer"""
        dst, src = [], []
        for p in self._members[b]:
            g = p.grad
            view = self._view[id(p)]
            if g is not None and g.data_ptr() != view.data_ptr():
                dst.append(view)
                src.append(g)
        if dst:
            torch._foreach_copy_(dst, src)

    def _fire(self, b):
        if self._fired[b]:
            return
        self._fired[b] = True
        if self.steal:
            self._collect(b)
        lo, hi, _ = self.buckets[b]
        if self.exchange and torch_dist.get_world_size() > 1 and hi > lo:
            self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))

    def _on_grad(self, p):
        if not self.enabled:
            return
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._fire(b)

    def finish(self):
        """Issues the all-reduce of every bucket that has not fired (parameters without a gradient this step, or
        overlap off), waits for all of them and returns the flat buffer holding the SUM over ranks."""
        for b in range(len(self.buckets)):
            self._fire(b)
        for w in self._works:
            w.wait()
        self._works = []
        if self.steal:
            for p in self.params:
                p.grad = self._view[id(p)]
        return self.flat

    def all_reduce_mean(self):
        world = torch_dist.get_world_size()
        self.finish()
        if world > 1:
            self.flat.mul_(1.0 / world)
        return self.flat


def clip_and_average_(flat, world, max_norm):
    """flat holds the gradient SUM over `world` ranks.  In place: average, then scale so that the global L2 norm is at
    most `max_norm` (torch.nn.utils.clip_grad_norm_ semantics, coefficient max_norm / (norm + 1e-6) clamped to 1; what
    pytorch-lightning's gradient_clip_val does after DDP's averaging).  One reduction + one scaling pass over the
    bucket, no host synchronisation.  Returns the (pre-clip) norm as a 0-dim tensor."""
    norm = torch.linalg.vector_norm(flat) / float(world)
    if max_norm is None:
        if world > 1:
            flat.mul_(1.0 / world)
        return norm
    coef = torch.clamp(max_norm / (norm + 1e-6), max=1.0) / float(world)
    flat.mul_(coef)
    return norm


def fold_frozen_batchnorm(module):
    """Folds every eval-mode BatchNorm2d that directly follows a Conv2d / ConvTranspose2d into that convolution
    (w' = w * gamma / sqrt(var + eps), b' = beta + (b - mean) * gamma / sqrt(var + eps)) and replaces the norm by
    nn.Identity.  For a FROZEN network only (the distillation teacher, reference ..._distill_lidar.py:419-433: eval(),
    requires_grad False): same function up to fp32 rounding, one full pass over every activation less.  Handles
    nn.Sequential runs and the convN / bnN attribute pairs of torchvision's ResNet blocks.  Returns the count."""
    from torch.nn.utils.fusion import fuse_conv_bn_eval

    def fusable(conv, bn):
        return (isinstance(conv, (nn.Conv2d, nn.ConvTranspose2d)) and type(bn) is nn.BatchNorm2d and not bn.training
                and bn.track_running_stats and bn.running_mean is not None
                and (not isinstance(conv, nn.ConvTranspose2d) or conv.groups == 1))

    n = 0
    for m in list(module.modules()):
        names = list(m._modules.keys())
        if isinstance(m, nn.Sequential):
            pairs = [(names[i], names[i + 1]) for i in range(len(names) - 1)]
        else:
            pairs = [(a, "bn" + a[4:]) for a in names if a.startswith("conv") and ("bn" + a[4:]) in m._modules]
        for a, b in pairs:
            conv, bn = m._modules[a], m._modules[b]
            if conv is None or bn is None or not fusable(conv, bn):
                continue
            fmt = torch.channels_last if conv.weight.is_contiguous(memory_format=torch.channels_last) and \
                not conv.weight.is_contiguous() else torch.contiguous_format
            fused = fuse_conv_bn_eval(conv.eval(), bn, transpose=isinstance(conv, nn.ConvTranspose2d))
            fused = fused.to(memory_format=fmt)
            for p in fused.parameters():
                p.requires_grad = False
            m._modules[a] = fused
            m._modules[b] = nn.Identity()
            n += 1
    return n


def gt_with_labels(gt_boxes, gt_labels):
    """(B,M,9) boxes + 0-based labels -> (B,M,10) with 1-based labels in the last column (:456-458)."""
    return torch.cat([gt_boxes, (gt_labels.to(gt_boxes.dtype) + 1).unsqueeze(2)], dim=2)


def normaliser_vector(targets, valid, mask_sum):
    """[num_pos (T), num_hm (T), #valid boxes, Gaussian-mask sum] of this rank as one (2T + 2,) vector."""
    tasks = sorted(targets["mask"].keys()) if isinstance(targets["mask"], dict) else list(range(len(targets["mask"])))
    parts = [targets["mask"][t].float().sum() for t in tasks]
    parts += [targets["heatmap"][t].eq(1).float().sum() for t in tasks]
    parts += [valid.float().sum(), mask_sum]
    return torch.stack([p.reshape(()) for p in parts]), len(tasks)


class GraphedTargets:
    """Everything of a step that depends on the ground truth only -- label column, valid-box prefix, BEV corners, the
    head's target assignment, the Gaussian response mask and this rank's normaliser vector -- is ~150 launches of a few
    microseconds with static shapes: captured once per ground-truth shape and replayed.  Outputs are static buffers,
    valid until the next call."""

    def __init__(self, step):
        self.step = step
        self.key = None
        self.graph = None

    def _run(self, gt_boxes, gt_labels):
        st = self.step
        gt = gt_with_labels(gt_boxes, gt_labels)
        valid = D.valid_box_prefix(gt_boxes)
        corners = D.box_corners_bev(gt, st.pc_range, st.voxel_size, st.osf)
        head = st.student.det_head.dense_head
        targets = head.assign_targets(gt)
        for _, enc in targets["box_encoding"].items():
            enc.masked_fill_(torch.isinf(enc), 0)
        h, w = targets["heatmap"][0].shape[-2:]
        mask = D.response_mask(gt, (h, w), st.pc_range, st.voxel_size, st.osf)
        vec, n_tasks = normaliser_vector(targets, valid, mask.sum())
        return gt, valid, corners, targets, vec, n_tasks

    def __call__(self, gt_boxes, gt_labels):
        key = (tuple(gt_boxes.shape), gt_boxes.dtype, tuple(gt_labels.shape), gt_labels.dtype, gt_boxes.device)
        if self.graph is None or key != self.key:
            self.static_in = (gt_boxes.detach().clone(), gt_labels.detach().clone())
            side = torch.cuda.Stream(device=gt_boxes.device)
            side.wait_stream(torch.cuda.current_stream(gt_boxes.device))
            with torch.cuda.stream(side), torch.no_grad():
                for _ in range(3):
                    self._run(*self.static_in)
            torch.cuda.current_stream(gt_boxes.device).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph), torch.no_grad():
                self.static_out = self._run(*self.static_in)
            self.key = key
        self.static_in[0].copy_(gt_boxes)
        self.static_in[1].copy_(gt_labels)
        self.graph.replay()
        return self.static_out


class AsyncNormalisers:
    """mean over ranks of [num_pos (T), num_hm (T), #valid boxes, Gaussian-mask sum] in ONE all-reduce that is in flight
    while the student forward runs; `get()` (first consumer: get_loss, after the whole forward) waits for it."""

    def __init__(self, targets, valid, mask_sum, vec=None, n_tasks=None):
        """`vec`, `n_tasks`: a ready normaliser_vector (GraphedTargets builds it inside its graph)."""
        if vec is None:
            vec, n_tasks = normaliser_vector(targets, valid, mask_sum)
        self.T = n_tasks
        self.vec = vec
        self.world = torch_dist.get_world_size()
        self.work = dist.all_reduce(self.vec, op=dist.ReduceOp.SUM, async_op=True) if self.world > 1 else None

    def get(self):
        if self.work is not None:
            self.work.wait()
            self.work = None
            self.vec = self.vec / float(self.world)
        T, v = self.T, self.vec
        return v[:T], v[T: 2 * T], v[2 * T], v[2 * T + 1]


def fused_normalisers(ret, valid, mask_sum):
    """Blocking form (kept for callers of round 1): -> (num_pos (T), num_hm (T), n_valid, mask_sum), rank-averaged."""
    return AsyncNormalisers(ret, valid, mask_sum).get()


class DistillStep:
    """student / teacher: BEVFusionCenterHead models on the same device.  `geometry`: dict(pc_range, voxel_size,
    out_size_factor).  Call with a device-resident batch dict (points list | None, imgs | None, mats_dict | None,
    gt_boxes (B,M,9), gt_labels (B,M)); returns a dict of detached loss terms.

    preset: one of PRESETS (weights / clamp_min / lr of the reference's four experiments); explicit arguments win.
    teacher_mode: "eval" -- frozen teacher in eval mode (what freeze_model + .eval() at :419-433 ask for), BatchNorm2d
    folded into its convolutions; "train" -- the teacher's BatchNorms use batch statistics, which is what happens in the
    reference when pytorch-lightning puts the whole LightningModule into train mode before the step (the
    reference reloads the teacher's state_dict every step to undo the running-statistics updates, :465): no folding,
    buffers restored from a device-side snapshot after the forward."""

    def __init__(self, student, teacher, geometry, lr=None, weight_decay=1e-7, clamp_min=None, optimizer=True,
                 fold_teacher_bn=True, teacher_cuda_graph=True, weights=None, preset="camera_exp_distill_lidar",
                 grad_clip=0.1, lr_milestones=(10, 15), lr_gamma=0.1, teacher_mode="eval", overlap_allreduce=True,
                 bucket_bytes=16 << 20, fuse_teacher_epilogues=True,
                 student_cuda_graph=True, overlap_teacher_camera=False):
        cfg = PRESETS[preset]
        self.weights = tuple(float(w) for w in (weights if weights is not None else cfg["weights"]))
        self.clamp_min = float(clamp_min if clamp_min is not None else cfg["clamp_min"])
        lr = float(lr if lr is not None else cfg["lr"])
        assert teacher_mode in ("eval", "train")
        self.teacher_mode = teacher_mode
        self._geo_stream = None
        self._geo_events = None
        self._geo_turn = 0
        self._prefetched = None
        self._geo_in_use = None
        self._geo_prev = None
        self.student, self.teacher = student, teacher
        self.pc_range = list(geometry["pc_range"])
        self.voxel_size = list(geometry["voxel_size"])
        self.osf = int(geometry["out_size_factor"])
        self.grad_clip = grad_clip
        self.keep_features = False          # tests: keep the student/teacher tensors of the last step in self.last
        self.last = None
        self.grad_norm = None
        for p in teacher.parameters():                         # freeze_model (:427-433)
            p.requires_grad = False
        teacher.det_head.dense_head.distill = True             # (:419)
        self.folded_bn = 0
        self.fused_epilogues = 0
        if teacher_mode == "eval":
            teacher.eval()
            # the frozen eval-mode teacher: BatchNorm folded into the dense convolutions
            self.folded_bn = fold_frozen_batchnorm(teacher) if fold_teacher_bn else 0
            if self.folded_bn and fuse_teacher_epilogues:
                from .fused_dense import fuse_frozen_conv_relu

                # bias (+ residual) + ReLU of the folded convolutions in cuDNN's epilogue (the towers of the head are
                # batched by CenterHead.run_towers, which fuses its own first layer)
                self.fused_epilogues = fuse_frozen_conv_relu(teacher, skip=("SepHead",))
        else:
            teacher.train()
        if teacher_cuda_graph and next(teacher.parameters()).is_cuda:
            teacher.use_cuda_graph(True)                       # dense trunk replayed from a CUDA graph
        self._teacher_buffers = [(b, b.detach().clone()) for b in teacher.buffers()]
        student.train()
        self.overlap_teacher_camera = bool(overlap_teacher_camera)
        self._graphed_targets = None
        if student_cuda_graph and next(student.parameters()).is_cuda:
            student.use_cuda_graph_training(True)              # dense trunk + loss, forward AND backward, from CUDA graphs
            self._graphed_targets = GraphedTargets(self)       # and the ground-truth-only preparation
        self.bucket = FlatGradBucket(student.parameters(), bucket_bytes=bucket_bytes, overlap=overlap_allreduce,
                                     steal=os.environ.get("UD_GRAD_STEAL", "1") != "0")
        self.opt = self.sched = None
        if optimizer:
            fused = self.bucket.flat.is_cuda
            self.opt = torch.optim.AdamW(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=fused)
            self.sched = torch.optim.lr_scheduler.MultiStepLR(self.opt, list(lr_milestones), gamma=lr_gamma)

    def epoch_end(self):
        """MultiStepLR([10, 15]) steps once per epoch (pytorch-lightning's default scheduler interval)."""
        if self.sched is not None:
            self.sched.step()

    def _restore_teacher(self):
        """The reference reloads the teacher's state_dict every step (:465) because pytorch-lightning may have put it
        back into train mode; an eval-mode teacher cannot change its buffers, so there is nothing to restore."""
        if self.teacher_mode == "eval" and not any(m.training for m in self.teacher.modules()):
            return
        for b, snap in self._teacher_buffers:
            b.copy_(snap)

    def losses(self, batch):
        points, imgs, metas = batch.get("points"), batch.get("imgs"), batch.get("mats_dict")
        # ---- everything that depends on the ground truth only: ONE target assignment (the reference runs the same
        # assignment twice, in the student's and the teacher's head) and the normalisers' all-reduce, in flight from here
        if self._graphed_targets is not None and batch["gt_boxes"].is_cuda:
            gt, valid, corners, targets, vec, n_tasks = self._graphed_targets(batch["gt_boxes"], batch["gt_labels"])
            pending = AsyncNormalisers(None, None, None, vec=vec, n_tasks=n_tasks)
            persistent = True
        else:
            gt = gt_with_labels(batch["gt_boxes"], batch["gt_labels"])
            valid = D.valid_box_prefix(batch["gt_boxes"])                               # (:447-455)
            corners = D.box_corners_bev(gt, self.pc_range, self.voxel_size, self.osf)   # (:470-483)
            head = self.student.det_head.dense_head
            with torch.no_grad():
                targets = head.assign_targets(gt)
                for _, enc in targets["box_encoding"].items():
                    enc.masked_fill_(torch.isinf(enc), 0)                           # log(0) of padded boxes (quirk 10)
                h, w = targets["heatmap"][0].shape[-2:]
                mask = D.response_mask(gt, (h, w), self.pc_range, self.voxel_size, self.osf)
                pending = AsyncNormalisers(targets, valid, mask.sum())
            persistent = False
        norm = {}

        def normalisers(ret):
            num_pos, num_hm, n_valid, mask_sum = pending.get()
            ret["num_pos_reduced"], ret["num_hm_reduced"] = num_pos, num_hm
            norm["n_valid"], norm["mask_sum"] = n_valid, mask_sum

        geometry = {}       # voxels + rulebooks of the step's clouds, shared by the two LiDAR encoders (if there are two)
        pf, self._prefetched = self._prefetched, None
        if pf is not None and pf[0] is points:          # built on the side stream under the previous step's backward
            torch.cuda.current_stream().wait_event(pf[2])
            geometry = pf[1]
            self._geo_in_use = geometry                 # kept alive until the side stream may reuse its memory
        cam_future = None
        if (self.overlap_teacher_camera and imgs is not None and imgs.is_cuda and self.teacher.with_camera_encoder
                and self.teacher.with_lidar_encoder):
            # the frozen teacher's camera branch on a side stream under the LiDAR encoders (which stall on host reads).
            # Off by default: measured on B200 at bs = 4 it changes nothing (59.16 ms per step either way) -- the bubbles it
            # fills are paid back by the kernels it shares the SMs with.
            cam_future = self.teacher.start_camera(imgs, metas)
        ret_dict, tb, feat_s, bev_s, resp_s, _ = self.student(points, imgs, metas, gt, normalisers=normalisers,
                                                              targets=targets, geometry=geometry,
                                                              persistent_targets=persistent)
        with torch.no_grad():
            feat_t, bev_t, resp_t = self.teacher(points, imgs, metas, gt, return_feature=True, targets=targets,
                                                 geometry=geometry, camera_future=cam_future)
            self._restore_teacher()
        loss_feature = L.FeatureDistillLoss(feat_s, feat_t, corners, valid, weight=norm["n_valid"])
        loss_rel = L.BEVDistillLoss(bev_s, bev_t, corners, valid, weight=norm["n_valid"])
        loss_cls, loss_reg = L.ResponseDistillLoss(resp_s, resp_t, gt, self.pc_range, self.voxel_size, self.osf,
                                                   clamp_min=self.clamp_min, mask_weight=norm["mask_sum"])
        wf, wr, wp = self.weights
        loss = ret_dict["loss"].mean() + wf * loss_feature + wr * loss_rel + wp * (loss_cls + loss_reg)
        terms = {"loss": loss, "loss_det": ret_dict["loss"].detach(), "loss_feature": loss_feature.detach(),
                 "loss_bev_rel": loss_rel.detach(), "loss_resp_cls": loss_cls.detach(),
                 "loss_resp_reg": loss_reg.detach()}
        if self.keep_features:
            self.last = dict(feat_s=feat_s, bev_s=bev_s, resp_s=resp_s, feat_t=feat_t, bev_t=bev_t, resp_t=resp_t,
                             gt=gt, valid=valid, corners=corners, norm=dict(norm))
        return terms

    def geometry_stream(self):
        if self._geo_stream is None:
            self._geo_stream = torch.cuda.Stream()
        return self._geo_stream

    def prefetch_geometry(self, points, ready_event=None):
        """Voxelization + rulebooks of the NEXT step's clouds on a side stream.  Call it right after `step(batch)` returned
        (the host has launched the whole step, the GPU is still in its backward): the five host reads of the geometry
        phase then wait on the side stream only, and the next step finds its geometry ready (`points` must be the very
        object the next batch carries; `ready_event`: recorded where the clouds were produced, if not on this stream)."""
        enc = None
        for net in (self.student, self.teacher):
            if getattr(net, "with_lidar_encoder", False):
                enc = net.lidar_encoder
                break
        if enc is None or not points or not points[0].is_cuda:
            return
        main = torch.cuda.current_stream()
        side = self.geometry_stream()
        if self._geo_events is None:         # reusable events: creating them per step inside a timed region contends with
            self._geo_events = [torch.cuda.Event() for _ in range(4)]        # NVML clock sampling for a driver lock
        self._geo_turn ^= 1
        done, step_end = self._geo_events[self._geo_turn], self._geo_events[2 + self._geo_turn]
        # Memory protocol (no record_stream: tensors parked by it made the allocator grow for several steps, each growth a
        # device-wide stall).  The geometry tensors live in the side stream's pool but are read on the main stream, so they
        # may only return to the pool once the main stream is past their step: the geometry consumed by the step BEFORE
        # the one just launched is released here, after the side stream has been told to wait for that step's end.
        if self._geo_prev is not None:
            side.wait_event(self._geo_prev[1])
            self._geo_prev = None
        step_end.record(main)                # end of the step that has just been launched (it consumed _geo_in_use)
        if self._geo_in_use is not None:
            self._geo_prev, self._geo_in_use = (self._geo_in_use, step_end), None
        if ready_event is not None:
            side.wait_event(ready_event)
        geometry = {}
        with torch.cuda.stream(side):
            enc.build_geometry(points, geometry)
            done.record(side)
        self._prefetched = (points, geometry, done)

    def __call__(self, batch):
        self.bucket.zero()
        terms = self.losses(batch)
        terms["loss"].backward()                 # bucket all-reduces fire from the gradient hooks, under the backward
        flat = self.bucket.finish()
        self.grad_norm = clip_and_average_(flat, torch_dist.get_world_size(), self.grad_clip)
        if self.opt is not None:
            self.opt.step()
        terms["loss"] = terms["loss"].detach()
        terms["grad_norm"] = self.grad_norm
        return terms
