"""One optimisation step of a cross-modality distillation experiment.

Restates `Exp.training_step` of the reference's distill scripts
(exps/multisensor_fusion/nuscenes/BEVFusion/BEVFusion_nuscenes_centerhead_camera_exp_distill_lidar.py:438-513 and
its three siblings) plus the optimiser/all-reduce work pytorch-lightning's DDP does around it:

    student forward (train) -> detection loss                        (:463-465)
    teacher forward (frozen, `distill=True` so it still assigns targets, return_feature=True)   (:466-469)
    GT boxes -> BEV corners in cell units, valid-box prefix          (:447-455, :470-483)
    loss = det + 100 * feature + 40 * relation + 10 * (resp_cls + resp_reg)                       (:484-510)
    backward, ONE gradient all-reduce over NCCL (flat bucket), AdamW

What changes operationally (same numbers): no `.cpu()` / numpy hop for the corners, no Python loops over the batch,
all GT-only normalisers (6 x num_pos, 6 x num_hm, #valid boxes, Gaussian-mask sum) travel in ONE 14-float all-reduce
instead of ~27 scalar `reduce_mean`s, and the teacher's buffers are restored from a device-side snapshot instead
of `load_state_dict` from host memory every step (:465).
"""
import torch
import torch.distributed as dist
import torch.nn as nn

from unidistill.layers.losses import distill as L
from unidistill.utils import torch_dist

from . import distill as D

DISTILL_WEIGHTS = (100.0, 40.0, 10.0)      # feature, relation, response (:505-509)


class FlatGradBucket:
    """All student gradients live in one flat fp32 buffer (param.grad are views), so the data-parallel exchange is a
    single all-reduce over NVLink.  Works with any backend (gloo in the CPU tests)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            # same strides as the parameter (NHWC-format conv weights stay NHWC): dense, so the segment is exact
            p.grad = self.flat[off: off + p.numel()].as_strided(p.size(), p.stride())
            off += p.numel()
        self.numel = n

    def zero(self):
        self.flat.zero_()

    def all_reduce_mean(self):
        world = torch_dist.get_world_size()
        if world > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.mul_(1.0 / world)
        return self.flat


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


def fused_normalisers(ret, valid, mask_sum):
    """-> mean over ranks of [num_pos (T), num_hm (T), #valid boxes, Gaussian-mask sum] in one all-reduce."""
    tasks = sorted(ret["mask"].keys()) if isinstance(ret["mask"], dict) else list(range(len(ret["mask"])))
    parts = [ret["mask"][t].float().sum() for t in tasks]
    parts += [ret["heatmap"][t].eq(1).float().sum() for t in tasks]
    parts += [valid.float().sum(), mask_sum]
    vec = torch.stack([p.reshape(()) for p in parts])
    vec = torch_dist.reduce_mean(vec)
    T = len(tasks)
    return vec[:T], vec[T: 2 * T], vec[2 * T], vec[2 * T + 1]


class DistillStep:
    """student / teacher: BEVFusionCenterHead models on the same device.  `geometry`: dict(pc_range, voxel_size,
    out_size_factor).  Call with a device-resident batch dict (points list | None, imgs | None, mats_dict | None,
    gt_boxes (B,M,9), gt_labels (B,M)); returns a dict of detached loss terms."""

    def __init__(self, student, teacher, geometry, lr=2e-4, weight_decay=1e-2, clamp_min=1e-4, optimizer=True,
                 fold_teacher_bn=True, teacher_cuda_graph=True):
        self.student, self.teacher = student, teacher
        self.pc_range = list(geometry["pc_range"])
        self.voxel_size = list(geometry["voxel_size"])
        self.osf = int(geometry["out_size_factor"])
        self.clamp_min = clamp_min
        self.keep_features = False          # tests: keep the student/teacher tensors of the last step in self.last
        self.last = None
        for p in teacher.parameters():                         # freeze_model (:427-433)
            p.requires_grad = False
        teacher.eval()
        teacher.det_head.dense_head.distill = True             # (:419)
        # the frozen eval-mode teacher: BatchNorm folded into the dense convolutions, dense trunk replayed from a CUDA graph
        self.folded_bn = fold_frozen_batchnorm(teacher) if fold_teacher_bn else 0
        if teacher_cuda_graph and next(teacher.parameters()).is_cuda:
            teacher.use_cuda_graph(True)
        self._teacher_buffers = [(b, b.detach().clone()) for b in teacher.buffers()]
        student.train()
        self.bucket = FlatGradBucket(student.parameters())
        self.opt = None
        if optimizer:
            fused = self.bucket.flat.is_cuda
            self.opt = torch.optim.AdamW(self.bucket.params, lr=lr, weight_decay=weight_decay, fused=fused)

    def _restore_teacher(self):
        """The reference reloads the teacher's state_dict every step (:465) because pytorch-lightning may have put it
        back into train mode; an eval-mode teacher cannot change its buffers, so there is nothing to restore."""
        if not any(m.training for m in self.teacher.modules()):
            return
        for b, snap in self._teacher_buffers:
            b.copy_(snap)

    def losses(self, batch):
        points, imgs, metas = batch.get("points"), batch.get("imgs"), batch.get("mats_dict")
        gt = gt_with_labels(batch["gt_boxes"], batch["gt_labels"])
        valid = D.valid_box_prefix(batch["gt_boxes"])                               # (:447-455)
        corners = D.box_corners_bev(gt, self.pc_range, self.voxel_size, self.osf)   # (:470-483)
        hw = None
        norm = {}

        def normalisers(ret):
            h, w = ret["heatmap"][0].shape[-2:]
            mask = D.response_mask(gt, (h, w), self.pc_range, self.voxel_size, self.osf)
            num_pos, num_hm, n_valid, mask_sum = fused_normalisers(ret, valid, mask.sum())
            ret["num_pos_reduced"], ret["num_hm_reduced"] = num_pos, num_hm
            norm["n_valid"], norm["mask_sum"] = n_valid, mask_sum

        ret_dict, tb, feat_s, bev_s, resp_s, _ = self.student(points, imgs, metas, gt, normalisers=normalisers)
        with torch.no_grad():
            self._restore_teacher()
            feat_t, bev_t, resp_t = self.teacher(points, imgs, metas, gt, return_feature=True)
        loss_feature = L.FeatureDistillLoss(feat_s, feat_t, corners, valid, weight=norm["n_valid"])
        loss_rel = L.BEVDistillLoss(bev_s, bev_t, corners, valid, weight=norm["n_valid"])
        loss_cls, loss_reg = L.ResponseDistillLoss(resp_s, resp_t, gt, self.pc_range, self.voxel_size, self.osf,
                                                   clamp_min=self.clamp_min, mask_weight=norm["mask_sum"])
        wf, wr, wp = DISTILL_WEIGHTS
        loss = ret_dict["loss"].mean() + wf * loss_feature + wr * loss_rel + wp * (loss_cls + loss_reg)
        terms = {"loss": loss, "loss_det": ret_dict["loss"].detach(), "loss_feature": loss_feature.detach(),
                 "loss_bev_rel": loss_rel.detach(), "loss_resp_cls": loss_cls.detach(),
                 "loss_resp_reg": loss_reg.detach()}
        del hw
        if self.keep_features:
            self.last = dict(feat_s=feat_s, bev_s=bev_s, resp_s=resp_s, feat_t=feat_t, bev_t=bev_t, resp_t=resp_t,
                             gt=gt, valid=valid, corners=corners, norm=dict(norm))
        return terms

    def __call__(self, batch):
        self.bucket.zero()
        terms = self.losses(batch)
        terms["loss"].backward()
        self.bucket.all_reduce_mean()
        if self.opt is not None:
            self.opt.step()
        terms["loss"] = terms["loss"].detach()
        return terms
