"""FCOSAssigner -- top-k nearest-anchor target assignment for the CenterPoint head.

Same constructor, same `assign_targets(gt_boxes) -> {heatmap, ind, mask, cat, box_encoding}` contract and
the same results as the reference (layers/head/det3d/target_assigner/fcos_assigner.py:73-285), but batched:
the reference runs B x n_tasks Python iterations with `.sum() == 0` host syncs and ~40 tiny kernels each
(executed for the student AND the teacher every step); here the whole batch and all tasks go through one
static-shape tensor program with no host synchronisation.

Semantics kept on purpose:
  * a sample's GT list ends at its last non-zero row, and row 0 is always kept (reference :105-110, quirk 8);
  * inside a task, GTs are ordered class by class (reference :127-141) -- ties in "nearest GT" go to the
    first one in that order;
  * positives = union over the task's GTs of their `assign_topk` nearest anchor points; each positive anchor
    regresses to its nearest GT; outputs are listed in ascending anchor order and zero padded to max_objs;
  * log(0) for degenerate boxes yields -inf in box_encoding exactly as in the reference (the caller zeroes
    infs, BEVFusion_nuscenes_centerhead_fusion_exp.py:125-127).
"""
import math
import os

import numpy as np
import torch

from .base_assigner import BaseAssigner

# CUDA inputs go through the fused assignment kernel (csrc/fcos_assign.cu); UD_FCOS_KERNEL=0 keeps the tensor program
# (which is also what runs for CPU tensors, e.g. in the golden-fixture tests)
USE_ASSIGN_KERNEL = os.environ.get("UD_FCOS_KERNEL", "1") != "0"
# the reference asserts `n_pos < max_objs` when padding (padding_to_maxobjs); checking it costs a host sync, so it is a
# debug switch here -- the un-truncated count is always returned as ret["n_pos"][task] (B,) for a sync-free check
CHECK_OVERFLOW = os.environ.get("UD_ASSIGN_CHECK", "0") != "0"


class FCOSAssigner(BaseAssigner):
    def __init__(self, out_size_factor, tasks, dense_reg, gaussian_overlap, max_objs, min_radius, mapping,
                 grid_size, pc_range, voxel_size, assign_topk, no_log=False, with_velocity=False):
        super().__init__()
        self.out_size_factor = out_size_factor
        self.tasks = tasks
        self.num_classes = sum(len(t["class_names"]) for t in self.tasks)
        self.dense_reg = dense_reg
        self.gaussian_overlap = gaussian_overlap
        self._max_objs = max_objs
        self._min_radius = min_radius
        self.class_to_idx = mapping
        self.grid_size = np.array(grid_size)
        self.pc_range = pc_range
        self.voxel_size = voxel_size
        self.no_log = no_log
        self.assign_topk = assign_topk
        self.with_velocity = with_velocity
        self.default_box_dims = 10 if with_velocity else 8
        self._lut_cache = {}

    # kept for API compatibility with the reference
    def generate_anchor_grid(self, featmap_size, offsets, stride, device):
        step_x, step_y = featmap_size
        shift = offsets * stride
        grid_x = torch.linspace(shift, (step_x - 1) * stride + shift, steps=step_x, device=device)
        grid_y = torch.linspace(shift, (step_y - 1) * stride + shift, steps=step_y, device=device)
        grids_x, grids_y = torch.meshgrid(grid_y, grid_x, indexing="ij")
        return grids_x.reshape(-1), grids_y.reshape(-1)

    def limit_period(self, val, offset=0.5, period=math.pi):
        return val - torch.floor(val / period + offset) * period

    def _luts(self, device):
        """class index (1-based, as in gt_boxes[..., -1]) -> (task id, class offset inside the task)."""
        key = str(device)
        if key not in self._lut_cache:
            max_cls = max(self.class_to_idx.values()) + 1
            task_of = torch.full((max_cls + 1,), -1, dtype=torch.long)
            offset_of = torch.zeros((max_cls + 1,), dtype=torch.long)
            for t, task in enumerate(self.tasks):
                for off, name in enumerate(task["class_names"]):
                    task_of[self.class_to_idx[name]] = t
                    offset_of[self.class_to_idx[name]] = off
            self._lut_cache[key] = (task_of.to(device), offset_of.to(device))
        return self._lut_cache[key]

    def _assign_kernel(self, boxes, task_id, cls_off, anchors, fw, fh, stride, max_objs):
        """The per-task part of assign_targets as ONE kernel launch for the whole batch and all tasks
        (csrc/fcos_assign.cu, `ud_fcos_assign`); bit-identical to the tensor program below."""
        from unidistill_b200 import native

        lib = native.load()
        dev = boxes.device
        B, M, NB = boxes.shape
        T = len(self.tasks)
        n_anchor = fw * fh
        prefix = [0]
        for task in self.tasks:
            prefix.append(prefix[-1] + len(task["class_names"]))
        boxes = boxes.contiguous()
        tid = task_id.to(torch.int32).contiguous()
        coff = cls_off.to(torch.int32).contiguous()
        anchors = anchors.float().contiguous()
        heat = torch.zeros((B * prefix[-1] * n_anchor,), dtype=torch.float32, device=dev)
        ind = torch.empty((T, B, max_objs), dtype=torch.long, device=dev)
        mask = torch.empty((T, B, max_objs), dtype=torch.bool, device=dev)
        cat = torch.empty((T, B, max_objs), dtype=torch.long, device=dev)
        enc = torch.empty((T, B, max_objs, NB + 1), dtype=torch.float32, device=dev)
        if NB + 1 != self.default_box_dims:
            raise RuntimeError("box encoding width %d != %d" % (NB + 1, self.default_box_dims))
        pre = native.host_ints(prefix)
        n_pos = torch.empty((T, B), dtype=torch.int32, device=dev)
        rc = lib.ud_fcos_assign(native.ptr(boxes), native.ptr(tid), native.ptr(coff), native.ptr(anchors), int(B), int(M),
                                int(NB), int(fw), int(fh), float(stride), T, native.host_ptr(pre), int(self.assign_topk),
                                int(max_objs), float(self.voxel_size[0]), float(self.voxel_size[1]),
                                float(self.out_size_factor), native.ptr(heat), native.ptr(ind), native.ptr(mask),
                                native.ptr(cat), native.ptr(enc), native.ptr(n_pos), native.stream_ptr(dev))
        native.check(rc, "ud_fcos_assign")
        if CHECK_OVERFLOW and int(n_pos.max()) >= max_objs:          # debug mode only: host sync
            raise AssertionError("FCOSAssigner: %d positives >= max_objs %d (reference padding_to_maxobjs asserts)"
                                 % (int(n_pos.max()), max_objs))
        heatmaps, inds, masks, cats, encs = {}, {}, {}, {}, {}
        for t, task in enumerate(self.tasks):
            n_cls = prefix[t + 1] - prefix[t]
            heatmaps[t] = heat[B * n_anchor * prefix[t]: B * n_anchor * prefix[t + 1]].view(B, n_cls, fh, fw)
            inds[t], masks[t], cats[t], encs[t] = ind[t], mask[t], cat[t], enc[t]
        return {"heatmap": heatmaps, "ind": inds, "mask": masks, "cat": cats, "box_encoding": encs,
                "n_pos": {t: n_pos[t] for t in range(T)}}

    @torch.no_grad()
    def assign_targets(self, gt_boxes):
        """gt_boxes (B, M, C + 1) with the 1-based class in the last column."""
        device = gt_boxes.device
        max_objs = self._max_objs * self.dense_reg
        fw, fh = [int(v) for v in (self.grid_size[:2] // self.out_size_factor)]
        stride = self.out_size_factor
        B, M = gt_boxes.shape[0], gt_boxes.shape[1]
        n_anchor = fw * fh
        topk = self.assign_topk
        dims = self.default_box_dims

        grids_x, grids_y = self.generate_anchor_grid((fw, fh), 0.0, stride, device)
        anchors = torch.stack([grids_y, grids_x], dim=1)                       # (A, 2) = (x, y), x fastest

        classes = gt_boxes[:, :, -1].int().long()
        boxes = gt_boxes[:, :, :-1].clone().float()
        # valid prefix: up to the last row whose sum is non-zero; row 0 always (reference :105-110)
        ar = torch.arange(M, device=device)
        nonzero = boxes.sum(dim=2) != 0
        last = (nonzero.long() * ar).max(dim=1).values
        valid = ar.unsqueeze(0) <= last.unsqueeze(1)

        task_of, offset_of = self._luts(device)
        cls_c = classes.clamp(0, task_of.numel() - 1)
        task_id = torch.where(valid, task_of[cls_c], torch.full_like(cls_c, -1))
        cls_off = offset_of[cls_c]

        # reorder every sample's GTs to (task, class offset, original index): the reference's concat order
        sort_key = torch.where(task_id >= 0, task_id * 64 + cls_off, torch.full_like(task_id, 1 << 20))
        order = torch.sort(sort_key, dim=1, stable=True).indices
        boxes = torch.gather(boxes, 1, order.unsqueeze(2).expand_as(boxes))
        task_id = torch.gather(task_id, 1, order)
        cls_off = torch.gather(cls_off, 1, order)

        # to feature-map (voxel) units (reference :149-160)
        boxes[:, :, 0] = (boxes[:, :, 0] - self.pc_range[0]) / self.voxel_size[0]
        boxes[:, :, 1] = (boxes[:, :, 1] - self.pc_range[1]) / self.voxel_size[1]
        boxes[:, :, 3] = boxes[:, :, 3] / self.voxel_size[0]
        boxes[:, :, 4] = boxes[:, :, 4] / self.voxel_size[1]
        boxes[:, :, 6] = self.limit_period(boxes[:, :, 6], offset=0.5, period=math.pi * 2)

        if gt_boxes.is_cuda and USE_ASSIGN_KERNEL and M <= 128 and n_anchor <= 32768 and topk <= 9:
            return self._assign_kernel(boxes, task_id, cls_off, anchors, fw, fh, stride, max_objs)

        # squared centre distance anchor <-> GT: (B, M, A)
        d2 = torch.pow(anchors.view(1, 1, n_anchor, 2) - boxes[:, :, None, :2], 2).sum(dim=3)
        near = torch.topk(d2, topk, dim=2, largest=False).indices              # (B, M, topk)

        heatmaps, inds, masks, cats, encs, n_poss = {}, {}, {}, {}, {}, {}
        inf = torch.full((), float("inf"), device=device)      # device-side fill: CUDA-graph capturable
        rank = torch.arange(max_objs, device=device).unsqueeze(0)
        for t, task in enumerate(self.tasks):
            n_cls = len(task["class_names"])
            member = task_id == t                                              # (B, M)
            hit = torch.zeros((B, n_anchor + 1), dtype=torch.bool, device=device)
            # anchors in any member GT's top-k; non-members are routed to the dummy column A
            tgt = torch.where(member.unsqueeze(2), near, torch.full_like(near, n_anchor)).reshape(B, -1)
            hit.scatter_(1, tgt, True)
            pos = hit[:, :n_anchor]                                            # (B, A)
            # nearest member GT of every anchor (ties -> first in task order)
            d2_t = torch.where(member.unsqueeze(2), d2, inf)
            gt_of_anchor = d2_t.argmin(dim=1)                                  # (B, A)
            n_pos = pos.sum(dim=1, keepdim=True)
            # positives first, ascending anchor index
            pos_order = torch.sort((~pos).to(torch.uint8), dim=1, stable=True).indices[:, :max_objs]
            if pos_order.shape[1] < max_objs:                                  # fewer anchors than slots (tiny maps)
                pos_order = torch.nn.functional.pad(pos_order, (0, max_objs - pos_order.shape[1]))
            keep = rank < n_pos                                                # (B, max_objs)
            gt_sel = torch.gather(gt_of_anchor, 1, pos_order)
            tgt_box = torch.gather(boxes, 1, gt_sel.unsqueeze(2).expand(B, max_objs, boxes.shape[2]))
            tgt_cls = torch.gather(cls_off, 1, gt_sel)
            anc = anchors[pos_order]                                           # (B, max_objs, 2)
            enc = torch.cat([
                (tgt_box[:, :, 0:2] - anc) / self.out_size_factor,
                tgt_box[:, :, 2:3],
                torch.log(tgt_box[:, :, 3:4] * self.voxel_size[0]),
                torch.log(tgt_box[:, :, 4:5] * self.voxel_size[1]),
                torch.log(tgt_box[:, :, 5:6]),
                torch.sin(tgt_box[:, :, 6:7]),
                torch.cos(tgt_box[:, :, 6:7]),
                tgt_box[:, :, 7:],
            ], dim=2)
            keep_f = keep.unsqueeze(2)
            encs[t] = torch.where(keep_f, enc, torch.zeros_like(enc)).float().contiguous()
            inds[t] = torch.where(keep, pos_order, torch.zeros_like(pos_order)).long().contiguous()
            masks[t] = keep.contiguous()
            n_poss[t] = n_pos[:, 0].to(torch.int32)
            cats[t] = torch.where(keep, tgt_cls, torch.zeros_like(tgt_cls)).long().contiguous()
            heat = torch.zeros((B, n_cls, n_anchor), dtype=torch.float32, device=device)
            # (B, A); a sample without a GT of this task points every anchor at row 0, whose class offset may not exist
            # in this task: clamp (those anchors are not positive, the scattered value is 0)
            cls_of_anchor = torch.gather(cls_off, 1, gt_of_anchor).clamp(max=n_cls - 1)
            heat.scatter_(1, cls_of_anchor.unsqueeze(1), pos.unsqueeze(1).float())
            heatmaps[t] = heat.view(B, n_cls, fh, fw).contiguous()
            if enc.shape[2] != dims:
                raise RuntimeError("box encoding width %d != %d" % (enc.shape[2], dims))
        if CHECK_OVERFLOW and max(int(v.max()) for v in n_poss.values()) >= max_objs:
            raise AssertionError("FCOSAssigner: positives >= max_objs %d (reference padding_to_maxobjs asserts)" % max_objs)
        return {"heatmap": heatmaps, "ind": inds, "mask": masks, "cat": cats, "box_encoding": encs, "n_pos": n_poss}
