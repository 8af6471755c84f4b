"""CenterPointGenProposals -- top-k decoding of the CenterHead maps + rotated-BEV NMS (inference post-processing).

Same constructor, helper methods and per-sample `proposal_layer` contract as the reference
(unidistill/layers/head/det3d/generate_proposals/centerpoint_gen_proposals.py:8-272); `_nms_gpu_3d` goes through the
module-level `iou3d_nms_cuda` drop-in exactly like the reference's call site (:97-100).  `generate_predicted_boxes`
additionally has a batched device-side path (`_decode_task` + unidistill_b200.nms.nms_bev_batched): all (task, sample)
pairs share one NMS launch and the only device->host read is the vector of kept counts at the very end.
"""
import torch
import torch.nn as nn

from unidistill_b200 import nms as _nms

from . import iou3d_nms_cuda
from .base_gen_proposals import BaseGenProposals


class CenterPointGenProposals(BaseGenProposals):
    # constructor arguments of the reference (:9-39), stored under the same attribute names
    _ARGS = ("dataset_name", "class_names", "post_center_limit_range", "score_threshold", "pc_range", "out_size_factor",
             "voxel_size", "no_log", "nms_iou_threshold_train", "nms_pre_max_size_train", "nms_post_max_size_train",
             "nms_iou_threshold_test", "nms_pre_max_size_test", "nms_post_max_size_test")

    def __init__(self, dataset_name, class_names, post_center_limit_range, score_threshold, pc_range, out_size_factor,
                 voxel_size, no_log, nms_iou_threshold_train, nms_pre_max_size_train, nms_post_max_size_train,
                 nms_iou_threshold_test, nms_pre_max_size_test, nms_post_max_size_test):
        super().__init__()
        given = locals()
        for name in self._ARGS:
            setattr(self, name, given[name])
        self.training = True

    # ---- map helpers (same names and results as the reference's, :41-83) ---------------------------
    def _gather_feat(self, feat, ind, mask=None):
        """feat (B, N, C), ind (B, K) -> (B, K, C); with `mask` (B, K) the selected rows, flattened to (n, C)."""
        picked = torch.take_along_dim(feat, ind.unsqueeze(-1), dim=1)
        return picked if mask is None else picked[mask]

    def _transpose_and_gather_feat(self, feat, ind):
        """feat (B, C, H, W), ind (B, K) flat positions -> (B, K, C), without materialising the NHWC copy."""
        flat = feat.flatten(2)                                                      # (B, C, H*W)
        return torch.take_along_dim(flat, ind.unsqueeze(1), dim=2).transpose(1, 2).contiguous()

    def _nms(self, heat, kernel=3):
        """Keeps the local maxima of a heat-map (kernel x kernel neighbourhood), zeroes the rest."""
        peak = nn.functional.max_pool2d(heat, kernel, stride=1, padding=(kernel - 1) // 2)
        return torch.where(peak == heat, heat, torch.zeros_like(heat))

    def _topk(self, scores, K=40):
        """Two-stage top-K as in the reference (per class, then across classes: tie order is part of the contract)
        -> (score, flat index, class (int32), y, x), each (B, K)."""
        B, n_cls, H, W = scores.shape
        per_cls_score, per_cls_pos = scores.flatten(2).topk(K, dim=2)               # (B, n_cls, K)
        score, pick = per_cls_score.flatten(1).topk(K, dim=1)                       # pick in [0, n_cls * K)
        pos = torch.take_along_dim(per_cls_pos.flatten(1), pick, dim=1)             # flat position inside the map
        cls = torch.div(pick, K, rounding_mode="floor").int()
        ys = torch.div(pos, W, rounding_mode="floor").float()
        xs = (pos - torch.div(pos, W, rounding_mode="floor") * W).float()
        return score, pos, cls, ys, xs

    # ---- per-sample NMS through the drop-in extension module (reference call site :97-100) ---------------------
    def _nms_gpu_3d(self, boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
        if boxes.shape[1] != 7:
            raise ValueError("rotated NMS takes (N, 7) boxes [x, y, z, dx, dy, dz, heading]")
        ranked = torch.argsort(scores, dim=0, descending=True)[:pre_maxsize]        # [:None] keeps everything
        kept_slots = torch.empty(ranked.numel(), dtype=torch.long)                  # CPU buffer, as the extension expects
        n_kept = iou3d_nms_cuda.nms_gpu(boxes[ranked].contiguous(), kept_slots, thresh)
        return ranked[kept_slots[:n_kept].to(boxes.device)][:post_max_size].contiguous()

    def _boxes3d_to_bevboxes_lidar_torch(self, boxes3d):
        """(N, 7+) [x, y, z, dx, dy, dz, ..., heading] -> (N, 5) axis-aligned BEV extents + heading."""
        centre, half = boxes3d[:, 0:2], boxes3d[:, 3:5] / 2
        return torch.cat([centre - half, centre + half, boxes3d[:, -1:]], dim=1)

    def get_nms_scores(self, scores):
        return scores

    def nms_options(self, boxes3d, labels, scores, nms_scores):
        if nms_scores.shape[0] == 0:
            keep = []
        else:
            keep = self._nms_gpu_3d(boxes3d[:, :7], nms_scores, self.nms_iou_threshold_use,
                                    pre_maxsize=self.nms_pre_max_size_use, post_max_size=self.nms_post_max_size_use)
        return boxes3d[keep], labels[keep], scores[keep]

    # ---- decoding shared by the per-sample and the batched path ---------------------------------------
    def _nms_score(self, scores, inds, task_id, iouhm):
        """CenterPoint: the NMS ranks by the class score itself (reference :164)."""
        return scores

    def _decode_task(self, heat, rots, rotc, hei, dim, vel, reg, task_id=-1, iouhm=None):
        """Top-k + box decoding of one task for the whole batch (reference proposal_layer :160-214 up to the mask):
        -> final_box_preds (B,K,7|9), final_scores (B,K), clses (B,K) float, nms_scores (B,K), mask (B,K) bool."""
        batch = heat.size(0)
        K = self.nms_pre_max_size_use
        scores, inds, clses, ys, xs = self._topk(heat, K=K)
        nms_scores = self._nms_score(scores, inds, task_id, iouhm)
        assert reg is not None
        reg = self._transpose_and_gather_feat(reg, inds).view(batch, K, 2)
        xs = xs.view(batch, K, 1) + reg[:, :, 0:1]
        ys = ys.view(batch, K, 1) + reg[:, :, 1:2]
        rots = self._transpose_and_gather_feat(rots, inds).view(batch, K, 1)
        rotc = self._transpose_and_gather_feat(rotc, inds).view(batch, K, 1)
        rot = torch.atan2(rots, rotc)
        hei = self._transpose_and_gather_feat(hei, inds).view(batch, K, 1)
        dim = self._transpose_and_gather_feat(dim, inds).view(batch, K, 3)
        clses = clses.view(batch, K).float()
        scores = scores.view(batch, K)
        xs = xs.view(batch, K, 1) * self.out_size_factor * self.voxel_size[0] + self.pc_range[0]
        ys = ys.view(batch, K, 1) * self.out_size_factor * self.voxel_size[1] + self.pc_range[1]
        if self.dataset_name == "nuscenes":
            vel = self._transpose_and_gather_feat(vel, inds).view(batch, K, 2)
            final_box_preds = torch.cat([xs, ys, hei, dim, rot, vel], dim=2)
        else:
            final_box_preds = torch.cat([xs, ys, hei, dim, rot], dim=2)
        assert self.post_center_limit_range is not None and self.score_threshold is not None
        rng = torch.tensor(self.post_center_limit_range, dtype=final_box_preds.dtype, device=final_box_preds.device)
        mask = (final_box_preds[..., :3] >= rng[:3]).all(2)
        mask &= (final_box_preds[..., :3] <= rng[3:]).all(2)
        mask &= scores > self.score_threshold
        return final_box_preds, scores, clses, nms_scores, mask

    @torch.no_grad()
    def proposal_layer(self, heat, rots, rotc, hei, dim, vel, reg=None, raw_rot=False, task_id=-1, iouhm=None):
        """Per-sample path with the reference's return contract (list of {"boxes","scores","labels"})."""
        assert raw_rot is False
        boxes, scores, clses, nms_scores, mask = self._decode_task(heat, rots, rotc, hei, dim, vel, reg, task_id, iouhm)
        out = []
        for i in range(heat.size(0)):
            cmask = mask[i, :]
            b, l, s = self.nms_options(boxes[i, cmask], clses[i, cmask], scores[i, cmask], nms_scores[i, cmask])
            out.append({"boxes": b, "scores": s, "labels": l.long()})
        return out

    def _use_sizes(self):
        if self.training:
            self.nms_iou_threshold_use = self.nms_iou_threshold_train
            self.nms_pre_max_size_use = self.nms_pre_max_size_train
            self.nms_post_max_size_use = self.nms_post_max_size_train
        else:
            self.nms_iou_threshold_use = self.nms_iou_threshold_test
            self.nms_pre_max_size_use = self.nms_pre_max_size_test
            self.nms_post_max_size_use = self.nms_post_max_size_test

    def _task_inputs(self, pred_dict):
        batch_hm = pred_dict["hm"].sigmoid()
        if not self.no_log:
            batch_dim = torch.clamp(torch.exp(pred_dict["dim"]), min=0.001, max=30)
        else:
            batch_dim = pred_dict["dim"]
        rots = pred_dict["rot"][:, 0].unsqueeze(1)
        rotc = pred_dict["rot"][:, 1].unsqueeze(1)
        vel = pred_dict["vel"] if self.dataset_name == "nuscenes" else None
        return batch_hm, rots, rotc, pred_dict["height"], batch_dim, vel, pred_dict["reg"], pred_dict.get("iou")

    @torch.no_grad()
    def generate_predicted_boxes(self, forward_ret_dict, data_dict):
        """Decode + NMS of every task and sample (reference :216-272 / iou_aware :141-247): fills data_dict with
        pred_dicts [{pred_boxes, pred_scores, pred_labels}], rois, roi_scores, roi_labels, has_class_labels."""
        pred_dicts = forward_ret_dict["multi_head_features"]
        self._use_sizes()
        T = len(pred_dicts)
        post = int(self.nms_post_max_size_use)
        boxes_t, scores_t, labels_t, nvalid_t = [], [], [], []
        for task_id, pred_dict in enumerate(pred_dicts):
            hm, rots, rotc, hei, dim, vel, reg, iou = self._task_inputs(pred_dict)
            boxes, scores, clses, nms_scores, mask = self._decode_task(hm, rots, rotc, hei, dim, vel, reg, task_id, iou)
            key = torch.where(mask, nms_scores, torch.full_like(nms_scores, float("-inf")))
            order = torch.sort(key, dim=1, descending=True, stable=True)[1]          # masked-out candidates last
            boxes_t.append(torch.gather(boxes, 1, order.unsqueeze(2).expand_as(boxes)))
            scores_t.append(torch.gather(scores, 1, order))
            labels_t.append(torch.gather(clses, 1, order))
            nvalid_t.append(mask.sum(1).int())
        B, K, D = boxes_t[0].shape
        all_boxes = torch.stack(boxes_t, 0)                                          # (T, B, K, D)
        n_valid = torch.stack(nvalid_t, 0).reshape(T * B)
        keep, num = _nms.nms_bev_batched(all_boxes.reshape(T * B, K, D), n_valid, self.nms_iou_threshold_use, post)
        idx = keep.clamp(min=0).long().reshape(T, B, post)
        sel_boxes = torch.gather(all_boxes, 2, idx.unsqueeze(3).expand(T, B, post, D))
        sel_scores = torch.gather(torch.stack(scores_t, 0), 2, idx)
        sel_labels = torch.gather(torch.stack(labels_t, 0), 2, idx).long()
        offsets = [1]
        for names in self.class_names[:-1]:
            offsets.append(offsets[-1] + len(names))
        sel_labels = sel_labels + torch.tensor(offsets, device=sel_labels.device).view(T, 1, 1)   # global, 1-based
        counts = num.reshape(T, B).cpu()                                             # the one device->host read
        num_rois = post * len(self.class_names)
        out, rois, roi_scores, roi_labels = [], [], [], []
        for b in range(B):
            fb = torch.cat([sel_boxes[t, b, : int(counts[t, b])] for t in range(T)])
            fs = torch.cat([sel_scores[t, b, : int(counts[t, b])] for t in range(T)])
            fl = torch.cat([sel_labels[t, b, : int(counts[t, b])] for t in range(T)])
            roi = fb.new_zeros(num_rois, D)
            roi_score = fs.new_zeros(num_rois)
            roi_label = fl.new_zeros(num_rois)
            n = fb.shape[0]
            roi[:n], roi_score[:n], roi_label[:n] = fb, fs, fl
            rois.append(roi)
            roi_scores.append(roi_score)
            roi_labels.append(roi_label)
            out.append({"pred_boxes": fb, "pred_scores": fs, "pred_labels": fl})
        data_dict["pred_dicts"] = out
        data_dict["rois"] = torch.stack(rois, dim=0)
        data_dict["roi_scores"] = torch.stack(roi_scores, dim=0)
        data_dict["roi_labels"] = torch.stack(roi_labels, dim=0)
        data_dict["has_class_labels"] = True
        data_dict.pop("batch_index", None)
        return data_dict
