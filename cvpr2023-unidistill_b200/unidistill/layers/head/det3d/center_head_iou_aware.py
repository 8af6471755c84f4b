"""CenterHeadIouAware -- CenterPoint head with an IoU-aware branch and uncertainty-weighted losses
(reference layers/head/det3d/center_head_iou_aware.py:12-298).

Differences from the reference are purely operational (same numbers):
  * no `.item()` / Python `if` on device values: `if loc_loss.item() < 1` (:132) becomes a torch.where;
    tb_dict holds detached 0-dim tensors instead of ~14 host round trips per task;
  * the IoU-aware target uses the row-wise BEV IoU directly instead of the diagonal of a
    (B*max_objs)^2 pairwise matrix (:289-293: 10^8 elements, ~400 MB temporaries per task at bs=4);
  * the per-call scalar `reduce_mean`s can be replaced by one fused all-reduce: pass
    forward_ret_dict["num_pos_reduced"] = tensor(n_tasks) (see unidistill_b200.step).
"""
import os

import torch

from unidistill.layers.head.det3d.center_head import CenterHead
from unidistill.layers.losses.det3d import (AutomaticWeightedLoss, CenterNetFocalLoss, CenterNetRegLoss, FocalLoss,
                                            _transpose_and_gather_feat)
from unidistill.utils.det3d_utils import box_utils
from unidistill.utils.torch_dist import reduce_mean


# The per-task loop of get_loss issues ~250 small kernels per task and as many again in its backward; with pre-reduced
# normalisers (unidistill_b200.step) the six tasks have identical shapes except for the class count of the heat-map, so
# the same arithmetic runs ONCE on tensors with a task dimension (`_get_loss_batched`).  UD_BATCH_LOSS=0 disables it.
BATCH_LOSS = os.environ.get("UD_BATCH_LOSS", "1") != "0"


class CenterHeadIouAware(CenterHead):
    def __init__(self, dataset_name, tasks, target_assigner, proposal_layer, out_size_factor, input_channels,
                 grid_size, point_cloud_range, code_weights, loc_weight, iou_weight, share_conv_channel,
                 common_heads, upsample_for_pedestrian=False, predict_boxes_when_training=False, mode="3d",
                 init_bias=-2.19):
        super().__init__(dataset_name, tasks, target_assigner, proposal_layer, input_channels, grid_size,
                         point_cloud_range, code_weights, loc_weight, share_conv_channel, common_heads,
                         upsample_for_pedestrian=upsample_for_pedestrian,
                         predict_boxes_when_training=predict_boxes_when_training, mode=mode, init_bias=init_bias)
        self.auto_loss = AutomaticWeightedLoss(num=len(self.code_weights) + 2)
        self.iou_weight = iou_weight
        self.out_size_factor = out_size_factor

    def _box_heads(self):
        return ["reg", "height", "dim", "rot"] + (["vel"] if self.dataset == "nuscenes" else []) + ["iou"]

    def _can_batch_loss(self, ret):
        if not (BATCH_LOSS and "num_pos_reduced" in ret and "num_hm_reduced" in ret):
            return False
        if type(self.crit) not in (FocalLoss, CenterNetFocalLoss) or type(self.crit_reg) is not CenterNetRegLoss \
                or type(self.crit_iou_aware) is not CenterNetRegLoss:
            return False
        preds = ret["multi_head_features"]
        n_code = 10 if self.dataset == "nuscenes" else 8
        try:
            widths = [[p[h].shape[1] for h in self._box_heads()] for p in preds]
        except KeyError:
            return False
        if any(w != widths[0] for w in widths) or sum(widths[0]) != n_code + 1:
            return False
        k = ret["ind"][0].shape
        return all(ret["ind"][t].shape == k and ret["box_encoding"][t].shape[-1] >= n_code for t in range(len(preds)))

    def _task_onehot(self, like, classes):
        """(sum classes, T) 0/1 matrix adding heat-map channels up per task; built once per device."""
        cache = self.__dict__.setdefault("_onehot_cache", {})
        key = (like.device, like.dtype, tuple(classes))
        if key not in cache:
            m = torch.zeros(sum(classes), len(classes))
            c0 = 0
            for t, c in enumerate(classes):
                m[c0: c0 + c, t] = 1
                c0 += c
            cache[key] = m.to(device=like.device, dtype=like.dtype)
        return cache[key]

    def _get_loss_batched(self, ret):
        """get_loss with a task dimension: same terms, same tb_dict, same side effects on the dict (quirk 6:
        pred_dict["hm"] becomes the clamped sigmoid; "pred_box_encoding" is filled) as the per-task loop below."""
        preds = ret["multi_head_features"]
        T = len(preds)
        n_code = 10 if self.dataset == "nuscenes" else 8
        stride, voxel_size = self.out_size_factor, self.proposal_layer.voxel_size
        num_pos, num_hm = ret["num_pos_reduced"], ret["num_hm_reduced"]                # (T,), (T,)
        # ---- heat-maps: all tasks' classes side by side
        classes = [p["hm"].shape[1] for p in preds]
        hm_all = self._sigmoid(torch.cat([p["hm"] for p in preds], dim=1))             # (B, sum classes, H, W)
        heat_all = torch.cat([ret["heatmap"][t] for t in range(T)], dim=1)
        c0 = 0
        for p, c in zip(preds, classes):
            p["hm"] = hm_all[:, c0: c0 + c]                                            # quirk 6
            c0 += c
        per_task = self.crit.channel_sums(hm_all, heat_all) @ self._task_onehot(hm_all, classes)
        hm_loss = -per_task / torch.where(num_hm == 0, torch.ones_like(num_hm), num_hm)
        # ---- box codes at the assigned cells: (B, T, K, n_code + 1)
        heads = self._box_heads()
        B, _, H, W = preds[0][heads[0]].shape
        codes = torch.cat([p[h] for p in preds for h in heads], dim=1).view(B, T, n_code + 1, H * W)
        ret["pred_box_encoding"] = {t: codes[:, t].view(B, n_code + 1, H, W) for t in range(T)}
        ind = torch.stack([ret["ind"][t] for t in range(T)], dim=1)                    # (B, T, K)
        mask = torch.stack([ret["mask"][t] for t in range(T)], dim=1)                  # (B, T, K)
        target = torch.stack([ret["box_encoding"][t][:, :, :n_code] for t in range(T)], dim=1)
        K = ind.shape[-1]
        pred = codes.gather(3, ind.unsqueeze(2).expand(B, T, n_code + 1, K)).permute(0, 1, 3, 2).contiguous()
        maskf = mask.float()

        def decode(t):
            x = (t[..., 0:1] * stride * voxel_size[0]).reshape(-1, 1)
            y = (t[..., 1:2] * stride * voxel_size[1]).reshape(-1, 1)
            whl = torch.clamp(torch.exp(t[..., 3:6].clamp(max=10.0)).reshape(-1, 3), min=0.001, max=30)   # see _get_iou_loss
            rot = torch.atan2(t[..., 6], t[..., 7]).reshape(-1, 1)
            z = t[..., 2].reshape(-1, 1)
            return x, y, z, whl, rot

        tx, ty, tz, t_whl, t_rot = decode(target)
        px, py, pz, p_whl, p_rot = decode(pred)
        iou = self._get_3d_iou(tx, ty, t_whl, tz, px, py, p_whl, pz)
        iou_pos = torch.clamp(iou[:, 0], 0, 1).view(B, T, K)
        iou_loss = ((1 - iou_pos) * maskf).sum(dim=(0, 2)) / torch.clamp(num_pos, min=1)
        target_box = torch.cat([tx, ty, tz, t_whl, t_rot], dim=-1)
        pred_box = torch.cat([px, py, pz, p_whl, p_rot], dim=-1).detach()
        bev_iou = box_utils.boxes3d_nearest_bev_iou_paired(target_box, pred_box)
        tar_iou_pred = 2 * (bev_iou.view(B, T, K, 1) - 0.5)

        def reg_loss(regr, gt_regr):          # CenterNetRegLoss._reg_loss with a task dimension -> (T, C)
            m = maskf.unsqueeze(3).expand_as(gt_regr) * (~torch.isnan(gt_regr)).float()
            return torch.abs(regr * m - gt_regr * m).sum(dim=(0, 2)) / (num_pos + 1e-4).unsqueeze(1)

        iou_aware_loss = reg_loss(pred[..., n_code:], tar_iou_pred).sum(dim=1)
        box_loss = reg_loss(pred[..., :n_code], target)                                # (T, n_code)
        loc_loss = (box_loss * self._code_weights_on(box_loss)).sum(dim=1)
        loss = self.auto_loss(hm_loss, loc_loss, iou_aware_loss)
        loss = torch.where(loc_loss.detach() < 1, loss + iou_loss * self.iou_weight, loss)
        n_positive = maskf.sum(dim=(0, 2))
        tb_dict = {}
        for t in range(T):
            key = "task_%d/" % t
            tb_dict.update({
                key + "loss": loss[t].detach(), key + "hm_loss": hm_loss[t].detach(), key + "loc_loss": loc_loss[t].detach(),
                key + "box_loss": box_loss[t].detach(), key + "iou_aware_loss": iou_aware_loss[t].detach(),
                key + "iou_loss": iou_loss[t].detach(), key + "num_positive": n_positive[t]})
        return loss.sum(), tb_dict

    def get_loss(self, forward_ret_dict):
        if self._can_batch_loss(forward_ret_dict):
            return self._get_loss_batched(forward_ret_dict)
        tb_dict = {}
        losses = []
        forward_ret_dict["pred_box_encoding"] = {}
        n_code = 10 if self.dataset == "nuscenes" else 8
        stride = self.out_size_factor
        voxel_size = self.proposal_layer.voxel_size
        pre = forward_ret_dict.get("num_pos_reduced", None)
        code_w = None
        for task_id, pred_dict in enumerate(forward_ret_dict["multi_head_features"]):
            mask = forward_ret_dict["mask"][task_id]
            ind = forward_ret_dict["ind"][task_id]
            target = forward_ret_dict["box_encoding"][task_id]
            heat = forward_ret_dict["heatmap"][task_id]
            if pre is not None:
                num_pos = pre[task_id]
                # heat-map positives are NOT the same count as regression slots (overflow truncation, duplicate classes):
                # without a pre-reduced value the criterion derives it from heatmap.eq(1).sum() itself
                num_hm = forward_ret_dict["num_hm_reduced"][task_id] if "num_hm_reduced" in forward_ret_dict else None
            else:
                num_pos = reduce_mean(mask.float().sum())
                num_hm = None
            # quirk 6: the sigmoid-clamped heat-map replaces the logits in the dict (read later by the
            # response-distillation loss, ..._distill_lidar.py:335)
            pred_dict["hm"] = self._sigmoid(pred_dict["hm"])
            hm_loss = self.crit(pred_dict["hm"], heat, num_hm)
            pred = self._box_codes(pred_dict, with_iou=True)                    # (B, n_code + 1, H, W)
            forward_ret_dict["pred_box_encoding"][task_id] = pred
            iou_loss, iou_aware_loss = self._get_iou_loss(pred[:, : n_code + 1], target[:, :, :n_code], ind, mask,
                                                          stride, voxel_size, num_pos)
            box_loss = self.crit_reg(pred[:, :n_code], mask, ind, target[:, :, :n_code], num_pos)
            if code_w is None:
                code_w = self._code_weights_on(box_loss)
            loc_loss = (box_loss * code_w).sum()
            loss = self.auto_loss(hm_loss, loc_loss, iou_aware_loss)
            # the IoU loss only joins once localisation is good enough (reference :131-133)
            loss = torch.where(loc_loss.detach() < 1, loss + iou_loss * self.iou_weight, loss)
            key = "task_%d/" % task_id
            tb_dict.update({
                key + "loss": loss.detach(), key + "hm_loss": hm_loss.detach(), key + "loc_loss": loc_loss.detach(),
                key + "box_loss": box_loss.detach(), key + "iou_aware_loss": iou_aware_loss.detach(),
                key + "iou_loss": iou_loss.detach(), key + "num_positive": mask.float().sum()})
            losses.append(loss)
        return sum(losses), tb_dict

    def _get_3d_iou(self, tx, ty, t_whl, tz, px, py, p_whl, pz):
        """Axis-aligned 3-D IoU with each extent clamped on its own (reference :185-233)."""

        def overlap(c1, e1, c2, e2):
            return torch.clamp(torch.min(c1 + e1 / 2, c2 + e2 / 2) - torch.max(c1 - e1 / 2, c2 - e2 / 2), min=1e-3)

        ix = overlap(px, p_whl[:, 0:1], tx, t_whl[:, 0:1])
        iy = overlap(py, p_whl[:, 2:3], ty, t_whl[:, 2:3])
        iz = overlap(pz, p_whl[:, 1:2], tz, t_whl[:, 1:2])
        inter = ix * iy * iz
        vol_p = torch.clamp(p_whl[:, 0:1] * p_whl[:, 2:3] * p_whl[:, 1:2], min=1e-3)
        vol_t = torch.clamp(t_whl[:, 0:1] * t_whl[:, 2:3] * t_whl[:, 1:2], min=1e-3)
        return inter / (vol_p + vol_t - inter)

    def _get_iou_loss(self, batch_preds, batch_targets, proposal_inds, pos_mask, stride, voxel_size, num_pos=None):
        pred = _transpose_and_gather_feat(batch_preds, proposal_inds)           # (B, K, n_code + 1)

        def decode(t):
            x = (t[..., 0:1] * stride * voxel_size[0]).reshape(-1, 1)
            y = (t[..., 1:2] * stride * voxel_size[1]).reshape(-1, 1)
            # exp(clamp(x, max=10)) == exp(x) wherever the outer clamp (max=30 < e^10) does not cut it off, values and
            # gradients alike; but it cannot overflow.  The reference only puts this branch into the graph when
            # loc_loss < 1 (a Python `if` on .item(), :131-133); the sync-free torch.where below always builds it, and a
            # zero gradient flowing into an exp() that overflowed to inf would be 0 * inf = NaN for the whole step.
            whl = torch.clamp(torch.exp(t[..., 3:6].clamp(max=10.0)).reshape(-1, 3), min=0.001, max=30)
            rot = torch.atan2(t[..., 6], t[..., 7]).reshape(-1, 1)
            z = t[..., 2].reshape(-1, 1)
            return x, y, z, whl, rot

        tx, ty, tz, t_whl, t_rot = decode(batch_targets)
        px, py, pz, p_whl, p_rot = decode(pred)
        iou = self._get_3d_iou(tx, ty, t_whl, tz, px, py, p_whl, pz)
        flat_mask = pos_mask.flatten()
        iou_pos = torch.clamp(iou[:, 0], 0, 1)
        if num_pos is None:
            num_pos = reduce_mean(pos_mask.float().sum())
        # == (1 - iou[pos]).sum() / max(1, num_pos)
        iou_loss = ((1 - iou_pos) * flat_mask.float()).sum() / torch.clamp(num_pos, min=1)

        target_box = torch.cat([tx, ty, tz, t_whl, t_rot], dim=-1)
        pred_box = torch.cat([px, py, pz, p_whl, p_rot], dim=-1).detach()
        bev_iou = box_utils.boxes3d_nearest_bev_iou_paired(target_box, pred_box)   # == diag of the N^2 matrix
        tar_iou_pred = 2 * (bev_iou.reshape(*pos_mask.shape, 1) - 0.5)
        iou_aware_loss = self.crit_iou_aware(batch_preds[:, -1:, :, :], pos_mask, proposal_inds, tar_iou_pred,
                                             num_pos)
        return iou_loss, iou_aware_loss.sum()
