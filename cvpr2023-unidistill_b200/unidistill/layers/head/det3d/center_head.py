"""CenterHead / SepHead -- CenterPoint detection head (reference layers/head/det3d/center_head.py:15-171,311-375).

Dense 3x3 convolutions stay on cuDNN; module tree and parameter names are the reference's
(`shared_conv.{0,1}`, `tasks.<t>.<head>.{0,1,3}`) so checkpoints load by key.
"""
import copy
import os
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

from unidistill.layers.losses import det3d as loss_utils
from unidistill.utils.det3d_utils.initialize_utils import kaiming_init


class Sequential(nn.Sequential):
    """nn.Sequential with the reference's `add(module, name=None)` convenience (center_head.py:266-308)."""

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)


class SepHead(nn.Module):
    """One small conv tower per regression / classification target of a task."""

    def __init__(self, in_channels, heads, head_conv=64, name="", final_kernel=1, bn=False, init_bias=-2.19,
                 directional_classifier=False, **kwargs):
        super().__init__(**kwargs)
        assert directional_classifier is False, "directional classifier is not supported (as in the reference)"
        self.heads = heads
        pad = final_kernel // 2
        for head, (classes, num_conv) in self.heads.items():
            tower = Sequential()
            for _ in range(num_conv - 1):
                tower.add(nn.Conv2d(in_channels, head_conv, kernel_size=final_kernel, stride=1, padding=pad,
                                    bias=True))
                if bn:
                    tower.add(nn.BatchNorm2d(head_conv))
                tower.add(nn.ReLU())
            tower.add(nn.Conv2d(head_conv, classes, kernel_size=final_kernel, stride=1, padding=pad, bias=True))
            if "hm" in head:
                tower[-1].bias.data.fill_(init_bias)
            else:
                for m in tower.modules():
                    if isinstance(m, nn.Conv2d):
                        kaiming_init(m)
            self.__setattr__(head, tower)

    def forward(self, x):
        return {head: self.__getattr__(head)(x) for head in self.heads}


# the 6 tasks x 7 targets = 42 towers of a CenterHead all read the same 64-channel map and have the same shape
# up to the class count of the last conv, so they can run as ONE 64 -> 42*64 convolution, ONE BatchNorm over 2688 channels
# (per-channel statistics: identical to 42 separate norms), ONE ReLU and ONE block-diagonal convolution 2688 -> 76 instead
# of 42 x (conv, BN, ReLU, conv).  Parameters stay where the reference keeps them
# (`tasks.<t>.<head>.{0,1,3}`), gradients flow back to them through the concatenations.  UD_BATCH_TOWERS=0 disables it.
BATCH_TOWERS = os.environ.get("UD_BATCH_TOWERS", "1") != "0"


_ROWS_CACHE = {}


class _ConvDgradAsFprop(torch.autograd.Function):
    """conv2d (stride 1, dilation 1) whose input gradient is computed as a FORWARD convolution of the output gradient with
    the flipped, transposed kernel.  For the towers' block-diagonal second layer (76 <- 2688 channels) cuDNN's dgrad pick is
    an indexed sm80 kernel at 1.17 ms on B200, while the same arithmetic as an fprop (2688 <- 76) runs on the sm100 implicit
    GEMM in ~0.3 ms (profiles/r03_step_timeline.txt).  Weight / bias gradients come from aten's convolution_backward."""

    @staticmethod
    def forward(ctx, x, w, b, padding):
        ctx.save_for_backward(x, w)
        ctx.padding = tuple(int(p) for p in padding)
        ctx.has_bias = b is not None
        return F.conv2d(x, w, b, stride=1, padding=ctx.padding)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        gx = gw = gb = None
        kh, kw = int(w.shape[2]), int(w.shape[3])
        if ctx.needs_input_grad[0]:
            wt = w.flip(2, 3).transpose(0, 1)
            gg = g
            if x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
                # the gradient must come out in the layout of x (NHWC in the trunk): the output gradient is the small tensor
                # here (76 channels), converting it costs microseconds; an NCHW result would cost a 700 MB layout pass
                gg = g.contiguous(memory_format=torch.channels_last)
                wt = wt.contiguous(memory_format=torch.channels_last)
            gx = F.conv2d(gg, wt, None, stride=1, padding=(kh - 1 - ctx.padding[0], kw - 1 - ctx.padding[1]))
        if ctx.needs_input_grad[1] or (ctx.has_bias and ctx.needs_input_grad[2]):
            _, gw, gb = torch.ops.aten.convolution_backward(
                g, x, w, [int(w.shape[0])] if ctx.has_bias else None, (1, 1), ctx.padding, (1, 1), False, (0, 0), 1,
                [False, True, ctx.has_bias])
            if not ctx.has_bias:
                gb = None
        return gx, gw, gb, None


def batched_towers(tasks, x):
    """tasks: the ModuleList of SepHeads of one CenterHead; x (B, C, H, W).  Returns [ {head: (B, classes, H, W)} per task ]
    == [task(x) for task in tasks], or None when the towers do not have the regular conv-BN-ReLU-conv shape."""
    towers, names = [], []
    for t, task in enumerate(tasks):
        for head in task.heads:
            tw = getattr(task, head)
            if len(tw) != 4 or not isinstance(tw[0], nn.Conv2d) or not isinstance(tw[3], nn.Conv2d) \
                    or not isinstance(tw[1], (nn.BatchNorm2d, nn.Identity)) or not isinstance(tw[2], nn.ReLU):
                return None
            towers.append(tw)
            names.append((t, head))
    c1, c2 = towers[0][0], towers[0][3]
    same = all(tw[0].weight.shape == c1.weight.shape and tw[0].padding == c1.padding and tw[0].stride == (1, 1)
               and tw[0].bias is not None and tw[3].kernel_size == c2.kernel_size and tw[3].padding == c2.padding
               and tw[3].in_channels == c2.in_channels and tw[3].bias is not None and type(tw[1]) is type(towers[0][1])
               for tw in towers)
    if not same:
        return None
    n, mid = len(towers), c1.out_channels
    # ---- layer 1: one convolution for all towers
    w1 = torch.cat([tw[0].weight for tw in towers], dim=0)
    b1 = torch.cat([tw[0].bias for tw in towers], dim=0)
    if w1.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last) and not x.is_contiguous():
        w1 = w1.contiguous(memory_format=torch.channels_last)
    frozen = isinstance(towers[0][1], nn.Identity) and not torch.is_grad_enabled() and x.is_cuda
    bns = [tw[1] for tw in towers]
    # A train-mode BatchNorm subtracts the batch mean, so the convolution bias in front of it cancels: BN(conv + b) ==
    # BN(conv) with the batch mean shifted by b.  Adding it anyway is a broadcast add over the (B, 42*64, H, W) tensor --
    # 704 MB read + written, 0.48 ms per step on B200 (profiles/r03_step_timeline.txt) -- so the bias is left out of the
    # tensor and put back where it is observable: the running mean (+ momentum * b).  Its gradient, which the reference
    # gets as round-off noise around zero, is exactly zero here (beta + 0 * b keeps the parameter in the autograd graph).
    absorb_bias = (isinstance(bns[0], nn.BatchNorm2d) and bns[0].training and bns[0].momentum is not None
                   and bns[0].track_running_stats and not frozen)
    if frozen:          # folded BatchNorm, no gradients: bias + ReLU in the convolution's epilogue
        y = torch.cudnn_convolution_relu(x, w1, b1, (1, 1), c1.padding, (1, 1), 1)
    else:
        y = F.conv2d(x, w1, None if absorb_bias else b1, stride=1, padding=c1.padding)
    # ---- one BatchNorm over all 42*64 channels (identity when the norms were folded into the convolutions)
    if isinstance(bns[0], nn.BatchNorm2d):
        training = bns[0].training
        gamma = torch.cat([bn.weight for bn in bns])
        beta = torch.cat([bn.bias for bn in bns])
        rm = torch.cat([bn.running_mean for bn in bns])
        rv = torch.cat([bn.running_var for bn in bns])
        if absorb_bias:
            beta = beta + 0.0 * b1
        y = F.batch_norm(y, rm, rv, gamma, beta, training, bns[0].momentum, bns[0].eps)
        if training:
            with torch.no_grad():                                   # running statistics back into the 42 modules
                if absorb_bias:                                     # out of place: autograd holds `rm` for the backward
                    rm = torch.add(rm, b1, alpha=float(bns[0].momentum))
                torch._foreach_copy_([bn.running_mean for bn in bns], list(rm.split(mid)))
                torch._foreach_copy_([bn.running_var for bn in bns], list(rv.split(mid)))
                torch._foreach_add_([bn.num_batches_tracked for bn in bns], 1)
    if not frozen:
        y = F.relu(y)
    # ---- layer 2: ONE dense convolution over all 42*64 channels with a block-diagonal weight (tower i's rows see only
    # tower i's 64 input channels, zeros elsewhere).  It spends 42x the FLOPs of the 42 small convolutions, but as one
    # tensor-core implicit GEMM (~0.4 PFLOP/s) it is still ~5x faster than they are, and -- unlike a groups=42
    # convolution, whose weight gradient cuDNN runs as 42 SIMT SGEMMs (13.5 ms per step measured on B200,
    # profiles/r02_head_batching.md) -- its backward stays on the tensor cores.  Only the diagonal blocks of the weight
    # gradient flow back to the parameters (index_put's backward gathers them).
    outs = [tw[3].out_channels for tw in towers]
    cin2 = c2.in_channels
    key = (str(x.device), cin2, tuple(outs))
    idx = _ROWS_CACHE.get(key)
    if idx is None:      # built once per layout (the first call must be outside a CUDA-graph capture: warm-up runs are)
        rows = torch.arange(sum(outs), device=x.device)
        tower_of_row = torch.tensor([i for i, c_ in enumerate(outs) for _ in range(c_)], device=x.device)
        cols = tower_of_row[:, None] * cin2 + torch.arange(cin2, device=x.device)[None, :]
        idx = (rows[:, None].expand_as(cols).contiguous(), cols.contiguous())
        _ROWS_CACHE[key] = idx
    w2_all = torch.cat([tw[3].weight for tw in towers], dim=0)              # (sum outs, 64, 3, 3)
    b2 = torch.cat([tw[3].bias for tw in towers], dim=0)
    w2 = w2_all.new_zeros((w2_all.shape[0], n * cin2) + tuple(w2_all.shape[2:])).index_put(idx, w2_all)
    if y.is_contiguous(memory_format=torch.channels_last) and not y.is_contiguous():
        w2 = w2.contiguous(memory_format=torch.channels_last)
    if torch.is_grad_enabled() and y.requires_grad:
        z = _ConvDgradAsFprop.apply(y, w2, b2, c2.padding)
    else:
        z = F.conv2d(y, w2, b2, stride=1, padding=c2.padding)
    starts = [0]
    for c_ in outs:
        starts.append(starts[-1] + c_)
    ret = [dict() for _ in tasks]
    for i, (t, head) in enumerate(names):
        ret[t][head] = z[:, starts[i]: starts[i + 1]]
    return [OrderedDict((h, d[h]) for h in task.heads) for task, d in zip(tasks, ret)]


class CenterHead(nn.Module):
    def __init__(self, dataset_name, tasks, target_assigner, proposal_layer, input_channels, grid_size,
                 point_cloud_range, code_weights, loc_weight, share_conv_channel, common_heads,
                 upsample_for_pedestrian=False, predict_boxes_when_training=False, mode="3d", init_bias=-2.19,
                 distill=False):
        super().__init__()
        self.in_channels = input_channels
        self.grid_size = grid_size
        self.point_cloud_range = point_cloud_range
        self.predict_boxes_when_training = predict_boxes_when_training
        self.num_classes = [len(t["class_names"]) for t in tasks]
        self.class_names = [t["class_names"] for t in tasks]
        self.code_weights = code_weights
        self.weight = loc_weight
        self.dataset = dataset_name
        self.box_n_dim = 9 if self.dataset == "nuscenes" else 7
        self.encode_background_as_zeros = True
        self.use_sigmoid_score = True
        self.no_log = False
        self.use_direction_classifier = False
        self.bev_only = mode == "bev"
        self.distill = distill
        self.shared_conv = nn.Sequential(
            nn.Conv2d(self.in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            nn.BatchNorm2d(share_conv_channel),
            nn.ReLU(inplace=True),
        )
        self.upsample_for_pedestrian = upsample_for_pedestrian
        if upsample_for_pedestrian:
            self.upsample_conv = nn.Sequential(
                nn.ConvTranspose2d(share_conv_channel, share_conv_channel, 2, stride=2, bias=False),
                nn.BatchNorm2d(share_conv_channel),
                nn.ReLU(),
            )
        self.common_heads = common_heads
        self.init_bias = init_bias
        self.tasks = nn.ModuleList()
        for num_cls in self.num_classes:
            heads = copy.deepcopy(dict(self.common_heads))
            heads.update(dict(hm=(num_cls, 2)))
            self.tasks.append(SepHead(share_conv_channel, OrderedDict(heads), bn=True, init_bias=init_bias,
                                      final_kernel=3, directional_classifier=False))
        self.target_assigner = target_assigner
        self.proposal_layer = proposal_layer
        self.build_losses()

    def train(self, mode: bool = True):
        super().train(mode=mode)
        self.proposal_layer.training = mode
        return self

    def eval(self):
        return self.train(False)

    def assign_targets(self, gt_boxes):
        return self.target_assigner.assign_targets(gt_boxes)

    def forward(self, spatial_features_2d, gt_boxes=None, targets=None):
        """`targets`: optional result of `assign_targets(gt_boxes)` computed by the caller (the assignment depends on
        the ground truth only; a distillation step shares one between student and teacher) -- same dict either way."""
        ret = {}
        x = self.shared_conv(spatial_features_2d)
        if self.upsample_for_pedestrian:
            x = self.upsample_conv(x)
        ret["multi_head_features"] = self.run_towers(x)
        if self.training or self.distill:
            ret.update(targets if targets is not None else self.assign_targets(gt_boxes))
            return ret
        return self.proposal_layer.generate_predicted_boxes(ret, {})

    def run_towers(self, x):
        """[task(x) for task in self.tasks]; on a CUDA device the 42 towers run batched (see batched_towers)."""
        if BATCH_TOWERS and x.is_cuda:
            out = batched_towers(self.tasks, x)
            if out is not None:
                return out
        return [task(x) for task in self.tasks]

    def build_losses(self):
        self.add_module("crit", loss_utils.CenterNetFocalLoss())
        self.add_module("crit_reg", loss_utils.CenterNetRegLoss())

    def _code_weights_on(self, like):
        """code_weights as a tensor next to `like`, built once per device (a host->device copy per call would also make
        get_loss impossible to capture into a CUDA graph)."""
        cache = self.__dict__.setdefault("_code_w_cache", {})
        key = (like.device, like.dtype)
        if key not in cache:
            cache[key] = like.new_tensor(self.code_weights)
        return cache[key]

    def _sigmoid(self, x):
        return torch.clamp(x.sigmoid(), min=1e-4, max=1 - 1e-4)

    def _box_codes(self, pred_dict, with_iou=False):
        names = ["reg", "height", "dim", "rot"] + (["vel"] if self.dataset == "nuscenes" else [])
        if with_iou:
            names.append("iou")
        return torch.cat([pred_dict[n] for n in names], dim=1).contiguous()

    def get_loss(self, forward_ret_dict):
        """hm focal loss + weighted L1 box loss per task (reference center_head.py:171-263).
        tb_dict values are 0-dim tensors (no host sync); call float() on them to log."""
        tb_dict = {}
        losses = []
        forward_ret_dict["pred_box_encoding"] = {}
        code_w = None
        for task_id, pred_dict in enumerate(forward_ret_dict["multi_head_features"]):
            pred_dict["hm"] = self._sigmoid(pred_dict["hm"])
            hm_loss = self.crit(pred_dict["hm"], forward_ret_dict["heatmap"][task_id])
            target = forward_ret_dict["box_encoding"][task_id]
            pred = self._box_codes(pred_dict)
            forward_ret_dict["pred_box_encoding"][task_id] = pred
            box_loss = self.crit_reg(pred, forward_ret_dict["mask"][task_id], forward_ret_dict["ind"][task_id],
                                     target)
            if code_w is None:
                code_w = self._code_weights_on(box_loss)
            loc_loss = (box_loss * code_w).sum()
            loss = hm_loss + self.weight * loc_loss
            key = "task_%d/" % task_id
            tb_dict.update({key + "loss": loss.detach(), key + "hm_loss": hm_loss.detach(),
                            key + "loc_loss": loc_loss.detach(),
                            key + "num_positive": forward_ret_dict["mask"][task_id].float().sum()})
            losses.append(loss)
        return sum(losses), tb_dict
