"""IouAwareGenProposals -- CenterPoint decoding whose NMS ranks candidates by the IoU-aware score
score^(1-a) * iou^a with iou = clamp(iou_head / 2 + 0.5, 0, 1) and a per-task exponent `iou_aware_list[task]`
(reference generate_proposals/iou_aware_gen_proposals.py:42-139).  The reported `scores` stay the class scores.
The head also reads `voxel_size` from this object during training (center_head_iou_aware.py:87)."""
import torch

from .centerpoint_gen_proposals import CenterPointGenProposals


class IouAwareGenProposals(CenterPointGenProposals):
    def __init__(self, dataset_name, class_names, post_center_limit_range, score_threshold, pc_range,
                 out_size_factor, voxel_size, no_log, nms_iou_threshold_train, nms_pre_max_size_train,
                 nms_post_max_size_train, nms_iou_threshold_test, nms_pre_max_size_test, nms_post_max_size_test,
                 iou_aware_list):
        super().__init__(dataset_name, class_names, post_center_limit_range, score_threshold, pc_range,
                         out_size_factor, voxel_size, no_log, nms_iou_threshold_train, nms_pre_max_size_train,
                         nms_post_max_size_train, nms_iou_threshold_test, nms_pre_max_size_test,
                         nms_post_max_size_test)
        self.iou_aware_list = iou_aware_list
        self.nms_cfg = dict(
            train=(nms_iou_threshold_train, nms_pre_max_size_train, nms_post_max_size_train),
            test=(nms_iou_threshold_test, nms_pre_max_size_test, nms_post_max_size_test))

    def _nms_score(self, scores, inds, task_id, iouhm):
        batch, K = scores.shape
        iouscore = self._transpose_and_gather_feat(iouhm, inds).view(batch, K)
        iouscore = torch.clamp(iouscore / 2 + 0.5, 0, 1)
        a = self.iou_aware_list[task_id]
        return (scores ** (1 - a)).mul(iouscore ** a)
