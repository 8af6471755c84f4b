"""IouAwareGenProposals -- constructor-compatible holder of the decode/NMS configuration.

The decode + rotated-BEV NMS post-processing (reference generate_proposals/iou_aware_gen_proposals.py:141-247,
backed by the binary-only iou3d_nms_cuda extension) only runs in eval/predict mode and is OUT OF SCOPE of
the training hot path (SURVEY 2.1 #14, 8f-2 "next").  The head reads `voxel_size` from this object during
training (center_head_iou_aware.py:87), which is provided.
"""
from .base_gen_proposals import BaseGenProposals


class IouAwareGenProposals(BaseGenProposals):
    def __init__(self, dataset_name, class_names, post_center_limit_range, score_threshold, pc_range,
                 out_size_factor, voxel_size, no_log, iou_aware_list, nms_iou_threshold_train,
                 nms_pre_max_size_train, nms_post_max_size_train, nms_iou_threshold_test,
                 nms_pre_max_size_test, nms_post_max_size_test):
        self.dataset_name = dataset_name
        self.class_names = class_names
        self.post_center_limit_range = post_center_limit_range
        self.score_threshold = score_threshold
        self.pc_range = pc_range
        self.out_size_factor = out_size_factor
        self.voxel_size = voxel_size
        self.no_log = no_log
        self.iou_aware_list = iou_aware_list
        self.nms_cfg = dict(
            train=(nms_iou_threshold_train, nms_pre_max_size_train, nms_post_max_size_train),
            test=(nms_iou_threshold_test, nms_pre_max_size_test, nms_post_max_size_test))
        self.training = True

    def generate_predicted_boxes(self, forward_ret_dict, data_dict):
        raise NotImplementedError(
            "eval-time proposal decoding / rotated NMS is outside the B200 training hot path "
            "(SURVEY.md 8f-2); run evaluation with the reference's post-processing")
