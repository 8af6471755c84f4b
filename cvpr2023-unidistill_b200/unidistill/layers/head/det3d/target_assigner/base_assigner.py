class BaseAssigner(object):
    """Interface of target assigners (reference target_assigner/base_assigner.py)."""

    def assign_targets(self, gt_boxes):
        raise NotImplementedError
