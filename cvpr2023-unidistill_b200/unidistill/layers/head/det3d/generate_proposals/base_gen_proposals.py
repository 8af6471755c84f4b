class BaseGenProposals(object):
    """Interface of the proposal (decode + NMS) layers (reference generate_proposals/base_gen_proposals.py)."""

    training = True

    def generate_predicted_boxes(self, forward_ret_dict, data_dict):
        raise NotImplementedError
