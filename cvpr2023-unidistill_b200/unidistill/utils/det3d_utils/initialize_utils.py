import torch.nn as nn


def kaiming_init(module, a=0, mode="fan_out", nonlinearity="relu", bias=0, distribution="normal"):
    """Kaiming initialisation of a conv / linear layer (reference utils/det3d_utils/initialize_utils.py:6-19)."""
    if distribution not in ("uniform", "normal"):
        raise ValueError(distribution)
    init = nn.init.kaiming_uniform_ if distribution == "uniform" else nn.init.kaiming_normal_
    init(module.weight, a=a, mode=mode, nonlinearity=nonlinearity)
    if getattr(module, "bias", None) is not None:
        nn.init.constant_(module.bias, bias)
