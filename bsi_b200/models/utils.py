import torch.nn as nn


def actfn_from_str(name: str):
    """Activation class by name (reference bsi/models/utils.py:4-12)."""
    return {"silu": nn.SiLU, "gelu": nn.GELU, "relu": nn.ReLU, "softplus": nn.Softplus, "tanh": nn.Tanh}[name]
