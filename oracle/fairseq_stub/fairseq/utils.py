"""fairseq.utils subset: index_put, get_activation_fn (reference modules/module.py:9,20)."""
import torch
import torch.nn.functional as F


def index_put(tensor, indices, value):
    tensor[indices] = value
    return tensor


def gelu(x):
    return F.gelu(x.float()).type_as(x)


def get_activation_fn(activation):
    if activation == "relu":
        return F.relu
    if activation == "gelu":
        return gelu
    if activation == "tanh":
        return torch.tanh
    if activation == "linear":
        return lambda x: x
    raise RuntimeError("--activation-fn {} not supported".format(activation))
