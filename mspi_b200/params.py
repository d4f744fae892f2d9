"""Parameter containers.

The reference's released checkpoints are plain ``state_dict``s whose keys come from its module
tree (model/model_utils.py:388-514).  The product keeps those names — `load_state_dict`,
`named_parameters()` prefixes (`audnet.`, `image_encoder.` are what train.py:151-155 freezes) and
attribute access all work — but the modules here only *hold* parameters: arithmetic happens in the
CUDA kernels driven by mspi_b200.engine.ForwardPlan, never in torch.nn forward passes.
"""
from __future__ import annotations

import math
from typing import Iterable, Tuple

import torch
import torch.nn as nn


class ParamNode(nn.Module):
    """A parameter-only module; children and tensors are registered by dotted name."""

    def put(self, dotted: str, tensor: torch.Tensor, buffer: bool = False):
        node = self
        parts = dotted.split(".")
        for part in parts[:-1]:
            if part not in node._modules:
                node.add_module(part, ParamNode())
            node = node._modules[part]
        if buffer:
            node.register_buffer(parts[-1], tensor)
        else:
            node.register_parameter(parts[-1], nn.Parameter(tensor))

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("mspi_b200 parameter containers have no PyTorch forward; the model's forward runs the "
                           "sm_100a kernels through mspi_b200.engine.ForwardPlan")


# ---- initialisers matching the reference's construction-time distributions --------------------
def _default_conv(shape, gen=None):
    """nn.Conv*/nn.Linear default: kaiming_uniform_(a=sqrt(5)) == U(+-1/sqrt(fan_in))."""
    fan_in = int(math.prod(shape[1:]))
    bound = 1.0 / math.sqrt(fan_in)
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _trunc_normal(shape, std=0.02, gen=None):
    t = torch.empty(shape)
    return nn.init.trunc_normal_(t, std=std, generator=gen)


def conv_bn(node: ParamNode, conv: str, bn: str, cout: int, cin: int, k: Tuple[int, ...], bias=False, init="default"):
    shape = (cout, cin) + tuple(k)
    if init == "kaiming_out":  # resnet.py:92-94
        fan_out = cout * int(math.prod(k))
        w = torch.randn(shape) * math.sqrt(2.0 / fan_out)
    elif init == "trunc":
        w = _trunc_normal(shape)
    else:
        w = _default_conv(shape)
    node.put(conv + ".weight", w)
    if bias:
        fan_in = cin * int(math.prod(k))
        node.put(conv + ".bias", torch.zeros(cout) if init == "trunc" else (torch.rand(cout) * 2 - 1) / math.sqrt(fan_in))
    if bn:
        node.put(bn + ".weight", torch.ones(cout) if init != "kaiming_out" else 1 + 0.02 * torch.randn(cout))
        node.put(bn + ".bias", torch.zeros(cout))
        node.put(bn + ".running_mean", torch.zeros(cout), buffer=True)
        node.put(bn + ".running_var", torch.ones(cout), buffer=True)
        node.put(bn + ".num_batches_tracked", torch.zeros((), dtype=torch.long), buffer=True)


def linear(node: ParamNode, name: str, cout: int, cin: int, bias=True, init="default"):
    if init == "xavier":  # SyncBlock._init_weights, model_utils.py:241-245
        bound = math.sqrt(6.0 / (cin + cout))
        w = (torch.rand(cout, cin) * 2 - 1) * bound
    elif init == "trunc":
        w = _trunc_normal((cout, cin))
    else:
        w = _default_conv((cout, cin))
    node.put(name + ".weight", w)
    if bias:
        node.put(name + ".bias", torch.zeros(cout) if init != "default" else (torch.rand(cout) * 2 - 1) / math.sqrt(cin))


def layer_norm(node: ParamNode, name: str, c: int):
    node.put(name + ".weight", torch.ones(c))
    node.put(name + ".bias", torch.zeros(c))
