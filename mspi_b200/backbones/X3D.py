"""X3D-L video encoder (parameter container).  Mirrors backbones/X3D.py:111-250 of the reference (PySlowFast X3D
built from configs/X3D_L.yaml): same state_dict keys (`s1.pathway0_stem.*`, `s{2..5}.pathway0_res{i}.{branch1,
branch1_bn,branch2.{a,a_bn,b,b_bn,se.fc1,se.fc2,c,c_bn}}`), `load_weight(path)` reading `['model_state']`.
The ~25 constants the reference pulls from the vendored yacs config are literals here.
Forward arithmetic: mspi_b200.engine.ForwardPlan.x3d."""
import torch

from ..params import ParamNode, conv_bn

DEPTHS = (5, 10, 25, 15)          # block_basis [1,2,5,3] x DEPTH_FACTOR 5.0 (X3D.py:157-163, X3D_L.yaml)
OUT = (24, 48, 96, 192)           # DIM_C1 12 x WIDTH_FACTOR 2.0, doubling per stage
INNER = (54, 108, 216, 432)       # int(BOTTLENECK_FACTOR 2.25 * out)


def se_width(dim_in: int, ratio: float = 0.0625, divisor: int = 8) -> int:
    """SE._round_width, SlowFast/resnet_helper.py:26-45"""
    w = dim_in * ratio
    out = max(divisor, int(w + divisor / 2) // divisor * divisor)
    if out < 0.9 * w:
        out += divisor
    return int(out)


def declare_res_block(node, q, cin, cout, inner, tk_a, k_b, depthwise, first, se_dim=0):
    if first:
        conv_bn(node, q + ".branch1", q + ".branch1_bn", cout, cin, (1, 1, 1))
    conv_bn(node, q + ".branch2.a", q + ".branch2.a_bn", inner, cin, (tk_a, 1, 1))
    conv_bn(node, q + ".branch2.b", q + ".branch2.b_bn", inner, 1 if depthwise else inner, k_b)
    if se_dim:
        conv_bn(node, q + ".branch2.se.fc1", None, se_dim, inner, (1, 1, 1), bias=True)
        conv_bn(node, q + ".branch2.se.fc2", None, inner, se_dim, (1, 1, 1), bias=True)
    conv_bn(node, q + ".branch2.c", q + ".branch2.c_bn", cout, inner, (1, 1, 1))


class X3D(ParamNode):
    embeds = OUT

    def __init__(self, path_to_config=None, features_only=True):
        super().__init__()
        conv_bn(self, "s1.pathway0_stem.conv_xy", None, 24, 3, (1, 3, 3))
        conv_bn(self, "s1.pathway0_stem.conv", "s1.pathway0_stem.bn", 24, 1, (5, 1, 1))
        cin = 24
        for si, (depth, cout, inner) in enumerate(zip(DEPTHS, OUT, INNER)):
            for i in range(depth):
                declare_res_block(self, f"s{si + 2}.pathway0_res{i}", cin, cout, inner, 1, (3, 3, 3), True, i == 0,
                                  se_width(inner) if (i + 1) % 2 == 1 else 0)
                cin = cout

    def load_weight(self, path):
        self.load_state_dict(torch.load(path, map_location="cpu")['model_state'], strict=False)
        print("LOAD!!!")
