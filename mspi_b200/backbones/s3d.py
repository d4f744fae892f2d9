"""S3D video encoder (parameter container).  Mirrors backbones/s3d.py:379-426 of the reference:
same state_dict keys, `load_weight(path)` with the same FileNotFoundError behaviour.  The forward
arithmetic is mspi_b200.engine.ForwardPlan.s3d (tcgen05 implicit-GEMM convs + pooling kernels)."""
import os

import torch

from ..params import ParamNode, conv_bn

# (cin, (b0, b1a, b1, b2a, b2, b3)) of Mixed_3b..5c, s3d.py:118-376
MIXED = {
    "base2.0": (192, (64, 96, 128, 16, 32, 32)),
    "base2.1": (256, (128, 128, 192, 32, 96, 64)),
    "base3.0": (480, (192, 96, 208, 16, 48, 64)),
    "base3.1": (512, (160, 112, 224, 24, 64, 64)),
    "base3.2": (512, (128, 128, 256, 24, 64, 64)),
    "base3.3": (512, (112, 144, 288, 32, 64, 64)),
    "base3.4": (528, (256, 160, 320, 32, 128, 128)),
    "base4.0": (832, (256, 160, 320, 32, 128, 128)),
    "base4.1": (832, (384, 192, 384, 48, 128, 128)),
}


def declare_basic(node, p, cin, cout, k=(1, 1, 1)):
    conv_bn(node, p + ".conv", p + ".bn", cout, cin, k)


def declare_sep(node, p, cin, cout, k):
    conv_bn(node, p + ".conv_s", p + ".bn_s", cout, cin, (1, k, k))
    conv_bn(node, p + ".conv_t", p + ".bn_t", cout, cout, (k, 1, 1))


def declare_mixed(node, p, cin, plan):
    b0, b1a, b1, b2a, b2, b3 = plan
    declare_basic(node, p + ".branch0.0", cin, b0)
    declare_basic(node, p + ".branch1.0", cin, b1a)
    declare_sep(node, p + ".branch1.1", b1a, b1, 3)
    declare_basic(node, p + ".branch2.0", cin, b2a)
    declare_sep(node, p + ".branch2.1", b2a, b2, 3)
    declare_basic(node, p + ".branch3.1", cin, b3)


class S3D_features_only(ParamNode):
    embeds = (192, 480, 832, 1024)

    def __init__(self, pool: int = 1):
        super().__init__()
        self.pool = pool
        declare_sep(self, "base1.0", 3, 64, 7)
        declare_basic(self, "base1.2", 64, 64)
        declare_sep(self, "base1.3", 64, 192, 3)
        for name, (cin, plan) in MIXED.items():
            declare_mixed(self, name, cin, plan)

    def load_weight(self, weight_path):
        if os.path.exists(weight_path):
            self.load_state_dict(torch.load(weight_path, map_location="cpu"))
            print("S3D Weight Loaded!")
        else:
            raise FileNotFoundError('S3D pretrained weight file ?')
