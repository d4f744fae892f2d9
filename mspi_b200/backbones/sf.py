"""SlowFast 4x16 R50 video encoder (parameter container).  Mirrors backbones/sf.py:162-388 of the reference
(PySlowFast SlowFast built from configs/SLOWFAST_4x16_R50.yaml): same state_dict keys (`s1.pathway{0,1}_stem.*`,
`s{1..4}_fuse.{conv_f2s,bn}`, `s{2..5}.pathway{0,1}_res{i}.*`).  Forward arithmetic:
mspi_b200.engine.ForwardPlan.slowfast."""
import pickle

import torch

from ..params import ParamNode, conv_bn
from .X3D import declare_res_block

DEPTHS = (3, 4, 6, 3)                                  # ResNet-50
TK_SLOW, TK_FAST = (1, 1, 3, 3), (3, 3, 3, 3)          # _TEMPORAL_KERNEL_BASIS["slowfast"], sf.py:31-100
SLOW_FRAMES = (0, 4, 12, -1)                           # model_utils.py:523
ALPHA, BETA_INV, FUSION_RATIO, FUSION_K = 4, 8, 2, 5   # SLOWFAST_4x16_R50.yaml


class SlowFast(ParamNode):
    embeds = (320, 640, 1280, 2048)

    def __init__(self, path_to_config=None):
        super().__init__()
        conv_bn(self, "s1.pathway0_stem.conv", "s1.pathway0_stem.bn", 64, 3, (1, 7, 7))
        conv_bn(self, "s1.pathway1_stem.conv", "s1.pathway1_stem.bn", 8, 3, (5, 7, 7))
        conv_bn(self, "s1_fuse.conv_f2s", "s1_fuse.bn", 16, 8, (FUSION_K, 1, 1))
        cs, cf = 64 + 16, 8
        for si, depth in enumerate(DEPTHS):
            outs, outf = 256 * 2 ** si, 32 * 2 ** si
            ins, inf_ = 64 * 2 ** si, 8 * 2 ** si
            for pw, (cin, cout, inner, tk) in enumerate(((cs, outs, ins, TK_SLOW[si]), (cf, outf, inf_, TK_FAST[si]))):
                c = cin
                for i in range(depth):
                    declare_res_block(self, f"s{si + 2}.pathway{pw}_res{i}", c, cout, inner, tk, (1, 3, 3), False, i == 0)
                    c = cout
            if si < 3:
                conv_bn(self, f"s{si + 2}_fuse.conv_f2s", f"s{si + 2}_fuse.bn", FUSION_RATIO * outf, outf, (FUSION_K, 1, 1))
            cs, cf = outs + (FUSION_RATIO * outf if si < 3 else 0), outf

    def load_weight(self, path):
        """The reference converts a Caffe2 pickle (`blobs`) through PySlowFast's name table (checkpoint.py:226-330);
        here a PyTorch state_dict (optionally under 'model_state') or an empty-blob pickle is accepted, anything else is
        refused loudly rather than silently mis-mapped."""
        with open(path, "rb") as f:
            head = f.read(2)
        if head[:1] == b"\x80" and head != b"PK":
            with open(path, "rb") as f:
                blob = pickle.load(f, encoding="latin1")
            if isinstance(blob, dict) and "blobs" in blob:
                if blob["blobs"]:
                    raise NotImplementedError("Caffe2 -> PyTorch key conversion is out of scope; convert the checkpoint "
                                              "with PySlowFast and load the resulting state_dict")
                return
        sd = torch.load(path, map_location="cpu")
        self.load_state_dict(sd.get("model_state", sd), strict=False)
