"""Eager forwards at B=32, 16x224x384 for ncu: python tools/prof_infer.py [forwards] [s3d|x3dl|slowfast4x16]; prints launches per forward."""
import contextlib, copy, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mspi_b200 import _lib
from mspi_b200.config import cfg as base_cfg, select_motion_encoder
from mspi_b200.model.model_utils import AudioVisualSaliencyModel

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
enc = sys.argv[2] if len(sys.argv) > 2 else "s3d"
torch.manual_seed(2023)
with contextlib.redirect_stdout(io.StringIO()):
    model = AudioVisualSaliencyModel(select_motion_encoder(enc, copy.deepcopy(base_cfg)), load_pretrained=False).cuda().eval()
B = 32
clips, audio = torch.randn(B, 3, 16, 224, 384, device="cuda"), torch.randn(B, 1, 257, 111, device="cuda")
lib = _lib.load()
counts = []
for _ in range(n):
    l0 = lib.mspi_launch_count()
    out, loss = model(clips, audio)
    torch.cuda.synchronize()
    counts.append(int(lib.mspi_launch_count() - l0))
print("launches_per_forward", counts[-1], "first", counts[0])
