"""Two eager training steps (B=2, 16x224x384) for ncu: python tools/prof_train.py [steps]"""
import contextlib, copy, io, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mspi_b200.config import cfg as base_cfg, select_motion_encoder
from mspi_b200.model.model_utils import AudioVisualSaliencyModel
from mspi_b200.train_engine import TrainPlan

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
torch.manual_seed(2023)
with contextlib.redirect_stdout(io.StringIO()):
    model = AudioVisualSaliencyModel(select_motion_encoder("s3d", copy.deepcopy(base_cfg)), load_pretrained=False)
B, T, H, W = 2, 16, 224, 384
plan = TrainPlan(model.state_dict(), B, T, H, W)
clips, audio, gt = torch.randn(B, 3, T, H, W, device="cuda"), torch.randn(B, 1, 257, 111, device="cuda"), torch.rand(B, H, W, device="cuda")
for _ in range(steps):
    out = plan.train_step(clips, audio, gt)
torch.cuda.synchronize()
print("loss", out.cpu().tolist(), "launches/step", plan.num_launches)
