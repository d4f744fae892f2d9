import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mspi_b200 import ops
from mspi_b200.ops import Act
cin, cout, k, shape = 32, 32, (1, 3, 3), (2, 1, 64, 64)
n, t, h, w = shape
x = Act(torch.randn(n, t, h, w, cin, device="cuda"))
dy = Act(torch.randn(n, t, h, w, cout, device="cuda"))
conv = ops.Conv(torch.zeros(cout, cin, *k), None, None, pad=(0, 1, 1), dtype=torch.float32, name="dbg")
dw = torch.zeros(cout, cin, *k, device="cuda")
run = conv.wgrad_plan(x, dy, dw)
d = run.desc
print("box", list(d.box), "o_dims", list(d.o_dims))
run()
torch.cuda.synchronize()
print("ok", float(dw.abs().sum()))
