"""How does tcgen05 kind::tf32 reduce fp32 operands to tf32?  (truncation vs round-to-nearest)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mspi_b200 import ops
from mspi_b200.ops import Act
for name, v in (("1+2^-11+2^-13", 1 + 2.0 ** -11 + 2.0 ** -13), ("1+2^-11", 1 + 2.0 ** -11), ("1+3*2^-11", 1 + 3 * 2.0 ** -11),
                ("1+2^-10-2^-20", 1 + 2.0 ** -10 - 2.0 ** -20), ("-(1+2^-11+2^-13)", -(1 + 2.0 ** -11 + 2.0 ** -13))):
    x = Act(torch.full((1, 1, 1, 128, 32), 0.0, device="cuda"))
    x.buf[..., 0] = v
    w = torch.zeros(16, 32)
    w[0, 0] = 1.0
    w[1, 0] = v          # weight side rounding
    conv = ops.Conv(w, None, None, dtype=torch.float32, name="t")
    y = Act(torch.zeros(1, 1, 1, 128, 16, device="cuda"))
    conv.plan(x, y)()
    torch.cuda.synchronize()
    a, b = float(y.buf[0, 0, 0, 0, 0]), float(y.buf[0, 0, 0, 0, 1])
    print(f"{name:20s} x*1 = 1+{(abs(a) - 1) * 2 ** 10:.6f}*2^-10   x*x = {b!r}  (exact fp32 {v * v!r})")
