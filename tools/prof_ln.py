"""Micro-benchmark of the row LayerNorm kernel on the ConvNeXt shapes."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import ops  # noqa: E402

nf = 512
for name, rows, c, idt, odt in (("stem_1 f32->bf16 C96", nf * 56 * 96, 96, torch.float32, torch.bfloat16),
                                ("ds1.0 bf16->bf16 C96", nf * 56 * 96, 96, torch.bfloat16, torch.bfloat16),
                                ("ds2.0 bf16->bf16 C192", nf * 28 * 48, 192, torch.bfloat16, torch.bfloat16),
                                ("ds3.0 bf16->bf16 C384", nf * 14 * 24, 384, torch.bfloat16, torch.bfloat16),
                                ("s3 LN bf16 C768 in place", nf * 7 * 12, 768, torch.bfloat16, torch.bfloat16)):
    x = torch.randn(rows, c, device="cuda").to(idt)
    y = torch.empty(rows, c, device="cuda", dtype=odt)
    run = ops.layernorm(x, y, rows, c, torch.ones(c), torch.zeros(c), 1e-6)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    byt = x.numel() * x.element_size() + y.numel() * y.element_size()
    print(f"{name:28s} {ms:7.3f} ms  {byt / ms / 1e6:7.1f} GB/s  {rows / ms / 1e3:8.1f} Mrows/s", flush=True)
