#!/usr/bin/env python
"""Micro-benchmark of the depthwise 7x7 + LayerNorm kernel on the ConvNeXt / decoder shapes."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import ops  # noqa: E402
from mspi_b200.ops import Act  # noqa: E402


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    bf, f32 = torch.bfloat16, torch.float32
    cases = [("s0 C96 56x96", nf, 56, 96, 96, bf, bf, (7, 7)), ("s1 C192 28x48", nf, 28, 48, 192, bf, bf, (7, 7)),
             ("s2 C384 14x24", nf, 14, 24, 384, bf, bf, (7, 7)), ("s3 C768 7x12", nf, 7, 12, 768, bf, bf, (7, 7)),
             ("lat0 C192 f32 56x96", nf // 4, 56, 96, 192, f32, f32, (7, 7)),
             ("lat0.t C192 f32", nf // 16, 56, 96, 192, f32, f32, (7, 1, 1))]
    for name, n, h, w, c, dt, odt, k in cases:
        if len(k) == 3:
            x = Act(torch.randn(n, 4, h, w, c, device="cuda").to(dt))
            y = Act(torch.empty(n, 4, h, w, c, device="cuda", dtype=odt))
            wgt = torch.randn(c, 1, 7, 1, 1) * 0.1
            run = ops.dwconv_ln(x, y, wgt, torch.zeros(c))
        else:
            x = Act(torch.randn(n, 1, h, w, c, device="cuda").to(dt))
            y = Act(torch.empty(n, 1, h, w, c, device="cuda", dtype=odt))
            wgt = torch.randn(c, 1, 7, 7) * 0.1
            run = ops.dwconv_ln(x, y, wgt, torch.zeros(c), torch.ones(c), torch.zeros(c), 1e-6)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        taps = 49 if len(k) == 2 else 7
        fma = x.buf.numel() * taps
        byt = x.buf.numel() * x.buf.element_size() + y.buf.numel() * y.buf.element_size()
        print(f"{name:22s} {ms:7.3f} ms  {fma / ms / 1e9:7.2f} TFMA/s  {byt / ms / 1e6:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
