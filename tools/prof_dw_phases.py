#!/usr/bin/env python
"""Where a block of the depthwise 7x7 + LayerNorm kernel spends its life (MSPI_DW_DEBUG=1: instrumented instance).
Prints, per ConvNeXt stage, the mean cycles thread 0 of a block waits for its tile / runs the stencil / runs LayerNorm."""
import ctypes as C
import os
import sys

os.environ["MSPI_DW_DEBUG"] = "1"
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import _lib, ops  # noqa: E402
from mspi_b200.ops import Act  # noqa: E402


def main():
    nf = 512
    lib = _lib.load()
    out = (C.c_uint64 * 8)()
    for name, h, w, c in (("s0 C96 56x96", 56, 96, 96), ("s1 C192 28x48", 28, 48, 192), ("s2 C384 14x24", 14, 24, 384)):
        x = Act(torch.randn(nf, 1, h, w, c, device="cuda").to(torch.bfloat16))
        y = Act(torch.empty(nf, 1, h, w, c, device="cuda", dtype=torch.bfloat16))
        run = ops.dwconv_ln(x, y, torch.randn(c, 1, 7, 7) * 0.1, torch.zeros(c), torch.ones(c), torch.zeros(c), 1e-6)
        run()
        lib.mspi_debug_dw_phase_cycles(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        lib.mspi_debug_dw_phase_cycles(C.cast(out, C.c_void_p), 1)
        n = max(1, out[3])
        print(f"{name:16s} {e0.elapsed_time(e1):.3f} ms  warps {out[3]}  per warp: tile wait {out[0] / n:7.0f}  stencil {out[1] / n:7.0f}  "
              f"barrier {out[4] / n:7.0f}  result store+barrier {out[5] / n:7.0f}  layernorm+stores {out[2] / n:7.0f} cycles", flush=True)


if __name__ == "__main__":
    main()
