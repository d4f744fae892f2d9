#!/usr/bin/env python
"""Per-phase cycles of the conv_gemm epilogue warps (study build of the library):
  nvcc ... -DMSPI_GEMM_STUDY -c mspi_b200/csrc/conv_gemm.cu ; link as mspi_b200/lib/libmspi_b200_study.so
  MSPI_LIB=mspi_b200/lib/libmspi_b200_study.so python tools/prof_gemm_phases.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import _lib, ops  # noqa: E402
from mspi_b200.ops import ACT_GELU, ACT_NONE, ACT_RELU, Act  # noqa: E402


def main():
    nf = 512
    bf, f32 = torch.bfloat16, torch.float32
    cases = [
        ("s2.fc1.gelu", (nf, 1, 14, 24), 384, 1536, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("s2.fc1.none", (nf, 1, 14, 24), 384, 1536, (1, 1, 1), (0, 0, 0), ACT_NONE, False, bf, bf),
        ("s3.fc1.gelu", (nf, 1, 7, 12), 768, 3072, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("s2.fc2.res", (nf, 1, 14, 24), 1536, 384, (1, 1, 1), (0, 0, 0), ACT_NONE, True, bf, bf),
        ("s0.fc1.gelu", (nf, 1, 56, 96), 96, 384, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("base1.3.conv_s", (nf // 16, 8, 56, 96), 64, 192, (1, 3, 3), (0, 1, 1), ACT_RELU, False, bf, bf),
        ("mixed.entry N=176", (nf // 16, 8, 28, 48), 192, 176, (1, 1, 1), (0, 0, 0), ACT_RELU, False, bf, bf),
        ("readout.1.tf32", (nf // 16, 4, 56, 96), 192, 192, (3, 3, 3), (1, 1, 1), ACT_RELU, False, f32, f32),
    ]
    lib = _lib.load()
    out = (C.c_uint64 * 8)()
    for name, shape, cin, cout, k, pad, act, res, dt, odt in cases:
        n, t, h, w = shape
        x = Act((torch.randn(n, t, h, w, cin, device="cuda")).to(dt))
        wgt = torch.randn(cout, cin, *k) / (cin * k[0] * k[1] * k[2]) ** 0.5
        conv = ops.Conv(wgt, None, torch.zeros(cout), pad=pad, act=act, dtype=dt, res_after_act=res, name=name)
        y = Act(torch.empty(n, t, h, w, cout, device="cuda", dtype=odt))
        r = Act(torch.randn(n, t, h, w, cout, device="cuda").to(odt)) if res else None
        run = conv.plan(x, y, r)
        run()
        lib.mspi_debug_gemm_epilogue_cycles(None, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        lib.mspi_debug_gemm_epilogue_cycles(C.cast(out, C.c_void_p), 1)
        ch, tiles = max(1, out[5]), max(1, out[6])
        nw = 16 * 148   # epilogue warps of a full grid
        print(f"{name:18s} {e0.elapsed_time(e1):.3f} ms  bn={conv.bn}  per warp: loop {out[7] / nw:9.0f} cyc | per tile: wait-acc {out[0] / tiles:7.0f} | "
              f"per chunk: tmem-ld {out[1] / ch:6.0f}  math {out[2] / ch:6.0f}  stage-free {out[3] / ch:6.0f}  stage+store {out[4] / ch:6.0f}  "
              f"(chunks/tile/warp {ch / tiles:.2f})", flush=True)


if __name__ == "__main__":
    main()
