#!/usr/bin/env python
"""Micro-benchmark of the implicit-GEMM kernel on the layer shapes that dominate the step
(ConvNeXt stage-0/2 MLP, S3D base1.3, readout.1).  Prints CUDA-event times; run the same command
under `ncu --set full -k regex:conv_gemm` for the stall breakdown (profiles/).

  python tools/prof_gemm.py [--frames 512] [--only NAME] [--reps 5]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import ops  # noqa: E402
from mspi_b200.ops import ACT_GELU, ACT_NONE, ACT_RELU, Act  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--only", default=None)
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    nf = args.frames
    bf, f32 = torch.bfloat16, torch.float32
    # name, (n,t,h,w), cin, cout, kernel, pad, act, residual, in dtype, out dtype
    cases = [
        ("s0.fc1.gelu", (nf, 1, 56, 96), 96, 384, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("s0.fc1.none", (nf, 1, 56, 96), 96, 384, (1, 1, 1), (0, 0, 0), ACT_NONE, False, bf, bf),
        ("s0.fc1.relu", (nf, 1, 56, 96), 96, 384, (1, 1, 1), (0, 0, 0), ACT_RELU, False, bf, bf),
        ("s0.fc2.res", (nf, 1, 56, 96), 384, 96, (1, 1, 1), (0, 0, 0), ACT_NONE, True, bf, bf),
        ("s0.fc2.nores", (nf, 1, 56, 96), 384, 96, (1, 1, 1), (0, 0, 0), ACT_NONE, False, bf, bf),
        ("s2.fc1.gelu", (nf, 1, 14, 24), 384, 1536, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("s2.fc1.none", (nf, 1, 14, 24), 384, 1536, (1, 1, 1), (0, 0, 0), ACT_NONE, False, bf, bf),
        ("s2.fc1.relu", (nf, 1, 14, 24), 384, 1536, (1, 1, 1), (0, 0, 0), ACT_RELU, False, bf, bf),
        ("s3.fc1.gelu", (nf, 1, 7, 12), 768, 3072, (1, 1, 1), (0, 0, 0), ACT_GELU, False, bf, bf),
        ("s3.fc1.none", (nf, 1, 7, 12), 768, 3072, (1, 1, 1), (0, 0, 0), ACT_NONE, False, bf, bf),
        ("s2.fc2.res", (nf, 1, 14, 24), 1536, 384, (1, 1, 1), (0, 0, 0), ACT_NONE, True, bf, bf),
        ("base1.3.conv_s", (nf // 16, 8, 56, 96), 64, 192, (1, 3, 3), (0, 1, 1), ACT_RELU, False, bf, bf),
        ("readout.1.tf32", (nf // 16, 4, 56, 96), 192, 192, (3, 3, 3), (1, 1, 1), ACT_RELU, False, f32, f32),
        ("lat0.pw1.tf32", (nf // 16, 4, 56, 96), 192, 768, (1, 1, 1), (0, 0, 0), ACT_GELU, False, f32, f32),
    ]
    for name, shape, cin, cout, k, pad, act, res, dt, odt in cases:
        if args.only and args.only != name:
            continue
        n, t, h, w = shape
        x = Act((torch.randn(n, t, h, w, cin, device="cuda") * 1.0).to(dt))
        wgt = torch.randn(cout, cin, *k) / (cin * k[0] * k[1] * k[2]) ** 0.5
        conv = ops.Conv(wgt, None, torch.zeros(cout), pad=pad, act=act, dtype=dt, res_after_act=res, name=name)
        y = Act(torch.empty(n, t, h, w, cout, device="cuda", dtype=odt))
        r = Act(torch.randn(n, t, h, w, cout, device="cuda").to(odt)) if res else None
        run = conv.plan(x, y, r)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.reps
        fl = conv.flops(x)
        es_i, es_o = x.buf.element_size(), y.buf.element_size()
        byt = x.buf.numel() * es_i + y.buf.numel() * es_o * (2 if res else 1)
        print(f"{name:18s} {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TF/s  {byt / ms / 1e6:8.1f} GB/s  bn={conv.bn} "
              f"elems/s={y.buf.numel() / ms / 1e6:.0f}G", flush=True)


def fused():
    """Fused MLP vs fc1 + fc2 on the ConvNeXt stage-0 / stage-1 shapes."""
    import torch.nn.functional as F  # noqa: F401
    for name, nf, h, w, c in (("s0.mlp.fused C96", 512, 56, 96, 96), ("s1.mlp.fused C192", 512, 28, 48, 192)):
        m = nf * h * w
        x = Act(torch.randn(1, 1, 1, m, c, device="cuda").to(torch.bfloat16))
        r = Act(torch.randn(1, 1, 1, m, c, device="cuda").to(torch.bfloat16))
        y = Act(torch.empty(1, 1, 1, m, c, device="cuda", dtype=torch.bfloat16))
        run = ops.mlp_fused(x, y, r, torch.randn(4 * c, c) / c ** 0.5, torch.zeros(4 * c), torch.randn(c, 4 * c) / (4 * c) ** 0.5,
                            torch.zeros(c), torch.ones(c))
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
        print(f"{name:18s} {ms:8.3f} ms  {run.flops / ms / 1e9:8.1f} TF/s  {3 * m * c * 2 / ms / 1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    if "--fused" in sys.argv:
        fused()
    else:
        main()
