"""Debug aid: delta-weight probes of ops.stem_conv (which (kh,kw,ch) lands where)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mspi_b200 import ops
from mspi_b200.ops import Act

def probe(k, stride, pad, odt):
    b, t, h, w = 1, 2, 32, 64
    g = torch.Generator().manual_seed(1)
    clip = torch.randn(b, 3, t, h, w, generator=g).to(torch.bfloat16).float()
    frames = torch.zeros(b * t, h + 8, w + 8, 4, dtype=torch.bfloat16, device="cuda")
    ops.clip_to_padded({"clips": clip.cuda()}, "clips", frames, b, t, h, w)()
    torch.cuda.synchronize()
    taps = [(kh, kw, ch) for kh in range(k) for kw in range(k) for ch in range(3)]
    bad = []
    for base in range(0, len(taps), 64):
        sel = taps[base:base + 64]
        cout = 64
        wgt = torch.zeros(cout, 3, k, k)
        for i, (kh, kw, ch) in enumerate(sel):
            wgt[i, ch, kh, kw] = 1.0
        oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
        y = Act.empty(b, t, oh, ow, cout, dtype=odt)
        ops.stem_conv(frames, h, w, wgt, None, None, k, stride, pad, 0, y)()
        torch.cuda.synchronize()
        x2 = clip.permute(0, 2, 1, 3, 4).reshape(b * t, 3, h, w)
        ref = F.conv2d(x2, wgt, None, stride, pad).view(b, t, cout, oh, ow).permute(0, 2, 1, 3, 4)
        got = y.to_ncdhw().cpu()
        for i, tp in enumerate(sel):
            e = (got[:, i] - ref[:, i]).abs().max().item()
            if e > 1e-3:
                bad.append((tp, round(e, 3)))
    print(f"k={k} s={stride}: {len(bad)} bad taps of {len(taps)}", bad[:40])

probe(4, 4, 0, torch.float32)
probe(7, 2, 3, torch.bfloat16)
