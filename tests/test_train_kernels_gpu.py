"""GPU parity tests of the training-step kernels (row a20) against PyTorch autograd on the CPU (fp32/fp64).

Each kernel is driven through the C ABI exactly as the training plan drives it.  Tolerances are stated at their use:
tf32 tensor-core paths round both operands to 10 mantissa bits (2^-11 relative per product, fp32 accumulation), bf16 paths
are compared on bf16-rounded operands.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rel_l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def _act(x, dtype, cs=None, c0=0):
    from mspi_b200.ops import Act
    n, c, t, h, w = x.shape
    cs = cs or c
    buf = torch.randn(n, t, h, w, cs).to(dtype).cuda()
    buf[..., c0:c0 + c] = x.permute(0, 2, 3, 4, 1).to(dtype).cuda()
    return Act(buf.contiguous(), c0, c)


WGRAD_CASES = [
    # cin, cout, kernel, stride, pad, (n,t,h,w), dtype
    (64, 64, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 4, 8, 16), torch.bfloat16),
    (64, 192, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 4, 14, 12), torch.bfloat16),
    (96, 208, (3, 1, 1), (1, 1, 1), (1, 0, 0), (2, 4, 7, 12), torch.bfloat16),
    (32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 1, 16, 24), torch.float32),
    (192, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 4, 8, 12), torch.float32),
    (480, 16, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 4, 7, 6), torch.float32),
    (192, 192, (2, 1, 1), (2, 1, 1), (0, 0, 0), (2, 8, 8, 12), torch.float32),   # lateral temporal conv
    (64, 64, (7, 1, 1), (2, 1, 1), (3, 0, 0), (1, 16, 8, 8), torch.float32),     # S3D stem conv_t
    (512, 2048, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1, 2), torch.float32),   # SimSiam head, 2 rows
    (1536, 192, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 4, 2, 3), torch.float32),
]


@pytest.mark.parametrize("cin,cout,k,stride,pad,shape,dtype", WGRAD_CASES)
def test_conv_wgrad(cin, cout, k, stride, pad, shape, dtype):
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(cin + cout)
    n, t, h, w = shape
    x = torch.randn(n, cin, t, h, w, generator=g)
    wgt = torch.zeros(cout, cin, *k)
    conv = ops.Conv(wgt, None, None, stride=stride, pad=pad, dtype=dtype, name="wg")
    ot, oh, ow = conv.out_shape(t, h, w)
    dy = torch.randn(n, cout, ot, oh, ow, generator=g)
    rnd = _bf if dtype == torch.bfloat16 else (lambda v: v)
    xa = _act(x, dtype, cs=cin + 8, c0=8)
    dya = _act(dy, dtype)
    base = torch.randn(cout, cin, *k, generator=g)
    dw = base.clone().cuda()
    conv.wgrad_plan(xa, dya, dw)()
    torch.cuda.synchronize()
    ref = torch.nn.grad.conv3d_weight(rnd(x).double(), wgt.shape, rnd(dy).double(), stride=stride, padding=pad).float()
    got = dw.cpu() - base
    # bf16 operands are exact in the MMA; tf32 rounds both operands to 11 bits: 2e-3 of the gradient norm covers both
    assert _rel_l2(got, ref) < 2e-3, (_rel_l2(got, ref), got.flatten()[:8], ref.flatten()[:8])
