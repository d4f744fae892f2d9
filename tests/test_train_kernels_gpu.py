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
    (32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 1, 64, 64), torch.float32),    # readout.10 at 64x64
    (64, 32, (4, 1, 1), (4, 1, 1), (0, 0, 0), (2, 4, 16, 16), torch.float32),    # readout.8
    (192, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 4, 16, 16), torch.float32),   # readout.4
    (768, 192, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 4, 16, 16), torch.float32),  # readout.0
    (512, 32, (3, 3, 3), (1, 1, 1), (1, 1, 1), (2, 4, 4, 4), torch.float32),     # SA conv_mask.0
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


# ------------------------------------------------------------------------------------------ fp32 training kernels
def _ptr(t, off=0):
    import ctypes as C
    return C.c_void_p(t.data_ptr() + off) if t is not None else None


_KEEP = []


def _dev(t):
    """Move to the GPU and keep the tensor alive until the test module is torn down (the kernels only get raw pointers)."""
    d = t.detach().contiguous().cuda()
    _KEEP.append(d)
    if len(_KEEP) > 64:
        torch.cuda.synchronize()
        del _KEEP[:32]
    return d


def _stream():
    import ctypes as C
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ck(rc, what=""):
    from mspi_b200 import _lib
    _lib.check(rc, what)


def _cl(x):
    """NCDHW cpu -> channels-last cuda [N,T,H,W,C] fp32"""
    return x.permute(0, 2, 3, 4, 1).contiguous().cuda()


def _nc(x):
    return x.permute(0, 4, 1, 2, 3).cpu()


@pytest.mark.parametrize("c,shape,relu", [(64, (2, 4, 8, 12), 1), (24, (2, 2, 5, 7), 1), (192, (1, 4, 6, 6), 1)])
def test_bn_train_fwd_bwd(c, shape, relu):
    from mspi_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(c)
    n, t, h, w = shape
    x = (torch.randn(n, c, t, h, w, generator=g) * 2 + 0.5).requires_grad_(True)
    wt = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    b = (torch.randn(c, generator=g) * 0.1).requires_grad_(True)
    rm, rv = torch.randn(c, generator=g), torch.rand(c, generator=g) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    y_ref = F.relu(F.batch_norm(x, rm_ref, rv_ref, wt, b, True, 0.001, 1e-3))
    dy = torch.randn(n, c, t, h, w, generator=g)
    y_ref.backward(dy)
    dev = "cuda"
    cs = c + 8
    xb = torch.randn(n, t, h, w, cs, device=dev)
    xb[..., 8:] = _cl(x.detach())
    yb = torch.zeros(n, t, h, w, c, device=dev)
    px = n * t * h * w
    rmd, rvd, nbt = rm.cuda(), rv.cuda(), torch.zeros((), dtype=torch.int64, device=dev)
    mean, invstd = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    work = torch.zeros(2 * c, dtype=torch.float64, device=dev)
    wd, bd = wt.detach().cuda(), b.detach().cuda()
    _ck(lib.mspi_bn_train_fwd(_ptr(xb, 32), cs, _ptr(yb), c, px, c, _ptr(wd), _ptr(bd), 1e-3, 0.001, _ptr(rmd), _ptr(rvd), _ptr(nbt),
                              _ptr(mean), _ptr(invstd), _ptr(work), relu, _stream()))
    torch.cuda.synchronize()
    assert (_nc(yb) - y_ref.detach()).abs().max() < 1e-4
    assert (rmd.cpu() - rm_ref).abs().max() < 1e-5 and (rvd.cpu() - rv_ref).abs().max() < 1e-5 and int(nbt) == 1
    work.zero_()   # the caller clears the batch sums between uses
    dyb = _cl(dy)
    dxb = torch.full((n, t, h, w, c), 1.0, device=dev)
    gw, gb = torch.zeros(c, device=dev), torch.zeros(c, device=dev)
    _ck(lib.mspi_bn_train_bwd(_ptr(xb, 32), cs, _ptr(yb), c, _ptr(dyb), c, _ptr(dxb), c, px, c, _ptr(wd), _ptr(mean), _ptr(invstd),
                              _ptr(gw), _ptr(gb), _ptr(work), relu, 1, _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(_nc(dxb) - 1.0, x.grad) < 1e-4
    assert _rel_l2(gw.cpu(), wt.grad) < 1e-4 and _rel_l2(gb.cpu(), b.grad) < 1e-4


@pytest.mark.parametrize("k,s,p,shape", [((3, 3, 3), (1, 1, 1), (1, 1, 1), (2, 4, 6, 8)), ((1, 3, 3), (1, 2, 2), (0, 1, 1), (1, 2, 8, 12)),
                                         ((3, 3, 3), (2, 2, 2), (1, 1, 1), (2, 4, 8, 8)), ((1, 2, 2), (1, 2, 2), (0, 0, 0), (1, 4, 6, 8))])
def test_maxpool_f32_fwd_bwd(k, s, p, shape):
    from mspi_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(5)
    n, t, h, w = shape
    c = 16
    x = torch.randn(n, c, t, h, w, generator=g).requires_grad_(True)
    y_ref = F.max_pool3d(x, k, s, p)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    d = _lib.PoolDesc()
    d.n, d.t, d.h, d.w, d.c = n, t, h, w, c
    d.in_cstride = d.out_cstride = c
    d.kt, d.kh, d.kw = k
    d.st, d.sh, d.sw = s
    d.pt, d.ph, d.pw = p
    d.ot, d.oh, d.ow = y_ref.shape[2:]
    import ctypes as C
    xb = _cl(x.detach())
    yb = torch.zeros(n, d.ot, d.oh, d.ow, c, device="cuda")
    _ck(lib.mspi_maxpool3d_f32(C.byref(d), _ptr(xb), _ptr(yb), _stream()))
    dxb = torch.zeros_like(xb)
    _ck(lib.mspi_maxpool3d_bwd(C.byref(d), _ptr(xb), _ptr(_dev(_cl(dy))), c, _ptr(dxb), c, _stream()))
    torch.cuda.synchronize()
    assert (_nc(yb) - y_ref.detach()).abs().max() == 0
    assert (_nc(dxb) - x.grad).abs().max() < 1e-5


@pytest.mark.parametrize("k,act", [(2, 0), (4, 1), (8, 0)])
def test_upsample_bwd(k, act):
    from mspi_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    g = torch.Generator().manual_seed(k)
    n, t, h, w, c = 2, 2, 5, 6, 16
    x = torch.randn(n, c, t, h, w, generator=g).requires_grad_(True)
    y_ref = F.interpolate(x, scale_factor=(1, k, k), mode="trilinear", align_corners=False)
    if act:
        y_ref = F.relu(y_ref)
    dy = torch.randn(y_ref.shape, generator=g)
    y_ref.backward(dy)
    d = _lib.UpDesc()
    d.nt, d.h, d.w, d.c, d.k = n * t, h, w, c, k
    d.in_cstride = d.out_cstride = c
    d.in_dtype = d.out_dtype = 1
    d.act = act
    base = torch.randn(n, t, h, w, c, device="cuda")
    dxb = base.clone()
    _ck(lib.mspi_upsample_bilinear_bwd(C.byref(d), _ptr(_dev(_cl(dy))), _ptr(_dev(_cl(y_ref.detach()))), _ptr(dxb), _stream()))
    torch.cuda.synchronize()
    assert (_nc(dxb - base) - x.grad).abs().max() < 1e-4


@pytest.mark.parametrize("rows,c,relu", [(37, 512, 0), (2, 2048, 1), (300, 192, 0)])
def test_layernorm_bwd(rows, c, relu):
    from mspi_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, c, generator=g).requires_grad_(True)
    w = (torch.rand(c, generator=g) + 0.5).requires_grad_(True)
    b = torch.randn(c, generator=g).requires_grad_(True)
    y = F.layer_norm(x, (c,), w, b, 1e-5)
    if relu:
        y = F.relu(y)
    dy = torch.randn(rows, c, generator=g)
    y.backward(dy)
    dx = torch.zeros(rows, c, device="cuda")
    dw, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    yd = y.detach().cuda()
    _ck(lib.mspi_layernorm_bwd(_ptr(_dev(x.detach())), c, _ptr(_dev(dy)), c, rows, 0, _ptr(yd) if relu else None,
                               _ptr(_dev(w.detach())), 1e-5, _ptr(dx), c, rows, c, 0, _ptr(dw), _ptr(db), _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(dx.cpu(), x.grad) < 1e-4 and _rel_l2(dw.cpu(), w.grad) < 1e-4 and _rel_l2(db.cpu(), b.grad) < 1e-4


def test_conv_c1_bwd():
    from mspi_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    n, t, h, w = 2, 2, 9, 11
    x = torch.randn(n, 32, t, h, w, generator=g).requires_grad_(True)
    wt = torch.randn(1, 32, 1, 3, 3, generator=g).requires_grad_(True)
    b = torch.randn(1, generator=g).requires_grad_(True)
    y = F.conv3d(x, wt, b, padding=(0, 1, 1))
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    xb = _cl(x.detach())
    # forward (fp32 and bf16 inputs, the latter as a channel slice of a wider buffer like the SA masks)
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    yb = Act(torch.zeros(n, t, h, w, 1, device="cuda"))
    ops.conv_c1(Act(xb), wt.detach().cuda(), b.detach().cuda(), yb)()
    wide = torch.randn(n, t, h, w, 96, device="cuda").to(torch.bfloat16)
    wide[..., 64:] = xb.to(torch.bfloat16)
    yb16 = Act(torch.zeros(n, t, h, w, 1, device="cuda"))
    ops.conv_c1(Act(wide, 64, 32), wt.detach().cuda(), b.detach().cuda(), yb16)()
    torch.cuda.synchronize()
    assert _rel_l2(yb.buf.cpu().reshape(y.shape), y.detach()) < 1e-5
    y16_ref = F.conv3d(_bf(x.detach()), wt.detach(), b.detach(), padding=(0, 1, 1))
    assert _rel_l2(yb16.buf.cpu().reshape(y.shape), y16_ref) < 1e-5
    dxb = torch.zeros_like(xb)
    dw, db = torch.zeros(288, device="cuda"), torch.zeros(1, device="cuda")
    _ck(lib.mspi_conv_c1_bwd(_ptr(xb), 32, _ptr(_dev(dy.reshape(-1))), _ptr(_dev(wt.detach().reshape(-1))), _ptr(dxb), 32, _ptr(dw),
                             _ptr(db), n * t, h, w, 32, 0, _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(_nc(dxb), x.grad) < 1e-5 and _rel_l2(dw.cpu(), wt.grad.reshape(-1)) < 1e-5
    assert abs(float(db) - float(b.grad)) < 1e-3


@pytest.mark.parametrize("k", [(7, 1, 1), (1, 7, 7)])
def test_dwconv_wgrad_and_flipped_dgrad(k):
    from mspi_b200 import _lib
    import ctypes as C
    lib = _lib.load()
    g = torch.Generator().manual_seed(9)
    n, t, h, w, c = 2, 4, 9, 10, 64
    x = torch.randn(n, c, t, h, w, generator=g).requires_grad_(True)
    wt = torch.randn(c, 1, *k, generator=g).requires_grad_(True)
    b = torch.randn(c, generator=g).requires_grad_(True)
    y = F.conv3d(x, wt, b, padding=(k[0] // 2, k[1] // 2, k[2] // 2), groups=c)
    dy = torch.randn(y.shape, generator=g)
    y.backward(dy)
    d = _lib.DwDesc()
    d.n, d.t, d.h, d.w, d.c = n, t, h, w, c
    d.kt, d.kh, d.kw = k
    d.ln_eps, d.out_dtype, d.in_dtype = 1e-5, 1, 1
    taps = k[0] * k[1] * k[2]
    dw, db = torch.zeros(c * taps, device="cuda"), torch.zeros(c, device="cuda")
    xb, dyb = _cl(x.detach()), _cl(dy)
    _ck(lib.mspi_dwconv_wgrad(C.byref(d), _ptr(xb), _ptr(dyb), _ptr(dw), _ptr(db), _stream()))
    # data gradient = the forward kernel on the flipped filter (packed by mspi_permute_copy, as the plan does)
    wsrc = wt.detach().reshape(c, taps).contiguous().cuda()
    wflip = torch.zeros(taps, c, device="cuda")
    pd = _lib.PermDesc()
    for j, (n_, s_, d_) in enumerate(zip((c, taps, 1, 1), (taps, -1, 0, 0), (1, c, 0, 0))):
        pd.n[j], pd.src_strides[j], pd.dst_strides[j] = n_, s_, d_
    pd.dst_dtype, pd.accumulate = 1, 0
    _ck(lib.mspi_permute_copy(C.byref(pd), _ptr(wsrc, (taps - 1) * 4), _ptr(wflip), _stream()))
    dxb = torch.zeros_like(xb)
    zb = torch.zeros(c, device="cuda")
    _ck(lib.mspi_dwconv_ln(C.byref(d), _ptr(dyb), _ptr(wflip), _ptr(zb), None, None, _ptr(dxb), _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(dw.cpu(), wt.grad.reshape(-1)) < 1e-4 and _rel_l2(db.cpu(), b.grad) < 1e-4
    assert _rel_l2(_nc(dxb), x.grad) < 1e-4


def test_salloss_and_simsiam_bwd():
    from mspi_b200 import _lib
    from oracle import mspi_oracle as orc
    lib = _lib.load()
    g = torch.Generator().manual_seed(1)
    b, h, w = 2, 16, 24
    logits = (torch.randn(b, h, w, generator=g) * 2).requires_grad_(True)
    gt = torch.rand(b, h, w, generator=g) ** 3
    logp = logits - torch.logsumexp(logits, dim=(1, 2), keepdim=True)
    parts = orc.sal_loss(logp, gt)
    parts["loss"].backward()
    dl = torch.zeros(b, h * w, device="cuda")
    out, work = torch.zeros(4, device="cuda"), torch.zeros(2 * b, device="cuda")
    va = torch.full((1,), 0.25, device="cuda")
    _ck(lib.mspi_salloss_bwd(_ptr(_dev(logp.detach())), _ptr(_dev(gt)), _ptr(va), 2.0, _ptr(dl), _ptr(out), _ptr(work), b, h * w, 1.0,
                             _stream()))
    torch.cuda.synchronize()
    o = out.cpu()
    assert abs(float(o[1]) - float(parts["kl"])) < 1e-5 and abs(float(o[2]) - float(parts["cc"])) < 1e-5
    assert abs(float(o[0]) - (float(parts["loss"]) + 0.5)) < 1e-5
    assert _rel_l2(dl.cpu().view(b, h, w), logits.grad) < 1e-4
    # SimSiam
    c = 256
    pv, pa = torch.randn(b, c, generator=g).requires_grad_(True), torch.randn(b, c, generator=g).requires_grad_(True)
    za, zv = torch.randn(b, c, generator=g), torch.randn(b, c, generator=g)
    d = lambda a, z: -F.cosine_similarity(a, z.detach(), dim=-1).mean()
    (0.5 * (d(pv, za) + d(pa, zv))).backward()
    dpv, dpa = torch.zeros(b, c, device="cuda"), torch.zeros(b, c, device="cuda")
    _ck(lib.mspi_simsiam_bwd(_ptr(_dev(pv.detach())), _ptr(_dev(za)), _ptr(_dev(pa.detach())), _ptr(_dev(zv)), _ptr(dpv), _ptr(dpa),
                             b, c, 1.0, _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(dpv.cpu(), pv.grad) < 1e-4 and _rel_l2(dpa.cpu(), pa.grad) < 1e-4


def test_sgemm_softmax_gate_adamw():
    from mspi_b200 import _lib
    from oracle import mspi_oracle as orc
    import ctypes as C
    lib = _lib.load()
    g = torch.Generator().manual_seed(2)
    # strided batched GEMM with a transposed A view:  C[b1][b0] = 0.5 * A^T B
    b0, b1, m, n, k = 3, 2, 45, 70, 37
    A = torch.randn(b1, b0, k, m, generator=g)
    Bm = torch.randn(b1, b0, k, n, generator=g)
    Cm = torch.zeros(b1, b0, m, n, device="cuda")
    d = _lib.SgemmDesc()
    d.m, d.n, d.k, d.batch0, d.batch1, d.accumulate, d.alpha = m, n, k, b0, b1, 0, 0.5
    for j, (a_, b_, c_) in enumerate(zip((1, m, k * m, b0 * k * m), (n, 1, k * n, b0 * k * n), (n, 1, m * n, b0 * m * n))):
        d.a_strides[j], d.b_strides[j], d.c_strides[j] = a_, b_, c_
    _ck(lib.mspi_sgemm_strided(C.byref(d), _ptr(_dev(A)), _ptr(_dev(Bm)), _ptr(Cm), _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(Cm.cpu(), 0.5 * A.transpose(2, 3) @ Bm) < 1e-5
    # softmax backward
    s = torch.randn(20, 33, generator=g).requires_grad_(True)
    p = torch.softmax(0.3 * s, -1)
    dp = torch.randn(20, 33, generator=g)
    p.backward(dp)
    dpd = dp.clone().cuda()
    _ck(lib.mspi_softmax_bwd_rows(_ptr(_dev(p.detach())), _ptr(dpd), 20, 33, 33, 0.3, _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(dpd.cpu(), s.grad) < 1e-5
    # SA gate backward
    px, c = 50, 64
    x = torch.randn(px, c, generator=g).requires_grad_(True)
    l = torch.randn(px, generator=g).requires_grad_(True)
    y = x * torch.sigmoid(l)[:, None] + x
    dy = torch.randn(px, c, generator=g)
    y.backward(dy)
    dx, dl = torch.zeros(px, c, device="cuda"), torch.zeros(px, device="cuda")
    _ck(lib.mspi_sa_gate_bwd(_ptr(_dev(x.detach())), c, _ptr(_dev(l.detach())), _ptr(_dev(dy)), c, _ptr(dx), c, _ptr(dl), px, c, 0,
                             _stream()))
    torch.cuda.synchronize()
    assert _rel_l2(dx.cpu(), x.grad) < 1e-5 and _rel_l2(dl.cpu(), l.grad) < 1e-5
    # AdamW, two steps
    nel = 1024
    p0, gr = torch.randn(nel, generator=g), torch.randn(nel, generator=g) * 1e-3
    pd_, md, vd = p0.clone().cuda(), torch.zeros(nel, device="cuda"), torch.zeros(nel, device="cuda")
    pr, mr, vr = p0.clone(), torch.zeros(nel), torch.zeros(nel)
    for step in (1, 2):
        _ck(lib.mspi_adamw_step(_ptr(pd_), _ptr(_dev(2 * gr)), _ptr(md), _ptr(vd), nel, 1e-4, 0.9, 0.999, 1e-8, 0.0, step, None, 0.5,
                                _stream()))
        pr, mr, vr = orc.adamw_step(pr, gr, mr, vr, step)
    torch.cuda.synchronize()
    assert (pd_.cpu() - pr).abs().max() < 1e-7
