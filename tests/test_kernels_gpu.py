"""GPU parity tests of the individual C-ABI kernels against plain PyTorch fp32 (CPU) references.

Tolerances: bf16 tensor-core paths are compared on bf16-rounded operands with fp32 accumulation,
so the only differences are accumulation order and the final bf16 rounding of the output (2^-8
relative); fp32 paths are compared at 1e-5-class tolerances.  Each tolerance is stated at its use.
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def _act_from_ncdhw(x, cs=None, c0=0, dtype=torch.bfloat16):
    """NCDHW fp32 CPU tensor -> Act on the GPU (optionally as a slice of a wider buffer)."""
    from mspi_b200.ops import Act
    n, c, t, h, w = x.shape
    cs = cs or c
    buf = torch.randn(n, t, h, w, cs).to(dtype).cuda()  # garbage around the slice
    buf[..., c0:c0 + c] = x.permute(0, 2, 3, 4, 1).to(dtype).cuda()
    return Act(buf.contiguous(), c0, c)


def _run_conv(cin, cout, k, stride, pad, shape, act=0, residual=False, out_dtype=torch.bfloat16, dtype=torch.bfloat16,
              in_slice=None, out_slice=None, res_after_act=False, seed=0):
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(seed)
    n, t, h, w = shape
    x = torch.randn(n, cin, t, h, w, generator=g)
    wgt = torch.randn(cout, cin, *k, generator=g) / (cin * k[0] * k[1] * k[2]) ** 0.5
    scale = torch.rand(cout, generator=g) + 0.5
    shift = torch.randn(cout, generator=g) * 0.1
    rnd = _bf if dtype == torch.bfloat16 else (lambda v: v)
    conv = ops.Conv(wgt, scale, shift, stride=stride, pad=pad, act=act, dtype=dtype, res_after_act=res_after_act, name="t")
    xa = _act_from_ncdhw(x, *(in_slice or (None, 0)), dtype=dtype)
    ot, oh, ow = conv.out_shape(t, h, w)
    ocs, oc0 = out_slice or (cout, 0)
    ybuf = torch.full((n, ot, oh, ow, ocs), 7.0, dtype=out_dtype, device="cuda")
    ya = Act(ybuf, oc0, cout)
    ra = None
    ref = F.conv3d(rnd(x), rnd(wgt), None, stride, pad) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    if residual:
        r = torch.randn(n, cout, ot, oh, ow, generator=g)
        ra = _act_from_ncdhw(r)
        if not res_after_act:
            ref = ref + _bf(r)
    if act == 1:
        ref = ref.relu()
    elif act == 2:
        ref = F.gelu(ref)
    elif act == 3:
        ref = ref.sigmoid()
    if residual and res_after_act:
        ref = ref + _bf(r)
    run = conv.plan(xa, ya, ra)
    run()
    torch.cuda.synchronize()
    got = ya.to_ncdhw().cpu()
    # untouched channels of a wider output buffer must keep their fill value
    if out_slice:
        rest = torch.cat([ybuf[..., :oc0], ybuf[..., oc0 + cout:]], -1)
        assert (rest.float() == 7.0).all(), "conv wrote outside its channel slice"
    return got, ref, run.mode


# bf16 out: 2^-8 relative rounding of the output + accumulation-order noise -> 1.5e-2 of max is ample
BF16_TOL = 1.5e-2


def test_gemm_plain_linear():
    got, ref, mode = _run_conv(192, 96, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 1, 300))
    assert mode == "shift"
    assert _rel(got, ref) < BF16_TOL


def test_gemm_multi_ntile_gelu_fp32out():
    got, ref, _ = _run_conv(96, 384, (1, 1, 1), (1, 1, 1), (0, 0, 0), (2, 1, 5, 77), act=2, out_dtype=torch.float32)
    assert _rel(got, ref) < 2e-3  # fp32 output: only accumulation order differs


@pytest.mark.parametrize("k,pad", [((1, 3, 3), (0, 1, 1)), ((3, 1, 1), (1, 0, 0)), ((3, 3, 3), (1, 1, 1)),
                                   ((7, 1, 1), (3, 0, 0))])
def test_conv_shift_mode(k, pad):
    got, ref, mode = _run_conv(64, 192, k, (1, 1, 1), pad, (2, 4, 14, 24), act=1)
    assert mode == "shift"
    assert _rel(got, ref) < BF16_TOL


def test_conv_small_channels_and_slices():
    # Cin=16 (quarter of a K chunk), input is a slice of a wider buffer, output goes into a concat slice
    got, ref, _ = _run_conv(16, 48, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 4, 7, 12), act=1, in_slice=(112, 96),
                            out_slice=(512, 400))
    assert _rel(got, ref) < BF16_TOL


def test_conv_cin_not_multiple_of_chunk():
    got, ref, _ = _run_conv(96, 208, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 4, 14, 24), act=1)
    assert _rel(got, ref) < BF16_TOL


@pytest.mark.parametrize("k,stride,pad,t", [((7, 1, 1), (2, 1, 1), (3, 0, 0), 16), ((2, 1, 1), (2, 1, 1), (0, 0, 0), 8),
                                            ((4, 1, 1), (4, 1, 1), (0, 0, 0), 4)])
def test_conv_temporal_stride(k, stride, pad, t):
    got, ref, mode = _run_conv(64, 64, k, stride, pad, (2, t, 12, 20), act=1)
    assert mode == "tstride"
    assert _rel(got, ref) < BF16_TOL


def test_conv_gather_mode_strided_odd():
    got, ref, mode = _run_conv(64, 128, (1, 3, 3), (1, 2, 2), (0, 1, 1), (2, 1, 65, 28), act=1)
    assert mode == "gather"
    assert _rel(got, ref) < BF16_TOL


def test_conv_residual_relu():
    got, ref, _ = _run_conv(64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), (2, 1, 33, 14), act=1, residual=True)
    assert _rel(got, ref) < BF16_TOL


def test_conv_residual_after_act_layer_scale():
    got, ref, _ = _run_conv(384, 96, (1, 1, 1), (1, 1, 1), (0, 0, 0), (1, 1, 20, 31), act=0, residual=True, res_after_act=True)
    assert _rel(got, ref) < BF16_TOL


def test_conv_single_output_channel_fp32():
    got, ref, _ = _run_conv(32, 1, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 32, 64), out_dtype=torch.float32)
    assert _rel(got, ref) < 2e-3


def test_conv_tf32_path():
    # kind::tf32: 10-bit mantissa operands (2^-11 relative), fp32 accumulate and fp32 output
    got, ref, _ = _run_conv(32, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), (1, 1, 24, 40), act=1, dtype=torch.float32,
                            out_dtype=torch.float32)
    assert _rel(got, ref) < 3e-3


def test_conv_large_k_many_tiles():
    got, ref, _ = _run_conv(192, 192, (3, 3, 3), (1, 1, 1), (1, 1, 1), (1, 4, 28, 48), act=1)
    assert _rel(got, ref) < BF16_TOL


def test_patch_gather_from_ncdhw_stem():
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 4, 20, 36, generator=g)
    k, s, p = (1, 7, 7), (1, 2, 2), (0, 3, 3)
    kk = 3 * 49
    k_pad = 192
    ot, oh, ow = 4, 10, 18
    out = torch.empty(2 * ot * oh * ow, k_pad, dtype=torch.bfloat16, device="cuda")
    ops.patch_gather_ncdhw(x.cuda(), k, s, p, k_pad, out)()
    torch.cuda.synchronize()
    cols = F.unfold(x.permute(0, 2, 1, 3, 4).reshape(8, 3, 20, 36), (7, 7), padding=3, stride=2)  # [NT, C*49, L]
    cols = cols.view(8, 3, 49, oh * ow).permute(0, 3, 2, 1).reshape(8 * oh * ow, kk)  # K order (kh,kw,c)
    got = out.float().cpu()
    assert torch.equal(got[:, :kk], _bf(cols)), "patch rows differ"  # pure data movement: bit-exact
    assert (got[:, kk:] == 0).all()


@pytest.mark.parametrize("k,stride,pad,cout,odt", [(7, 2, 3, 64, torch.bfloat16), (4, 4, 0, 96, torch.float32)])
def test_stem_conv_on_padded_frames(k, stride, pad, cout, odt):
    """Cin=3 stems as implicit GEMMs off the padded 4-channel frames (MspiConvDesc.k_rows): S3D conv_s 7x7/s2
    (s3d.py:383) and the ConvNeXt 4x4/s4 stem, against F.conv2d on the bf16-rounded clip."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(21)
    b, t, h, w = 2, 3, 32, 64
    clip = torch.randn(b, 3, t, h, w, generator=g)
    wgt = torch.randn(cout, 3, k, k, generator=g) / (3 * k * k) ** 0.5
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    frames = torch.zeros(b * t, h + ops.PAD_EXTRA, w + ops.PAD_EXTRA, 4, dtype=torch.bfloat16, device="cuda")
    holder = {"clips": clip.cuda()}
    ops.clip_to_padded(holder, "clips", frames, b, t, h, w)()
    torch.cuda.synchronize()
    inner = frames[:, ops.PAD_T:ops.PAD_T + h, ops.PAD_L:ops.PAD_L + w].float().cpu()
    want = _bf(clip).permute(0, 2, 3, 4, 1).reshape(b * t, h, w, 3)
    assert torch.equal(inner[..., :3], want) and (inner[..., 3] == 0).all()
    assert frames.float().abs().sum().item() == pytest.approx(inner.abs().sum().item(), rel=1e-6)  # borders stay zero
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    y = Act.empty(b, t, oh, ow, cout, dtype=odt)
    ops.stem_conv(frames, h, w, wgt, scale, shift, k, stride, pad, 1, y)()
    torch.cuda.synchronize()
    x2 = _bf(clip).permute(0, 2, 1, 3, 4).reshape(b * t, 3, h, w)
    ref = F.conv2d(x2, _bf(wgt), None, stride, pad) * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)
    ref = ref.relu().view(b, t, cout, oh, ow).permute(0, 2, 1, 3, 4)
    got = y.to_ncdhw().cpu()
    assert _rel(got, ref) < (BF16_TOL if odt == torch.bfloat16 else 1e-4)


@pytest.mark.parametrize("b,t,h,w", [(2, 3, 32, 64), (1, 5, 44, 72), (3, 2, 224, 384)])
def test_stem_conv_with_layernorm_epilogue(b, t, h, w):
    """ConvNeXt stem (Conv2d 4x4/s4 + LayerNorm2d, timm convnext_tiny) in one launch: mspi_conv_gemm_ln normalises each of
    the two output pixels of a GEMM row over its 96 channels.  Inputs with a large common offset (|mean| >> sigma per pixel,
    what the stem sees on bright frames) check the two-pass statistics; reference = F.conv2d + F.layer_norm in fp32."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(23)
    cout = 96
    clip = torch.randn(b, 3, t, h, w, generator=g) + 2.0
    wgt = torch.randn(cout, 3, 4, 4, generator=g) / 48 ** 0.5 + 0.05
    bias = torch.randn(cout, generator=g) * 0.1
    lw, lb = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.2
    assert ops.stem_conv_ln_ok(w, cout, 4, 4, 0)
    frames = torch.zeros(b * t, h + ops.PAD_EXTRA, w + ops.PAD_EXTRA, 4, dtype=torch.bfloat16, device="cuda")
    ops.clip_to_padded({"clips": clip.cuda()}, "clips", frames, b, t, h, w)()
    oh, ow = h // 4, w // 4
    y = Act(torch.full((b, t, oh, ow, cout), 7.0, dtype=torch.bfloat16, device="cuda"))
    ops.stem_conv(frames, h, w, wgt, None, bias, 4, 4, 0, 0, y, ln=(lw, lb, 1e-6))()
    torch.cuda.synchronize()
    x2 = _bf(clip).permute(0, 2, 1, 3, 4).reshape(b * t, 3, h, w)
    conv = F.conv2d(x2, _bf(wgt), bias, 4, 0).permute(0, 2, 3, 1)            # [bt, oh, ow, c]
    ref = F.layer_norm(conv, (cout,), lw, lb, 1e-6).view(b, t, oh, ow, cout)
    got = y.buf.float().cpu()
    assert _rel(got, ref) < BF16_TOL
    assert (got - ref).abs().max() < 0.04 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("c,n,t,h,w", [(8, 2, 5, 14, 24), (16, 2, 3, 9, 13), (8, 1, 2, 56, 96), (16, 1, 2, 28, 48)])
def test_conv133_small_channels(c, n, t, h, w):
    """SlowFast fast-pathway branch2.b (resnet_helper.py:323-341, dim_inner 8 / 16): (1,3,3) conv + BN + ReLU on CUDA cores
    (mspi_conv133_small) against F.conv3d on the bf16-rounded input; ragged widths, channel-slice views on both sides, and the
    plan's dispatch (ForwardPlan.conv sends exactly these layers here)."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(17)
    x = torch.randn(n, c, t, h, w, generator=g)
    wgt = torch.randn(c, c, 1, 3, 3, generator=g) / (9 * c) ** 0.5
    scale, shift = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    xa = _act_from_ncdhw(x, c + 8, 8)
    buf = torch.zeros(n, t, h, w, c + 16, dtype=torch.bfloat16, device="cuda")
    ya = Act(buf, 8, c)
    assert ops.conv133_small_ok(wgt, (1, 1, 1), (0, 1, 1), xa, torch.bfloat16, torch.bfloat16, None)
    ops.conv133_small(xa, ya, wgt, scale, shift, 1)()
    torch.cuda.synchronize()
    ref = (F.conv3d(_bf(x), wgt, None, 1, (0, 1, 1)) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)).relu()
    assert _rel(ya.to_ncdhw().cpu(), ref) < BF16_TOL
    assert (buf[..., :8] == 0).all() and (buf[..., 8 + c:] == 0).all()
    assert not ops.conv133_small_ok(wgt, (1, 2, 2), (0, 1, 1), xa, torch.bfloat16, torch.bfloat16, None)


def test_conv_1x1_spatial_stride_pick_mode():
    """ResBlock.branch1 (1x1x1 conv, stride (1,2,2), resnet_helper.py:556-566) as a strided TMA view, no gather."""
    got, ref, mode = _run_conv(24, 48, (1, 1, 1), (1, 2, 2), (0, 0, 0), (2, 3, 12, 20), seed=31)
    assert mode == "pick" and _rel(got, ref) < BF16_TOL


@pytest.mark.parametrize("cin,cout,shape", [(96, 192, (5, 1, 8, 12)), (384, 768, (3, 1, 14, 24)), (64, 96, (2, 3, 6, 10))])
def test_conv_patch2_mode(cin, cout, shape):
    """2x2 / stride-2 patch convolution (ConvNeXt downsample) as an implicit GEMM: row / column parity as tensor axes."""
    got, ref, mode = _run_conv(cin, cout, (1, 2, 2), (1, 2, 2), (0, 0, 0), shape, seed=cin)
    assert mode == "patch2" and _rel(got, ref) < BF16_TOL


def test_clip_frame_map_and_time_padding():
    """SlowFast inputs: slow pathway = frames [0,4,12,T-1] (model_utils.py:523); fast pathway time-padded by 2."""
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(22)
    b, t, h, w = 2, 16, 8, 12
    clip = torch.randn(b, 3, t, h, w, generator=g)
    fm = (0, 4, 12, -1)
    fr = torch.zeros(b * 4, h + 8, w + 8, 4, dtype=torch.bfloat16, device="cuda")
    ops.clip_to_padded({"c": clip.cuda()}, "c", fr, b, t, h, w, fm, 0)()
    fr2 = torch.zeros(b * (t + 4), h + 8, w + 8, 4, dtype=torch.bfloat16, device="cuda")
    ops.clip_to_padded({"c": clip.cuda()}, "c", fr2, b, t, h, w, None, 2)()
    torch.cuda.synchronize()
    inner = fr[:, ops.PAD_T:ops.PAD_T + h, ops.PAD_L:ops.PAD_L + w, :3].float().cpu().view(b, 4, h, w, 3)
    want = _bf(clip)[:, :, [0, 4, 12, 15]].permute(0, 2, 3, 4, 1)
    assert torch.equal(inner, want)
    v = fr2.view(b, t + 4, h + 8, w + 8, 4).float().cpu()
    assert (v[:, :2] == 0).all() and (v[:, -2:] == 0).all()
    assert torch.equal(v[:, 2:-2, ops.PAD_T:ops.PAD_T + h, ops.PAD_L:ops.PAD_L + w, :3], _bf(clip).permute(0, 2, 3, 4, 1))


def test_clip_u8_conversion_is_bit_exact_and_gather_rows():
    """mspi_clip_u8_to_padded_nhwc4 == mspi_clip_to_padded_nhwc4 of the host-normalised fp32 clip (bit for bit), with and
    without a frame map; mspi_gather_rows copies / clamps rows."""
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(9)
    n, t, h, w = 2, 8, 20, 24
    u8 = torch.randint(0, 256, (n, t, h, w, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(ops.IMAGENET_MEAN).view(1, 1, 1, 1, 3), torch.tensor(ops.IMAGENET_STD).view(1, 1, 1, 1, 3)
    clip = ((u8.float() / 255.0 - mean) / std).permute(0, 4, 1, 2, 3).contiguous().cuda()
    for fmap, tpad in ((None, 0), ([0, 3, 6, 7], 0), (None, 2)):
        t_out = t if fmap is None else len(fmap)
        shape = (n * (t_out + 2 * tpad), h + ops.PAD_EXTRA, w + ops.PAD_EXTRA, 4)
        a = torch.zeros(shape, dtype=torch.bfloat16, device="cuda")
        b = torch.zeros(shape, dtype=torch.bfloat16, device="cuda")
        ops.clip_to_padded({"c": clip}, "c", a, n, t, h, w, fmap, tpad)()
        ops.clip_u8_to_padded({"c": u8.cuda()}, "c", b, n, t, h, w, frame_map=fmap, t_pad=tpad)()
        torch.cuda.synchronize()
        assert torch.equal(a.view(torch.int16), b.view(torch.int16))
    src = torch.randn(11, 40, generator=g).to(torch.bfloat16).cuda()
    idx = torch.tensor([3, 3, 0, 10, 7, -2, 99], dtype=torch.int32, device="cuda")
    dst = torch.empty(7, 40, dtype=torch.bfloat16, device="cuda")
    ops.gather_rows({"s": src}, "s", idx, dst, 40)()
    torch.cuda.synchronize()
    assert torch.equal(dst, src[idx.clamp(0, 10).long()])


def test_stem_conv_temporal_taps_per_clip():
    """SlowFast fast stem (5,7,7)/s(1,2,2) p(2,3,3), 3->8 (stem_helper.py:128-204): 35 taps, one launch per clip."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(23)
    b, t, h, w, cout = 2, 16, 32, 64, 8
    clip = torch.randn(b, 3, t, h, w, generator=g)
    wgt = torch.randn(cout, 3, 5, 7, 7, generator=g) / (3 * 5 * 49) ** 0.5
    scale, shift = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    fr = torch.zeros(b * (t + 4), h + 8, w + 8, 4, dtype=torch.bfloat16, device="cuda")
    ops.clip_to_padded({"c": clip.cuda()}, "c", fr, b, t, h, w, None, 2)()
    y = Act.empty(b, t, h // 2, w // 2, cout)
    ops.stem_conv(fr, h, w, wgt, scale, shift, 7, 2, 3, 1, y, clips=b)()
    torch.cuda.synchronize()
    ref = F.conv3d(_bf(clip), _bf(wgt), None, (1, 2, 2), (2, 3, 3)) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    assert _rel(y.to_ncdhw().cpu(), ref.relu()) < BF16_TOL


@pytest.mark.parametrize("c,act,n,t,h,w", [(56, 4, 2, 4, 20, 32), (112, 0, 1, 3, 9, 48), (216, 4, 2, 2, 14, 24), (432, 1, 1, 4, 7, 12),
                                           (56, 0, 1, 1, 5, 7), (80, 4, 1, 5, 11, 40)])
def test_x3d_depthwise_3x3x3_tiled(c, act, n, t, h, w):
    """The shared-memory tiled 3x3x3 depthwise kernel (stride 1; x3d.cu dw3d_tile_kernel): channel groups (C = 112 / 216 / 432),
    1-3 strips per tile row, ragged tiles, single-frame clips (both temporal neighbours are padding), every activation."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(31)
    creal = c - 2 if c in (56, 112, 216, 432) else c
    x = torch.zeros(n, c, t, h, w)
    x[:, :creal] = torch.randn(n, creal, t, h, w, generator=g)
    wgt = torch.randn(creal, 1, 3, 3, 3, generator=g) * 0.3
    scale, shift = torch.rand(creal, generator=g) + 0.5, torch.randn(creal, generator=g) * 0.1
    xa = _act_from_ncdhw(x)
    ya = Act.empty(n, t, h, w, c)
    ops.dwconv3d_bn(xa, ya, wgt, scale, shift, 1, act)()
    torch.cuda.synchronize()
    ref = F.conv3d(_bf(x[:, :creal]), wgt, None, 1, 1, 1, creal) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    ref = ref * torch.sigmoid(ref) if act == 4 else (ref.relu() if act == 1 else ref)
    got = ya.to_ncdhw().cpu()
    assert _rel(got[:, :creal], ref) < BF16_TOL and (got[:, creal:] == 0).all()


@pytest.mark.parametrize("c,stride,act", [(56, 2, 4), (112, 1, 0), (24, 1, 1)])
def test_x3d_depthwise_and_se(c, stride, act):
    """X3DTransform.b + b_bn (+Swish) and the SE path (resnet_helper.py:47-73,213-351) against PyTorch."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(24)
    creal = c - 2 if c % 8 == 0 and c > 24 else c          # weights narrower than the padded buffers (54 -> 56)
    n, t, h, w = 2, 4, 10, 12
    x = torch.zeros(n, c, t, h, w)
    x[:, :creal] = torch.randn(n, creal, t, h, w, generator=g)
    k = (5, 1, 1) if c == 24 else (3, 3, 3)
    wgt = torch.randn(creal, 1, *k, generator=g) * 0.3
    scale, shift = torch.rand(creal, generator=g) + 0.5, torch.randn(creal, generator=g) * 0.1
    xa = _act_from_ncdhw(x)
    oh, ow = (h - 1) // stride + 1 if k[1] == 3 else h, (w - 1) // stride + 1 if k[2] == 3 else w
    ya = Act.empty(n, t, oh, ow, c)
    ops.dwconv3d_bn(xa, ya, wgt, scale, shift, stride, act)()
    torch.cuda.synchronize()
    ref = F.conv3d(_bf(x[:, :creal]), wgt, None, (1, stride, stride) if k[1] == 3 else 1, tuple(kk // 2 for kk in k), 1, creal)
    ref = ref * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    ref = ref * torch.sigmoid(ref) if act == 4 else (ref.relu() if act == 1 else ref)
    got = ya.to_ncdhw().cpu()
    assert _rel(got[:, :creal], ref) < BF16_TOL and (got[:, creal:] == 0).all()
    # SE + Swish in place on the conv output
    cfc = 8
    w1, b1 = torch.randn(cfc, creal, 1, 1, 1, generator=g) * 0.3, torch.randn(cfc, generator=g) * 0.1
    w2, b2 = torch.randn(creal, cfc, 1, 1, 1, generator=g) * 0.3, torch.randn(creal, generator=g) * 0.1
    for fn in ops.se_block(ya, w1, b1, w2, b2):
        fn()
    torch.cuda.synchronize()
    y0 = got[:, :creal]
    s = y0.mean((2, 3, 4), keepdim=True)
    s = torch.sigmoid(F.conv3d(F.relu(F.conv3d(s, w1, b1)), w2, b2))
    z = y0 * s
    z = z * torch.sigmoid(z)
    got2 = ya.to_ncdhw().cpu()
    assert _rel(got2[:, :creal], z) < BF16_TOL and (got2[:, creal:] == 0).all()


@pytest.mark.parametrize("c,stride,n,t,h,w", [(56, 1, 2, 4, 20, 32), (112, 1, 2, 3, 9, 48), (216, 1, 3, 2, 14, 24),
                                              (432, 1, 2, 4, 7, 12), (56, 2, 2, 4, 20, 32), (56, 1, 1, 1, 5, 7)])
def test_x3d_depthwise_with_fused_se_mean(c, stride, n, t, h, w):
    """mspi_dwconv3d_bn_mean: the SE squeeze (mean over t, h, w per sample and channel, resnet_helper.py:47-73) accumulated by
    the tiled depthwise kernel while it stores its output; stride-2 and tiny shapes fall back to the separate reduction.
    The conv output must be the plain call's bits, the mean PyTorch's mean of it, and se_block(mean=) the plain SE result."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(33)
    creal = c - 2
    x = torch.zeros(n, c, t, h, w)
    x[:, :creal] = torch.randn(n, creal, t, h, w, generator=g)
    wgt = torch.randn(creal, 1, 3, 3, 3, generator=g) * 0.3
    scale, shift = torch.rand(creal, generator=g) + 0.5, torch.randn(creal, generator=g) * 0.1 + 0.3
    xa = _act_from_ncdhw(x)
    oh, ow = (h - 1) // stride + 1, (w - 1) // stride + 1
    y0, y1 = Act.empty(n, t, oh, ow, c), Act.empty(n, t, oh, ow, c)
    mean = torch.full((n, c), 7.0, dtype=torch.float32, device="cuda")
    ops.dwconv3d_bn(xa, y0, wgt, scale, shift, stride, 0)()
    ops.dwconv3d_bn(xa, y1, wgt, scale, shift, stride, 0, mean=mean)()
    torch.cuda.synchronize()
    assert torch.equal(y0.buf, y1.buf)
    ref = F.conv3d(_bf(x[:, :creal]), wgt, None, (1, stride, stride), 1, 1, creal) * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)
    m = mean.cpu()
    assert (m[:, :creal] - ref.mean((2, 3, 4))).abs().max() < 2e-3 * max(1.0, ref.mean((2, 3, 4)).abs().max().item())
    assert (m[:, creal:] == 0).all()
    cfc = 8
    w1, b1 = torch.randn(cfc, creal, 1, 1, 1, generator=g) * 0.3, torch.randn(cfc, generator=g) * 0.1
    w2, b2 = torch.randn(creal, cfc, 1, 1, 1, generator=g) * 0.3, torch.randn(creal, generator=g) * 0.1
    fns = ops.se_block(y1, w1, b1, w2, b2, mean=mean)
    assert len(fns) == 2
    for fn in fns:
        fn()
    for fn in ops.se_block(y0, w1, b1, w2, b2):
        fn()
    torch.cuda.synchronize()
    assert _rel(y1.buf.float().cpu(), y0.buf.float().cpu()) < BF16_TOL


def test_maxpool_variants():
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(2)
    for (k, s, p) in [((1, 3, 3), (1, 2, 2), (0, 1, 1)), ((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)),
                      ((1, 2, 2), (1, 2, 2), (0, 0, 0)), ((4, 1, 1), (4, 1, 1), (0, 0, 0))]:
        x = torch.randn(2, 24, 8, 14, 12, generator=g)
        xa = _act_from_ncdhw(x, 40, 8)
        ref = F.max_pool3d(_bf(x), k, s, p)
        ya = Act.empty(2, *ref.shape[2:], 24)
        ops.maxpool3d(xa, ya, k, s, p)()
        torch.cuda.synchronize()
        assert torch.equal(ya.to_ncdhw().cpu(), ref), f"maxpool {k} {s} {p}"  # selection only: bit-exact


@pytest.mark.parametrize("n,c,t,h,w", [(2, 32, 8, 28, 48), (3, 48, 4, 14, 24), (2, 64, 2, 7, 12), (1, 16, 1, 5, 20),
                                       (2, 32, 8, 9, 48), (1, 32, 3, 30, 128)])
def test_maxpool333_tile_kernel(n, c, t, h, w):
    """Inception pooling branch (3,3,3)/1 pad 1 (s3d.py:134) from a shared-memory tile: 16-channel groups, all frames of a row
    strip per block, partial last strips, channel-slice strides on both sides; selection only, so bit-exact against
    F.max_pool3d (the sliding-window kernel it replaces stays reachable with MSPI_POOL_TILE=0 and for odd channel-group
    counts, test_maxpool_variants)."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, c, t, h, w, generator=g)
    xa = _act_from_ncdhw(x, c + 16, 8)
    ref = F.max_pool3d(_bf(x), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    buf = torch.zeros(n, t, h, w, c + 24, dtype=torch.bfloat16, device="cuda")
    ya = Act(buf, 16, c)
    ops.maxpool3d(xa, ya, (3, 3, 3), (1, 1, 1), (1, 1, 1))()
    torch.cuda.synchronize()
    assert torch.equal(ya.to_ncdhw().cpu(), ref)
    assert (buf[..., :16] == 0).all() and (buf[..., 16 + c:] == 0).all()     # neighbouring channels untouched


@pytest.mark.parametrize("k", [2, 4, 8])
def test_upsample_bilinear(k):
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 16, 3, 7, 12, generator=g)
    xa = _act_from_ncdhw(x)
    ref = F.interpolate(_bf(x), scale_factor=(1, k, k), mode="trilinear", align_corners=False)
    ya = Act.zeros(2, 3, 7 * k, 12 * k, 16, dtype=torch.float32)
    ops.upsample(xa, ya, k)()
    torch.cuda.synchronize()
    assert (ya.to_ncdhw().cpu() - ref).abs().max() < 1e-5  # fp32 interpolation weights, exact powers of two
    # accumulate into a bf16 buffer
    base = torch.randn(2, 16, 3, 7 * k, 12 * k, generator=g)
    yb = _act_from_ncdhw(base)
    ops.upsample(xa, yb, k, accumulate=True)()
    torch.cuda.synchronize()
    assert (yb.to_ncdhw().cpu() - (ref + _bf(base))).abs().max() < 4e-2  # one bf16 rounding of values up to ~6


@pytest.mark.parametrize("c,kern", [(192, (7, 1, 1)), (192, (1, 7, 7)), (96, (1, 7, 7)), (768, (1, 7, 7))])
def test_dwconv_ln(c, kern):
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(4)
    x = torch.randn(1, c, 4, 9, 11, generator=g)
    w = torch.randn(c, 1, *kern, generator=g) * 0.2
    b = torch.randn(c, generator=g) * 0.1
    lw, lb = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    xa = _act_from_ncdhw(x)
    pad = tuple(k // 2 for k in kern)
    conv = F.conv3d(_bf(x), w, b, padding=pad, groups=c)
    ya = Act.empty(1, 4, 9, 11, c, dtype=torch.float32)
    ops.dwconv_ln(xa, ya, w, b)()
    torch.cuda.synchronize()
    assert (ya.to_ncdhw().cpu() - conv).abs().max() < 1e-4  # fp32 math, order only
    ref = F.layer_norm(conv.permute(0, 2, 3, 4, 1), (c,), lw, lb, 1e-6).permute(0, 4, 1, 2, 3)
    ops.dwconv_ln(xa, ya, w, b, lw, lb, eps=1e-6)()
    torch.cuda.synchronize()
    assert (ya.to_ncdhw().cpu() - ref).abs().max() < 2e-4


@pytest.mark.parametrize("c,h,w,dt", [(96, 16, 32, torch.bfloat16), (384, 6, 24, torch.bfloat16), (192, 7, 12, torch.float32),
                                      (40, 5, 16, torch.float32), (768, 7, 12, torch.bfloat16), (384, 14, 24, torch.bfloat16),
                                      (768, 2, 2, torch.bfloat16), (192, 10, 20, torch.bfloat16), (96, 13, 21, torch.bfloat16),
                                      (192, 28, 48, torch.bfloat16)])
def test_dwconv7x7_ln_strip_kernel(c, h, w, dt):
    """Register-tiled 7x7 depthwise + LayerNorm (dwconv.cu): strip lengths 16/12/8, bf16 and fp32 inputs, bf16 output."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(14)
    x = torch.randn(3, c, 1, h, w, generator=g)
    wgt = torch.randn(c, 1, 7, 7, generator=g) * 0.2
    b = torch.randn(c, generator=g) * 0.1
    lw, lb = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    xa = _act_from_ncdhw(x, dtype=dt)
    xr = _bf(x) if dt == torch.bfloat16 else x
    conv = F.conv3d(xr, wgt[:, :, None], b, padding=(0, 3, 3), groups=c)
    ref = F.layer_norm(conv.permute(0, 2, 3, 4, 1), (c,), lw, lb, 1e-6).permute(0, 4, 1, 2, 3)
    y32 = Act.empty(3, 1, h, w, c, dtype=torch.float32)
    ops.dwconv_ln(xa, y32, wgt, b, lw, lb, eps=1e-6)()
    torch.cuda.synchronize()
    assert (y32.to_ncdhw().cpu() - ref).abs().max() < 2e-4  # fp32 math, summation order only
    y16 = Act.empty(3, 1, h, w, c, dtype=torch.bfloat16)
    ops.dwconv_ln(xa, y16, wgt, b, lw, lb, eps=1e-6)()
    torch.cuda.synchronize()
    # one bf16 rounding of values up to ~5; the channel-grouped path (C = 384 / 768) also rounds the convolution to bf16
    # before the LayerNorm kernel
    assert (y16.to_ncdhw().cpu() - ref).abs().max() < 4e-2


@pytest.mark.parametrize("c", [64, 96, 192, 384, 768, 1024, 100])
@pytest.mark.parametrize("idt,odt", [(torch.bfloat16, torch.bfloat16), (torch.float32, torch.bfloat16),
                                      (torch.float32, torch.float32), (torch.bfloat16, torch.float32)])
def test_layernorm_rows(c, idt, odt):
    """Row LayerNorm over the channel vector (convnext.py LayerNorm channels_last / timm LayerNorm2d): the 8-channels-per-lane
    kernel (C % 8 == 0) and the 4-channel fallback (C = 100), ragged row count, also in place."""
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(50 + c)
    rows = 1037
    x = (torch.randn(rows, c, generator=g) * 2 + 0.7).to(idt)
    w, bb = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    xd = x.cuda()
    y = torch.empty(rows, c, dtype=odt, device="cuda")
    ops.layernorm(xd, y, rows, c, w, bb, 1e-6)()
    torch.cuda.synchronize()
    ref = F.layer_norm(x.float(), (c,), w, bb, 1e-6)
    tol = 3e-2 if odt == torch.bfloat16 else 2e-5
    assert (y.float().cpu() - ref).abs().max() < tol
    if idt == odt:
        ops.layernorm(xd, xd, rows, c, w, bb, 1e-6)()
        torch.cuda.synchronize()
        assert torch.equal(xd, y)


def test_layernorm_pos_groups():
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(5)
    b, nv, na, c = 2, 10, 4, 512
    xv = torch.randn(b * nv, c, generator=g)
    w, bb = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.1
    pos = torch.randn(nv, c, generator=g)
    out = torch.zeros(b, nv + na, c, dtype=torch.bfloat16, device="cuda")
    ops.layernorm(xv.cuda(), out, b * nv, c, w, bb, 1e-5, pos=pos.cuda(), rows_per_group=nv, out_gstride=(nv + na) * c)()
    torch.cuda.synchronize()
    ref = F.layer_norm(xv, (c,), w, bb, 1e-5).view(b, nv, c) + pos
    assert (out[:, :nv].float().cpu() - ref).abs().max() < 3e-2  # bf16 output of values up to ~5
    assert (out[:, nv:] == 0).all()


def test_attention_small():
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(6)
    b, n, heads, hd = 2, 52, 4, 128
    qkv = torch.randn(b, n, 3, heads, hd, generator=g)
    out = torch.empty(b, n, heads * hd, dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv.to(torch.bfloat16).cuda(), out, b, n, heads, hd)()
    torch.cuda.synchronize()
    q, k, v = _bf(qkv).permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(b, n, heads * hd)
    assert (out.float().cpu() - ref).abs().max() < 2e-2  # bf16 output rounding


@pytest.mark.parametrize("b,n", [(2, 52), (1, 372), (1, 1380)])
def test_attention_as_batched_gemms(b, n):
    """scores = scale Q K^T and out = P V as batched kind::tf32 GEMMs + row softmax (model_utils.py:97-109); N = 372 is
    the S3D/SlowFast token count at 224x384, 1380 the X3D-L one.  tf32 operands (10-bit mantissa), fp32 accumulate."""
    from mspi_b200 import ops
    g = torch.Generator().manual_seed(16)
    heads, hd = 4, 128
    qkv = torch.randn(b, n, 3, heads, hd, generator=g)
    out = torch.empty(b * n, heads * hd, dtype=torch.float32, device="cuda")
    for _nm, fn in ops.attention_gemm(qkv.view(b * n, -1).cuda(), out, b, n, heads, hd):
        fn()
    torch.cuda.synchronize()
    q, k, v = qkv.permute(2, 0, 3, 1, 4)
    ref = ((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(-1) @ v
    ref = ref.transpose(1, 2).reshape(b * n, heads * hd)
    assert _rel(out.cpu(), ref) < 5e-3


@pytest.mark.parametrize("c,m", [(96, 128 * 5 + 37), (192, 128 * 3 + 5), (96, 128 * 300)])
def test_mlp_fused(c, m):
    """Fused ConvNeXt MLP (fc1 + GELU + fc2 + layer scale + residual, hidden tile on chip) against PyTorch on
    bf16-rounded operands; the hidden activation is rounded to bf16 exactly like the unfused path does."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(41)
    x = torch.randn(m, c, generator=g)
    r = torch.randn(m, c, generator=g)
    w1 = torch.randn(4 * c, c, generator=g) / c ** 0.5
    b1 = torch.randn(4 * c, generator=g) * 0.1
    w2 = torch.randn(c, 4 * c, generator=g) / (4 * c) ** 0.5
    b2 = torch.randn(c, generator=g) * 0.1
    gamma = torch.rand(c, generator=g) + 0.5
    xa = Act(x.to(torch.bfloat16).cuda().view(1, 1, 1, m, c))
    ra = Act(r.to(torch.bfloat16).cuda().view(1, 1, 1, m, c))
    ya = Act(torch.full((1, 1, 1, m, c), 7.0, dtype=torch.bfloat16, device="cuda"))
    ops.mlp_fused(xa, ya, ra, w1, b1, w2, b2, gamma)()
    torch.cuda.synchronize()
    h = _bf(F.gelu(_bf(x) @ _bf(w1).t() + b1))
    ref = _bf(r) + gamma * (h @ _bf(w2).t() + b2)
    got = ya.buf.view(m, c).float().cpu()
    assert _rel(got, ref) < BF16_TOL
    # in place (y aliases the residual, as the ConvNeXt block's `x = input + drop_path(x)` allows): every tile reads its
    # residual rows before it writes them
    ops.mlp_fused(xa, ra, ra, w1, b1, w2, b2, gamma)()
    torch.cuda.synchronize()
    assert torch.equal(ra.buf.view(m, c), ya.buf.view(m, c))


@pytest.mark.parametrize("c,m", [(96, 128 * 5 + 37), (192, 128 * 3 + 5), (96, 128 * 300), (192, 128 * 299 + 1)])
def test_mlp_fused_with_layernorm_store(c, m):
    """Last block of ConvNeXt stages 0 / 1: the next stage's downsample.0 LayerNorm2d is applied to the bf16-rounded block
    output before it is stored (mspi_mlp_fused_ln).  Reference: the plain fused kernel followed by the LayerNorm kernel the
    plan used before (both already tested against PyTorch), and PyTorch's layer_norm on the same rounded rows."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(43)
    x = torch.randn(m, c, generator=g)
    r = torch.randn(m, c, generator=g) * 3 + 0.7
    w1 = torch.randn(4 * c, c, generator=g) / c ** 0.5
    b1 = torch.randn(4 * c, generator=g) * 0.1
    w2 = torch.randn(c, 4 * c, generator=g) / (4 * c) ** 0.5
    b2 = torch.randn(c, generator=g) * 0.1
    gamma = torch.rand(c, generator=g) + 0.5
    lw, lb = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.2
    xa = Act(x.to(torch.bfloat16).cuda().view(1, 1, 1, m, c))
    ra = Act(r.to(torch.bfloat16).cuda().view(1, 1, 1, m, c))
    y0 = Act(torch.zeros((1, 1, 1, m, c), dtype=torch.bfloat16, device="cuda"))
    y1 = Act(torch.full((1, 1, 1, m, c), 7.0, dtype=torch.bfloat16, device="cuda"))
    y2 = Act(torch.full((1, 1, 1, m, c), 7.0, dtype=torch.bfloat16, device="cuda"))
    ops.mlp_fused(xa, y0, ra, w1, b1, w2, b2, gamma)()
    ops.layernorm(y0.buf, y1.buf, m, c, lw, lb, 1e-6)()
    ops.mlp_fused(xa, y2, ra, w1, b1, w2, b2, gamma, ln=(lw, lb, 1e-6))()
    torch.cuda.synchronize()
    two_step = y1.buf.view(m, c).float().cpu()
    got = y2.buf.view(m, c).float().cpu()
    ref = F.layer_norm(y0.buf.view(m, c).float().cpu(), (c,), lw, lb, 1e-6)
    assert _rel(got, ref) < BF16_TOL and _rel(two_step, ref) < BF16_TOL
    # same statistics up to fp32 summation order: the two results differ by at most one bf16 step on a few elements
    d = (got - two_step).abs()
    assert d.max() <= 2 ** -6 * max(1.0, ref.abs().max().item()) and (d > 0).float().mean() < 0.02, (d.max(), (d > 0).float().mean())


def test_sa_gate_token_mean_simsiam():
    from mspi_b200 import _lib, ops
    from mspi_b200.ops import Act
    lib = _lib.load()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(1, 192, 2, 5, 6, generator=g)
    m = torch.randn(1 * 2 * 5 * 6, generator=g)
    xa = _act_from_ncdhw(x)
    ya = Act.empty(1, 2, 5, 6, 192)
    ops.sa_gate(xa, m.cuda(), ya)()
    torch.cuda.synchronize()
    ref = _bf(x) * (1 + torch.sigmoid(m).view(1, 1, 2, 5, 6))
    assert (ya.to_ncdhw().cpu() - ref).abs().max() < 4e-2
    # token mean
    t = torch.randn(3, 20, 512, generator=g).to(torch.bfloat16)
    y = torch.empty(3, 512, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    td = t.cuda()
    _lib.check(lib.mspi_token_mean(C.c_void_p(td.data_ptr()), 0, C.c_void_p(y.data_ptr()), 3, 20, 4, 17, 512, st))
    torch.cuda.synchronize()
    assert (y.cpu() - t.float()[:, 4:17].mean(1)).abs().max() < 1e-5
    # simsiam
    pv, za, pa, zv = (torch.randn(3, 2048, generator=g) for _ in range(4))
    o = torch.empty(1, device="cuda")
    dev = [v.cuda() for v in (pv, za, pa, zv)]
    _lib.check(lib.mspi_simsiam_loss(*(C.c_void_p(v.data_ptr()) for v in dev), C.c_void_p(o.data_ptr()), 3, 2048, st))
    torch.cuda.synchronize()
    ref = -0.5 * (F.cosine_similarity(pv, za, dim=-1).mean() + F.cosine_similarity(pa, zv, dim=-1).mean())
    assert abs(o.item() - ref.item()) < 1e-5


def test_metrics_kernel_against_golden_and_oracle():
    """Metric kernel on identical fp32 maps: KATs produced by the reference's own functions (tests/golden/metrics.pt)
    within 1e-5 relative, far inside the 1e-3 budget."""
    import os
    from mspi_b200.utils import compute_saliency_metrics as m
    from mspi_b200.utils.loss import SalLoss
    cases = torch.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.pt"), weights_only=False)
    for nm, c in cases.items():
        s, g, f = c["s"].cuda(), c["gt"].cuda(), c["fix"].cuda()
        for key, val in (("kld", m.kldiv(s, g)), ("cc", m.cc(s, g)), ("sim", m.similarity(s, g)), ("nss", m.nss(s, f))):
            assert abs(val.item() - c[key]) <= 1e-5 * max(1.0, abs(c[key])), (nm, key, val.item(), c[key])
        logp = torch.log(c["s"] / c["s"].sum((1, 2), keepdim=True)).cuda()
        crit = SalLoss()
        assert abs(crit(logp, g).item() - c["loss"]) <= 2e-5 * max(1.0, abs(c["loss"]))
        assert abs(crit(logp, g, f).item() - c["loss_fix"]) <= 2e-5 * max(1.0, abs(c["loss_fix"]))
        assert crit.log["loss"].count == 2


def test_logsoftmax_full_size():
    from mspi_b200 import ops
    x = torch.randn(3, 224 * 384, generator=torch.Generator().manual_seed(8)) * 3
    y = torch.empty(3, 224 * 384, device="cuda")
    ops.logsoftmax2d(x.cuda(), y, 3, 224 * 384)()
    torch.cuda.synchronize()
    ref = x - torch.logsumexp(x, 1, keepdim=True)
    assert (y.cpu() - ref).abs().max() < 2e-5


def test_logspec_against_reference_fixtures():
    import os
    from mspi_b200.audio import log_spectrogram
    fx = torch.load(os.path.join(os.path.dirname(__file__), "golden", "audio.pt"), weights_only=False)
    for nm, c in fx.items():
        got = log_spectrogram(c["wave"].cuda())[0].cpu()
        assert got.shape == c["feat"].shape
        # fp32 direct DFT vs torch's fp32 FFT, after log and standardisation: 2e-3 absolute on O(1) values
        assert (got - c["feat"]).abs().max().item() < 2e-3, (nm, (got - c["feat"]).abs().max().item())
    # ragged / edge cases: a batch, a signal shorter than one hop beyond the half window
    w = torch.randn(3, 400, generator=torch.Generator().manual_seed(4))
    from oracle import mspi_oracle as orc
    assert (log_spectrogram(w.cuda()).cpu() - orc.log_spectrogram(w)).abs().max() < 2e-3


def test_postprocess_maps_against_cv2():
    """GPU blur/exp/resize/normalise/uint8 vs the reference's cv2 + numpy sequence (inference.py:84-91).  uint8 output:
    float summation order can flip a rounding at an exact .5 boundary, so at most 1 LSB on a small fraction of pixels."""
    pytest.importorskip("cv2")
    from oracle import mspi_oracle as orc
    from mspi_b200.postprocess import postprocess_maps
    g = torch.Generator().manual_seed(33)
    yy, xx = torch.meshgrid(torch.arange(224.), torch.arange(384.), indexing="ij")
    maps = []
    for i in range(3):
        m = torch.zeros(224, 384)
        for _ in range(4):
            cy, cx = torch.rand(1, generator=g) * 224, torch.rand(1, generator=g) * 384
            m += torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * 25.0 ** 2))
        m = m + 0.05 * torch.rand(224, 384, generator=g)
        maps.append(torch.log(m / m.sum()))
    x = torch.stack(maps)
    got = postprocess_maps(x.cuda(), (640, 480)).cpu().numpy()
    assert got.shape == (3, 480, 640) and got.dtype.name == "uint8"
    for i in range(3):
        ref = orc.postprocess(x[i].numpy(), (640, 480))
        d = abs(got[i].astype(int) - ref.astype(int))
        assert d.max() <= 1 and (d > 0).mean() < 0.02, (d.max(), (d > 0).mean())
        assert got[i].max() == 255 and got[i].min() == 0


@pytest.mark.parametrize("scales,w,mask,inplace", [((2, 4, 8), 24, True, False), ((2, 4), 24, True, False),
                                                   ((2,), 8, True, False), ((2, 4, 8), 40, False, True),
                                                   ((4, 2), 24, True, False), ((2, 4, 8), 8, True, False),
                                                   ((2,), 6, False, False), ((), 24, True, False)])
def test_sa_gate_fused_with_topdown_sums(scales, w, mask, inplace):
    """y = x*sigmoid(l) + x + up2(a) + up4(b) + up8(c) in one kernel (model_utils.py:167-170,566-568) vs PyTorch: the
    scale-ladder kernel (2, 4, 8 in order, width a multiple of 4: four pixels per thread, border columns clamped), the row
    kernel for everything else, the mask-free in-place form readout.0 uses."""
    from mspi_b200 import ops
    from mspi_b200.ops import Act
    g = torch.Generator().manual_seed(8)
    n, t, h, c = 2, 2, 16, 16
    x = torch.randn(n, c, t, h, w, generator=g)
    l = torch.randn(n, 1, t, h, w, generator=g)
    srcs = [(torch.randn(n, c, t, h // k, w // k, generator=g), k) for k in scales]
    ref = x * torch.sigmoid(l) + x if mask else x.clone()
    for a, k in srcs:
        ref = ref + F.interpolate(a, scale_factor=(1, k, k), mode="trilinear", align_corners=False)
    xa = _act_from_ncdhw(x, 24, 8, dtype=torch.float32)
    ya = xa if inplace else Act(torch.zeros(n, t, h, w, c, device="cuda"))
    sa = [(_act_from_ncdhw(a, dtype=torch.float32), k) for a, k in srcs]
    lg = l.permute(0, 2, 3, 4, 1).contiguous().cuda().view(-1) if mask else None
    ops.sa_gate_fused(xa, lg, ya, sa)()
    torch.cuda.synchronize()
    assert (ya.to_ncdhw().cpu() - ref).abs().max() < 1e-5
