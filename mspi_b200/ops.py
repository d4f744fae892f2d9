"""Host-side operator layer: turns PyTorch tensors + layer geometry into C-ABI calls.

PyTorch is used for device memory and streams only; every arithmetic op below is a call into
``libmspi_b200.so`` (hand-written sm_100a kernels).  Activations are channels-last
([N, T, H, W, C], C contiguous) views described by :class:`Act`; a view may be a channel slice
of a wider buffer, which is how concatenations (torch.cat in the reference, e.g.
backbones/s3d.py:143, model/model_utils.py:198,559,570) are produced without a copy.
"""
from __future__ import annotations

import ctypes as C
import functools
import math
import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (ACT_GELU, ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SWISH, MSPI_BF16, MSPI_F32, ConvDesc, Dw3dDesc, DwDesc,
                   LnDesc, PatchDesc, PoolDesc, UpDesc)

_DT = {torch.bfloat16: MSPI_BF16, torch.float32: MSPI_F32}
_ES = {torch.bfloat16: 2, torch.float32: 4}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor], byte_off: int = 0) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    return C.c_void_p(t.data_ptr() + byte_off)


class Act:
    """Channels-last activation view: channels [c0, c0+c) of a [N,T,H,W,Cs] buffer."""

    __slots__ = ("buf", "n", "t", "h", "w", "cs", "c0", "c")

    def __init__(self, buf: torch.Tensor, c0: int = 0, c: Optional[int] = None):
        assert buf.dim() == 5 and buf.is_contiguous(), "Act needs a contiguous [N,T,H,W,C] buffer"
        self.buf = buf
        self.n, self.t, self.h, self.w, self.cs = buf.shape
        self.c0 = c0
        self.c = self.cs - c0 if c is None else c
        assert 0 <= c0 and c0 + self.c <= self.cs

    @staticmethod
    def empty(n, t, h, w, c, dtype=torch.bfloat16, device="cuda") -> "Act":
        return Act(torch.empty((n, t, h, w, c), dtype=dtype, device=device))

    @staticmethod
    def zeros(n, t, h, w, c, dtype=torch.bfloat16, device="cuda") -> "Act":
        return Act(torch.zeros((n, t, h, w, c), dtype=dtype, device=device))

    def slice(self, c0: int, c: int) -> "Act":
        return Act(self.buf, self.c0 + c0, c)

    @property
    def dtype(self):
        return self.buf.dtype

    @property
    def ptr(self) -> C.c_void_p:
        return _ptr(self.buf, self.c0 * _ES[self.buf.dtype])

    @property
    def pixels(self) -> int:
        return self.n * self.t * self.h * self.w

    def rows(self) -> torch.Tensor:
        """[pixels, c] strided torch view (for tests / hand-off)."""
        return self.buf.view(-1, self.cs)[:, self.c0:self.c0 + self.c]

    def to_ncdhw(self) -> torch.Tensor:
        """fp32 NCDHW copy (the reference's layout) made by the CUDA transpose kernel."""
        out = torch.empty((self.n, self.c, self.t, self.h, self.w), dtype=torch.float32, device=self.buf.device)
        lib = _lib.load()
        _lib.check(lib.mspi_ndhwc_to_ncdhw(self.ptr, _DT[self.dtype], self.cs, _ptr(out), self.n, self.c,
                                           self.t * self.h * self.w, _stream()), "ndhwc_to_ncdhw")
        return out


# ------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def choose_box(dims: Tuple[int, int, int, int], max_rows: int = 128) -> Tuple[int, int, int, int]:
    """Pick the output-position box (b1..b4), b1*b2*b3*b4 <= 128, that wastes the fewest MMA rows.

    Cost = number of tiles (each tile costs one 128-row MMA pass whatever its fill); ties prefer
    a long innermost run (contiguous TMA rows)."""
    best, best_key = None, None

    def cands(d):
        out = {d} if d <= max_rows else set()
        for b in range(1, min(d, max_rows) + 1):
            # only sizes that tile d evenly, or powers of two / the remainder-minimal splits
            if d % b == 0 or (b & (b - 1)) == 0:
                out.add(b)
        k = 1
        while k <= d:  # ceil splits: d/2, d/3, ...
            b = -(-d // k)
            if b <= max_rows:
                out.add(b)
            k += 1
            if k > 64:
                break
        return sorted(out)

    c = [cands(d) for d in dims]
    for b1 in c[0]:
        for b2 in c[1]:
            if b1 * b2 > max_rows:
                break
            for b3 in c[2]:
                if b1 * b2 * b3 > max_rows:
                    break
                for b4 in c[3]:
                    if b1 * b2 * b3 * b4 > max_rows:
                        break
                    tiles = 1
                    for d, b in zip(dims, (b1, b2, b3, b4)):
                        tiles *= -(-d // b)
                    key = (tiles, -b1, -b2, -b3)
                    if best_key is None or key < best_key:
                        best, best_key = (b1, b2, b3, b4), key
    return best


@functools.lru_cache(maxsize=None)
def choose_box_k(dims: Tuple[int, int, int, int], mult: int, max_rows: int = 128) -> Tuple[int, int, int, int]:
    """Position box for the weight-gradient GEMM: the box is the MMA's K extent, so its volume must be a multiple of
    `mult` (one MMA K step: 16 bf16 / 8 tf32 positions); it may overhang the tensor (TMA zero-fills dY there).
    Fewest boxes first, then the least overhang."""
    best, best_key = None, None

    def cands(d):
        out = set()
        for b in range(1, min(max_rows, 2 * d) + 1):
            if b <= d or (b & (b - 1)) == 0 or b % mult == 0:
                out.add(b)
        return sorted(out)

    c = [cands(d) for d in dims]
    for b1 in c[0]:
        for b2 in c[1]:
            if b1 * b2 > max_rows:
                break
            for b3 in c[2]:
                if b1 * b2 * b3 > max_rows:
                    break
                for b4 in c[3]:
                    v = b1 * b2 * b3 * b4
                    if v > max_rows:
                        break
                    if v % mult:
                        continue
                    tiles = 1
                    for d, b in zip(dims, (b1, b2, b3, b4)):
                        tiles *= -(-d // b)
                    key = (tiles, tiles * v, -b1)
                    if best_key is None or key < best_key:
                        best, best_key = (b1, b2, b3, b4), key
    assert best is not None
    return best


def choose_bn(cout: int, chunk: int = 64) -> int:
    """N tile of the UMMA (multiple of 16, <= 256).  With several N tiles every tile must end on a 128-byte
    output chunk (`chunk` columns: 64 bf16 / 32 fp32) so the epilogue can bulk-store whole chunks."""
    if cout <= 256:
        return max(16, -(-cout // 16) * 16)
    best = None
    for bn in range(256, 127, -chunk):
        key = (-(-cout // bn) * bn, -bn)  # least padded work, then the widest tile
        if best is None or key < best[0]:
            best = (key, bn)
    return best[1]


def pack_conv_weight(w: torch.Tensor, dtype: torch.dtype) -> Tuple[torch.Tensor, int, int]:
    """[Cout, Cin, kt, kh, kw] fp32 -> K-major GEMM matrix [rows16, taps*cin_pad] (tap-major K,
    zero padded so that one tap's K extent is a whole number of 128-byte chunks)."""
    cout, cin, kt, kh, kw = w.shape
    bk = 128 // _ES[dtype]
    cin_pad = -(-cin // bk) * bk
    rows = -(-cout // 16) * 16
    taps = kt * kh * kw
    m = torch.zeros((rows, taps, cin_pad), dtype=torch.float32, device=w.device)
    m[:cout, :, :cin] = w.permute(0, 2, 3, 4, 1).reshape(cout, taps, cin)
    return m.reshape(rows, taps * cin_pad).to(dtype).contiguous(), taps, cin_pad


def fold_bn(bn_w, bn_b, bn_mean, bn_var, eps: float, conv_bias=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm (+ optional conv bias) as per-channel scale/shift applied to the fp32 accumulator."""
    scale = bn_w.float() / torch.sqrt(bn_var.float() + eps)
    shift = bn_b.float() - bn_mean.float() * scale
    if conv_bias is not None:
        shift = shift + conv_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


class Conv:
    """A convolution / linear layer with its epilogue, prepared once and planned per input view.

    weight: [Cout, Cin, kt, kh, kw] (Conv3d), [Cout, Cin, kh, kw] (Conv2d) or [Cout, Cin] (Linear), fp32.
    y = act(scale * conv(x) + shift (+res))  or  act(scale*conv(x)+shift) + res  (res_after_act).
    """

    def __init__(self, weight: torch.Tensor, scale: Optional[torch.Tensor] = None,
                 shift: Optional[torch.Tensor] = None, stride=(1, 1, 1), pad=(0, 0, 0), act: int = ACT_NONE,
                 dtype: torch.dtype = torch.bfloat16, res_after_act: bool = False, device="cuda", name: str = "",
                 split_weights: bool = False, cache: Optional[dict] = None):
        w = weight.detach().float()
        self.cache = cache
        if w.dim() == 2:
            w = w[:, :, None, None, None]
        elif w.dim() == 4:
            w = w[:, :, None]
        self.name = name
        self.cout, self.cin, self.kt, self.kh, self.kw = w.shape
        self.stride = tuple(stride) if len(stride) == 3 else (1,) + tuple(stride)
        self.pad = tuple(pad) if len(pad) == 3 else (0,) + tuple(pad)
        self.act = act
        self.dtype = dtype
        self.res_after_act = res_after_act
        # split_weights: W = bf16(W) + bf16(W - bf16(W)); the low half rides along as a second set of taps at the
        # same offsets, so a bf16 activation meets ~16-bit-mantissa weights at 2x the MMA work (used where the
        # result is min-max normalised downstream and the input is already exact in bf16).
        self.split = split_weights and dtype == torch.bfloat16
        self.device = device
        self.w_raw = w
        self.packed = None  # built lazily, the packing depends on the execution mode
        self.scale = None if scale is None else scale.detach().float().contiguous().to(device)
        self.shift = None if shift is None else shift.detach().float().contiguous().to(device)

    # -- geometry ---------------------------------------------------------------------------
    def out_shape(self, t, h, w):
        (st, sh, sw), (pt, ph, pw) = self.stride, self.pad
        return ((t + 2 * pt - self.kt) // st + 1, (h + 2 * ph - self.kh) // sh + 1, (w + 2 * pw - self.kw) // sw + 1)

    def _mode(self, x: Act) -> str:
        st, sh, sw = self.stride
        if x.c % 8 != 0 or x.cs % 8 != 0 or x.c0 % 8 != 0 or x.dtype != self.dtype:
            return "gather"
        if (st, sh, sw) == (1, 1, 1):
            return "shift"
        if sh == 1 and sw == 1 and self.kh == 1 and self.kw == 1 and x.t % st == 0:
            return "tstride"
        if (self.kt, self.kh, self.kw) == (1, 1, 1) and st == 1 and self.pad == (0, 0, 0):
            return "pick"  # 1x1x1 conv with spatial stride (ResBlock.branch1, resnet_helper.py:556-566): a strided view
        if (self.kt, self.kh, self.kw) == (1, 2, 2) and (st, sh, sw) == (1, 2, 2) and self.pad == (0, 0, 0) \
                and x.h % 2 == 0 and x.w % 2 == 0:
            return "patch2"  # non-overlapping 2x2/s2 patches (ConvNeXt downsample, timm convnext): the row / column parity
            #                   of the input is its own tensor axis, each of the 4 taps a box at parity (kh, kw)
        return "gather"

    def _pack(self, w5: torch.Tensor):
        """Packing runs where the weights live (host tensors for an inference plan: no kernel launches, only the packed
        matrix is uploaded; device tensors for the training plan) and is cached per (layer, layout) in `self.cache`."""
        key = (self.name, str(self.dtype), self.split, tuple(w5.shape))
        if self.cache is not None and key in self.cache:
            return self.cache[key]
        if not self.split:
            packed, taps, cin_pad = pack_conv_weight(w5, self.dtype)
        else:
            hi = w5.to(torch.bfloat16).float()
            p_hi, taps, cin_pad = pack_conv_weight(hi, self.dtype)
            p_lo, _, _ = pack_conv_weight(w5 - hi, self.dtype)
            packed, taps = torch.cat([p_hi, p_lo], 1).contiguous(), 2 * taps
        res = (packed.to(self.device), taps, cin_pad)
        if self.cache is not None:
            self.cache[key] = res
        return res

    def _geom(self, mode: str, x: Act, ot: int, oh: int, ow: int):
        """TMA view of the input (dims/strides, inner -> outer), output extents, per-tap box offsets and the output
        row-stride function for the implicit-GEMM modes (shared by the forward plan and the weight-gradient plan)."""
        (st, sh, sw), (pt, ph, pw) = self.stride, self.pad
        if mode == "shift":
            a_dims = (x.c, x.w, x.h, x.t, x.n)
            a_str = (1, x.cs, x.w * x.cs, x.h * x.w * x.cs, x.t * x.h * x.w * x.cs)
            o_dims = (ow, oh, ot, x.n)
            offs = [(kw - pw, kh - ph, kt - pt, 0) for kt in range(self.kt) for kh in range(self.kh) for kw in range(self.kw)]
            ostr = lambda a: (a.cs, a.w * a.cs, a.h * a.w * a.cs, a.t * a.h * a.w * a.cs)
        elif mode == "pick":
            a_dims = (x.c, ow, oh, x.t, x.n)
            a_str = (1, sw * x.cs, sh * x.w * x.cs, x.h * x.w * x.cs, x.t * x.h * x.w * x.cs)
            o_dims = (ow, oh, ot, x.n)
            offs = [(0, 0, 0, 0)]
            ostr = lambda a: (a.cs, a.w * a.cs, a.h * a.w * a.cs, a.t * a.h * a.w * a.cs)
        elif mode == "tstride":
            hw = x.h * x.w
            a_dims = (x.c, hw, st, x.t // st, x.n)
            a_str = (1, x.cs, hw * x.cs, st * hw * x.cs, x.t * hw * x.cs)
            o_dims = (hw, 1, ot, x.n)
            offs = []
            for kt in range(self.kt):
                dlt = kt - pt
                q = dlt // st  # floor
                offs.append((0, dlt - q * st, q, 0))
            ostr = lambda a: (a.cs, 0, a.h * a.w * a.cs, a.t * a.h * a.w * a.cs)
        elif mode == "patch2":
            rows = x.n * x.t * (x.h // 2)     # (n, t, h/2) merge: stride(t) = H*W*cs = (H/2) * stride(h/2)
            a_dims = (x.c, 2, x.w // 2, 2, rows)
            a_str = (1, x.cs, 2 * x.cs, x.w * x.cs, 2 * x.w * x.cs)
            o_dims = (1, ow, 1, rows)
            offs = [(kw, 0, kh, 0) for kh in range(2) for kw in range(2)]
            ostr = lambda a: (0, a.cs, 0, a.w * a.cs)
        else:
            raise ValueError(mode)
        return a_dims, a_str, o_dims, offs, ostr

    def wgrad_plan(self, x: Act, dy: Act, dw: torch.Tensor) -> Callable[[], None]:
        """dw (fp32, the layer's own [Cout, Cin, kt, kh, kw] / [Cout, Cin] gradient tensor, contiguous) +=
        sum_p dy[p] (x) x[p + tap]  on the tensor cores (mspi_conv_wgrad).  x and dy share the layer dtype."""
        lib = _lib.load()
        ot, oh, ow = self.out_shape(x.t, x.h, x.w)
        assert (dy.n, dy.t, dy.h, dy.w, dy.c) == (x.n, ot, oh, ow, self.cout) and x.c == self.cin
        assert x.dtype == dy.dtype == self.dtype and dw.dtype == torch.float32 and dw.is_contiguous()
        assert dw.numel() == self.cout * self.cin * self.kt * self.kh * self.kw
        mode = self._mode(x)
        assert mode in ("shift", "pick", "tstride"), f"{self.name}: wgrad needs an implicit-GEMM layer, got {mode}"
        es = _ES[self.dtype]
        assert (dy.cs * es) % 16 == 0 and (dy.c0 * es) % 16 == 0, f"{self.name}: dy slice not 16-byte aligned"
        a_dims, a_str, o_dims, offs, ostr = self._geom(mode, x, ot, oh, ow)
        d = ConvDesc()
        d.a_dtype = d.o_dtype = _DT[self.dtype]
        d.cout = self.cout
        # tf32 chunks are 128 B per position like bf16 ones but hold half the channels: 64-position stages keep >= 3
        # pipeline stages in shared memory at the widest tiles (128 x 128 channels)
        box = choose_box_k(tuple(o_dims), 16, 128) if self.dtype == torch.bfloat16 else choose_box_k(tuple(o_dims), 8, 64)
        for j in range(5):
            d.a_dims[j], d.a_strides[j] = a_dims[j], a_str[j]
        d.box[0] = 0
        for j in range(4):
            d.box[j + 1], d.o_dims[j], d.o_strides[j] = box[j], o_dims[j], ostr(dy)[j]
        taps = len(offs)
        d.ntaps = taps
        for i, o in enumerate(offs):
            for j in range(4):
                d.tap_off[i][j] = o[j]
        s_co, s_ci, s_tap = self.cin * taps, taps, 1
        xp, dyp, dwp = x.ptr, dy.ptr, _ptr(dw)
        name = self.name

        def run(_keep=(x.buf, dy.buf, dw, d)):
            _lib.check(lib.mspi_conv_wgrad(C.byref(d), xp, dyp, dwp, s_co, s_ci, s_tap, _stream()), f"conv_wgrad[{name}]")

        run.desc = d
        run.flops = self.flops(x)
        return run

    def plan(self, x: Act, y: Act, residual: Optional[Act] = None) -> Callable[[], None]:
        """Build the descriptor(s) for this input/output pair; returns a closure that enqueues the kernels."""
        lib = _lib.load()
        ot, oh, ow = self.out_shape(x.t, x.h, x.w)
        assert (y.n, y.t, y.h, y.w) == (x.n, ot, oh, ow), f"{self.name}: out {(y.n, y.t, y.h, y.w)} vs {(x.n, ot, oh, ow)}"
        assert y.c == self.cout and x.c == self.cin, f"{self.name}: channels x={x.c}/{self.cin} y={y.c}/{self.cout}"
        mode = self._mode(x)
        d = ConvDesc()
        d.a_dtype = _DT[self.dtype]
        self.bn = choose_bn(self.cout, 64 if y.dtype == torch.bfloat16 else 32)
        d.cout, d.bn = self.cout, self.bn
        d.o_dtype = _DT[y.dtype]
        d.act = self.act
        d.has_residual = 0 if residual is None else 1
        d.res_after_act = 1 if self.res_after_act else 0
        d.r_dtype = _DT[residual.dtype] if residual is not None else 0
        es = _ES[self.dtype]
        pre = None
        (st, sh, sw), (pt, ph, pw) = self.stride, self.pad
        if mode in ("shift", "pick", "tstride", "patch2"):
            packed, taps, cin_pad = self._pack(self.w_raw)
            a_dims, a_str, o_dims, offs, ostr = self._geom(mode, x, ot, oh, ow)
            x_ptr = x.ptr
        else:  # explicit patch gather + flat GEMM
            k = self.kt * self.kh * self.kw * self.cin
            bk = 128 // es
            k_pad = -(-k // bk) * bk
            w_flat = self.w_raw.permute(0, 2, 3, 4, 1).reshape(self.cout, k)
            packed, taps, cin_pad = self._pack(w_flat[:, :, None, None, None])
            m = x.n * ot * oh * ow
            patches = torch.empty((m, k_pad), dtype=self.dtype, device=self.device)
            assert self.dtype == torch.bfloat16, "patch gather writes bf16"
            pd = PatchDesc()
            pd.src_layout = 1
            pd.n, pd.c, pd.t, pd.h, pd.w = x.n, x.c, x.t, x.h, x.w
            pd.src_cstride = x.cs
            pd.kt, pd.kh, pd.kw = self.kt, self.kh, self.kw
            pd.st, pd.sh, pd.sw = st, sh, sw
            pd.pt, pd.ph, pd.pw = pt, ph, pw
            pd.ot, pd.oh, pd.ow = ot, oh, ow
            pd.k_pad = k_pad
            assert x.dtype == torch.bfloat16
            xp = x.ptr

            def pre():
                _lib.check(lib.mspi_patch_gather(C.byref(pd), xp, _ptr(patches), _stream()), f"{self.name}: patch_gather")

            a_dims = (k_pad, m, 1, 1, 1)
            a_str = (1, k_pad, m * k_pad, m * k_pad, m * k_pad)
            o_dims = (m, 1, 1, 1)
            offs = [(0, 0, 0, 0)]
            ostr = lambda a: (a.cs, 0, 0, 0)
            x_ptr = _ptr(patches)
            self._patches = patches
        if self.split:
            offs = offs + offs
        box = choose_box(tuple(o_dims))
        for j in range(5):
            d.a_dims[j] = a_dims[j]
            d.a_strides[j] = a_str[j]
        d.box[0] = 0
        for j in range(4):
            d.box[j + 1] = box[j]
            d.o_dims[j] = o_dims[j]
            d.o_strides[j] = ostr(y)[j]
            d.r_strides[j] = ostr(residual)[j] if residual is not None else 0
        assert len(offs) == taps and taps <= _lib.MAX_TAPS, f"{self.name}: {len(offs)} taps"
        d.ntaps = taps
        for i, o in enumerate(offs):
            for j in range(4):
                d.tap_off[i][j] = o[j]
        d.cin_pad = cin_pad
        d.w_rows = packed.shape[0]
        self.packed = packed  # keep alive
        w_ptr, sc_ptr, sh_ptr = _ptr(packed), _ptr(self.scale), _ptr(self.shift)
        y_ptr = y.ptr
        r_ptr = residual.ptr if residual is not None else C.c_void_p(0)
        keep = (packed, x.buf, y.buf, None if residual is None else residual.buf, d)
        name = self.name

        def run(_keep=keep):
            if pre is not None:
                pre()
            _lib.check(lib.mspi_conv_gemm(C.byref(d), x_ptr, w_ptr, sc_ptr, sh_ptr, r_ptr, y_ptr, _stream()),
                       f"conv_gemm[{name}]")

        run.mode = mode
        run.desc = d
        run.flops = self.flops(x)
        return run

    def flops(self, x: Act) -> float:
        ot, oh, ow = self.out_shape(x.t, x.h, x.w)
        return 2.0 * x.n * ot * oh * ow * self.cout * self.cin * self.kt * self.kh * self.kw


# ------------------------------------------------------------------------------------------ Cin=3 stems
PAD_T, PAD_L, PAD_EXTRA = 3, 4, 8   # padded frame: rows 3 above / 5 below, columns 4 left / 4 right (ops.stem_conv)


def clip_to_padded(holder: dict, key: str, frames: torch.Tensor, n: int, t: int, h: int, w: int,
                   frame_map: Optional[Sequence[int]] = None, t_pad: int = 0) -> Callable[[], None]:
    """fp32 NCDHW clip (looked up as holder[key] at run time) -> bf16 zero-padded [N*T', H+8, W+8, 4] frames.
    frame_map selects / re-orders source frames (SlowFast slow pathway); t_pad leaves that many zero frames at both
    ends of every clip (temporal padding of the SlowFast fast stem): T' = len(frame_map or range(t)) + 2*t_pad."""
    lib = _lib.load()
    hp, wp = h + PAD_EXTRA, w + PAD_EXTRA
    fm = list(range(t)) if frame_map is None else [f % t for f in frame_map]
    t_out = len(fm)
    fpc = t_out + 2 * t_pad
    assert tuple(frames.shape) == (n * fpc, hp, wp, 4) and frames.dtype == torch.bfloat16
    fp = _ptr(frames)
    arr = (C.c_int32 * t_out)(*fm)

    def run(_keep=(frames, arr)):
        if frame_map is None and t_pad == 0:
            rc = lib.mspi_clip_to_padded_nhwc4(_ptr(holder[key]), fp, n, t, h, w, PAD_T, PAD_L, hp, wp, _stream())
        else:
            rc = lib.mspi_clip_frames_to_padded_nhwc4(_ptr(holder[key]), fp, n, t, h, w, PAD_T, PAD_L, hp, wp,
                                                      C.cast(arr, C.c_void_p), t_out, fpc, t_pad, _stream())
        _lib.check(rc, "clip_to_padded_nhwc4")

    return run


IMAGENET_MEAN, IMAGENET_STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)   # timm.data.constants (inference.py:8,161)


def clip_u8_to_padded(holder: dict, key: str, frames: torch.Tensor, n: int, t: int, h: int, w: int,
                      mean=IMAGENET_MEAN, std=IMAGENET_STD, frame_map: Optional[Sequence[int]] = None,
                      t_pad: int = 0) -> Callable[[], None]:
    """uint8 frames [N,T,H,W,3] (holder[key] at run time) -> the same padded bf16 frames as clip_to_padded, with
    ToTensor + Normalize (inference.py:154-165) folded in (fp32, IEEE division: bit-identical to converting the normalised
    fp32 clip)."""
    lib = _lib.load()
    hp, wp = h + PAD_EXTRA, w + PAD_EXTRA
    fm = list(range(t)) if frame_map is None else [f % t for f in frame_map]
    t_out = len(fm)
    fpc = t_out + 2 * t_pad
    assert tuple(frames.shape) == (n * fpc, hp, wp, 4) and frames.dtype == torch.bfloat16
    fp = _ptr(frames)
    arr = (C.c_int32 * t_out)(*fm)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)

    def run(_keep=(frames, arr, m3, s3)):
        src = holder[key]
        _lib.check(lib.mspi_clip_u8_to_padded_nhwc4(_ptr(src), fp, n, t, h, w, PAD_T, PAD_L, hp, wp, C.cast(arr, C.c_void_p), t_out,
                                                    fpc, t_pad, C.cast(m3, C.c_void_p), C.cast(s3, C.c_void_p), _stream()),
                   "clip_u8_to_padded_nhwc4")

    return run


def gather_rows(holder: dict, key: str, index: torch.Tensor, dst: torch.Tensor, row_elems: int) -> Callable[[], None]:
    """dst[i] = holder[key][index[i]] over rows of `row_elems` elements (per-window views of the per-frame feature cache)."""
    lib = _lib.load()
    assert index.dtype == torch.int32 and index.is_cuda and dst.is_contiguous()
    n_rows = index.numel()
    row_bytes = row_elems * dst.element_size()
    assert dst.numel() == n_rows * row_elems

    def run(_keep=(index, dst)):
        src = holder[key]
        assert src.dtype == dst.dtype and src.is_contiguous() and src.numel() % row_elems == 0
        _lib.check(lib.mspi_gather_rows(_ptr(src), _ptr(index), _ptr(dst), n_rows, row_bytes, src.numel() // row_elems, _stream()),
                   "gather_rows")

    return run


def stem_conv_ln_ok(w: int, cout: int, k: int, stride: int, pad: int) -> bool:
    """Whether stem_conv's GEMM for this geometry is one N tile the LayerNorm epilogue takes (ConvNeXt stem: two output pixels
    per row, N = 192)."""
    if (k, stride, pad) != (4, 4, 0) or os.environ.get("MSPI_STEM_WIDE", "1") == "0":
        return False
    ow = (w + 2 * pad - k) // stride + 1
    wide = 2 if ow % 2 == 0 and cout * 2 <= 256 else 1
    return (cout * wide) % 64 == 0 and cout * wide <= 192


def stem_conv(frames: torch.Tensor, h: int, w: int, weight: torch.Tensor, scale, shift, k: int, stride: int, pad: int,
              act: int, y: "Act", name: str = "stem", clips: int = 0, allow_wide: bool = True, ln=None) -> Callable[[], None]:
    """(1,k,k)/stride conv with Cin=3 straight off the padded 4-channel frames (no im2col buffer).
    ln = (weight, bias, eps): LayerNorm over the output channels as the GEMM's epilogue (mspi_conv_gemm_ln; bf16 output);
    stem_conv_ln_ok() says whether a geometry qualifies.

    One row of the filter (k taps x 4 channels, at most 8 pixels) is a contiguous 16-byte aligned run of the frame;
    the filter rows are the GEMM's taps and each is fetched by one TMA box whose inner extent is that run
    (MspiConvDesc.k_row_bytes = 64 or 32: 64B / 32B-swizzled operand tiles).
      S3D   conv_s 7x7/s2 p3 (s3d.py:383):  run = 8 px starting at 2*ow-4 (+PAD_L), 7 taps, frame row 2*(oh+j)+r for
                                            filter row kh = 2j+r: the row parity r is its own tensor axis
      ConvNeXt stem 4x4/s4 p0 (timm):       run = 4 px starting at 4*ow   (+PAD_L), 4 taps, frame row 4*oh+kh
    """
    lib = _lib.load()
    nf, hp, wp, _ = frames.shape
    cout = weight.shape[0]
    w5 = weight.detach().float()      # packed where the weights live, uploaded once
    if w5.dim() == 4:
        w5 = w5[:, :, None]
    kt = w5.shape[2]
    # kt > 1 (SlowFast fast stem, (5,7,7)): `frames` holds kt//2 zero frames at both ends of each of the `clips` clips
    # and the conv runs once per clip, so a temporal tap is a plain offset along the clip's own frame axis.
    assert w5.shape[1] == 3 and w5.shape[3] == w5.shape[4] == k and (kt == 1 or clips > 0)
    if kt > 1:
        fpc = nf // clips
        nf = fpc - (kt - 1)  # frames produced per launch
    oh, ow = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    # outputs per GEMM row: the S3D stem computes 4 neighbouring output pixels from one 16-pixel (128-byte) window, see below
    # (the temporal taps of the SlowFast fast stem are offsets along the clip's own frame axis: orthogonal to the wide rows)
    wide_ok = allow_wide and (kt == 1 or os.environ.get("MSPI_STEM_WIDE_KT", "1") != "0") and y.c0 == 0 and y.cs == cout \
        and os.environ.get("MSPI_STEM_WIDE", "1") != "0"
    wide = 1
    if wide_ok and (k, stride, pad) == (7, 2, 3) and ow % 4 == 0 and cout * 4 <= 256:
        wide = 4      # S3D conv_s: 16-pixel window (128 B), 4 outputs, N = 256
    elif wide_ok and (k, stride, pad) == (4, 4, 0) and ow % 2 == 0 and cout * 2 <= 256:
        wide = 2      # ConvNeXt stem: 8-pixel window (64 B instead of 32 B rows), 2 outputs, N = 192
    if (k, stride, pad) == (7, 2, 3):
        run_px, x_lead = (16, 1) if wide == 4 else (8, 1)     # window starts one pixel before tap 0 (alignment)
        row0, col0 = PAD_T - pad, PAD_L - pad - x_lead
    elif (k, stride, pad) == (4, 4, 0):
        run_px, x_lead = 4 * wide, 0
        row0, col0 = PAD_T, PAD_L
    elif (k, stride, pad) == (3, 2, 1):   # X3D stem conv_xy (stem_helper.py:262-270): 4-px window from 2*ow-2
        run_px, x_lead = 4, 1
        row0, col0 = PAD_T - pad, PAD_L - pad - x_lead
    else:
        raise ValueError(f"stem_conv: unsupported geometry k={k} stride={stride} pad={pad}")
    run_el = run_px * 4
    assert (col0 * 4 * 2) % 16 == 0 and (stride * 4 * 2) % 16 == 0 and row0 >= 0 and col0 >= 0
    # weight matrix [cout16][k taps][run_px][4]
    gemm_n = cout * wide
    rows16 = -(-gemm_n // 16) * 16
    wm = torch.zeros((rows16, kt, k, run_px, 4), dtype=torch.float32, device=w5.device)
    for j in range(wide):
        # wide == 4 (S3D stem): GEMM column (j, co) is output pixel 4g + j of the row's group g; its 7 taps sit at pixels
        # 2j + kw + x_lead of the 16-pixel window, the rest of its K entries are zero.  The TMA engine issues one request
        # per (row, tap) whatever the row length, and those requests — not bytes, not MMA time — bound the 8-pixel form
        # (896 per 128 outputs); four outputs per row need a quarter of them at twice the (cheap) MMA work.
        wm[j * cout:(j + 1) * cout, :, :, stride * j + x_lead:stride * j + x_lead + k, :3] = w5.permute(0, 2, 3, 4, 1)
    packed = wm.reshape(rows16, kt * k * run_el).to(torch.bfloat16).contiguous().to(frames.device)
    # frame row of (oh, kh) = row0 + stride*oh + kh = row0 + stride*(oh + kh // stride) + kh % stride
    n_r, n_j = min(stride, k), (k - 1) // stride + 1
    assert row0 + stride * (oh - 1) + k - 1 < hp and col0 + wide * stride * (ow // wide - 1) + run_px - 1 < wp
    d = ConvDesc()
    d.a_dtype = MSPI_BF16
    d.k_row_bytes = run_el * 2
    a_dims = (run_el, n_r, ow // wide, oh + n_j - 1, nf + kt - 1)
    a_str = (1, wp * 4, wide * stride * 4, stride * wp * 4, hp * wp * 4)
    for j in range(5):
        d.a_dims[j] = a_dims[j]
        d.a_strides[j] = a_str[j]
    o_dims = (1, ow // wide, oh, nf)
    box = (1,) + choose_box((ow // wide, oh, nf, 1))[:3]
    ostr = (0, wide * y.cs, y.w * y.cs, y.h * y.w * y.cs)  # frames (n, t) are contiguous in y; a wide row = `wide` pixels
    d.box[0] = 0
    for j in range(4):
        d.box[j + 1] = box[j]
        d.o_dims[j] = o_dims[j]
        d.o_strides[j] = ostr[j]
        d.r_strides[j] = 0
    d.ntaps = kt * k
    assert d.ntaps <= _lib.MAX_TAPS
    for it in range(kt):
        for kh in range(k):
            i = it * k + kh
            d.tap_off[i][0], d.tap_off[i][1], d.tap_off[i][2], d.tap_off[i][3] = kh % stride, 0, kh // stride, it
    d.cin_pad = run_el
    d.cout, d.w_rows = gemm_n, rows16
    d.bn = choose_bn(gemm_n, 64 if y.dtype == torch.bfloat16 else 32)
    d.o_dtype = _DT[y.dtype]
    d.act = act
    if ln is not None:
        assert y.dtype == torch.bfloat16 and act == ACT_NONE and kt == 1 and gemm_n % 64 == 0 and gemm_n <= 192 \
            and wide in (1, 2, 4), "stem_conv: geometry does not qualify for the LayerNorm epilogue"
        d.bn = gemm_n
        lw = ln[0].detach().float().contiguous().to(frames.device)
        lb = ln[1].detach().float().contiguous().to(frames.device)
        ln_eps = float(ln[2])
        assert lw.numel() == cout and lb.numel() == cout
    sc = None if scale is None else scale.detach().float().repeat(wide).contiguous().to(frames.device)
    sh = None if shift is None else shift.detach().float().repeat(wide).contiguous().to(frames.device)
    launches = 1 if kt == 1 else clips
    x_ptrs = [_ptr(frames, ((b * (nf + kt - 1) * hp + row0) * wp + col0) * 4 * 2) for b in range(launches)]
    es_y = _ES[y.dtype]
    y_ptrs = [_ptr(y.buf, (b * nf * oh * ow * y.cs + y.c0) * es_y) for b in range(launches)]
    w_ptr = _ptr(packed)
    assert y.pixels == launches * nf * oh * ow and y.c == cout

    if ln is not None:
        def run(_keep=(frames, packed, sc, sh, y.buf, d, lw, lb)):
            for xp_, yp_ in zip(x_ptrs, y_ptrs):
                _lib.check(lib.mspi_conv_gemm_ln(C.byref(d), xp_, w_ptr, _ptr(sc), _ptr(sh), _ptr(lw), _ptr(lb), ln_eps, wide,
                                                 yp_, _stream()), f"conv_gemm_ln[{name}]")
    else:
        def run(_keep=(frames, packed, sc, sh, y.buf, d)):
            for xp_, yp_ in zip(x_ptrs, y_ptrs):
                _lib.check(lib.mspi_conv_gemm(C.byref(d), xp_, w_ptr, _ptr(sc), _ptr(sh), C.c_void_p(0), yp_, _stream()),
                           f"conv_gemm[{name}]")

    run.mode = "stem"
    run.desc = d
    run.packed, run.geom = packed, (k, kt, run_px, x_lead, rows16)   # the training plan re-packs the weights in place
    run.flops = 2.0 * launches * nf * oh * ow * cout * 3 * kt * k * k
    return run


# ------------------------------------------------------------------------------------------ few-channel (1,3,3) conv
def conv133_small_ok(w: torch.Tensor, stride, pad, x: "Act", dtype, out_dtype, residual) -> bool:
    """Whether a conv is the few-channel (1,3,3) layer mspi_conv133_small serves (SlowFast fast pathway, dim_inner 8 / 16)."""
    return (w.dim() == 5 and tuple(w.shape[2:]) == (1, 3, 3) and w.shape[0] == w.shape[1] and w.shape[0] in (8, 16)
            and tuple(stride) == (1, 1, 1) and tuple(pad) == (0, 1, 1) and residual is None and dtype == torch.bfloat16
            and out_dtype == torch.bfloat16 and x.dtype == torch.bfloat16 and x.c == w.shape[1] and x.c0 % 8 == 0
            and x.cs % 8 == 0 and os.environ.get("MSPI_SMALLC_CONV", "1") != "0")


def conv133_small(x: "Act", y: "Act", weight: torch.Tensor, scale, shift, act: int) -> Callable[[], None]:
    """y = act(conv(1,3,3)(x) * scale + shift) for Cin = Cout = 8 / 16 on CUDA cores (mspi_conv133_small); the BatchNorm scale
    is multiplied into the fp32 weights on the host."""
    lib = _lib.load()
    c = weight.shape[0]
    assert (y.n, y.t, y.h, y.w, y.c) == (x.n, x.t, x.h, x.w, c) and y.dtype == torch.bfloat16 and y.c0 % 8 == 0 and y.cs % 8 == 0
    w = weight.detach().float()[:, :, 0]                       # [co, ci, 3, 3]
    if scale is not None:
        w = w * scale.detach().float().view(-1, 1, 1, 1)
    wp = w.permute(2, 3, 1, 0).contiguous().view(-1).to(x.buf.device)     # [kh][kw][ci][co]
    sh = None if shift is None else shift.detach().float().contiguous().to(x.buf.device)
    xp, yp, xcs, ycs = x.ptr, y.ptr, x.cs, y.cs
    planes, h, wd = x.n * x.t, x.h, x.w

    def run(_keep=(x.buf, y.buf, wp, sh)):
        _lib.check(lib.mspi_conv133_small(xp, xcs, _ptr(wp), _ptr(sh), yp, ycs, planes, h, wd, c, act, _stream()), "conv133_small")

    run.flops = 2.0 * x.pixels * c * c * 9
    return run


# ------------------------------------------------------------------------------------------ fused ConvNeXt MLP
def mlp_fused(x: Act, y: Act, residual: Act, fc1_w, fc1_b, fc2_w, fc2_b, gamma, ln=None) -> Callable[[], None]:
    """y = residual + gamma * (fc2(gelu(fc1(x)))) in one kernel (mspi_mlp_fused); C = 96 / 192, bf16.
    ln = (weight, bias, eps): y = LayerNorm_c(that) instead — the next stage's downsample.0 norm fused into the store
    (mspi_mlp_fused_ln)."""
    lib = _lib.load()
    dev = x.buf.device
    c = x.c
    assert c in (96, 192) and x.c0 == 0 and x.cs == c and x.dtype == y.dtype == residual.dtype == torch.bfloat16
    assert y.c == c and residual.c == c and y.pixels == x.pixels == residual.pixels
    c_pad = -(-c // 64) * 64
    w1 = torch.zeros((4 * c, c_pad), dtype=torch.float32, device=fc1_w.device)
    w1[:, :c] = fc1_w.detach().float()
    w1 = w1.to(torch.bfloat16).contiguous().to(dev)
    w2 = fc2_w.detach().float().to(torch.bfloat16).contiguous().to(dev)      # [c, 4c], K contiguous
    b1 = fc1_b.detach().float().contiguous().to(dev)
    g = gamma.detach().float().contiguous().to(dev)
    sh = (gamma.detach().float() * fc2_b.detach().float()).contiguous().to(dev)
    m = x.pixels
    xp, yp, rp = x.ptr, y.ptr, residual.ptr
    rs, ys = residual.cs, y.cs

    if ln is not None:
        lw = ln[0].detach().float().contiguous().to(dev)
        lb = ln[1].detach().float().contiguous().to(dev)
        eps = float(ln[2])
        assert lw.numel() == c and lb.numel() == c

        def run(_keep=(x.buf, y.buf, residual.buf, w1, w2, b1, g, sh, lw, lb)):
            _lib.check(lib.mspi_mlp_fused_ln(xp, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(g), _ptr(sh), rp, yp, m, c, c_pad, rs, ys,
                                             _ptr(lw), _ptr(lb), eps, _stream()), "mlp_fused_ln")
    else:
        def run(_keep=(x.buf, y.buf, residual.buf, w1, w2, b1, g, sh)):
            _lib.check(lib.mspi_mlp_fused(xp, _ptr(w1), _ptr(b1), _ptr(w2), _ptr(g), _ptr(sh), rp, yp, m, c, c_pad, rs, ys,
                                          _stream()), "mlp_fused")

    run.flops = 2.0 * m * c * 4 * c * 2
    return run


# ------------------------------------------------------------------------------------------ X3D pieces
def dwconv3d_bn(x: Act, y: Act, weight: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
                stride_hw: int = 1, act: int = ACT_NONE, mean: Optional[torch.Tensor] = None) -> Callable[[], None]:
    """Depthwise Conv3d [C,1,kt,kh,kw] ("same" padding, stride (1,s,s)) + per-channel scale/shift + activation.
    The channel count of x / y may exceed the weight's (buffers padded to a multiple of 8): extra channels get zero
    weights and zero shift, so they stay zero."""
    lib = _lib.load()
    w = weight.detach().float()
    c, _, kt, kh, kw = w.shape
    cp = x.c
    assert cp >= c and cp % 8 == 0 and y.c == cp and x.dtype == y.dtype == torch.bfloat16
    dev = x.buf.device
    wt = torch.zeros((kt * kh * kw, cp), dtype=torch.float32, device=w.device)
    sc = torch.ones(c, device=w.device) if scale is None else scale.detach().float().to(w.device)
    wt[:, :c] = (w.reshape(c, -1) * sc[:, None]).t()
    wt = wt.to(dev)
    sh = torch.zeros(cp, dtype=torch.float32, device=w.device)
    if shift is not None:
        sh[:c] = shift.detach().float().to(w.device)
    sh = sh.to(dev)
    d = Dw3dDesc()
    d.n, d.t, d.h, d.w, d.c = x.n, x.t, x.h, x.w, cp
    d.in_cstride, d.out_cstride = x.cs, y.cs
    d.kt, d.kh, d.kw, d.sh, d.sw = kt, kh, kw, stride_hw, stride_hw
    d.oh, d.ow = y.h, y.w
    d.act = act
    assert (y.n, y.t) == (x.n, x.t) and x.c0 % 8 == 0 and y.c0 % 8 == 0
    xp, yp = x.ptr, y.ptr

    if mean is not None:      # SE squeeze accumulated by the depthwise kernel (mspi_dwconv3d_bn_mean): fp32 [n][cp]
        assert mean.dtype == torch.float32 and tuple(mean.shape) == (x.n, cp) and y.c0 == 0 and y.cs == cp
        work = torch.empty((x.n, 16, cp), dtype=torch.float32, device=dev)     # partial means (kSeSlots = 16)

        def run(_keep=(x.buf, y.buf, wt, sh, d, mean, work)):
            _lib.check(lib.mspi_dwconv3d_bn_mean(C.byref(d), xp, _ptr(wt), _ptr(sh), yp, _ptr(mean), _ptr(work), _stream()),
                       "dwconv3d_bn_mean")
    else:
        def run(_keep=(x.buf, y.buf, wt, sh, d)):
            _lib.check(lib.mspi_dwconv3d_bn(C.byref(d), xp, _ptr(wt), _ptr(sh), yp, _stream()), "dwconv3d_bn")

    return run


def se_block(x: Act, fc1_w, fc1_b, fc2_w, fc2_b, act: int = ACT_SWISH, mean: Optional[torch.Tensor] = None) -> List[Callable[[], None]]:
    """SE (+ the Swish that follows it in X3DTransform): x <- act(x * sigmoid(fc2(relu(fc1(mean_thw(x)))))), in place.
    resnet_helper.py:47-73,327-333.  Three launches: channel mean, the two FCs, scale+act; with `mean` (fp32 [n][c], already
    filled by the layer that produced x: dwconv3d_bn(..., mean=)) the first one is dropped."""
    lib = _lib.load()
    dev = x.buf.device
    cp, n = x.c, x.n
    assert x.c0 == 0 and x.cs == cp and x.dtype == torch.bfloat16
    w1 = fc1_w.detach().float().reshape(fc1_w.shape[0], -1)
    cfc, c = w1.shape
    w1p = torch.zeros((cfc, cp), dtype=torch.float32, device=w1.device)
    w1p[:, :c] = w1
    w1p = w1p.to(dev)
    w2p = torch.zeros((cp, cfc), dtype=torch.float32, device=w1.device)
    w2p[:c] = fc2_w.detach().float().reshape(c, cfc)
    w2p = w2p.to(dev)
    b1 = fc1_b.detach().float().contiguous().to(dev)
    b2 = torch.zeros(cp, dtype=torch.float32, device=w1.device)
    b2[:c] = fc2_b.detach().float()
    b2 = b2.to(dev)
    have_mean = mean is not None
    if not have_mean:
        mean = torch.empty((n, cp), dtype=torch.float32, device=dev)
    assert mean.dtype == torch.float32 and tuple(mean.shape) == (n, cp)
    gate = torch.empty((n, cp), dtype=torch.float32, device=dev)
    rows = x.t * x.h * x.w
    xp = x.ptr
    keep = (x.buf, w1p, w2p, b1, b2, mean, gate)

    def k_mean(_keep=keep):
        _lib.check(lib.mspi_channel_mean(xp, _ptr(mean), n, rows, cp, cp, _stream()), "channel_mean")

    def k_gate(_keep=keep):
        _lib.check(lib.mspi_se_gate(_ptr(mean), _ptr(w1p), _ptr(b1), _ptr(w2p), _ptr(b2), _ptr(gate), n, cp, cfc, _stream()),
                   "se_gate")

    def k_scale(_keep=keep):
        _lib.check(lib.mspi_scale_act(xp, _ptr(gate), xp, n, rows, cp, act, _stream()), "scale_act")

    return [k_gate, k_scale] if have_mean else [k_mean, k_gate, k_scale]


# ------------------------------------------------------------------------------------------ other ops
def patch_gather_ncdhw(src: torch.Tensor, kernel, stride, pad, k_pad: int, out: torch.Tensor) -> Callable[[], None]:
    """fp32 NCDHW input (the model's input contract) -> bf16 patch rows [M, k_pad]."""
    lib = _lib.load()
    n, c, t, h, w = src.shape
    pd = PatchDesc()
    pd.src_layout = 0
    pd.n, pd.c, pd.t, pd.h, pd.w = n, c, t, h, w
    pd.src_cstride = 0
    pd.kt, pd.kh, pd.kw = kernel
    pd.st, pd.sh, pd.sw = stride
    pd.pt, pd.ph, pd.pw = pad
    pd.ot = (t + 2 * pad[0] - kernel[0]) // stride[0] + 1
    pd.oh = (h + 2 * pad[1] - kernel[1]) // stride[1] + 1
    pd.ow = (w + 2 * pad[2] - kernel[2]) // stride[2] + 1
    pd.k_pad = k_pad
    assert out.shape == (n * pd.ot * pd.oh * pd.ow, k_pad) and out.dtype == torch.bfloat16
    sp, op = _ptr(src), _ptr(out)

    def run(_keep=(src, out, pd)):
        _lib.check(lib.mspi_patch_gather(C.byref(pd), sp, op, _stream()), "patch_gather")

    return run


def maxpool3d(x: Act, y: Act, kernel, stride, pad) -> Callable[[], None]:
    lib = _lib.load()
    d = PoolDesc()
    d.n, d.t, d.h, d.w, d.c = x.n, x.t, x.h, x.w, x.c
    d.in_cstride, d.out_cstride = x.cs, y.cs
    d.kt, d.kh, d.kw = kernel
    d.st, d.sh, d.sw = stride
    d.pt, d.ph, d.pw = pad
    d.ot, d.oh, d.ow = y.t, y.h, y.w
    assert y.c == x.c and y.n == x.n
    assert d.ot == (x.t + 2 * pad[0] - kernel[0]) // stride[0] + 1
    assert d.oh == (x.h + 2 * pad[1] - kernel[1]) // stride[1] + 1
    assert d.ow == (x.w + 2 * pad[2] - kernel[2]) // stride[2] + 1
    xp, yp = x.ptr, y.ptr

    def run(_keep=(x.buf, y.buf, d)):
        _lib.check(lib.mspi_maxpool3d(C.byref(d), xp, yp, _stream()), "maxpool3d")

    return run


def upsample(x: Act, y: Act, k: int, accumulate: bool = False, act: int = ACT_NONE) -> Callable[[], None]:
    lib = _lib.load()
    d = UpDesc()
    d.nt, d.h, d.w, d.c, d.k = x.n * x.t, x.h, x.w, x.c, k
    d.in_cstride, d.out_cstride = x.cs, y.cs
    d.in_dtype, d.out_dtype = _DT[x.dtype], _DT[y.dtype]
    d.accumulate = 1 if accumulate else 0
    d.act = act
    assert (y.n, y.t, y.h, y.w, y.c) == (x.n, x.t, x.h * k, x.w * k, x.c)
    xp, yp = x.ptr, y.ptr

    def run(_keep=(x.buf, y.buf, d)):
        _lib.check(lib.mspi_upsample_bilinear(C.byref(d), xp, yp, _stream()), "upsample")

    return run


def dwconv_ln(x: Act, y: Act, weight: torch.Tensor, bias: torch.Tensor, ln_w=None, ln_b=None,
              eps: float = 1e-5) -> Callable[[], None]:
    """weight: [C, 1, kt, kh, kw] or [C, 1, kh, kw] depthwise conv weight (fp32)."""
    lib = _lib.load()
    w = weight.detach().float()
    if w.dim() == 4:
        w = w[:, :, None]
    c, _, kt, kh, kw = w.shape
    assert x.c == c and x.c0 == 0 and x.cs == c and y.c0 == 0 and y.cs == c, "dwconv works on whole buffers"
    wt = w.reshape(c, kt * kh * kw).t().contiguous().to(x.buf.device)
    b = bias.detach().float().contiguous().to(x.buf.device)
    lw = None if ln_w is None else ln_w.detach().float().contiguous().to(x.buf.device)
    lb = None if ln_b is None else ln_b.detach().float().contiguous().to(x.buf.device)
    d = DwDesc()
    d.n, d.t, d.h, d.w, d.c = x.n, x.t, x.h, x.w, c
    d.kt, d.kh, d.kw = kt, kh, kw
    d.ln_eps = eps
    d.out_dtype = _DT[y.dtype]
    d.in_dtype = _DT[x.dtype]
    xp, yp = x.ptr, y.ptr

    def run(_keep=(x.buf, y.buf, wt, b, lw, lb, d)):
        _lib.check(lib.mspi_dwconv_ln(C.byref(d), xp, _ptr(wt), _ptr(b), _ptr(lw), _ptr(lb), yp, _stream()), "dwconv_ln")

    return run


def layernorm(x: torch.Tensor, y: torch.Tensor, rows: int, c: int, w: torch.Tensor, b: torch.Tensor, eps: float,
              in_rstride: Optional[int] = None, out_rstride: Optional[int] = None, relu: bool = False,
              pos: Optional[torch.Tensor] = None, rows_per_group: int = 0, out_gstride: int = 0,
              x_off: int = 0, y_off: int = 0) -> Callable[[], None]:
    """LayerNorm over the last dim of `rows` rows; x_off / y_off are element offsets into x / y."""
    lib = _lib.load()
    d = LnDesc()
    d.rows, d.c = rows, c
    d.in_rstride = c if in_rstride is None else in_rstride
    d.out_rstride = c if out_rstride is None else out_rstride
    d.in_dtype, d.out_dtype = _DT[x.dtype], _DT[y.dtype]
    d.eps = eps
    d.relu = 1 if relu else 0
    d.pos_rows = 0 if pos is None else pos.shape[0]
    d.rows_per_group = rows_per_group if rows_per_group else rows
    d.out_gstride = out_gstride
    wf = w.detach().float().contiguous().to(x.device)
    bf = b.detach().float().contiguous().to(x.device)
    xp, yp = _ptr(x, x_off * _ES[x.dtype]), _ptr(y, y_off * _ES[y.dtype])

    def run(_keep=(x, y, wf, bf, pos, d)):
        _lib.check(lib.mspi_layernorm(C.byref(d), xp, _ptr(wf), _ptr(bf), _ptr(pos), yp, _stream()), "layernorm")

    return run


def attention(qkv: torch.Tensor, out: torch.Tensor, b: int, n: int, heads: int, hd: int) -> Callable[[], None]:
    lib = _lib.load()
    scale = float(hd) ** -0.5

    def run(_keep=(qkv, out)):
        _lib.check(lib.mspi_attention(_ptr(qkv), _ptr(out), _DT[qkv.dtype], b, n, heads, hd, scale, _stream()), "attention")

    return run


def attention_gemm(qkv: torch.Tensor, out: torch.Tensor, b: int, n: int, heads: int, hd: int) -> List[Tuple[str, Callable[[], None]]]:
    """Multi-head self-attention (model_utils.py:97-109) as tensor-core work: scores = scale * Q K^T and out = P V are
    batched kind::tf32 GEMMs (one weight matrix per (head, sample), MspiConvDesc.w_batch_dims); softmax and the K-major
    copy of V are small fp32 kernels.  qkv fp32 [B*N, 3*heads*hd] (layout [3][heads][hd] per token), out fp32 [B*N, heads*hd].
    Returns the four launches as (name, closure)."""
    lib = _lib.load()
    assert qkv.dtype == torch.float32 and out.dtype == torch.float32 and hd % 32 == 0
    dev = qkv.device
    c = heads * hd
    n_pad = -(-n // 4) * 4                      # fp32 rows must start 16-byte aligned
    scores = torch.empty((b, heads, n, n_pad), dtype=torch.float32, device=dev)
    vt = torch.zeros((b, heads, hd, n_pad), dtype=torch.float32, device=dev)
    scale = torch.full((n,), float(hd) ** -0.5, dtype=torch.float32, device=dev)
    row = 3 * c                                  # elements between consecutive tokens of qkv

    def desc(k, m_stride, h_stride, b_stride, w_rows, w_strides, cout, o_strides):
        d = ConvDesc()
        d.a_dtype = MSPI_F32
        for j, (dim, st) in enumerate(zip((k, n, 1, heads, b), (1, m_stride, m_stride * n, h_stride, b_stride))):
            d.a_dims[j], d.a_strides[j] = dim, st
        d.box[0], d.box[1], d.box[2], d.box[3], d.box[4] = 0, min(n, 128), 1, 1, 1
        d.ntaps = 1
        d.cin_pad = -(-k // 32) * 32
        d.cout, d.w_rows = cout, w_rows
        d.bn = choose_bn(cout, 32)
        for j, (dim, st) in enumerate(zip((n, 1, heads, b), o_strides)):
            d.o_dims[j], d.o_strides[j], d.r_strides[j] = dim, st, 0
        d.o_dtype = MSPI_F32
        d.w_batch_dims[0], d.w_batch_dims[1] = heads, b
        for j in range(3):
            d.w_strides[j] = w_strides[j]
        return d

    # scores[b][h][q][k] = scale * sum_d Q[b,q,h,d] K[b,k,h,d]
    d1 = desc(hd, row, hd, n * row, n, (row, hd, n * row), n, (n_pad, 0, n * n_pad, heads * n * n_pad))
    # out[b,q,h,:] = sum_k P[b][h][q][k] Vt[b][h][:, k]
    d2 = desc(n, n_pad, n * n_pad, heads * n * n_pad, hd, (n_pad, hd * n_pad, heads * hd * n_pad), hd, (c, 0, hd, n * c))
    q_ptr, k_ptr = _ptr(qkv), _ptr(qkv, c * 4)
    keep = (qkv, out, scores, vt, scale, d1, d2)
    null = C.c_void_p(0)

    def k_scores(_keep=keep):
        _lib.check(lib.mspi_conv_gemm(C.byref(d1), q_ptr, k_ptr, _ptr(scale), null, null, _ptr(scores), _stream()), "attn.scores")

    def k_softmax(_keep=keep):
        _lib.check(lib.mspi_softmax_rows(_ptr(scores), b * heads * n, n, n_pad, _stream()), "attn.softmax")

    def k_vt(_keep=keep):
        _lib.check(lib.mspi_transpose_v(_ptr(qkv), _ptr(vt), b, n, heads, hd, n_pad, _stream()), "attn.transpose_v")

    def k_out(_keep=keep):
        _lib.check(lib.mspi_conv_gemm(C.byref(d2), _ptr(scores), _ptr(vt), null, null, null, _ptr(out), _stream()), "attn.out")

    k_scores.desc, k_scores.flops = d1, 2.0 * b * heads * n * n * hd
    k_out.desc, k_out.flops = d2, 2.0 * b * heads * n * n * hd
    return [("scores", k_scores), ("softmax", k_softmax), ("transpose_v", k_vt), ("out", k_out)]


def conv_c1(x: Act, weight: torch.Tensor, bias: Optional[torch.Tensor], y: Act) -> Callable[[], None]:
    """Conv3d(32, 1, (1,3,3), padding (0,1,1)) + bias -> fp32 [N,T,H,W,1] (SA.conv_mask.2, model_utils.py:163; readout.12,
    model_utils.py:503) on the direct kernel (mspi_conv_c1_fwd).  weight / bias are used in place (live fp32 tensors)."""
    lib = _lib.load()
    assert x.c == 32 and tuple(weight.shape[-3:]) == (1, 3, 3) and weight.shape[0] == 1 and weight.shape[1] == 32
    assert y.dtype == torch.float32 and y.c == 1 and y.cs == 1 and y.pixels == x.pixels
    w = weight.detach()
    w = w if (w.dtype == torch.float32 and w.is_contiguous() and w.device == x.buf.device) else w.float().contiguous().to(x.buf.device)
    b = None if bias is None else (bias.detach() if (bias.dtype == torch.float32 and bias.device == x.buf.device)
                                   else bias.detach().float().to(x.buf.device))
    planes, h, wd = x.n * x.t, x.h, x.w
    xp, yp, dt, xcs = x.ptr, y.ptr, _DT[x.dtype], x.cs

    def run(_keep=(x.buf, y.buf, w, b)):
        _lib.check(lib.mspi_conv_c1_fwd(xp, dt, xcs, _ptr(w), _ptr(b), yp, planes, h, wd, 32, _stream()), "conv_c1_fwd")

    return run


def sa_gate(x: Act, mask_logits: torch.Tensor, y: Act) -> Callable[[], None]:
    lib = _lib.load()
    assert mask_logits.dtype == torch.float32 and x.c == y.c and x.dtype == y.dtype
    pixels, c = x.pixels, x.c
    dt = _DT[x.dtype]
    assert mask_logits.numel() == pixels and y.pixels == pixels
    xp, yp, xcs, ycs = x.ptr, y.ptr, x.cs, y.cs

    def run(_keep=(x.buf, y.buf, mask_logits)):
        _lib.check(lib.mspi_sa_gate(xp, xcs, _ptr(mask_logits), yp, ycs, pixels, c, dt, _stream()), "sa_gate")

    return run


def sa_gate_fused(x: Act, mask_logits: torch.Tensor, y: Act, sources) -> Callable[[], None]:
    """y = x * sigmoid(mask) + x + sum up_k(src) for (src Act, k) in sources (fp32): SA gate + top-down fusion in one pass
    (model_utils.py:167-170,566-568)."""
    lib = _lib.load()
    assert x.dtype == y.dtype == torch.float32 and len(sources) <= 3
    assert mask_logits is None or (mask_logits.dtype == torch.float32 and mask_logits.numel() == x.pixels)
    assert (y.n, y.t, y.h, y.w, y.c) == (x.n, x.t, x.h, x.w, x.c)
    n = len(sources)
    ptrs = (C.c_void_p * max(n, 1))(*[a.ptr for a, _k in sources])
    css = (C.c_int64 * max(n, 1))(*[a.cs for a, _k in sources])
    ks = (C.c_int32 * max(n, 1))(*[k for _a, k in sources])
    for a, k in sources:
        assert a.dtype == torch.float32 and (a.n, a.t, a.h * k, a.w * k, a.c) == (x.n, x.t, x.h, x.w, x.c)
    xp, yp, xcs, ycs = x.ptr, y.ptr, x.cs, y.cs
    nt, h, w, c = x.n * x.t, x.h, x.w, x.c

    def run(_keep=(x.buf, y.buf, mask_logits, [a.buf for a, _k in sources], ptrs, css, ks)):
        _lib.check(lib.mspi_sa_gate_fused(xp, xcs, _ptr(mask_logits), yp, ycs, nt, h, w, c, n, ptrs, css, ks, _stream()),
                   "sa_gate_fused")

    return run


def token_mean(x: torch.Tensor, y: torch.Tensor, b: int, rows: int, r0: int, r1: int, c: int) -> Callable[[], None]:
    lib = _lib.load()

    def run(_keep=(x, y)):
        _lib.check(lib.mspi_token_mean(_ptr(x), _DT[x.dtype], _ptr(y), b, rows, r0, r1, c, _stream()), "token_mean")

    return run


def cast_rows(src: torch.Tensor, dst: torch.Tensor, groups: int, rows: int, c: int, src_rstride: int, src_gstride: int,
              dst_rstride: int, dst_gstride: int, src_off: int = 0, dst_off: int = 0) -> Callable[[], None]:
    lib = _lib.load()
    sp, dp = _ptr(src, src_off * _ES[src.dtype]), _ptr(dst, dst_off * _ES[dst.dtype])

    def run(_keep=(src, dst)):
        _lib.check(lib.mspi_cast_rows(sp, _DT[src.dtype], src_rstride, src_gstride, dp, _DT[dst.dtype], dst_rstride,
                                      dst_gstride, groups, rows, c, _stream()), "cast_rows")

    return run


def simsiam_loss(pv, za, pa, zv, out, b: int, c: int) -> Callable[[], None]:
    lib = _lib.load()

    def run(_keep=(pv, za, pa, zv, out)):
        _lib.check(lib.mspi_simsiam_loss(_ptr(pv), _ptr(za), _ptr(pa), _ptr(zv), _ptr(out), b, c, _stream()), "simsiam")

    return run


def logsoftmax2d(x: torch.Tensor, y: torch.Tensor, b: int, pixels: int) -> Callable[[], None]:
    lib = _lib.load()

    def run(_keep=(x, y)):
        _lib.check(lib.mspi_logsoftmax2d(_ptr(x), _ptr(y), b, pixels, _stream()), "logsoftmax2d")

    return run
