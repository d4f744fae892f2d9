"""TEST INFRASTRUCTURE (see oracle/mspi_oracle.py): the oracle's training step re-evaluated with the operand rounding of
the CUDA path, for the training parity test only.

Why this exists.  Train-mode BatchNorm at random initialisation makes the S3D forward *chaotic*: a relative perturbation
of the activations grows by ~1.5x per Inception block (measured: 6e-3 after base1 -> 4e-1 after base4.1 at 64x64, the
well-known mean-field behaviour of batch-normalised ReLU networks at init).  The CUDA path multiplies in tf32 (bf16 in the
frozen encoders and the Cin=3 stem), so against the plain fp32 oracle (mspi_oracle.train_grads, pinned to the live
reference) its gradients can only agree to the amplified rounding noise — exactly as the reference itself run on a GPU
with PyTorch's default TF32 convolutions would.  To still check every trainable layer's forward and backward wiring to a
tight tolerance, this module evaluates THE SAME restated algorithm (mspi_oracle._forward + autograd, nothing else) with

  * every trainable Conv3d / Linear computed on tf32-rounded operands, forward, data gradient and weight gradient
    (custom autograd functions: products of rounded operands are exact in fp32, accumulation stays fp32);
  * the Cin=3 S3D stem on bf16-rounded clip and weights (forward and weight gradient);
  * the 32->1 mask / readout convolutions left in fp32 (direct CUDA-core kernels on the CUDA path);
  * the outputs of the frozen encoders (image encoder features, audio features) injected from the CUDA run, since those
    run the bf16 inference pipeline and need no gradient.

Only tests/ may import this module.
"""
from __future__ import annotations

import contextlib
from typing import Optional

import torch
import torch.nn.functional as F_real

from . import mspi_oracle as orc


def round_tf32(x: torch.Tensor, trunc: bool = False) -> torch.Tensor:
    """fp32 -> tf32 (10 explicit mantissa bits) the way the CUDA path's operands reach the tensor core: round to nearest,
    ties to even (measured on B200 with tools/tf32_round.py: 1+2^-11 -> 1, 1+3*2^-11 -> 1+2^-9, 1+2^-11+2^-13 -> 1+2^-10;
    the TMA TFLOAT32 load performs it).  trunc=True gives plain truncation, for comparison."""
    i = x.contiguous().view(torch.int32)
    if not trunc:
        i = i + 0xFFF + ((i >> 13) & 1)     # sign-magnitude: adding to the raw bits grows the magnitude for either sign
    return (i & -8192).view(torch.float32)


def round_bf16(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32)


def _tup(v, n=3):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _ConvR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, stride, padding, rnd_f, rnd_b):
        rx, rw = rnd_f(x), rnd_f(w)
        ctx.save_for_backward(rx, rw)
        ctx.cfg = (stride, padding, rnd_b, bias is not None, x.requires_grad)
        return F_real.conv3d(rx, rw, bias, stride, padding)

    @staticmethod
    def backward(ctx, dy):
        rx, rw = ctx.saved_tensors
        stride, padding, rnd_b, has_bias, need_dx = ctx.cfg
        if rnd_b is None:   # backward in exact fp32 on the unrounded-gradient (CUDA-core kernels)
            rdy = dy
        else:
            rdy = rnd_b(dy)
        dx = torch.nn.grad.conv3d_input(rx.shape, rw, rdy, stride, padding) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv3d_weight(rx, rw.shape, rdy, stride, padding)
        db = dy.sum((0, 2, 3, 4)) if has_bias else None
        return dx, dw, db, None, None, None, None


class _LinearR(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, rnd):
        rx, rw = rnd(x), rnd(w)
        ctx.save_for_backward(rx, rw)
        ctx.rnd, ctx.has_bias = rnd, bias is not None
        y = rx @ rw.t()
        return y + bias if bias is not None else y

    @staticmethod
    def backward(ctx, dy):
        rx, rw = ctx.saved_tensors
        rdy = ctx.rnd(dy)
        dx = rdy @ rw if ctx.needs_input_grad[0] else None
        dw = rdy.reshape(-1, rdy.shape[-1]).t() @ rx.reshape(-1, rx.shape[-1])
        db = dy.reshape(-1, dy.shape[-1]).sum(0) if ctx.has_bias else None
        return dx, dw, db, None


class _MatmulR(torch.autograd.Function):
    """Attention products: tf32 operands in the forward (batched tensor-core GEMMs), exact fp32 backward (strided SGEMM)."""

    @staticmethod
    def forward(ctx, a, b, rnd):
        ctx.save_for_backward(a, b)
        return rnd(a) @ rnd(b)

    @staticmethod
    def backward(ctx, dy):
        a, b = ctx.saved_tensors
        return dy @ b.transpose(-2, -1), a.transpose(-2, -1) @ dy, None


class _FProxy:
    """torch.nn.functional with conv3d / linear replaced by the rounding versions."""

    def __init__(self, trunc: bool):
        self._tf32 = lambda t: round_tf32(t, trunc)

    def __getattr__(self, name):
        return getattr(F_real, name)

    def conv3d(self, x, w, bias=None, stride=1, padding=0, dilation=1, groups=1):
        if groups != 1:
            return F_real.conv3d(x, w, bias, stride, padding, dilation, groups)   # depthwise: fp32 FMA on the CUDA path
        stride, padding = _tup(stride), _tup(padding)
        if w.shape[1] == 3:       # S3D stem on the clip: bf16 operands, bf16 weight gradient
            return _ConvR.apply(x, w, bias, stride, padding, round_bf16, round_bf16)
        if w.shape[0] == 1:       # 32 -> 1 convs: fp32 CUDA-core kernels, forward and backward
            return F_real.conv3d(x, w, bias, stride, padding)
        return _ConvR.apply(x, w, bias, stride, padding, self._tf32, self._tf32)

    def linear(self, x, w, bias=None):
        return _LinearR.apply(x, w, bias, self._tf32)


@contextlib.contextmanager
def product_numerics(trunc: bool = False, image_feats=None, audio_feats=None):
    """Inside: mspi_oracle's conv3d / linear round their operands like the CUDA path, and (optionally) the frozen
    encoders return the given features instead of being evaluated."""
    saved = (orc.F, orc.image_encoder, orc.resnet18_audio, orc.vit_block)
    proxy = _FProxy(trunc)
    orc.F = proxy
    if image_feats is not None:
        orc.image_encoder = lambda sd, p, frames: image_feats
    if audio_feats is not None:
        orc.resnet18_audio = lambda sd, p, x: audio_feats
    tf32 = proxy._tf32

    def vit_block_r(sd, p, x, heads=4):
        """mspi_oracle.vit_block with the two attention products on tf32-rounded operands."""
        b, n, c = x.shape
        h = F_real.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        qkv = proxy.linear(h, sd[p + "attn.qkv.weight"]).reshape(b, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = (_MatmulR.apply(q, k.transpose(-2, -1), tf32) * (c // heads) ** -0.5).softmax(-1)
        h = _MatmulR.apply(a, v, tf32).transpose(1, 2).reshape(b, n, c)
        x = x + proxy.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
        h = F_real.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        h = proxy.linear(F_real.gelu(proxy.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                         sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
        return x + h

    orc.vit_block = vit_block_r
    try:
        yield
    finally:
        orc.F, orc.image_encoder, orc.resnet18_audio, orc.vit_block = saved


def train_grads_product_numerics(sd, clips, audios, gt, image_feats, audio_feats, trunc: bool = False, gamma: float = 1.0):
    """mspi_oracle.train_grads under product_numerics()."""
    with product_numerics(trunc, image_feats, audio_feats):
        return orc.train_grads(sd, clips, audios, gt, "s3d", gamma)
