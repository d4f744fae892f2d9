"""TEST INFRASTRUCTURE — CPU fp32 restatement of the reference's clip forward (oraclefina/MSPI).

This file is the parity ORACLE.  It is not product code: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import it, and only as the checker or the
CPU baseline.  The product (mspi_b200/) never imports it and has no CPU fallback.

It restates, function by function, what the reference's nn.Modules compute, as stateless
functions over a flat ``state_dict`` that uses the reference's own key names (so the very same
dict loads into the live reference through oracle/ref_shim.py).  Every function cites the
reference file:line it follows (paths relative to the reference root).

Pinned how: the reference ships NO tests, golden vectors or fixtures for this path (SURVEY.md §4,
§8c), so the oracle is pinned against the live reference itself: oracle/gen_golden.py imports the
unmodified reference in the build container, runs it on seeded inputs/weights produced by
``make_state_dict`` / ``make_inputs`` below and commits its outputs under tests/golden/;
tests/test_oracle_cpu.py checks this restatement against those vectors.  The image encoder's
arithmetic (timm==0.6.12 convnext_tiny, README.md:34) is not in the reference tree; its
restatement here is additionally pinned against torchvision's independent convnext_tiny.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# Inception-style block channel plans, (b0, b1a, b1, b2a, b2, b3): backbones/s3d.py:118-376
S3D_MIXED = {
    "base2.0": (192, (64, 96, 128, 16, 32, 32)),
    "base2.1": (256, (128, 128, 192, 32, 96, 64)),
    "base3.0": (480, (192, 96, 208, 16, 48, 64)),
    "base3.1": (512, (160, 112, 224, 24, 64, 64)),
    "base3.2": (512, (128, 128, 256, 24, 64, 64)),
    "base3.3": (512, (112, 144, 288, 32, 64, 64)),
    "base3.4": (528, (256, 160, 320, 32, 128, 128)),
    "base4.0": (832, (256, 160, 320, 32, 128, 128)),
    "base4.1": (832, (384, 192, 384, 48, 128, 128)),
}
ADAPTER_MIXED = (416, (192, 96, 208, 16, 48, 64))  # model/model_utils.py:173-191
S3D_EMBEDS = (192, 480, 832, 1024)                 # config.py:65
CONVNEXT_DIMS, CONVNEXT_DEPTHS = (96, 192, 384, 768), (3, 3, 9, 3)
DE = 192                                           # de_embed_dim, model_utils.py:389


# ============================================================================ building blocks
# Train mode (engine_train.py:19-20: model.train(); model.frozen_encoder()): every BatchNorm outside `audnet.` /
# `image_encoder.` normalises with the statistics of the current batch (biased variance) and moves its running statistics
# by `momentum` (unbiased variance), torch.nn.BatchNorm semantics.  _TRAIN["stats"] collects the updated buffers.
_TRAIN = {"on": False, "stats": None}
FROZEN_PREFIXES = ("audnet.", "image_encoder.")  # train.py:151-155, model_utils.py:516-518


def _bn(sd: SD, p: str, x, eps):
    shape = [1, -1] + [1] * (x.dim() - 2)
    if _TRAIN["on"] and not p.startswith(FROZEN_PREFIXES):
        # momentum: the S3D-style blocks (eps 1e-3) use 0.001 (backbones/s3d.py:45,99,103), nn.BatchNorm3d defaults
        # (eps 1e-5, readout: model_utils.py:493,496) use 0.1
        mom = 0.001 if eps == 1e-3 else 0.1
        rm, rv = sd[p + ".running_mean"].clone(), sd[p + ".running_var"].clone()
        # y = (x - mean_batch) / sqrt(var_batch_biased + eps) * w + b; rm/rv <- (1-mom)*old + mom*(mean, unbiased var)
        y = F.batch_norm(x, rm, rv, sd[p + ".weight"], sd[p + ".bias"], True, mom, eps)
        if _TRAIN["stats"] is not None:
            _TRAIN["stats"][p + ".running_mean"], _TRAIN["stats"][p + ".running_var"] = rm, rv
            _TRAIN["stats"][p + ".num_batches_tracked"] = sd[p + ".num_batches_tracked"] + 1
        return y
    scale = sd[p + ".weight"] / torch.sqrt(sd[p + ".running_var"] + eps)
    return (x - sd[p + ".running_mean"].view(shape)) * scale.view(shape) + sd[p + ".bias"].view(shape)


def basic_conv3d(sd: SD, p: str, x, stride=1, padding=0):
    """BasicConv3d: conv(bias=False) -> BN(eps 1e-3) -> ReLU.  backbones/s3d.py:41-52"""
    x = F.conv3d(x, sd[p + ".conv.weight"], None, stride, padding)
    return F.relu(_bn(sd, p + ".bn", x, 1e-3))


def sep_conv3d(sd: SD, p: str, x, k, stride, padding):
    """SepConv3d: (1,k,k) conv+BN+ReLU then (k,1,1) conv+BN+ReLU.  backbones/s3d.py:95-116"""
    x = F.conv3d(x, sd[p + ".conv_s.weight"], None, (1, stride, stride), (0, padding, padding))
    x = F.relu(_bn(sd, p + ".bn_s", x, 1e-3))
    x = F.conv3d(x, sd[p + ".conv_t.weight"], None, (stride, 1, 1), (padding, 0, 0))
    return F.relu(_bn(sd, p + ".bn_t", x, 1e-3))


def mixed_block(sd: SD, p: str, x):
    """Four-branch Inception block, concat on channels.  backbones/s3d.py:118-145 (all Mixed_*),
    model/model_utils.py:173-199 (Adapter's Inception)."""
    b0 = basic_conv3d(sd, p + ".branch0.0", x)
    b1 = sep_conv3d(sd, p + ".branch1.1", basic_conv3d(sd, p + ".branch1.0", x), 3, 1, 1)
    b2 = sep_conv3d(sd, p + ".branch2.1", basic_conv3d(sd, p + ".branch2.0", x), 3, 1, 1)
    b3 = basic_conv3d(sd, p + ".branch3.1", F.max_pool3d(x, 3, 1, 1))
    return torch.cat((b0, b1, b2, b3), 1)


def s3d_features(sd: SD, p: str, x, pool: int = 1) -> List[torch.Tensor]:
    """S3D_features_only.forward.  backbones/s3d.py:379-418"""
    x = sep_conv3d(sd, p + "base1.0", x, 7, 2, 3)
    x = F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    x = basic_conv3d(sd, p + "base1.2", x)
    v1 = sep_conv3d(sd, p + "base1.3", x, 3, 1, 1)
    x = F.max_pool3d(v1, (1, 3, 3), (1, 2, 2), (0, 1, 1))
    x = mixed_block(sd, p + "base2.0", x)
    v2 = mixed_block(sd, p + "base2.1", x)
    x = F.max_pool3d(v2, 3, 2, 1)
    for i in range(5):
        x = mixed_block(sd, p + f"base3.{i}", x)
    v3 = x
    x = F.max_pool3d(v3, (pool, 2, 2), (pool, 2, 2))
    x = mixed_block(sd, p + "base4.0", x)
    v4 = mixed_block(sd, p + "base4.1", x)
    return [v1, v2, v3, v4]


# ------------------------------------------------------------------ PySlowFast-style ResNet pieces
# (vendored SlowFast/resnet_helper.py, SlowFast/stem_helper.py; all BatchNorm eps 1e-5)
X3D_DEPTHS, X3D_OUT, X3D_INNER = (5, 10, 25, 15), (24, 48, 96, 192), (54, 108, 216, 432)  # X3D_L.yaml + X3D.py:139-163


def _se_width(dim_in: int, ratio: float = 0.0625, divisor: int = 8) -> int:
    """SE._round_width, resnet_helper.py:26-45"""
    w = dim_in * ratio
    out = max(divisor, int(w + divisor / 2) // divisor * divisor)
    if out < 0.9 * w:
        out += divisor
    return int(out)


def _se(sd: SD, p: str, x):
    """SE: avg-pool -> fc1 -> ReLU -> fc2 -> sigmoid -> scale.  resnet_helper.py:47-73"""
    s = x.mean((2, 3, 4), keepdim=True)
    s = F.relu(F.conv3d(s, sd[p + ".fc1.weight"], sd[p + ".fc1.bias"]))
    s = torch.sigmoid(F.conv3d(s, sd[p + ".fc2.weight"], sd[p + ".fc2.bias"]))
    return x * s


def _res_block(sd: SD, p: str, x, stride: int, transform):
    """ResBlock.forward: relu(shortcut + branch2(x)); projection shortcut when present.  resnet_helper.py:580-590"""
    f = transform(sd, p + ".branch2", x, stride)
    if (p + ".branch1.weight") in sd:
        x = _bn(sd, p + ".branch1_bn", F.conv3d(x, sd[p + ".branch1.weight"], None, (1, stride, stride)), 1e-5)
    return F.relu(x + f)


def _x3d_transform(use_se: bool):
    def tr(sd: SD, p: str, x, stride: int):
        """X3DTransform: 1x1x1+BN+ReLU -> depthwise 3x3x3 (stride on H,W)+BN -> [SE] -> Swish -> 1x1x1+BN.
        resnet_helper.py:213-351 (children() order)"""
        c = sd[p + ".b.weight"].shape[0]
        x = F.relu(_bn(sd, p + ".a_bn", F.conv3d(x, sd[p + ".a.weight"]), 1e-5))
        x = _bn(sd, p + ".b_bn", F.conv3d(x, sd[p + ".b.weight"], None, (1, stride, stride), (1, 1, 1), 1, c), 1e-5)
        if use_se:
            x = _se(sd, p + ".se", x)
        x = x * torch.sigmoid(x)
        return _bn(sd, p + ".c_bn", F.conv3d(x, sd[p + ".c.weight"]), 1e-5)
    return tr


def x3d_features(sd: SD, p: str, x) -> List[torch.Tensor]:
    """X3D(features_only).forward for X3D-L: stem conv_xy (1,3,3)/s(1,2,2) -> depthwise (5,1,1) -> BN -> ReLU; four
    stages of 5/10/25/15 ResBlocks (first of each stage has spatial stride 2), SE on every other block.
    backbones/X3D.py:111-250, stem_helper.py:207-290"""
    q = p + "s1.pathway0_stem"
    x = F.conv3d(x, sd[q + ".conv_xy.weight"], None, (1, 2, 2), (0, 1, 1))
    x = F.conv3d(x, sd[q + ".conv.weight"], None, 1, (2, 0, 0), 1, x.shape[1])
    x = F.relu(_bn(sd, q + ".bn", x, 1e-5))
    feats = []
    for si, depth in enumerate(X3D_DEPTHS):
        for i in range(depth):
            x = _res_block(sd, f"{p}s{si + 2}.pathway0_res{i}", x, 2 if i == 0 else 1, _x3d_transform((i + 1) % 2 == 1))
        feats.append(x)
    return feats


SF_DEPTHS = (3, 4, 6, 3)                       # ResNet-50, SLOWFAST_4x16_R50.yaml
SF_TK_SLOW, SF_TK_FAST = (1, 1, 3, 3), (3, 3, 3, 3)   # _TEMPORAL_KERNEL_BASIS["slowfast"], sf.py:31-100
SF_SLOW_FRAMES = (0, 4, 12, -1)                # model_utils.py:523


def _bottleneck(tk: int):
    def tr(sd: SD, p: str, x, stride: int):
        """BottleneckTransform: (tk,1,1)+BN+ReLU -> (1,3,3)/stride+BN+ReLU -> 1x1x1+BN.  resnet_helper.py:354-487"""
        x = F.relu(_bn(sd, p + ".a_bn", F.conv3d(x, sd[p + ".a.weight"], None, 1, (tk // 2, 0, 0)), 1e-5))
        x = F.relu(_bn(sd, p + ".b_bn", F.conv3d(x, sd[p + ".b.weight"], None, (1, stride, stride), (0, 1, 1)), 1e-5))
        return _bn(sd, p + ".c_bn", F.conv3d(x, sd[p + ".c.weight"]), 1e-5)
    return tr


def slowfast_features(sd: SD, p: str, clips) -> List[torch.Tensor]:
    """SlowFast 4x16 R50 features: slow pathway = frames [0,4,12,-1], fast = all 16.  Stems (k,7,7)/s(1,2,2)+BN+ReLU+
    MaxPool(1,3,3)/s(1,2,2); FuseFastToSlow = (5,1,1)/s(4,1,1) conv + BN + ReLU concatenated to the slow pathway
    after the stem and after res2..res4.  backbones/sf.py:101-389, model_utils.py:521-524"""
    xs = torch.stack([clips[:, :, i] for i in SF_SLOW_FRAMES], 2)
    xf = clips

    def stem(q, x, kt):
        x = F.relu(_bn(sd, q + ".bn", F.conv3d(x, sd[q + ".conv.weight"], None, (1, 2, 2), (kt // 2, 3, 3)), 1e-5))
        return F.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))

    def fuse(q, xs, xf):
        f = F.relu(_bn(sd, q + ".bn", F.conv3d(xf, sd[q + ".conv_f2s.weight"], None, (4, 1, 1), (2, 0, 0)), 1e-5))
        return torch.cat([xs, f], 1)

    xs, xf = stem(p + "s1.pathway0_stem", xs, 1), stem(p + "s1.pathway1_stem", xf, 5)
    xs = fuse(p + "s1_fuse", xs, xf)
    feats = []
    for si, depth in enumerate(SF_DEPTHS):
        stride = 1 if si == 0 else 2
        for i in range(depth):
            st = stride if i == 0 else 1
            xs = _res_block(sd, f"{p}s{si + 2}.pathway0_res{i}", xs, st, _bottleneck(SF_TK_SLOW[si]))
            xf = _res_block(sd, f"{p}s{si + 2}.pathway1_res{i}", xf, st, _bottleneck(SF_TK_FAST[si]))
        if si < 3:
            xs = fuse(f"{p}s{si + 2}_fuse", xs, xf)
        feats.append(xs)
    return feats


def resnet18_audio(sd: SD, p: str, x):
    """ResNet18 trunk without pool/fc on a 1-channel spectrogram.  backbones/resnet.py:17-54,131-143"""
    x = F.conv2d(x, sd[p + "conv1.weight"], None, 2, 3)
    x = F.relu(_bn(sd, p + "bn1", x, 1e-5))
    x = F.max_pool2d(x, 3, 2, 1)
    for li in range(1, 5):
        for bi in range(2):
            q = f"{p}layer{li}.{bi}"
            stride = 2 if (li > 1 and bi == 0) else 1
            idn = x
            o = F.conv2d(x, sd[q + ".conv1.weight"], None, stride, 1)
            o = F.relu(_bn(sd, q + ".bn1", o, 1e-5))
            o = F.conv2d(o, sd[q + ".conv2.weight"], None, 1, 1)
            o = _bn(sd, q + ".bn2", o, 1e-5)
            if q + ".downsample.0.weight" in sd:
                idn = _bn(sd, q + ".downsample.1", F.conv2d(x, sd[q + ".downsample.0.weight"], None, stride), 1e-5)
            x = F.relu(o + idn)
    return x


def _ln_nchw(x, w, b, eps):
    return F.layer_norm(x.permute(0, 2, 3, 1), (x.shape[1],), w, b, eps).permute(0, 3, 1, 2)


def convnext_tiny_features(sd: SD, p: str, x) -> List[torch.Tensor]:
    """timm==0.6.12 convnext_tiny(features_only=True) (README.md:34; model_utils.py:361,380):
    stem conv4x4/s4 + LayerNorm2d(1e-6); stage = [LayerNorm2d + conv2x2/s2] + blocks;
    block = dw7x7 -> LN(1e-6) -> fc1 -> GELU(erf) -> fc2 -> *gamma -> + shortcut."""
    x = F.conv2d(x, sd[p + "stem_0.weight"], sd[p + "stem_0.bias"], 4)
    x = _ln_nchw(x, sd[p + "stem_1.weight"], sd[p + "stem_1.bias"], 1e-6)
    outs = []
    for s, depth in enumerate(CONVNEXT_DEPTHS):
        q = f"{p}stages_{s}."
        if s > 0:
            x = _ln_nchw(x, sd[q + "downsample.0.weight"], sd[q + "downsample.0.bias"], 1e-6)
            x = F.conv2d(x, sd[q + "downsample.1.weight"], sd[q + "downsample.1.bias"], 2)
        for j in range(depth):
            b = f"{q}blocks.{j}."
            c = x.shape[1]
            h = F.conv2d(x, sd[b + "conv_dw.weight"], sd[b + "conv_dw.bias"], 1, 3, 1, c).permute(0, 2, 3, 1)
            h = F.layer_norm(h, (c,), sd[b + "norm.weight"], sd[b + "norm.bias"], 1e-6)
            h = F.linear(F.gelu(F.linear(h, sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"])),
                         sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"])
            x = x + (h * sd[b + "gamma"]).permute(0, 3, 1, 2)
        outs.append(x)
    return outs


def image_encoder(sd: SD, p: str, frames):
    """StaticSaliencyModelConvNext.forward: taps s16/s32 -> smooth convs.  model_utils.py:357-385"""
    _, _, o1, o0 = convnext_tiny_features(sd, p + "encoder.", frames)
    o0 = F.relu(_bn(sd, p + "smooth_0.1", F.conv2d(o0, sd[p + "smooth_0.0.weight"], sd[p + "smooth_0.0.bias"], 1, 1), 1e-5))
    o1 = F.relu(_bn(sd, p + "smooth_1.1", F.conv2d(o1, sd[p + "smooth_1.0.weight"], sd[p + "smooth_1.0.bias"], 1, 1), 1e-5))
    return o1, o0


def up_hw(x, k):
    """nn.Upsample(scale_factor=(1,k,k), mode='trilinear', align_corners=False).  model_utils.py:486-488"""
    return F.interpolate(x, scale_factor=(1, k, k), mode="trilinear", align_corners=False)


def adapter(sd: SD, p: str, o3, o2, num_frames: int):
    """Adapter.forward: (b t) c h w -> b c t h w, temporal max-pool to 4 frames, up2 the coarse map,
    concat, Inception.  model_utils.py:202-220"""
    stride = num_frames // 4
    bt, c3, h3, w3 = o3.shape
    b = bt // num_frames
    o3 = o3.view(b, num_frames, c3, h3, w3).permute(0, 2, 1, 3, 4)
    o2 = o2.view(b, num_frames, *o2.shape[1:]).permute(0, 2, 1, 3, 4)
    o3 = F.max_pool3d(o3, (stride, 1, 1), (stride, 1, 1))
    o2 = F.max_pool3d(o2, (stride, 1, 1), (stride, 1, 1))
    return mixed_block(sd, p + "conv", torch.cat([o3, up_hw(o2, 2)], 1))


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """get_sinusoid_encoding_table: fp64 numpy, cast to fp32.  model_utils.py:18-29"""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    tab = pos / np.power(10000, 2 * (j // 2) / d_hid)
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.tensor(tab, dtype=torch.float)


def vit_block(sd: SD, p: str, x, heads=4):
    """Pre-LN ViT block, qkv without bias, proj with bias, MLP x4 GELU.  model_utils.py:84-152"""
    b, n, c = x.shape
    h = F.layer_norm(x, (c,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
    qkv = F.linear(h, sd[p + "attn.qkv.weight"]).reshape(b, n, 3, heads, c // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    a = ((q @ k.transpose(-2, -1)) * (c // heads) ** -0.5).softmax(-1)
    h = (a @ v).transpose(1, 2).reshape(b, n, c)
    x = x + F.linear(h, sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"])
    h = F.layer_norm(x, (c,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"])),
                 sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + h


def sync_block(sd: SD, p: str, v4, aud):
    """SyncBlock.forward: tokens (t h w) / (h t), Linear+LN / LN, + sinusoid table, 3 ViT blocks.
    model_utils.py:257-282"""
    vis = v4.flatten(2).transpose(1, 2)                       # b (t h w) c
    au = aud.flatten(2).transpose(1, 2)                       # b (h t) c   ('b c h t -> b (h t) c')
    vis = F.linear(vis, sd[p + "vis_proj.weight"], sd[p + "vis_proj.bias"])
    vis = F.layer_norm(vis, (512,), sd[p + "vis_norm.weight"], sd[p + "vis_norm.bias"], 1e-5)
    au = F.layer_norm(au, (512,), sd[p + "aud_norm.weight"], sd[p + "aud_norm.bias"], 1e-5)
    vis = vis + sinusoid_table(vis.shape[1], 512).to(vis.device)   # (.to: the same functions time PyTorch eager on a GPU,
    au = au + sinusoid_table(au.shape[1], 512).to(au.device)        #  bench.py --impl eager; a no-op on the CPU oracle)
    x = torch.cat([vis, au], 1)
    for i in range(3):
        x = vit_block(sd, f"{p}blocks.{i}.", x)
    return x


def _head(sd: SD, p: str, x, idx: Tuple[int, ...], last_norm: bool):
    """Linear -> LayerNorm -> ReLU chains of the SimSiam projector/predictor.  model_utils.py:404-435"""
    for n, i in enumerate(idx):
        x = F.linear(x, sd[f"{p}.{i}.weight"], sd[f"{p}.{i}.bias"])
        is_last = n == len(idx) - 1
        if not is_last or last_norm:
            w = sd[f"{p}.{i + 1}.weight"]
            x = F.layer_norm(x, (w.shape[0],), w, sd[f"{p}.{i + 1}.bias"], 1e-5)
            if not is_last:
                x = F.relu(x)
    return x


def simsiam_loss(sd: SD, vis_fea, aud_fea):
    """vis/aud pooled embeddings -> projector -> predictor -> negative cosine.  model_utils.py:285-290,545-552"""
    zv = _head(sd, "vis_projector", vis_fea.mean(1), (0, 3, 6), True)
    za = _head(sd, "aud_projector", aud_fea.mean(1), (0, 3, 6), True)
    pv = _head(sd, "mlp_vis", zv, (0, 3), False)
    pa = _head(sd, "mlp_aud", za, (0, 3), False)
    d = lambda a, b: -F.cosine_similarity(a, b.detach(), dim=-1).mean()  # stop-gradient on z, model_utils.py:285-290
    return 0.5 * (d(pv, za) + d(pa, zv))


def convnext_block3d(sd: SD, p: str, x):
    """ConvNextBlock: dw(7,1,1) -> dw(1,7,7) -> LN over C -> 1x1x1 C->4C -> GELU -> 1x1x1 -> + input.
    model_utils.py:306-354 (no layer scale)."""
    c = x.shape[1]
    h = F.conv3d(x, sd[p + ".dwconv_t.weight"], sd[p + ".dwconv_t.bias"], 1, (3, 0, 0), 1, c)
    h = F.conv3d(h, sd[p + ".dwconv_s.weight"], sd[p + ".dwconv_s.bias"], 1, (0, 3, 3), 1, c)
    h = F.layer_norm(h.permute(0, 2, 3, 4, 1), (c,), sd[p + ".norm.norm.weight"], sd[p + ".norm.norm.bias"], 1e-5)
    h = h.permute(0, 4, 1, 2, 3)
    h = F.gelu(F.conv3d(h, sd[p + ".pwconv1.weight"], sd[p + ".pwconv1.bias"]))
    return x + F.conv3d(h, sd[p + ".pwconv2.weight"], sd[p + ".pwconv2.bias"])


def lateral(sd: SD, p: str, x, temporal_stride: Optional[int]):
    """latlayer_k: 1x1x1 (+bias) -> [ (s,1,1)/s conv, no bias ] -> ConvNextBlock.  model_utils.py:437-484"""
    x = F.conv3d(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
    i = 1
    if temporal_stride:
        x = F.conv3d(x, sd[p + ".1.weight"], None, (temporal_stride, 1, 1))
        i = 2
    return convnext_block3d(sd, f"{p}.{i}", x)


def sa_gate(sd: SD, p: str, x, masks, k: int):
    """SA: BasicConv3d 512->32 k3 -> up k -> conv(1,3,3) 32->1 (+bias) -> sigmoid; x*m + x.  model_utils.py:155-170"""
    m = basic_conv3d(sd, p + ".conv_mask.0", masks, 1, 1)
    if k != 1:
        m = up_hw(m, k)
    m = torch.sigmoid(F.conv3d(m, sd[p + ".conv_mask.2.weight"], sd[p + ".conv_mask.2.bias"], 1, (0, 1, 1)))
    return x * m + x


def readout(sd: SD, p: str, x):
    """readout Sequential.  model_utils.py:490-504"""
    x = F.conv3d(x, sd[p + ".0.weight"], sd[p + ".0.bias"])
    x = F.conv3d(x, sd[p + ".1.weight"], sd[p + ".1.bias"], 1, 1)
    x = F.relu(_bn(sd, p + ".2", x, 1e-5))
    x = F.conv3d(x, sd[p + ".4.weight"], sd[p + ".4.bias"], 1, (0, 1, 1))
    x = F.relu(_bn(sd, p + ".5", x, 1e-5))
    x = up_hw(x, 4)
    x = F.relu(F.conv3d(x, sd[p + ".8.weight"], sd[p + ".8.bias"], (4, 1, 1)))
    x = F.relu(F.conv3d(x, sd[p + ".10.weight"], sd[p + ".10.bias"], 1, (0, 1, 1)))
    return F.conv3d(x, sd[p + ".12.weight"], sd[p + ".12.bias"], 1, (0, 1, 1))


# ================================================================================== the model
LATERAL_BOOL_S3D = (True, True, False, False)  # config.py:41
LATERAL_STRIDE = 2                             # config.py:63
# per motion encoder: tap channels (config.py:65-74), lateral temporal convs (config.py:38-47) and their stride (:63)
ENCODERS = {
    "s3d": {"embeds": S3D_EMBEDS, "lateral_bool": LATERAL_BOOL_S3D, "lateral_stride": 2},
    "x3dl": {"embeds": (24, 48, 96, 192), "lateral_bool": (True, True, True, True), "lateral_stride": 4},
    "slowfast4x16": {"embeds": (320, 640, 1280, 2048), "lateral_bool": (False, False, False, False), "lateral_stride": 2},
}


def motion_features(sd: SD, encoder: str, clips):
    if encoder == "s3d":
        return s3d_features(sd, "visnet.", clips)
    if encoder == "x3dl":
        return x3d_features(sd, "visnet.", clips)
    if encoder == "slowfast4x16":
        return slowfast_features(sd, "visnet.", clips)
    raise Exception("Invalid Motion Encoder!")  # get_video_backbones.py:28-29


def forward(sd: SD, clips: torch.Tensor, audios: Optional[torch.Tensor], taps: Optional[dict] = None,
            encoder: str = "s3d"):
    with torch.no_grad():
        return _forward(sd, clips, audios, taps, encoder)


def _forward(sd: SD, clips: torch.Tensor, audios: Optional[torch.Tensor], taps: Optional[dict] = None,
             encoder: str = "s3d"):
    """AudioVisualSaliencyModel.forward (audios given) / VisualSaliencyModel.forward (audios None) for the motion
    encoders s3d / x3dl / slowfast4x16.  model_utils.py:520-574, 685-702.  Returns (log_map [B,H,W], loss_av)."""
    enc = ENCODERS[encoder]
    lat = lambda k: enc["lateral_stride"] if enc["lateral_bool"][k] else None
    rec = (lambda k, v: taps.__setitem__(k, v)) if taps is not None else (lambda k, v: None)
    b, _, t, h, w = clips.shape
    frames = clips.permute(0, 2, 1, 3, 4).reshape(b * t, 3, h, w)
    o1, o0 = image_encoder(sd, "image_encoder.", frames)
    rec("image_encoder.o1", o1), rec("image_encoder.o0", o0)
    masks = adapter(sd, "adapter.", o1, o0, t)
    rec("adapter", masks)
    v1, v2, v3, v4 = motion_features(sd, encoder, clips)
    for i, v in enumerate((v1, v2, v3, v4)):
        rec(f"visnet.base{i + 1}", v)
    loss = torch.zeros((), device=clips.device)
    if audios is not None:
        aud = resnet18_audio(sd, "audnet.", audios)
        rec("audnet", aud)
        tok = sync_block(sd, "aud_vis_sync_block.", v4, aud)
        rec("aud_vis_sync_block", tok)
        nv = v4.shape[2] * v4.shape[3] * v4.shape[4]
        vis_tok, aud_tok = tok[:, :nv], tok[:, nv:]
        loss = simsiam_loss(sd, vis_tok, aud_tok)
        vis_sync = vis_tok.transpose(1, 2).reshape(b, 512, *v4.shape[2:])
        v4 = torch.cat([v4, vis_sync], 1)
    s3 = lateral(sd, "latlayer_3", v4, lat(3))
    s0 = lateral(sd, "latlayer_0", v1, lat(0))
    s1 = lateral(sd, "latlayer_1", v2, lat(1))
    s2 = lateral(sd, "latlayer_2", v3, lat(2))
    for i, s in enumerate((s0, s1, s2, s3)):
        rec(f"latlayer_{i}", s)
    s2 = sa_gate(sd, "sa_2", s2, masks, 1) + up_hw(s3, 2)
    s1 = sa_gate(sd, "sa_1", s1, masks, 2) + up_hw(s2, 2) + up_hw(s3, 4)
    s0 = sa_gate(sd, "sa_0", s0, masks, 4) + up_hw(s1, 2) + up_hw(s2, 4) + up_hw(s3, 8)
    rec("fuse.s2", s2), rec("fuse.s1", s1), rec("fuse.s0", s0)
    out = readout(sd, "readout", torch.cat([s0, up_hw(s1, 2), up_hw(s2, 4), up_hw(s3, 8)], 1))
    out = out.squeeze(1).squeeze(1)
    rec("readout", out)
    out = out - torch.logsumexp(out, dim=(1, 2), keepdim=True)
    return out, loss


# ============================================================================ training step
def trainable_keys(sd: SD) -> List[str]:
    """Parameters the optimiser sees (train.py:151-158): everything with a gradient outside audnet / image_encoder,
    in state_dict order (= named_parameters() order, which is the AdamW parameter order)."""
    return [k for k, v in sd.items() if v.is_floating_point() and not k.startswith(FROZEN_PREFIXES)
            and not k.endswith((".running_mean", ".running_var"))]


def train_grads(sd: SD, clips: torch.Tensor, audios: Optional[torch.Tensor], gt: torch.Tensor, encoder: str = "s3d",
                gamma: float = 1.0):
    """loss and gradients of one training step: model.train() + frozen_encoder(); output, loss_va = model(imgs, audio);
    loss = SalLoss()(output, label) + gamma * loss_va; loss.backward().  engine_train.py:19-20,37-38,74-75, train.py:31.
    Returns dict(loss, kl, cc, loss_va, out, grads{name: tensor}, stats{name: updated BN buffer})."""
    keys = trainable_keys(sd)
    work = dict(sd)
    for k in keys:
        work[k] = sd[k].detach().clone().requires_grad_(True)
    _TRAIN["on"], _TRAIN["stats"] = True, {}
    try:
        out, loss_va = _forward(work, clips, audios, None, encoder)
        parts = sal_loss(out, gt)
        loss = parts["loss"] + gamma * loss_va
        loss.backward()
        stats = _TRAIN["stats"]
    finally:
        _TRAIN["on"], _TRAIN["stats"] = False, None
    grads = {k: (work[k].grad if work[k].grad is not None else torch.zeros_like(work[k])) for k in keys}
    return {"loss": loss.detach(), "kl": parts["kl"].detach(), "cc": parts["cc"].detach(), "loss_va": loss_va.detach(),
            "out": out.detach(), "grads": grads, "stats": stats}


def adamw_step(p, g, m, v, step: int, lr: float = 1e-4, beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8,
               weight_decay: float = 0.0):
    """torch.optim.AdamW(lr=cfg.SOLVER.LR, weight_decay=0) single-tensor update (train.py:157-158); returns (p, m, v)."""
    p = p * (1 - lr * weight_decay)
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    denom = (v.sqrt() / math.sqrt(1 - beta2 ** step)) + eps
    return p - (lr / (1 - beta1 ** step)) * m / denom, m, v


# ============================================================================ loss and metrics
_EPS = 2.2204e-16


def kldiv(s, g):
    """utils/compute_saliency_metrics.py:9-31"""
    b = s.shape[0]
    s = s.reshape(b, -1)
    g = g.reshape(b, -1)
    s = s / s.sum(1, keepdim=True)
    g = g / g.sum(1, keepdim=True)
    return (g * torch.log(_EPS + g / (s + _EPS))).sum(1).mean()


def _minmax(x):
    """normalize_map, utils/compute_saliency_metrics.py:33-43"""
    mn, mx = x.min(1, keepdim=True)[0], x.max(1, keepdim=True)[0]
    return (x - mn) / (mx - mn)


def similarity(s, g):
    """utils/compute_saliency_metrics.py:46-72"""
    b = s.shape[0]
    s, g = _minmax(s.reshape(b, -1)), _minmax(g.reshape(b, -1))
    s = s / s.sum(1, keepdim=True)
    g = g / g.sum(1, keepdim=True)
    return torch.minimum(s, g).sum(1).mean()


def cc(s, g):
    """utils/compute_saliency_metrics.py:75-92 (torch.std is unbiased; it cancels)"""
    b = s.shape[0]
    s, g = s.reshape(b, -1), g.reshape(b, -1)
    s = (s - s.mean(1, keepdim=True)) / s.std(1, keepdim=True)
    g = (g - g.mean(1, keepdim=True)) / g.std(1, keepdim=True)
    return ((s * g).sum(1) / torch.sqrt((s * s).sum(1) * (g * g).sum(1))).mean()


def nss(s, fix):
    """utils/compute_saliency_metrics.py:95-108"""
    b = s.shape[0]
    s, fix = s.reshape(b, -1), fix.reshape(b, -1)
    s = (s - s.mean(1, keepdim=True)) / (s.std(1, keepdim=True) + _EPS)
    return ((s * fix).sum(1) / fix.sum(1)).mean()


def sal_loss(log_map, gt, fixations=None):
    """SalLoss.forward on exp(log_map): KLD - CC (- 0.1 NSS).  utils/loss.py:26-49"""
    p = log_map.exp()
    out = {"kl": kldiv(p, gt), "cc": cc(p, gt), "sim": similarity(p, gt)}
    out["nss"] = nss(p, fixations) if fixations is not None else torch.zeros(())
    out["loss"] = out["kl"] - out["cc"] - (0.1 * out["nss"] if fixations is not None else 0.0)
    return out


# ================================================================================ audio front end
def log_spectrogram(wave: torch.Tensor, frames_out: int = 111) -> torch.Tensor:
    """get_audio_feature after slicing: Spectrogram(n_fft=512, hop=160) = |stft|^2 with a periodic Hann
    window, center=True, reflect padding; log(x+1e-6); standardise each frame over the 257 bins
    (unbiased std, +1e-6); pad with 0.02 / crop to `frames_out` columns.  inference.py:44-60,
    avsp_dataloader.py:51-80.  wave: [B, n] -> [B, 1, 257, frames_out]"""
    win = torch.hann_window(512, periodic=True)
    spec = torch.stft(wave, 512, 160, 512, win, center=True, pad_mode="reflect", return_complex=True)
    p = torch.log(spec.abs() ** 2 + 1e-6)                     # [B, 257, frames]
    p = (p - p.mean(1, keepdim=True)) / (p.std(1, keepdim=True) + 1e-6)
    out = torch.full((wave.shape[0], 257, frames_out), 0.02)
    n = min(frames_out, p.shape[-1])
    out[:, :, :n] = p[:, :, :n]
    return out.unsqueeze(1)


# ================================================================================ post-processing
def postprocess(pred_map, img_size=(640, 480)):
    """process() after the forward, verbatim (cv2 is the reference's own dependency): blur the LOG map, exp, resize,
    min-max, round to uint8.  inference.py:65-91.  pred_map: [H,W] float32 numpy / tensor -> uint8 [h,w] numpy."""
    import cv2
    m = np.asarray(pred_map, dtype=np.float32)
    m = cv2.GaussianBlur(m, (11, 11), 0)
    m = np.exp(m)
    m = cv2.resize(m, img_size)
    m = (m - m.min()) / (m.max() - m.min())
    return np.round(m * 255).astype(np.uint8)


# ============================================================== seeded weights and inputs (shared)
def _spec_conv_bn(spec, p, cout, cin, k, bn_prefix=None, bias=False):
    spec[p + ".weight"] = ("conv", (cout, cin) + tuple(k))
    if bias:
        spec[p + ".bias"] = ("bias", (cout,))
    if bn_prefix:
        for s, kind in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
            spec[f"{bn_prefix}.{s}"] = (kind, (cout,))
        spec[f"{bn_prefix}.num_batches_tracked"] = ("count", ())


def _spec_basic(spec, p, cin, cout, k):
    _spec_conv_bn(spec, p + ".conv", cout, cin, k, p + ".bn")


def _spec_sep(spec, p, cin, cout, k):
    _spec_conv_bn(spec, p + ".conv_s", cout, cin, (1, k, k), p + ".bn_s")
    _spec_conv_bn(spec, p + ".conv_t", cout, cout, (k, 1, 1), p + ".bn_t")


def _spec_mixed(spec, p, cin, plan):
    b0, b1a, b1, b2a, b2, b3 = plan
    _spec_basic(spec, p + ".branch0.0", cin, b0, (1, 1, 1))
    _spec_basic(spec, p + ".branch1.0", cin, b1a, (1, 1, 1))
    _spec_sep(spec, p + ".branch1.1", b1a, b1, 3)
    _spec_basic(spec, p + ".branch2.0", cin, b2a, (1, 1, 1))
    _spec_sep(spec, p + ".branch2.1", b2a, b2, 3)
    _spec_basic(spec, p + ".branch3.1", cin, b3, (1, 1, 1))


def _spec_linear(spec, p, cout, cin, bias=True):
    spec[p + ".weight"] = ("linear", (cout, cin))
    if bias:
        spec[p + ".bias"] = ("bias", (cout,))


def _spec_ln(spec, p, c):
    spec[p + ".weight"] = ("ln_w", (c,))
    spec[p + ".bias"] = ("ln_b", (c,))


def _spec_res_block(sp, q, cin, cout, inner, tk_a, k_b, groups_b, first, se_dim=0):
    """ResBlock keys in module-registration order (resnet_helper.py:540-578, 270-351, 400-487)."""
    if first:
        _spec_conv_bn(sp, q + ".branch1", cout, cin, (1, 1, 1), q + ".branch1_bn")
    _spec_conv_bn(sp, q + ".branch2.a", inner, cin, (tk_a, 1, 1), q + ".branch2.a_bn")
    if groups_b == inner:
        sp[q + ".branch2.b.weight"] = ("dw", (inner, 1) + tuple(k_b))
        for s_, kind in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
            sp[f"{q}.branch2.b_bn.{s_}"] = (kind, (inner,))
        sp[q + ".branch2.b_bn.num_batches_tracked"] = ("count", ())
    else:
        _spec_conv_bn(sp, q + ".branch2.b", inner, inner, k_b, q + ".branch2.b_bn")
    if se_dim:
        _spec_conv_bn(sp, q + ".branch2.se.fc1", se_dim, inner, (1, 1, 1), None, bias=True)
        _spec_conv_bn(sp, q + ".branch2.se.fc2", inner, se_dim, (1, 1, 1), None, bias=True)
    _spec_conv_bn(sp, q + ".branch2.c", cout, inner, (1, 1, 1), q + ".branch2.c_bn")


def _spec_x3d(sp, v):
    _spec_conv_bn(sp, v + "s1.pathway0_stem.conv_xy", 24, 3, (1, 3, 3), None)
    sp[v + "s1.pathway0_stem.conv.weight"] = ("dw", (24, 1, 5, 1, 1))
    for s_, kind in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_m"), ("running_var", "bn_v")):
        sp[f"{v}s1.pathway0_stem.bn.{s_}"] = (kind, (24,))
    sp[v + "s1.pathway0_stem.bn.num_batches_tracked"] = ("count", ())
    cin = 24
    for si, (depth, cout, inner) in enumerate(zip(X3D_DEPTHS, X3D_OUT, X3D_INNER)):
        for i in range(depth):
            _spec_res_block(sp, f"{v}s{si + 2}.pathway0_res{i}", cin, cout, inner, 1, (3, 3, 3), inner, i == 0,
                            _se_width(inner) if (i + 1) % 2 == 1 else 0)
            cin = cout


def _spec_slowfast(sp, v):
    _spec_conv_bn(sp, v + "s1.pathway0_stem.conv", 64, 3, (1, 7, 7), v + "s1.pathway0_stem.bn")
    _spec_conv_bn(sp, v + "s1.pathway1_stem.conv", 8, 3, (5, 7, 7), v + "s1.pathway1_stem.bn")
    _spec_conv_bn(sp, v + "s1_fuse.conv_f2s", 16, 8, (5, 1, 1), v + "s1_fuse.bn")
    cs, cf = 64 + 16, 8
    for si, depth in enumerate(SF_DEPTHS):
        outs, outf = 256 * 2 ** si, 32 * 2 ** si
        ins, inf_ = 64 * 2 ** si, 8 * 2 ** si
        for pw, (cin, cout, inner, tk) in enumerate(((cs, outs, ins, SF_TK_SLOW[si]), (cf, outf, inf_, SF_TK_FAST[si]))):
            c = cin
            for i in range(depth):
                _spec_res_block(sp, f"{v}s{si + 2}.pathway{pw}_res{i}", c, cout, inner, tk, (1, 3, 3), 1, i == 0)
                c = cout
        if si < 3:
            _spec_conv_bn(sp, f"{v}s{si + 2}_fuse.conv_f2s", 2 * outf, outf, (5, 1, 1), f"{v}s{si + 2}_fuse.bn")
        cs, cf = outs + (2 * outf if si < 3 else 0), outf


def param_spec(audio: bool = True, encoder: str = "s3d") -> Dict[str, Tuple[str, tuple]]:
    """Every state_dict entry of the reference model for the given motion encoder: name -> (kind, shape).
    Order and names follow the reference module tree (model_utils.py:388-514)."""
    enc = ENCODERS[encoder]
    embeds = enc["embeds"]
    sp: Dict[str, Tuple[str, tuple]] = {}
    if audio:
        _spec_conv_bn(sp, "audnet.conv1", 64, 1, (7, 7), "audnet.bn1")
        cin = 64
        for li, c in enumerate((64, 128, 256, 512), 1):
            for bi in range(2):
                q = f"audnet.layer{li}.{bi}"
                _spec_conv_bn(sp, q + ".conv1", c, cin if bi == 0 else c, (3, 3), q + ".bn1")
                _spec_conv_bn(sp, q + ".conv2", c, c, (3, 3), q + ".bn2")
                if bi == 0 and li > 1:
                    _spec_conv_bn(sp, q + ".downsample.0", c, cin, (1, 1), q + ".downsample.1")
            cin = c
    e = "image_encoder.encoder."
    _spec_conv_bn(sp, e + "stem_0", 96, 3, (4, 4), None, bias=True)
    _spec_ln(sp, e + "stem_1", 96)
    prev = 96
    for s, (d, n) in enumerate(zip(CONVNEXT_DIMS, CONVNEXT_DEPTHS)):
        q = f"{e}stages_{s}."
        if s > 0:
            _spec_ln(sp, q + "downsample.0", prev)
            _spec_conv_bn(sp, q + "downsample.1", d, prev, (2, 2), None, bias=True)
        for j in range(n):
            b = f"{q}blocks.{j}."
            sp[b + "gamma"] = ("gamma", (d,))
            sp[b + "conv_dw.weight"] = ("dw", (d, 1, 7, 7))
            sp[b + "conv_dw.bias"] = ("bias", (d,))
            _spec_ln(sp, b + "norm", d)
            _spec_linear(sp, b + "mlp.fc1", 4 * d, d)
            _spec_linear(sp, b + "mlp.fc2", d, 4 * d)
        prev = d
    _spec_conv_bn(sp, "image_encoder.smooth_0.0", 320, 768, (3, 3), "image_encoder.smooth_0.1", bias=True)
    _spec_conv_bn(sp, "image_encoder.smooth_1.0", 96, 384, (3, 3), "image_encoder.smooth_1.1", bias=True)
    v = "visnet."
    if encoder == "s3d":
        _spec_sep(sp, v + "base1.0", 3, 64, 7)
        _spec_basic(sp, v + "base1.2", 64, 64, (1, 1, 1))
        _spec_sep(sp, v + "base1.3", 64, 192, 3)
        for name, (cin, plan) in S3D_MIXED.items():
            _spec_mixed(sp, v + name, cin, plan)
    elif encoder == "x3dl":
        _spec_x3d(sp, v)
    else:
        _spec_slowfast(sp, v)
    if audio:
        a = "aud_vis_sync_block."
        _spec_linear(sp, a + "vis_proj", 512, embeds[3])
        _spec_ln(sp, a + "vis_norm", 512)
        _spec_ln(sp, a + "aud_norm", 512)
        for i in range(3):
            b = f"{a}blocks.{i}."
            _spec_ln(sp, b + "norm1", 512)
            _spec_linear(sp, b + "attn.qkv", 1536, 512, bias=False)
            _spec_linear(sp, b + "attn.proj", 512, 512)
            _spec_ln(sp, b + "norm2", 512)
            _spec_linear(sp, b + "mlp.fc1", 2048, 512)
            _spec_linear(sp, b + "mlp.fc2", 512, 2048)
        # module order in the reference: vis_projector, mlp_vis, aud_projector, mlp_aud
        for proj, pred in (("vis_projector", "mlp_vis"), ("aud_projector", "mlp_aud")):
            dims = (512, 2048, 2048, 2048)
            for n, i in enumerate((0, 3, 6)):
                _spec_linear(sp, f"{proj}.{i}", dims[n + 1], dims[n])
                _spec_ln(sp, f"{proj}.{i + 1}", dims[n + 1])
            _spec_linear(sp, f"{pred}.0", 512, 2048)
            _spec_ln(sp, f"{pred}.1", 512)
            _spec_linear(sp, f"{pred}.3", 2048, 512)
    for k in range(4):
        cin = embeds[k] + (512 if (k == 3 and audio) else 0)
        p = f"latlayer_{k}"
        _spec_conv_bn(sp, p + ".0", DE, cin, (1, 1, 1), None, bias=True)
        i = 1
        if enc["lateral_bool"][k]:
            _spec_conv_bn(sp, p + ".1", DE, DE, (enc["lateral_stride"], 1, 1), None)
            i = 2
        q = f"{p}.{i}"
        sp[q + ".dwconv_t.weight"] = ("dw", (DE, 1, 7, 1, 1))
        sp[q + ".dwconv_t.bias"] = ("bias", (DE,))
        sp[q + ".dwconv_s.weight"] = ("dw", (DE, 1, 1, 7, 7))
        sp[q + ".dwconv_s.bias"] = ("bias", (DE,))
        _spec_ln(sp, q + ".norm.norm", DE)
        _spec_conv_bn(sp, q + ".pwconv1", 4 * DE, DE, (1, 1, 1), None, bias=True)
        _spec_conv_bn(sp, q + ".pwconv2", DE, 4 * DE, (1, 1, 1), None, bias=True)
    r = "readout"
    _spec_conv_bn(sp, r + ".0", DE, 4 * DE, (1, 1, 1), None, bias=True)
    _spec_conv_bn(sp, r + ".1", DE, DE, (3, 3, 3), r + ".2", bias=True)
    _spec_conv_bn(sp, r + ".4", 64, DE, (1, 3, 3), r + ".5", bias=True)
    _spec_conv_bn(sp, r + ".8", 32, 64, (4, 1, 1), None, bias=True)
    _spec_conv_bn(sp, r + ".10", 32, 32, (1, 3, 3), None, bias=True)
    _spec_conv_bn(sp, r + ".12", 1, 32, (1, 3, 3), None, bias=True)
    _spec_mixed(sp, "adapter.conv", *ADAPTER_MIXED)
    for k in range(3):
        _spec_basic(sp, f"sa_{k}.conv_mask.0", 512, 32, (3, 3, 3))
        _spec_conv_bn(sp, f"sa_{k}.conv_mask.2", 1, 32, (1, 3, 3), None, bias=True)
    return sp


def make_state_dict(seed: int = 0, init: str = "calibrated", audio: bool = True, encoder: str = "s3d") -> SD:
    """Deterministic random-init weights for every entry of param_spec().

    init='default'    mimics PyTorch's default init scale (U(-1/sqrt(fan_in), 1/sqrt(fan_in)) convs/linears,
                      identity norms, fresh BN statistics) — with it the S3D activations vanish;
    init='calibrated' He-normal fan-in weights (x ~1.0 gain) and non-trivial BN/LN statistics so that every
                      kernel on the path sees O(1) activations and the BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    cal = init == "calibrated"
    for name, (kind, shape) in param_spec(audio, encoder).items():
        if kind in ("conv", "linear", "dw"):
            fan_in = int(np.prod(shape[1:]))
            if cal:
                gain = 2.0 if kind != "linear" else 1.0
                if name.startswith(("readout.12", "sa_")) and kind == "conv" and shape[0] == 1:
                    gain = 1.0
                t = torch.randn(shape, generator=g) * math.sqrt(gain / fan_in)
            else:
                bound = 1.0 / math.sqrt(fan_in)
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "bias":
            t = torch.randn(shape, generator=g) * (0.05 if cal else 0.01)
        elif kind in ("bn_w", "ln_w"):
            t = 1.0 + (torch.rand(shape, generator=g) - 0.5) * (0.4 if cal else 0.0)
            if cal and name.endswith("branch2.c_bn.weight"):
                t = t * 0.25  # residual branches of the 16..55-block ResNets: keeps the trunk O(1) (cf. zero-init final BN)
        elif kind in ("bn_b", "ln_b"):
            t = torch.randn(shape, generator=g) * (0.1 if cal else 0.0)
        elif kind == "bn_m":
            t = torch.randn(shape, generator=g) * (0.1 if cal else 0.0)
        elif kind == "bn_v":
            t = 1.0 + (torch.rand(shape, generator=g) - 0.5) * (0.6 if cal else 0.0)
        elif kind == "gamma":
            t = torch.full(shape, 1e-6) if not cal else 0.1 + 0.1 * torch.rand(shape, generator=g)
        elif kind == "count":
            t = torch.zeros((), dtype=torch.long)
        else:
            raise KeyError(kind)
        sd[name] = t
    return sd


def make_inputs(batch: int, height: int, width: int, seed: int = 2023, frames: int = 16):
    """Synthetic clips [B,3,T,H,W] ~ N(0,1) (ImageNet-normalised frames) and spectrograms [B,1,257,111]
    (model_utils.py:708 uses exactly these random shapes); seed 2023 is the reference's (train.py:36)."""
    g = torch.Generator().manual_seed(seed)
    clips = torch.randn(batch, 3, frames, height, width, generator=g)
    audio = torch.randn(batch, 1, 257, 111, generator=g)
    return clips, audio


def make_gt(log_map: torch.Tensor, seed: int = 11):
    """Prediction-correlated ground truth + dense fixations (recipe B): gt = 0.7*minmax(exp(map))
    + 0.3*blobs; fixations ~ Bernoulli(0.1) over the map's top-2% pixels."""
    g = torch.Generator().manual_seed(seed)
    b, h, w = log_map.shape
    p = log_map.exp()
    flat = p.reshape(b, -1)
    mm = ((flat - flat.min(1, keepdim=True)[0]) / (flat.max(1, keepdim=True)[0] - flat.min(1, keepdim=True)[0])).view(b, h, w)
    yy, xx = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    blobs = torch.zeros(b, h, w)
    for i in range(b):
        for _ in range(3):
            cy, cx = torch.rand(1, generator=g).item() * h, torch.rand(1, generator=g).item() * w
            blobs[i] += torch.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * (min(h, w) / 10.0) ** 2))
        blobs[i] /= blobs[i].max()
    gt = 0.7 * mm + 0.3 * blobs
    thr = torch.quantile(flat, 0.98, dim=1).view(b, 1, 1)
    fix = ((p >= thr) & (torch.rand(b, h, w, generator=g) < 0.1)).float()
    for i in range(b):
        if fix[i].sum() == 0:
            fix[i].view(-1)[flat[i].argmax()] = 1.0
    return gt, fix
