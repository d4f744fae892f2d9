"""AudioVisualSaliencyModel / VisualSaliencyModel — drop-in nn.Module surface of the reference's
model/model_utils.py:388-702, executed by the B200 kernels.

Same constructor argument (`cfg`), same attribute / state_dict names (`audnet`, `image_encoder`,
`visnet`, `aud_vis_sync_block`, `vis_projector`, `mlp_vis`, `aud_projector`, `mlp_aud`, `latlayer_k`,
`readout`, `adapter`, `sa_k`), same `forward(clips[B,3,T,H,W] fp32, audios[B,1,257,111] fp32) ->
(log_map[B,H,W] fp32, loss_av)`, same `frozen_encoder()`.  The forward is inference-only (eval-mode
BatchNorm); the modules below hold parameters, the arithmetic runs in libmspi_b200.so.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from ..backbones.resnet import get_resnet18
from ..backbones.s3d import declare_basic, declare_mixed
from ..engine import ForwardPlan
from ..params import ParamNode, conv_bn, layer_norm, linear
from .get_video_backbones import video_motion_extractor

CONVNEXT_DIMS, CONVNEXT_DEPTHS = (96, 192, 384, 768), (3, 3, 9, 3)


class StaticSaliencyModelConvNext(ParamNode):
    """model_utils.py:357-385; `encoder` uses timm==0.6.12 FeatureListNet key names (stem_0, stem_1,
    stages_k.downsample.{0,1}, stages_k.blocks.j.{conv_dw,norm,mlp.fc1,mlp.fc2,gamma})."""

    def __init__(self):
        super().__init__()
        e = "encoder."
        conv_bn(self, e + "stem_0", None, 96, 3, (4, 4), bias=True, init="trunc")
        layer_norm(self, e + "stem_1", 96)
        prev = 96
        for s, (d, n) in enumerate(zip(CONVNEXT_DIMS, CONVNEXT_DEPTHS)):
            q = f"{e}stages_{s}."
            if s > 0:
                layer_norm(self, q + "downsample.0", prev)
                conv_bn(self, q + "downsample.1", None, d, prev, (2, 2), bias=True, init="trunc")
            for j in range(n):
                b = f"{q}blocks.{j}."
                self.put(b + "gamma", torch.full((d,), 1e-6))
                conv_bn(self, b + "conv_dw", None, d, 1, (7, 7), bias=True, init="trunc")
                layer_norm(self, b + "norm", d)
                linear(self, b + "mlp.fc1", 4 * d, d, init="trunc")
                linear(self, b + "mlp.fc2", d, 4 * d, init="trunc")
            prev = d
        conv_bn(self, "smooth_0.0", "smooth_0.1", 320, 768, (3, 3), bias=True)
        conv_bn(self, "smooth_1.0", "smooth_1.1", 96, 384, (3, 3), bias=True)


def load_image_encoder_checkpoint(module: nn.Module, ckpt: Dict[str, torch.Tensor], verbose: bool = True):
    """`image_encoder.load_state_dict(ckpt, strict=False)` of model_utils.py:514, made loud: strict=False silently leaves
    every unmatched tensor at its random init, so a checkpoint written under other key names (another timm release, a
    bare timm ConvNeXt instead of the FeatureListNet wrapper) would load as noise.  Keys of the plain timm ConvNeXt
    layout (`stem.0`, `stages.k.`) are mapped onto the FeatureListNet names (`stem_0`, `stages_k.`) the module declares;
    the matched / missing / unexpected counts are reported and a checkpoint that matches nothing raises."""
    import re
    own = module.state_dict()
    fixed = {}
    for k, v in ckpt.items():
        k2 = k[7:] if k.startswith("module.") else k
        k2 = re.sub(r"^(encoder\.)?stem\.(\d)\.", r"encoder.stem_\2.", k2) if re.match(r"^(encoder\.)?stem\.\d\.", k2) else k2
        k2 = re.sub(r"^(encoder\.)?stages\.(\d)\.", r"encoder.stages_\2.", k2) if re.match(r"^(encoder\.)?stages\.\d\.", k2) else k2
        fixed[k2] = v
    usable = {k: v for k, v in fixed.items() if k in own and tuple(own[k].shape) == tuple(v.shape)}
    shape_mismatch = sorted(k for k, v in fixed.items() if k in own and tuple(own[k].shape) != tuple(v.shape))
    res = module.load_state_dict(usable, strict=False)
    matched = len(usable)
    unexpected = sorted(k for k in fixed if k not in own)
    report = {"matched": matched, "of": len(own), "missing": list(res.missing_keys), "unexpected": unexpected,
              "shape_mismatch": shape_mismatch}
    if verbose:
        print(f"image_encoder checkpoint: matched {matched}/{len(own)} tensors, missing {len(res.missing_keys)}, "
              f"unexpected {len(unexpected)}, shape mismatch {len(shape_mismatch)}")
        for name, keys in (("missing", res.missing_keys), ("unexpected", unexpected), ("shape mismatch", shape_mismatch)):
            if keys:
                print(f"  {name}: {', '.join(list(keys)[:6])}{' ...' if len(keys) > 6 else ''}")
    if len(ckpt) and matched == 0:
        raise RuntimeError("image saliency encoder checkpoint matched none of the module's tensors (strict=False would have "
                           f"left the encoder at random init); first checkpoint keys: {list(ckpt)[:4]}, expected e.g. "
                           f"{list(own)[:2]}")
    module.load_report = report
    return report


class SyncBlock(ParamNode):
    """model_utils.py:223-282 (the sinusoid tables are not parameters and not in the state_dict)."""

    def __init__(self, num_blocks=3, num_vis_tokens=336, num_aud_tokens=36, vis_in_embed=1024, embed_dim=512):
        super().__init__()
        self.num_vis_tokens, self.num_aud_tokens = num_vis_tokens, num_aud_tokens
        linear(self, "vis_proj", 512, vis_in_embed, init="xavier")
        layer_norm(self, "vis_norm", 512)
        layer_norm(self, "aud_norm", 512)
        for i in range(num_blocks):
            b = f"blocks.{i}."
            layer_norm(self, b + "norm1", embed_dim)
            linear(self, b + "attn.qkv", 3 * embed_dim, embed_dim, bias=False, init="xavier")
            linear(self, b + "attn.proj", embed_dim, embed_dim, init="xavier")
            layer_norm(self, b + "norm2", embed_dim)
            linear(self, b + "mlp.fc1", 4 * embed_dim, embed_dim, init="xavier")
            linear(self, b + "mlp.fc2", embed_dim, 4 * embed_dim, init="xavier")


def _projector(cin, hidden):
    n = ParamNode()
    dims = (cin, hidden, hidden, hidden)
    for k, i in enumerate((0, 3, 6)):
        linear(n, str(i), dims[k + 1], dims[k])
        layer_norm(n, str(i + 1), dims[k + 1])
    return n


def _predictor(hidden, mid):
    n = ParamNode()
    linear(n, "0", mid, hidden)
    layer_norm(n, "1", mid)
    linear(n, "3", hidden, mid)
    return n


def _latlayer(cin, de, temporal_stride: Optional[int]):
    """model_utils.py:437-484 + ConvNextBlock :306-337 (trunc_normal .02 weights, zero bias)."""
    n = ParamNode()
    conv_bn(n, "0", None, de, cin, (1, 1, 1), bias=True)
    i = 1
    if temporal_stride:
        conv_bn(n, "1", None, de, de, (temporal_stride, 1, 1))
        i = 2
    q = str(i)
    conv_bn(n, q + ".dwconv_t", None, de, 1, (7, 1, 1), bias=True, init="trunc")
    conv_bn(n, q + ".dwconv_s", None, de, 1, (1, 7, 7), bias=True, init="trunc")
    layer_norm(n, q + ".norm.norm", de)
    conv_bn(n, q + ".pwconv1", None, 4 * de, de, (1, 1, 1), bias=True, init="trunc")
    conv_bn(n, q + ".pwconv2", None, de, 4 * de, (1, 1, 1), bias=True, init="trunc")
    return n


def _readout(de):
    n = ParamNode()
    conv_bn(n, "0", None, de, 4 * de, (1, 1, 1), bias=True)
    conv_bn(n, "1", "2", de, de, (3, 3, 3), bias=True)
    conv_bn(n, "4", "5", 64, de, (1, 3, 3), bias=True)
    conv_bn(n, "8", None, 32, 64, (4, 1, 1), bias=True)
    conv_bn(n, "10", None, 32, 32, (1, 3, 3), bias=True)
    conv_bn(n, "12", None, 1, 32, (1, 3, 3), bias=True)
    return n


def _sa(in_embed=512):
    n = ParamNode()
    declare_basic(n, "conv_mask.0", in_embed, in_embed // 16, (3, 3, 3))
    conv_bn(n, "conv_mask.2", None, 1, in_embed // 16, (1, 3, 3), bias=True)
    return n


def _adapter():
    n = ParamNode()
    declare_mixed(n, "conv", 416, (192, 96, 208, 16, 48, 64))
    return n


class _SaliencyBase(nn.Module):
    has_audio = True

    def __init__(self, cfg, aud_embed_dim=512, de_embed_dim=192, load_pretrained: bool = True):
        super().__init__()
        print("Motion Encoder is {}.".format(cfg.MODEL.MOTION_ENCODER))
        self.cfg = cfg
        vis_embed_dims = cfg.MODEL.MOTION_ENCODER_EMBEDS[cfg.MODEL.MOTION_ENCODER]
        num_vis_tokens = cfg.MODEL.NUM_VIS_TOKENS[cfg.MODEL.MOTION_ENCODER]
        if self.has_audio:
            self.audnet = get_resnet18(path=cfg.MODEL.AUDIO_ENCODER_WEIGHT, pretrained=load_pretrained)
        self.image_encoder = StaticSaliencyModelConvNext()
        self.visnet = video_motion_extractor(cfg)
        if self.has_audio:
            self.aud_vis_sync_block = SyncBlock(num_blocks=3, num_vis_tokens=num_vis_tokens,
                                                vis_in_embed=vis_embed_dims[-1], embed_dim=aud_embed_dim)
            hidden = 2048
            self.vis_projector = _projector(aud_embed_dim, hidden)
            self.mlp_vis = _predictor(hidden, 512)
            self.aud_projector = _projector(aud_embed_dim, hidden)
            self.mlp_aud = _predictor(hidden, 512)
        lb, ls = cfg.MODEL.LATERAL_BOOL, cfg.MODEL.LATERAL_STRIDE
        extra = aud_embed_dim if self.has_audio else 0
        self.latlayer_0 = _latlayer(vis_embed_dims[0], de_embed_dim, ls[0] if lb[0] else None)
        self.latlayer_1 = _latlayer(vis_embed_dims[1], de_embed_dim, ls[1] if lb[1] else None)
        self.latlayer_2 = _latlayer(vis_embed_dims[2], de_embed_dim, ls[2] if lb[2] else None)
        self.latlayer_3 = _latlayer(vis_embed_dims[3] + extra, de_embed_dim, ls[3] if lb[3] else None)
        self.readout = _readout(de_embed_dim)
        self.adapter = _adapter()
        self.sa_0 = _sa(512)
        self.sa_1 = _sa(512)
        self.sa_2 = _sa(512)
        self._plans: Dict[Tuple, ForwardPlan] = {}
        self._wcache: Dict = {}      # packed, uploaded weights per device: shared by the plans of every batch shape
        self._liveness: Dict = {}    # buffer life times per plan kind (activation arena, mspi_b200/arena.py)
        self.use_cuda_graph = False
        self.keep_taps = False
        if load_pretrained:
            # Load Pretrained Weights — same files, same failure mode as model_utils.py:512-514
            self.visnet.load_weight(cfg.MODEL.MOTION_ENCODER_WEIGHT)
            if self.has_audio:
                self.audnet.load_state_dict(torch.load(cfg.MODEL.AUDIO_ENCODER_WEIGHT, map_location="cpu"))
            load_image_encoder_checkpoint(self.image_encoder,
                                          torch.load(cfg.MODEL.IMAGE_SALIENCY_ENCODER_WEIGHT, map_location="cpu"))

    def frozen_encoder(self):
        if self.has_audio:
            self.audnet.eval()
        self.image_encoder.eval()

    # -- plan cache ---------------------------------------------------------------------------
    def load_state_dict(self, *a, **k):
        self._plans.clear()  # packed weights are derived from the parameters
        self._wcache.clear()
        self._train_state = None   # the fp32 master copy / AdamW moments belong to the old parameters (see training_state)
        return super().load_state_dict(*a, **k)

    def invalidate_plans(self):
        """Call after mutating parameters in place (the packed bf16 weights are cached per input shape)."""
        self._plans.clear()
        self._wcache.clear()
        self._train_state = None

    def plan_for(self, clips: torch.Tensor, mode: str = "full") -> ForwardPlan:
        """The plan for this input: fp32 clips [B,3,T,H,W] (the reference's contract) or uint8 frames [B,T,H,W,3]."""
        u8 = clips.dtype == torch.uint8
        if u8:
            b, t, h, w, c = clips.shape
        else:
            b, c, t, h, w = clips.shape
        if c != 3:
            raise RuntimeError(f"expected clips [B,3,T,H,W] (fp32) or [B,T,H,W,3] (uint8), got {tuple(clips.shape)}")
        if mode != "image_encoder" and t != self.cfg.DATA.NUM_FRAMES:
            raise RuntimeError(f"clip has {t} frames, cfg.DATA.NUM_FRAMES is {self.cfg.DATA.NUM_FRAMES}")
        graph = self.use_cuda_graph and mode == "full"
        key = (mode, u8, b, t, h, w, clips.device.index, graph, self.keep_taps)
        plan = self._plans.get(key)
        if plan is None:
            m = self.cfg.MODEL
            sd = self.state_dict()
            kw = dict(audio=self.has_audio, lateral_bool=tuple(m.LATERAL_BOOL), lateral_stride=tuple(m.LATERAL_STRIDE),
                      pool_stride=m.S3D.POOL_STRIDE, device=clips.device, keep_taps=self.keep_taps, encoder=m.MOTION_ENCODER,
                      weight_cache=self._wcache.setdefault(clips.device.index, {}), mode=mode, input_u8=u8)
            liveness = None
            if os.environ.get("MSPI_ARENA", "1") != "0":
                # activation memory re-use: the life time of every buffer is read off a probe build of the same plan at batch 1
                # (the sequence of allocations and launches does not depend on the batch size; checked below)
                lkey = (mode, u8, t, h, w, clips.device.index, self.keep_taps)
                liveness = self._liveness.get(lkey)
                if liveness is None:
                    probe = ForwardPlan(sd, 1, t, h, w, **kw)
                    liveness = self._liveness[lkey] = probe.compute_liveness()
                    del probe
            plan = ForwardPlan(sd, b, t, h, w, liveness=liveness, **kw)
            if liveness is not None and len(plan._allocs) != liveness["__count__"]:
                raise RuntimeError("activation arena: the plan's allocation sequence differs from its probe build "
                                   f"({len(plan._allocs)} vs {liveness['__count__']}); set MSPI_ARENA=0")
            if graph:
                plan.capture()
            self._plans[key] = plan
        return plan

    # -- per-frame image-encoder cache (sliding-window inference, inference.py:120-150) ---------------------------
    @torch.no_grad()
    def encode_frames(self, frames: torch.Tensor, chunk: int = 128):
        """The image saliency encoder (model_utils.py:357-385: ConvNeXt-T + smooth convs) on N frames, once per frame:
        frames fp32 [N,3,H,W] (normalised) or uint8 [N,H,W,3].  Returns the bf16 channels-last maps
        (o1 [N,H/16,W/16,96], o0 [N,H/32,W/32,320]) that forward_cached() indexes."""
        if not frames.is_cuda:
            raise RuntimeError("mspi_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        n = frames.shape[0]
        u8 = frames.dtype == torch.uint8
        h, w = (frames.shape[1], frames.shape[2]) if u8 else (frames.shape[2], frames.shape[3])
        o1 = torch.empty((n, h // 16, w // 16, 96), dtype=torch.bfloat16, device=frames.device)
        o0 = torch.empty((n, h // 32, w // 32, 320), dtype=torch.bfloat16, device=frames.device)
        with torch.cuda.device(frames.device):
            for i in range(0, n, chunk):
                part = frames[i:i + chunk].contiguous()
                k = part.shape[0]
                x = part.view(k, 1, h, w, 3) if u8 else part.float().view(k, 3, 1, h, w)
                plan = self.plan_for(x, mode="image_encoder")
                f1, f0 = plan.run(x, None)
                o1[i:i + k].copy_(f1.buf.view(k, h // 16, w // 16, 96))
                o0[i:i + k].copy_(f0.buf.view(k, h // 32, w // 32, 320))
        return o1, o0

    @torch.no_grad()
    def forward_cached(self, clips, audios, feats, frame_index):
        """forward() with the image-encoder features of every frame taken from `feats` = encode_frames(...) instead of being
        recomputed: frame_index [B,T] (int) names, for window b and position t, the row of the cache holding that frame
        (a time-flipped window simply lists its frames backwards).  Bit-identical to forward() on the same clips."""
        if not clips.is_cuda:
            raise RuntimeError("mspi_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        if self.training:
            raise RuntimeError("forward_cached() is an inference path: call model.eval()")
        clips = clips.contiguous() if clips.dtype == torch.uint8 else clips.contiguous().float()
        audios = audios.contiguous().float() if (self.has_audio and audios is not None) else None
        with torch.cuda.device(clips.device):
            plan = self.plan_for(clips, mode="cached")
            out, loss = plan.run(clips, audios, feats, frame_index.to(clips.device))
            return out.clone(), (loss[0].clone() if self.has_audio else 0)

    # -- training step (engine_train.py:27-76) ------------------------------------------------------
    def training_state(self, device):
        """The TrainState (fp32 master parameters, AdamW moments, step count, live BatchNorm buffers) every training plan of
        this model shares.  Created from the module's parameters at the first training call; load_state_dict() and
        invalidate_plans() drop it (the next call starts again from the module's parameters with zero moments — restore
        moments with load_optimizer_state_dict())."""
        from ..train_engine import TrainState
        st = getattr(self, "_train_state", None)
        if st is None or st.device != device:
            with torch.cuda.device(device):
                st = TrainState(self.state_dict(), device)
            self._train_state = st
        return st

    def training_plan(self, clips: torch.Tensor, lr: float = 1e-4, gamma: float = 1.0, world_size: int = 1,
                      betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        """The TrainPlan (mspi_b200/train_engine.py) for this batch shape.  Plans of different shapes share one TrainState, so
        a tail batch or a second resolution continues from the current weights, moments and step count."""
        from ..train_engine import TrainPlan
        if self.cfg.MODEL.MOTION_ENCODER != "s3d":
            raise NotImplementedError("the training step is implemented for the S3D motion encoder (BASELINE config 5)")
        b, c, t, h, w = clips.shape
        key = ("train", b, t, h, w, clips.device.index)
        plan = self._plans.get(key)
        if plan is None:
            m = self.cfg.MODEL
            with torch.cuda.device(clips.device):
                plan = TrainPlan(self.training_state(clips.device), b, t, h, w, audio=self.has_audio, lr=lr, gamma=gamma,
                                 device=clips.device, lateral_bool=tuple(m.LATERAL_BOOL), lateral_stride=tuple(m.LATERAL_STRIDE),
                                 world_size=world_size, pool_stride=m.S3D.POOL_STRIDE, betas=betas, eps=eps,
                                 weight_decay=weight_decay)
            self._plans[key] = plan
        plan.lr, plan.betas, plan.adam_eps, plan.weight_decay, plan.world_size = lr, betas, eps, weight_decay, world_size
        plan.set_gamma(gamma)
        self._last_train_plan = plan
        return plan

    def train_step(self, clips, audios, labels, lr: float = 1e-4, gamma: float = 1.0, allreduce=None, world_size: int = 1,
                   betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        """One optimisation step of engine_train.py:27-76 — `model.train(); model.frozen_encoder()` forward,
        `SalLoss()(output, label) + gamma*loss_va`, backward, (gradient all-reduce,) AdamW(lr, betas, eps, weight_decay).
        Returns a device tensor [loss, kld, cc, loss_va]."""
        if not clips.is_cuda:
            raise RuntimeError("mspi_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        with torch.cuda.device(clips.device):
            plan = self.training_plan(clips, lr, gamma, world_size, betas, eps, weight_decay)
            audios = audios.contiguous().float() if (self.has_audio and audios is not None) else None
            return plan.train_step(clips.contiguous().float(), audios, labels.contiguous().float(), allreduce).clone()

    def sync_from_training(self):
        """Copy the trained parameters and BatchNorm buffers of the shared training state back into this module."""
        st = getattr(self, "_train_state", None)
        if st is None:
            return
        with torch.no_grad():
            own = super().state_dict()
            for k, v in st.live.items():
                if k in own:
                    own[k].copy_(v)
        for k in [k for k in self._plans if k[0] != "train"]:
            del self._plans[k]
        self._wcache.clear()

    def optimizer_state_dict(self, lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0) -> dict:
        """AdamW state of the training step in torch.optim.AdamW.state_dict() form (per-parameter `step` / `exp_avg` /
        `exp_avg_sq` in named_parameters() order of the trainable tensors): what the reference's checkpoints store under
        'optimizer' (utils/optim.py:40-50) and `torch.optim.AdamW.load_state_dict` accepts."""
        st = getattr(self, "_train_state", None)
        if st is None:
            return {"state": {}, "param_groups": []}
        return st.optimizer_state_dict(lr, betas, eps, weight_decay)

    def load_optimizer_state_dict(self, osd: dict, device=None):
        """Restore AdamW moments and the step count (a torch.optim.AdamW.state_dict()) into the shared training state."""
        dev = device if device is not None else next(self.parameters()).device
        self.training_state(dev).load_optimizer_state_dict(osd)

    def _forward(self, clips, audios):
        if not clips.is_cuda:
            raise RuntimeError("mspi_b200 runs on CUDA (sm_100a) only: move the model inputs to the GPU; "
                               "there is no CPU fallback")
        if self.training:
            # engine_train.py:37 `output, loss_va = model(imgs, audio)` under model.train(); model.frozen_encoder(): the
            # train-mode forward (batch-statistics BatchNorm outside the frozen encoders, running buffers updated).  The
            # tensors carry no autograd graph — loss.backward() has no counterpart here; model.train_step() does forward,
            # loss, backward and AdamW in one call.
            if clips.dtype == torch.uint8:
                raise RuntimeError("the training plan takes normalised fp32 clips")
            with torch.cuda.device(clips.device):
                plan = self.training_plan(clips.float())
                aud = audios.contiguous().float() if (self.has_audio and audios is not None) else None
                out, loss = plan.forward_train(clips.contiguous().float(), aud)
                return out.clone(), loss[0].clone()
        # fp32 [B,3,T,H,W] normalised clips (the reference's contract) or uint8 [B,T,H,W,3] frames (normalised on the device)
        clips = clips.contiguous() if clips.dtype == torch.uint8 else clips.contiguous().float()
        if audios is not None:
            audios = audios.contiguous().float()
        with torch.cuda.device(clips.device):
            plan = self.plan_for(clips)
            out, loss = plan.run(clips, audios)
            return out.clone(), loss[0].clone()


class AudioVisualSaliencyModel(_SaliencyBase):
    """model_utils.py:388-574"""
    has_audio = True

    @torch.no_grad()
    def forward(self, clips, audios):
        return self._forward(clips, audios)


class VisualSaliencyModel(_SaliencyBase):
    """model_utils.py:576-702; returns (log_map, 0)."""
    has_audio = False

    @torch.no_grad()
    def forward(self, clips):
        out, _ = self._forward(clips, None)
        return out, 0
