"""Parity harness shared by the GPU tests and __graft_entry__.smoke(): runs the CUDA path
(mspi_b200, through the C ABI) and the CPU oracle on identical seeded weights and inputs and
reports per-tap and final-map errors."""
from __future__ import annotations

import copy
import io
import contextlib
import time

import torch

from oracle import mspi_oracle as orc


def build_product_model(sd, audio=True, height=None, encoder="s3d"):
    from mspi_b200.config import cfg as base_cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel, VisualSaliencyModel
    cfg = select_motion_encoder(encoder, copy.deepcopy(base_cfg))
    cls = AudioVisualSaliencyModel if audio else VisualSaliencyModel
    with contextlib.redirect_stdout(io.StringIO()):
        model = cls(cfg, load_pretrained=False)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    return model.cuda().eval()


def minmax(x):
    b = x.shape[0]
    f = x.reshape(b, -1)
    mn, mx = f.min(1, keepdim=True)[0], f.max(1, keepdim=True)[0]
    return ((f - mn) / (mx - mn)).view_as(x)


def rel_l2(a, b):
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def run_forward_parity(height=64, width=64, batch=1, init="calibrated", seed=0, audio=True, input_seed=2023,
                       tap_tol=3e-2, map_tol=1e-2, verbose=False, encoder="s3d"):
    sd = orc.make_state_dict(seed, init, audio=audio, encoder=encoder)
    clips, aud = orc.make_inputs(batch, height, width, input_seed)
    taps_ref = {}
    t0 = time.time()
    ref_out, ref_loss = orc.forward(sd, clips, aud if audio else None, taps_ref, encoder=encoder)
    t_cpu = time.time() - t0
    model = build_product_model(sd, audio, encoder=encoder)
    model.keep_taps = True
    if audio:
        out, loss = model(clips.cuda(), aud.cuda())
    else:
        out, loss = model(clips.cuda())
    torch.cuda.synchronize()
    plan = next(iter(model._plans.values()))
    res = {"taps": {}, "cpu_seconds": t_cpu, "launches": len(plan.steps)}
    worst = 0.0
    for name, act in plan.taps.items():
        if name not in taps_ref:
            continue
        got = act.to_ncdhw().cpu()
        ref = taps_ref[name]
        if ref.dim() == 4:  # 2-D feature maps ([B*T, C, H, W] or [B, C, H, W]) vs our T-major buffers
            got = got.permute(0, 2, 1, 3, 4).reshape(ref.shape) if got.shape[2] != 1 or got.shape[0] != ref.shape[0] \
                else got.squeeze(2)
        e = rel_l2(got, ref)
        res["taps"][name] = e
        worst = max(worst, e)
        if verbose:
            print(f"  tap {name:24s} rel-L2 {e:.3e}  |ref| {ref.abs().mean():.3e}")
    if audio and hasattr(plan, "taps_stream"):
        e = rel_l2(plan.taps_stream.cpu(), taps_ref["aud_vis_sync_block"])
        res["taps"]["aud_vis_sync_block"] = e
        worst = max(worst, e)
        if verbose:
            print(f"  tap aud_vis_sync_block      rel-L2 {e:.3e}")
    out_c = out.cpu()
    res["logit_maxabs"] = (out_c - ref_out).abs().max().item()
    res["map_maxabs_minmax"] = (minmax(out_c.exp()) - minmax(ref_out.exp())).abs().max().item()
    res["loss_abs"] = abs(float(loss) - float(ref_loss))
    res["sum_exp"] = out_c.exp().sum((1, 2)).tolist()
    res["worst_tap"] = worst
    res["ok"] = bool(worst < tap_tol and res["map_maxabs_minmax"] < map_tol and res["loss_abs"] < 1e-2)
    res["ref_out"], res["out"] = ref_out, out_c
    return res
