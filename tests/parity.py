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


def run_train_parity(height=64, width=64, batch=2, init="calibrated", seed=0, input_seed=2023, verbose=False,
                     grad_tol=3e-2, optimizer=True, emulate=True, audio=True):
    """One training step (row a20) on the CUDA path vs the oracle's autograd step on identical weights / inputs / GT.

    Gradients are compared per parameter tensor by relative L2 error, tiny tensors against the global gradient norm
    (the CUDA path multiplies in tf32: 2^-11 relative per product, fp32 accumulation)."""
    from mspi_b200.train_engine import TrainPlan
    sd = orc.make_state_dict(seed, init, audio=audio, encoder="s3d")
    clips, aud = orc.make_inputs(batch, height, width, input_seed)
    if not audio:
        aud = None
    log_map = orc.forward(sd, clips, aud)[0]
    gt, _ = orc.make_gt(log_map)
    t0 = time.time()
    ref = orc.train_grads(sd, clips, aud, gt)
    t_cpu = time.time() - t0
    plan = TrainPlan(sd, batch, clips.shape[2], height, width, keep_taps=True, audio=audio)
    loss_out = plan.forward_backward(clips.cuda(), aud.cuda() if audio else None, gt.cuda())
    torch.cuda.synchronize()
    ref_fp32 = ref
    if emulate:
        # the same oracle step with the CUDA path's operand rounding and its frozen-encoder features (oracle/precision.py)
        from oracle import precision as prec
        o1 = plan.taps["image_encoder.o1"].to_ncdhw().cpu().squeeze(2)
        o0 = plan.taps["image_encoder.o0"].to_ncdhw().cpu().squeeze(2)
        af = plan.taps["audnet"].to_ncdhw().cpu().squeeze(2) if audio else None
        ref = prec.train_grads_product_numerics(sd, clips, aud, gt, (o1, o0), af)
    if verbose:   # train-mode forward taps against the oracle's train-mode forward
        taps_ref = {}
        orc._TRAIN["on"], orc._TRAIN["stats"] = True, {}
        mixed_orig = orc.mixed_block

        def mixed_rec(sd_, p_, x_):
            y_ = mixed_orig(sd_, p_, x_)
            taps_ref["mixed:" + p_] = y_
            return y_

        orc.mixed_block = mixed_rec
        import contextlib as _cl
        ctx = prec.product_numerics(False, (o1, o0), af) if emulate else _cl.nullcontext()
        try:
            with torch.no_grad(), ctx:
                orc._forward(dict(sd), clips, aud, taps_ref, "s3d")
        finally:
            orc._TRAIN["on"], orc._TRAIN["stats"] = False, None
            orc.mixed_block = mixed_orig
        for name, act in plan.taps.items():
            if name in taps_ref:
                got, ref_t = act.to_ncdhw().cpu(), taps_ref[name]
                if ref_t.dim() == 4:
                    got = got.permute(0, 2, 1, 3, 4).reshape(ref_t.shape)
                extra = ""
                if name.startswith("mixed:"):
                    q = name[6:]
                    cs_ = [sd[q + k_].shape[0] for k_ in (".branch0.0.conv.weight", ".branch1.1.conv_t.weight",
                                                          ".branch2.1.conv_t.weight", ".branch3.1.conv.weight")]
                    o_ = 0
                    for c_ in cs_:
                        extra += f" {rel_l2(got[:, o_:o_ + c_], ref_t[:, o_:o_ + c_]):.2e}"
                        o_ += c_
                print(f"  tap {name:24s} rel-L2 {rel_l2(got, ref_t):.3e}{extra}")
        if audio:
            print(f"  tap aud_vis_sync_block      rel-L2 {rel_l2(plan.taps_stream.cpu(), taps_ref['aud_vis_sync_block']):.3e}")
    lo = loss_out.cpu().tolist()
    res = {"cpu_seconds": t_cpu, "launches": plan.num_launches,
           "loss": lo[0], "kl": lo[1], "cc": lo[2], "loss_va": lo[3],
           "ref_loss": float(ref["loss"]), "ref_kl": float(ref["kl"]), "ref_cc": float(ref["cc"]),
           "ref_loss_va": float(ref["loss_va"])}
    res["out_maxabs"] = (plan.out.cpu() - ref["out"]).abs().max().item()
    grads = plan.grads()
    total = sum(float(g.norm()) ** 2 for g in ref["grads"].values()) ** 0.5
    errs = {}
    for k in plan.param_keys:
        g, r = grads[k].cpu(), ref["grads"][k]
        errs[k] = float((g - r).norm()) / max(float(r.norm()), 1e-4 * total)
    res["grad_errs"] = errs
    if emulate:
        tot32 = sum(float(g.norm()) ** 2 for g in ref_fp32["grads"].values()) ** 0.5
        e32 = sorted(float((grads[k].cpu() - ref_fp32["grads"][k]).norm()) / max(float(ref_fp32["grads"][k].norm()), 1e-4 * tot32)
                     for k in plan.param_keys)
        res["vs_fp32_oracle"] = {"median_grad_err": e32[len(e32) // 2], "max_grad_err": e32[-1],
                                 "loss": float(ref_fp32["loss"]), "out_maxabs": (plan.out.cpu() - ref_fp32["out"]).abs().max().item()}
    res["worst_grad"] = max(errs.values())
    res["worst_key"] = max(errs, key=errs.get)
    got_total = sum(float(g.norm()) ** 2 for g in grads.values()) ** 0.5
    res["grad_norm"], res["ref_grad_norm"] = got_total, total
    # BatchNorm running buffers after the step
    berr = 0.0
    for k, v in ref["stats"].items():
        if k.endswith("num_batches_tracked"):
            assert int(plan.sd[k]) == int(v), k
            continue
        berr = max(berr, float((plan.sd[k].cpu() - v).abs().max()) / max(1.0, float(v.abs().max())))
    res["bn_buffer_err"] = berr
    if verbose:
        for k in reversed(plan.param_keys):
            print(f"  {errs[k]:9.3e}  |g| {float(ref['grads'][k].norm()):9.3e}  {k}")
    if optimizer:
        plan.optimizer_step()
        torch.cuda.synchronize()
        perr = 0.0
        for k in plan.param_keys[::17]:
            p1, _, _ = orc.adamw_step(sd[k], grads[k].cpu(), torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), 1)
            perr = max(perr, float((plan.sd[k].cpu() - p1).abs().max()))
        res["adamw_err"] = perr
    tol_l = 2e-3
    res["ok"] = bool(res["worst_grad"] < grad_tol and abs(res["loss"] - res["ref_loss"]) < tol_l * max(1.0, abs(res["ref_loss"]))
                     and berr < 1e-3 and res.get("adamw_err", 0.0) < 1e-6)
    return res


def run_train_segment_parity(height=64, width=64, batch=2, init="calibrated", seed=3, input_seed=2023, verbose=False):
    """Training-step parity under teacher forcing (row a20).

    Train-mode BatchNorm at random init makes the S3D forward chaotic (oracle/precision.py: a 1e-4 input perturbation of the
    fp32 oracle itself is 3e-2 after base4.1, and tf32 operand rounding alone moves the oracle's gradients by ~95%), so a
    whole-model gradient comparison cannot be tight for ANY tf32 implementation.  This harness therefore checks the step
    segment by segment: every segment of the oracle (stem..base1, each Mixed block, the Adapter's Inception, and the whole
    SyncBlock + heads + laterals + SA + fusion + readout + loss "decoder") is evaluated by mspi_oracle's own functions with
    the CUDA path's operand rounding (oracle/precision.py) on the CUDA run's INPUT of that segment, back-propagated from the
    CUDA run's gradient at the segment OUTPUT, and its output and parameter gradients are compared with the CUDA run's.
    Together the segments cover all 411 trainable tensors."""
    from mspi_b200.train_engine import TrainPlan
    from oracle import precision as prec
    sd = orc.make_state_dict(seed, init, audio=True, encoder="s3d")
    clips, aud = orc.make_inputs(batch, height, width, input_seed)
    gt, _ = orc.make_gt(orc.forward(sd, clips, aud)[0])
    plan = TrainPlan(sd, batch, clips.shape[2], height, width, keep_taps=True)
    loss_out = plan.forward_backward(clips.cuda(), aud.cuda(), gt.cuda())
    torch.cuda.synchronize()
    grads = {k: v.cpu() for k, v in plan.grads().items()}
    total = sum(float(g.norm()) ** 2 for g in grads.values()) ** 0.5
    T = lambda a: a.to_ncdhw().cpu()
    G = lambda a: plan.act_grad(a).to_ncdhw().cpu()
    taps = plan.taps
    res = {"segments": {}, "covered": set()}

    def params_under(prefixes):
        return [k for k in plan.param_keys if k.startswith(tuple(prefixes))]

    def run_segment(name, prefixes, fn, y_gpu, dy_gpu):
        """fn(work_sd) -> y (oracle, rounded numerics, train mode).  Compares y and d(params under prefixes)."""
        keys = params_under(prefixes)
        work = dict(sd)
        for k in keys:
            work[k] = sd[k].detach().clone().requires_grad_(True)
        orc._TRAIN["on"], orc._TRAIN["stats"] = True, {}
        try:
            with prec.product_numerics(False):
                y = fn(work)
                y.backward(dy_gpu)
            stats = orc._TRAIN["stats"]
        finally:
            orc._TRAIN["on"], orc._TRAIN["stats"] = False, None
        out_err = rel_l2(y_gpu, y.detach())
        for bk, bv in stats.items():    # BatchNorm running buffers after the step (torch.nn.BatchNorm semantics)
            if bk.endswith("num_batches_tracked"):
                assert int(plan.sd[bk]) == int(bv), bk
            else:
                res["bn_buffer_err"] = max(res.get("bn_buffer_err", 0.0),
                                           float((plan.sd[bk].cpu() - bv).abs().max()) / max(1.0, float(bv.abs().max())))
        worst, wkey = 0.0, None
        for k in keys:
            g_ref = work[k].grad if work[k].grad is not None else torch.zeros_like(sd[k])
            scale = max(float(g_ref.norm()), 1e-5 * total)
            # a bias feeding a batch-statistics BatchNorm has an exactly zero gradient: compare against the floor only
            e = float((grads[k] - g_ref).norm()) / scale
            if e > worst:
                worst, wkey = e, k
            res["covered"].add(k)
        res["segments"][name] = {"out_err": out_err, "worst_grad": worst, "worst_key": wkey, "n": len(keys)}
        if verbose:
            print(f"  segment {name:22s} out {out_err:.2e}  worst grad {worst:.2e}  ({wkey})")

    # ---- S3D: stem..base1, then every Mixed block on its own input --------------------------------------------------
    def seg_base1(w):
        x = orc.sep_conv3d(w, "visnet.base1.0", clips, 7, 2, 3)
        x = torch.nn.functional.max_pool3d(x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        x = orc.basic_conv3d(w, "visnet.base1.2", x)
        return orc.sep_conv3d(w, "visnet.base1.3", x, 3, 1, 1)

    v1 = taps["visnet.base1"]
    run_segment("visnet.base1", ["visnet.base1."], seg_base1, T(v1), G(v1))
    for name in [k[6:] for k in taps if k.startswith("mixed:")]:
        x_in = T(taps["mixed_in:" + name])
        out = taps["mixed:" + name]
        run_segment(name, [name + "."], lambda w, n=name, x=x_in: orc.mixed_block(w, n, x), T(out), G(out))

    # ---- everything after the encoders: SyncBlock, SimSiam heads, laterals, SA, fusion, readout, loss ----------------
    vs = [T(taps[f"visnet.base{i}"]) for i in range(1, 5)]
    masks = T(taps["adapter"])
    af = T(taps["audnet"]).squeeze(2)
    dec_prefixes = ["aud_vis_sync_block.", "vis_projector.", "mlp_vis.", "aud_projector.", "mlp_aud.", "latlayer_", "sa_", "readout."]
    keys = params_under(dec_prefixes)
    work = dict(sd)
    for k in keys:
        work[k] = sd[k].detach().clone().requires_grad_(True)
    saved = (orc.motion_features, orc.adapter)
    orc.motion_features = lambda sd_, enc, c: vs
    orc.adapter = lambda sd_, p, o3, o2, nf: masks
    dummy = (torch.zeros(1), torch.zeros(1))
    orc._TRAIN["on"], orc._TRAIN["stats"] = True, {}
    try:
        with prec.product_numerics(False, dummy, af):
            out, loss_va = orc._forward(work, clips, aud, None, "s3d")
            parts = orc.sal_loss(out, gt)
            loss = parts["loss"] + loss_va
            loss.backward()
    finally:
        orc._TRAIN["on"], orc._TRAIN["stats"] = False, None
        orc.motion_features, orc.adapter = saved
    lo = loss_out.cpu().tolist()
    res["loss"], res["ref_loss"] = lo[0], float(loss.detach())
    res["kl"], res["ref_kl"], res["cc"], res["ref_cc"] = lo[1], float(parts["kl"]), lo[2], float(parts["cc"])
    res["loss_va"], res["ref_loss_va"] = lo[3], float(loss_va)
    res["out_maxabs"] = (plan.out.cpu() - out.detach()).abs().max().item()
    worst, wkey, errs = 0.0, None, {}
    for k in keys:
        g_ref = work[k].grad if work[k].grad is not None else torch.zeros_like(sd[k])
        e = float((grads[k] - g_ref).norm()) / max(float(g_ref.norm()), 1e-5 * total)
        errs[k] = e
        if e > worst:
            worst, wkey = e, k
        res["covered"].add(k)
    res["segments"]["decoder"] = {"out_err": res["out_maxabs"], "worst_grad": worst, "worst_key": wkey, "n": len(keys),
                                  "median_grad": sorted(errs.values())[len(errs) // 2]}
    res["decoder_errs"] = errs
    # AdamW on the flat buffers (train.py:157-158), first step, given the CUDA path's own gradients
    plan.optimizer_step()
    torch.cuda.synchronize()
    perr = 0.0
    for k in plan.param_keys[::13]:
        p1, _, _ = orc.adamw_step(sd[k], grads[k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), 1)
        perr = max(perr, float((plan.sd[k].cpu() - p1).abs().max()))
    res["adamw_err"] = perr
    if verbose:
        for k in keys:
            print(f"  {errs[k]:9.3e}  |g| {float(work[k].grad.norm()) if work[k].grad is not None else 0.0:9.3e}  {k}")
        print(f"  segment decoder                out(maxabs) {res['out_maxabs']:.2e}  worst grad {worst:.2e}  ({wkey})")
    res["n_covered"], res["n_params"] = len(res["covered"]), len(plan.param_keys)
    res["worst_grad"] = max(s["worst_grad"] for s in res["segments"].values())
    res["worst_out"] = max(s["out_err"] for n, s in res["segments"].items() if n != "decoder")
    del res["covered"]
    return res
