"""CPU tests around the training step (row a20): host-side bookkeeping of the plan and the rounded-numerics oracle
(oracle/precision.py) that the GPU parity test uses.  No GPU compute here."""
import torch

from oracle import mspi_oracle as orc
from oracle import precision as prec


def test_trainable_keys_match_the_oracle_and_the_reference_counts():
    from mspi_b200.train_engine import trainable_keys
    sd = orc.make_state_dict(0, "default")
    keys = trainable_keys(sd)
    assert keys == orc.trainable_keys(sd)
    assert len(keys) == 411 and sum(sd[k].numel() for k in keys) == 46042052   # SURVEY 8(a) a20 [probe]
    assert not any(k.startswith(("audnet.", "image_encoder.")) for k in keys)   # train.py:151-155


def test_tf32_rounding_matches_the_b200_probe():
    """tools/tf32_round.py on B200: operands reach the tensor core rounded to nearest, ties to even (10 mantissa bits)."""
    f = lambda v: float(prec.round_tf32(torch.tensor([v], dtype=torch.float32))[0])
    assert f(1 + 2.0 ** -11 + 2.0 ** -13) == 1 + 2.0 ** -10
    assert f(1 + 2.0 ** -11) == 1.0                       # tie -> even
    assert f(1 + 3 * 2.0 ** -11) == 1 + 2 * 2.0 ** -10    # tie -> even
    assert f(1 + 2.0 ** -10 - 2.0 ** -20) == 1 + 2.0 ** -10
    assert f(-(1 + 2.0 ** -11 + 2.0 ** -13)) == -(1 + 2.0 ** -10)
    x = torch.randn(10000)
    r = prec.round_tf32(x)
    assert ((r - x).abs() <= x.abs() * 2.0 ** -11 + 1e-45).all() and (r.view(torch.int32) & 0x1FFF == 0).all()


def _decoder_grads(sd, clips, aud, gt, feats, emulate):
    import contextlib
    o1, o0, af, masks, vs = feats
    keys = [k for k in orc.trainable_keys(sd) if not k.startswith(("visnet.", "adapter."))]
    work = dict(sd)
    for k in keys:
        work[k] = sd[k].detach().clone().requires_grad_(True)
    saved = (orc.motion_features, orc.adapter, orc.image_encoder, orc.resnet18_audio)
    orc.motion_features = lambda sd_, enc, c: vs
    orc.adapter = lambda sd_, p, o3, o2, nf: masks
    orc.image_encoder = lambda *a: (o1, o0)
    orc.resnet18_audio = lambda *a: af
    orc._TRAIN["on"], orc._TRAIN["stats"] = True, {}
    try:
        with (prec.product_numerics(False, (o1, o0), af) if emulate else contextlib.nullcontext()):
            out, lva = orc._forward(work, clips, aud, None, "s3d")
            loss = orc.sal_loss(out, gt)["loss"] + lva
            loss.backward()
    finally:
        orc._TRAIN["on"], orc._TRAIN["stats"] = False, None
        orc.motion_features, orc.adapter, orc.image_encoder, orc.resnet18_audio = saved
    return {k: work[k].grad for k in keys}, float(loss.detach())


def test_rounded_oracle_sensitivity():
    """Documents why the GPU parity test is segment-wise: with the features feeding the decoder held fixed, rounding the GEMM
    operands to tf32 moves the oracle's decoder gradients by ~1e-2 (median); through the train-mode S3D (batch-statistics
    BatchNorm at random init) the same rounding moves the whole model's gradients by order one."""
    sd = orc.make_state_dict(3, "calibrated")
    clips, aud = orc.make_inputs(2, 64, 64, 2023)
    gt, _ = orc.make_gt(orc.forward(sd, clips, aud)[0])
    b, _, t, h, w = clips.shape
    frames = clips.permute(0, 2, 1, 3, 4).reshape(b * t, 3, h, w)
    orc._TRAIN["on"] = True
    try:
        with torch.no_grad():
            o1, o0 = orc.image_encoder(sd, "image_encoder.", frames)
            af = orc.resnet18_audio(sd, "audnet.", aud)
            masks = orc.adapter(sd, "adapter.", o1, o0, t)
            vs = orc.motion_features(sd, "s3d", clips)
    finally:
        orc._TRAIN["on"] = False
    feats = (o1, o0, af, masks, vs)
    g0, l0 = _decoder_grads(sd, clips, aud, gt, feats, False)
    g1, l1 = _decoder_grads(sd, clips, aud, gt, feats, True)
    tot = sum(float(g.norm()) ** 2 for g in g0.values()) ** 0.5
    errs = sorted(float((g1[k] - g0[k]).norm()) / max(float(g0[k].norm()), 1e-5 * tot) for k in g0)
    assert abs(l1 - l0) < 1e-3 * max(1.0, abs(l0))
    assert errs[len(errs) // 2] < 5e-2 and errs[-1] < 0.5, (errs[len(errs) // 2], errs[-1])
    # whole model: same rounding, the chaotic S3D in the loop
    ref = orc.train_grads(sd, clips, aud, gt)
    emu = prec.train_grads_product_numerics(sd, clips, aud, gt, (o1, o0), af)
    tot = sum(float(g.norm()) ** 2 for g in ref["grads"].values()) ** 0.5
    e = sorted(float((emu["grads"][k] - ref["grads"][k]).norm()) / max(float(ref["grads"][k].norm()), 1e-4 * tot)
               for k in ref["grads"])
    assert e[len(e) // 2] > 0.2, e[len(e) // 2]      # order-one deviation of the fp32 oracle from its own tf32 evaluation
    assert abs(float(emu["loss"]) - float(ref["loss"])) < 0.1
