"""CPU tests: the oracle restatement (oracle/mspi_oracle.py) against fixtures produced by the live,
unmodified reference (oracle/gen_golden.py -> tests/golden/).  fp32 on both sides; tolerances cover
only summation-order differences of the same arithmetic."""
import os

import pytest
import torch

from oracle import mspi_oracle as orc

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def _tap_samples(t, summ):
    flat = t.float().reshape(-1)
    return flat[::summ["stride"]][:summ["samples"].numel()]


TAP_MAP = {  # reference module hook name -> oracle tap name
    "visnet.base1": "visnet.base1", "visnet.base2": "visnet.base2", "visnet.base3": "visnet.base3",
    "visnet.base4": "visnet.base4",
    "visnet.0": "visnet.base1", "visnet.1": "visnet.base2", "visnet.2": "visnet.base3", "visnet.3": "visnet.base4",
    "image_encoder.0": "image_encoder.o1", "image_encoder.1": "image_encoder.o0", "audnet": "audnet",
    "adapter": "adapter", "aud_vis_sync_block": "aud_vis_sync_block", "latlayer_0": "latlayer_0",
    "latlayer_1": "latlayer_1", "latlayer_2": "latlayer_2", "latlayer_3": "latlayer_3",
}


@pytest.mark.parametrize("name", ["s3d_av_64x64_b2_cal", "s3d_av_64x96_b1_def", "s3d_v_64x64_b1_cal",
                                  "x3dl_av_64x64_b1_cal", "x3dl_av_64x96_b1_def", "sf_av_64x64_b1_cal", "sf_av_64x96_b1_def"])
def test_forward_matches_reference_small(name):
    """S3D, X3D-L (BASELINE config 3) and SlowFast 4x16 R50 (config 4) model variants vs the live reference."""
    fx = _load(name + ".pt")
    c = fx["case"]
    enc = c.get("encoder", "s3d")
    sd = orc.make_state_dict(c["wseed"], c["init"], audio=c["audio"], encoder=enc)
    clips, aud = orc.make_inputs(c["b"], c["h"], c["w"], c["iseed"])
    taps = {}
    out, loss = orc.forward(sd, clips, aud if c["audio"] else None, taps, encoder=enc)
    # same fp32 arithmetic, different kernels/summation order: 1e-4 on O(10) log-probabilities
    assert (out - fx["out"]).abs().max().item() < 2e-4 * max(1.0, fx["out"].abs().max().item())
    assert abs(float(loss) - fx["loss"]) < 1e-5
    for ref_name, summ in fx["taps"].items():
        if ref_name not in TAP_MAP:
            continue
        t = taps[TAP_MAP[ref_name]]
        assert tuple(t.shape) == summ["shape"], ref_name
        diff = (_tap_samples(t, summ) - summ["samples"]).abs().max().item()
        assert diff < 1e-4 * max(1.0, summ["absmax"]), (ref_name, diff)


def test_forward_matches_reference_default_shape():
    """B=1 at the reference's default 16x224x384 shape (config.py:12,14)."""
    fx = _load("s3d_av_224x384_b1_cal.pt")
    c = fx["case"]
    torch.set_num_threads(os.cpu_count() or 1)
    sd = orc.make_state_dict(c["wseed"], c["init"])
    clips, aud = orc.make_inputs(c["b"], c["h"], c["w"], c["iseed"])
    out, loss = orc.forward(sd, clips, aud)
    assert (out - fx["out"]).abs().max().item() < 5e-4 * fx["out"].abs().max().item()
    assert abs(float(loss) - fx["loss"]) < 1e-5


def test_metrics_match_reference_kats():
    cases = _load("metrics.pt")
    # KAT values quoted in SURVEY.md §8c (computed by the reference's own functions)
    assert abs(cases["kat1"]["kld"] - 0.121777266) < 1e-7 and abs(cases["kat2"]["nss"] + 0.168924779) < 1e-7
    for nm, c in cases.items():
        s, g, f = c["s"], c["gt"], c["fix"]
        for key, fn, args in (("kld", orc.kldiv, (s, g)), ("cc", orc.cc, (s, g)), ("sim", orc.similarity, (s, g)),
                              ("nss", orc.nss, (s, f))):
            assert abs(fn(*args).item() - c[key]) <= 1e-6 * max(1.0, abs(c[key])), (nm, key)
        logp = torch.log(s / s.sum((1, 2), keepdim=True))
        assert abs(orc.sal_loss(logp, g)["loss"].item() - c["loss"]) < 2e-6
        assert abs(orc.sal_loss(logp, g, f)["loss"].item() - c["loss_fix"]) < 2e-6


def test_audio_front_end_matches_reference():
    for nm, c in _load("audio.pt").items():
        got = orc.log_spectrogram(c["wave"])[0]
        assert got.shape == c["feat"].shape
        assert (got - c["feat"]).abs().max().item() < 1e-4, nm


def test_convnext_restatement_matches_torchvision():
    """The image encoder's arithmetic lives in un-vendored timm==0.6.12; pin our restatement against
    torchvision's independent ConvNeXt-T by mapping the timm-style keys onto torchvision's."""
    tv = pytest.importorskip("torchvision")
    m = tv.models.convnext_tiny(weights=None).eval()
    sd = {k: v for k, v in orc.make_state_dict(3, "calibrated").items() if k.startswith("image_encoder.encoder.")}
    p = "image_encoder.encoder."
    tsd = {}
    tsd["features.0.0.weight"], tsd["features.0.0.bias"] = sd[p + "stem_0.weight"], sd[p + "stem_0.bias"]
    tsd["features.0.1.weight"], tsd["features.0.1.bias"] = sd[p + "stem_1.weight"], sd[p + "stem_1.bias"]
    for s, depth in enumerate(orc.CONVNEXT_DEPTHS):
        q = f"{p}stages_{s}."
        if s > 0:
            for i, nm in ((0, "downsample.0"), (1, "downsample.1")):
                tsd[f"features.{2 * s}.{i}.weight"] = sd[q + nm + ".weight"]
                tsd[f"features.{2 * s}.{i}.bias"] = sd[q + nm + ".bias"]
        for j in range(depth):
            b, t = f"{q}blocks.{j}.", f"features.{2 * s + 1}.{j}."
            tsd[t + "layer_scale"] = sd[b + "gamma"].view(-1, 1, 1)
            tsd[t + "block.0.weight"], tsd[t + "block.0.bias"] = sd[b + "conv_dw.weight"], sd[b + "conv_dw.bias"]
            tsd[t + "block.2.weight"], tsd[t + "block.2.bias"] = sd[b + "norm.weight"], sd[b + "norm.bias"]
            tsd[t + "block.3.weight"], tsd[t + "block.3.bias"] = sd[b + "mlp.fc1.weight"], sd[b + "mlp.fc1.bias"]
            tsd[t + "block.5.weight"], tsd[t + "block.5.bias"] = sd[b + "mlp.fc2.weight"], sd[b + "mlp.fc2.bias"]
    missing, unexpected = m.load_state_dict(tsd, strict=False)
    assert not unexpected and all(k.startswith("classifier") for k in missing)
    x = torch.randn(2, 3, 64, 96, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        feats, h = [], x
        for i, layer in enumerate(m.features):
            h = layer(h)
            if i in (1, 3, 5, 7):
                feats.append(h)
        mine = orc.convnext_tiny_features(sd, p, x)
    for a, b in zip(mine, feats):
        assert (a - b).abs().max().item() < 1e-4 * max(1.0, b.abs().max().item())


def test_param_spec_other_encoders():
    assert len(orc.param_spec(True, "x3dl")) == 1670 and len(orc.param_spec(True, "slowfast4x16")) == 1189
    assert orc._se_width(54) == 8 and orc._se_width(216) == 16 and orc._se_width(432) == 32


def test_param_spec_counts():
    sp = orc.param_spec(True)
    assert len(sp) == 993  # 993 state_dict keys for the S3D variant
    n = sum(int(torch.tensor(s).prod()) if s else 1 for _, s in sp.values())
    assert n == 87609972


@pytest.mark.parametrize("name", ["train_s3d_av_64x64_b2_cal", "train_s3d_av_64x96_b2_def"])
def test_train_step_matches_reference(name):
    """Row a20: the oracle's train-mode forward (batch-statistics BatchNorm outside the frozen encoders), loss, autograd
    gradients of all 411 trainable tensors, BatchNorm buffer updates and the AdamW update against one optimisation step
    of the live reference (oracle/gen_golden.py: run_train_case)."""
    fx = _load(name + ".pt")
    c = fx["case"]
    sd = orc.make_state_dict(c["wseed"], c["init"], audio=c["audio"], encoder=c["encoder"])
    clips, aud = orc.make_inputs(c["b"], c["h"], c["w"], c["iseed"])
    log_map = _load(c["fwd_fixture"] + ".pt")["out"] if c["fwd_fixture"] else orc.forward(sd, clips, aud)[0]
    gt, _ = orc.make_gt(log_map)
    r = orc.train_grads(sd, clips, aud, gt)
    assert abs(float(r["loss"]) - fx["loss"]) < 2e-5 and abs(float(r["loss_va"]) - fx["loss_va"]) < 1e-5
    assert abs(float(r["kl"]) - fx["kl"]) < 2e-5 and abs(float(r["cc"]) - fx["cc"]) < 2e-5
    assert (r["out"] - fx["out"]).abs().max().item() < 2e-4 * fx["out"].abs().max().item()
    keys = orc.trainable_keys(sd)
    assert keys == list(fx["grads"].keys()) and len(keys) == 411
    total = sum(g["norm"] ** 2 for g in fx["grads"].values()) ** 0.5
    worst = 0.0
    for k in keys:
        g, summ = r["grads"][k], fx["grads"][k]
        assert tuple(g.shape) == summ["shape"], k
        # same fp32 arithmetic in a different summation order; tiny gradients are judged against the global scale
        tol = 2e-3 * max(summ["absmax"], 1e-6 * total)
        diff = (_tap_samples(g, summ) - summ["samples"]).abs().max().item()
        worst = max(worst, diff / tol)
        assert diff < tol, (k, diff, summ["absmax"])
        assert abs(g.norm().item() - summ["norm"]) < 2e-3 * summ["norm"] + 1e-6 * total, k
    for k, summ in fx["buffers_after"].items():
        diff = (_tap_samples(r["stats"][k], summ) - summ["samples"]).abs().max().item()
        assert diff < 1e-5 * max(1.0, summ["absmax"]), (k, diff)
    # AdamW, first step (train.py:157): m = v = 0
    for k in keys[::37]:
        p1, _, _ = orc.adamw_step(sd[k], r["grads"][k], torch.zeros_like(sd[k]), torch.zeros_like(sd[k]), 1)
        summ = fx["params_after"][k]
        got, ref0 = _tap_samples(p1, summ), _tap_samples(sd[k], summ)
        # the first AdamW step is lr*sign(g) wherever |g| >> eps: compare where the reference moved by the full lr
        moved = (summ["samples"] - ref0).abs() > 0.99e-4
        assert ((got - summ["samples"]).abs()[moved] < 2e-6).all(), k
