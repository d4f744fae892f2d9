"""TEST INFRASTRUCTURE — generates tests/golden/*.pt by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden
The reference is imported through oracle/ref_shim.py, loaded with the seeded weights of
oracle.mspi_oracle.make_state_dict and fed make_inputs; what it returns is stored as small fixtures
(full output maps, per-tap statistics and strided samples).  tests/test_oracle_cpu.py then holds the
oracle restatement to these vectors, and the GPU parity tests compare the CUDA path with the oracle.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import mspi_oracle as orc  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

# (name, audio, B, H, W, init, weight seed, input seed[, encoder])
CASES = [
    ("x3dl_av_64x64_b1_cal", True, 1, 64, 64, "calibrated", 3, 2023, "x3dl"),
    ("x3dl_av_64x96_b1_def", True, 1, 64, 96, "default", 4, 2023, "x3dl"),
    ("sf_av_64x64_b1_cal", True, 1, 64, 64, "calibrated", 5, 2023, "slowfast4x16"),
    ("sf_av_64x96_b1_def", True, 1, 64, 96, "default", 6, 2023, "slowfast4x16"),
    ("s3d_av_64x64_b2_cal", True, 2, 64, 64, "calibrated", 0, 2023),
    ("s3d_av_64x96_b1_def", True, 1, 64, 96, "default", 1, 2023),
    ("s3d_v_64x64_b1_cal", False, 1, 64, 64, "calibrated", 2, 2023),
    ("s3d_av_224x384_b1_cal", True, 1, 224, 384, "calibrated", 0, 2023),
]

TAP_MODULES = ["audnet", "adapter", "aud_vis_sync_block", "latlayer_0", "latlayer_1", "latlayer_2", "latlayer_3",
               "sa_0", "sa_1", "sa_2", "readout"]


def summarize(t: torch.Tensor, n_samples: int = 256) -> dict:
    """Shape, moments and a deterministic strided sample of a tap (keeps fixtures small)."""
    flat = t.detach().float().reshape(-1)
    stride = max(1, flat.numel() // n_samples)
    return {"shape": tuple(t.shape), "mean": flat.mean().item(), "std": flat.std().item(),
            "absmax": flat.abs().max().item(), "stride": stride, "samples": flat[::stride][:n_samples].clone()}


def run_case(name, audio, b, h, w, init, wseed, iseed, encoder="s3d"):
    sd = orc.make_state_dict(wseed, init, audio=audio, encoder=encoder)
    tokens = (16 if encoder == "x3dl" else 4) * (h // 32) * (w // 32)
    model = ref_shim.build_reference_model(sd, encoder, num_vis_tokens=tokens, audio=audio)
    ref_sd = model.state_dict()
    assert list(ref_sd.keys()) == list(sd.keys()), "param_spec order/names differ from the reference state_dict"
    for k in sd:
        assert tuple(ref_sd[k].shape) == tuple(sd[k].shape), k
    clips, aud = orc.make_inputs(b, h, w, iseed)
    taps = {}
    hooks = []

    def hook(nm):
        def f(_m, _i, o):
            if isinstance(o, (list, tuple)):
                for i, oo in enumerate(o):
                    taps[f"{nm}.{i}"] = summarize(oo)
            else:
                taps[nm] = summarize(o)
        return f

    for nm in TAP_MODULES + ["visnet", "image_encoder"]:
        if hasattr(model, nm):
            hooks.append(getattr(model, nm).register_forward_hook(hook(nm)))
    if encoder != "s3d":  # the PySlowFast-style encoders return lists per stage: tap the four feature maps
        orig = model.visnet.forward

        def wrapped(x, _o=orig):
            feats = _o(x)
            for i, f in enumerate(feats):
                taps[f"visnet.base{i + 1}"] = summarize(f)
            return feats
        model.visnet.forward = wrapped
    with torch.no_grad():
        out, loss = model(clips, aud) if audio else model(clips)
    for hk in hooks:
        hk.remove()
    fix = {"case": dict(name=name, audio=audio, b=b, h=h, w=w, init=init, wseed=wseed, iseed=iseed, encoder=encoder),
           "out": out.clone(), "loss": float(loss), "taps": taps}
    torch.save(fix, os.path.join(GOLDEN, name + ".pt"))
    print(f"{name}: out range [{out.min():.4f}, {out.max():.4f}] loss {float(loss):.6f} taps {len(taps)}")


# (name, forward fixture the ground truth is derived from, audio, B, H, W, init, weight seed, input seed)
TRAIN_CASES = [
    ("train_s3d_av_64x64_b2_cal", "s3d_av_64x64_b2_cal", True, 2, 64, 64, "calibrated", 0, 2023),
    ("train_s3d_av_64x96_b2_def", None, True, 2, 64, 96, "default", 1, 2023),
]


def run_train_case(name, fwd_fixture, audio, b, h, w, init, wseed, iseed, encoder="s3d"):
    """One optimisation step of the UNMODIFIED reference (engine_train.py:19-20,37-38,74-76; train.py:151-158):
    model.train(); frozen_encoder(); SalLoss()(out, gt) + loss_va; backward; AdamW(lr 1e-4, wd 0).step().
    Stored: the losses, a summary (norm + strided samples) of every gradient, of every updated parameter and of every
    BatchNorm buffer after the step."""
    import importlib
    sd = orc.make_state_dict(wseed, init, audio=audio, encoder=encoder)
    tokens = (16 if encoder == "x3dl" else 4) * (h // 32) * (w // 32)
    model = ref_shim.build_reference_model(sd, encoder, num_vis_tokens=tokens, audio=audio)
    lossmod = importlib.import_module("utils.loss")
    clips, aud = orc.make_inputs(b, h, w, iseed)
    if fwd_fixture:
        log_map = torch.load(os.path.join(GOLDEN, fwd_fixture + ".pt"), weights_only=False)["out"]
    else:
        log_map = orc.forward(sd, clips, aud if audio else None, encoder=encoder)[0]
    gt, _fix = orc.make_gt(log_map)
    model.train()
    model.frozen_encoder()
    for n, p in model.named_parameters():
        if n.startswith("audnet") or n.startswith("image_encoder"):
            p.requires_grad_(False)
    opt = torch.optim.AdamW(filter(lambda p: p.requires_grad, model.parameters()), 1e-4, weight_decay=0)
    crit = lossmod.SalLoss()
    out, loss_va = model(clips, aud) if audio else model(clips)
    loss = crit(out, gt) + 1.0 * loss_va
    opt.zero_grad()
    loss.backward()
    grads = {n: summarize(p.grad, 64) | {"norm": p.grad.norm().item()} for n, p in model.named_parameters() if p.requires_grad}
    opt.step()
    after = {n: summarize(p, 32) for n, p in model.named_parameters() if p.requires_grad}
    bufs = {n: summarize(bf, 32) for n, bf in model.named_buffers()
            if n.endswith(("running_mean", "running_var")) and not n.startswith(("audnet", "image_encoder"))}
    fix = {"case": dict(name=name, fwd_fixture=fwd_fixture, audio=audio, b=b, h=h, w=w, init=init, wseed=wseed, iseed=iseed,
                        encoder=encoder),
           "loss": float(loss), "loss_va": float(loss_va), "kl": crit.log["kl"].val, "cc": crit.log["cc"].val,
           "out": out.detach().clone(), "grads": grads, "params_after": after, "buffers_after": bufs}
    torch.save(fix, os.path.join(GOLDEN, name + ".pt"))
    print(f"{name}: loss {float(loss):.6f} (kl {crit.log['kl'].val:.6f} cc {crit.log['cc'].val:.6f} va {float(loss_va):.6f}) "
          f"{len(grads)} grads, total norm {sum(g['norm'] ** 2 for g in grads.values()) ** 0.5:.4e}")


def run_metrics():
    ref_shim.install()
    import importlib
    m = importlib.import_module("utils.compute_saliency_metrics")
    lossmod = importlib.import_module("utils.loss")
    g = torch.Generator().manual_seed(5)
    cases = {}
    # KAT1 / KAT2 of SURVEY §8c
    s1 = torch.tensor([[[1., 2.], [3., 4.]]]); g1 = torch.tensor([[[0., 1.], [1., 2.]]]); f1 = torch.tensor([[[0., 0.], [1., 1.]]])
    i = torch.arange(48).reshape(2, 4, 6)
    s2 = (i % 7 + 1).float(); g2 = ((3 * i) % 5).float(); f2 = (i % 11 == 0).float()
    s3 = torch.rand(3, 56, 96, generator=g) ** 3; g3 = torch.rand(3, 56, 96, generator=g) ** 2
    f3 = (torch.rand(3, 56, 96, generator=g) < 0.01).float()
    for nm, (s, gt, fx) in {"kat1": (s1, g1, f1), "kat2": (s2, g2, f2), "rand": (s3, g3, f3)}.items():
        logp = torch.log(s / s.sum((1, 2), keepdim=True))
        crit = lossmod.SalLoss()
        l0 = crit(logp, gt).item()
        l1 = crit(logp, gt, fx).item()
        cases[nm] = {"s": s, "gt": gt, "fix": fx, "kld": m.kldiv(s, gt).item(), "cc": m.cc(s, gt).item(),
                     "sim": m.similarity(s, gt).item(), "nss": m.nss(s, fx).item(), "loss": l0, "loss_fix": l1}
        print(nm, {k: v for k, v in cases[nm].items() if isinstance(v, float)})
    torch.save(cases, os.path.join(GOLDEN, "metrics.pt"))


def run_audio():
    """inference.get_audio_feature run unmodified: torchaudio.load / os.path.exists are pointed at a
    synthetic 16 kHz waveform so no file or codec is needed."""
    ref_shim.install()
    import importlib
    import torchaudio
    with ref_shim.scratch_cwd():
        inf = importlib.import_module("inference")
    g = torch.Generator().manual_seed(9)
    wave = torch.randn(1, 16000 * 4, generator=g) * 0.1
    orig_load, orig_exists = torchaudio.load, os.path.exists
    torchaudio.load = lambda p, *a, **k: (wave.clone(), 16000)
    os.path.exists = lambda p: True if str(p).endswith("fake.wav") else orig_exists(p)
    out = {}
    try:
        for nm, (start, fps, snip) in {"train17": (10, 25.0, 16), "infer33": (3, 25.0, 32), "long": (0, 10.0, 32)}.items():
            feat = inf.get_audio_feature("fake.wav", start, fps, len_snippet=snip)
            s = int(round(start / fps * 16000)); e = int(round((start + snip + 1) / fps * 16000))
            out[nm] = {"wave": wave[:, s:e].clone(), "feat": feat.clone()}
            print("audio", nm, tuple(out[nm]["wave"].shape), tuple(feat.shape))
    finally:
        torchaudio.load, os.path.exists = orig_load, orig_exists
    torch.save(out, os.path.join(GOLDEN, "audio.pt"))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count() or 1)
    only = sys.argv[1:]
    if not only:
        run_metrics()
        run_audio()
    for c in CASES:
        if not only or any(o in c[0] for o in only):
            run_case(*c)
    for c in TRAIN_CASES:
        if not only or any(o in c[0] for o in only):
            run_train_case(*c)


if __name__ == "__main__":
    main()
