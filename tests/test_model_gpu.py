"""GPU parity of the whole clip forward (CUDA kernels behind the C ABI) against the CPU oracle.

Tolerances (north_star): saliency maps within 1e-2 max-abs after min-max normalisation of exp(out);
CC/NSS/KLD/SIM within 1e-3 relative.  Intermediate taps are held to rel-L2 <= 3e-2 (bf16 operands with
fp32 accumulation give 4e-3..9e-3); they are checked because with the default init the
S3D activations vanish and the final map alone would not exercise those kernels."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def test_forward_parity_default_init_literal_contract():
    """north_star's gate: identical random-init weights (the reference's default init scale), identical synthetic
    clips; exp(out) min-max normalised within 1e-2 max-abs of the fp32 reference."""
    from tests.parity import run_forward_parity
    res = run_forward_parity(64, 96, 1, init="default", seed=1, verbose=True)
    print({k: v for k, v in res.items() if k not in ("ref_out", "out", "taps")})
    assert res["map_maxabs_minmax"] < 1e-2
    assert res["worst_tap"] < 3e-2, res["taps"]
    assert all(abs(s - 1.0) < 1e-3 for s in res["sum_exp"])  # log-softmax: probabilities sum to one
    assert res["loss_abs"] < 1e-3


def test_forward_parity_calibrated_init_taps():
    """He-normal weights + non-trivial BN statistics: every tap carries O(1) signal, so every kernel on the path is
    exercised.  The log-map spans ~20 nats here, so exp/min-max amplifies any bf16-level trunk error; the
    criterion is per-tap rel-L2 and the logit error relative to the map's range."""
    from tests.parity import run_forward_parity
    res = run_forward_parity(64, 64, 2, init="calibrated", seed=0, verbose=True)
    print({k: v for k, v in res.items() if k not in ("ref_out", "out", "taps")})
    assert res["worst_tap"] < 3e-2, res["taps"]
    rng = (res["ref_out"].max() - res["ref_out"].min()).item()
    assert res["logit_maxabs"] / rng < 5e-2
    assert res["loss_abs"] < 1e-3


def test_forward_parity_visual_only():
    from tests.parity import run_forward_parity
    res = run_forward_parity(64, 64, 1, init="default", seed=2, audio=False)
    assert res["worst_tap"] < 3e-2 and res["map_maxabs_minmax"] < 1e-2


def test_forward_default_shape_and_metrics():
    """B=1 at the reference's default 16x224x384 shape + end-to-end KLD/CC/SIM/NSS within 1e-3 relative, on a
    prediction-correlated ground truth with dense fixations (SURVEY.md §8d recipe B)."""
    from oracle import mspi_oracle as orc
    from tests.parity import run_forward_parity
    from mspi_b200.utils.loss import SalLoss
    res = run_forward_parity(224, 384, 1, init="default", seed=1)
    assert res["map_maxabs_minmax"] < 1e-2 and res["worst_tap"] < 3e-2
    gt, fix = orc.make_gt(res["ref_out"])
    ref = orc.sal_loss(res["ref_out"], gt, fix)
    crit = SalLoss()
    loss = crit(res["out"].cuda(), gt.cuda(), fix.cuda())
    got = {"kl": crit.log["kl"].val, "cc": crit.log["cc"].val, "sim": crit.log["sim"].val, "nss": crit.log["nss"].val}
    for k, v in got.items():
        r = ref[k].item()
        assert abs(v - r) <= 1e-3 * abs(r), (k, v, r)
    assert abs(loss.item() - ref["loss"].item()) <= 1e-3 * abs(ref["loss"].item())


@pytest.mark.parametrize("encoder,init,seed", [("x3dl", "calibrated", 3), ("x3dl", "default", 4),
                                               ("slowfast4x16", "calibrated", 5), ("slowfast4x16", "default", 6)])
def test_forward_parity_other_motion_encoders(encoder, init, seed):
    """BASELINE configs 3 and 4: MSPI with the X3D-L and SlowFast 4x16 R50 motion encoders (config.py:29-74),
    CUDA path vs the oracle (itself pinned to the live reference by tests/golden/{x3dl,sf}_*.pt)."""
    from tests.parity import run_forward_parity
    res = run_forward_parity(64, 96 if init == "default" else 64, 1, init=init, seed=seed, encoder=encoder, verbose=True)
    print({k: v for k, v in res.items() if k not in ("ref_out", "out", "taps")})
    assert res["worst_tap"] < 3e-2, res["taps"]
    assert res["loss_abs"] < 1e-3
    if init == "default":
        assert res["map_maxabs_minmax"] < 1e-2
    else:
        rng = (res["ref_out"].max() - res["ref_out"].min()).item()
        assert res["logit_maxabs"] / rng < 5e-2


@pytest.mark.parametrize("encoder,seed", [("x3dl", 4), ("slowfast4x16", 6)])
def test_forward_default_shape_other_motion_encoders(encoder, seed):
    """BASELINE configs 3 / 4 at the reference's default 16x224x384 shape (X3D-L: 16*7*12 = 1344 visual tokens + 36 audio
    tokens through the batched-GEMM attention; SlowFast: 336 + 36), B = 1, default init: the literal map contract and
    every tap."""
    from tests.parity import run_forward_parity
    res = run_forward_parity(224, 384, 1, init="default", seed=seed, encoder=encoder, verbose=True)
    print({k: v for k, v in res.items() if k not in ("ref_out", "out", "taps")})
    assert res["map_maxabs_minmax"] < 1e-2
    assert res["worst_tap"] < 3e-2, res["taps"]
    assert res["loss_abs"] < 1e-3
    assert all(abs(s - 1.0) < 1e-3 for s in res["sum_exp"])


def test_forward_parity_at_the_benchmarked_configuration():
    """The configuration bench.py times (B = 32 clips of 16x224x384 per GPU, one captured CUDA graph, CTA-pair GEMM
    instances, 2368-block grids, 30 GB of buffers): first and last clip of the graph-replayed batch against the fp32
    oracle on the same weights and inputs, min-max-normalised map within 1e-2."""
    import contextlib, copy, io
    import bench
    from mspi_b200.config import cfg as base_cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel
    torch.manual_seed(2023)
    with contextlib.redirect_stdout(io.StringIO()):
        model = AudioVisualSaliencyModel(select_motion_encoder("s3d", copy.deepcopy(base_cfg)), load_pretrained=False)
    model = model.cuda().eval()
    model.use_cuda_graph = True
    g = torch.Generator(device="cuda").manual_seed(2023)
    B = 32
    clips = torch.randn(B, 3, 16, 224, 384, device="cuda", generator=g)
    audio = torch.randn(B, 1, 257, 111, device="cuda", generator=g)
    model(clips, audio)                      # builds the plan, captures the graph
    out, _ = model(clips, audio)             # a replay, like the timed steps
    torch.cuda.synchronize()
    res = bench.parity_at_bench_config(model, clips, audio, out, "s3d")
    print(res)
    assert res["ok"], res
    assert all(abs(s - 1.0) < 1e-3 for s in res["sum_exp"])
    # every clip of the batch went through the same kernels: all 32 maps are finite, normalised log-probabilities
    assert torch.isfinite(out).all()
    assert (out.exp().sum((1, 2)) - 1.0).abs().max() < 1e-3
    # clips are independent in eval mode: the same clip at another batch position gives the same map (to rounding of the
    # batch-position-dependent tile schedule: identical kernels, identical per-row arithmetic)
    perm = torch.arange(B - 1, -1, -1, device="cuda")
    out2, _ = model(clips[perm].contiguous(), audio[perm].contiguous())
    assert (out2[perm] - out).abs().max() < 1e-4


def _small_model(encoder="s3d", seed=7):
    from oracle import mspi_oracle as orc
    from tests.parity import build_product_model
    sd = orc.make_state_dict(seed, "calibrated", encoder=encoder)
    return build_product_model(sd, True, encoder=encoder)


def test_feature_cache_matches_plain_forward_bit_for_bit():
    """Row f1: sliding windows over a 40-frame synthetic video.  The image encoder runs once per frame
    (model.encode_frames), every window — plain and time-flipped — indexes the cache (model.forward_cached); the maps must
    equal the plain forward's, which re-encodes all 16 frames of every window (inference.py:120-150), bit for bit."""
    model = _small_model()
    g = torch.Generator().manual_seed(3)
    n, T, H, W = 40, 16, 64, 96
    frames = torch.randn(n, 3, H, W, generator=g)
    cache = model.encode_frames(frames.cuda(), chunk=16)      # 3 chunks, the last one ragged
    assert cache[0].shape == (n, H // 16, W // 16, 96) and cache[1].shape == (n, H // 32, W // 32, 320)
    jobs = [(s, False) for s in range(0, n - T + 1, 3)] + [(0, True), (5, True), (24, True)]
    for j0 in range(0, len(jobs), 4):
        chunk = jobs[j0:j0 + 4]
        clips = torch.stack([torch.flip(frames[s:s + T].permute(1, 0, 2, 3), [1]) if f else frames[s:s + T].permute(1, 0, 2, 3)
                             for s, f in chunk]).contiguous().cuda()
        aud = torch.randn(len(chunk), 1, 257, 111, generator=g).cuda()
        index = torch.tensor([[s + (T - 1 - t if f else t) for t in range(T)] for s, f in chunk], dtype=torch.int32)
        ref, ref_loss = model(clips, aud)
        got, loss = model.forward_cached(clips, aud, cache, index)
        assert torch.equal(got, ref), (got - ref).abs().max()
        assert abs(float(loss) - float(ref_loss)) < 1e-6   # SimSiam loss: an atomic sum over (sample, pair) blocks


def test_activation_arena_reuses_memory_without_changing_results(monkeypatch):
    """mspi_b200/arena.py: the memory of dead activation buffers is handed to later allocations of the same graph branch.
    Same maps bit for bit with and without it (eager and captured graph with its three branches), several times less memory."""
    model = _small_model(seed=11)
    g = torch.Generator().manual_seed(6)
    clips = torch.randn(3, 3, 16, 96, 128, generator=g).cuda()
    aud = torch.randn(3, 1, 257, 111, generator=g).cuda()
    outs, mem = {}, {}
    for arena in ("0", "1"):
        for graph in (False, True):
            monkeypatch.setenv("MSPI_ARENA", arena)
            model.invalidate_plans()
            model._liveness.clear()
            model.use_cuda_graph = graph
            model(clips, aud)
            out, _ = model(clips, aud)
            torch.cuda.synchronize()
            outs[(arena, graph)] = out.clone()
            mem[(arena, graph)] = next(iter(model._plans.values())).bytes_alloc
    ref = outs[("0", False)]
    for k, o in outs.items():
        assert torch.equal(o, ref), (k, (o - ref).abs().max())
    assert mem[("1", True)] < 0.45 * mem[("0", True)], mem
    model.use_cuda_graph = False


@pytest.mark.parametrize("encoder", ["s3d", "x3dl"])
def test_programmatic_dependent_launch_does_not_change_results(encoder):
    """mspi_set_pdl: with programmatic dependent launch the next kernel's prologue overlaps the previous kernel's drain, and every
    kernel waits (griddepcontrol.wait) before its first global access — so the maps of the captured three-branch graph must
    be the same bits with and without it, replay after replay."""
    from mspi_b200 import _lib
    lib = _lib.load()
    model = _small_model(encoder=encoder, seed=13)
    g = torch.Generator().manual_seed(9)
    clips = torch.randn(3, 3, 16, 64, 96, generator=g).cuda()
    aud = torch.randn(3, 1, 257, 111, generator=g).cuda()
    prev = lib.mspi_set_pdl(0)
    try:
        outs = {}
        for on in (0, 1):
            lib.mspi_set_pdl(on)
            model.invalidate_plans()
            model.use_cuda_graph = True
            for _ in range(3):
                out, _l = model(clips, aud)
            torch.cuda.synchronize()
            outs[on] = out.clone()
            if on == 0:      # run-to-run spread without PDL: X3D's SE means are fp32 atomic sums, so its maps are not bit-stable
                again, _l = model(clips, aud)
                torch.cuda.synchronize()
                spread = (again - outs[0]).abs().max().item()
        diff = (outs[0] - outs[1]).abs().max().item()
        if encoder == "s3d":
            assert spread == 0.0 and diff == 0.0, (spread, diff)
        else:
            assert diff <= max(4 * spread, 1e-6 * outs[0].abs().max().item()), (spread, diff)
    finally:
        lib.mspi_set_pdl(prev)
        model.use_cuda_graph = False


def test_forward_accepts_uint8_frames():
    """uint8 [B,T,H,W,3] frames, normalised on the device (ToTensor + Normalize of inference.py:154-165 folded into the clip
    conversion kernel), give exactly the forward of the host-normalised fp32 clip."""
    from mspi_b200 import ops
    model = _small_model(seed=8)
    g = torch.Generator().manual_seed(5)
    u8 = torch.randint(0, 256, (2, 16, 64, 96, 3), generator=g, dtype=torch.uint8)
    mean, std = torch.tensor(ops.IMAGENET_MEAN).view(1, 1, 1, 1, 3), torch.tensor(ops.IMAGENET_STD).view(1, 1, 1, 1, 3)
    clips = ((u8.float() / 255.0 - mean) / std).permute(0, 4, 1, 2, 3).contiguous()
    aud = torch.randn(2, 1, 257, 111, generator=g)
    ref, _ = model(clips.cuda(), aud.cuda())
    got, _ = model(u8.cuda(), aud.cuda())
    assert torch.equal(got, ref), (got - ref).abs().max()


def test_inference_entry_point_on_synthetic_dataset(tmp_path):
    """inference.py (the reference's entry point, same CLI / dataset layout / output naming) end to end on a tiny
    synthetic AVAD-style dataset: 33 frames -> 18 forward windows + 15 time-flipped ones = one image per frame."""
    cv2 = pytest.importorskip("cv2")
    import numpy as np
    from scipy.io import wavfile
    import inference
    root, vname = tmp_path / "data", "V01"
    (root / "fold_lists").mkdir(parents=True)
    (root / "video_frames" / "AVAD" / vname).mkdir(parents=True)
    (root / "video_audio" / "AVAD" / vname).mkdir(parents=True)
    n = 33
    rng = np.random.default_rng(0)
    for i in range(n):
        img = (rng.random((90, 120, 3)) * 255).astype(np.uint8)
        cv2.imwrite(str(root / "video_frames" / "AVAD" / vname / f"img_{i + 1:05d}.jpg"), img)
    wavfile.write(str(root / "video_audio" / "AVAD" / vname / f"{vname}.wav"), 22050,
                  (rng.standard_normal(22050 * 3) * 0.1).astype(np.float32))
    (root / "fold_lists" / "AVAD_list_test_2_fps.txt").write_text(f"{vname} {n} 25\n")
    out = tmp_path / "out"
    inference.main(["--path_data", str(root), "--save_path", str(out), "--dataset", "AVAD", "--split", "2",
                    "--random_init", "--batch", "6"])
    files = sorted(os.listdir(out / vname))
    assert files == [f"img_{i + 1:05d}.jpg" for i in range(n)]
    img = cv2.imread(str(out / vname / files[20]), cv2.IMREAD_GRAYSCALE)
    assert img.shape == (480, 640) and img.max() > 200 and img.min() < 30


def test_inference_audio_features_match_oracle():
    """Per-window spectrograms of the driver (GPU STFT on slices of a once-resampled wav) vs the oracle front end."""
    import inference
    from oracle import mspi_oracle as orc
    g = torch.Generator().manual_seed(8)
    audio = torch.randn(16000 * 4, generator=g) * 0.1
    wins = [(0, False), (3, False), (3, True), (40, False)]
    feats = inference.audio_features(audio, wins, 25.0, torch.device("cuda")).cpu()
    for k, (s, flip) in enumerate(wins):
        w = inference.audio_window(audio, s, 25.0, flip=flip)
        ref = orc.log_spectrogram(w[None])[0]
        assert (feats[k] - ref).abs().max() < 2e-3
