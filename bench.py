#!/usr/bin/env python
"""Headline benchmark: MSPI-S3D clip forward, clips/s on N x B200 (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

One process per GPU (the driver launches N>1 under torch.distributed.run).  A "step" is one forward of
`--batch` synthetic clips [B,3,16,224,384] + spectrograms [B,1,257,111] per GPU through the C-ABI kernels,
followed (N>1) by the NCCL all-gather of the saliency maps.  Rank 0 prints ONE JSON line.

  value      clips/s, whole job, inputs resident in HBM before the timed region (CUDA-event timed, max over ranks)
  e2e        clips/s through the public nn.Module API with HOST inputs: pinned fp32 clips are copied H2D every
             step and the [B,H,W] maps are copied back D2H every step, inside the timed region
  roofline   dominant kernel = the tcgen05 implicit-GEMM conv kernel (bf16 instance): algorithmic FLOPs of its
             launches / their CUDA-event durations, against the measured bf16 peak (MEASURED_PEAKS.json)
  cpu_baseline  the oracle (CPU port of the reference forward) timed on this box's host cores, rank 0, N=1
"""
from __future__ import annotations

import argparse
import contextlib
import copy
import io
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T, H, W = 16, 224, 384
ALGO_GFLOP_PER_CLIP = 416.05  # 2xMAC of the reference forward at 16x224x384 (FlopCounterMode)
ALGO_GFLOP = {"s3d": 416.05, "x3dl": 414.75, "slowfast4x16": 434.61}
ENC_NAME = {"s3d": "MSPI-S3D", "x3dl": "MSPI-X3D-L", "slowfast4x16": "MSPI-SlowFast4x16-R50"}


def read_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"bf16_burst": d["bf16_tflops"], "bf16_sustained": d["bf16_tflops_sustained"], "hbm": d["hbm_gbs"],
                "source": "measured"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons for one GPU while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def mark(self):
        """Start of the timed region: samples taken before it (the sampler starts ahead of the warm-up so that nvidia-smi is
        already streaming when the short timed region begins) are only used if the region itself yields none."""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def wait_for_sample(self, load_fn, timeout: float = 4.0):
        """The timed region (10 steps = 0.3 s) can be shorter than nvidia-smi's first report on an 8-GPU box: keep the same
        load running, untimed and collective-free, until one sample taken under load exists."""
        self.extra = 0
        t_end = time.time() + timeout
        while self.proc and not any(r[0] >= self.t0 for r in self.rows) and time.time() < t_end:
            load_fn()
            self.extra += 1

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        t0 = getattr(self, "t0", 0.0)
        timed = [r[1:] for r in self.rows if r[0] >= t0]
        rows = timed if timed else [r[1:] for r in self.rows]
        t1 = getattr(self, "t1", None)
        inside = t1 is None or any(t0 <= r[0] <= t1 for r in self.rows)
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [c for c in sm if c > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm),
                "during": ("timed region" if inside else f"timed region + {getattr(self, 'extra', 0)} untimed steps of the same load")
                if timed else "warm-up + timed region"}


def read_traffic():
    """DRAM bytes per launch of the dominant kernel, from the committed ncu pass (profiles/r02_roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "r02_roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def cpu_baseline(seconds_budget: float = 20.0):
    """The oracle (CPU fp32 port of the reference forward) on the host cores: B=1 at the default shape."""
    import torch
    from oracle import mspi_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = orc.make_state_dict(1, "default")
    clips, aud = orc.make_inputs(1, H, W, 2023)
    orc.forward(sd, clips, aud)  # warm-up
    times = []
    t_end = time.time() + seconds_budget
    while len(times) < 3 or (time.time() < t_end and len(times) < 10):
        t0 = time.perf_counter()
        orc.forward(sd, clips, aud)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": 1.0 / best, "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": f"B=1 clip 16x{H}x{W} + audio, fp32, oracle port of the reference forward, best of {len(times)} "
                      f"({best:.3f} s/clip, torch {torch.__version__} CPU, {torch.get_num_threads()} threads)"}


def parity_at_bench_config(model, clips, audio, out, encoder, picks=None):
    """The checker for the configuration that is actually timed: clips `picks` (first and last by default) of the
    B-clip graph-replayed forward against the fp32 oracle on the same weights and inputs (outside every timed region).
    north_star gate: exp(out) min-max normalised within 1e-2 max-abs."""
    import torch
    from oracle import mspi_oracle as orc
    b = clips.shape[0]
    picks = sorted(set(picks if picks is not None else (0, b - 1)))
    sd = {k: v.detach().float().cpu() if v.is_floating_point() else v.detach().cpu() for k, v in model.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    c = clips[picks].float().cpu()
    a = audio[picks].float().cpu() if audio is not None else None
    t0 = time.time()
    ref, _ = orc.forward(sd, c, a, encoder=encoder)
    got = out[picks].float().cpu()

    def mm(x):
        f = x.exp().reshape(x.shape[0], -1)
        mn, mx = f.min(1, keepdim=True)[0], f.max(1, keepdim=True)[0]
        return (f - mn) / (mx - mn)

    err = (mm(got) - mm(ref)).abs().max(1)[0]
    return {"batch": b, "clips_checked": picks, "map_maxabs_minmax": [float(e) for e in err], "tolerance": 1e-2,
            "logit_maxabs": float((got - ref).abs().max()), "sum_exp": [float(v) for v in got.exp().sum((1, 2))],
            "ok": bool((err < 1e-2).all()), "oracle_seconds": time.time() - t0,
            "how": "graph-replayed forward of the timed batch vs oracle.forward (fp32, CPU) on the same weights and inputs"}


def gpu_eager_baseline(dev, batch, steps=5, warmup=3, modes=("tf32", "bf16_autocast_channels_last"), encoder="s3d"):
    """Same-silicon baseline: the reference forward as PyTorch eager ops (cuDNN / cuBLAS) on this GPU.
    /root/reference cannot travel to the GPU box, so the op sequence is the oracle's — a functional restatement of the
    reference's nn.Modules (same torch.nn.functional calls, pinned to the live reference by tests/golden) — executed on
    CUDA tensors: (i) fp32 storage with TF32 tensor cores allowed, (ii) bf16 autocast with channels-last weights and inputs;
    cudnn.benchmark on in both, as inference.py:189 sets it."""
    import torch
    from oracle import mspi_oracle as orc
    res = {}
    sd0 = orc.make_state_dict(1, "default", encoder=encoder)
    g = torch.Generator(device=dev).manual_seed(7)
    sets = [(torch.randn(batch, 3, T, H, W, device=dev, generator=g), torch.randn(batch, 1, 257, 111, device=dev, generator=g))
            for _ in range(2)]
    old = (torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        for mode in modes:
            sd = {}
            for k, v in sd0.items():
                v = v.to(dev)
                if mode.startswith("bf16") and v.dim() == 5:
                    v = v.contiguous(memory_format=torch.channels_last_3d)
                elif mode.startswith("bf16") and v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
                sd[k] = v
            ctx = (lambda: torch.autocast("cuda", dtype=torch.bfloat16)) if mode.startswith("bf16") else contextlib.nullcontext

            def fwd(i):
                c, a = sets[i % 2]
                if mode.startswith("bf16"):
                    c = c.contiguous(memory_format=torch.channels_last_3d)
                with ctx():
                    return orc.forward(sd, c, a, encoder=encoder)

            try:
                for i in range(warmup):
                    fwd(i)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(steps):
                    out, _ = fwd(i)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                res[mode] = {"value": batch / (ms / 1e3), "unit": "clips/s", "ms_per_step": ms, "batch": batch, "steps": steps,
                             "warmup": warmup, "finite": bool(torch.isfinite(out.float()).all())}
            except Exception as e:  # e.g. out of memory at this batch: report, do not fail the product's line
                res[mode] = {"unavailable": f"{type(e).__name__}: {str(e)[:200]}"}
            del sd
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.benchmark, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    res["what"] = ("the reference forward as PyTorch-eager cuDNN/cuBLAS ops on this GPU (oracle.forward on CUDA tensors, "
                   f"torch {torch.__version__}, cudnn.benchmark): tf32 = fp32 storage + TF32 tensor cores; "
                   "bf16_autocast_channels_last = torch.autocast(bf16) + channels_last(_3d) weights and clips")
    return res


def run_eager(args, rank, world, local):
    """--impl eager: the same-silicon baseline alone (one JSON line, rank 0)."""
    if rank != 0:
        return
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    with ClockSampler(local) as clocks:
        clocks.mark()
        res = gpu_eager_baseline(dev, args.batch, steps=max(1, args.steps), warmup=max(3, args.warmup), encoder=args.encoder)
    best = max((m["value"] for m in res.values() if isinstance(m, dict) and "value" in m), default=None)
    line = {"impl": "eager", "metric": f"clips/sec {ENC_NAME[args.encoder]} inference", "value": best, "unit": "clips/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(3, args.warmup), "higher_is_better": True, "dtype": "tf32 / bf16 autocast",
            "data": "synthetic", "config": {"workload": f"{ENC_NAME[args.encoder]} forward, {args.batch} clips/step of 16x{H}x{W}, "
                                                        "PyTorch eager (cuDNN/cuBLAS) on the same GPU"},
            "gpu_eager_baseline": res, "clocks": clocks.summary()}
    print(json.dumps(line), flush=True)


def run_sliding(args, rank, world, local):
    """--sliding: the reference's real workload (inference.py:120-150): one saliency map per video frame from stride-1 sliding
    windows of 16 frames.  Output frames/s with the image-encoder feature cache (one ConvNeXt pass per FRAME) against the plain
    forward (one per WINDOW: 15 of 16 frames re-encoded), on a synthetic device-resident video."""
    if rank != 0:
        return
    import torch
    assert torch.cuda.is_available()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from mspi_b200.config import cfg as base_cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel
    torch.manual_seed(2023)
    with contextlib.redirect_stdout(io.StringIO()):
        model = AudioVisualSaliencyModel(select_motion_encoder(args.encoder, copy.deepcopy(base_cfg)), load_pretrained=False)
    model = model.to(dev).eval()
    B, n_frames = args.batch, args.frames
    g = torch.Generator(device=dev).manual_seed(1)
    video = torch.randn(n_frames, 3, H, W, device=dev, generator=g)
    starts = list(range(0, n_frames - T + 1))
    starts = starts[: (len(starts) // B) * B]          # whole batches only (same work in both arms)
    aud = torch.randn(B, 1, 257, 111, device=dev, generator=g)
    base = torch.arange(T, dtype=torch.int32)

    def windows(j0):
        idx = torch.stack([base + s for s in starts[j0:j0 + B]])
        clips = video[idx.long().to(dev)].permute(0, 2, 1, 3, 4).contiguous()      # [B,3,T,H,W] assembled on the device
        return clips, idx

    def plain():
        for j0 in range(0, len(starts), B):
            clips, _ = windows(j0)
            out, _ = model(clips, aud)
        return out

    def cached():
        cache = model.encode_frames(video[: starts[-1] + T], chunk=16 * B)
        for j0 in range(0, len(starts), B):
            clips, idx = windows(j0)
            out, _ = model.forward_cached(clips, aud, cache, idx)
        return out

    res = {}
    with ClockSampler(local) as clocks:
        for name, fn in (("plain", plain), ("cached", cached)):
            o = fn()
            torch.cuda.synchronize()
            clocks.mark()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(max(1, args.steps)):
                o = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / max(1, args.steps)
            res[name] = {"frames_per_s": len(starts) / (ms / 1e3), "ms_per_video": ms, "out": o}
    same = bool(torch.equal(res["plain"].pop("out"), res["cached"].pop("out")))
    line = {"metric": f"output frames/sec {ENC_NAME[args.encoder]} sliding-window inference", "value": res["cached"]["frames_per_s"],
            "unit": "frames/s", "n_gpus": 1, "steps": args.steps, "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"stride-1 windows of {T} frames over a {n_frames}-frame synthetic 224x384 video, {B} windows per "
                                   f"forward, {len(starts)} output frames; eager launches (no CUDA graph)"},
            "feature_cache": res["cached"], "plain_forward": res["plain"],
            "speedup_from_cache": res["cached"]["frames_per_s"] / res["plain"]["frames_per_s"],
            "last_batch_bit_identical": same, "clocks": clocks.summary()}
    print(json.dumps(line), flush=True)


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Python and cannot
    travel to the GPU box, so its CPU port (oracle/, pinned to the live reference by tests/golden) is timed."""
    if rank != 0:
        return
    import torch
    from oracle import mspi_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = orc.make_state_dict(1, "default")
    clips, aud = orc.make_inputs(1, H, W, 2023)
    for _ in range(max(1, min(args.warmup, 2))):
        orc.forward(sd, clips, aud)
    steps = max(1, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(steps):
        orc.forward(sd, clips, aud)
    dt = (time.perf_counter() - t0) / steps
    v = 1.0 / dt
    line = {"impl": "reference", "metric": "clips/sec MSPI-S3D inference", "value": v, "unit": "clips/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"MSPI-S3D clip forward, bounded sample: B=1 clip 16x{H}x{W} per step on host cores"},
            "cpu_baseline": {"value": v, "unit": "clips/s", "cores": cores, "kind": "port",
                             "sample": f"B=1 clip per step, {steps} steps, oracle port of the reference forward"},
            "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def train_cpu_baseline(batch: int):
    """The oracle's training step (train-mode forward, loss, autograd backward of the 411 trainable tensors) on the host."""
    import torch
    from oracle import mspi_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = orc.make_state_dict(1, "default")
    clips, aud = orc.make_inputs(batch, H, W, 2023)
    gt = torch.rand(batch, H, W)
    t0 = time.perf_counter()
    orc.train_grads(sd, clips, aud, gt)
    dt = time.perf_counter() - t0
    return {"value": batch / dt, "unit": "clips/s", "cores": cores, "kind": "port",
            "sample": f"one step, B={batch} clips 16x{H}x{W} + audio + GT maps, fp32 oracle port of the reference training step "
                      f"(forward + autograd backward, optimiser excluded), {dt:.2f} s, {torch.get_num_threads()} threads"}


def run_train(args, rank, world, local):
    """--train: BASELINE config 5 — MSPI-S3D training step (KLD - CC + SimSiam loss, backward, NCCL gradient all-reduce,
    AdamW) with `--batch` clips per GPU (cfg.TRAIN.BATCH_SIZE = 2, config.py:18)."""
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W_, K, B = max(3, args.warmup), args.steps, args.batch
    from mspi_b200 import _lib
    from mspi_b200.config import cfg as base_cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel
    from mspi_b200.train_engine import TrainPlan
    torch.manual_seed(2023)
    with contextlib.redirect_stdout(io.StringIO()):
        model = AudioVisualSaliencyModel(select_motion_encoder("s3d", copy.deepcopy(base_cfg)), load_pretrained=False)
    plan = TrainPlan(model.state_dict(), B, T, H, W, device=dev, world_size=world)
    g = torch.Generator(device=dev).manual_seed(2023 + rank)
    n_sets = 4
    clips = [torch.randn(B, 3, T, H, W, device=dev, generator=g) for _ in range(n_sets)]
    audio = [torch.randn(B, 1, 257, 111, device=dev, generator=g) for _ in range(n_sets)]
    gts = [torch.rand(B, H, W, device=dev, generator=g) for _ in range(n_sets)]
    from mspi_b200.distributed import allreduce_gradients
    allreduce = allreduce_gradients if world > 1 else None      # one NCCL SUM all-reduce of the flat gradient buffer
    lib = _lib.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    plan.train_step(clips[0], audio[0], gts[0], allreduce)   # builds the lazy weight-gradient plans
    torch.cuda.synchronize()
    l0 = lib.mspi_launch_count()
    plan.train_step(clips[1], audio[1], gts[1], allreduce)
    torch.cuda.synchronize()
    launches = int(lib.mspi_launch_count() - l0)
    if not args.no_graph:
        plan.capture_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for i in range(W_):
            plan.train_step(clips[i % n_sets], audio[i % n_sets], gts[i % n_sets], allreduce)
        barrier()
        clocks.mark()
        ev0.record()
        for i in range(K):
            out = plan.train_step(clips[i % n_sets], audio[i % n_sets], gts[i % n_sets], allreduce)
        ev1.record()
        barrier()
        clocks.mark_end()
        if rank == 0:   # forward+backward replays (no collective, no parameter update) until nvidia-smi has reported under load
            def _load():
                plan.forward_backward(clips[0], audio[0], gts[0])
                torch.cuda.synchronize()
            clocks.wait_for_sample(_load)
        barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * K / (ms / 1e3)
    loss_last = out.cpu().tolist()
    # ---- e2e: pinned host clips / audio / GT copied in every step, the 4 loss scalars copied back every step
    pin = [(torch.randn(B, 3, T, H, W).pin_memory(), torch.randn(B, 1, 257, 111).pin_memory(), torch.rand(B, H, W).pin_memory())
           for _ in range(2)]
    dbuf = [(torch.empty(B, 3, T, H, W, device=dev), torch.empty(B, 1, 257, 111, device=dev), torch.empty(B, H, W, device=dev))
            for _ in range(2)]
    host_loss = torch.empty(4).pin_memory()

    def e2e_run(n):
        for i in range(n):
            s = i % 2
            for d_, p_ in zip(dbuf[s], pin[s]):
                d_.copy_(p_, non_blocking=True)
            o = plan.train_step(*dbuf[s], allreduce)
            host_loss.copy_(o, non_blocking=True)
        torch.cuda.synchronize()

    e2e_run(2)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    e2e_run(K)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (t2.item() / 1e3)
    # ---- the step's only collective, timed alone: NCCL SUM all-reduce of the flat fp32 gradient buffer
    allreduce_ms = None
    if world > 1:
        for _ in range(2):
            allreduce(plan.flat_g)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(10):
            allreduce(plan.flat_g)
        a1.record()
        barrier()
        ta = torch.tensor([a0.elapsed_time(a1) / 10], device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_ms = ta.item()
    # ---- per-kernel breakdown of one step (CUDA events, eager replay), rank 0
    roofline, kinds = None, None
    if rank == 0:
        peaks = read_peaks()
        plan.bind(clips[0], audio[0])
        acc = {}
        phases = (("pack", [("pack(batched)", plan.pack)]), ("fwd", plan.steps), ("bwd", plan.bwd_steps))
        for rep in range(3):
            plan.flat_g.zero_()
            plan._bn_arena.zero_()
            evs = [torch.cuda.Event(enable_timing=True)]
            evs[0].record()
            names = []
            for ph, steps in phases:
                for name, fn in steps:
                    fn()
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    evs.append(e)
                    names.append((ph, name, fn))
            torch.cuda.synchronize()
            if rep:
                for j in range(len(names)):
                    acc.setdefault(j, []).append(evs[j].elapsed_time(evs[j + 1]))
        rows = []
        for j, (ph, name, fn) in enumerate(names):
            kind = "other"
            if name.startswith("conv_dgrad"):
                kind = "conv_gemm_tf32(dgrad)"
            elif name.endswith(".wgrad") and "dwconv" not in name:
                kind = "conv_wgrad"
            elif getattr(fn, "desc", None) is not None:
                kind = "conv_gemm_bf16" if fn.desc.a_dtype == 0 else "conv_gemm_tf32(fwd)"
            elif ph == "pack":
                kind = "pack"
            rows.append({"phase": ph, "step": name, "kind": kind, "ms": sum(acc[j]) / len(acc[j])})
        tot = sum(r["ms"] for r in rows)
        kinds = {}
        for r in rows:
            k = kinds.setdefault(r["kind"], {"launches": 0, "ms": 0.0})
            k["launches"] += 1
            k["ms"] += r["ms"]
        for k in kinds.values():
            k["share_of_step"] = k["ms"] / tot
        # algorithmic FLOPs: forward 416.05 GFLOP/clip; backward = 2x the trainable part (416.05 - 253.89 - 2.17), SURVEY 8(a) a20
        train_gflop_per_clip = ALGO_GFLOP_PER_CLIP + 2.0 * (ALGO_GFLOP_PER_CLIP - 253.89 - 2.17)
        tc_ms = sum(v["ms"] for k, v in kinds.items() if k.startswith("conv_"))
        tc_gflop = B * train_gflop_per_clip
        roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel<tf32|bf16> + wgrad_kernel<tf32> (tcgen05 implicit GEMMs: forward, "
                                                 "data gradient, weight gradient)",
                    "achieved": tc_gflop / tc_ms, "peak": peaks["bf16_sustained"] / 2, "unit": "TFLOP/s",
                    "frac": tc_gflop / tc_ms / (peaks["bf16_sustained"] / 2), "traffic": None,
                    "peak_source": f"{peaks['source']} bf16_tflops_sustained / 2 (tf32 issues at half the bf16 rate)",
                    "algorithmic_gflop_per_step": tc_gflop, "tensor_core_ms": tc_ms, "share_of_step": tc_ms / tot,
                    "note": "B=2 clips per GPU (the reference's batch size): the step is launch- and latency-bound, not tensor-bound"}
        if args.breakdown:
            with open(args.breakdown, "w") as f:
                json.dump({"rows": rows, "total_ms": tot, "kinds": kinds}, f, indent=1)
    if rank == 0:
        line = {"metric": "clips/sec MSPI-S3D training step", "value": value, "unit": "clips/s", "n_gpus": world, "steps": K,
                "warmup": W_, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32 (trainable part: fp32 storage, tf32 tensor cores) + bf16 (frozen encoders)", "data": "synthetic",
                "config": {"workload": f"MSPI-S3D training step (BASELINE config 5): train-mode forward, KLD - CC + SimSiam loss, backward "
                                       f"of the 411 trainable tensors (46.0 M parameters), "
                                       f"{'NCCL all-reduce of the flat fp32 gradient buffer, ' if world > 1 else ''}AdamW; "
                                       f"{B} clips/GPU/step of 16x{H}x{W} + spectrograms + GT maps, random init",
                           "batch_per_gpu": B, "parallelism": f"dp{world}", "cuda_graph": plan.graph is not None,
                           "l2": f"{n_sets} rotating input sets; activations + gradients of one step ({plan.bytes_alloc / 2**30:.1f} GiB) "
                                 "exceed L2"},
                "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": B * (3 * T * H * W + 257 * 111 + H * W) * 4,
                        "d2h_bytes_per_step": 16},
                "gpu_launches": launches * K, "launches_per_step": launches, "loss_last_step": loss_last,
                "trainable_parameters": plan.n_params, "grad_allreduce_bytes": plan.n_flat * 4 if world > 1 else 0,
                "grad_allreduce_ms": allreduce_ms,
                "grad_allreduce_busbw_gbs": (plan.n_flat * 4 * 2 * (world - 1) / world / allreduce_ms / 1e6) if allreduce_ms else None,
                "clocks": clocks.summary(), "roofline": roofline, "kernel_kinds": kinds,
                "activation_bytes_allocated": plan.bytes_alloc}
        line["cpu_baseline"] = train_cpu_baseline(B) if (world == 1 and not args.no_cpu_baseline) else None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--impl", default="mspi_b200", choices=["mspi_b200", "reference", "eager"])
    ap.add_argument("--encoder", default="s3d", choices=["s3d", "x3dl", "slowfast4x16"],
                    help="motion encoder (BASELINE configs: s3d = headline, x3dl = config 3, slowfast4x16 = config 4)")
    ap.add_argument("--train", action="store_true", help="BASELINE config 5: the training step (use --batch 2)")
    ap.add_argument("--no-graph", action="store_true", help="replay the kernel list eagerly instead of a CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-eager-baseline", action="store_true", help="skip the PyTorch-eager (cuDNN/cuBLAS) run on the same GPU")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle check of the timed batch")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-input end-to-end arms (multi-GPU sweeps of the other encoders)")
    ap.add_argument("--sliding", action="store_true", help="sliding-window workload with / without the per-frame feature cache")
    ap.add_argument("--frames", type=int, default=272, help="--sliding: frames of the synthetic video")
    ap.add_argument("--breakdown", default=None, help="write the per-kernel CUDA-event breakdown to this file")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)
    if args.impl == "eager":
        return run_eager(args, rank, world, local)
    if args.train:
        return run_train(args, rank, world, local)
    if args.sliding:
        return run_sliding(args, rank, world, local)

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product has no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W_ = max(3, args.warmup)
    K, B = args.steps, args.batch

    from mspi_b200 import _lib
    from mspi_b200.config import cfg as base_cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel
    torch.manual_seed(2023)
    algo_gflop = ALGO_GFLOP[args.encoder]
    with contextlib.redirect_stdout(io.StringIO()):
        model = AudioVisualSaliencyModel(select_motion_encoder(args.encoder, copy.deepcopy(base_cfg)),
                                         load_pretrained=False)  # random init, synthetic data
    model = model.to(dev).eval()
    model.use_cuda_graph = not args.no_graph

    g = torch.Generator(device=dev).manual_seed(2023 + rank)
    n_sets = 2  # rotate input sets; each set (B*16.5 MB) is already larger than... see config.l2
    clips_sets = [torch.randn(B, 3, T, H, W, device=dev, generator=g) for _ in range(n_sets)]
    audio_sets = [torch.randn(B, 1, 257, 111, device=dev, generator=g) for _ in range(n_sets)]
    from mspi_b200.distributed import gather_maps

    pending = [None]

    def step(i):
        out, loss = model(clips_sets[i % n_sets], audio_sets[i % n_sets])
        if world > 1:
            # NCCL all-gather of the [B,H,W] maps (global clip order) on the communicator's stream: it overlaps the next
            # step's kernels; the previous step's gather is waited for here, the last one by drain()
            h = gather_maps(out, world * B, async_op=True)
            if pending[0] is not None:
                pending[0].wait()
            pending[0] = h
        return out, loss

    def drain():
        if pending[0] is not None:
            pending[0].wait()
            pending[0] = None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    lib = _lib.load()
    l0 = lib.mspi_launch_count()
    step(0)  # builds the plan (and captures the graph)
    torch.cuda.synchronize()
    plan = next(iter(model._plans.values()))
    # launches per forward, counted on an eager replay of the plan's step list
    l0 = lib.mspi_launch_count()
    plan.bind(clips_sets[0], audio_sets[0])
    plan.run_eager()
    torch.cuda.synchronize()
    launches_per_fwd = int(lib.mspi_launch_count() - l0)
    # ------------------------------------------------------------------ timed region: device-resident inputs
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        for i in range(W_):
            step(i)
        barrier()
        clocks.mark()
        ev0.record()
        for i in range(K):
            step(i)
        drain()
        ev1.record()
        barrier()
        clocks.mark_end()
        if rank == 0:   # local replays of the same forward (no collectives) until nvidia-smi has reported under load
            def _load():
                plan.run(clips_sets[0], audio_sets[0])
                torch.cuda.synchronize()
            clocks.wait_for_sample(_load)
        barrier()
    ms = ev0.elapsed_time(ev1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    value = world * B * K / (ms / 1e3)

    # ------------------------------------------------------------------ parity of the timed configuration (untimed)
    parity = None
    if rank == 0 and not args.no_parity:
        out_chk, _ = model(clips_sets[0], audio_sets[0])      # the same graph replay the timed region ran
        torch.cuda.synchronize()
        parity = parity_at_bench_config(model, clips_sets[0], audio_sets[0], out_chk, args.encoder)

    # ------------------------------------------------------------------ e2e: host inputs through the public API
    # Two input formats of the same public forward(): (a) uint8 [B,T,H,W,3] frames — what a decoder produces and what the
    # reference's own pipeline starts from (inference.py:154-165 converts and normalises them on the host); the normalisation
    # runs on the device, fused into the clip-conversion kernel, bit-identical to the host-normalised clip (tests) — and
    # (b) the reference's literal contract, normalised fp32 [B,3,T,H,W] clips (4x the PCIe / host-memory bytes).
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    pin_a = [torch.randn(B, 1, 257, 111).pin_memory() for _ in range(2)]
    host_out = [torch.empty(B, H, W).pin_memory() for _ in range(2)]
    daud = [torch.empty(B, 1, 257, 111, device=dev) for _ in range(2)]

    e2e_pending = [None]

    def e2e_measure(pin, dclips):
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def upload(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[s])
                dclips[s].copy_(pin[s], non_blocking=True)
                daud[s].copy_(pin_a[s], non_blocking=True)
                ready[s].record(copy_stream)

        def e2e_run(n):
            for s in range(2):
                freed[s].record(main_stream)
            upload(0)
            for i in range(n):
                s = i % 2
                if i + 1 < n:
                    upload(i + 1)  # overlaps the H2D of the next step's inputs with this step's kernels
                main_stream.wait_event(ready[s])
                out, loss = model(dclips[s], daud[s])
                freed[s].record(main_stream)
                if world > 1:   # NCCL all-gather of the maps on the communicator's stream, as in the device-resident loop
                    h = gather_maps(out, world * B, async_op=True)
                    if e2e_pending[0] is not None:
                        e2e_pending[0].wait()
                    e2e_pending[0] = h
                host_out[s].copy_(out, non_blocking=True)  # D2H of this step's maps
            if e2e_pending[0] is not None:
                e2e_pending[0].wait()
                e2e_pending[0] = None
            torch.cuda.synchronize()

        e2e_run(2)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        e2e_run(K)
        e1.record()
        barrier()
        t2 = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        return world * B * K / (t2.item() / 1e3)

    h2d = B * (3 * T * H * W + 257 * 111 * 4)
    h2d_fp32 = B * (3 * T * H * W + 257 * 111) * 4
    d2h = B * H * W * 4
    e2e_value = e2e_fp32 = None
    if not args.no_e2e:
        pin_u8 = [torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8).pin_memory() for _ in range(2)]
        d_u8 = [torch.empty(B, T, H, W, 3, dtype=torch.uint8, device=dev) for _ in range(2)]
        e2e_value = e2e_measure(pin_u8, d_u8)
        del pin_u8, d_u8
        model._plans = {k: v for k, v in model._plans.items() if not (k[0] == "full" and k[1] is True)}   # drop the uint8 plan's 30 GB of buffers
        torch.cuda.empty_cache()
        pin = [torch.randn(B, 3, T, H, W).pin_memory() for _ in range(2)]
        dclips = [torch.empty(B, 3, T, H, W, device=dev) for _ in range(2)]
        e2e_fp32 = e2e_measure(pin, dclips)
        del pin, dclips

    # ------------------------------------------------------------------ per-kernel breakdown (CUDA events, eager replay)
    roofline, tf32_info, breakdown = None, None, None
    if rank == 0:
        peaks = read_peaks()
        plan.bind(clips_sets[0], audio_sets[0])
        reps = 2
        acc = {}
        for rep in range(reps + 1):
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(plan.steps) + 1)]
            evs[0].record()
            for j, (_name, fn) in enumerate(plan.steps):
                fn()
                evs[j + 1].record()
            torch.cuda.synchronize()
            if rep == 0:
                continue  # warm
            for j, (name, fn) in enumerate(plan.steps):
                acc.setdefault(j, []).append(evs[j].elapsed_time(evs[j + 1]))
        rows = []
        for j, (name, fn) in enumerate(plan.steps):
            d = getattr(fn, "desc", None)
            kind = "other"
            flops = 0.0
            if d is not None:
                kind = "conv_gemm_bf16" if d.a_dtype == 0 else "conv_gemm_tf32"
                flops = getattr(fn, "flops", 0.0)
            rows.append({"step": name, "kind": kind, "ms": sum(acc[j]) / len(acc[j]), "gflop": flops / 1e9})
        tot = sum(r["ms"] for r in rows)
        for kind in ("conv_gemm_bf16", "conv_gemm_tf32"):
            sel = [r for r in rows if r["kind"] == kind]
            ms_k, gf_k = sum(r["ms"] for r in sel), sum(r["gflop"] for r in sel)
            info = {"launches": len(sel), "ms": ms_k, "share_of_step": ms_k / tot if tot else None,
                    "gflop": gf_k, "tflops": gf_k / ms_k if ms_k else None}
            if kind == "conv_gemm_bf16":
                peak = peaks["bf16_sustained"]
                tr = read_traffic()
                traffic = tr["conv_gemm_bf16"]["dram_bytes_per_launch"] if tr else None
                roofline = {"bound": "tensor", "kernel": "conv_gemm_kernel<bf16> (tcgen05 implicit GEMM)",
                            "achieved": info["tflops"], "peak": peak, "unit": "TFLOP/s",
                            "frac": (info["tflops"] / peak) if info["tflops"] else None, "traffic": traffic,
                            "traffic_note": f"dram read+write bytes per launch, mean over the {tr['conv_gemm_bf16']['launches_per_step']} "
                                            "bf16 launches of one step (ncu, profiles/r02_launches.csv)" if tr else None,
                            "hbm_gbs_from_traffic": (tr["conv_gemm_bf16"]["dram_bytes_per_step"] / ms_k / 1e6) if tr else None,
                            "hbm_peak_gbs": peaks["hbm"],
                            "peak_source": f"{peaks['source']} bf16_tflops_sustained (kernel timed inside a long step)",
                            "launches_per_step": info["launches"], "share_of_step": info["share_of_step"],
                            "algorithmic_gflop_per_step": gf_k}
            else:
                tf32_info = info
        breakdown = {"total_ms_eager_sum": tot, "rows": sorted(rows, key=lambda r: -r["ms"])[:40]}
        if args.breakdown:
            with open(args.breakdown, "w") as f:
                json.dump({"rows": rows, "total_ms": tot}, f, indent=1)

    if rank == 0:
        peaks = read_peaks()
        line = {
            "metric": f"clips/sec {ENC_NAME[args.encoder]} inference", "value": value, "unit": "clips/s", "n_gpus": world, "steps": K,
            "warmup": W_, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{ENC_NAME[args.encoder]} ({args.encoder} motion encoder + ResNet18 audio + ConvNeXt-T image encoder + fusion decoder) inference, "
                                   f"{B} clips/GPU/step of 16x{H}x{W} fp32 + [1,257,111] spectrograms, random init",
                       "batch_per_gpu": B, "clip": [3, T, H, W], "parallelism": f"dp{world} (clip sharding, NCCL all-gather of maps)",
                       "precision": "bf16 tensor cores (encoders), tf32 tensor cores + fp32 storage (fusion/decoder), fp32 accumulate",
                       "cuda_graph": bool(model.use_cuda_graph),
                       "l2": f"inputs larger than L2: {B * 3 * T * H * W * 4 / 2**20:.0f} MiB of clips per step, {n_sets} rotating sets"},
            "tensor_frac_of_peak_whole_step": value / world * algo_gflop / 1e3 / peaks["bf16_sustained"],
            "algorithmic_gflop_per_clip": algo_gflop,
            "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "input": "uint8 [B,T,H,W,3] frames + fp32 spectrograms in pinned host memory",
                    "note": "model(frames_u8, spectrograms): H2D of step i+1 overlapped with step i on a copy stream, /255 + ImageNet "
                            "normalisation fused into the clip-conversion kernel (bit-identical to the host-normalised clip), "
                            "D2H of the [B,H,W] maps every step"},
            "e2e_fp32_clips": {"value": e2e_fp32, "unit": "clips/s", "h2d_bytes_per_step": h2d_fp32, "d2h_bytes_per_step": d2h,
                               "input": "host-normalised fp32 [B,3,T,H,W] clips (the reference forward's literal contract)"},
            "gpu_launches": launches_per_fwd * K,
            "launches_per_step": launches_per_fwd,
            "clocks": clocks.summary(),
            "roofline": roofline,
            "tf32_kernel": tf32_info,
            "activation_bytes_allocated": plan.bytes_alloc,
            "parity_at_bench_config": parity,
        }
        if world == 1 and not args.no_eager_baseline:
            with ClockSampler(local) as eclocks:
                eclocks.mark()
                line["gpu_eager_baseline"] = gpu_eager_baseline(dev, B, encoder=args.encoder)
            line["gpu_eager_baseline"]["clocks"] = eclocks.summary()
            best = max((m["value"] for m in line["gpu_eager_baseline"].values() if isinstance(m, dict) and "value" in m),
                       default=None)
            line["vs_gpu_eager"] = (value / best) if best else None
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        else:
            line["cpu_baseline"] = None
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
