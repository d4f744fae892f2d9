"""Forward plan of the MSPI clip forward pass on B200.

A :class:`ForwardPlan` is built once per (batch, frames, height, width): it allocates every
activation buffer (channels-last bf16, fp32 where the result feeds a normalisation or the final
log-softmax), packs the weights for the tcgen05 implicit-GEMM kernel (BatchNorm folded into the
epilogue's scale/shift), and records the forward as a flat list of C-ABI kernel launches.
``run()`` replays the list on the current CUDA stream — eagerly or, with ``use_graph=True``,
as one captured CUDA graph.

The op order follows the reference forward (model/model_utils.py:556-574) and each builder
cites the reference lines it re-implements.  Nothing here calls a PyTorch operator on the data
path; PyTorch only owns the memory.
"""
from __future__ import annotations

import math
import os
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch

from . import ops
from .ops import ACT_GELU, ACT_NONE, ACT_RELU, Act, Conv, fold_bn


class ForwardPlan:
    def __init__(self, sd: Dict[str, torch.Tensor], batch: int, frames: int, height: int, width: int,
                 audio: bool = True, lateral_bool=(True, True, False, False), lateral_stride=(2, 2, 2, 2),
                 pool_stride: int = 1, device="cuda", keep_taps: bool = False, encoder: str = "s3d",
                 pack_on_host: bool = True, weight_cache: Optional[dict] = None, mode: str = "full", input_u8: bool = False,
                 liveness: Optional[dict] = None):
        assert (frames % 4 == 0 or mode == "image_encoder") and height % 32 == 0 and width % 32 == 0, \
            "T must be a multiple of 4 and H, W multiples of 32 (model_utils.py:506,566-570)"
        # Inference plans fold BatchNorm and pack the weights on the HOST (plain fp32 tensor arithmetic) and upload the packed
        # matrices: building a plan enqueues no kernels, so the first launches a profiler sees are the product's own, and
        # `weight_cache` (owned by the nn.Module, dropped when its parameters change) lets plans for other batch shapes
        # reuse the uploaded matrices.  The training plan keeps live device tensors (it re-packs them every step).
        self.sd = {k: (v.detach().cpu() if pack_on_host else v.detach().to(device)) for k, v in sd.items()}
        self.wcache = weight_cache if pack_on_host else None
        self.B, self.T, self.H, self.W = batch, frames, height, width
        self.audio = audio
        self.device = device
        self.lateral_bool, self.lateral_stride, self.pool_stride = lateral_bool, lateral_stride, pool_stride
        if encoder not in ("s3d", "x3dl", "slowfast4x16"):
            raise Exception("Invalid Motion Encoder!")  # get_video_backbones.py:28-29
        self.encoder = encoder
        # mode "full": the reference forward.  "image_encoder": only the per-frame ConvNeXt-T + smooth convs over `batch` frames
        # (frames = 1).  "cached": the forward with the image-encoder features taken from a per-frame cache through a
        # [B*T] frame index (sliding windows share 15 of 16 frames: inference.py:120-150).
        # input_u8: clips arrive as uint8 [B,T,H,W,3] frames and are normalised on the device (inference.py:154-165).
        assert mode in ("full", "image_encoder", "cached")
        self.mode, self.input_u8 = mode, input_u8
        self.fuse_mlp = os.environ.get("MSPI_FUSE_MLP", "1") != "0"
        # downsample.0 LayerNorm inside the last fused MLP of stages 0 / 1 (mspi_mlp_fused_ln): measured SLOWER than the
        # separate LayerNorm kernel (C = 96: +0.29 ms on the MLP launch against a 0.245 ms LayerNorm; C = 192: +0.14 against
        # 0.13) — the fused kernel is bound by its epilogue warps' instruction stream, the LayerNorm kernel by HBM.  Off.
        self.fuse_ds_norm = os.environ.get("MSPI_FUSE_DS_NORM", "0") != "0"
        self.fuse_stem_norm = os.environ.get("MSPI_FUSE_STEM_NORM", "1") != "0"
        self.fuse_mixed = os.environ.get("MSPI_FUSE_MIXED", "1") != "0"
        self.steps: List[Tuple[str, Callable[[], None]]] = []
        # Branch schedule (see _build / run_branched): `sched` interleaves ("step", index) entries with ("edge", src, dst)
        # entries meaning "branch dst waits for everything branch src has enqueued so far".  The list order of `steps` stays a
        # valid serial order, which is what run_eager(), the per-kernel breakdown and the training plan use.
        self.sched: List[Tuple] = []
        # activation memory: every self.new() is recorded; with `liveness` (from compute_liveness() of a probe build of the
        # same plan) the memory of dead buffers is handed to later allocations of the same graph branch (arena.py)
        from .arena import Arena
        self._allocs: List[Tuple] = []
        self._arena = Arena(liveness, device) if liveness is not None else None
        self.input_steps = set()
        self.step_branch: List[int] = []
        self._branch = 0
        self.branches = os.environ.get("MSPI_GRAPH_BRANCHES", "1") != "0"
        self.flops = 0.0
        self.bytes_alloc = 0
        self.taps: Dict[str, Act] = {}
        self.keep_taps = keep_taps
        self._keep = []
        self._inputs = {"clips": None, "audio": None, "feat_o1": None, "feat_o0": None}
        self.graph = None
        self._build()

    # ------------------------------------------------------------------ small helpers
    def P(self, name: str) -> torch.Tensor:
        return self.sd[name]

    def new(self, n, t, h, w, c, dtype=torch.bfloat16) -> Act:
        k, step_now = len(self._allocs), len(self.steps)
        if self._arena is not None:
            before = self._arena.fresh_bytes
            a = Act(self._arena.alloc(k, (n, t, h, w, c), dtype, step_now, self._branch))
            self.bytes_alloc += self._arena.fresh_bytes - before
        else:
            a = Act.empty(n, t, h, w, c, dtype=dtype, device=self.device)
            self.bytes_alloc += a.buf.numel() * a.buf.element_size()
        self._allocs.append((k, a.buf, step_now, self._branch))
        return a

    def compute_liveness(self) -> dict:
        """{allocation index: (branch, last step touching it)} for the buffers whose memory may be re-used (arena.py)."""
        from .arena import compute_liveness
        pinned = [getattr(self, "logits", None), getattr(self, "out", None)]
        pinned += [a.buf for a in self.taps.values()]
        for nm in ("feat_o1", "feat_o0"):
            if getattr(self, nm, None) is not None:
                pinned.append(getattr(self, nm).buf)
        live = compute_liveness(self._allocs, self.steps, self.step_branch, [t for t in pinned if t is not None])
        live["__count__"] = len(self._allocs)
        return live

    def add(self, name: str, fn: Callable[[], None], reads_input: bool = False):
        """reads_input: the step reads the caller's clips / spectrograms.  Those steps stay OUTSIDE the captured graph and run
        eagerly on the caller's tensors before every replay, so a replay needs no copy of the inputs into static buffers."""
        if reads_input:
            self.input_steps.add(len(self.steps))
        self.sched.append(("step", len(self.steps)))
        self.step_branch.append(self._branch)
        self.steps.append((name, fn))

    def on_branch(self, b: int):
        """Steps added from now on belong to branch b (0 = the main stream)."""
        self._branch = b

    def edge(self, src: int, dst: int):
        """Branch dst waits for the work branch src has been given so far."""
        if src != dst:
            self.sched.append(("edge", src, dst))

    def tap(self, name: str, a: Act):
        if self.keep_taps:
            self.taps[name] = a

    def bn(self, prefix: str, eps: float, conv_bias: Optional[torch.Tensor] = None):
        return fold_bn(self.P(prefix + ".weight"), self.P(prefix + ".bias"), self.P(prefix + ".running_mean"),
                       self.P(prefix + ".running_var"), eps, conv_bias)

    def conv(self, name: str, x: Act, w: torch.Tensor, scale=None, shift=None, stride=(1, 1, 1), pad=(0, 0, 0),
             act=ACT_NONE, out: Optional[Act] = None, residual: Optional[Act] = None, res_after_act=False,
             out_dtype=torch.bfloat16, dtype=torch.bfloat16, split_weights=False) -> Act:
        if act in (ACT_NONE, ACT_RELU) and not split_weights and not self.keep_taps and \
                ops.conv133_small_ok(w, stride, pad, x, dtype, out_dtype, residual):
            # SlowFast fast pathway, dim_inner 8 / 16: nine K = 8 taps are no work for a tensor-core tile (4 TFLOP/s)
            if out is None:
                out = self.new(x.n, x.t, x.h, x.w, w.shape[0], out_dtype)
            run = ops.conv133_small(x, out, w, scale, shift, act)
            self.add(name, run)
            self.flops += run.flops
            return out
        c = Conv(w, scale, shift, stride=stride, pad=pad, act=act, dtype=dtype, res_after_act=res_after_act,
                 device=self.device, name=name, split_weights=split_weights, cache=getattr(self, "wcache", None))
        if out is None:
            ot, oh, ow = c.out_shape(x.t, x.h, x.w)
            out = self.new(x.n, ot, oh, ow, c.cout, out_dtype)
        self.add(name, c.plan(x, out, residual))
        self.flops += c.flops(x)
        self._keep.append(c)
        return out

    def conv_bn_relu(self, conv_key: str, bn_key: str, x: Act, eps: float, stride=(1, 1, 1), pad=(0, 0, 0),
                     out: Optional[Act] = None, bias_key: Optional[str] = None, relu=True) -> Act:
        bias = self.P(bias_key) if bias_key else None
        sc, sh = self.bn(bn_key, eps, bias)
        return self.conv(conv_key, x, self.P(conv_key + ".weight"), sc, sh, stride, pad,
                         ACT_RELU if relu else ACT_NONE, out)

    def pool(self, name: str, x: Act, k, s, p, out: Optional[Act] = None) -> Act:
        o = [(d + 2 * pp - kk) // ss + 1 for d, kk, ss, pp in zip((x.t, x.h, x.w), k, s, p)]
        if out is None:
            out = self.new(x.n, *o, x.c)
        self.add(name, ops.maxpool3d(x, out, k, s, p))
        return out

    def up(self, name: str, x: Act, k: int, out: Optional[Act] = None, accumulate=False, dtype=None,
           act=ACT_NONE) -> Act:
        if out is None:
            out = self.new(x.n, x.t, x.h * k, x.w * k, x.c, dtype or x.dtype)
        self.add(name, ops.upsample(x, out, k, accumulate, act))
        return out

    def rows_act(self, t: torch.Tensor, c: Optional[int] = None) -> Act:
        """View a [M, C] row-major tensor as an Act with N=T=H=1, W=M (flat GEMM operand)."""
        m, cs = t.shape
        return Act(t.view(1, 1, 1, m, cs), 0, cs if c is None else c)

    def flat(self, a: Act) -> Act:
        return Act(a.buf.view(1, 1, 1, a.pixels, a.cs), a.c0, a.c)

    def linear(self, name: str, x: Act, w_key: str, b_key: Optional[str], act=ACT_NONE, out: Optional[Act] = None,
               residual: Optional[Act] = None, out_dtype=torch.bfloat16, scale=None) -> Act:
        w = self.P(w_key)
        shift = self.P(b_key) if b_key else None
        if scale is not None and shift is not None:
            shift = shift * scale
        return self.conv(name, x, w, scale, shift, act=act, out=out, residual=residual,
                         res_after_act=residual is not None, out_dtype=out_dtype)

    # ------------------------------------------------------------------ stems reading the fp32 NCDHW inputs
    def padded_frames(self) -> torch.Tensor:
        """The clip as bf16 zero-padded 4-channel frames, converted once per forward and shared by both video stems."""
        if getattr(self, "_frames", None) is None:
            B, T, H, W = self.B, self.T, self.H, self.W
            self._frames = torch.zeros((B * T, H + ops.PAD_EXTRA, W + ops.PAD_EXTRA, 4), dtype=torch.bfloat16,
                                       device=self.device)
            self.bytes_alloc += self._frames.numel() * 2
            conv = ops.clip_u8_to_padded if self.input_u8 else ops.clip_to_padded
            self.add("clips.to_padded_nhwc4", conv(self._inputs, "clips", self._frames, B, T, H, W), reads_input=True)
        return self._frames

    def padded_frames_variant(self, tag: str, frame_map, t_pad: int) -> torch.Tensor:
        """Frame-selected / time-padded copies of the clip for the SlowFast stems (ops.clip_to_padded)."""
        B, T, H, W = self.B, self.T, self.H, self.W
        t_out = len(frame_map) if frame_map is not None else T
        fr = torch.zeros((B * (t_out + 2 * t_pad), H + ops.PAD_EXTRA, W + ops.PAD_EXTRA, 4), dtype=torch.bfloat16,
                         device=self.device)
        self.bytes_alloc += fr.numel() * 2
        if self.input_u8:
            self.add(f"clips.to_padded_nhwc4[{tag}]", ops.clip_u8_to_padded(self._inputs, "clips", fr, B, T, H, W,
                                                                           frame_map=frame_map, t_pad=t_pad), reads_input=True)
        else:
            self.add(f"clips.to_padded_nhwc4[{tag}]", ops.clip_to_padded(self._inputs, "clips", fr, B, T, H, W, frame_map, t_pad),
                     reads_input=True)
        return fr

    def stem_direct(self, name: str, w: torch.Tensor, scale, shift, k: int, stride: int, pad: int, act,
                    out: Act, frames: Optional[torch.Tensor] = None, clips: int = 0, ln=None) -> Act:
        fr = self.padded_frames() if frames is None else frames
        run = ops.stem_conv(fr, self.H, self.W, w, scale, shift, k, stride, pad, act, out, name, clips=clips, ln=ln)
        self.add(name, run)
        self.flops += run.flops
        return out

    def stem_gemm(self, name: str, which: str, shape5, w: torch.Tensor, scale, shift, kernel, stride, pad, act,
                  out_dtype=torch.bfloat16) -> Act:
        """Patch-gather (fp32 NCDHW -> bf16 rows) + flat GEMM for the Cin<=3 stems."""
        n, c, t, h, wd = shape5
        cout = w.shape[0]
        k = c * kernel[0] * kernel[1] * kernel[2]
        k_pad = -(-k // 64) * 64
        ot = (t + 2 * pad[0] - kernel[0]) // stride[0] + 1
        oh = (h + 2 * pad[1] - kernel[1]) // stride[1] + 1
        ow = (wd + 2 * pad[2] - kernel[2]) // stride[2] + 1
        m = n * ot * oh * ow
        patches = torch.empty((m, k_pad), dtype=torch.bfloat16, device=self.device)
        self.bytes_alloc += patches.numel() * 2
        lib = ops._lib.load()
        pd = ops.PatchDesc()
        pd.src_layout = 0
        pd.n, pd.c, pd.t, pd.h, pd.w = n, c, t, h, wd
        pd.kt, pd.kh, pd.kw = kernel
        pd.st, pd.sh, pd.sw = stride
        pd.pt, pd.ph, pd.pw = pad
        pd.ot, pd.oh, pd.ow = ot, oh, ow
        pd.k_pad = k_pad
        holder = self._inputs
        import ctypes as C

        def gather(_keep=(patches, pd)):
            src = holder[which]
            ops._lib.check(lib.mspi_patch_gather(C.byref(pd), ops._ptr(src), ops._ptr(patches), ops._stream()),
                           f"patch_gather[{name}]")

        self.add(name + ".gather", gather, reads_input=True)
        w2 = torch.zeros((cout, k_pad), dtype=torch.float32, device=w.device)
        w5 = w if w.dim() == 5 else w[:, :, None]
        w2[:, :k] = w5.permute(0, 2, 3, 4, 1).reshape(cout, k)
        out = self.new(n, ot, oh, ow, cout, out_dtype)
        self.conv(name, self.rows_act(patches), w2, scale, shift, act=act, out=self.flat(out))
        self.flops -= 2.0 * m * cout * (k_pad - k)  # count algorithmic flops only
        return out

    # ------------------------------------------------------------------ S3D  (backbones/s3d.py)
    def sep(self, p: str, x: Act, k: int, stride: int, pad: int, out: Optional[Act] = None) -> Act:
        """SepConv3d, s3d.py:95-116"""
        y = self.conv_bn_relu(p + ".conv_s", p + ".bn_s", x, 1e-3, (1, stride, stride), (0, pad, pad))
        return self.conv_bn_relu(p + ".conv_t", p + ".bn_t", y, 1e-3, (stride, 1, 1), (pad, 0, 0), out)

    def basic(self, p: str, x: Act, out: Optional[Act] = None, stride=(1, 1, 1), pad=(0, 0, 0)) -> Act:
        """BasicConv3d, s3d.py:41-52"""
        return self.conv_bn_relu(p + ".conv", p + ".bn", x, 1e-3, stride, pad, out)

    def mixed(self, p: str, x: Act, out: Optional[Act] = None) -> Act:
        """Mixed_* / Inception block: four branches written straight into their slices of the
        concatenated output.  s3d.py:118-376, model_utils.py:173-199

        The three branch-entry 1x1x1 convs (branch0.0, branch1.0, branch2.0: same input, BN + ReLU each) run as ONE GEMM
        with their weights stacked along N.  Its output columns [t1 | t2 | b0] must be contiguous channels, so the block's
        buffer is laid out [t1 | t2 | b0 | b1 | b2 | b3] and the block output is the channel slice behind the two
        temporaries (every consumer addresses activations as channel slices anyway).  When the caller dictates the output
        buffer (the fp32 v4 slice of the decoder input), t1 and t2 share one GEMM and branch0 keeps its own."""
        c0 = self.P(p + ".branch0.0.conv.weight").shape[0]
        c1 = self.P(p + ".branch1.1.conv_t.weight").shape[0]
        c2 = self.P(p + ".branch2.1.conv_t.weight").shape[0]
        c3 = self.P(p + ".branch3.1.conv.weight").shape[0]
        c1a = self.P(p + ".branch1.0.conv.weight").shape[0]
        c2a = self.P(p + ".branch2.0.conv.weight").shape[0]
        fuse = self.fuse_mixed and c1a % 8 == 0 and c2a % 8 == 0 and c0 % 8 == 0

        def entry(branches, dst: Act, name: str):
            ws = [self.P(f"{p}.{b}.conv.weight") for b in branches]
            folded = [self.bn(f"{p}.{b}.bn", 1e-3) for b in branches]
            self.conv(name, x, torch.cat(ws, 0), torch.cat([f[0] for f in folded]), torch.cat([f[1] for f in folded]),
                      act=ACT_RELU, out=dst)

        if not fuse:
            if out is None:
                out = self.new(x.n, x.t, x.h, x.w, c0 + c1 + c2 + c3)
            self.basic(p + ".branch0.0", x, out.slice(0, c0))
            t1 = self.basic(p + ".branch1.0", x)
            t2 = self.basic(p + ".branch2.0", x)
        elif out is None:
            wide = self.new(x.n, x.t, x.h, x.w, c1a + c2a + c0 + c1 + c2 + c3)
            out = wide.slice(c1a + c2a, c0 + c1 + c2 + c3)
            entry(("branch1.0", "branch2.0", "branch0.0"), wide.slice(0, c1a + c2a + c0), p + ".branch{1.0,2.0,0.0}")
            t1, t2 = wide.slice(0, c1a), wide.slice(c1a, c2a)
        else:
            self.basic(p + ".branch0.0", x, out.slice(0, c0))
            t12 = self.new(x.n, x.t, x.h, x.w, c1a + c2a)
            entry(("branch1.0", "branch2.0"), t12, p + ".branch{1.0,2.0}")
            t1, t2 = t12.slice(0, c1a), t12.slice(c1a, c2a)
        self.sep(p + ".branch1.1", t1, 3, 1, 1, out.slice(c0, c1))
        self.sep(p + ".branch2.1", t2, 3, 1, 1, out.slice(c0 + c1, c2))
        pooled = self.pool(p + ".branch3.0", x, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        self.basic(p + ".branch3.1", pooled, out.slice(c0 + c1 + c2, c3))
        return out

    def s3d(self, v4_out: Optional[Act]):
        """S3D_features_only.forward, s3d.py:406-418"""
        p = "visnet."
        B, T, H, W = self.B, self.T, self.H, self.W
        sc, sh = self.bn(p + "base1.0.bn_s", 1e-3)
        x = self.stem_direct(p + "base1.0.conv_s", self.P(p + "base1.0.conv_s.weight"), sc, sh, 7, 2, 3, ACT_RELU,
                             self.new(B, T, H // 2, W // 2, 64))
        x = self.conv_bn_relu(p + "base1.0.conv_t", p + "base1.0.bn_t", x, 1e-3, (2, 1, 1), (3, 0, 0))
        x = self.pool(p + "base1.1", x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        x = self.basic(p + "base1.2", x)
        v1 = self.sep(p + "base1.3", x, 3, 1, 1)
        x = self.pool(p + "maxpooling2", v1, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        x = self.mixed(p + "base2.0", x)
        v2 = self.mixed(p + "base2.1", x)
        x = self.pool(p + "maxpooling3", v2, (3, 3, 3), (2, 2, 2), (1, 1, 1))
        for i in range(5):
            x = self.mixed(p + f"base3.{i}", x)
        v3 = x
        ps = self.pool_stride
        x = self.pool(p + "maxpooling4", v3, (ps, 2, 2), (ps, 2, 2), (0, 0, 0))
        x = self.mixed(p + "base4.0", x)
        v4 = self.mixed(p + "base4.1", x, v4_out)
        for i, v in enumerate((v1, v2, v3, v4)):
            self.tap(f"visnet.base{i + 1}", v)
        return v1, v2, v3, v4

    # ------------------------------------------------------------------ PySlowFast-style ResNets (X3D-L, SlowFast)
    @staticmethod
    def _pad_rows(t: torch.Tensor, n: int) -> torch.Tensor:
        if t.shape[0] == n:
            return t
        out = torch.zeros((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        out[: t.shape[0]] = t
        return out

    @staticmethod
    def _pad_cin(w: torch.Tensor, n: int) -> torch.Tensor:
        if w.shape[1] == n:
            return w
        out = torch.zeros((w.shape[0], n) + tuple(w.shape[2:]), dtype=w.dtype, device=w.device)
        out[:, : w.shape[1]] = w
        return out

    def res_shortcut(self, q: str, x: Act, stride: int) -> Act:
        """ResBlock projection shortcut: 1x1x1 conv with spatial stride + BN (resnet_helper.py:556-570, 586-587)."""
        if (q + ".branch1.weight") not in self.sd:
            return x
        sc, sh = self.bn(q + ".branch1_bn", 1e-5)
        return self.conv(q + ".branch1", x, self.P(q + ".branch1.weight"), sc, sh, stride=(1, stride, stride))

    def x3d_block(self, q: str, x: Act, stride: int, out: Optional[Act] = None) -> Act:
        """ResBlock(X3DTransform): a 1x1x1+BN+ReLU -> b depthwise 3x3x3 (stride s)+BN -> [SE] -> Swish -> c 1x1x1+BN;
        relu(shortcut + .).  resnet_helper.py:213-351, 580-590.  The inner width (54/108) is padded to a multiple of 8
        with zero weights so every tensor keeps 16-byte channel vectors."""
        inner = self.P(q + ".branch2.a.weight").shape[0]
        ip = -(-inner // 8) * 8
        sc, sh = self.bn(q + ".branch2.a_bn", 1e-5)
        a = self.conv(q + ".branch2.a", x, self._pad_rows(self.P(q + ".branch2.a.weight"), ip), self._pad_rows(sc, ip),
                      self._pad_rows(sh, ip), act=ACT_RELU)
        has_se = (q + ".branch2.se.fc1.weight") in self.sd
        sc, sh = self.bn(q + ".branch2.b_bn", 1e-5)
        b = self.new(x.n, x.t, x.h // stride, x.w // stride, ip)
        # SE squeeze (mean over t, h, w of b) accumulated by the depthwise kernel itself: one pass over b less per SE block
        mean = None
        if has_se and os.environ.get("MSPI_SE_MEAN_FUSED", "1") != "0":
            mean = torch.empty((x.n, ip), dtype=torch.float32, device=self.device)
        self.add(q + (".branch2.b+mean" if mean is not None else ".branch2.b"),
                 ops.dwconv3d_bn(a, b, self.P(q + ".branch2.b.weight"), sc, sh, stride,
                                 ops.ACT_NONE if has_se else ops.ACT_SWISH, mean=mean))
        self.flops += 2.0 * b.pixels * inner * 27
        if has_se:
            se = q + ".branch2.se."
            fns = ops.se_block(b, self.P(se + "fc1.weight"), self.P(se + "fc1.bias"), self.P(se + "fc2.weight"),
                               self.P(se + "fc2.bias"), mean=mean)
            for nm, fn in zip(("mean", "fc", "scale+swish")[3 - len(fns):], fns):
                self.add(se + nm, fn)
        short = self.res_shortcut(q, x, stride)
        sc, sh = self.bn(q + ".branch2.c_bn", 1e-5)
        return self.conv(q + ".branch2.c", b, self._pad_cin(self.P(q + ".branch2.c.weight"), ip), sc, sh, act=ACT_RELU,
                         out=out, residual=short, out_dtype=out.dtype if out is not None else torch.bfloat16)

    def x3d(self, v4_out: Optional[Act]):
        """X3D(features_only).forward, backbones/X3D.py:232-246; stem: stem_helper.py:207-290."""
        from .backbones.X3D import DEPTHS
        p = "visnet."
        B, T, H, W = self.B, self.T, self.H, self.W
        q = p + "s1.pathway0_stem"
        xy = self.stem_direct(q + ".conv_xy", self.P(q + ".conv_xy.weight"), None, None, 3, 2, 1, ACT_NONE,
                              self.new(B, T, H // 2, W // 2, 24))
        sc, sh = self.bn(q + ".bn", 1e-5)
        x = self.new(B, T, H // 2, W // 2, 24)
        self.add(q + ".conv+bn", ops.dwconv3d_bn(xy, x, self.P(q + ".conv.weight"), sc, sh, 1, ACT_RELU))
        self.flops += 2.0 * x.pixels * 24 * 5
        feats = []
        for si, depth in enumerate(DEPTHS):
            for i in range(depth):
                last = si == 3 and i == depth - 1
                x = self.x3d_block(f"{p}s{si + 2}.pathway0_res{i}", x, 2 if i == 0 else 1, v4_out if last else None)
            feats.append(x)
            self.tap(f"visnet.base{si + 1}", x)
        return feats

    def bottleneck_block(self, q: str, x: Act, stride: int, tk: int, out: Optional[Act] = None) -> Act:
        """ResBlock(BottleneckTransform): (tk,1,1)+BN+ReLU -> (1,3,3)/s+BN+ReLU -> 1x1x1+BN; relu(shortcut + .).
        resnet_helper.py:354-487, 580-590"""
        sc, sh = self.bn(q + ".branch2.a_bn", 1e-5)
        a = self.conv(q + ".branch2.a", x, self.P(q + ".branch2.a.weight"), sc, sh, pad=(tk // 2, 0, 0), act=ACT_RELU)
        sc, sh = self.bn(q + ".branch2.b_bn", 1e-5)
        b = self.conv(q + ".branch2.b", a, self.P(q + ".branch2.b.weight"), sc, sh, stride=(1, stride, stride), pad=(0, 1, 1),
                      act=ACT_RELU)
        short = self.res_shortcut(q, x, stride)
        sc, sh = self.bn(q + ".branch2.c_bn", 1e-5)
        return self.conv(q + ".branch2.c", b, self.P(q + ".branch2.c.weight"), sc, sh, act=ACT_RELU, out=out, residual=short,
                         out_dtype=out.dtype if out is not None else torch.bfloat16)

    def slowfast(self, v4_out: Optional[Act]):
        """SlowFast.forward (backbones/sf.py:364-385) on [slow = frames 0,4,12,T-1 ; fast = all frames]
        (model_utils.py:521-524)."""
        from .backbones.sf import DEPTHS, SLOW_FRAMES, TK_FAST, TK_SLOW
        p = "visnet."
        B, T, H, W = self.B, self.T, self.H, self.W
        ts = len(SLOW_FRAMES)
        # stems: (k,7,7)/s(1,2,2) + BN + ReLU + MaxPool(1,3,3)/s(1,2,2)   stem_helper.py:128-204
        fr_s = self.padded_frames_variant("slow", SLOW_FRAMES, 0)
        sc, sh = self.bn(p + "s1.pathway0_stem.bn", 1e-5)
        xs = self.stem_direct(p + "s1.pathway0_stem.conv", self.P(p + "s1.pathway0_stem.conv.weight"), sc, sh, 7, 2, 3,
                              ACT_RELU, self.new(B, ts, H // 2, W // 2, 64), frames=fr_s)
        fr_f = self.padded_frames_variant("fast", None, 2)
        sc, sh = self.bn(p + "s1.pathway1_stem.bn", 1e-5)
        xf = self.stem_direct(p + "s1.pathway1_stem.conv", self.P(p + "s1.pathway1_stem.conv.weight"), sc, sh, 7, 2, 3,
                              ACT_RELU, self.new(B, T, H // 2, W // 2, 8), frames=fr_f, clips=B)
        cat = self.new(B, ts, H // 4, W // 4, 64 + 16)
        self.pool(p + "s1.pathway0_stem.pool", xs, (1, 3, 3), (1, 2, 2), (0, 1, 1), cat.slice(0, 64))
        xf = self.pool(p + "s1.pathway1_stem.pool", xf, (1, 3, 3), (1, 2, 2), (0, 1, 1))

        def fuse(q: str, xf: Act, dst: Act):
            """FuseFastToSlow: (5,1,1)/s(4,1,1) conv + BN + ReLU written next to the slow features (sf.py:101-159)."""
            sc, sh = self.bn(q + ".bn", 1e-5)
            self.conv(q + ".conv_f2s", xf, self.P(q + ".conv_f2s.weight"), sc, sh, stride=(4, 1, 1), pad=(2, 0, 0),
                      act=ACT_RELU, out=dst)

        fuse(p + "s1_fuse", xf, cat.slice(64, 16))
        xs = cat
        feats = []
        for si, depth in enumerate(DEPTHS):
            outs, outf = 256 * 2 ** si, 32 * 2 ** si
            stride = 1 if si == 0 else 2
            nxt = None
            for i in range(depth):
                st = stride if i == 0 else 1
                lastb = i == depth - 1
                dst = None
                if lastb and si < 3:
                    nxt = self.new(B, ts, xs.h // st, xs.w // st, outs + 2 * outf)
                    dst = nxt.slice(0, outs)
                elif lastb and si == 3:
                    dst = v4_out
                xs = self.bottleneck_block(f"{p}s{si + 2}.pathway0_res{i}", xs, st, TK_SLOW[si], dst)
                if si < 3 or not lastb or True:
                    xf = self.bottleneck_block(f"{p}s{si + 2}.pathway1_res{i}", xf, st, TK_FAST[si])
            if si < 3:
                fuse(f"{p}s{si + 2}_fuse", xf, nxt.slice(outs, 2 * outf))
                xs = nxt
            feats.append(xs)
            self.tap(f"visnet.base{si + 1}", xs)
        return feats

    # ------------------------------------------------------------------ ResNet18 audio (backbones/resnet.py)
    def resnet18(self) -> Act:
        p = "audnet."
        B = self.B
        sc, sh = self.bn(p + "bn1", 1e-5)
        x = self.stem_gemm(p + "conv1", "audio", (B, 1, 1, 257, 111), self.P(p + "conv1.weight"), sc, sh,
                           (1, 7, 7), (1, 2, 2), (0, 3, 3), ACT_RELU)
        x = self.pool(p + "maxpool", x, (1, 3, 3), (1, 2, 2), (0, 1, 1))
        for li in range(1, 5):
            for bi in range(2):
                q = f"{p}layer{li}.{bi}"
                s = 2 if (li > 1 and bi == 0) else 1
                idn = x
                if (q + ".downsample.0.weight") in self.sd:
                    dsc, dsh = self.bn(q + ".downsample.1", 1e-5)
                    idn = self.conv(q + ".downsample.0", x, self.P(q + ".downsample.0.weight"), dsc, dsh, (1, s, s))
                sc1, sh1 = self.bn(q + ".bn1", 1e-5)
                o = self.conv(q + ".conv1", x, self.P(q + ".conv1.weight"), sc1, sh1, (1, s, s), (0, 1, 1), ACT_RELU)
                sc2, sh2 = self.bn(q + ".bn2", 1e-5)
                x = self.conv(q + ".conv2", o, self.P(q + ".conv2.weight"), sc2, sh2, (1, 1, 1), (0, 1, 1), ACT_RELU,
                              residual=idn)  # relu(bn2(conv2) + identity), resnet.py:44-52
        self.tap("audnet", x)
        return x

    # ------------------------------------------------------------------ ConvNeXt-T image encoder
    def convnext(self) -> Tuple[Act, Act]:
        """timm convnext_tiny(features_only) on the B*T frames + smooth convs, model_utils.py:357-385.
        Frames are the (b t) order of rearrange('b c t h w -> (b t) c h w'), i.e. exactly the
        N,T-major rows of the clip tensor, so N=B*T, T=1 here."""
        p = "image_encoder.encoder."
        B, T, H, W = self.B, self.T, self.H, self.W
        nf = B * T
        h, w = H // 4, W // 4
        if self.fuse_stem_norm and not self.keep_taps and ops.stem_conv_ln_ok(W, 96, 4, 4, 0):
            # stem_1 (LayerNorm2d) as the epilogue of the stem GEMM: the fp32 conv output (1 GB at B = 32) is never written
            x = self.stem_direct(p + "stem_0+1", self.P(p + "stem_0.weight"), None, self.P(p + "stem_0.bias"), 4, 4, 0,
                                 ACT_NONE, self.new(nf, 1, h, w, 96),
                                 ln=(self.P(p + "stem_1.weight"), self.P(p + "stem_1.bias"), 1e-6))
        else:
            x32 = self.stem_direct(p + "stem_0", self.P(p + "stem_0.weight"), None, self.P(p + "stem_0.bias"), 4, 4, 0,
                                   ACT_NONE, self.new(nf, 1, h, w, 96, torch.float32))
            x = self.new(nf, 1, h, w, 96)
            self.add(p + "stem_1", ops.layernorm(x32.buf, x.buf, nf * h * w, 96, self.P(p + "stem_1.weight"),
                                                 self.P(p + "stem_1.bias"), 1e-6))
        dims, depths = (96, 192, 384, 768), (3, 3, 9, 3)
        feats = []
        normed = False
        for s, (d, depth) in enumerate(zip(dims, depths)):
            q = f"{p}stages_{s}."
            if s > 0:
                if normed:      # the previous stage's last fused MLP already stored LayerNorm(x)
                    xn, normed = x, False
                else:
                    xn = self.new(nf, 1, h, w, dims[s - 1])
                    self.add(q + "downsample.0", ops.layernorm(x.buf, xn.buf, nf * h * w, dims[s - 1],
                                                               self.P(q + "downsample.0.weight"),
                                                               self.P(q + "downsample.0.bias"), 1e-6))
                x = self.conv(q + "downsample.1", xn, self.P(q + "downsample.1.weight"), None,
                              self.P(q + "downsample.1.bias"), (1, 2, 2))
                h, w = h // 2, w // 2
            for j in range(depth):
                b = f"{q}blocks.{j}."
                y = self.new(nf, 1, h, w, d)
                self.add(b + "conv_dw+norm", ops.dwconv_ln(x, y, self.P(b + "conv_dw.weight"), self.P(b + "conv_dw.bias"),
                                                           self.P(b + "norm.weight"), self.P(b + "norm.bias"), 1e-6))
                if d in (96, 192) and self.fuse_mlp:
                    # stages 0 and 1: fc1 + GELU + fc2 + layer scale + residual in one kernel, hidden tile on chip
                    out = self.new(nf, 1, h, w, d)
                    # last block of the stage: its output feeds only the next stage's downsample.0 LayerNorm2d, which the
                    # kernel applies before storing (a training plan keeps the un-normalised output for the backward pass)
                    ln = None
                    if j == depth - 1 and self.fuse_ds_norm and not self.keep_taps:
                        nq = f"{p}stages_{s + 1}.downsample.0."
                        ln = (self.P(nq + "weight"), self.P(nq + "bias"), 1e-6)
                        normed = True
                    run = ops.mlp_fused(y, out, x, self.P(b + "mlp.fc1.weight"), self.P(b + "mlp.fc1.bias"),
                                        self.P(b + "mlp.fc2.weight"), self.P(b + "mlp.fc2.bias"), self.P(b + "gamma"), ln=ln)
                    self.add(b + ("mlp(fused)+norm" if ln else "mlp(fused)"), run)
                    self.flops += run.flops
                    x = out
                    continue
                hid = self.linear(b + "mlp.fc1", y, b + "mlp.fc1.weight", b + "mlp.fc1.bias", ACT_GELU)
                # x + gamma * (fc2(h) + bias): layer scale folded into the epilogue scale/shift
                x = self.linear(b + "mlp.fc2", hid, b + "mlp.fc2.weight", b + "mlp.fc2.bias", residual=x,
                                scale=self.P(b + "gamma"))
            feats.append(x)
        o1, o0 = feats[2], feats[3]
        i = "image_encoder."
        sc, sh = self.bn(i + "smooth_1.1", 1e-5, self.P(i + "smooth_1.0.bias"))
        s1 = self.conv(i + "smooth_1.0", o1, self.P(i + "smooth_1.0.weight"), sc, sh, pad=(0, 1, 1), act=ACT_RELU)
        sc, sh = self.bn(i + "smooth_0.1", 1e-5, self.P(i + "smooth_0.0.bias"))
        s0 = self.conv(i + "smooth_0.0", o0, self.P(i + "smooth_0.0.weight"), sc, sh, pad=(0, 1, 1), act=ACT_RELU)
        self.tap("image_encoder.o1", s1)
        self.tap("image_encoder.o0", s0)
        return s1, s0

    def cached_features(self) -> Tuple[Act, Act]:
        """The (b t)-ordered image-encoder maps of this batch of windows, gathered from the per-frame cache
        (self._inputs["feat_o1"] [N,H/16,W/16,96], ["feat_o0"] [N,H/32,W/32,320], bf16) through self.frame_index [B*T]."""
        nf = self.B * self.T
        self.frame_index = torch.zeros((nf,), dtype=torch.int32, device=self.device)
        o1 = self.new(nf, 1, self.H // 16, self.W // 16, 96)
        o0 = self.new(nf, 1, self.H // 32, self.W // 32, 320)
        self.add("image_encoder.o1<-cache", ops.gather_rows(self._inputs, "feat_o1", self.frame_index, o1.buf, o1.h * o1.w * o1.c))
        self.add("image_encoder.o0<-cache", ops.gather_rows(self._inputs, "feat_o0", self.frame_index, o0.buf, o0.h * o0.w * o0.c))
        self.tap("image_encoder.o1", o1)
        self.tap("image_encoder.o0", o0)
        return o1, o0

    def adapter(self, o1: Act, o0: Act) -> Act:
        """Adapter.forward, model_utils.py:202-220"""
        B, T = self.B, self.T
        st = T // 4
        o1 = Act(o1.buf.view(B, T, o1.h, o1.w, o1.cs))  # (b t) frames are already b-major, t-minor
        o0 = Act(o0.buf.view(B, T, o0.h, o0.w, o0.cs))
        cat = self.new(B, 4, o1.h, o1.w, o1.c + o0.c)
        self.pool("adapter.pool_time.o3", o1, (st, 1, 1), (st, 1, 1), (0, 0, 0), cat.slice(0, o1.c))
        p0 = self.pool("adapter.pool_time.o2", o0, (st, 1, 1), (st, 1, 1), (0, 0, 0))
        self.up("adapter.up", p0, 2, cat.slice(o1.c, o0.c))
        masks = self.mixed("adapter.conv", cat)
        self.tap("adapter", masks)
        return masks

    # ------------------------------------------------------------------ SyncBlock + SimSiam heads
    @staticmethod
    def sinusoid(n: int, d: int) -> torch.Tensor:
        """model_utils.py:18-29 (fp64 table cast to fp32)"""
        pos = np.arange(n, dtype=np.float64)[:, None]
        j = np.arange(d)[None, :]
        tab = pos / np.power(10000, 2 * (j // 2) / d)
        tab[:, 0::2] = np.sin(tab[:, 0::2])
        tab[:, 1::2] = np.cos(tab[:, 1::2])
        return torch.tensor(tab, dtype=torch.float32)

    def sync_block(self, v4: Act, aud: Act, v4cat: Act):
        """SyncBlock.forward + forward_encoder's token split and SimSiam loss, model_utils.py:257-282,540-552"""
        p = "aud_vis_sync_block."
        B = self.B
        nv, na, c = v4.t * v4.h * v4.w, aud.h * aud.w, 512
        n = nv + na
        dev = self.device
        f32 = torch.float32
        stream = torch.empty((B, n, c), dtype=f32, device=dev)  # residual stream, fp32
        self._keep.append(stream)
        # v4 arrives in fp32 (the S3D head writes it that way, see _build) and the block runs in tf32: the fused
        # tokens go straight back into the decoder, which is precision critical (see "decoder" below).
        vp = self.conv(p + "vis_proj", self.flat(v4), self.P(p + "vis_proj.weight"), None, self.P(p + "vis_proj.bias"),
                       out_dtype=f32, dtype=v4.dtype)
        pos_v, pos_a = self.sinusoid(nv, c).to(dev), self.sinusoid(na, c).to(dev)
        self.add(p + "vis_norm", ops.layernorm(vp.buf, stream, B * nv, c, self.P(p + "vis_norm.weight"),
                                               self.P(p + "vis_norm.bias"), 1e-5, pos=pos_v, rows_per_group=nv,
                                               out_gstride=n * c))
        assert aud.c0 == 0 and aud.cs == c
        self.add(p + "aud_norm", ops.layernorm(aud.buf, stream, B * na, c, self.P(p + "aud_norm.weight"),
                                               self.P(p + "aud_norm.bias"), 1e-5, pos=pos_a, rows_per_group=na,
                                               out_gstride=n * c, y_off=nv * c))
        s_act = self.rows_act(stream.view(B * n, c))
        ln = torch.empty((B * n, c), dtype=f32, device=dev)
        ln_act = self.rows_act(ln)
        attn_out = torch.empty((B * n, c), dtype=f32, device=dev)

        def lin(name, x, wk, bk, act=ACT_NONE, out=None, residual=None):
            return self.conv(name, x, self.P(wk), None, self.P(bk) if bk else None, act=act, out=out, residual=residual,
                             res_after_act=residual is not None, out_dtype=f32, dtype=f32)

        for i in range(3):
            b = f"{p}blocks.{i}."
            self.add(b + "norm1", ops.layernorm(stream, ln, B * n, c, self.P(b + "norm1.weight"), self.P(b + "norm1.bias"), 1e-5))
            qkv = lin(b + "attn.qkv", ln_act, b + "attn.qkv.weight", None)
            for nm, fn in ops.attention_gemm(qkv.buf.view(B * n, 3 * c), attn_out, B, n, 4, c // 4):
                self.add(b + "attn." + nm, fn)
            self.flops += 4.0 * B * n * n * c
            lin(b + "attn.proj", self.rows_act(attn_out), b + "attn.proj.weight", b + "attn.proj.bias", out=s_act, residual=s_act)
            self.add(b + "norm2", ops.layernorm(stream, ln, B * n, c, self.P(b + "norm2.weight"), self.P(b + "norm2.bias"), 1e-5))
            hid = lin(b + "mlp.fc1", ln_act, b + "mlp.fc1.weight", b + "mlp.fc1.bias", ACT_GELU)
            lin(b + "mlp.fc2", hid, b + "mlp.fc2.weight", b + "mlp.fc2.bias", out=s_act, residual=s_act)
        self.taps_stream = stream
        # vis tokens -> channels [1024, 1536) of the concatenated v4 (torch.cat([v4, vis_sync]), :559)
        self.add("vis_sync.cat", ops.cast_rows(stream, v4cat.buf, B, nv, c, c, n * c, v4cat.cs, nv * v4cat.cs,
                                               dst_off=v4.c))
        # SimSiam heads (:404-435, 545-552)
        pooled_v = torch.empty((B, c), dtype=torch.float32, device=dev)
        pooled_a = torch.empty((B, c), dtype=torch.float32, device=dev)
        self.add("vis_pool", ops.token_mean(stream, pooled_v, B, n, 0, nv, c))
        self.add("aud_pool", ops.token_mean(stream, pooled_a, B, n, nv, n, c))

        def head(prefix: str, x32: torch.Tensor, idx, last_norm: bool) -> torch.Tensor:
            cur32 = x32
            for k, i in enumerate(idx):
                cin = cur32.shape[1]
                xb = torch.empty((B, cin), dtype=torch.bfloat16, device=dev)
                self.add(f"{prefix}.{i}.cast", ops.cast_rows(cur32, xb, 1, B, cin, cin, 0, cin, 0))
                y = self.linear(f"{prefix}.{i}", self.rows_act(xb), f"{prefix}.{i}.weight", f"{prefix}.{i}.bias",
                                out_dtype=torch.float32)
                cur32 = y.buf.view(B, -1)
                last = k == len(idx) - 1
                if not last or last_norm:
                    cout = cur32.shape[1]
                    z = torch.empty((B, cout), dtype=torch.float32, device=dev)
                    self.add(f"{prefix}.{i + 1}", ops.layernorm(cur32, z, B, cout, self.P(f"{prefix}.{i + 1}.weight"),
                                                                self.P(f"{prefix}.{i + 1}.bias"), 1e-5, relu=not last))
                    cur32 = z
            return cur32

        zv = head("vis_projector", pooled_v, (0, 3, 6), True)
        za = head("aud_projector", pooled_a, (0, 3, 6), True)
        pv = head("mlp_vis", zv, (0, 3), False)
        pa = head("mlp_aud", za, (0, 3), False)
        self.loss = torch.zeros((1,), dtype=torch.float32, device=dev)
        self.add("simsiam", ops.simsiam_loss(pv, za, pa, zv, self.loss, B, zv.shape[1]))

    # ------------------------------------------------------------------ decoder
    # Precision: everything after the encoders is min-max normalised by its consumers (inference.py:88,
    # compute_saliency_metrics.normalize_map) and, with freshly initialised weights, carries a pixel-to-pixel
    # signal far smaller than its bias-dominated mean.  bf16 storage there costs 1.1e-2 of the normalised map
    # (measured: DESIGN.md "precision"), over the 1e-2 budget, so the decoder keeps fp32 activations and runs
    # its GEMMs as kind::tf32 (fp32 accumulate); the encoders (87% of the FLOPs) stay bf16.
    def convnext_block3d(self, p: str, x: Act) -> Act:
        """ConvNextBlock, model_utils.py:306-354.  x: fp32."""
        f32 = torch.float32
        a = self.new(x.n, x.t, x.h, x.w, x.c, f32)
        self.add(p + ".dwconv_t", ops.dwconv_ln(x, a, self.P(p + ".dwconv_t.weight"), self.P(p + ".dwconv_t.bias")))
        b = self.new(x.n, x.t, x.h, x.w, x.c, f32)
        self.add(p + ".dwconv_s+norm", ops.dwconv_ln(a, b, self.P(p + ".dwconv_s.weight"), self.P(p + ".dwconv_s.bias"),
                                                     self.P(p + ".norm.norm.weight"), self.P(p + ".norm.norm.bias"), 1e-5))
        # the 4C hidden tensor is the block's largest (2.1 GB per step at latlayer_0): it is stored in bf16 (one rounding,
        # like every hidden activation of the encoders) and multiplied by hi/lo-split bf16 weights; the LayerNorm output,
        # the residual stream and the block output stay fp32.
        hid = self.conv(p + ".pwconv1", b, self.P(p + ".pwconv1.weight"), None, self.P(p + ".pwconv1.bias"), act=ACT_GELU,
                        out_dtype=torch.bfloat16, dtype=f32)
        return self.conv(p + ".pwconv2", hid, self.P(p + ".pwconv2.weight"), None, self.P(p + ".pwconv2.bias"),
                         residual=x, res_after_act=True, out_dtype=f32, split_weights=True)

    def lateral(self, k: int, x: Act) -> Act:
        """latlayer_k, model_utils.py:437-484.  Conv1x1(+bias) followed by the bias-free (s,1,1)/s temporal conv
        is one linear map: W[kt] = W1[:,:,kt] @ W0, b = sum_kt W1[:,:,kt] @ b0 — composed on the host in fp32 and
        run as ONE strided conv on the bf16 encoder tap, with hi/lo-split bf16 weights and an fp32 result."""
        p = f"latlayer_{k}"
        w0 = self.P(p + ".0.weight").float()[:, :, 0, 0, 0]  # [de, cin]
        b0 = self.P(p + ".0.bias").float()
        i = 1
        if self.lateral_bool[k]:
            s = self.lateral_stride[k]
            w1 = self.P(p + ".1.weight").float()[:, :, :, 0, 0]  # [de, de, s]
            w = torch.einsum("omk,mi->oik", w1, w0)[:, :, :, None, None]  # [de, cin, s, 1, 1]
            bias = torch.einsum("omk,m->o", w1, b0)
            y = self.conv(p + ".0+1", x, w, None, bias, stride=(s, 1, 1), out_dtype=torch.float32, dtype=x.dtype,
                          split_weights=x.dtype == torch.bfloat16)
            self.flops += 2.0 * y.pixels * w1.shape[0] * w1.shape[1] * s  # the reference's separate temporal conv
            i = 2
        elif x.dtype == torch.float32:
            y = self.conv(p + ".0", x, w0[:, :, None, None, None], None, b0, out_dtype=torch.float32, dtype=torch.float32)
        else:
            y = self.conv(p + ".0", x, w0[:, :, None, None, None], None, b0, out_dtype=torch.float32, split_weights=True)
        out = self.convnext_block3d(f"{p}.{i}", y)
        self.tap(p, out)
        return out

    def sa_masks(self, masks: Act) -> Act:
        """The three SA 512->32 3x3x3 BasicConv3d run on the same `masks` (model_utils.py:161, Appendix C.9):
        one GEMM with the three weight sets stacked along N."""
        ws, scs, shs = [], [], []
        for k in range(3):
            ws.append(self.P(f"sa_{k}.conv_mask.0.conv.weight"))
            sc, sh = self.bn(f"sa_{k}.conv_mask.0.bn", 1e-3)
            scs.append(sc), shs.append(sh)
        return self.conv("sa_*.conv_mask.0", masks, torch.cat(ws, 0), torch.cat(scs), torch.cat(shs), pad=(1, 1, 1),
                         act=ACT_RELU)

    def sa_gate(self, k: int, x: Act, m96: Act, scale: int, out: Optional[Act] = None, sources=()) -> Act:
        """SA.forward (model_utils.py:167-170); `sources` = [(Act, k)] adds the top-down terms up_k(src) in the same pass
        (model_utils.py:566-568)."""
        p = f"sa_{k}"
        m = m96.slice(32 * k, 32)
        if scale != 1:
            m = self.up(p + ".up", m, scale)
        logit = self.new(m.n, m.t, m.h, m.w, 1, torch.float32)
        self.add(p + ".conv_mask.2", ops.conv_c1(m, self.P(p + ".conv_mask.2.weight"), self.P(p + ".conv_mask.2.bias"), logit))
        self.flops += 2.0 * m.pixels * 32 * 9
        if out is None:
            out = self.new(x.n, x.t, x.h, x.w, x.c, x.dtype)
        if sources and x.dtype == torch.float32:
            self.add(p + ".gate+fuse", ops.sa_gate_fused(x, logit.buf.view(-1), out, list(sources)))
            return out
        self.add(p + ".gate", ops.sa_gate(x, logit.buf.view(-1), out))
        for src, kk in sources:
            self.up(f"{p}.fuse+=up{kk}", src, kk, out, accumulate=True)
        return out

    def readout(self, g0: Act, g1: Act, g2: Act, s3: Act) -> torch.Tensor:
        """readout Sequential + log-softmax, model_utils.py:490-504,570-572.

        readout.0 is a 1x1x1 conv over cat[s0, up2(s1), up4(s2), up8(s3)]: linear and pointwise in space, so it
        commutes with the bilinear upsamples.  Each 192-channel group is multiplied by its slice of the weight at
        its own resolution and the products are upsample-accumulated into one 192-channel map; the 768-channel
        concatenation (66 MB per clip in fp32) is never materialised."""
        p = "readout"
        f32 = torch.float32
        w0 = self.P(p + ".0.weight")
        c = g0.c
        x = self.conv(p + ".0[s0]", g0, w0[:, 0:c], None, self.P(p + ".0.bias"), out_dtype=f32, dtype=f32)
        parts = []
        for i, (g, k) in enumerate(((g1, 2), (g2, 4), (s3, 8)), 1):
            parts.append((self.conv(f"{p}.0[s{i}]", g, w0[:, i * c:(i + 1) * c], out_dtype=f32, dtype=f32), k))
        # x += up2(part1) + up4(part2) + up8(part3) in ONE pass over x (in place; the gate kernel without a mask)
        self.add(f"{p}.0+=up2,4,8", ops.sa_gate_fused(x, None, x, parts))
        sc, sh = self.bn(p + ".2", 1e-5, self.P(p + ".1.bias"))
        if os.environ.get("MSPI_READOUT1_BF16", "0") != "0":
            # study switch: readout.1 (K = 5184, the largest tf32 layer) on bf16 operands
            xb = self.new(x.n, x.t, x.h, x.w, x.c, torch.bfloat16)
            self.add(p + ".0.cast", ops.cast_rows(x.buf, xb.buf, 1, x.pixels, x.c, x.cs, 0, xb.cs, 0))
            x = self.conv(p + ".1", xb, self.P(p + ".1.weight"), sc, sh, pad=(1, 1, 1), act=ACT_RELU, out_dtype=f32,
                          dtype=torch.bfloat16, split_weights=os.environ.get("MSPI_READOUT1_BF16") == "2")
        else:
            x = self.conv(p + ".1", x, self.P(p + ".1.weight"), sc, sh, pad=(1, 1, 1), act=ACT_RELU, out_dtype=f32, dtype=f32)
        sc, sh = self.bn(p + ".5", 1e-5, self.P(p + ".4.bias"))
        x = self.conv(p + ".4", x, self.P(p + ".4.weight"), sc, sh, pad=(0, 1, 1), act=ACT_RELU, out_dtype=f32, dtype=f32)
        # readout.7-9: Upsample(1,4,4) -> Conv(4,1,1)/s4 -> ReLU.  The conv mixes T and C only, the upsample H and
        # W only: both linear, so conv first (16x fewer pixels, no 64ch full-resolution tensor), ReLU after.
        x = self.conv(p + ".8", x, self.P(p + ".8.weight"), None, self.P(p + ".8.bias"), stride=(4, 1, 1),
                      out_dtype=f32, dtype=f32)
        x = self.up(p + ".7", x, 4, act=ACT_RELU)
        x = self.conv(p + ".10", x, self.P(p + ".10.weight"), None, self.P(p + ".10.bias"), pad=(0, 1, 1), act=ACT_RELU,
                      out_dtype=f32, dtype=f32)
        x12 = self.new(x.n, x.t, x.h, x.w, 1, f32)
        self.add(p + ".12", ops.conv_c1(x, self.P(p + ".12.weight"), self.P(p + ".12.bias"), x12))
        self.flops += 2.0 * x.pixels * 32 * 9
        x = x12
        assert x.t == 1 and x.c == 1
        self.logits = x.buf.view(self.B, self.H * self.W)
        self.out = torch.empty((self.B, self.H, self.W), dtype=torch.float32, device=self.device)
        self.add("log_softmax", ops.logsoftmax2d(self.logits, self.out, self.B, self.H * self.W))
        return self.out

    # ------------------------------------------------------------------ whole forward
    def _build(self):
        """The forward as three branches that only meet where the reference's data flow does (model_utils.py:556-570):
             branch 0  clip conversion -> image encoder (ConvNeXt-T) -> adapter -> SA mask conv ........ -> gates, readout
             branch 1  motion encoder (S3D / X3D-L / SlowFast) -> [wait 2] SyncBlock + heads -> laterals -^
             branch 2  audio encoder (ResNet18)
        Captured into one CUDA graph, the branches become parallel paths: the small-grid kernels (audio net, SyncBlock,
        SimSiam heads, the deep S3D stages) fill SMs that the tails of the big encoder kernels leave idle."""
        B = self.B
        if self.mode == "image_encoder":
            self.feat_o1, self.feat_o0 = self.convnext()
            return
        if self.encoder != "slowfast4x16":
            self.padded_frames()       # shared by the image-encoder stem and the motion-encoder stem
        self.edge(0, 1), self.edge(0, 2)
        if self.mode == "cached":
            o1, o0 = self.cached_features()
        else:
            o1, o0 = self.convnext()
        masks = self.adapter(o1, o0)
        m96 = self.sa_masks(masks)
        h32, w32 = self.H // 32, self.W // 32
        visnet = {"s3d": self.s3d, "x3dl": self.x3d, "slowfast4x16": self.slowfast}[self.encoder]
        c4 = {"s3d": 1024, "x3dl": 192, "slowfast4x16": 2048}[self.encoder]
        t4 = self.T if self.encoder == "x3dl" else self.T // 4   # X3D keeps all 16 frames (X3D.py), the others end on 4
        self.on_branch(1)
        if self.audio:
            v4cat = self.new(B, t4, h32, w32, c4 + 512, torch.float32)  # cat([v4, vis_sync]) feeds the decoder
            v1, v2, v3, v4 = visnet(v4cat.slice(0, c4))
            self.on_branch(2)
            aud = self.resnet18()
            self.on_branch(1)
            self.edge(2, 1)
            self.sync_block(v4, aud, v4cat)
            v4in = v4cat
        else:
            v1, v2, v3, v4 = visnet(None)
            v4in = v4
            self.loss = torch.zeros((1,), dtype=torch.float32, device=self.device)
        s3 = self.lateral(3, v4in)
        s0 = self.lateral(0, v1)
        s1 = self.lateral(1, v2)
        s2 = self.lateral(2, v3)
        assert s0.t == s1.t == s2.t == s3.t == masks.t, "laterals must land on the adapter's 4 frames (model_utils.py:169)"
        self.on_branch(0)
        self.edge(1, 0)
        # top-down fusion, model_utils.py:566-568 (fp32)
        g2 = self.sa_gate(2, s2, m96, 1, sources=[(s3, 2)])
        g1 = self.sa_gate(1, s1, m96, 2, sources=[(g2, 2), (s3, 4)])
        g0 = self.sa_gate(0, s0, m96, 4, sources=[(g1, 2), (g2, 4), (s3, 8)])
        self.tap("fuse.s2", g2), self.tap("fuse.s1", g1), self.tap("fuse.s0", g0)
        self.readout(g0, g1, g2, s3)

    # ------------------------------------------------------------------ execution
    def bind(self, clips: torch.Tensor, audio: Optional[torch.Tensor], feats=None, frame_index=None):
        if self.input_u8:
            assert clips.is_cuda and clips.dtype == torch.uint8 and clips.is_contiguous()
            assert tuple(clips.shape) == (self.B, self.T, self.H, self.W, 3), f"plan is for uint8 {(self.B, self.T, self.H, self.W, 3)}"
        else:
            assert clips.is_cuda and clips.dtype == torch.float32 and clips.is_contiguous()
            assert tuple(clips.shape) == (self.B, 3, self.T, self.H, self.W), f"plan is for {(self.B, 3, self.T, self.H, self.W)}"
        self._inputs["clips"] = clips
        if self.audio and self.mode != "image_encoder":
            assert audio is not None and audio.is_cuda and audio.dtype == torch.float32 and audio.is_contiguous()
            assert tuple(audio.shape) == (self.B, 1, 257, 111), "audio must be [B,1,257,111] (inference.py:26)"
            self._inputs["audio"] = audio
        if self.mode == "cached":
            o1, o0 = feats
            assert o1.is_cuda and o0.is_cuda and o1.dtype == o0.dtype == torch.bfloat16 and o1.shape[0] == o0.shape[0]
            assert tuple(o1.shape[1:]) == (self.H // 16, self.W // 16, 96) and tuple(o0.shape[1:]) == (self.H // 32, self.W // 32, 320)
            self._inputs["feat_o1"], self._inputs["feat_o0"] = o1.contiguous(), o0.contiguous()
            assert frame_index.numel() == self.B * self.T
            self.frame_index.copy_(frame_index.reshape(-1).to(torch.int32), non_blocking=True)

    def run_eager(self):
        for _name, fn in self.steps:
            fn()

    def run_inputs(self):
        """The steps that read the caller's tensors (clip conversion, spectrogram patch gather), on the current stream."""
        for i in sorted(self.input_steps):
            self.steps[i][1]()

    def run_branched(self, skip_inputs: bool = False):
        """Enqueue the steps on one stream per branch (branch 0 = the current stream) with event edges between them.
        Under stream capture this records the fork / join structure of _build into the graph."""
        main = torch.cuda.current_stream()
        if getattr(self, "_side_streams", None) is None:
            self._side_streams = {}
        streams = {0: main}

        def stream_of(b):
            if b not in streams:
                if b not in self._side_streams:
                    self._side_streams[b] = torch.cuda.Stream(device=self.device)
                streams[b] = self._side_streams[b]
            return streams[b]

        used = set()
        for ent in self.sched:
            if ent[0] == "edge":
                _, src, dst = ent
                ev = torch.cuda.Event()
                ev.record(stream_of(src))
                stream_of(dst).wait_event(ev)
                used.add(dst)
            else:
                if skip_inputs and ent[1] in self.input_steps:
                    continue
                b = self.step_branch[ent[1]]
                fn = self.steps[ent[1]][1]
                if b == 0:
                    fn()
                else:
                    with torch.cuda.stream(stream_of(b)):
                        fn()
        for b in used:   # every side branch must have been joined back (a capture cannot end with unjoined streams)
            if b != 0:
                main.wait_stream(streams[b])

    def capture(self):
        """Capture the step list into a CUDA graph bound to static input buffers."""
        dev = self.device
        assert self.mode == "full", "graph capture is for the whole forward (the cached / image-encoder plans run eagerly)"
        if self.input_u8:
            self.static_clips = torch.zeros((self.B, self.T, self.H, self.W, 3), dtype=torch.uint8, device=dev)
        else:
            self.static_clips = torch.zeros((self.B, 3, self.T, self.H, self.W), dtype=torch.float32, device=dev)
        self.static_audio = torch.zeros((self.B, 1, 257, 111), dtype=torch.float32, device=dev) if self.audio else None
        self.bind(self.static_clips, self.static_audio)
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            self.run_eager()  # warm-up (module loading, attribute setting) outside capture
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # The graph holds everything EXCEPT the steps that read the caller's tensors: those run eagerly in front of every
        # replay (run()), straight from the caller's buffers — a replay copies no inputs.
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if self.branches and any(e[0] == "edge" for e in self.sched):
                self.run_branched(skip_inputs=True)
            else:
                for i, (_name, fn) in enumerate(self.steps):
                    if i not in self.input_steps:
                        fn()
        self.graph = g
        self.static_clips = self.static_audio = None   # only needed for the warm-up above

    def run(self, clips: torch.Tensor, audio: Optional[torch.Tensor], feats=None, frame_index=None):
        if self.graph is not None:
            self.bind(clips, audio)
            self.run_inputs()          # clip conversion / patch gather from the caller's tensors (not in the graph)
            self.graph.replay()
        else:
            self.bind(clips, audio, feats, frame_index)
            self.run_eager()
        if self.mode == "image_encoder":
            return self.feat_o1, self.feat_o0
        return self.out, self.loss

    @property
    def num_launches(self) -> int:
        """Kernels per forward (patch-gather steps of non-stem convs launch two)."""
        return getattr(self, "_launches", len(self.steps))
