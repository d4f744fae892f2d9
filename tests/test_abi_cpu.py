"""CPU tests of the C-ABI boundary and the host-side operator logic (no GPU compute).

  * the shared library loads and exports every function include/mspi_b200.h declares, and the ctypes binding
    (mspi_b200/_lib.py) covers exactly that set;
  * the ctypes mirrors of the descriptor structs have the C compiler's size (gcc on the header);
  * compute entry points fail loudly without a usable GPU (no CPU fallback), and the nn.Module refuses CPU tensors;
  * host logic: tile-box choice, N-tile choice, BatchNorm folding, weight packing, clip sharding is in
    tests/test_distributed_cpu.py.
"""
import ctypes as C
import os
import re
import subprocess
import tempfile

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mspi_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mspi_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from mspi_b200 import _lib
    lib = _lib.load()
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mspi_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == names, set(names) ^ set(_lib.EXPORTED_SYMBOLS)
    assert lib.mspi_arch().decode() == "sm_100a" and lib.mspi_version() >= 1


def test_ctypes_structs_match_the_c_layout():
    from mspi_b200 import _lib
    structs = {"MspiConvDesc": _lib.ConvDesc, "MspiPatchDesc": _lib.PatchDesc, "MspiPoolDesc": _lib.PoolDesc,
               "MspiUpDesc": _lib.UpDesc, "MspiDwDesc": _lib.DwDesc, "MspiDw3dDesc": _lib.Dw3dDesc, "MspiLnDesc": _lib.LnDesc,
               "MspiSgemmDesc": _lib.SgemmDesc, "MspiPermDesc": _lib.PermDesc}
    prog = '#include <stdio.h>\n#include "mspi_b200.h"\nint main(void){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in structs) + "return 0;}"
    with tempfile.TemporaryDirectory() as d:
        src, exe = os.path.join(d, "s.c"), os.path.join(d, "s")
        open(src, "w").write(prog)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), src, "-o", exe], check=True)
        out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    sizes = dict(line.split() for line in out.strip().splitlines())
    for n, cls in structs.items():
        assert C.sizeof(cls) == int(sizes[n]), (n, C.sizeof(cls), sizes[n])
    assert _lib.MAX_TAPS == int(re.search(r"#define MSPI_MAX_TAPS (\d+)", open(HEADER).read()).group(1))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_entry_points_fail_loudly_without_gpu():
    from mspi_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(4, 8)
    rc = lib.mspi_logsoftmax2d(C.c_void_p(x.data_ptr()), C.c_void_p(x.data_ptr()), 4, 8, C.c_void_p(0))
    assert rc == -2 and b"CUDA" in lib.mspi_last_error()  # MSPI_ERR_CUDA: no device, no fallback
    with pytest.raises(_lib.MspiError):
        _lib.check(rc, "logsoftmax2d")


def test_module_refuses_cpu_tensors_and_train_mode():
    import contextlib
    import copy
    import io
    from mspi_b200.config import cfg
    from mspi_b200.model.model_utils import VisualSaliencyModel
    with contextlib.redirect_stdout(io.StringIO()):
        m = VisualSaliencyModel(copy.deepcopy(cfg), load_pretrained=False).eval()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 3, 16, 64, 64))
    with pytest.raises(Exception, match="Invalid Motion Encoder"):
        c = copy.deepcopy(cfg)
        c.MODEL.MOTION_ENCODER = "mvitv2s"
        with contextlib.redirect_stdout(io.StringIO()):
            VisualSaliencyModel(c, load_pretrained=False)


def test_state_dict_keys_match_the_reference_for_every_encoder():
    """Key names + shapes are the checkpoint compatibility contract: oracle.param_spec was checked against
    the live reference's state_dict when the golden fixtures were generated."""
    import contextlib
    import copy
    import io
    from oracle import mspi_oracle as orc
    from mspi_b200.config import cfg, select_motion_encoder
    from mspi_b200.model.model_utils import AudioVisualSaliencyModel
    for enc in ("s3d", "x3dl", "slowfast4x16"):
        with contextlib.redirect_stdout(io.StringIO()):
            m = AudioVisualSaliencyModel(select_motion_encoder(enc, copy.deepcopy(cfg)), load_pretrained=False)
        spec = orc.param_spec(True, enc)
        sd = m.state_dict()
        assert set(sd) == set(spec), (enc, list(set(sd) ^ set(spec))[:5])
        assert all(tuple(sd[k].shape) == tuple(spec[k][1]) for k in sd)
        frozen = [n for n, _ in m.named_parameters() if n.startswith(("audnet", "image_encoder"))]
        assert frozen, "train.py:151-155 freezes by these prefixes"


def test_host_tiling_and_packing_logic():
    from mspi_b200 import ops
    for dims in [(96, 56, 1, 512), (12, 7, 4, 32), (192, 112, 16, 2), (1, 1, 1, 1), (372, 1, 4, 32), (24, 14, 4, 1)]:
        box = ops.choose_box(dims)
        rows = box[0] * box[1] * box[2] * box[3]
        assert 1 <= rows <= 128 and all(1 <= b <= d for b, d in zip(box, dims))
    assert ops.choose_box((96, 56, 1, 512)) == (32, 4, 1, 1) or ops.choose_box((96, 56, 1, 512))[0] * ops.choose_box((96, 56, 1, 512))[1] == 128
    for cout in (1, 16, 24, 96, 192, 256, 320, 384, 480, 768, 1536, 3072):
        for chunk in (64, 32):
            bn = ops.choose_bn(cout, chunk)
            assert 16 <= bn <= 256 and bn % 16 == 0
            if cout > 256:
                assert bn % chunk == 0  # every N tile ends on a 128-byte output chunk (bulk stores)
    w = torch.randn(10, 5, 1, 3, 3)
    packed, taps, cin_pad = ops.pack_conv_weight(w, torch.bfloat16)
    assert taps == 9 and cin_pad == 64 and packed.shape == (16, 9 * 64)
    assert torch.equal(packed[:10].view(10, 9, 64)[:, :, :5].float(), w.permute(0, 2, 3, 4, 1).reshape(10, 9, 5).to(torch.bfloat16).float())
    assert (packed[10:] == 0).all() and (packed.view(16, 9, 64)[:, :, 5:] == 0).all()
    g, b, m, v = torch.rand(7) + 0.5, torch.randn(7), torch.randn(7), torch.rand(7) + 0.5
    sc, sh = ops.fold_bn(g, b, m, v, 1e-3, conv_bias=torch.ones(7))
    x = torch.randn(7)
    assert torch.allclose((x + 1 - m) / torch.sqrt(v + 1e-3) * g + b, x * sc + sh, atol=1e-5)


def test_image_encoder_checkpoint_load_is_loud():
    """model_utils.py:514 loads the image-encoder checkpoint with strict=False: a checkpoint whose keys do not match would
    silently leave the encoder at random init.  The product reports matched / missing / unexpected counts, maps the plain
    timm ConvNeXt key layout (stem.0 / stages.k.) onto the FeatureListNet names, and raises when nothing matches."""
    import pytest
    import torch
    from mspi_b200.model.model_utils import StaticSaliencyModelConvNext, load_image_encoder_checkpoint
    m = StaticSaliencyModelConvNext()
    own = {k: v.clone() + 1 for k, v in m.state_dict().items()}
    rep = load_image_encoder_checkpoint(StaticSaliencyModelConvNext(), own, verbose=False)
    assert rep["matched"] == rep["of"] == len(own) and not rep["missing"] and not rep["unexpected"]
    plain = {}
    for k, v in own.items():
        k2 = k.replace("encoder.stem_0", "stem.0").replace("encoder.stem_1", "stem.1")
        for i in range(4):
            k2 = k2.replace(f"encoder.stages_{i}.", f"stages.{i}.")
        plain[k2] = v
    plain["head.fc.weight"] = torch.zeros(3)
    tgt = StaticSaliencyModelConvNext()
    rep = load_image_encoder_checkpoint(tgt, plain, verbose=False)
    assert rep["matched"] == len(own) and rep["unexpected"] == ["head.fc.weight"]
    assert torch.equal(tgt.state_dict()["encoder.stages_2.blocks.4.mlp.fc1.weight"], own["encoder.stages_2.blocks.4.mlp.fc1.weight"])
    bad = dict(own)
    bad["smooth_0.0.weight"] = torch.zeros(1, 2, 3, 3)
    rep = load_image_encoder_checkpoint(StaticSaliencyModelConvNext(), bad, verbose=False)
    assert rep["shape_mismatch"] == ["smooth_0.0.weight"] and "smooth_0.0.weight" in rep["missing"]
    with pytest.raises(RuntimeError):
        load_image_encoder_checkpoint(StaticSaliencyModelConvNext(), {"backbone.conv1.weight": torch.zeros(1)}, verbose=False)
    assert load_image_encoder_checkpoint(StaticSaliencyModelConvNext(), {}, verbose=False)["matched"] == 0   # the tests' empty file


def test_dispatch_predicates_of_the_specialised_kernels(monkeypatch):
    """Host logic that routes a layer to a specialised kernel: the few-channel (1,3,3) conv of the SlowFast fast pathway
    (ops.conv133_small_ok), the ConvNeXt stem with its LayerNorm epilogue (ops.stem_conv_ln_ok), the launch-protocol switch
    (mspi_set_pdl returns the previous setting)."""
    from mspi_b200 import _lib, ops

    class FakeAct:           # the predicate only reads these attributes
        def __init__(self, c, c0=0, cs=None, dtype=torch.bfloat16):
            self.c, self.c0, self.cs, self.dtype = c, c0, cs or c, dtype

    bf, f32 = torch.bfloat16, torch.float32
    w8 = torch.zeros(8, 8, 1, 3, 3)
    ok = lambda *a, **k: ops.conv133_small_ok(*a, **k)
    assert ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8), bf, bf, None)
    assert ok(torch.zeros(16, 16, 1, 3, 3), (1, 1, 1), (0, 1, 1), FakeAct(16, 8, 32), bf, bf, None)
    assert not ok(w8, (1, 2, 2), (0, 1, 1), FakeAct(8), bf, bf, None)                # strided: implicit GEMM
    assert not ok(w8, (1, 1, 1), (1, 1, 1), FakeAct(8), bf, bf, None)                # temporal padding: not this layer
    assert not ok(torch.zeros(8, 8, 3, 3, 3), (1, 1, 1), (0, 1, 1), FakeAct(8), bf, bf, None)
    assert not ok(torch.zeros(32, 32, 1, 3, 3), (1, 1, 1), (0, 1, 1), FakeAct(32), bf, bf, None)
    assert not ok(torch.zeros(16, 8, 1, 3, 3), (1, 1, 1), (0, 1, 1), FakeAct(8), bf, bf, None)
    assert not ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8), f32, bf, None) and not ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8), bf, f32, None)
    assert not ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8), bf, bf, object())            # residual: implicit GEMM epilogue
    assert not ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8, 4, 16), bf, bf, None)         # pixel vectors must stay 16-byte aligned
    monkeypatch.setenv("MSPI_SMALLC_CONV", "0")
    assert not ok(w8, (1, 1, 1), (0, 1, 1), FakeAct(8), bf, bf, None)
    monkeypatch.delenv("MSPI_SMALLC_CONV")

    assert ops.stem_conv_ln_ok(384, 96, 4, 4, 0) and ops.stem_conv_ln_ok(96, 96, 4, 4, 0)
    assert not ops.stem_conv_ln_ok(388, 96, 4, 4, 0)          # odd output width: one pixel per row, N = 96 is not a 64-multiple
    assert not ops.stem_conv_ln_ok(384, 96, 7, 2, 3) and not ops.stem_conv_ln_ok(384, 160, 4, 4, 0)
    monkeypatch.setenv("MSPI_STEM_WIDE", "0")
    assert not ops.stem_conv_ln_ok(384, 96, 4, 4, 0)
    monkeypatch.delenv("MSPI_STEM_WIDE")

    lib = _lib.load()
    prev = lib.mspi_set_pdl(0)
    assert lib.mspi_set_pdl(1) == 0 and lib.mspi_set_pdl(prev) == 1
