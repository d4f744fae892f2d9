"""TEST INFRASTRUCTURE — not product code.

Makes the *unmodified* reference (/root/reference, oraclefina/MSPI) importable in
this container, where several of its third-party dependencies are not installed
(timm, fvcore, easydict, mmcv, iopath, pytorchvideo, simplejson, torch._six).
It installs small stand-in modules into ``sys.modules`` and never edits the
reference.  Only ``oracle/gen_golden.py`` (run HERE, where /root/reference
exists) uses this file; nothing on the GPU box imports it.

The one stand-in that carries arithmetic is ``timm.models.create_model
("convnext_tiny", features_only=True)``: timm==0.6.12 is pinned by the reference
(README.md:34, call site model/model_utils.py:361) but absent.  We restate its
published architecture (ConvNeXt-T: stem Conv4x4/s4 + LayerNorm2d(eps 1e-6);
block = dw Conv7x7 p3 -> LayerNorm(eps 1e-6) -> Linear C->4C -> exact GELU ->
Linear 4C->C -> *gamma -> +x; downsample = LayerNorm2d + Conv2x2/s2; depths
3/3/9/3, dims 96/192/384/768) with timm's FeatureListNet key names, and
``tests/test_oracle_cpu.py`` pins that restatement against torchvision's
independent ``convnext_tiny`` implementation.
"""
from __future__ import annotations

import contextlib
import importlib
import json
import os
import sys
import tempfile
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("MSPI_REF", "/root/reference")


# --------------------------------------------------------------------------- helpers
def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so sub-imports resolve
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


class EasyDict(dict):
    """Attribute-access dict (stand-in for easydict.EasyDict)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in {**(d or {}), **kw}.items():
            self[k] = v

    def __setitem__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            v = EasyDict(v)
        super().__setitem__(k, v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    __setattr__ = __setitem__


class CfgNode(dict):
    """Minimal yacs/fvcore CfgNode: attribute access, clone, YAML merge."""

    def __init__(self, init=None, **_):
        super().__init__()
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def clone(self):
        import copy
        return copy.deepcopy(self)

    def _merge(self, other):
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict):
                self[k]._merge(v)
            else:
                self[k] = CfgNode(v) if isinstance(v, dict) else v

    def merge_from_file(self, path):
        import yaml
        with open(path) as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, lst):
        for k, v in zip(lst[0::2], lst[1::2]):
            node = self
            parts = k.split(".")
            for p in parts[:-1]:
                node = node[p]
            node[parts[-1]] = v

    def merge_from_other_cfg(self, other):
        self._merge(other)

    def freeze(self):
        pass

    def defrost(self):
        pass

    def dump(self, **kw):
        return json.dumps(self)


class AverageMeter:
    """timm.utils.AverageMeter"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


class DropPath(nn.Module):
    def __init__(self, drop_prob=0.0, scale_by_keep=True):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):  # inference / p=0 only
        return x


def to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


# ------------------------------------------------------------- ConvNeXt-T (timm 0.6.12)
class _LayerNorm2d(nn.LayerNorm):
    """LayerNorm over C of an NCHW tensor (timm.models.layers.LayerNorm2d)."""

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class _Mlp(nn.Module):
    def __init__(self, d, h):
        super().__init__()
        self.fc1 = nn.Linear(d, h)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(h, d)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _CNBlock(nn.Module):
    def __init__(self, dim, ls_init=1e-6):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, 7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, 4 * dim)
        self.gamma = nn.Parameter(ls_init * torch.ones(dim))

    def forward(self, x):
        s = x
        x = self.conv_dw(x).permute(0, 2, 3, 1)
        x = self.mlp(self.norm(x)).permute(0, 3, 1, 2)
        return x * self.gamma.reshape(1, -1, 1, 1) + s


class _CNStage(nn.Module):
    def __init__(self, cin, cout, depth, down):
        super().__init__()
        self.downsample = (nn.Sequential(_LayerNorm2d(cin, eps=1e-6), nn.Conv2d(cin, cout, 2, 2))
                           if down else nn.Identity())
        self.blocks = nn.Sequential(*[_CNBlock(cout) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class ConvNeXtTinyFeatures(nn.Module):
    """Stand-in for timm FeatureListNet(convnext_tiny, out_indices=(0,1,2,3)).
    Module names follow timm's flatten_sequential naming: stem_0, stem_1, stages_k."""

    def __init__(self):
        super().__init__()
        dims, depths = (96, 192, 384, 768), (3, 3, 9, 3)
        self.stem_0 = nn.Conv2d(3, dims[0], 4, 4)
        self.stem_1 = _LayerNorm2d(dims[0], eps=1e-6)
        prev = dims[0]
        for i, (d, n) in enumerate(zip(dims, depths)):
            setattr(self, f"stages_{i}", _CNStage(prev, d, n, down=i > 0))
            prev = d

    def forward(self, x):
        x = self.stem_1(self.stem_0(x))
        outs = []
        for i in range(4):
            x = getattr(self, f"stages_{i}")(x)
            outs.append(x)
        return outs


def _create_model(name, pretrained=False, features_only=False, **kw):
    assert name == "convnext_tiny" and features_only, name
    return ConvNeXtTinyFeatures()


# ------------------------------------------------------------------------ installation
_installed = False


def install():
    """Install stand-in modules and put the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not os.path.isdir(REF_ROOT):
        raise FileNotFoundError(f"reference tree not found at {REF_ROOT}")

    _mod("easydict", EasyDict=EasyDict)

    timm = _mod("timm")
    _mod("timm.models", create_model=_create_model)
    timm.create_model = _create_model
    _mod("timm.models.layers", to_2tuple=to_2tuple, DropPath=DropPath,
         trunc_normal_=nn.init.trunc_normal_)
    _mod("timm.models.vision_transformer", VisionTransformer=type("VisionTransformer", (nn.Module,), {}),
         _cfg=lambda **kw: dict(kw))
    _mod("timm.utils", AverageMeter=AverageMeter, get_state_dict=lambda m, *a, **k: m.state_dict())
    _mod("timm.data")
    _mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
         IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))

    _mod("fvcore")
    _mod("fvcore.common")
    _mod("fvcore.common.config", CfgNode=CfgNode)
    _mod("fvcore.nn", FlopCountAnalysis=lambda *a, **k: None, flop_count_table=lambda *a, **k: "")

    _mod("mmcv")
    _mod("mmcv.utils", get_logger=lambda *a, **k: __import__("logging").getLogger("mmcv"))
    _mod("mmcv.runner", load_checkpoint=lambda *a, **k: None)

    class _PM:
        def __getattr__(self, k):
            if k == "open":
                return open
            if k == "exists":
                return os.path.exists
            if k == "mkdirs":
                return lambda p: os.makedirs(p, exist_ok=True)
            if k == "ls":
                return os.listdir
            return lambda *a, **kw: None

    class _PMF:
        @staticmethod
        def get(*a, **k):
            return _PM()

    _mod("iopath")
    _mod("iopath.common")
    _mod("iopath.common.file_io", PathManagerFactory=_PMF, g_pathmgr=_PM())

    ident = lambda *a, **k: None
    _mod("pytorchvideo")
    _mod("pytorchvideo.layers")
    _mod("pytorchvideo.layers.distributed", cat_all_gather=ident, get_local_process_group=ident,
         get_local_rank=lambda: 0, get_local_size=lambda: 1, get_world_size=lambda: 1,
         init_distributed_training=ident)

    class Swish(nn.Module):
        def forward(self, x):
            return x * torch.sigmoid(x)

    _mod("pytorchvideo.layers.swish", Swish=Swish)
    sys.modules.setdefault("simplejson", json)
    if "torch._six" not in sys.modules:
        _mod("torch._six", inf=float("inf"))

    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    _installed = True


@contextlib.contextmanager
def scratch_cwd():
    """Reference code uses cwd-relative paths (./weights, ./configs, ./checkpoints);
    run it from a scratch directory so nothing is written into the repo."""
    old = os.getcwd()
    with tempfile.TemporaryDirectory(prefix="mspi_ref_") as d:
        os.symlink(os.path.join(REF_ROOT, "configs"), os.path.join(d, "configs"))
        os.makedirs(os.path.join(d, "weights"))
        os.chdir(d)
        try:
            yield d
        finally:
            os.chdir(old)


def build_reference_model(state_dict, encoder="s3d", num_vis_tokens=None, audio=True):
    """Construct the reference AudioVisualSaliencyModel with `state_dict` loaded
    (strict) and eval() set.  Pretrained-file loads in the constructor
    (model_utils.py:512-514, resnet.py:151-152) are satisfied with files written
    to the scratch cwd; the constructor itself is untouched."""
    install()
    with scratch_cwd():
        cfgmod = importlib.import_module("config")
        cfg = cfgmod.cfg
        cfg.MODEL.MOTION_ENCODER = encoder
        cfg.MODEL.LATERAL_BOOL = cfgmod._LATERAL_BOOL[encoder]
        cfg.MODEL.LATERAL_STRIDE = [4, 4, 4, 4] if encoder == "x3dl" else [2, 2, 2, 2]
        cfg.MODEL.MOTION_ENCODER_WEIGHT = cfgmod._MOTION_WEIGHTS[encoder]
        if num_vis_tokens is not None:
            cfg.MODEL.NUM_VIS_TOKENS[encoder] = num_vis_tokens
        mu = importlib.import_module("model.model_utils")

        def sub(prefix):
            return {k[len(prefix):]: v for k, v in state_dict.items() if k.startswith(prefix)}

        if encoder == "s3d":
            torch.save(sub("visnet."), cfg.MODEL.MOTION_ENCODER_WEIGHT)
        elif encoder == "x3dl":       # X3D.load_weight: torch.load(path)['model_state'], strict=False (X3D.py:248-250)
            torch.save({"model_state": sub("visnet.")}, cfg.MODEL.MOTION_ENCODER_WEIGHT)
        elif encoder == "slowfast4x16":  # caffe2 pickle path of load_checkpoint (checkpoint.py:226-233): empty blobs
            import pickle
            with open(cfg.MODEL.MOTION_ENCODER_WEIGHT, "wb") as f:
                pickle.dump({"blobs": {}}, f)
        else:
            raise NotImplementedError(encoder)
        torch.save(sub("audnet."), cfg.MODEL.AUDIO_ENCODER_WEIGHT)
        torch.save({}, cfg.MODEL.IMAGE_SALIENCY_ENCODER_WEIGHT)
        cls = mu.AudioVisualSaliencyModel if audio else mu.VisualSaliencyModel
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            model = cls(cfg)
        if not audio:
            state_dict = {k: v for k, v in state_dict.items() if k in model.state_dict()}
        model.load_state_dict(state_dict, strict=True)
        model.eval()
    return model
