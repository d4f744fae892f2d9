"""KLD / CC / SIM / NSS on the GPU — same names, arguments and return type (0-d tensor, batch mean) as
utils/compute_saliency_metrics.py:9-108 of the reference.  One fused multi-reduction kernel
(mspi_saliency_metrics) computes all four for a batch of maps; each function below returns its entry."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib


def saliency_metrics(pred: torch.Tensor, gt: torch.Tensor, fixations: torch.Tensor = None, pred_is_log: bool = False):
    """Returns a float32[5] device tensor {kld, cc, sim, nss, loss = kld - cc (- 0.1 nss)} (batch means)."""
    if not pred.is_cuda:
        raise RuntimeError("mspi_b200 metrics run on CUDA tensors only (no CPU fallback)")
    assert pred.shape == gt.shape and pred.dim() == 3, "expected [B,H,W] maps"
    lib = _lib.load()
    b = pred.shape[0]
    pixels = pred.shape[1] * pred.shape[2]
    p = pred.contiguous().float()
    g = gt.contiguous().float()
    f = None if fixations is None else fixations.contiguous().float()
    out = torch.empty(5, dtype=torch.float32, device=pred.device)
    work = torch.empty(16 * b, dtype=torch.float32, device=pred.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mspi_saliency_metrics(C.c_void_p(p.data_ptr()), 1 if pred_is_log else 0, C.c_void_p(g.data_ptr()),
                                         C.c_void_p(f.data_ptr() if f is not None else 0), C.c_void_p(out.data_ptr()),
                                         C.c_void_p(work.data_ptr()), b, pixels, st), "saliency_metrics")
    return out


def kldiv(s_map, gt):
    return saliency_metrics(s_map, gt)[0]


def cc(s_map, gt):
    return saliency_metrics(s_map, gt)[1]


def similarity(s_map, gt):
    return saliency_metrics(s_map, gt)[2]


def nss(s_map, gt):
    """gt is the binary fixation map (compute_saliency_metrics.py:95-108)."""
    assert s_map.size() == gt.size()
    return saliency_metrics(s_map, s_map, gt)[3]
