"""GPU post-processing of saliency maps: what `process()` does on the CPU in the reference (inference.py:84-91):
GaussianBlur(11x11) of the log map -> exp -> resize -> min-max -> uint8."""
import ctypes as C

import torch

from . import _lib


def postprocess_maps(log_maps: torch.Tensor, img_size=(640, 480)) -> torch.Tensor:
    """log_maps: float32 CUDA [B,H,W] (the model output); img_size = (width, height) as cv2.resize takes it.
    Returns uint8 CUDA [B, height, width]."""
    if not log_maps.is_cuda:
        raise RuntimeError("mspi_b200.postprocess needs a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    x = log_maps.contiguous().float()
    b, h, w = x.shape
    ow, oh = int(img_size[0]), int(img_size[1])
    out = torch.empty((b, oh, ow), dtype=torch.uint8, device=x.device)
    work = torch.empty(b * h * w + b * oh * ow + 2 * b, dtype=torch.float32, device=x.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mspi_postprocess_maps(C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(work.data_ptr()),
                                         b, h, w, oh, ow, st), "postprocess_maps")
    return out
