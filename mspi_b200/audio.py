"""Audio front end on the GPU: waveform slice -> log power spectrogram [B,1,257,111] (inference.py:44-60,
avsp_dataloader.py:51-80).  File decoding / resampling stay on the host."""
import ctypes as C

import torch

from . import _lib

SPECTRO_SHAPE = (257, 111)


def log_spectrogram(wave: torch.Tensor, frames_out: int = SPECTRO_SHAPE[1]) -> torch.Tensor:
    """wave: float32 CUDA tensor [B, n] at 16 kHz -> [B, 1, 257, frames_out]."""
    if not wave.is_cuda:
        raise RuntimeError("mspi_b200.audio.log_spectrogram needs a CUDA tensor (no CPU fallback)")
    lib = _lib.load()
    w = wave.contiguous().float()
    b, n = w.shape
    out = torch.empty((b, 1, SPECTRO_SHAPE[0], frames_out), dtype=torch.float32, device=w.device)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.mspi_logspec(C.c_void_p(w.data_ptr()), C.c_void_p(out.data_ptr()), b, n, frames_out, st), "logspec")
    return out
