"""mspi_b200 — B200-native (sm_100a) implementation of MSPI's batched clip forward pass.

Host code is Python/PyTorch (memory, streams, torch.distributed); all arithmetic runs in the
hand-written CUDA kernels of ``mspi_b200/lib/libmspi_b200.so`` behind the C ABI declared in
``include/mspi_b200.h``.  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
