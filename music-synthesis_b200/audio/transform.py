"""Drop-in mirror of the FFT band split / merge in featuresynth/audio/transform.py:50-115.

`fft_frequency_decompose(x, min_size) -> {size: band}` (keys ascending, as the reference
builds them) and `fft_frequency_recompose(d, desired_size) -> tensor`, on CUDA tensors,
through `ms_fft_frequency_decompose / recompose` (hand-written Stockham FFT passes; the
reference uses the removed `torch.rfft / irfft`).  Forward only.
"""
import ctypes

import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr


def _workspace(batch, n, device):
    nbytes = _lib.lib().ms_fft_bands_workspace_bytes(batch, n)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def fft_frequency_decompose(x, min_size):
    """audio/transform.py:50-82.  x: (B, C, N) -> {size: (B, C, size)}"""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, C, N = x.shape
    sizes = []
    s = min_size
    while s <= N:
        sizes.append(s)
        s *= 2
    if not sizes:
        return {}
    out = {s: torch.empty((B, C, s), dtype=torch.float32, device=x.device) for s in sizes}
    ptrs = (ctypes.c_void_p * len(sizes))(*[out[s].data_ptr() for s in sizes])
    ws = _workspace(B * C, N, x.device)
    check(_lib.lib().ms_fft_frequency_decompose(ptr(x), B * C, N, min_size, ptrs, len(sizes),
                                                ptr(ws), ws.numel(), stream_ptr()),
          "ms_fft_frequency_decompose")
    return out


def fft_frequency_recompose(d, desired_size):
    """audio/transform.py:107-115.  d: {size: (B, C, size)} -> (B, C, desired_size)"""
    items = list(d.items())
    first = items[0][1]
    _lib.require_cuda(first, "band")
    B, C = first.shape[0], first.shape[1]
    keep = [b.contiguous() for _, b in items]
    ptrs = (ctypes.c_void_p * len(keep))(*[b.data_ptr() for b in keep])
    sizes = (ctypes.c_int * len(keep))(*[int(s) for s, _ in items])
    out = torch.empty((B, C, desired_size), dtype=torch.float32, device=first.device)
    ws = _workspace(B * C, desired_size, first.device)
    check(_lib.lib().ms_fft_frequency_recompose(ptrs, sizes, len(keep), B * C, desired_size,
                                                ptr(out), ptr(ws), ws.numel(), stream_ptr()),
          "ms_fft_frequency_recompose")
    return out
