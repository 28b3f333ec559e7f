"""Drop-in mirror of the FFT band split / merge in featuresynth/audio/transform.py:50-115.

`fft_frequency_decompose(x, min_size) -> {size: band}` (keys ascending, as the reference
builds them) and `fft_frequency_recompose(d, desired_size) -> tensor`, on CUDA tensors,
through `ms_fft_frequency_decompose / recompose` (hand-written Stockham FFT passes; the
reference uses the removed `torch.rfft / irfft`).

Both functions are differentiable (training with `decompose=True` / `recompose=True`,
generator/multiscale.py:166-178, discriminator/multiscale.py:212-252): with orthonormal
transforms the merge is the adjoint of the split and vice versa up to one bin per band, which
`ms_fft_decompose_adjoint_fix` / `ms_fft_recompose_adjoint_fix` correct in place (include/msb200.h).
"""
import ctypes

import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr


def _workspace(batch, n, device):
    nbytes = _lib.lib().ms_fft_bands_workspace_bytes(batch, n)
    return torch.empty(nbytes, dtype=torch.uint8, device=device)


def _band_sizes(n, min_size):
    sizes = []
    s = min_size
    while s <= n:
        sizes.append(s)
        s *= 2
    return sizes


class _Decompose(torch.autograd.Function):
    """x -> bands (ascending sizes); backward = band merge + rank-1 Nyquist terms."""

    @staticmethod
    def forward(ctx, x, min_size):
        out = _decompose_raw(x, min_size)
        ctx.n = x.shape[-1]
        ctx.sizes = tuple(out.keys())
        ctx.shape = tuple(x.shape)
        return tuple(out.values())

    @staticmethod
    def backward(ctx, *dbands):
        B, C, N = ctx.shape
        g = {}
        for s_, d_ in zip(ctx.sizes, dbands):
            g[s_] = (torch.zeros((B, C, s_), dtype=torch.float32, device=dbands_device(dbands))
                     if d_ is None else d_.contiguous())
        dx = _recompose_raw(g, N)
        keep = list(g.values())
        ptrs = (ctypes.c_void_p * len(keep))(*[b.data_ptr() for b in keep])
        sizes = (ctypes.c_int * len(keep))(*[int(s_) for s_ in g])
        check(_lib.lib().ms_fft_decompose_adjoint_fix(ptrs, sizes, len(keep), B * C, N, ptr(dx),
                                                      stream_ptr()), "ms_fft_decompose_adjoint_fix")
        return dx, None


def dbands_device(dbands):
    return next(d.device for d in dbands if d is not None)


class _Recompose(torch.autograd.Function):
    """bands -> waveform; backward = band split of the gradient + rank-1 Nyquist terms."""

    @staticmethod
    def forward(ctx, desired_size, sizes, *bands):
        ctx.sizes = tuple(int(s_) for s_ in sizes)
        ctx.desired = int(desired_size)
        return _recompose_raw(dict(zip(ctx.sizes, bands)), desired_size)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.contiguous()
        B, C, N = dy.shape
        dec = _decompose_raw(dy, min(ctx.sizes))
        keep = [dec[s_] for s_ in ctx.sizes]
        ptrs = (ctypes.c_void_p * len(keep))(*[b.data_ptr() for b in keep])
        sizes = (ctypes.c_int * len(keep))(*ctx.sizes)
        check(_lib.lib().ms_fft_recompose_adjoint_fix(ptr(dy), B * C, N, ptrs, sizes, len(keep),
                                                      stream_ptr()), "ms_fft_recompose_adjoint_fix")
        return (None, None) + tuple(keep)


def fft_frequency_decompose(x, min_size):
    """audio/transform.py:50-82.  x: (B, C, N) -> {size: (B, C, size)}"""
    if torch.is_grad_enabled() and x.requires_grad:
        sizes = _band_sizes(x.shape[-1], min_size)
        if not sizes:
            return {}
        return dict(zip(sizes, _Decompose.apply(x, min_size)))
    return _decompose_raw(x, min_size)


def fft_frequency_recompose(d, desired_size):
    """audio/transform.py:107-115.  d: {size: (B, C, size)} -> (B, C, desired_size)"""
    if torch.is_grad_enabled() and any(b.requires_grad for b in d.values()):
        return _Recompose.apply(desired_size, tuple(d.keys()), *d.values())
    return _recompose_raw(d, desired_size)


def _decompose_raw(x, min_size):
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


def _recompose_raw(d, desired_size):
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
