"""Helpers shared by the -m gpu parity tests."""
import numpy as np
import torch


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd16(t, operand="f16"):
    """Round to the 16-bit operand format and back (what the kernels see)."""
    return (t.half() if operand == "f16" else t.bfloat16()).float()


def randn(seed, *shape, scale=1.0):
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.standard_normal(shape) * scale).astype(np.float32))


def strided_weight_view_ref(w, stride):
    """index map of ms_strided_weight_view (forward): (Cout, C, k) -> (Cout, s*C, taps) with
    w1[co][i*C + c][j - jmin] = w[co][c][kk], kk - k//2 = s*j + i"""
    cout, c, k = w.shape
    half = k // 2
    j_min = -((half + stride - 1) // stride)
    taps = half // stride - j_min + 1
    out = torch.zeros((cout, stride * c, taps), dtype=w.dtype)
    for kk in range(k):
        m = kk - half
        j, i = m // stride, m % stride
        out[:, i * c:(i + 1) * c, j - j_min] = w[:, :, kk]
    return out
