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
