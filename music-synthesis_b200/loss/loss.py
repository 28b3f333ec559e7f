"""Drop-in mirror of featuresynth/loss/loss.py:5-79 (forward values).

Same function names and signatures; each returns a 0-dim CUDA tensor.  The composite
losses accumulate every term into ONE device scalar through `ms_reduce_fwd` (deterministic
two-stage reductions, no atomics) instead of materialising 18-38 intermediate tensors.
Forward only in this round (no autograd).
"""
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr

L1, HINGE_D, HINGE_G, LSQ_D, LSQ_G = 0, 1, 2, 3, 4


class _Acc:
    """A device scalar that sums weighted reductions."""

    def __init__(self, device):
        self.out = torch.zeros(1, dtype=torch.float32, device=device)
        self.ws = torch.empty(_lib.lib().ms_reduce_workspace_bytes(), dtype=torch.uint8,
                              device=device)

    def add(self, mode, a, b=None, weight=1.0):
        _lib.require_cuda(a, "a")
        a = a.contiguous()
        if b is not None:
            _lib.require_cuda(b, "b")
            b = b.contiguous()
            if b.numel() != a.numel():
                raise _lib.MsbError("loss operands differ in size")
        check(_lib.lib().ms_reduce_fwd(mode, ptr(a), ptr(b), a.numel(), float(weight),
                                       ptr(self.out), 1, ptr(self.ws), stream_ptr()),
              "ms_reduce_fwd")
        return self

    def value(self):
        return self.out.reshape(())


def least_squares_generator_loss(j, _acc=None, _w=1.0):
    """loss.py:5-6"""
    return (_acc or _Acc(j.device)).add(LSQ_G, j, None, _w).value()


def hinge_generator_loss(j, _acc=None, _w=1.0):
    """loss.py:9-10"""
    return (_acc or _Acc(j.device)).add(HINGE_G, j, None, _w).value()


def least_squares_disc_loss(r_j, f_j, _acc=None, _w=1.0):
    """loss.py:13-14"""
    return (_acc or _Acc(r_j.device)).add(LSQ_D, r_j, f_j, _w).value()


def hinge_discriminator_loss(r_j, f_j, _acc=None, _w=1.0):
    """loss.py:17-18"""
    return (_acc or _Acc(r_j.device)).add(HINGE_D, r_j, f_j, _w).value()


def mel_gan_disc_loss(real_judgements, fake_judgements, gan_loss=hinge_discriminator_loss):
    """loss.py:21-25"""
    acc = _Acc(real_judgements[0].device)
    for r, f in zip(real_judgements, fake_judgements):
        gan_loss(r, f, _acc=acc)
    return acc.value()


def _feature_terms(acc, real_features, fake_features, scale):
    nd = 1 / len(real_features)
    for r_group, f_group in zip(real_features, fake_features):
        nl = 1 / len(r_group)
        for r_f, f_f in zip(r_group, f_group):
            acc.add(L1, r_f, f_f, scale * nl * nd)


def mel_gan_feature_loss(real_features, fake_features):
    """loss.py:28-65: sum over discriminators / layers of (1/n_disc)(1/n_layers) mean|r-f|."""
    acc = _Acc(real_features[0][0].device)
    _feature_terms(acc, real_features, fake_features, 1.0)
    return acc.value()


def mel_gan_gen_loss(real_features, fake_features, real_judgements, fake_judgements,
                     gan_loss=hinge_generator_loss, feature_loss_weight=10):
    """loss.py:68-79"""
    acc = _Acc(fake_judgements[0].device)
    for _, f in zip(real_judgements, fake_judgements):
        gan_loss(f, _acc=acc)
    _feature_terms(acc, real_features, fake_features, float(feature_loss_weight))
    return acc.value()
