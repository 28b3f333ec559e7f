"""Drop-in mirror of featuresynth/loss/loss.py:5-79 (forward values).

Same function names and signatures; each returns a 0-dim CUDA tensor.  The composite
losses accumulate every term into ONE device scalar through `ms_reduce_fwd` (deterministic
two-stage reductions, no atomics) instead of materialising 18-38 intermediate tensors.
Differentiable: the gradient of every term is one `ms_reduce_bwd` launch.
"""
import torch

from .. import _lib
from ..autograd import WeightedLoss

L1, HINGE_D, HINGE_G, LSQ_D, LSQ_G = 0, 1, 2, 3, 4


class _Acc:
    """Collects weighted reduction terms; `value()` evaluates them into ONE device scalar
    through `WeightedLoss` (forward: ms_reduce_fwd per term; backward: ms_reduce_bwd)."""

    def __init__(self, device=None):
        self.tensors = []
        self.index = {}
        self.spec = []

    def _slot(self, t):
        _lib.require_cuda(t, "loss operand")
        k = id(t)
        if k not in self.index:
            self.index[k] = len(self.tensors)
            self.tensors.append(t)
        return self.index[k]

    def add(self, mode, a, b=None, weight=1.0):
        if b is not None and b.numel() != a.numel():
            raise _lib.MsbError("loss operands differ in size")
        self.spec.append((mode, self._slot(a), self._slot(b) if b is not None else -1,
                          float(weight)))
        return self

    def value(self):
        return WeightedLoss.apply(tuple(self.spec), *self.tensors)


def _term(acc, mode, a, b, w):
    """sub-loss called stand-alone (returns its value) or as a term of a composite loss"""
    if acc is not None:
        acc.add(mode, a, b, w)
        return None
    return _Acc().add(mode, a, b, w).value()


def least_squares_generator_loss(j, _acc=None, _w=1.0):
    """loss.py:5-6"""
    return _term(_acc, LSQ_G, j, None, _w)


def hinge_generator_loss(j, _acc=None, _w=1.0):
    """loss.py:9-10"""
    return _term(_acc, HINGE_G, j, None, _w)


def least_squares_disc_loss(r_j, f_j, _acc=None, _w=1.0):
    """loss.py:13-14"""
    return _term(_acc, LSQ_D, r_j, f_j, _w)


def hinge_discriminator_loss(r_j, f_j, _acc=None, _w=1.0):
    """loss.py:17-18"""
    return _term(_acc, HINGE_D, r_j, f_j, _w)


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
