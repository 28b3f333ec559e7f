"""Drop-in mirror of featuresynth/discriminator/melgan.py:7-27 (MelGanDiscriminator):
ONE shared FullDiscriminator applied at 3 scales, `avg_pool1d(x, 4, 2, padding=2)`
(count_include_pad=True, lengths N/2+1) between scales.  State-dict keys `disc.*`."""
import torch
from torch import nn

from .. import autograd as ag
from .. import ops
from .full import FullDiscriminator

POOL = dict(ksize=4, stride=2, pad=2)      # between scales


def _downsample(x):
    """average pooling through the C ABI; recorded on the tape only when a gradient can flow"""
    if torch.is_grad_enabled() and x.requires_grad:
        return ag.AvgPool.apply(x, POOL["ksize"], POOL["stride"], POOL["pad"], True)
    return ops.avg_pool1d(x, POOL["ksize"], POOL["stride"], POOL["pad"])


class MelGanDiscriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.disc = FullDiscriminator()
        self.scales = 2                     # extra scales after the full-rate one

    def forward(self, x, feat=None):
        """-> (features per scale, judgement per scale).  `feat` is accepted and ignored: the
        trainers call discriminator(audio, features) (featuresynth/train/train.py:30-31), this
        unconditioned discriminator only reads x."""
        per_scale = [self.disc(x)]
        for _ in range(self.scales):
            x = _downsample(x)
            per_scale.append(self.disc(x))
        return [f for f, _ in per_scale], [j for _, j in per_scale]
