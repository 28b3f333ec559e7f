"""Drop-in mirror of featuresynth/discriminator/melgan.py:7-27 (MelGanDiscriminator):
ONE shared FullDiscriminator applied at 3 scales, `avg_pool1d(x, 4, 2, padding=2)`
(count_include_pad=True, lengths N/2+1) between scales.  State-dict keys `disc.*`."""
from torch import nn

import torch

from .. import autograd as ag
from .. import ops
from .full import FullDiscriminator


class MelGanDiscriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.disc = FullDiscriminator()
        self.scales = 2

    def forward(self, x, feat=None):
        """`feat` is accepted and ignored: the trainers call discriminator(audio, features)
        (featuresynth/train/train.py:30-31), this unconditioned discriminator only reads x."""
        features = []
        judgements = []
        f, j = self.disc(x)
        features.append(f)
        judgements.append(j)
        for _ in range(self.scales):
            if torch.is_grad_enabled() and x.requires_grad:
                x = ag.AvgPool.apply(x, 4, 2, 2, True)
            else:
                x = ops.avg_pool1d(x, 4, 2, 2)
            f, j = self.disc(x)
            features.append(f)
            judgements.append(j)
        return features, judgements
