"""Drop-in mirror of featuresynth/discriminator/melgan.py:7-27 (MelGanDiscriminator):
ONE shared FullDiscriminator applied at 3 scales, `avg_pool1d(x, 4, 2, padding=2)`
(count_include_pad=True, lengths N/2+1) between scales.  State-dict keys `disc.*`."""
from torch import nn

from .. import ops
from .full import FullDiscriminator


class MelGanDiscriminator(nn.Module):
    def __init__(self):
        super().__init__()
        self.disc = FullDiscriminator()
        self.scales = 2

    def forward(self, x):
        features = []
        judgements = []
        f, j = self.disc(x)
        features.append(f)
        judgements.append(j)
        for _ in range(self.scales):
            x = ops.avg_pool1d(x, 4, 2, 2)
            f, j = self.disc(x)
            features.append(f)
            judgements.append(j)
        return features, judgements
