"""Drop-in mirror of featuresynth/discriminator/filterbank.py:114-202 (`FilterBankDiscriminator`):
the discriminator of FilterBankExperiment / ConditionalFilterBankExperiment
(experiment/filterbank.py:14-128).

Same constructor, state dict (`main.{0..7}`, `judge`, `medium_res.stack.main.{i}`,
`medium_res.judge`, `low_res...`) and forward contract: (x (B,1,N), feat (B,C,T)) ->
(features: [8 maps, 3 maps, 3 maps], judgements: [(B,1,N/256), (B,1,16), (B,1,4)]).

x -> 511-tap Morlet analysis bank (tcgen05: 16-wide sliding expansion, 32 taps of dilation 16)
-> three heads on the same analysis: eight stride-2 k7 convs (space-to-depth + stride-1 tcgen05
convs, up to 2048 -> 1024 channels) + judge; two LowResSpectrogramDiscriminators (relu + window
means, then stride-2 convs).  One path for training and inference: every block is an autograd
Function over the C ABI (see ../autograd.py).
"""
import torch
from torch import nn

from .. import autograd as ag
from .. import grad_ops, ops
from .._lib import MsbError
from ..util.modules import LowResSpectrogramDiscriminator, nearest_upsample


class FilterBankDiscriminator(nn.Module):
    def __init__(self, filter_bank, input_size, conditioning_channels=0, log_scaling=False):
        super().__init__()
        self.log_scaling = log_scaling
        self.conditioning_channels = conditioning_channels
        self.input_size = input_size
        self._filter_bank = [filter_bank]
        in_channels = self.filter_bank.n_bands
        self.main = nn.Sequential(
            nn.Conv1d(in_channels + conditioning_channels, 256, 7, 2, 3),
            nn.Conv1d(256, 256, 7, 2, 3),
            nn.Conv1d(256, 512, 7, 2, 3),
            nn.Conv1d(512, 512, 7, 2, 3),
            nn.Conv1d(512, 1024, 7, 2, 3),
            nn.Conv1d(1024, 1024, 7, 2, 3),
            nn.Conv1d(1024, 1024, 7, 2, 3),
            nn.Conv1d(1024, 1024, 7, 2, 3))
        self.judge = nn.Conv1d(1024, 1, 3, 1, 1)
        self.medium_res = LowResSpectrogramDiscriminator(
            freq_bins=128, time_steps=128, n_judgements=16, kernel_size=7, max_channels=1024,
            conditioning_channels=conditioning_channels, log_scaling=log_scaling)
        self.low_res = LowResSpectrogramDiscriminator(
            freq_bins=32, time_steps=32, n_judgements=4, kernel_size=7, max_channels=512,
            conditioning_channels=conditioning_channels, log_scaling=log_scaling)
        self._sc = [ag.StridedCache() for _ in self.main]

    @property
    def filter_bank(self):
        return self._filter_bank[0]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.filter_bank.to(fn(torch.zeros(1)).device)
        return out

    def full_resolution(self, a32, a16, feat):
        h32, h16 = a32, a16
        if self.conditioning_channels > 0:
            up = nearest_upsample(feat, a16.shape[2])
            h32 = torch.cat([h32, grad_ops.pack_ncl32(up)], dim=1)
            h16 = torch.cat([h16, ops.pack_ncl(up)], dim=1)
        features = []
        length = h16.shape[2]
        for conv, sc in zip(self.main, self._sc):
            h32, h16 = ag.StridedConvBlk.apply(h32, h16, conv.weight, conv.bias, sc, 2, length)
            features.append(ag.UnpackBlk32.apply(h32))
            length = h16.shape[2]
        return features, ag.MonoConv.apply(h32, self.judge.weight, self.judge.bias, 3, 1, False)

    def forward(self, x, feat):
        if x.dim() != 3 or x.shape[1] != 1:
            raise MsbError("expected (B, 1, N) audio")
        if self.conditioning_channels > 0 and feat is None:
            raise MsbError("this discriminator is conditioned: pass the features")
        a32, a16 = ag.BankAnalysis.apply(x.contiguous(), self.filter_bank)
        features, judgements = [], []
        for head in (lambda: self.full_resolution(a32, a16, feat),
                     lambda: self.medium_res.forward_blocked(a32, feat),
                     lambda: self.low_res.forward_blocked(a32, feat)):
            f, j = head()
            features.append(f)
            judgements.append(j)
        return features, judgements
