"""Drop-in mirror of featuresynth/generator/filterbank.py:93-128 (`FilterBankGenerator`): the
generator of FilterBankExperiment ("probably the best audio quality yet",
experiment/filterbank.py:14-78).

Same constructor, forward contract (x (B, in_channels, in_size) -> (B, 1, out_size)) and state
dict (`main.main.{i}.conv.weight`, `to_frames.{weight,bias}`; the fixed bank is held in a list so
that it is neither a sub-module nor in the state dict, generator/filterbank.py:100,119-125).

log2(out/in) x [ConvTranspose1d(C, 256, k 8, stride 2, padding 3, no bias) + LeakyReLU] as
four-tap polyphase implicit GEMMs, Conv1d(256, n_bands, 7), then the 511-tap bank synthesis
(16 phase channels x 32 taps of dilation 16 + anti-diagonal sum), all on the tcgen05 conv kernel.

Precision: ten fp16-operand layers in a row would put ~1.2e-3 of forward error on the waveform
(3.9e-4 per layer in quadrature, see generator/multiscale.py).  The first `exact_layers`
upsamplers run in full split precision (their sequences are short: 3 % of the FLOPs), the rest in
weight-split form (~2.3e-4 each); the bank synthesis -- a heavily cancelling sum, 1.9e-3 if
only its input is rounded to fp16 -- runs in full split precision.
"""
import torch
from torch import nn

from .. import autograd as ag
from .. import ops
from .._lib import MS_CONV, MS_CONVT, MsbError
from ..util.modules import (LearnedUpSample, ResidualStack, UpsamplingStack, layer_weight,
                            weight_norm)


class ResidualStackFilterBankGenerator(nn.Module):
    """featuresynth/generator/filterbank.py:8-90 (generator of AlternateFilterBankExperiment,
    experiment/filterbank.py:131-206): the MelGAN trunk at 256 channels built from (weight-normed)
    ConvTranspose1d layers -- the k 7 / stride 1 ones included -- and weight-normed ResidualStacks,
    then a harmonic head (`to_frames` -> filter-bank synthesis) plus a noise head (`to_noise`
    times the bank analysis of one white-noise row, summed over the bands).

    Same constructor and state dict (`main.{0,2,5,8,11}.{bias,weight_g,weight_v}`,
    `main.{4,7,10,13}.main.{a}.main.{c}...`, `to_frames...`, `to_noise...`).  The noise row is drawn
    exactly as the reference draws it -- `torch.normal(0, 1, (1, 1, time))` on the host generator,
    then moved -- so seeded runs agree."""

    def __init__(self, filter_bank, in_size, out_size, in_channels, add_weight_norm=True):
        super().__init__()
        self._filter_bank = [filter_bank]      # plain attribute in the reference, not a Module
        self.in_channels = in_channels
        self.out_size = out_size
        self.in_size = in_size
        self.add_weight_norm = add_weight_norm
        self.main = nn.Sequential(
            self._conv_layer(in_channels, 512, 7, 1, 3),
            nn.LeakyReLU(0.2),
            self._conv_layer(512, 256, 16, 8, 4),
            nn.LeakyReLU(0.2),
            ResidualStack(256, [1, 3, 9], add_weight_norm),
            self._conv_layer(256, 256, 16, 8, 4),
            nn.LeakyReLU(0.2),
            ResidualStack(256, [1, 3, 9], add_weight_norm),
            self._conv_layer(256, 256, 4, 2, 1),
            nn.LeakyReLU(0.2),
            ResidualStack(256, [1, 3, 9], add_weight_norm),
            self._conv_layer(256, 256, 4, 2, 1),
            nn.LeakyReLU(0.2),
            ResidualStack(256, [1, 3, 9], add_weight_norm),
        )
        self.to_frames = self._conv_layer(256, 128, 7, 1, 3)
        self.to_noise = self._conv_layer(256, 128, 7, 1, 3)
        self._cache = {}

    @property
    def filter_bank(self):
        return self._filter_bank[0]

    def _conv_layer(self, *args, **kwargs):
        conv = nn.ConvTranspose1d(*args, **kwargs)
        return weight_norm(conv) if self.add_weight_norm else conv

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.filter_bank.to(fn(torch.zeros(1)).device)
        return out

    def _convt(self, name, m, h32, h16, leaky):
        cache = self._cache.setdefault(name, ag.WeightCache())
        return ag.conv_blk(h32, h16, layer_weight(m), m.bias, cache, MS_CONVT, 1, m.padding[0],
                           m.stride[0], leaky)

    def forward(self, x):
        if x.requires_grad:
            raise MsbError("gradients w.r.t. the conditioning features are not on this path")
        h32, h16 = None, ops.pack_ncl(x)
        for i, m in enumerate(self.main):
            if isinstance(m, nn.ConvTranspose1d):
                h32, h16 = self._convt(i, m, h32, h16, True)      # + the LeakyReLU that follows
            elif isinstance(m, ResidualStack):
                h32, h16 = m.forward_blocked_train(h32, h16)
        time = h16.shape[2]
        n32, _ = self._convt("noise", self.to_noise, h32, h16, False)
        raw_noise = torch.normal(0, 1, (1, 1, time)).to(x.device)
        filtered = self.filter_bank._analysis(raw_noise, False, True)[1]     # BLK f32 (1,n/8,time,8)
        f32, f16 = self._convt("frames", self.to_frames, h32, h16, False)
        harmonic = ag.BankSynthesis.apply(f32, f16, self.filter_bank, 2)
        return ag.NoiseMix.apply(n32, filtered, harmonic)


class FilterBankGenerator(nn.Module):
    #: upsamplers (from the input side) that run in full split precision
    exact_layers = 6

    def __init__(self, filter_bank, in_size, out_size, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.out_size = out_size
        self.in_size = in_size
        self.channels = 256
        self._filter_bank = [filter_bank]
        self.main = UpsamplingStack(self.in_size, self.out_size, 2, self._build_layer)
        self.to_frames = nn.Conv1d(self.channels, self.filter_bank.n_bands, 7, 1, 3)
        self._cache = [ag.WeightCache() for _ in range(len(self.main.main) + 1)]

    def _build_layer(self, i, curr_size, out_size, first, last):
        return LearnedUpSample(self.in_channels if first else self.channels, self.channels, 8, 2,
                               None)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self.filter_bank.to(fn(torch.zeros(1)).device)
        return out

    @property
    def filter_bank(self):
        return self._filter_bank[0]

    def forward(self, x):
        if x.requires_grad:
            raise MsbError("gradients w.r.t. the conditioning features are not on this path")
        layers = list(self.main.main)
        n_exact = min(self.exact_layers, len(layers) - 1)
        h32 = ag.PackBlk32.apply(x) if n_exact > 0 else None
        h16 = ops.pack_ncl(x)
        for i, up in enumerate(layers):
            k, s = up.kernel_size, up.scale_factor
            h32, h16 = ag.conv_blk(h32, h16, up.conv.weight, None, self._cache[i], MS_CONVT, 1,
                                   (k - s) // 2, s, True, None, 2 if i < n_exact else 1)
        tf = self.to_frames
        h32, h16 = ag.conv_blk(h32, h16, tf.weight, tf.bias, self._cache[-1], MS_CONV, 1, 3, 1,
                               False, None, 1)
        return ag.BankSynthesis.apply(h32, h16, self.filter_bank, 2)
