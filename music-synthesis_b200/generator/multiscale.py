"""Drop-in mirrors of featuresynth/generator/multiscale.py:59-178:
`FilterBankChannelGenerator` and `FilterBankMultiScaleGenerator`.

Same constructor signatures, forward contract (x (B,128,T) -> {band_size: (B,1,band_size)}
or, with recompose=True, the summed full-rate waveform) and state-dict keys
(`channel_{size}.main.0.0.{weight,bias}`, `channel_{size}.main.{1..4}.conv.weight`; the
fixed Morlet banks are plain attributes, not in the state dict).

Per band the whole chain stays channel-blocked on the tcgen05 path: Conv1d(128,128,k7)+LReLU
-> 4 x [ConvTranspose1d(k=2s, stride s, no bias)+LReLU] (polyphase implicit GEMMs) ->
filter-bank synthesis (128 -> 8 phase channels, 16 taps, + anti-diagonal sum).
Forward (inference) only in this round.
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from .._lib import MS_CONV, MS_CONVT, MS_F16, MsbError
from ..audio.filterbank import FilterBank, SampleRate, linear_center_frequencies
from ..audio.transform import fft_frequency_recompose
from ..util.modules import LearnedUpSample, _PackedConv


def _as_samplerate(sr):
    if isinstance(sr, SampleRate):
        return sr
    if hasattr(sr, "nyquist") and hasattr(sr, "__int__"):      # e.g. a zounds SampleRate
        return SampleRate(float(getattr(sr, "rate", None) or 2.0 * float(sr.nyquist)))
    return SampleRate(float(sr))


class FilterBankChannelGenerator(nn.Module):
    def __init__(self, scale_factors, channels, filter_bank, operand=MS_F16):
        super().__init__()
        self.filter_bank = filter_bank
        self.channels = channels
        self.scale_factors = scale_factors
        self.operand = operand
        layers = []
        for i in range(len(scale_factors)):
            if i == 0:
                layers.append(nn.Sequential(
                    nn.Conv1d(channels[i], channels[i + 1], 7, 1, 3), nn.LeakyReLU(0.2)))
            else:
                layers.append(LearnedUpSample(
                    in_channels=channels[i], out_channels=channels[i + 1],
                    kernel_size=scale_factors[i] * 2, scale_factor=scale_factors[i],
                    activation=None, operand=operand))
        self.main = nn.Sequential(*layers)
        self._packed = [_PackedConv() for _ in layers]

    def forward_blocked(self, x16, T):
        """x16: BLK 16-bit (B, Cin/8, T, 8) features (shared by all bands)."""
        B = x16.shape[0]
        emb = self.main[0][0]
        d = ops.conv_desc(MS_CONV, B, emb.in_channels, emb.out_channels, T, 7, 1, 3, leaky=True,
                          operand=self.operand)
        h16, _ = ops.conv_fwd(d, x16, self._packed[0].get(d, emb.weight), emb.bias)
        L = T
        for i in range(1, len(self.main)):
            up = self.main[i]
            s = up.scale_factor
            d = ops.conv_desc(MS_CONVT, B, up.in_channels, up.out_channels, L, 2 * s, 1,
                              (2 * s - s) // 2, s, leaky=True, operand=self.operand)
            h16, _ = ops.conv_fwd(d, h16, self._packed[i].get(d, up.conv.weight), None)
            L *= s
        # F.pad(x, (0, 1)) + transposed_convolve, generator/multiscale.py:90-91
        return self.filter_bank.transposed_convolve_blocked(h16, L)

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise MsbError("sm_100a path is forward-only in this build: use torch.no_grad()")
        return self.forward_blocked(ops.pack_ncl(x, operand=self.operand), x.shape[-1])


class FilterBankMultiScaleGenerator(nn.Module):
    def __init__(self, samplerate, feature_channels, input_size, output_size, recompose=True):
        super().__init__()
        self.samplerate = samplerate
        self.recompose = recompose
        self.output_size = output_size
        self.input_size = input_size
        self.feature_channels = feature_channels
        band_sizes = [int(2 ** (np.log2(output_size) - i)) for i in range(5)]
        self.upsample_ratio = output_size // input_size
        sr = _as_samplerate(samplerate)
        factors = ([1, 4, 4, 4, 4], [1, 4, 4, 4, 2], [1, 4, 4, 2, 2], [1, 4, 2, 2, 2],
                   [1, 2, 2, 2, 2])
        self.spec = {}
        self.channel_generators = {}
        for i, (size, sf) in enumerate(zip(band_sizes, factors)):
            self.spec[size] = {
                "scale_factors": sf,
                "channels": [feature_channels] + [128] * 5 if feature_channels != 128 else [128] * 6,
                "filter_bank": self._filter_bank(sr * (2 ** i), 128, zero_start=(i == 4)),
            }
            gen = FilterBankChannelGenerator(**self.spec[size])
            self.add_module(f"channel_{size}", gen)
            self.channel_generators[size] = gen

    def _filter_bank(self, samplerate, bands, zero_start=False):
        start = 0 if zero_start else samplerate.nyquist / 2
        return FilterBank(samplerate, 128,
                          linear_center_frequencies(start, samplerate.nyquist, bands),
                          scaling_factors=0.05, normalize_filters=True, a_weighting=False)

    def _apply(self, fn, *args, **kwargs):
        # the banks are plain attributes: move them with the module (.to / .cuda)
        out = super()._apply(fn, *args, **kwargs)
        probe = fn(torch.zeros(1))
        for gen in self.channel_generators.values():
            gen.filter_bank.to(probe.device)
        return out

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise MsbError("sm_100a path is forward-only in this build: use torch.no_grad()")
        input_size = x.shape[-1]
        x16 = ops.pack_ncl(x)
        results = {}
        for size, layer in self.channel_generators.items():
            results[size] = layer.forward_blocked(x16, input_size)
        if self.recompose:
            return fft_frequency_recompose(results, input_size * self.upsample_ratio)
        return results
