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

from .. import autograd as ag
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
    #: weight-split forward ([x, x] * [W_hi, W_lo], include/msb200.h): each of the six layers of a
    #: band contributes ~3.9e-4 of forward error with both operands rounded to fp16 -- 9.5e-4 to
    #: 1.08e-3 in quadrature, ON the 1e-3 bar -- and ~2.8e-4 with exact weights (6.9e-4 total)
    weight_split = True

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
        self._cache = [ag.WeightCache() for _ in layers]

    def forward_blocked_train(self, x16):
        """autograd-recorded form (training): every block is a Function over the C ABI"""
        emb = self.main[0][0]
        ws = self.weight_split
        h32, h16 = ag.conv_blk(None, x16, emb.weight, emb.bias, self._cache[0], MS_CONV, 1, 3,
                               1, True, None, ws)
        for i in range(1, len(self.main)):
            up = self.main[i]
            s = up.scale_factor
            h32, h16 = ag.conv_blk(h32, h16, up.conv.weight, None, self._cache[i], MS_CONVT, 1,
                                   s // 2, s, True, None, ws)
        return ag.BankSynthesis.apply(h32, h16, self.filter_bank, ws)

    def forward_blocked(self, x16, T):
        """x16: BLK 16-bit (B, Cin/8, T, 8) features (shared by all bands)."""
        B = x16.shape[0]
        emb = self.main[0][0]
        ws = self.weight_split and self.operand == MS_F16 and not ops.relaxed()
        mult, alpha = (2, 1.0 / ops.W_SPLIT_SCALE) if ws else (1, 1.0)

        def run(i, d, h16, w, bias, kind):
            if ws:       # the K loop wraps over h16 twice (x_repeat): [x, x] * [W_hi, W_lo]
                return ops.conv_fwd(d, h16, self._cache[i].fwd_wsplit(d, w, kind), bias)[0]
            return ops.conv_fwd(d, h16, self._packed[i].get(d, w), bias)[0]

        d = ops.conv_desc(MS_CONV, B, mult * emb.in_channels, emb.out_channels, T, 7, 1, 3,
                          leaky=True, operand=self.operand, alpha=alpha, x_repeat=mult)
        h16 = run(0, d, x16, emb.weight, emb.bias, MS_CONV)
        L = T
        for i in range(1, len(self.main)):
            up = self.main[i]
            s = up.scale_factor
            d = ops.conv_desc(MS_CONVT, B, mult * up.in_channels, up.out_channels, L, 2 * s, 1,
                              (2 * s - s) // 2, s, leaky=True, operand=self.operand, alpha=alpha,
                              x_repeat=mult)
            h16 = run(i, d, h16, up.conv.weight, None, MS_CONVT)
            L *= s
        # F.pad(x, (0, 1)) + transposed_convolve, generator/multiscale.py:90-91
        return self.filter_bank.transposed_convolve_blocked(h16, L, ws)

    def forward(self, x):
        if ag.needs_grad(self, x):
            if x.requires_grad:
                raise MsbError("gradients w.r.t. the conditioning features are not on this path")
            return self.forward_blocked_train(ops.pack_ncl(x))
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
        train = ag.needs_grad(self, x)
        if train and x.requires_grad:
            raise MsbError("gradients w.r.t. the conditioning features are not on this path")
        input_size = x.shape[-1]
        x16 = ops.pack_ncl(x)
        results = {}
        for size, layer in self.channel_generators.items():
            results[size] = layer.forward_blocked_train(x16) if train else \
                layer.forward_blocked(x16, input_size)
        if self.recompose:
            return fft_frequency_recompose(results, input_size * self.upsample_ratio)
        return results


# ---------------------------------------------------------------------------------------
# Non-filterbank multiscale generator, featuresynth/generator/multiscale.py:10-57, 180-251
# ---------------------------------------------------------------------------------------
class DilatedStack(nn.Module):
    """featuresynth/util/modules.py:79-139 as configured by ChannelGenerator: bias-free k3 convs
    with zero padding = dilation, x <- LeakyReLU(conv_d(x) + x) (activation AFTER the residual
    add: epilogue mode 2 of the tcgen05 conv kernel)."""

    def __init__(self, in_channels, channels, kernel_size, dilations, activation=None,
                 residual=True, groups=None, reflection_padding=False):
        super().__init__()
        if kernel_size != 3 or groups is not None or reflection_padding or in_channels != channels \
                or not residual:
            raise NotImplementedError("DilatedStack: only the ChannelGenerator configuration")
        self.in_channels = in_channels
        self.channels = channels
        self.kernel_size = kernel_size
        self.dilations = dilations
        self.residual = residual
        self.main = nn.Sequential(*[
            nn.Conv1d(channels, channels, kernel_size, padding=d, dilation=d, bias=False)
            for d in dilations])
        self._packed = [_PackedConv() for _ in dilations]
        self._cache = [ag.WeightCache() for _ in dilations]

    def forward_blocked_train(self, x32, x16):
        for conv, cache, d in zip(self.main, self._cache, self.dilations):
            x32, x16 = ag.DilatedLayerBlk.apply(x32, x16, conv.weight, cache, d)
        return x32, x16

    def forward_blocked(self, x16, x32):
        B, _, L, _ = x16.shape
        for conv, pk, d in zip(self.main, self._packed, self.dilations):
            desc = ops.conv_desc(MS_CONV, B, self.channels, self.channels, L, 3, d, d, leaky=2)
            x16, x32 = ops.conv_fwd(desc, x16, pk.get(desc, conv.weight), None, res32=x32,
                                    want16=True, want32=True)
        return x16, x32


class ChannelGenerator(nn.Module):
    """generator/multiscale.py:10-57 (transposed_conv=True, the configuration of every experiment
    in experiment/multiscale.py): 4 x [LearnedUpSample + DilatedStack] then Conv1d(C,1,7,1,3)."""

    def __init__(self, scale_factors, channels, transposed_conv=False, kernel_size=40):
        super().__init__()
        if not transposed_conv:
            raise NotImplementedError("ChannelGenerator: nearest-neighbour UpSample variant is "
                                      "not on this path (the experiments use transposed_conv=True)")
        self.kernel_size = kernel_size
        self.transposed_conv = transposed_conv
        self.channels = channels
        self.scale_factors = scale_factors
        layers = []
        for i in range(len(scale_factors)):
            layers.append(LearnedUpSample(in_channels=channels[i], out_channels=channels[i + 1],
                                          kernel_size=scale_factors[i] * 2,
                                          scale_factor=scale_factors[i], activation=None))
            layers.append(DilatedStack(channels[i + 1], channels[i + 1], 3, [1, 3, 9]))
        self.main = nn.Sequential(*layers)
        self.to_samples = nn.Conv1d(channels[-1], 1, 7, 1, 3)
        self._packed = [_PackedConv() for _ in scale_factors]
        self._cache = [ag.WeightCache() for _ in scale_factors]

    def forward_blocked_train(self, e32, e16):
        """autograd-recorded form; (e32, e16) = the shared embedding"""
        h32, h16 = e32, e16
        for i in range(len(self.scale_factors)):
            up, stack = self.main[2 * i], self.main[2 * i + 1]
            s = up.scale_factor
            h32, h16 = ag.conv_blk(h32, h16, up.conv.weight, None, self._cache[i], MS_CONVT, 1,
                                        s // 2, s, True)
            h32, h16 = stack.forward_blocked_train(h32, h16)
        return ag.MonoConv.apply(h32, self.to_samples.weight, self.to_samples.bias, 7, 3, False)

    def forward_blocked(self, x16, T):
        B = x16.shape[0]
        L = T
        h16, h32 = x16, None
        for i in range(len(self.scale_factors)):
            up, stack = self.main[2 * i], self.main[2 * i + 1]
            s = up.scale_factor
            d = ops.conv_desc(MS_CONVT, B, up.in_channels, up.out_channels, L, 2 * s, 1, s // 2, s,
                              leaky=True)
            h16, h32 = ops.conv_fwd(d, h16, self._packed[i].get(d, up.conv.weight), None,
                                    want16=True, want32=True)
            L *= s
            h16, h32 = stack.forward_blocked(h16, h32)
        return ops.conv_to_mono(h32, self.to_samples.weight, self.to_samples.bias, 7, 3, False)

    def forward(self, x):
        if ag.needs_grad(self, x):
            return self.forward_blocked_train(ag.PackBlk32.apply(x), ops.pack_ncl(x.detach()))
        return self.forward_blocked(ops.pack_ncl(x), x.shape[-1])


def _fwd_only_g(module, x):
    if torch.is_grad_enabled() and (x.requires_grad or
                                    any(p.requires_grad for p in module.parameters())):
        raise MsbError("this model has no backward on the sm_100a path yet: use torch.no_grad()")


class MultiScaleGenerator(nn.Module):
    """generator/multiscale.py:180-251: ReflectionPad1d(3) + Conv1d(C,512,7) + LeakyReLU embedding
    shared by five ChannelGenerators (one per octave band, 512->256->128->64->32 channels);
    state-dict keys `embedding.*`, `channel_{size}.main.{0,2,4,6}.conv.weight`,
    `channel_{size}.main.{1,3,5,7}.main.{0,1,2}.weight`, `channel_{size}.to_samples.*`."""

    def __init__(self, feature_channels, input_size, output_size, transposed_conv=False,
                 recompose=True, kernel_size=40):
        super().__init__()
        self.kernel_size = kernel_size
        self.recompose = recompose
        self.input_size = input_size
        self.output_size = output_size
        self.feature_channels = feature_channels
        self.embedding = nn.Conv1d(feature_channels, 512, 7, 1, padding=0)
        self.upsample_ratio = output_size // input_size
        band_sizes = [int(2 ** (np.log2(output_size) - i)) for i in range(5)]
        factors = ([4, 4, 4, 4], [4, 4, 4, 2], [4, 4, 2, 2], [4, 2, 2, 2], [2, 2, 2, 2])
        self.spec = {bs: {"scale_factors": sf, "channels": [512, 256, 128, 64, 32]}
                     for bs, sf in zip(band_sizes, factors)}
        self.channel_generators = {}
        for key, value in self.spec.items():
            generator = ChannelGenerator(**value, transposed_conv=transposed_conv)
            self.add_module(f"channel_{key}", generator)
            self.channel_generators[key] = generator
        self._pe = _PackedConv()
        self._ce = ag.WeightCache()

    def _forward_train(self, x):
        if x.requires_grad:
            raise MsbError("gradients w.r.t. the conditioning features are not on this path")
        x16 = ops.pack_ncl(x, 3, 1)
        e32, e16 = ag.conv_blk(None, x16, self.embedding.weight, self.embedding.bias, self._ce,
                                    MS_CONV, 1, 0, 1, True)
        results = {size: layer.forward_blocked_train(e32, e16)
                   for size, layer in self.channel_generators.items()}
        if self.recompose:      # differentiable band merge (generator/multiscale.py:248-251)
            return fft_frequency_recompose(results, x.shape[-1] * self.upsample_ratio)
        return results

    def forward(self, x):
        if ag.needs_grad(self, x):
            return self._forward_train(x.contiguous())
        B, _, T = x.shape
        x16 = ops.pack_ncl(x, 3, 1)                                   # ReflectionPad1d(3)
        d = ops.conv_desc(MS_CONV, B, self.feature_channels, 512, T + 6, 7, 1, 0, leaky=True)
        e16, _ = ops.conv_fwd(d, x16, self._pe.get(d, self.embedding.weight), self.embedding.bias)
        results = {}
        for size, layer in self.channel_generators.items():
            results[size] = layer.forward_blocked(e16, T)
        if self.recompose:
            return fft_frequency_recompose(results, T * self.upsample_ratio)
        return results
