"""Drop-in mirrors of featuresynth/discriminator/multiscale.py:70-252:
`FilterBankChannelDiscriminator` and `FilterBankMultiScaleDiscriminator`.

Same constructors, state-dict keys (`channel_{size}.main.{0..3}`, `.mj.{0..2}`, `.judge`,
`final.{0..2}`, `judge`) and forward contract: (x: {size: (B,1,size)} | (B,1,N), feat (B,C,T))
-> (features: 5 x [7 maps] + [3 maps], judgements: 6 x (B,1,T)).

Everything runs channel-blocked on the tcgen05 kernel: the Morlet analysis bank (sliding-
window expansion -> 16->128 ch, 8 taps, dilation 16), the stride-s k7 convs as stride-1 convs
over the space-to-depth input (s*128 channels, 2 or 4 taps), the k3 convs directly; the 1-ch
judges use the fp32 CUDA-core conv.  Every layer also emits its fp32 feature map, converted
to the reference's NCL layout because the feature-matching loss consumes them.
Forward (inference) only in this round.
"""
import numpy as np
import torch
from torch import nn

from .. import autograd as ag
from .. import grad_ops, ops
from .._lib import MS_CONV, MS_F16, MsbError
from ..audio.filterbank import FilterBank, linear_center_frequencies
from ..audio.transform import fft_frequency_decompose
from ..generator.multiscale import _as_samplerate
from ..util.modules import _PackedConv


class _PackedStrided:
    """packed weights of a stride-s conv rewritten as a stride-1 conv over space-to-depth"""

    def __init__(self):
        self.key = None
        self.val = None

    def get(self, weight, stride, make_desc):
        key = (weight.data_ptr(), weight._version, stride)
        if key != self.key:
            w1, taps, pad = ops.strided_conv_weight(weight.detach(), stride)
            self.val = (w1, taps, pad, None)
            self.key = key
        w1, taps, pad, packed = self.val
        d = make_desc(taps, pad)
        if packed is None:
            packed = ops.pack_conv_weight(d, w1)
            self.val = (w1, taps, pad, packed)
        return d, packed


def _k3(conv, packed, x16, leaky=True, operand=MS_F16):
    B, _, L, _ = x16.shape
    d = ops.conv_desc(MS_CONV, B, conv.in_channels, conv.out_channels, L, 3, 1, 1, leaky=leaky,
                      operand=operand)
    return ops.conv_fwd(d, x16, packed.get(d, conv.weight), conv.bias, want16=True, want32=True)


class FilterBankChannelDiscriminator(nn.Module):
    def __init__(self, scale_factors, channels, filter_bank, conditioning_channels=0,
                 operand=MS_F16):
        super().__init__()
        self.conditioning_channels = conditioning_channels
        self.filter_bank = filter_bank
        self.channels = channels
        self.scale_factors = scale_factors
        self.kernel_size = 7
        self.operand = operand
        self.main = nn.Sequential(*[
            nn.Conv1d(channels[i], channels[i + 1], self.kernel_size, scale_factors[i],
                      padding=self.kernel_size // 2) for i in range(len(scale_factors))])
        start = channels[-1] + (conditioning_channels if conditioning_channels > 0 else 0)
        self.mj = nn.Sequential(
            nn.Conv1d(start, channels[-1], 3, 1, 1),
            nn.Conv1d(channels[-1], channels[-1], 3, 1, 1),
            nn.Conv1d(channels[-1], channels[-1], 3, 1, 1))
        self.judge = nn.Conv1d(channels[-1], 1, 3, 1, 1)
        self._pm = [_PackedStrided() for _ in self.main]
        self._pj = [_PackedConv() for _ in self.mj]
        self._sc = [ag.StridedCache() for _ in self.main]
        self._cj = [ag.WeightCache() for _ in self.mj]

    def forward_blocked_train(self, x, feat32, feat16):
        """autograd-recorded form: x (B,1,L) (may require grad); conditioning as BLK f32 / f16.
        Returns ([7 NCL maps], h32, h16 of the last map, judgement)."""
        L = x.shape[-1]
        features = []
        if x.requires_grad:
            h32, h16 = ag.BankAnalysis.apply(x, self.filter_bank)
        else:
            h16, h32 = self.filter_bank._analysis(x, True, False)
        length = L
        for conv, sc in zip(self.main, self._sc):
            s = conv.stride[0]
            h32, h16 = ag.StridedConvBlk.apply(h32, h16, conv.weight, conv.bias, sc, s, length)
            features.append(ag.UnpackBlk32.apply(h32))
            length = h16.shape[2]
        if self.conditioning_channels > 0:
            h32 = torch.cat([h32, feat32], dim=1)
            h16 = torch.cat([h16, feat16], dim=1)
        for conv, cj in zip(self.mj, self._cj):
            h32, h16 = ag.conv_blk(h32, h16, conv.weight, conv.bias, cj, MS_CONV, 1, 1, 1, True)
            features.append(ag.UnpackBlk32.apply(h32))
        j = ag.MonoConv.apply(h32, self.judge.weight, self.judge.bias, 3, 1, False)
        return features, h32, h16, j

    def forward_blocked(self, x, feat16):
        """x (B,1,L) f32; feat16 BLK 16-bit conditioning or None.
        Returns ([7 NCL f32 maps], x16, x32 of the last map, judgement)."""
        B, _, L = x.shape
        features = []
        # analysis bank: (B,128,L+1), sliced to L by reading only L rows below
        a16 = self.filter_bank.convolve_blocked(x)
        h16, length = a16, L
        for conv, pm in zip(self.main, self._pm):
            s = conv.stride[0]
            xs = ops.space_to_depth(h16, s, length)
            lx = xs.shape[2]

            def make(taps, pad, B=B, conv=conv, s=s, lx=lx):
                # stride-1 conv over s*C channels; drop the one extra row of symmetric padding
                lout_full = lx + 2 * pad - (taps - 1)
                return ops.conv_desc(MS_CONV, B, s * conv.in_channels, conv.out_channels, lx, taps,
                                     1, pad, leaky=True, operand=self.operand,
                                     crop=lout_full - lx)
            d, packed = pm.get(conv.weight, s, make)
            h16, h32 = ops.conv_fwd(d, xs, packed, conv.bias, want16=True, want32=True)
            features.append(ops.unpack_blk32(h32))
            length = lx
        if self.conditioning_channels > 0:
            h16 = torch.cat([h16, feat16], dim=1)
        h32 = None
        for conv, pj in zip(self.mj, self._pj):
            h16, h32 = _k3(conv, pj, h16, operand=self.operand)
            features.append(ops.unpack_blk32(h32))
        j = ops.conv_to_mono(h32, self.judge.weight, self.judge.bias, 3, 1, False)
        return features, h16, j

    def forward(self, x, feat):
        _fwd_only(self, x)
        feat16 = ops.pack_ncl(feat, operand=self.operand) if self.conditioning_channels > 0 else None
        f, h16, j = self.forward_blocked(x, feat16)
        return f, f[-1], j


def _fwd_only(module, x):
    probe = x if isinstance(x, torch.Tensor) else next(iter(x.values()))
    if torch.is_grad_enabled() and (probe.requires_grad or
                                    any(p.requires_grad for p in module.parameters())):
        raise MsbError("sm_100a path is forward-only in this build: use torch.no_grad()")


class FilterBankMultiScaleDiscriminator(nn.Module):
    def __init__(self, input_size, samplerate, decompose=True, conditioning_channels=0):
        super().__init__()
        self.samplerate = samplerate
        self.conditioning_channels = conditioning_channels
        self.decompose = decompose
        self.input_size = input_size
        band_sizes = [int(2 ** (np.log2(self.input_size) - i)) for i in range(5)]
        sr = _as_samplerate(samplerate)
        strides = ([4, 4, 4, 4], [4, 4, 4, 2], [4, 4, 2, 2], [4, 2, 2, 2], [2, 2, 2, 2])
        self.spec = {}
        self.channel_discs = {}
        for i, (size, sf) in enumerate(zip(band_sizes, strides)):
            srb = sr * (2 ** i)
            start = 0 if i == 4 else srb.nyquist / 2
            self.spec[size] = {
                "scale_factors": sf, "channels": [128] * 5,
                "filter_bank": FilterBank(srb, 128,
                                          linear_center_frequencies(start, srb.nyquist, 128),
                                          scaling_factors=0.05),
                "conditioning_channels": conditioning_channels}
            disc = FilterBankChannelDiscriminator(**self.spec[size])
            self.add_module(f"channel_{size}", disc)
            self.channel_discs[size] = disc
        self.smallest_band = min(self.spec.keys())
        final_channels = sum(v["channels"][-1] for v in self.spec.values())
        channels = 512
        self.final = nn.Sequential(
            nn.Conv1d(final_channels + self.conditioning_channels, channels, 3, 1, 1),
            nn.Conv1d(channels, channels, 3, 1, 1),
            nn.Conv1d(channels, channels, 3, 1, 1))
        self.judge = nn.Conv1d(channels, 1, 3, 1, 1)
        self._pf = [_PackedConv() for _ in self.final]
        self._cf = [ag.WeightCache() for _ in self.final]

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        probe = fn(torch.zeros(1))
        for d in self.channel_discs.values():
            d.filter_bank.to(probe.device)
        return out

    def _forward_train(self, bands, feat):
        """autograd-recorded path (training): Functions over the C ABI, see ../autograd.py"""
        cond = self.conditioning_channels > 0
        feat32 = grad_ops.pack_ncl32(feat) if cond else None
        feat16 = ops.pack_ncl(feat) if cond else None
        features, c32, c16, judgements = [], [], [], []
        for size, layer in self.channel_discs.items():
            f, h32, h16, j = layer.forward_blocked_train(bands[size], feat32, feat16)
            features.append(f)
            c32.append(h32)
            c16.append(h16)
            judgements.append(j)
        x32, x16 = torch.cat(c32, dim=1), torch.cat(c16, dim=1)
        if cond:
            T = x16.shape[2]
            if feat.shape[-1] != T:      # F.upsample(feat, size=T): nearest neighbour
                idx = (torch.arange(T, device=feat.device) * feat.shape[-1]) // T
                up = feat[..., idx].contiguous()
                feat32, feat16 = grad_ops.pack_ncl32(up), ops.pack_ncl(up)
            x32, x16 = torch.cat([x32, feat32], dim=1), torch.cat([x16, feat16], dim=1)
        final_features = []
        for conv, cf in zip(self.final, self._cf):
            x32, x16 = ag.conv_blk(x32, x16, conv.weight, conv.bias, cf, MS_CONV, 1, 1, 1, True)
            final_features.append(ag.UnpackBlk32.apply(x32))
        features.append(final_features)
        judgements.append(ag.MonoConv.apply(x32, self.judge.weight, self.judge.bias, 3, 1, False))
        return features, judgements

    def forward(self, x, feat):
        probe = x if isinstance(x, torch.Tensor) else next(iter(x.values()))
        if ag.needs_grad(self, probe):
            if self.decompose:     # differentiable band split (discriminator/multiscale.py:212-216)
                x = fft_frequency_decompose(x, self.smallest_band)
            return self._forward_train(x, feat)
        bands = fft_frequency_decompose(x, self.smallest_band) if self.decompose else x
        cond = self.conditioning_channels > 0
        feat16 = ops.pack_ncl(feat) if cond else None
        features, channels, judgements = [], [], []
        for size, layer in self.channel_discs.items():
            f, h16, j = layer.forward_blocked(bands[size], feat16)
            features.append(f)
            channels.append(h16)
            judgements.append(j)
        x16 = torch.cat(channels, dim=1)
        if cond:
            T = x16.shape[2]
            if feat.shape[-1] != T:      # F.upsample(feat, size=T): nearest neighbour
                idx = (torch.arange(T, device=feat.device) * feat.shape[-1]) // T
                feat16 = ops.pack_ncl(feat[..., idx].contiguous())
            x16 = torch.cat([x16, feat16], dim=1)
        final_features = []
        h32 = None
        for conv, pf in zip(self.final, self._pf):
            x16, h32 = _k3(conv, pf, x16)
            final_features.append(ops.unpack_blk32(h32))
        features.append(final_features)
        judgements.append(ops.conv_to_mono(h32, self.judge.weight, self.judge.bias, 3, 1, False))
        return features, judgements


# ---------------------------------------------------------------------------------------
# Non-filterbank multiscale discriminator, featuresynth/discriminator/multiscale.py:10-67,
# 255-410 (MultiScaleDiscriminator, MultiScaleMultiResDiscriminator)
# ---------------------------------------------------------------------------------------
class ChannelDiscriminator(nn.Module):
    """discriminator/multiscale.py:10-67: four dense strided convs (kernel_size 41 or 9, stride 4 or
    2, channels 1 -> 32 -> 64 -> 128 -> 256) + optional judgement head.  The 1-input-channel first
    layer is a direct fp32 conv; the others run on the tcgen05 kernel as stride-1 convs over the
    space-to-depth input (k41: 11 taps at stride 4, 21 taps at stride 2)."""

    def __init__(self, scale_factors, channels, return_judgements=False, conditioning_channels=0,
                 kernel_size=41):
        super().__init__()
        self.kernel_size = kernel_size
        self.conditioning_channels = conditioning_channels
        self.return_judgements = return_judgements
        self.channels = channels
        self.scale_factors = scale_factors
        self.main = nn.Sequential(*[
            nn.Conv1d(channels[i], channels[i + 1], kernel_size, scale_factors[i],
                      padding=kernel_size // 2) for i in range(len(scale_factors))])
        if return_judgements:
            start = channels[-1] + (conditioning_channels if conditioning_channels > 0 else 0)
            self.mj = nn.Sequential(
                nn.Conv1d(start, channels[-1], 3, 1, 1),
                nn.Conv1d(channels[-1], channels[-1], 3, 1, 1),
                nn.Conv1d(channels[-1], channels[-1], 3, 1, 1))
            self.judge = nn.Conv1d(channels[-1], 1, 3, 1, 1)
            self._pj = [_PackedConv() for _ in self.mj]
            self._cj = [ag.WeightCache() for _ in self.mj]
        self._pm = [_PackedStrided() for _ in self.main]
        self._sc = [ag.StridedCache() for _ in self.main]

    def forward_blocked_train(self, x, feat32, feat16):
        """autograd-recorded form -> ([maps NCL f32], h32, h16 of the last map, judgement|None)"""
        features = []
        first = self.main[0]
        h = ag.DirectConv.apply(x, first.weight, first.bias, first.stride[0], first.padding[0], 1,
                                True)
        features.append(h)
        h32, h16 = ag.PackBlk32.apply(h), ops.pack_ncl(h.detach())
        length = h.shape[-1]
        for conv, sc in list(zip(self.main, self._sc))[1:]:
            h32, h16 = ag.StridedConvBlk.apply(h32, h16, conv.weight, conv.bias, sc, conv.stride[0],
                                               length)
            features.append(ag.UnpackBlk32.apply(h32))
            length = h16.shape[2]
        if not self.return_judgements:
            return features, h32, h16, None
        if self.conditioning_channels > 0:
            h32, h16 = torch.cat([h32, feat32], dim=1), torch.cat([h16, feat16], dim=1)
        for conv, cj in zip(self.mj, self._cj):
            h32, h16 = ag.conv_blk(h32, h16, conv.weight, conv.bias, cj, MS_CONV, 1, 1, 1, True)
            features.append(ag.UnpackBlk32.apply(h32))
        j = ag.MonoConv.apply(h32, self.judge.weight, self.judge.bias, 3, 1, False)
        return features, h32, h16, j

    def forward_blocked(self, x, feat16):
        """x (B,1,L) f32 -> ([maps NCL f32], last map BLK16, last map NCL f32, judgement|None)"""
        B = x.shape[0]
        features = []
        first = self.main[0]
        h = ops.conv1d_direct(x, first.weight, first.bias, first.stride[0], first.padding[0], 1,
                              leaky=True)
        features.append(h)
        h16, length = ops.pack_ncl(h), h.shape[-1]
        for conv, pm in list(zip(self.main, self._pm))[1:]:
            s = conv.stride[0]
            xs = ops.space_to_depth(h16, s, length)
            lx = xs.shape[2]

            def make(taps, pad, B=B, conv=conv, s=s, lx=lx):
                lout_full = lx + 2 * pad - (taps - 1)
                return ops.conv_desc(MS_CONV, B, s * conv.in_channels, conv.out_channels, lx, taps,
                                     1, pad, leaky=True, crop=lout_full - lx)
            d, packed = pm.get(conv.weight, s, make)
            h16, h32 = ops.conv_fwd(d, xs, packed, conv.bias, want16=True, want32=True)
            features.append(ops.unpack_blk32(h32))
            length = lx
        if not self.return_judgements:
            return features, h16, features[-1], None
        if self.conditioning_channels > 0:
            h16 = torch.cat([h16, feat16], dim=1)
        h32 = None
        for conv, pj in zip(self.mj, self._pj):
            h16, h32 = _k3(conv, pj, h16)
            features.append(ops.unpack_blk32(h32))
        j = ops.conv_to_mono(h32, self.judge.weight, self.judge.bias, 3, 1, False)
        return features, h16, features[-1], j

    def forward(self, x, feat=None):
        _fwd_only(self, x)
        feat16 = ops.pack_ncl(feat) if (self.return_judgements and
                                        self.conditioning_channels > 0) else None
        f, _, last, j = self.forward_blocked(x, feat16)
        return (f, last, j) if self.return_judgements else (f, last)


class MultiScaleDiscriminator(nn.Module):
    """discriminator/multiscale.py:255-375."""

    def __init__(self, input_size, decompose=True, channel_judgements=False,
                 conditioning_channels=0, kernel_size=41):
        super().__init__()
        self.kernel_size = kernel_size
        self.conditioning_channels = conditioning_channels
        self.channel_judgements = channel_judgements
        self.decompose = decompose
        self.input_size = input_size
        band_sizes = [int(2 ** (np.log2(self.input_size) - i)) for i in range(5)]
        factors = ([4, 4, 4, 4], [4, 4, 4, 2], [4, 4, 2, 2], [4, 2, 2, 2], [2, 2, 2, 2])
        self.spec = {bs: {"scale_factors": sf, "channels": [1, 32, 64, 128, 256]}
                     for bs, sf in zip(band_sizes, factors)}
        self.smallest_band = min(self.spec.keys())
        self.channel_discs = {}
        for key, value in self.spec.items():
            disc = ChannelDiscriminator(**value, return_judgements=self.channel_judgements,
                                        conditioning_channels=self.conditioning_channels,
                                        kernel_size=kernel_size)
            self.add_module(f"channel_{key}", disc)
            self.channel_discs[key] = disc
        final_channels = sum(v["channels"][-1] for v in self.spec.values())
        channels = 512
        self.final = nn.Sequential(
            nn.Conv1d(final_channels + self.conditioning_channels, channels, 3, 1, 1),
            nn.Conv1d(channels, channels, 3, 1, 1),
            nn.Conv1d(channels, channels, 3, 1, 1))
        self.judge = nn.Conv1d(channels, 1, 3, 1, 1)
        self.recon = None
        self._pf = [_PackedConv() for _ in self.final]
        self._cf = [ag.WeightCache() for _ in self.final]

    def _forward_train(self, bands, feat):
        cond = self.conditioning_channels > 0
        feat32 = grad_ops.pack_ncl32(feat) if cond else None
        feat16 = ops.pack_ncl(feat) if cond else None
        features, c32, c16, judgements = [], [], [], []
        for size, layer in self.channel_discs.items():
            f, h32, h16, j = layer.forward_blocked_train(bands[size], feat32, feat16)
            features.append(f)
            c32.append(h32)
            c16.append(h16)
            if self.channel_judgements:
                judgements.append(j)
        x32, x16 = torch.cat(c32, dim=1), torch.cat(c16, dim=1)
        if cond:
            T = x16.shape[2]
            if feat.shape[-1] != T:
                idx = (torch.arange(T, device=feat.device) * feat.shape[-1]) // T
                up = feat[..., idx].contiguous()
                feat32, feat16 = grad_ops.pack_ncl32(up), ops.pack_ncl(up)
            x32, x16 = torch.cat([x32, feat32], dim=1), torch.cat([x16, feat16], dim=1)
        final_features = []
        for conv, cf in zip(self.final, self._cf):
            x32, x16 = ag.conv_blk(x32, x16, conv.weight, conv.bias, cf, MS_CONV, 1, 1, 1, True)
            final_features.append(ag.UnpackBlk32.apply(x32))
        features.append(final_features)
        judgements.append(ag.MonoConv.apply(x32, self.judge.weight, self.judge.bias, 3, 1, False))
        return features, judgements

    def forward(self, x, feat):
        probe = x if isinstance(x, torch.Tensor) else next(iter(x.values()))
        if ag.needs_grad(self, probe):
            if self.decompose:     # differentiable band split (discriminator/multiscale.py:395-397)
                x = fft_frequency_decompose(x, self.smallest_band)
            return self._forward_train(x, feat)
        bands = fft_frequency_decompose(x, self.smallest_band) if self.decompose else x
        cond = self.conditioning_channels > 0
        feat16 = ops.pack_ncl(feat) if cond else None
        features, channels, judgements = [], [], []
        for size, layer in self.channel_discs.items():
            f, h16, _, j = layer.forward_blocked(bands[size], feat16)
            features.append(f)
            channels.append(h16)
            if self.channel_judgements:
                judgements.append(j)
        x16 = torch.cat(channels, dim=1)
        if cond:
            T = x16.shape[2]
            if feat.shape[-1] != T:      # F.upsample(feat, size=T): nearest neighbour
                idx = (torch.arange(T, device=feat.device) * feat.shape[-1]) // T
                feat16 = ops.pack_ncl(feat[..., idx].contiguous())
            x16 = torch.cat([x16, feat16], dim=1)
        final_features = []
        h32 = None
        for conv, pf in zip(self.final, self._pf):
            x16, h32 = _k3(conv, pf, x16)
            final_features.append(ops.unpack_blk32(h32))
        features.append(final_features)
        judgements.append(ops.conv_to_mono(h32, self.judge.weight, self.judge.bias, 3, 1, False))
        return features, judgements


class MultiScaleMultiResDiscriminator(nn.Module):
    """discriminator/multiscale.py:378-410."""

    def __init__(self, input_size, flatten_multiscale_features=False, decompose=True,
                 channel_judgements=False, conditioning_channels=0, kernel_size=41):
        super().__init__()
        self.kernel_size = kernel_size
        self.conditioning_channels = conditioning_channels
        self.input_size = input_size
        self.flatten_multiscale_features = flatten_multiscale_features
        self.multiscale = MultiScaleDiscriminator(input_size, decompose, channel_judgements,
                                                  conditioning_channels, kernel_size)

    def forward(self, x, feat):
        features, judgements = [], []
        f, j = self.multiscale(x, feat)
        if self.flatten_multiscale_features:
            features.append([m for group in f for m in group])
        else:
            features.extend(f)
        judgements.extend(j)
        return features, judgements
