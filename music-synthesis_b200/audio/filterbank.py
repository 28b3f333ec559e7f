"""Fixed Morlet filter bank on the B200 path.

Stand-in for `zounds.learn.FilterBank` exactly as the reference's hot path uses it
(featuresynth/generator/multiscale.py:86-92,151-164; discriminator/multiscale.py:109-112,
197-210): a (n_bands, 1, kernel_size) float32 tensor held as a PLAIN attribute (not a
Parameter / buffer), `convolve` (analysis) and `transposed_convolve` (synthesis).  The bank
is built on the host (real Morlet wavelets on a linear scale, unit norm); both operations run
on the tcgen05 implicit-GEMM conv kernel (see include/msb200.h, "filter bank").

zounds itself is not available offline; the construction follows its published algorithm
 and is checked in tests against the bank tensors the reference modules
built through the oracle harness.
"""
import numpy as np
import torch

from .. import _lib, ops
from .._lib import MS_CONV, MS_F16, check, ptr, stream_ptr


class SampleRate:
    """`int()` = samples/s, `.nyquist`; `rate * k` multiplies the sample PERIOD (zounds)."""

    def __init__(self, rate):
        self.rate = float(rate)

    def __int__(self):
        return int(self.rate)

    @property
    def nyquist(self):
        return self.rate / 2.0

    def __mul__(self, k):
        return SampleRate(self.rate / k)


def linear_center_frequencies(start_hz, stop_hz, n_bands):
    w = (float(stop_hz) - float(start_hz)) / n_bands
    return [float(start_hz) + w * i + w / 2 for i in range(n_bands)]


def morlet_bank(samplerate, kernel_size, center_frequencies, scaling_factor=0.05,
                normalize=True):
    """(n, kernel_size) float32: real part of Morlet wavelets (old scipy.signal.morlet)."""
    sr = float(samplerate.rate) if hasattr(samplerate, "rate") else float(int(samplerate))
    s = float(scaling_factor)
    x = np.linspace(-s * 2 * np.pi, s * 2 * np.pi, kernel_size)
    env = np.exp(-0.5 * x ** 2) * np.pi ** (-0.25)
    bank = np.zeros((len(center_frequencies), kernel_size), dtype=np.float64)
    for i, cf in enumerate(center_frequencies):
        w = cf / (s * 2 * sr / kernel_size)
        bank[i] = ((np.exp(1j * w * x) - np.exp(-0.5 * w ** 2)) * env).real
    if normalize:
        bank = bank / (np.linalg.norm(bank, axis=-1, keepdims=True) + 1e-8)
    return bank.astype(np.float32)


class FilterBank:
    """kernel_size: any length up to 512 taps.  Internally the bank is zero-padded at the end to
    kp = the next multiple of 32 taps (a no-op on both operations).  Synthesis runs with 32 phase
    channels (N = 32 GEMM, kp / 32 taps of dilation 32): with the first version's 8 phases the
    tensor core saw N = 16 MMAs (half of them zero columns) and 16 taps -- 4x the MMA count for the
    same arithmetic; the 65 536-sample band of config 5 took 1.74 ms per call.
    torch conventions kept: `convolve` = conv1d(padding = k // 2) -> L + 1 rows for even k, L for
    odd k; `transposed_convolve` = conv_transpose1d(padding = k // 2)."""

    def __init__(self, samplerate, kernel_size, center_frequencies, scaling_factors=0.05,
                 normalize_filters=True, a_weighting=False, operand=MS_F16, bank=None):
        if kernel_size < 16 or kernel_size > 512:
            raise NotImplementedError("FilterBank: 16 <= kernel_size <= 512")
        self.samplerate = samplerate
        self.kernel_size = kernel_size
        self.operand = operand
        if bank is None:
            bank = morlet_bank(samplerate, kernel_size, center_frequencies, scaling_factors,
                               normalize_filters)
        bank = torch.as_tensor(bank, dtype=torch.float32)
        self.n_bands = bank.shape[0]
        if self.n_bands % 16 != 0:
            raise NotImplementedError("FilterBank: n_bands must be a multiple of 16")
        self.filter_bank = bank.reshape(self.n_bands, 1, kernel_size)
        self.pad = kernel_size // 2                       # torch padding of both operations
        self.kp = (kernel_size + 31) // 32 * 32           # padded tap count
        self.extra = 1 - kernel_size % 2                  # convolve returns L + extra rows
        # synthesis as a conv: nph phase channels, kp / nph taps of dilation nph, padding syn_pad,
        # then y[t] = sum_i z[t + i + 1, i] (ms_diag_sum, skew 1)
        self.syn_nph = 32
        self.syn_ch = 32                                  # channels of the phase tensor
        self.syn_taps = self.kp // self.syn_nph
        self.syn_pad = kernel_size - self.pad
        self._packed = {}

    def to(self, device):
        self.filter_bank = self.filter_bank.to(device)
        self._packed = {}
        return self

    def _padded(self, device, flip=False):
        n, k = self.n_bands, self.kernel_size
        b = self.filter_bank.to(device).reshape(n, k)
        if flip:
            b = torch.flip(b, dims=[1])
        if self.kp != k:
            b = torch.cat([b, torch.zeros((n, self.kp - k), dtype=b.dtype, device=device)], dim=1)
        return b

    # -- analysis: conv1d(x, bank, padding=k//2) -> (B, n, L + extra) -------------------------
    def analysis_len(self, L):
        return L + 2 * self.pad - self.kernel_size + 1

    def _analysis(self, x, want16, want32):
        _lib.require_cuda(x, "x")
        L = x.shape[-1]
        x = x.reshape(-1, 1, L).contiguous()
        B, n = x.shape[0], self.n_bands
        taps = self.kp // 16
        Lx = self.analysis_len(L) + 16 * (taps - 1)
        x16 = torch.empty((B, 2, Lx, 8), dtype=torch.int16, device=x.device)
        check(_lib.lib().ms_expand_mono_to_blk16(ptr(x), ptr(x16), B, L, Lx, self.pad,
                                                 self.operand, stream_ptr()),
              "ms_expand_mono_to_blk16")
        d = ops.conv_desc(MS_CONV, B, 16, n, Lx, taps, 16, 0, operand=self.operand)
        key = ("a", x.device)
        if key not in self._packed:
            self._packed[key] = ops.pack_conv_weight(d, self.analysis_weight(x.device))
        return ops.conv_fwd(d, x16, self._packed[key], None, want16=want16, want32=want32)

    def analysis_weight(self, device):
        """(n, 16, kp/16) weight of the analysis conv over the 16-wide sliding-window expansion:
        W[f, i, j] = bank[f, 16 j + i]"""
        n, taps = self.n_bands, self.kp // 16
        return self._padded(device).reshape(n, taps, 16).permute(0, 2, 1).contiguous()

    def synthesis_weight(self, device):
        """(syn_ch, n, kp/nph) weight of the synthesis conv (nph phase channels used):
        Wg[i, c, j] = flip(bank)[c, nph j + i]"""
        n, nph, taps = self.n_bands, self.syn_nph, self.syn_taps
        wf = self._padded(device, flip=True)
        w = torch.zeros((self.syn_ch, n, taps), dtype=torch.float32, device=device)
        w[:nph] = wf.reshape(n, taps, nph).permute(2, 0, 1)
        return w.contiguous()

    def convolve(self, x):
        return ops.unpack_blk32(self._analysis(x, False, True)[1])

    def convolve_blocked(self, x):
        """Same as `convolve` but returns the channel-blocked 16-bit tensor (B, n/8, L+extra, 8)
        the next tcgen05 conv consumes directly."""
        return self._analysis(x, True, False)[0]

    # -- synthesis: conv_transpose1d(x(B,n,L+extra), bank, padding=k//2) -> (B,1,L) -----------
    #: power-of-two operand scales of the full-split synthesis (activations of ~1e-2, unit-norm
    #: filters of ~5e-2 per tap: keeps the lo terms out of the fp16 subnormals)
    SPLIT_SX, SPLIT_SW = 64.0, 256.0

    def _synth_weights(self, d, device, wsplit=0):
        key = ("s%d" % int(wsplit), device)
        if key not in self._packed:
            w = self.synthesis_weight(device)
            if wsplit == 2:
                w = ops.weight_split(w, self.operand, self.SPLIT_SW, terms=3)
            elif wsplit:
                w = ops.weight_split(w, self.operand, ops.W_SPLIT_SCALE, terms=2)
            self._packed[key] = ops.pack_conv_weight(d, w)
        return self._packed[key]

    def synthesis_rows(self, L):
        """rows of the phase tensor z for L input rows"""
        return L + 2 * self.syn_pad - self.syn_nph * (self.syn_taps - 1)

    def _synthesis(self, x16, Lin, Lout, wsplit, x32=None):
        """wsplit: 0 = 16-bit operands, 1 = weight split over the duplicated operand, 2 = full
        split precision from the fp32 activations x32.  The synthesis output is a heavily
        cancelling sum (narrow-band filters over a broadband input): rounding its INPUT to fp16
        alone puts 1.9e-3 on the waveform of the 511-tap bank (5e-4 for the 128-tap banks)."""
        B, n = x16.shape[0], self.n_bands
        xrep = 1
        if wsplit == 2:
            if x32 is None:
                raise _lib.MsbError("full-split synthesis needs the fp32 activations")
            mult, alpha = 3, 1.0 / (self.SPLIT_SX * self.SPLIT_SW)
            x16 = ops.blk32_split(x32, operand=self.operand, terms=3, scale=self.SPLIT_SX)
        elif wsplit:
            mult, alpha, xrep = 2, 1.0 / ops.W_SPLIT_SCALE, 2    # the K loop wraps over x16 twice
        else:
            mult, alpha = 1, 1.0
        d = ops.conv_desc(MS_CONV, B, mult * n, self.syn_ch, Lin, self.syn_taps, self.syn_nph, self.syn_pad,
                          operand=self.operand, alpha=alpha, x_repeat=xrep)
        _, z32 = ops.conv_fwd(d, x16, self._synth_weights(d, x16.device, wsplit), None,
                              want16=False, want32=True)
        y = torch.empty((B, 1, Lout), dtype=torch.float32, device=x16.device)
        check(_lib.lib().ms_diag_sum(ptr(z32), ptr(y), B, self.syn_ch, z32.shape[2], Lout, self.syn_nph, 1,
                                     stream_ptr()), "ms_diag_sum")
        return y

    def transposed_convolve_blocked(self, x16, L, wsplit=0, x32=None):
        """x16: BLK 16-bit (B, n/8, L, 8) holding rows 0..L-1 of the input of
        `transposed_convolve` (for an even kernel the reference appends one zero row first,
        generator/multiscale.py:90 -- it contributes nothing); returns (B, 1, L) f32.  wsplit:
        bank weights as a (hi, lo) pair over the duplicated operand (no weight rounding)."""
        return self._synthesis(x16, L, L, wsplit, x32)

    def transposed_convolve(self, x):
        _lib.require_cuda(x, "x")
        B, n, Lp = x.shape
        x16 = ops.pack_ncl(x.contiguous(), operand=self.operand)
        return self._synthesis(x16, Lp, Lp - self.extra, 0)
