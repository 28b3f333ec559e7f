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
    def __init__(self, samplerate, kernel_size, center_frequencies, scaling_factors=0.05,
                 normalize_filters=True, a_weighting=False, operand=MS_F16, bank=None):
        if kernel_size % 16 != 0 or kernel_size > 128:
            raise NotImplementedError("FilterBank: kernel_size must be a multiple of 16, <= 128")
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
        self._packed = {}

    def to(self, device):
        self.filter_bank = self.filter_bank.to(device)
        self._packed = {}
        return self

    # -- analysis: conv1d(x, bank, padding=k/2) -> (B, n, L+1) ------------------------------
    def _analysis(self, x, want16, want32):
        _lib.require_cuda(x, "x")
        L = x.shape[-1]
        x = x.reshape(-1, 1, L).contiguous()
        B, k, n = x.shape[0], self.kernel_size, self.n_bands
        taps = k // 16
        Lx = L + 1 + 16 * (taps - 1)
        x16 = torch.empty((B, 2, Lx, 8), dtype=torch.int16, device=x.device)
        check(_lib.lib().ms_expand_mono_to_blk16(ptr(x), ptr(x16), B, L, Lx, k // 2, self.operand,
                                                 stream_ptr()), "ms_expand_mono_to_blk16")
        d = ops.conv_desc(MS_CONV, B, 16, n, Lx, taps, 16, 0, operand=self.operand)
        key = ("a", x.device)
        if key not in self._packed:
            self._packed[key] = ops.pack_conv_weight(d, self.analysis_weight(x.device))
        return ops.conv_fwd(d, x16, self._packed[key], None, want16=want16, want32=want32)

    def analysis_weight(self, device):
        """(n, 16, k/16) weight of the analysis conv over the 16-wide sliding-window expansion:
        W[f, i, j] = bank[f, 16 j + i]"""
        n, taps = self.n_bands, self.kernel_size // 16
        return self.filter_bank.to(device).reshape(n, taps, 16).permute(0, 2, 1).contiguous()

    def synthesis_weight(self, device):
        """(16, n, k/8) weight of the synthesis conv (8 phase channels used):
        Wg[i, c, j] = flip(bank)[c, 8 j + i]"""
        n, k = self.n_bands, self.kernel_size
        taps = k // 8
        wf = torch.flip(self.filter_bank.to(device).reshape(n, k), dims=[1])
        w = torch.zeros((16, n, taps), dtype=torch.float32, device=device)
        w[:8] = wf.reshape(n, taps, 8).permute(2, 0, 1)
        return w.contiguous()

    def convolve(self, x):
        return ops.unpack_blk32(self._analysis(x, False, True)[1])

    def convolve_blocked(self, x):
        """Same as `convolve` but returns the channel-blocked 16-bit tensor (B, n/8, L+1, 8)
        the next tcgen05 conv consumes directly."""
        return self._analysis(x, True, False)[0]

    # -- synthesis: conv_transpose1d(x(B,n,L+1), bank, padding=k/2) -> (B,1,L) --------------
    def _synth_weights(self, d, device, wsplit=False):
        key = ("s2" if wsplit else "s", device)
        if key not in self._packed:
            w = self.synthesis_weight(device)
            if wsplit:
                w = ops.weight_split(w, self.operand, ops.W_SPLIT_SCALE, terms=2)
            self._packed[key] = ops.pack_conv_weight(d, w)
        return self._packed[key]

    def transposed_convolve_blocked(self, x16, L, wsplit=False):
        """x16: BLK 16-bit (B, n/8, L, 8) holding rows 0..L-1 of the (zero-padded to L+1)
        input of `transposed_convolve`; returns (B, 1, L) f32.  wsplit: bank weights as a
        (hi, lo) pair over the duplicated operand (no weight rounding)."""
        B = x16.shape[0]
        n, k = self.n_bands, self.kernel_size
        if wsplit:
            d = ops.conv_desc(MS_CONV, B, 2 * n, 16, L, k // 8, 8, k // 2, operand=self.operand,
                              alpha=1.0 / ops.W_SPLIT_SCALE)
            x16 = ops.dup_channels(x16)
        else:
            d = ops.conv_desc(MS_CONV, B, n, 16, L, k // 8, 8, k // 2, operand=self.operand)
        _, z32 = ops.conv_fwd(d, x16, self._synth_weights(d, x16.device, wsplit), None,
                              want16=False, want32=True)
        Lz = z32.shape[2]
        y = torch.empty((B, 1, L), dtype=torch.float32, device=x16.device)
        check(_lib.lib().ms_diag_sum(ptr(z32), ptr(y), B, 16, Lz, L, 8, 1, stream_ptr()),
              "ms_diag_sum")
        return y

    def transposed_convolve(self, x):
        _lib.require_cuda(x, "x")
        B, n, Lp = x.shape
        # the last input row only ever meets taps that fall outside the output: it is the
        # reference's F.pad(x, (0, 1)) zero (generator/multiscale.py:90) -- but keep general
        x16 = ops.pack_ncl(x.contiguous(), operand=self.operand)
        d = ops.conv_desc(MS_CONV, B, n, 16, Lp, self.kernel_size // 8, 8, self.kernel_size // 2,
                          operand=self.operand)
        _, z32 = ops.conv_fwd(d, x16, self._synth_weights(d, x.device), None,
                              want16=False, want32=True)
        y = torch.empty((B, 1, Lp - 1), dtype=torch.float32, device=x.device)
        check(_lib.lib().ms_diag_sum(ptr(z32), ptr(y), B, 16, z32.shape[2], Lp - 1, 8, 1,
                                     stream_ptr()), "ms_diag_sum")
        return y
