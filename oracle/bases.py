"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatements of the two THIRD-PARTY fixed bases the reference's hot path
depends on.  Their source is NOT under /root/reference (requirements.txt:1-3 names
`zounds>=1.55.0`, `lws`, `librosa`, all un-pinned and absent here), so these follow
the packages' published algorithms and are anchored on the reference's call sites.

PARITY UNPINNED for these two functions: the reference holds no golden vectors for
them.  Mitigations: (1) `librosa_mel` is cross-checked in tests against
`torchaudio.functional.melscale_fbanks(norm='slaney', mel_scale='slaney')`;
(2) kernels are always parity-tested with the *basis tensor* shared between oracle
and CUDA path, so the construction below cannot produce a false parity result.

Call sites:
  librosa.filters.mel  -> featuresynth/feature/feature.py:27-29 (positional:
                          sr, n_fft, n_mels, fmin, fmax  => librosa < 0.10 API)
  zounds.learn.FilterBank / LinearScale / FrequencyBand / SR22050 ->
                          featuresynth/generator/multiscale.py:151-164,
                          featuresynth/discriminator/multiscale.py:197-210,
                          test_multiscale_filterbank.py:14-35
"""
import numpy as np


# --------------------------------------------------------------------------- mel
def _hz_to_mel_slaney(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    with np.errstate(divide="ignore", invalid="ignore"):
        log_part = min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep
    return np.where(f >= min_log_hz, log_part, mels)


def _mel_to_hz_slaney(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel,
                    min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def librosa_mel(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax, htk=False, norm='slaney').

    Slaney mel scale, triangular filters, Slaney area normalisation, float32
    (n_mels, 1 + n_fft//2).
    """
    if fmax is None:
        fmax = float(sr) / 2
    n_bins = 1 + n_fft // 2
    fftfreqs = np.linspace(0, float(sr) / 2, n_bins, endpoint=True)
    mel_pts = np.linspace(_hz_to_mel_slaney(fmin), _hz_to_mel_slaney(fmax), n_mels + 2)
    mel_f = _mel_to_hz_slaney(mel_pts)
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights.astype(np.float32)


# ------------------------------------------------------------------ zounds pieces
class SampleRate:
    """zounds.SampleRate stand-in: `int()` = samples/s, `.nyquist`, and
    `rate * k` multiplies the sample PERIOD (rate / k), as zounds does."""

    def __init__(self, rate):
        self.rate = float(rate)

    def __int__(self):
        return int(self.rate)

    @property
    def samples_per_second(self):
        return int(self.rate)

    @property
    def nyquist(self):
        return self.rate / 2.0

    def __mul__(self, k):
        return SampleRate(self.rate / k)


class FrequencyBand:
    def __init__(self, start_hz, stop_hz):
        self.start_hz = float(start_hz)
        self.stop_hz = float(stop_hz)

    @property
    def bandwidth(self):
        return self.stop_hz - self.start_hz

    @property
    def center_frequency(self):
        return self.start_hz + self.bandwidth / 2


class LinearScale:
    """n equal-width contiguous bands on [lo, hi); centre = start + width/2."""

    def __init__(self, band, n_bands):
        self.band = band
        self.n_bands = n_bands

    def __len__(self):
        return self.n_bands

    @property
    def center_frequencies(self):
        w = self.band.bandwidth / self.n_bands
        return [self.band.start_hz + w * i + w / 2 for i in range(self.n_bands)]


def _morlet(M, w, s):
    """Old scipy.signal.morlet(M, w, s, complete=True) (removed from SciPy)."""
    x = np.linspace(-s * 2 * np.pi, s * 2 * np.pi, M)
    out = np.exp(1j * w * x)
    out = out - np.exp(-0.5 * (w ** 2))
    out = out * np.exp(-0.5 * (x ** 2)) * np.pi ** (-0.25)
    return out


def morlet_filter_bank(samplerate, kernel_size, scale, scaling_factors,
                       normalize=True):
    """(len(scale), kernel_size) float32 bank of real Morlet filters."""
    sr = float(int(samplerate)) if not isinstance(samplerate, SampleRate) \
        else samplerate.rate
    cfs = scale.center_frequencies
    if np.isscalar(scaling_factors):
        scaling_factors = [scaling_factors] * len(cfs)
    bank = np.zeros((len(cfs), kernel_size), dtype=np.complex128)
    for i, (cf, s) in enumerate(zip(cfs, scaling_factors)):
        w = cf / (s * 2 * sr / kernel_size)
        bank[i] = _morlet(kernel_size, w, s)
    bank = bank.real
    if normalize:
        bank = bank / (np.linalg.norm(bank, axis=-1, keepdims=True) + 1e-8)
    return bank.astype(np.float32)


class FilterBank:
    """zounds.learn.FilterBank stand-in (fixed bank; plain attribute, not a
    Parameter or buffer -- the reference relies on that, generator/generator.py:87)."""

    def __init__(self, samplerate, kernel_size, scale, scaling_factors,
                 normalize_filters=True, a_weighting=False):
        import torch
        self.samplerate = samplerate
        self.kernel_size = kernel_size
        self.scale = scale
        self.n_bands = len(scale)
        bank = morlet_filter_bank(samplerate, kernel_size, scale,
                                  scaling_factors, normalize_filters)
        self.filter_bank = torch.from_numpy(bank).view(len(scale), 1, kernel_size)

    def to(self, device):
        self.filter_bank = self.filter_bank.to(device)
        return self

    def convolve(self, x):
        from torch.nn import functional as F
        x = x.view(-1, 1, x.shape[-1])
        return F.conv1d(x, self.filter_bank, padding=self.kernel_size // 2)

    def transposed_convolve(self, x):
        from torch.nn import functional as F
        return F.conv_transpose1d(x, self.filter_bank,
                                  padding=self.kernel_size // 2)

    def temporal_pooling(self, x, kernel_size, stride):
        from torch.nn import functional as F
        return F.avg_pool1d(x, kernel_size, stride, padding=kernel_size // 2)
