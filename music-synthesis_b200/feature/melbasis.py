"""Host-side construction of the Slaney mel filterbank the reference obtains from
`librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` (feature/feature.py:27-29;
htk=False, norm='slaney').  Vectorised numpy; float32 (n_mels, 1 + n_fft//2)."""
import numpy as np

_F_SP = 200.0 / 3
_MIN_LOG_HZ = 1000.0
_MIN_LOG_MEL = _MIN_LOG_HZ / _F_SP
_LOGSTEP = np.log(6.4) / 27.0


def hz_to_mel(hz):
    hz = np.asarray(hz, dtype=np.float64)
    lin = hz / _F_SP
    log = _MIN_LOG_MEL + np.log(np.maximum(hz, 1e-30) / _MIN_LOG_HZ) / _LOGSTEP
    return np.where(hz >= _MIN_LOG_HZ, log, lin)


def mel_to_hz(mel):
    mel = np.asarray(mel, dtype=np.float64)
    return np.where(mel >= _MIN_LOG_MEL,
                    _MIN_LOG_HZ * np.exp(_LOGSTEP * (mel - _MIN_LOG_MEL)), _F_SP * mel)


def mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    fmax = float(sr) / 2 if fmax is None else float(fmax)
    bin_hz = np.linspace(0.0, float(sr) / 2, 1 + n_fft // 2)
    edges = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    width = np.diff(edges)
    ramps = edges[:, None] - bin_hz[None, :]
    rising = -ramps[:-2] / width[:-1, None]
    falling = ramps[2:] / width[1:, None]
    tri = np.maximum(0.0, np.minimum(rising, falling))
    tri *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return tri.astype(np.float32)
