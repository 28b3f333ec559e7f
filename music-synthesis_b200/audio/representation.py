"""Audio representations used by the hot path (featuresynth/audio/representation.py:12-54,
82-103): `RawAudio` and `MultiScale`.

Same contract as the reference -- numpy batches in, numpy batches out, `data` is what the
experiment feeds its networks -- but `MultiScale.from_audio / to_audio` run the octave-band
split / merge on the GPU (`ms_fft_frequency_decompose / recompose`) instead of torch's CPU FFT
(the reference does this step on the host inside its batch
pre-processing).  `from_audio(..., device_bands=True)` additionally keeps the bands as CUDA
tensors so a training loop can hand them to the discriminator without a host round trip.
The display / listen helpers of the reference (zounds, matplotlib) are out of scope.
"""
import numpy as np
import torch

from .transform import fft_frequency_decompose, fft_frequency_recompose


class BaseAudioRepresentation(object):
    def __init__(self, data, samplerate):
        super().__init__()
        self.samplerate = samplerate
        self.data = data

    @classmethod
    def from_audio(cls, samples, samplerate):
        raise NotImplementedError()

    def to_audio(self):
        raise NotImplementedError()


class RawAudio(BaseAudioRepresentation):
    @classmethod
    def from_audio(cls, samples, samplerate):
        return cls(samples, samplerate)

    def to_audio(self):
        batch, _, samples = self.data.shape
        return self.data.reshape((batch, samples))


def _device_tensor(v, device):
    if isinstance(v, torch.Tensor):
        return v.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).to(device)


class MultiScale(BaseAudioRepresentation):
    """{band size: (batch, 1, size)}: five octave bands, the largest as long as the clip."""
    N_BANDS = 5

    @classmethod
    def from_audio(cls, samples, samplerate, device="cuda", device_bands=None):
        """device_bands: keep the bands as CUDA tensors (default: when `samples` already is one,
        i.e. inside the training loop) instead of the reference's numpy arrays."""
        if device_bands is None:
            device_bands = isinstance(samples, torch.Tensor) and samples.is_cuda
        with torch.no_grad():
            time = samples.shape[-1]
            start = int(np.log2(time))
            smallest = 2 ** (start - cls.N_BANDS + 1)
            x = _device_tensor(samples, device)
            lead = x.shape[:-1]
            bands = fft_frequency_decompose(x.reshape(-1, 1, time), smallest)
            bands = {k: v.reshape(*lead, k) for k, v in bands.items()}
            if not device_bands:
                bands = {k: v.cpu().numpy() for k, v in bands.items()}
            return cls(bands, samplerate)

    def to_audio(self, device="cuda"):
        with torch.no_grad():
            mx = max(v.shape[-1] for v in self.data.values())
            bands = {k: _device_tensor(v, device) for k, v in self.data.items()}
            bands = {k: v.reshape(-1, 1, v.shape[-1]) for k, v in bands.items()}
            samples = fft_frequency_recompose(bands, mx)
            return samples.cpu().numpy().reshape((-1, mx))
