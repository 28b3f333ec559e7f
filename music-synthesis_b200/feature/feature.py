"""Drop-in mirror of featuresynth/feature/feature.py:11-59 (Audio2Mel).

Same constructor signature, buffers (`mel_basis`, `window`) and forward contract:
audio (B,1,N) tensor or (N,) ndarray -> log10-mel (B, n_mel_channels, F), with the
reference's one-sided right zero padding of (n_fft-hop)//2 (F = (N-640)//256+1 for
1024/256).  The whole transform is one fused CUDA kernel (csrc/audio2mel.cu).
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from .._lib import MsbError
from .melbasis import mel_filterbank


class Audio2Mel(nn.Module):
    def __init__(self, n_fft=1024, hop_length=256, win_length=1024, sampling_rate=22050,
                 n_mel_channels=80, mel_fmin=0.0, mel_fmax=None):
        super().__init__()
        if win_length != n_fft:
            raise NotImplementedError("Audio2Mel: win_length must equal n_fft on this path")
        window = torch.hann_window(win_length).float()
        mel_basis = torch.from_numpy(
            mel_filterbank(sampling_rate, n_fft, n_mel_channels, mel_fmin, mel_fmax)).float()
        self.register_buffer("mel_basis", mel_basis)
        self.register_buffer("window", window)
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.win_length = win_length
        self.sampling_rate = sampling_rate
        self.n_mel_channels = n_mel_channels
        self._ranges = None          # (key, device tensor): banded structure of mel_basis

    def forward(self, audio):
        if isinstance(audio, np.ndarray):
            audio = torch.from_numpy(audio).view(1, 1, -1).to(self.window.device)
        if audio.dim() != 3 or audio.shape[1] != 1:
            raise MsbError("Audio2Mel expects (B, 1, N) audio")
        key = (self.mel_basis.data_ptr(), self.mel_basis._version)
        if self._ranges is None or self._ranges[0] != key:
            self._ranges = (key, ops.mel_row_ranges(self.mel_basis).to(self.mel_basis.device))
        return ops.audio2mel(audio.float(), self.window, self.mel_basis, self.n_fft,
                             self.hop_length, self._ranges[1])
