"""Short program for `ncu --set full -k regex:'audio2mel_r16|fft_pass'`: one warm + one profiled
call of Audio2Mel (B=4096 x 16384) and of the FFT band split (64 x 65536, 5 bands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200.feature.feature import Audio2Mel
from music_synthesis_b200.audio.transform import fft_frequency_decompose
import numpy as np

torch.set_grad_enabled(False)


def uniform_audio(seed, batch, samples):
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.random_sample((batch, 1, samples)) * 2 - 1).astype(np.float32))


def randn(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))

a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
a = uniform_audio(3, 64, 16384).repeat(64, 1, 1).cuda()
x = (randn(4, 8, 1, 65536) * 0.1).repeat(8, 1, 1).cuda()
for _ in range(2):
    a2m(a)
    fft_frequency_decompose(x, 4096)
torch.cuda.synchronize()
