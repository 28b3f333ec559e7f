"""Short program for `ncu --set full -k regex:'audio2mel_r16|fft_pass'`: one warm + one profiled
call of Audio2Mel (B=4096 x 16384) and of the FFT band split (64 x 65536, 5 bands)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200.feature.feature import Audio2Mel
from music_synthesis_b200.audio.transform import fft_frequency_decompose
from oracle import synth

torch.set_grad_enabled(False)
a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
a = synth.uniform_audio(3, 64, 16384).repeat(64, 1, 1).cuda()
x = (synth.randn(4, 8, 1, 65536) * 0.1).repeat(8, 1, 1).cuda()
for _ in range(2):
    a2m(a)
    fft_frequency_decompose(x, 4096)
torch.cuda.synchronize()
