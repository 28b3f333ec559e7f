"""One warm and one profiled band split + merge at 512 x 65536 (for ncu: -k regex:fft_pass -s 24 -c 24)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose

torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
x = torch.from_numpy((np.random.RandomState(4).standard_normal((8, 1, 65536)) * 0.1).astype(np.float32))
x = x.repeat(B // 8, 1, 1).cuda()
for _ in range(2):
    bands = fft_frequency_decompose(x, 4096)
    y = fft_frequency_recompose(bands, 65536)
    torch.cuda.synchronize()
