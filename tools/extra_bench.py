"""Timings of the other hot-path rows (CUDA events, warm, back-to-back), one JSON line each."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200.feature.feature import Audio2Mel
from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
from oracle import restate, synth

torch.set_grad_enabled(False)
PEAK_HBM = 6536.7e9
PEAK_TF = 1627.2e12

def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n

out = []
a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
for B in (64, 4096):
    a = synth.uniform_audio(3, min(B, 64), 16384).repeat(B // min(B, 64), 1, 1).cuda()
    t = timeit(lambda: a2m(a))
    bytes_ = 4 * B * 16384 + 4 * B * 128 * 62
    out.append({"row": "a1 Audio2Mel", "workload": "B=%d x 16384 samples (cfg2%s)" % (B, "" if B == 64 else " scaled"),
                "us": t * 1e6, "samples_per_s": B * 16384 / t, "bound": "hbm", "achieved_GBs": bytes_ / t / 1e9,
                "frac": bytes_ / t / PEAK_HBM})
x = (synth.randn(4, 8, 1, 65536) * 0.1).repeat(8, 1, 1).cuda()
t = timeit(lambda: fft_frequency_decompose(x, 4096))
out.append({"row": "a8 fft_frequency_decompose", "workload": "B=64 x 65536, 5 bands (cfg5 preprocess)", "us": t * 1e6,
            "bound": "hbm", "achieved_GBs": (x.numel() * 4 * (1 + 31 / 16)) / t / 1e9})
g = FilterBankMultiScaleGenerator(22050, 128, 256, 65536, recompose=False).eval()
g.load_state_dict(restate.fb_generator_state(73, 65536)); g = g.cuda()
feat = synth.mel_features(74, 8, 256).cuda()
t = timeit(lambda: g(feat), n=10)
out.append({"row": "a10 FilterBankMultiScaleGenerator fwd", "workload": "8 clips x 256 frames -> 5 bands (cfg5 per-GPU shard)",
            "us": t * 1e6, "bound": "tensor", "achieved_TFLOPs": 17.172e9 * 8 / t / 1e12, "frac": 17.172e9 * 8 / t / PEAK_TF})
d = FilterBankMultiScaleDiscriminator(65536, 22050, decompose=False, conditioning_channels=128).eval()
d.load_state_dict(restate.fb_discriminator_state(81, 65536)); d = d.cuda()
bands = g(feat)
t = timeit(lambda: d(bands, feat), n=10)
out.append({"row": "a11 FilterBankMultiScaleDiscriminator fwd", "workload": "8 clips x 65536 (cfg5 per-GPU shard)",
            "us": t * 1e6, "bound": "tensor", "achieved_TFLOPs": 16.704e9 * 8 / t / 1e12, "frac": 16.704e9 * 8 / t / PEAK_TF})
md = MelGanDiscriminator().eval()
md.load_state_dict(restate.melgan_discriminator_state(41)); md = md.cuda()
xa = (synth.randn(5, 32, 1, 8192) * 0.1).cuda()
t = timeit(lambda: md(xa), n=10)
out.append({"row": "a6 MelGanDiscriminator fwd", "workload": "32 clips x 8192 (cfg4 batch)", "us": t * 1e6,
            "bound": "tensor", "achieved_TFLOPs": 0.861e9 * 32 / t / 1e12})
for o in out:
    print(json.dumps(o))
