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
from music_synthesis_b200.experiment.realmelgan import Generator as RealG, Discriminator as RealD
rg = RealG(128, 32, n_residual_layers=3).eval()
rg.load_state_dict(restate.realmelgan_generator_state(101)); rg = rg.cuda()
f64 = synth.mel_features(6, 32, 64).cuda()
t = timeit(lambda: rg(f64), n=10)
out.append({"row": "a5 realmelgan.Generator fwd", "workload": "32 clips x 64 frames -> 16384 samples", "us": t * 1e6,
            "bound": "tensor", "achieved_TFLOPs": 5.804e9 * 32 / t / 1e12, "frac": 5.804e9 * 32 / t / PEAK_TF})
rd = RealD(3, 16, 4, 4).eval()
rd.load_state_dict(restate.realmelgan_discriminator_state(103)); rd = rd.cuda()
t = timeit(lambda: rd(xa, None), n=10)
out.append({"row": "a7 realmelgan.Discriminator fwd", "workload": "32 clips x 8192 (cfg4 batch)", "us": t * 1e6,
            "bound": "tensor", "achieved_TFLOPs": 0.839e9 * 32 / t / 1e12})
from music_synthesis_b200.generator.multiscale import MultiScaleGenerator
from music_synthesis_b200.discriminator.multiscale import MultiScaleMultiResDiscriminator
mg = MultiScaleGenerator(128, 256, 65536, transposed_conv=True, recompose=False).eval()
mg.load_state_dict(restate.multiscale_generator_state(171, 65536)); mg = mg.cuda()
t = timeit(lambda: mg(feat), n=10)
out.append({"row": "a12 MultiScaleGenerator fwd", "workload": "8 clips x 256 frames -> 5 bands of 65536..4096", "us": t * 1e6,
            "bound": "tensor", "achieved_TFLOPs": 23.084e9 * 8 / t / 1e12, "frac": 23.084e9 * 8 / t / PEAK_TF})
mdisc = MultiScaleMultiResDiscriminator(65536, decompose=False, channel_judgements=True,
                                        conditioning_channels=128).eval()
mdisc.load_state_dict(restate.multiscale_discriminator_state(173, 65536)); mdisc = mdisc.cuda()
mb = mg(feat)
t = timeit(lambda: mdisc(mb, feat), n=10)
out.append({"row": "a12 MultiScaleMultiResDiscriminator fwd", "workload": "8 clips x 65536 (band dict)", "us": t * 1e6,
            "bound": "tensor", "achieved_TFLOPs": 10.814e9 * 8 / t / 1e12, "frac": 10.814e9 * 8 / t / PEAK_TF})
for o in out:
    print(json.dumps(o))
