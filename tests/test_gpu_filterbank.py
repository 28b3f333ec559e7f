"""-m gpu: Morlet filter-bank analysis / synthesis vs the golden
vectors produced through the reference harness."""
import numpy as np
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def test_bank_construction_matches_reference_fixture(golden):
    from music_synthesis_b200.audio.filterbank import (SampleRate, linear_center_frequencies,
                                                       morlet_bank)
    g = golden("filterbank_n1024")
    sr = SampleRate(22050)
    mine = morlet_bank(sr, 128, linear_center_frequencies(sr.nyquist / 2, sr.nyquist, 128), 0.05)
    assert np.abs(mine - g["bank"][:, 0, :]).max() < 2e-7


def test_analysis_and_synthesis_match_golden(golden):
    from music_synthesis_b200.audio.filterbank import FilterBank
    g = golden("filterbank_n1024")
    fb = FilterBank(None, 128, None, bank=g["bank"][:, 0, :]).to("cuda")
    x = synth.randn(61, 2, 1, 1024) * 0.1
    conv = fb.convolve(x.cuda())
    assert tuple(conv.shape) == tuple(g["conv_shape"]) == (2, 128, 1025)
    assert rel_l2(conv.reshape(-1)[::29], g["conv_sub"]) < 1e-3
    # synthesis on the reference's own analysis output (same input on both sides)
    ref_conv = restate.filterbank_convolve(x, torch.from_numpy(g["bank"]))
    back = fb.transposed_convolve(ref_conv.cuda())
    assert back.shape == (2, 1, 1024)
    assert rel_l2(back, g["back"]) < 1e-3


@pytest.mark.parametrize("B,L", [(1, 4096), (3, 1000), (2, 129)])
def test_filterbank_vs_oracle(B, L):
    from music_synthesis_b200.audio.filterbank import FilterBank, SampleRate, linear_center_frequencies
    sr = SampleRate(22050) * 4
    fb = FilterBank(sr, 128, linear_center_frequencies(sr.nyquist / 2, sr.nyquist, 128)).to("cuda")
    bank = fb.filter_bank.cpu()
    x = synth.randn(62, B, 1, L) * 0.1
    ref = restate.filterbank_convolve(x, bank)
    got = fb.convolve(x.cuda())
    assert got.shape == ref.shape == (B, 128, L + 1)
    assert rel_l2(got, ref) < 1e-3
    z = synth.randn(63, B, 128, L + 1) * 0.1
    rb = restate.filterbank_transposed_convolve(z, bank)
    gb = fb.transposed_convolve(z.cuda())
    assert gb.shape == rb.shape == (B, 1, L)
    assert rel_l2(gb, rb) < 1e-3


def _fb_generator(recompose):
    from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
    g = FilterBankMultiScaleGenerator(22050, 128, 8, 2048, recompose=recompose).eval()
    g.load_state_dict(restate.fb_generator_state(71, 2048))
    return g.cuda()


def test_fb_multiscale_generator_matches_golden(golden):
    """SURVEY section 8 row a10 vs the unmodified reference's output."""
    gold = golden("fb_generator_t8")
    g = _fb_generator(False)
    assert list(g.state_dict()) == list(restate.fb_generator_state(71, 2048))
    # the product's own bank construction equals the banks the reference modules built
    for size, cg in g.channel_generators.items():
        chk = float(cg.filter_bank.filter_bank.double().abs().sum())
        assert abs(chk - float(gold[f"bank_checksum_{size}"])) < 1e-4
    x = synth.mel_features(72, 2, 8)
    with torch.no_grad():
        y = g(x.cuda())
    assert list(y) == [2048, 1024, 512, 256, 128]
    for k, v in y.items():
        assert v.shape == (2, 1, k)
        err = rel_l2(v, gold[f"band_{k}"])
        print("band", k, "rel_l2", err)
        assert err < 1e-3
    with torch.no_grad():
        r = _fb_generator(True)(x.cuda())
    assert rel_l2(r, golden("fb_generator_recomposed_t8")["y"]) < 1e-3


def test_fb_multiscale_generator_cfg5_size_vs_oracle():
    """BASELINE config-5 geometry (T=256 -> 65536 samples), 2 clips."""
    from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
    sd = restate.fb_generator_state(73, 65536)
    g = FilterBankMultiScaleGenerator(22050, 128, 256, 65536, recompose=False).eval()
    g.load_state_dict(sd)
    g = g.cuda()
    x = synth.mel_features(74, 2, 256)
    with torch.no_grad():
        y = g(x.cuda())
    ref = restate.filterbank_multiscale_generator(x, sd, restate.fb_banks(), 65536)
    assert list(y) == list(ref) == [65536, 32768, 16384, 8192, 4096]
    for k in ref:
        assert rel_l2(y[k], ref[k]) < 1e-3


def test_space_to_depth_strided_conv_equals_strided_conv():
    """stride-s k7 conv == stride-1 conv over the space-to-depth input (exact re-indexing)."""
    from music_synthesis_b200 import ops
    from torch.nn import functional as F
    from tests.gpu_util import rnd16, randn
    for s, L in ((4, 1000), (2, 333), (4, 17)):
        x, w, b = randn(1, 2, 128, L), randn(2, 128, 128, 7, scale=0.05), randn(3, 128, scale=0.1)
        ref = F.leaky_relu(F.conv1d(rnd16(x).double(), rnd16(w).double(), b.double(), stride=s, padding=3), 0.2)
        x16 = ops.pack_ncl(x.cuda())
        xs = ops.space_to_depth(x16, s)
        w1, taps, pad = ops.strided_conv_weight(w.cuda(), s)
        lx = xs.shape[2]
        d = ops.conv_desc(ops.MS_CONV, 2, s * 128, 128, lx, taps, 1, pad, leaky=True,
                          crop=lx + 2 * pad - (taps - 1) - lx)
        _, y32 = ops.conv_fwd(d, xs, ops.pack_conv_weight(d, w1), b.cuda(), want16=False, want32=True)
        got = ops.unpack_blk32(y32).cpu()
        assert got.shape == ref.shape, (got.shape, ref.shape)
        assert rel_l2(got, ref) < 2e-5


def test_fb_multiscale_discriminator_matches_golden(golden):
    """SURVEY section 8 row a11 vs the unmodified reference's outputs."""
    from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
    g = golden("fb_discriminator_n2048")
    sd = restate.fb_discriminator_state(81, 2048)
    d = FilterBankMultiScaleDiscriminator(2048, 22050, decompose=False, conditioning_channels=128).eval()
    assert list(d.state_dict()) == list(sd)
    d.load_state_dict(sd)
    d = d.cuda()
    bands = {s: (synth.randn(82 + i, 2, 1, s) * 0.1).cuda()
             for i, s in enumerate(restate.fb_band_sizes(2048))}
    feat = synth.mel_features(90, 2, 8).cuda()
    with torch.no_grad():
        feats, judg = d(bands, feat)
    assert len(judg) == 6 and [len(f) for f in feats] == [7, 7, 7, 7, 7, 3]
    for i, j in enumerate(judg):
        assert j.shape == (2, 1, 8)
        assert rel_l2(j, g[f"j{i}"]) < 5e-3, i
    for gi, fl in enumerate(feats):
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(g[f"f{gi}_{i}_shape"])
            assert rel_l2(f.reshape(-1)[::13], g[f"f{gi}_{i}_sub"]) < 3e-3, (gi, i)
    # decompose=True path: full-band audio in, FFT octave split on the device
    d2 = FilterBankMultiScaleDiscriminator(2048, 22050, decompose=True, conditioning_channels=128).eval()
    d2.load_state_dict(sd)
    d2 = d2.cuda()
    audio = synth.randn(95, 2, 1, 2048) * 0.1
    with torch.no_grad():
        f2, j2 = d2(audio.cuda(), feat)
    rb = restate.fft_frequency_decompose(audio, 128)
    rf, rj = restate.filterbank_multiscale_discriminator(rb, feat.cpu(), sd, restate.fb_banks(), 2048)
    for a, b in zip(j2, rj):
        assert rel_l2(a, b) < 5e-3
