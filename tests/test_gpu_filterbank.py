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
        assert err < 2e-3
    with torch.no_grad():
        r = _fb_generator(True)(x.cuda())
    assert rel_l2(r, golden("fb_generator_recomposed_t8")["y"]) < 2e-3


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
        assert rel_l2(y[k], ref[k]) < 2e-3
