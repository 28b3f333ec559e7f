"""-m gpu: FFT octave-band split / merge vs the golden vectors of the
unmodified reference and vs the oracle at BASELINE config-5 size."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def test_bands_match_golden(golden):
    from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
    g = golden("fft_bands_n8192")
    x = synth.randn(51, 2, 1, 8192) * 0.1
    bands = fft_frequency_decompose(x.cuda(), 512)
    assert list(bands) == [512, 1024, 2048, 4096, 8192]
    for k, v in bands.items():
        assert v.shape == (2, 1, k)
        assert rel_l2(v, g[f"band_{k}"]) < 5e-6
    rec = fft_frequency_recompose(bands, 8192)
    assert rel_l2(rec, g["recomposed"]) < 5e-6


@pytest.mark.parametrize("B,N,min_size", [(4, 65536, 4096), (3, 2048, 64), (1, 4096, 4096)])
def test_bands_match_oracle(B, N, min_size):
    from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
    x = synth.randn(52, B, 1, N) * 0.1
    ref = restate.fft_frequency_decompose(x, min_size)
    got = fft_frequency_decompose(x.cuda(), min_size)
    assert list(got) == list(ref)
    for k in ref:
        assert rel_l2(got[k], ref[k]) < 5e-6
    rr = restate.fft_frequency_recompose(ref, N)
    gr = fft_frequency_recompose(got, N)
    assert rel_l2(gr, rr) < 5e-6
    # size-independent property: the split is linear
    y = synth.randn(53, B, 1, N) * 0.1
    gy = fft_frequency_decompose(y.cuda(), min_size)
    gxy = fft_frequency_decompose((x + 2 * y).cuda(), min_size)
    for k in got:
        assert rel_l2(gxy[k], got[k] + 2 * gy[k]) < 5e-6
