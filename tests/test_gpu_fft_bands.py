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


def test_merge_of_a_sparse_set_of_bands():
    """fft_frequency_recompose takes any dictionary of bands (audio/transform.py:85-115): sizes
    with gaps between them, smaller than the output, in any order -- the merge loader's band
    lookup must keep exactly the bins the reference keeps."""
    from music_synthesis_b200.audio.transform import fft_frequency_recompose
    for sizes, N in (((256, 1024, 4096), 8192), ((2048, 64), 2048), ((512,), 4096)):
        bands = {s: synth.randn(60 + s % 7, 3, 1, s) * 0.1 for s in sizes}
        want = restate.fft_frequency_recompose(bands, N)
        got = fft_frequency_recompose({s: v.cuda() for s, v in bands.items()}, N)
        assert rel_l2(got, want) < 5e-6, (sizes, N)


def test_multiscale_representation_matches_golden(golden):
    """`MultiScale.from_audio / to_audio` (audio/representation.py:82-103)
    on the GPU.  The reference's MultiScale run on this input reproduces the fft_bands_n8192
    fixture bit for bit (tests/test_oracle_golden.py pins that in the build container)."""
    from music_synthesis_b200.audio.representation import MultiScale, RawAudio
    g = golden("fft_bands_n8192")
    x = (synth.randn(51, 2, 1, 8192) * 0.1).numpy()
    ms = MultiScale.from_audio(x, 22050)
    assert sorted(ms.data) == [512, 1024, 2048, 4096, 8192]
    for k, v in ms.data.items():
        assert v.shape == (2, 1, k) and v.dtype.name == "float32"
        assert rel_l2(torch.from_numpy(v), g[f"band_{k}"]) < 5e-6
    audio = ms.to_audio()
    assert audio.shape == (2, 8192)
    assert rel_l2(torch.from_numpy(audio), g["recomposed"].reshape(2, 8192)) < 5e-6
    dev = MultiScale.from_audio(x, 22050, device_bands=True)
    assert all(v.is_cuda for v in dev.data.values())
    assert torch.equal(dev.data[512].cpu(), torch.from_numpy(ms.data[512]))
    assert RawAudio.from_audio(x, 22050).to_audio().shape == (2, 8192)


@pytest.mark.parametrize("knob", ["MSB_FFT_FUSE=0", "MSB_FFT_TABLE=0", "MSB_FFT_STAGED=0", "MSB_FFT_MERGE_GATHER=0", "MSB_FFT_PACKED=0", "MSB_FFT_LEGACY=1"])
def test_pass_variants_agree(monkeypatch, knob):
    """The library reads its knobs per call: one pass per launch (MSB_FFT_FUSE=0) agrees with the
    default two passes per launch to rounding; twiddles computed in the loaders (MSB_FFT_TABLE=0)
    are the same bits as the per-call table; the first radix-16 pass with direct stores
    (MSB_FFT_STAGED=0) is bit-identical to the shared-memory staged one; full-length complex
    transforms (MSB_FFT_PACKED=0) and the first-version radix-4 path (MSB_FFT_LEGACY=1) agree
    with the default real-packed half-length transforms to rounding."""
    from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
    x = (synth.randn(54, 3, 1, 65536) * 0.1).cuda()
    a = fft_frequency_decompose(x, 4096)
    ra = fft_frequency_recompose(a, 65536)
    name, value = knob.split("=")
    monkeypatch.setenv(name, value)
    b = fft_frequency_decompose(x, 4096)
    rb = fft_frequency_recompose(b, 65536)
    for k in a:
        if name in ("MSB_FFT_STAGED", "MSB_FFT_TABLE"):
            assert torch.equal(a[k], b[k])
        else:
            assert rel_l2(a[k], b[k]) < 3e-6
    assert rel_l2(ra, rb) < 3e-6


@pytest.mark.parametrize("B,N,min_size", [(2, 256, 16), (3, 8192, 512), (2, 65536, 4096)])
def test_band_split_and_merge_are_differentiable(B, N, min_size):
    """Training with decompose=True / recompose=True (generator/multiscale.py:166-178,
    discriminator/multiscale.py:212-252): gradients of both maps against torch autograd on the
    oracle restatement (fp64).  The adjoint of the split is the merge up to one bin per band."""
    from music_synthesis_b200.audio.transform import (fft_frequency_decompose,
                                                      fft_frequency_recompose)
    from oracle import restate
    from tests.gpu_util import randn, rel_l2
    x = randn(700 + N % 97, B, 1, N)
    with torch.enable_grad():
        # ---- split
        xr = x.double().requires_grad_(True)
        ref = restate.fft_frequency_decompose(xr, min_size)
        g = {s: randn(710 + i, B, 1, s) for i, s in enumerate(ref)}
        sum((ref[s] * g[s].double()).sum() for s in ref).backward()
        xd = x.cuda().requires_grad_(True)
        got = fft_frequency_decompose(xd, min_size)
        assert list(got) == list(ref)
        sum((got[s] * g[s].cuda()).sum() for s in got).backward()
        assert rel_l2(xd.grad, xr.grad) < 2e-5
        # ---- merge (all bands, and a subset with a gap)
        for keep in (list(ref), list(ref)[::2]):
            br = {s: g[s].double().requires_grad_(True) for s in keep}
            gy = randn(720, B, 1, N)
            (restate.fft_frequency_recompose(br, N) * gy.double()).sum().backward()
            bd = {s: g[s].cuda().requires_grad_(True) for s in keep}
            (fft_frequency_recompose(bd, N) * gy.cuda()).sum().backward()
            for s in keep:
                assert rel_l2(bd[s].grad, br[s].grad) < 2e-5, s
    # without a tape both stay on the plain kernels
    assert not fft_frequency_decompose(x.cuda(), min_size)[min_size].requires_grad
