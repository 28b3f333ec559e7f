"""-m gpu: MelGanGenerator / Audio2Mel parity through the module mirrors (which call
the C ABI) against (a) the committed golden vectors produced by the unmodified
reference and (b) the oracle restatement on larger seeded inputs.

Tolerances (north star): waveform relative L2 <= 1e-3; log-mel max-abs <= 1e-3."""
import numpy as np
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2
from music_synthesis_b200._lib import MsbError

pytestmark = pytest.mark.gpu

WAVEFORM_TOL = 1e-3
LOGMEL_TOL = 1e-3


def _module(sd):
    from music_synthesis_b200.generator.full import MelGanGenerator
    g = MelGanGenerator(64, 128).eval()
    g.load_state_dict(sd)
    return g.cuda()


@pytest.mark.parametrize("name", ["gen_b2_t8", "gen_b2_t8_bias", "gen_b3_t20_bias",
                                  "gen_cfg1_b1_t64"])
def test_generator_matches_golden(golden, name):
    g = golden(name)
    seed, B, T = int(g["seed"]), int(g["B"]), int(g["T"])
    sd = restate.melgan_generator_state(seed)
    if int(g["biased"]):
        sd = restate.randomize_biases(sd, seed + 1000)
    m = _module(sd)
    with torch.no_grad():
        y = m(synth.mel_features(seed, B, T).cuda())
    torch.cuda.synchronize()
    assert y.shape == (B, 1, 256 * T)
    err = rel_l2(y, g["y"])
    print(name, "rel_l2 =", err)
    assert err < WAVEFORM_TOL


def test_generator_matches_oracle_multi_pass():
    """More clips than one workspace pass holds, odd T: exercises the pass loop and
    the ragged last time tile of every layer."""
    sd = restate.randomize_biases(restate.melgan_generator_state(77), 1077)
    m = _module(sd)
    m.clips_per_pass = 2
    x = synth.mel_features(78, 5, 37)
    with torch.no_grad():
        y = m(x.cuda())
        ref = restate.melgan_generator(x, sd)
    err = rel_l2(y, ref)
    print("multi-pass rel_l2 =", err)
    assert err < WAVEFORM_TOL
    # per-clip results must not depend on how the batch was split into passes
    m.clips_per_pass = 16
    with torch.no_grad():
        y2 = m(x.cuda())
    assert torch.equal(y, y2)


def test_generator_reacts_to_weight_updates():
    sd = restate.melgan_generator_state(5)
    m = _module(sd)
    x = synth.mel_features(6, 1, 8).cuda()
    with torch.no_grad():
        y0 = m(x).clone()
        m.main[1].weight.mul_(0.5)          # in-place update must invalidate the pack
        y1 = m(x)
    assert not torch.equal(y0, y1)
    sd2 = dict(sd)
    sd2["main.1.weight"] = sd["main.1.weight"] * 0.5
    assert rel_l2(y1, restate.melgan_generator(x.cpu(), sd2)) < WAVEFORM_TOL


def test_generator_refuses_cpu_and_feature_gradients():
    from music_synthesis_b200._lib import MsbError
    m = _module(restate.melgan_generator_state(5))
    with pytest.raises(MsbError):
        m(torch.zeros(1, 128, 8))                      # CPU input: no CPU path
    with torch.enable_grad(), pytest.raises(MsbError):
        m(torch.zeros(1, 128, 8, device="cuda", requires_grad=True))   # no d/d(features)


def test_generator_training_forward_equals_inference_forward():
    """grad mode records the layer-wise autograd path; its output must agree with the fused
    inference kernels (same fp16 operands / fp32 accumulation, different tiling)"""
    m = _module(restate.melgan_generator_state(5))
    x = synth.mel_features(6, 2, 8).cuda()
    with torch.no_grad():
        y0 = m(x)
    with torch.enable_grad():
        y1 = m(x)
    assert y1.requires_grad and rel_l2(y1, y0) < 2e-4


@pytest.mark.parametrize("C,L", [(32, 96), (128, 64)])
def test_residual_stack_module_matches_golden(golden, C, L):
    from music_synthesis_b200.util.modules import ResidualStack
    g = golden(f"resstack_c{C}")
    seed = int(g["seed"])
    sd = synth.residual_stack_state(seed, C)
    rs = ResidualStack(C, [1, 3, 9]).eval()
    rs.load_state_dict({k[len("s."):]: v for k, v in sd.items()})
    rs = rs.cuda()
    with torch.no_grad():
        y = rs(synth.randn(seed + 1, 2, C, L).cuda())
    assert rel_l2(y, g["y"]) < WAVEFORM_TOL


@pytest.mark.parametrize("name", ["a2m_b2_n16384", "a2m_b3_n4000"])
def test_audio2mel_matches_golden(golden, name):
    from music_synthesis_b200.feature.feature import Audio2Mel
    g = golden(name)
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    b = golden("mel_basis_22050_1024_128")
    assert np.array_equal(a2m.mel_basis.cpu().numpy(), b["mel_basis"])
    a = synth.uniform_audio(int(g["seed"]), int(g["B"]), int(g["N"]))
    y = a2m(a.cuda()).cpu().numpy()
    assert y.shape == g["y"].shape
    err = np.abs(y - g["y"]).max()
    print(name, "max |d log10 mel| =", err)
    assert err < LOGMEL_TOL


def test_audio2mel_cfg2_and_edge_cases():
    from music_synthesis_b200.feature.feature import Audio2Mel
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    a = synth.uniform_audio(3, 64, 16384)                       # BASELINE config 2
    y = a2m(a.cuda()).cpu()
    ref = restate.audio2mel(a, a2m.mel_basis.cpu(), a2m.window.cpu())
    assert y.shape == (64, 128, 62)
    assert (y - ref).abs().max() < LOGMEL_TOL
    # silence hits the 1e-5 clamp exactly; numpy input path; shortest legal clip
    z = a2m(torch.zeros(1, 1, 2048, device="cuda")).cpu()
    assert torch.all(z == -5.0)
    n = a2m(a[0, 0].numpy()).cpu()
    assert (n - ref[:1]).abs().max() < LOGMEL_TOL
    s = a2m(a[:2, :, :640].contiguous().cuda())
    assert s.shape == (2, 128, 1)
    assert (s.cpu() - restate.audio2mel(a[:2, :, :640], a2m.mel_basis.cpu(),
                                        a2m.window.cpu())).abs().max() < LOGMEL_TOL


def test_generate_host_pipeline_matches_forward():
    """The chunked, copy-overlapped host-to-host path returns exactly forward()'s result."""
    sd = restate.melgan_generator_state(9)
    m = _module(sd)
    x = synth.mel_features(10, 7, 12)
    with torch.no_grad():
        ref = m(x.cuda()).cpu()
    # no synchronize here: generate() is host-to-host, its result must be readable on return
    got = m.generate(x.pin_memory(), chunk_clips=3)
    assert got.shape == ref.shape and torch.equal(got, ref)
    got2 = m.generate(x, chunk_clips=64)          # pageable input, single chunk
    assert torch.equal(got2, ref)
    got3 = m.generate(x.pin_memory(), chunks=[1, 4, 2])     # explicit chunk sizes
    assert torch.equal(got3, ref)
    with pytest.raises(MsbError):
        m.generate(x, chunks=[3, 3])


def test_generate_sees_every_kind_of_weight_change():
    """generate() checks the packed weights against the Parameter objects of the previous call
    before its first kernel and walks the module tree only while the GPU works: an in-place
    update, a load_state_dict and a REPLACED Parameter object must all reach the output."""
    m = _module(restate.melgan_generator_state(9))
    x = synth.mel_features(10, 5, 12)
    xp = x.pin_memory()

    def both():
        with torch.no_grad():
            ref = m(x.cuda()).cpu()
        return ref, m.generate(xp, chunk_clips=2, edge_clips=1)

    ref0, got0 = both()
    assert torch.equal(got0, ref0)
    with torch.no_grad():
        m.main[15].bias.add_(0.25)                                  # in place: version bump
    got1 = m.generate(xp, chunk_clips=2, edge_clips=1)
    assert not torch.equal(got1, got0)
    with torch.no_grad():
        assert torch.equal(got1, m(x.cuda()).cpu())
    m.main[15].bias = torch.nn.Parameter(m.main[15].bias.detach() - 0.5)   # new Parameter object
    got2 = m.generate(xp, chunk_clips=2, edge_clips=1)              # before any forward() call
    with torch.no_grad():
        ref2 = m(x.cuda()).cpu()
    assert torch.equal(got2, ref2) and not torch.equal(got2, got1)
    m.load_state_dict(restate.melgan_generator_state(11))
    ref3, got3 = both()
    assert torch.equal(got3, ref3) and not torch.equal(got3, got2)


def test_generator_full_size_config3_properties():
    """BASELINE config 3 size (256 clips x 256 frames): oracle parity on sampled clips and the
    size-independent property that a clip's waveform does not depend on its batch-mates."""
    sd = restate.melgan_generator_state(0)
    m = _module(sd)
    x = synth.mel_features(1000, 256, 256)
    with torch.no_grad():
        y = m(x.cuda())
        alone = m(x[[0, 137, 255]].cuda())
    assert y.shape == (256, 1, 65536)
    assert torch.isfinite(y).all()
    assert torch.equal(y[[0, 137, 255]], alone)          # bit-identical per clip
    ref = restate.melgan_generator(x[[0, 255]], sd)
    err = rel_l2(y[[0, 255]], ref)
    print("config-3 sampled clips rel_l2 =", err)
    assert err < WAVEFORM_TOL


@pytest.mark.parametrize("B,T", [(1, 4), (1, 5), (2, 512), (7, 3)])
def test_generator_edge_shapes(B, T):
    """shortest legal input (reflection pad 3 needs T >= 4), odd T, the stage-one caller's
    T = 512 (featureexperiment.py:310-312); T < 4 is rejected like nn.ReflectionPad1d does."""
    from music_synthesis_b200._lib import MsbError
    sd = restate.randomize_biases(restate.melgan_generator_state(3), 1003)
    m = _module(sd)
    x = synth.mel_features(5, B, T)
    if T < 4:
        with pytest.raises(MsbError), torch.no_grad():
            m(x.cuda())
        return
    with torch.no_grad():
        y = m(x.cuda())
    assert y.shape == (B, 1, 256 * T)
    assert rel_l2(y, restate.melgan_generator(x, sd)) < WAVEFORM_TOL


def test_audio2mel_banded_projection_is_bit_identical_to_dense():
    from music_synthesis_b200 import ops
    from music_synthesis_b200.feature.feature import Audio2Mel
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    a = synth.uniform_audio(8, 5, 12000).cuda()
    banded = a2m(a)
    dense = ops.audio2mel(a, a2m.window, a2m.mel_basis, 1024, 256, None)
    assert torch.equal(banded, dense)
    # a basis with arbitrary (non-banded) rows still works through the ranges path
    a2m.mel_basis[3, 500] = 0.01
    a2m.mel_basis[3, 2] = 0.02
    assert torch.equal(a2m(a), ops.audio2mel(a, a2m.window, a2m.mel_basis, 1024, 256, None))


def test_audio2mel_fft_forms_agree(monkeypatch):
    """n_fft = 1024 runs the register-resident 16x16x4 FFT (csrc/a2m_fft.cuh) by default;
    MSB_A2M_RADIX4=1 (read per call) selects the shared-memory radix-4 form.  Both against the
    oracle, and against each other, on ragged shapes (last frame group partial, clip shorter
    than the padded frames)."""
    from music_synthesis_b200.feature.feature import Audio2Mel
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    for seed, B, N in ((21, 3, 16384), (22, 2, 5000), (23, 1, 700)):
        a = synth.uniform_audio(seed, B, N)
        ref = restate.audio2mel(a, a2m.mel_basis.cpu(), a2m.window.cpu())
        monkeypatch.delenv("MSB_A2M_RADIX4", raising=False)
        new = a2m(a.cuda())
        monkeypatch.setenv("MSB_A2M_RADIX4", "1")
        old = a2m(a.cuda())
        monkeypatch.delenv("MSB_A2M_RADIX4", raising=False)
        assert new.shape == ref.shape
        assert (new.cpu() - ref).abs().max().item() < 1e-4
        assert (old.cpu() - ref).abs().max().item() < 1e-4
        assert (new - old).abs().max().item() < 1e-4


def test_neural_vocoder_long_sequence():
    """the stage-one caller runs the generator through `NeuralVocoder` on
    512-frame feature sequences (featureexperiment.py:94-100) -- 131072 samples per clip."""
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.experiment.featureexperiment import NeuralVocoder
    sd = restate.melgan_generator_state(31)
    g = MelGanGenerator(512, 128).eval()
    g.load_state_dict(sd)
    vocoder = NeuralVocoder(g.cuda())
    x = synth.mel_features(32, 1, 512)
    y = vocoder(x.cuda())
    assert y.shape == (1, 1, 131072) and not y.requires_grad
    assert rel_l2(y, restate.melgan_generator(x, sd)) < WAVEFORM_TOL
