"""-m gpu: the non-filterbank multiscale pair vs golden vectors from the
unmodified reference (generator/multiscale.py:180-251, discriminator/multiscale.py:255-410)."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu
T, N = 8, 2048


def _gen(recompose):
    from music_synthesis_b200.generator.multiscale import MultiScaleGenerator
    sd = restate.multiscale_generator_state(171, N)
    g = MultiScaleGenerator(128, T, N, transposed_conv=True, recompose=recompose).eval()
    assert list(g.state_dict()) == list(sd)
    g.load_state_dict(sd)
    return g.cuda()


def test_multiscale_generator_bands_match_golden(golden):
    gold = golden("ms_generator_t8")
    with torch.no_grad():
        y = _gen(False)(synth.mel_features(172, 2, T).cuda())
    assert list(y) == restate.fb_band_sizes(N)
    for size, v in y.items():
        assert v.shape == (2, 1, size)
        e = rel_l2(v, gold[f"band_{size}"])
        print("ms generator band", size, e)
        assert e < 1e-3


def test_multiscale_generator_recomposed_matches_golden(golden):
    with torch.no_grad():
        y = _gen(True)(synth.mel_features(172, 2, T).cuda())
    assert y.shape == (2, 1, N)
    assert rel_l2(y, golden("ms_generator_recomposed_t8")["y"]) < 1e-3


def test_multiscale_generator_larger_shape_matches_oracle():
    from music_synthesis_b200.generator.multiscale import MultiScaleGenerator
    n, t = 8192, 32
    sd = restate.multiscale_generator_state(181, n)
    g = MultiScaleGenerator(128, t, n, transposed_conv=True, recompose=True).eval()
    g.load_state_dict(sd)
    x = synth.mel_features(182, 3, t)
    with torch.no_grad():
        y = g.cuda()(x.cuda())
    assert rel_l2(y, restate.multiscale_generator(x, sd, n, recompose=True)) < 1e-3


@pytest.mark.parametrize("name,kw", [
    ("ms_discriminator_cond_n2048", dict(decompose=True, channel_judgements=True, conditioning_channels=128)),
    ("ms_discriminator_k9_n2048", dict(decompose=False, channel_judgements=True, kernel_size=9))])
def test_multiscale_discriminator_matches_golden(golden, name, kw):
    from music_synthesis_b200.discriminator.multiscale import MultiScaleMultiResDiscriminator
    gd = golden(name)
    d = MultiScaleMultiResDiscriminator(N, flatten_multiscale_features=False, **kw).eval()
    dsd = restate.multiscale_discriminator_state(173, N, kw.get("conditioning_channels", 0),
                                                 kw.get("kernel_size", 41))
    assert list(d.state_dict()) == list(dsd)
    d.load_state_dict(dsd)
    d = d.cuda()
    if kw["decompose"]:
        x = (synth.randn(174, 2, 1, N) * 0.1).cuda()
    else:
        x = {s: (synth.randn(175 + i, 2, 1, s) * 0.1).cuda() for i, s in enumerate(restate.fb_band_sizes(N))}
    with torch.no_grad():
        feats, judg = d(x, synth.mel_features(180, 2, T).cuda())
    assert len(feats) == 6 and len(judg) == 6
    worst_f = worst_j = 0.0
    for i, j in enumerate(judg):
        worst_j = max(worst_j, rel_l2(j, gd[f"j{i}"]))
    for gi, fl in enumerate(feats):
        assert len(fl) == int(gd[f"n_f{gi}"])
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(gd[f"f{gi}_{i}_shape"])
            worst_f = max(worst_f, rel_l2(f.reshape(-1)[::13], gd[f"f{gi}_{i}_sub"]))
    print(name, "worst feature rel_l2", worst_f, "worst judgement rel_l2", worst_j)
    # same bars as the filter-bank discriminator (fp16 operands; judgements are cancelling sums)
    assert worst_f < 3e-3 and worst_j < 5e-3
