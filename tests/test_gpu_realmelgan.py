"""-m gpu: the official-MelGAN pair vs golden vectors from the
unmodified reference (experiment/realmelgan.py) and the oracle at larger sizes."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def _gen(seed):
    from music_synthesis_b200.experiment.realmelgan import Generator
    sd = restate.realmelgan_generator_state(seed)
    g = Generator(128, 32, n_residual_layers=3).eval()
    assert list(g.state_dict()) == list(sd)
    g.load_state_dict(sd)
    return g.cuda(), sd


def test_realmelgan_generator_matches_golden(golden):
    g, _ = _gen(101)
    with torch.no_grad():
        y = g(synth.mel_features(102, 2, 8).cuda())
    assert y.shape == (2, 1, 2048)
    err = rel_l2(y, golden("realmelgan_gen_t8")["y"])
    print("realmelgan G rel_l2", err)
    # 16-bit operands on the weight-normed G measure ~1.2-1.4e-3 in emulation
    assert err < 1e-3


@pytest.mark.parametrize("B,T", [(3, 64), (1, 33)])
def test_realmelgan_generator_matches_oracle(B, T):
    g, sd = _gen(105)
    x = synth.mel_features(106, B, T)
    with torch.no_grad():
        y = g(x.cuda())
    assert rel_l2(y, restate.realmelgan_generator(x, sd)) < 1e-3


def test_realmelgan_discriminator_matches_golden(golden):
    from music_synthesis_b200.experiment.realmelgan import Discriminator
    gd = golden("realmelgan_disc_n4096")
    sd = restate.realmelgan_discriminator_state(103)
    d = Discriminator(3, 16, 4, 4).eval()
    assert list(d.state_dict()) == list(sd)
    d.load_state_dict(sd)
    d = d.cuda()
    with torch.no_grad():
        feats, judg = d((synth.randn(104, 2, 1, 4096) * 0.1).cuda(), None)
    assert [j.shape[-1] for j in judg] == [16, 8, 4]
    for i, j in enumerate(judg):
        assert rel_l2(j, gd[f"j{i}"]) < 5e-3
        for k, f in enumerate(feats[i]):
            assert tuple(f.shape) == tuple(gd[f"f{i}_{k}_shape"])
            assert rel_l2(f.reshape(-1)[::41], gd[f"f{i}_{k}_sub"]) < (1e-5 if k < 5 else 2e-3)


def test_realmelgan_losses_match_oracle():
    from music_synthesis_b200.experiment.realmelgan import Discriminator, mel_gan_gen_loss
    sd = restate.realmelgan_discriminator_state(103)
    d = Discriminator(3, 16, 4, 4).eval()
    d.load_state_dict(sd)
    d = d.cuda()
    a, b = synth.randn(107, 2, 1, 8192) * 0.1, synth.randn(108, 2, 1, 8192) * 0.1
    with torch.no_grad():
        f1, j1 = d(a.cuda(), None)
        f2, j2 = d(b.cuda(), None)
    rf1, rj1 = restate.realmelgan_discriminator(a, sd)
    rf2, rj2 = restate.realmelgan_discriminator(b, sd)
    ref = sum(restate.hinge_generator_loss(f) for f in rj2) + 10 * restate.real_mel_gan_feature_loss(rf1, rf2)
    got = mel_gan_gen_loss(f1, f2, j1, j2)
    assert abs(float(got) - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))
