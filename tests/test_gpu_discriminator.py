"""-m gpu: discriminator forward parity against the golden vectors
produced by the unmodified reference and against the oracle restatement."""
import pytest
import torch
from torch.nn import functional as F

from oracle import restate, synth
from tests.gpu_util import rel_l2, randn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", [
    # (B, cin, cout, L, k, stride, pad, groups)
    (2, 1, 16, 1000, 15, 1, 7, 1),
    (2, 16, 64, 1000, 41, 4, 20, 4),
    (1, 64, 256, 777, 41, 4, 20, 16),
    (2, 256, 1024, 300, 41, 4, 20, 64),
    (1, 1024, 1024, 70, 41, 4, 20, 256),
    (2, 8, 8, 50, 3, 2, 1, 2),
])
def test_direct_conv(case):
    from music_synthesis_b200 import ops
    B, ci, co, L, k, s, p, g = case
    x, w, b = randn(1, B, ci, L), randn(2, co, ci // g, k, scale=0.05), randn(3, co, scale=0.1)
    ref = F.leaky_relu(F.conv1d(x.double(), w.double(), b.double(), stride=s, padding=p, groups=g), 0.2)
    got = ops.conv1d_direct(x.cuda(), w.cuda(), b.cuda(), s, p, g, leaky=True).cpu()
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < 2e-6


@pytest.mark.parametrize("L", [16384, 8193, 33, 5])
def test_avg_pool(L):
    from music_synthesis_b200 import ops
    x = randn(4, 3, 2, L)
    ref = F.avg_pool1d(x, 4, 2, 2)
    got = ops.avg_pool1d(x.cuda(), 4, 2, 2).cpu()
    assert got.shape == ref.shape == (3, 2, L // 2 + 1)
    assert torch.allclose(got, ref, atol=1e-6)


def _disc():
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    sd = restate.randomize_biases(restate.melgan_discriminator_state(41), 1041)
    d = MelGanDiscriminator().eval()
    d.load_state_dict(sd)
    return d.cuda(), sd


def test_melgan_discriminator_matches_golden(golden):
    g = golden("disc_melgan_n4096")
    d, _ = _disc()
    x = synth.randn(42, 2, 1, 4096) * 0.1
    with torch.no_grad():
        feats, judg = d(x.cuda())
    assert [j.shape[-1] for j in judg] == [16, 9, 5]
    for s, (fl, j) in enumerate(zip(feats, judg)):
        assert len(fl) == 6
        assert rel_l2(j, g[f"j{s}"]) < 5e-3
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(g[f"f{s}_{i}_shape"])
            tol = 1e-5 if i < 5 else 2e-3        # fp32 direct convs vs the fp16-operand GEMM
            assert rel_l2(f.reshape(-1)[::37], g[f"f{s}_{i}_sub"]) < tol, (s, i)


def test_melgan_discriminator_matches_oracle_n16384():
    d, sd = _disc()
    x = synth.randn(44, 2, 1, 16384) * 0.1
    with torch.no_grad():
        feats, judg = d(x.cuda())
    rf, rj = restate.melgan_discriminator(x, sd)
    assert [j.shape[-1] for j in judg] == [64, 33, 17]
    for a, b in zip(judg, rj):
        # scores are a cancelling sum over 3072 products of the fp16-operand layer's output
        assert rel_l2(a, b) < 5e-3
    for fl, rl in zip(feats, rf):
        for a, b in zip(fl, rl):
            assert a.shape == b.shape and rel_l2(a, b) < 2e-3
    # the losses the trainers compute from these outputs agree too (loss.py:21-79)
    x2 = synth.randn(45, 2, 1, 16384) * 0.1
    with torch.no_grad():
        f2, j2 = d(x2.cuda())
    rf2, rj2 = restate.melgan_discriminator(x2, sd)
    cpu = lambda t: [[u.cpu() for u in v] for v in t] if isinstance(t[0], list) else [u.cpu() for u in t]
    got = restate.mel_gan_gen_loss(cpu(feats), cpu(f2), cpu(judg), cpu(j2))
    ref = restate.mel_gan_gen_loss(rf, rf2, rj, rj2)
    assert abs(float(got) - float(ref)) < 2e-3 * max(1.0, abs(float(ref)))


def test_discriminator_state_dict_layout():
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    ref = restate.melgan_discriminator_state(0)
    sd = MelGanDiscriminator().state_dict()
    assert list(sd) == list(ref)
    for k in sd:
        assert tuple(sd[k].shape) == tuple(ref[k].shape)


def test_loss_functions_match_golden(golden):
    """loss.py mirrors on the GPU discriminator outputs vs the reference's loss values."""
    from music_synthesis_b200.loss import loss as L
    g = golden("losses_melgan")
    d, _ = _disc()
    with torch.no_grad():
        f1, j1 = d((synth.randn(42, 2, 1, 4096) * 0.1).cuda())
        f2, j2 = d((synth.randn(43, 2, 1, 4096) * 0.1).cuda())
    got = dict(
        disc_hinge=L.mel_gan_disc_loss(j1, j2),
        disc_lsq=L.mel_gan_disc_loss(j1, j2, gan_loss=L.least_squares_disc_loss),
        feature=L.mel_gan_feature_loss(f1, f2),
        gen_hinge=L.mel_gan_gen_loss(f1, f2, j1, j2),
        gen_lsq=L.mel_gan_gen_loss(f1, f2, j1, j2, gan_loss=L.least_squares_generator_loss))
    for k, v in got.items():
        assert v.is_cuda and v.dim() == 0
        assert abs(float(v) - float(g[k])) <= 1e-3 * max(1.0, abs(float(g[k]))), (k, float(v), float(g[k]))


def test_loss_reductions_exact_on_same_inputs():
    """Same tensors on both sides: only fp32-vs-double accumulation differs."""
    from music_synthesis_b200.loss import loss as L
    a, b = randn(7, 3, 5, 1001), randn(8, 3, 5, 1001)
    ac, bc = a.cuda(), b.cuda()
    ref = [restate.hinge_discriminator_loss(a, b), restate.hinge_generator_loss(a),
           restate.least_squares_disc_loss(a, b), restate.least_squares_generator_loss(a),
           F.l1_loss(a, b)]
    got = [L.hinge_discriminator_loss(ac, bc), L.hinge_generator_loss(ac),
           L.least_squares_disc_loss(ac, bc), L.least_squares_generator_loss(ac),
           L.mel_gan_feature_loss([[ac]], [[bc]])]
    for r, v in zip(ref, got):
        assert abs(float(r) - float(v)) < 2e-6 * max(1.0, abs(float(r)))
    # deterministic: bit-identical on repeat
    assert float(L.mel_gan_feature_loss([[ac]], [[bc]])) == float(got[4])
