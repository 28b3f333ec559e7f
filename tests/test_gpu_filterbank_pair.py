"""-m gpu: SURVEY section 8(f) rank 1 -- FilterBankGenerator / FilterBankDiscriminator
(featuresynth/generator/filterbank.py:93-128, discriminator/filterbank.py:114-202, wired by
experiment/filterbank.py:14-128) vs golden vectors from the UNMODIFIED reference classes and vs
the oracle's restated trainers for one training cycle."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu


def _bank():
    from music_synthesis_b200.experiment.wirings import FilterBankExperiment
    return FilterBankExperiment.make_filter_bank()


def _generator(seed=401):
    from music_synthesis_b200.generator.filterbank import FilterBankGenerator
    sd = restate.filterbank_generator_state(seed)
    g = FilterBankGenerator(_bank(), 32, 8192, 128).eval()
    assert list(g.state_dict()) == list(sd)
    g.load_state_dict(sd)
    return g.cuda(), sd


def _discriminator(seed, cond):
    from music_synthesis_b200.discriminator.filterbank import FilterBankDiscriminator
    sd = restate.filterbank_discriminator_state(seed, conditioning_channels=cond)
    d = FilterBankDiscriminator(_bank(), 8192, conditioning_channels=cond).eval()
    assert list(d.state_dict()) == list(sd)
    d.load_state_dict(sd)
    return d.cuda(), sd


def test_learned_upsample_k8_s2_matches_torch():
    """ConvTranspose1d with k = 4 * stride (four polyphase taps, J - 1 = 3 tail rows)"""
    from music_synthesis_b200.util.modules import LearnedUpSample
    from torch.nn import functional as F
    from tests.gpu_util import rnd16, randn
    for cin, cout, L in ((128, 256, 37), (256, 256, 256), (64, 32, 5)):
        up = LearnedUpSample(cin, cout, 8, 2, None)
        w = randn(7, cin, cout, 8, scale=0.05)
        with torch.no_grad():
            up.conv.weight.copy_(w)
        x = randn(8, 2, cin, L)
        with torch.no_grad():
            y = up.cuda()(x.cuda())
        ref = F.leaky_relu(F.conv_transpose1d(rnd16(x).double(), rnd16(w).double(), stride=2,
                                              padding=3), 0.2)
        assert y.shape == (2, cout, 2 * L)
        assert rel_l2(y, ref) < 2e-6, (cin, cout, L)


def test_511_tap_bank_analysis_and_synthesis_match_torch():
    from torch.nn import functional as F
    from tests.gpu_util import rnd16
    fb = _bank().to("cuda")
    bank = fb.filter_bank.cpu()
    x = synth.randn(9, 2, 1, 3000) * 0.1
    a = fb.convolve(x.cuda())
    ref = F.conv1d(rnd16(x).double(), rnd16(bank).double(), padding=255)
    assert a.shape == (2, 128, 3000) and rel_l2(a, ref) < 2e-5
    h = synth.randn(10, 2, 128, 2048) * 0.1
    y = fb.transposed_convolve(h.cuda())
    ref = F.conv_transpose1d(rnd16(h).double(), rnd16(bank).double(), padding=255)
    assert y.shape == (2, 1, 2048) and rel_l2(y, ref) < 2e-5


def test_filterbank_generator_matches_golden(golden):
    g, _ = _generator()
    with torch.no_grad():
        y = g(synth.mel_features(402, 2, 32).cuda())
    assert y.shape == (2, 1, 8192)
    err = rel_l2(y, golden("filterbank_generator_t32")["y"])
    print("FilterBankGenerator rel_l2 vs reference:", err)
    assert err < 1e-3


@pytest.mark.parametrize("B,T,N", [(3, 64, 16384), (1, 256, 65536)])
def test_filterbank_generator_other_sizes_vs_oracle(B, T, N):
    from music_synthesis_b200.generator.filterbank import FilterBankGenerator
    sd = restate.filterbank_generator_state(411, T, N)
    g = FilterBankGenerator(_bank(), T, N, 128).eval()
    g.load_state_dict(sd)
    g = g.cuda()
    x = synth.mel_features(412, B, T)
    with torch.no_grad():
        y = g(x.cuda())
    ref = restate.filterbank_generator(x, sd, restate.filterbank_experiment_bank())
    err = rel_l2(y, ref)
    print("FilterBankGenerator", (B, T, N), "rel_l2", err)
    assert y.shape == (B, 1, N) and err < 1e-3


@pytest.mark.parametrize("cond", [0, 128])
def test_filterbank_discriminator_matches_golden(golden, cond):
    gd = golden("filterbank_discriminator_n8192" + ("_cond" if cond else ""))
    d, _ = _discriminator(403 + cond, cond)
    a = synth.randn(404, 2, 1, 8192) * 0.1
    feat = synth.mel_features(405, 2, 32)
    with torch.no_grad():
        feats, judg = d(a.cuda(), feat.cuda())
    assert [len(f) for f in feats] == [8, 3, 3]
    worst_f = worst_j = 0.0
    for i, j in enumerate(judg):
        assert tuple(j.shape) == tuple(gd[f"j{i}"].shape)
        worst_j = max(worst_j, rel_l2(j, gd[f"j{i}"]))
    for gi, fl in enumerate(feats):
        for i, f in enumerate(fl):
            assert tuple(f.shape) == tuple(gd[f"f{gi}_{i}_shape"])
            worst_f = max(worst_f, rel_l2(f.reshape(-1)[::53], gd[f"f{gi}_{i}_sub"]))
    print("FilterBankDiscriminator(cond=%d): worst feature rel_l2 %.3e, judgement %.3e"
          % (cond, worst_f, worst_j))
    # same bars as the multiscale discriminators (the north star states none for D outputs)
    assert worst_f < 3e-3 and worst_j < 5e-3


def test_filterbank_experiment_train_cycle_matches_oracle():
    """FilterBankExperiment wiring (least-squares sub-losses, raw audio): one D step + one G step
    through the trainer mirrors vs the oracle's restated trainers, Adam and CPU autograd"""
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    B = 4
    bank = restate.filterbank_experiment_bank()
    with torch.enable_grad():
        g, g_sd = _generator(421)
        d, d_sd = _discriminator(422, 0)
        g.train(), d.train()
        g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
        d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
        d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, least_squares_disc_loss)
        g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, least_squares_generator_loss)
        real = synth.randn(423, B, 1, 8192) * 0.1
        feats = synth.mel_features(424, B, 32)

        def gen_fn(f, sd):
            return restate.filterbank_generator(f, sd, bank)

        def disc_fn(x, f, sd):
            return restate.filterbank_discriminator(x, f, sd, bank, 0)

        rd = d_tr.train(real.cuda(), feats.cuda())
        d_loss, d_grads, d_new = restate.discriminator_train_step(
            g_sd, d_sd, real, feats, {}, sub_loss=restate.least_squares_disc_loss,
            gen_fn=gen_fn, disc_fn=disc_fn)
        assert abs(rd["d_loss"] - d_loss) < 2e-3 * abs(d_loss), (rd["d_loss"], d_loss)
        da = torch.cat([p.grad.detach().cpu().reshape(-1).double() for _, p in d.named_parameters()])
        db = torch.cat([d_grads[k].reshape(-1).double() for k, _ in d.named_parameters()])
        d_rel = float((da - db).norm() / db.norm())
        d.load_state_dict(d_new)
        rg = g_tr.train(real.cuda(), feats.cuda())
        g_loss, fake, g_grads, _ = restate.generator_train_step(
            g_sd, d_new, real, feats, {}, sub_loss=restate.least_squares_generator_loss,
            gen_fn=gen_fn, disc_fn=disc_fn)
        assert abs(rg["g_loss"] - g_loss) < 2e-3 * max(1.0, abs(g_loss)), (rg["g_loss"], g_loss)
        wave = rel_l2(rg["fake"], fake)
        ga = torch.cat([p.grad.detach().cpu().reshape(-1).double() for _, p in g.named_parameters()])
        gb = torch.cat([g_grads[k].reshape(-1).double() for k, _ in g.named_parameters()])
        g_rel = float((ga - gb).norm() / gb.norm())
    print("FilterBankExperiment cycle: fake rel_l2 %.3e, gradient vector rel_l2 D %.3e G %.3e"
          % (wave, d_rel, g_rel))
    assert wave < 1e-3
    assert d_rel < 3e-2 and g_rel < 5e-2


def test_resstack_filterbank_generator_matches_golden(golden):
    """ResidualStackFilterBankGenerator (generator/filterbank.py:8-90): weight-normed
    ConvTranspose1d layers (k 7 / stride 1 ones included), weight-normed ResidualStacks, harmonic +
    noise heads; the white-noise row is drawn from the seeded host generator as in the reference"""
    from music_synthesis_b200.generator.filterbank import ResidualStackFilterBankGenerator
    gold = golden("resstack_filterbank_generator_t8")
    sd = restate.resstack_filterbank_generator_state(411)
    g = ResidualStackFilterBankGenerator(_bank(), 8, 2048, 128, add_weight_norm=True).eval()
    assert list(g.state_dict()) == list(sd)
    g.load_state_dict(sd)
    g = g.cuda()
    torch.manual_seed(413)
    with torch.no_grad():
        y = g(synth.mel_features(412, 2, 8).cuda())
    err = rel_l2(y, gold["y"])
    print("ResidualStackFilterBankGenerator rel_l2 vs reference:", err)
    assert y.shape == (2, 1, 2048) and err < 1e-3


def test_resstack_filterbank_generator_backward_runs_and_matches_oracle():
    """gradients reach weight_g / weight_v of every layer through the weight-norm fold"""
    from music_synthesis_b200.generator.filterbank import ResidualStackFilterBankGenerator
    sd = restate.resstack_filterbank_generator_state(421)
    g = ResidualStackFilterBankGenerator(_bank(), 8, 2048, 128, add_weight_norm=True)
    g.load_state_dict(sd)
    g = g.cuda()
    x = synth.mel_features(422, 2, 8)
    target = synth.randn(423, 2, 1, 2048) * 0.1
    with torch.enable_grad():
        torch.manual_seed(424)
        y = g(x.cuda())
        loss = ((y - target.cuda()) ** 2).mean()
        loss.backward()
        leaf = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
        torch.manual_seed(424)
        raw = torch.normal(0, 1, (1, 1, 2048))
        ref = restate.resstack_filterbank_generator(x, leaf, restate.filterbank_experiment_bank(), raw)
        rl = ((ref - target) ** 2).mean()
        grads = torch.autograd.grad(rl, list(leaf.values()))
    assert abs(float(loss) - float(rl)) < 2e-3 * abs(float(rl))
    a = torch.cat([p.grad.detach().cpu().reshape(-1).double() for _, p in g.named_parameters()])
    b = torch.cat([gr.reshape(-1).double() for gr in grads])
    rel = float((a - b).norm() / b.norm())
    print("ResidualStackFilterBankGenerator gradient vector rel_l2:", rel)
    assert rel < 5e-2
