"""-m gpu: GAN training cycle of the filter-bank multiscale pair (BASELINE config 5 wiring,
experiment/multiscale.py:16-67: FilterBankMultiScaleGenerator(recompose=False) +
FilterBankMultiScaleDiscriminator(decompose=False, conditioning 128), least-squares sub-losses)
through the trainer mirrors vs the oracle's restated trainers, Adam and CPU autograd.  The fixed
Morlet banks are shared as tensors between both paths."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

# Gradient deviations here are LeakyReLU mask flips caused by the fp16 FORWARD (see
# test_oracle_golden.py::test_generator_gradient_tolerance_is_set_by_forward_rounding and
# test_gpu_backward.py::test_strided_conv_block_backward), not backward rounding; they average out
# over a map's elements, so the 128-sample band of this small test case (2 clips) is the noisiest:
# measured worst 7.4e-2 there, < 5e-2 on every other discriminator tensor.
D_TOL, G_TOL = 1e-1, 1e-1


@pytest.fixture(autouse=True)
def _grad_on():
    with torch.enable_grad():
        yield


def _bank_tensors(module_dict):
    return [m.filter_bank.filter_bank.detach().cpu() for m in module_dict.values()]


def test_filterbank_pair_train_cycle_matches_oracle():
    from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
    from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    B, T, N = 2, 8, 2048
    g_sd = restate.fb_generator_state(191, N)
    d_sd = restate.fb_discriminator_state(192, N)
    g = FilterBankMultiScaleGenerator(22050, 128, T, N, recompose=False)
    g.load_state_dict(g_sd)
    d = FilterBankMultiScaleDiscriminator(N, 22050, decompose=False, conditioning_channels=128)
    d.load_state_dict(d_sd)
    g, d = g.cuda(), d.cuda()
    g_banks = _bank_tensors(g.channel_generators)
    d_banks = _bank_tensors(d.channel_discs)
    sizes = restate.fb_band_sizes(N)

    def gen_fn(features, sd):
        return restate.filterbank_multiscale_generator(features, sd, g_banks, N)

    def disc_fn(x, features, sd):
        return restate.filterbank_multiscale_discriminator(x, features, sd, d_banks, N)

    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, least_squares_disc_loss)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, least_squares_generator_loss)
    real = {s: synth.randn(193 + i, B, 1, s) * 0.1 for i, s in enumerate(sizes)}
    feats = synth.mel_features(199, B, T)
    real_gpu = {s: v.cuda() for s, v in real.items()}

    rd = d_tr.train(real_gpu, feats.cuda())
    d_loss, d_grads, d_new = restate.discriminator_train_step(
        g_sd, d_sd, real, feats, {}, sub_loss=restate.least_squares_disc_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rd["d_loss"] - d_loss) < 2e-3 * abs(d_loss), (rd["d_loss"], d_loss)
    errs = {k: rel_l2(p.grad, d_grads[k]) for k, p in d.named_parameters()}
    bad = {k: round(e, 4) for k, e in errs.items() if not e < D_TOL}
    worst_d = max(errs.values())
    assert not bad, ("D", bad)
    d.load_state_dict(d_new)

    rg = g_tr.train(real_gpu, feats.cuda())
    g_loss, fake, g_grads, _ = restate.generator_train_step(
        g_sd, d_new, real, feats, {}, sub_loss=restate.least_squares_generator_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rg["g_loss"] - g_loss) < 2e-3 * max(1.0, abs(g_loss)), (rg["g_loss"], g_loss)
    for s in sizes:
        assert rel_l2(rg["fake"][s], fake[s]) < 1e-3
    errs = {k: rel_l2(p.grad, g_grads[k]) for k, p in g.named_parameters()}
    bad = {k: round(e, 4) for k, e in errs.items() if not e < G_TOL}
    worst_g = max(errs.values())
    assert not bad, ("G", bad)
    print("filter-bank pair: worst grad rel_l2 D %.4f G %.4f" % (worst_d, worst_g))


@pytest.mark.parametrize("in_graph", [False, True], ids=["band_dicts", "fft_in_graph"])
def test_multiscale_pair_train_cycle_matches_oracle(in_graph):
    """the non-filterbank multiscale pair as wired by experiment/multiscale.py:120-160
    (MultiScaleNoDeRecompose: band dictionaries, conditioned k41 discriminator, least squares);
    in_graph: the same pair with recompose=True / decompose=True -- the FFT band merge and split
    sit inside the autograd graph (generator/multiscale.py:248-251, discriminator/multiscale.py:
    395-397) and the trainers exchange waveforms"""
    from music_synthesis_b200.generator.multiscale import MultiScaleGenerator
    from music_synthesis_b200.discriminator.multiscale import MultiScaleMultiResDiscriminator
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    B, T, N = 2, 8, 2048
    g_sd = restate.multiscale_generator_state(201, N)
    d_sd = restate.multiscale_discriminator_state(202, N)
    g = MultiScaleGenerator(128, T, N, transposed_conv=True, recompose=in_graph)
    g.load_state_dict(g_sd)
    d = MultiScaleMultiResDiscriminator(N, flatten_multiscale_features=False, decompose=in_graph,
                                        channel_judgements=True, conditioning_channels=128)
    d.load_state_dict(d_sd)
    g, d = g.cuda(), d.cuda()
    sizes = restate.fb_band_sizes(N)

    def gen_fn(features, sd):
        return restate.multiscale_generator(features, sd, N, recompose=in_graph)

    def disc_fn(x, features, sd):
        return restate.multiscale_multires_discriminator(x, features, sd, N, decompose=in_graph)

    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, least_squares_disc_loss)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, least_squares_generator_loss)
    real = {s: synth.randn(203 + i, B, 1, s) * 0.1 for i, s in enumerate(sizes)}
    if in_graph:
        real = synth.randn(203, B, 1, N) * 0.1
    feats = synth.mel_features(209, B, T)
    real_gpu = real.cuda() if in_graph else {s: v.cuda() for s, v in real.items()}
    rd = d_tr.train(real_gpu, feats.cuda())
    d_loss, d_grads, d_new = restate.discriminator_train_step(
        g_sd, d_sd, real, feats, {}, sub_loss=restate.least_squares_disc_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rd["d_loss"] - d_loss) < 2e-3 * abs(d_loss), (rd["d_loss"], d_loss)
    errs = {k: rel_l2(p.grad, d_grads[k]) for k, p in d.named_parameters()}
    bad = {k: round(e, 4) for k, e in errs.items() if not e < D_TOL}
    worst_d = max(errs.values())
    assert not bad, ("D", bad)
    d.load_state_dict(d_new)
    rg = g_tr.train(real_gpu, feats.cuda())
    g_loss, fake, g_grads, _ = restate.generator_train_step(
        g_sd, d_new, real, feats, {}, sub_loss=restate.least_squares_generator_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rg["g_loss"] - g_loss) < 2e-3 * max(1.0, abs(g_loss)), (rg["g_loss"], g_loss)
    if in_graph:
        assert rel_l2(rg["fake"], fake) < 1e-3
    else:
        for s in sizes:
            assert rel_l2(rg["fake"][s], fake[s]) < 1e-3
    errs = {k: rel_l2(p.grad, g_grads[k]) for k, p in g.named_parameters()}
    if in_graph:
        # the band merge keeps bins >= S/4 of every band but the lowest: a constant offset of
        # those bands never reaches the output, so d loss / d to_samples.bias is analytically ZERO
        # (both sides hold rounding noise ~1e-9): compare those on the scale of the lowest band's
        lowest = "channel_%d.to_samples.bias" % min(sizes)
        scale = g_grads[lowest].abs().max().item()
        for k, p in g.named_parameters():
            if k.endswith("to_samples.bias") and k != lowest:
                assert (p.grad.cpu() - g_grads[k]).abs().max().item() < 1e-3 * scale, k
                assert g_grads[k].abs().max().item() < 1e-3 * scale, k
                errs.pop(k)
    bad = {k: round(e, 4) for k, e in errs.items() if not e < G_TOL}
    worst_g = max(errs.values())
    assert not bad, ("G", bad)
    print("multiscale pair: worst grad rel_l2 D %.4f G %.4f" % (worst_d, worst_g))


def test_realmelgan_pair_train_cycle_matches_oracle():
    """the official-MelGAN pair (experiment/realmelgan.py: weight-normed, reflection-padded
    generator; three independent NLayerDiscriminators) through the trainers: gradients reach
    weight_g / weight_v through the weight-norm fold"""
    from music_synthesis_b200.experiment.realmelgan import Generator, Discriminator
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    B, T = 2, 8
    g_sd = restate.realmelgan_generator_state(211)
    d_sd = restate.realmelgan_discriminator_state(212)
    g = Generator(128, 32, n_residual_layers=3)
    g.load_state_dict(g_sd)
    d = Discriminator(3, 16, 4, 4)
    d.load_state_dict(d_sd)
    g, d = g.cuda(), d.cuda()

    def gen_fn(features, sd):
        return restate.realmelgan_generator(features, sd)

    def disc_fn(x, features, sd):
        return restate.realmelgan_discriminator(x, sd)

    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, least_squares_disc_loss)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, least_squares_generator_loss)
    real = synth.randn(213, B, 1, 256 * T) * 0.1
    feats = synth.mel_features(214, B, T)
    rd = d_tr.train(real.cuda(), feats.cuda())
    d_loss, d_grads, d_new = restate.discriminator_train_step(
        g_sd, d_sd, real, feats, {}, sub_loss=restate.least_squares_disc_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rd["d_loss"] - d_loss) < 2e-3 * abs(d_loss), (rd["d_loss"], d_loss)
    errs = {k: rel_l2(p.grad, d_grads[k]) for k, p in d.named_parameters()}
    bad = {k: round(e, 4) for k, e in errs.items() if not e < D_TOL}
    worst_d = max(errs.values())
    assert not bad, ("D", bad)
    d.load_state_dict(d_new)
    rg = g_tr.train(real.cuda(), feats.cuda())
    g_loss, fake, g_grads, _ = restate.generator_train_step(
        g_sd, d_new, real, feats, {}, sub_loss=restate.least_squares_generator_loss,
        gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rg["g_loss"] - g_loss) < 2e-3 * max(1.0, abs(g_loss)), (rg["g_loss"], g_loss)
    assert rel_l2(rg["fake"], fake) < 1e-3
    errs = {k: rel_l2(p.grad, g_grads[k]) for k, p in g.named_parameters()}
    bad = {k: round(e, 4) for k, e in errs.items() if not e < G_TOL}
    worst_g = max(errs.values())
    assert not bad, ("G", bad)
    print("realmelgan pair: worst grad rel_l2 D %.4f G %.4f" % (worst_d, worst_g))
