"""-m gpu: the GAN training cycle at BASELINE.json sizes:

  * config 4: batch 32, 32 frames / 8192 samples -- the MelGanGenerator + MelGanDiscriminator
    pair and the official-MelGAN pair (RealMelGanExperiment wiring);
  * config 5: the filter-bank multiscale pair on one GPU's shard of the 8-GPU run
    (8 clips x 65536 samples, band dictionaries, least-squares sub-losses);
  * a 20-cycle trajectory (D loss, G loss) against the oracle's restated trainers + Adam.

What is asserted and why.  Losses and the generated batch are forward quantities: 2e-3 relative /
1e-3 rel-L2 (the north-star bar).  A per-tensor gradient rel-L2 is ill-conditioned at the
reference's init for SOME tensors whatever the arithmetic (LeakyReLU masks flip
where activations lie within the forward rounding error of zero -- reproduced on the CPU by
rounding only the forward operands, test_oracle_golden.py::
test_generator_gradient_tolerance_is_set_by_forward_rounding), so at size the whole gradient is
judged as ONE vector per network (rel-L2 and cosine against the oracle's autograd gradient) and
the post-step weights by the direction of the Adam update: the first Adam step moves every weight
by lr * g / (|g| + eps) ~ +-lr, so for weights whose reference update is ~lr the signs must
agree.  The multi-step check replaces per-tensor bounds: both trajectories run on their OWN
weights for 20 cycles and the losses must stay inside a stated band of each other."""
import pytest
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2

pytestmark = pytest.mark.gpu

LOSS_TOL = 2e-3
WAVE_TOL = 1e-3
GRAD_VEC_TOL = 2e-2       # rel-L2 of the whole gradient vector of a network (measured on B200:
                          # 6e-5 .. 8.4e-3 with the least-squares sub-losses at these sizes)
SIGN_AGREE = 0.97         # Adam update direction, weights with |reference update| > 0.9 lr


@pytest.fixture(autouse=True)
def _grad_on():
    with torch.enable_grad():
        yield


def _flat(named, ref):
    a = torch.cat([p.detach().cpu().reshape(-1).double() for _, p in named])
    b = torch.cat([ref[k].detach().reshape(-1).double() for k, _ in named])
    return a, b


def _grad_metrics(module, ref_grads):
    named = [(k, p.grad) for k, p in module.named_parameters()]
    a, b = _flat(named, ref_grads)
    rel = float((a - b).norm() / b.norm())
    cos = float(torch.dot(a, b) / (a.norm() * b.norm()))
    return rel, cos


def _update_agreement(module, old_sd, new_sd, lr=1e-4):
    named = list(module.named_parameters())
    got = torch.cat([(p.detach().cpu() - old_sd[k]).reshape(-1) for k, p in named])
    ref = torch.cat([(new_sd[k] - old_sd[k]).reshape(-1) for k, _ in named])
    big = ref.abs() > 0.9 * lr
    agree = float((torch.sign(got[big]) == torch.sign(ref[big])).float().mean())
    return agree, int(big.sum()), float((got - ref).abs().max())


def _cycle_at_size(g, d, g_sd, d_sd, samples, feats, gen_fn, disc_fn, lsq, tag):
    """one D step + one G step through the trainer mirrors vs the oracle; returns the metrics"""
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss,
                                                hinge_discriminator_loss, hinge_generator_loss)
    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    sub_d = least_squares_disc_loss if lsq else hinge_discriminator_loss
    sub_g = least_squares_generator_loss if lsq else hinge_generator_loss
    o_sub_d = restate.least_squares_disc_loss if lsq else restate.hinge_discriminator_loss
    o_sub_g = restate.least_squares_generator_loss if lsq else restate.hinge_generator_loss
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, sub_d)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, sub_g)

    def dev(x):
        return {k: v.cuda() for k, v in x.items()} if isinstance(x, dict) else x.cuda()

    rd = d_tr.train(dev(samples), feats.cuda())
    d_loss, d_grads, d_new = restate.discriminator_train_step(
        g_sd, d_sd, samples, feats, {}, sub_loss=o_sub_d, gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rd["d_loss"] - d_loss) < LOSS_TOL * abs(d_loss), (tag, rd["d_loss"], d_loss)
    d_rel, d_cos = _grad_metrics(d, d_grads)
    d_agree, d_n, d_maxdiff = _update_agreement(d, d_sd, d_new)
    d.load_state_dict(d_new)            # the G step of both paths sees the same discriminator

    rg = g_tr.train(dev(samples), feats.cuda())
    g_loss, fake, g_grads, g_new = restate.generator_train_step(
        g_sd, d_new, samples, feats, {}, sub_loss=o_sub_g, gen_fn=gen_fn, disc_fn=disc_fn)
    assert abs(rg["g_loss"] - g_loss) < LOSS_TOL * max(1.0, abs(g_loss)), (tag, rg["g_loss"], g_loss)
    if isinstance(fake, dict):
        wave = max(rel_l2(rg["fake"][s], fake[s]) for s in fake)
    else:
        wave = rel_l2(rg["fake"], fake)
    g_rel, g_cos = _grad_metrics(g, g_grads)
    g_agree, g_n, g_maxdiff = _update_agreement(g, g_sd, g_new)
    print("%s: d_loss %.6f (oracle %.6f)  g_loss %.6f (oracle %.6f)  fake rel_l2 %.2e\n"
          "   D gradient vector rel_l2 %.3e cos %.6f | Adam update signs agree %.4f on %d weights, "
          "max |dw - dw_ref| %.2e\n"
          "   G gradient vector rel_l2 %.3e cos %.6f | Adam update signs agree %.4f on %d weights, "
          "max |dw - dw_ref| %.2e"
          % (tag, rd["d_loss"], d_loss, rg["g_loss"], g_loss, wave, d_rel, d_cos, d_agree, d_n,
             d_maxdiff, g_rel, g_cos, g_agree, g_n, g_maxdiff))
    assert wave < WAVE_TOL, (tag, wave)
    return {"d_rel": d_rel, "d_cos": d_cos, "d_agree": d_agree, "g_rel": g_rel, "g_cos": g_cos,
            "g_agree": g_agree}


def _melgan_pair(T, g_seed, d_seed):
    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    g_sd = restate.randomize_biases(restate.melgan_generator_state(g_seed), 1000 + g_seed)
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(d_seed), 1000 + d_seed)
    g = MelGanGenerator(T, 128)
    g.load_state_dict(g_sd)
    d = MelGanDiscriminator()
    d.load_state_dict(d_sd)
    return g.cuda(), d.cuda(), g_sd, d_sd


@pytest.mark.parametrize("lsq", [True, False], ids=["least_squares", "hinge"])
def test_cfg4_melgan_pair_batch32(lsq):
    """BASELINE config 4: batch 32 x 8192 samples, MelGanGenerator + MelGanDiscriminator"""
    B, T = 32, 32
    g, d, g_sd, d_sd = _melgan_pair(T, 301, 302)
    samples = synth.randn(303, B, 1, 256 * T) * 0.1
    feats = synth.mel_features(304, B, T)
    m = _cycle_at_size(g, d, g_sd, d_sd, samples, feats, restate._melgan_gen, restate._melgan_disc,
                       lsq, "cfg4 melgan pair (%s)" % ("lsq" if lsq else "hinge"))
    assert m["g_rel"] < GRAD_VEC_TOL and m["g_agree"] > SIGN_AGREE
    if lsq:
        # with the hinge loss at init the real / fake terms cancel in the top layers of D
        # (DESIGN 5.6): the D gradient is asserted on the well-conditioned least-squares loss
        assert m["d_rel"] < GRAD_VEC_TOL and m["d_agree"] > SIGN_AGREE
    else:
        # measured 2.9e-2 / cos 0.99984: the hinge D gradient at init is the ill-conditioned one
        assert m["d_rel"] < 6e-2 and m["d_cos"] > 0.998 and m["d_agree"] > SIGN_AGREE


def test_cfg4_realmelgan_pair_batch32():
    """BASELINE config 4 on the RealMelGanExperiment wiring (experiment/realmelgan.py)"""
    from music_synthesis_b200.experiment.realmelgan import Generator, Discriminator
    B, T = 32, 32
    g_sd = restate.realmelgan_generator_state(311)
    d_sd = restate.realmelgan_discriminator_state(312)
    g = Generator(128, 32, n_residual_layers=3)
    g.load_state_dict(g_sd)
    d = Discriminator(3, 16, 4, 4)
    d.load_state_dict(d_sd)
    g, d = g.cuda(), d.cuda()
    samples = synth.randn(313, B, 1, 256 * T) * 0.1
    feats = synth.mel_features(314, B, T)
    m = _cycle_at_size(g, d, g_sd, d_sd, samples, feats,
                       lambda f, sd: restate.realmelgan_generator(f, sd),
                       lambda x, f, sd: restate.realmelgan_discriminator(x, sd),
                       True, "cfg4 realmelgan pair (lsq)")
    assert m["d_rel"] < GRAD_VEC_TOL and m["g_rel"] < GRAD_VEC_TOL
    assert m["d_agree"] > SIGN_AGREE and m["g_agree"] > SIGN_AGREE


def test_cfg5_filterbank_pair_one_gpu_shard():
    """BASELINE config 5, one GPU's shard of the 8-GPU run: 8 clips x 65536 samples"""
    from music_synthesis_b200.generator.multiscale import FilterBankMultiScaleGenerator
    from music_synthesis_b200.discriminator.multiscale import FilterBankMultiScaleDiscriminator
    B, T, N = 8, 256, 65536
    g_sd = restate.fb_generator_state(321, N)
    d_sd = restate.fb_discriminator_state(322, N)
    g = FilterBankMultiScaleGenerator(22050, 128, T, N, recompose=False)
    g.load_state_dict(g_sd)
    d = FilterBankMultiScaleDiscriminator(N, 22050, decompose=False, conditioning_channels=128)
    d.load_state_dict(d_sd)
    g, d = g.cuda(), d.cuda()
    g_banks = [m.filter_bank.filter_bank.detach().cpu() for m in g.channel_generators.values()]
    d_banks = [m.filter_bank.filter_bank.detach().cpu() for m in d.channel_discs.values()]
    sizes = restate.fb_band_sizes(N)
    real = {s: synth.randn(323 + i, B, 1, s) * 0.1 for i, s in enumerate(sizes)}
    feats = synth.mel_features(329, B, T)
    m = _cycle_at_size(
        g, d, g_sd, d_sd, real, feats,
        lambda f, sd: restate.filterbank_multiscale_generator(f, sd, g_banks, N),
        lambda x, f, sd: restate.filterbank_multiscale_discriminator(x, f, sd, d_banks, N),
        True, "cfg5 filter-bank pair, 8 x 65536 (lsq)")
    assert m["d_rel"] < GRAD_VEC_TOL and m["g_rel"] < GRAD_VEC_TOL
    assert m["d_agree"] > SIGN_AGREE and m["g_agree"] > SIGN_AGREE


def test_twenty_cycle_trajectory_follows_the_oracle_trainers():
    """20 D/G cycles, each path on its OWN weights (nothing is re-synchronised): the loss
    trajectories of the sm_100a trainers and of the oracle's restated trainers + Adam must stay
    within a band of each other.  Band: 1e-3 of the loss scale at every step (measured on B200:
    6.4e-6 for D, 1.7e-5 for G) -- Adam's normalised first steps turn every LeakyReLU-mask flip
    into a +-lr weight difference, so the weights decorrelate slowly (the generators differ by
    1.6e-2 rel-L2 on a fresh input after the 20 cycles, bound 5e-2); a wrong gradient, a skipped
    step or a stale weight image shows up as a divergence of order 1 within a few cycles."""
    from music_synthesis_b200.train import GeneratorTrainer, DiscriminatorTrainer, Adam
    from music_synthesis_b200.loss.loss import (mel_gan_disc_loss, mel_gan_gen_loss,
                                                least_squares_disc_loss,
                                                least_squares_generator_loss)
    B, T, CYCLES = 4, 8, 20
    g, d, g_sd, d_sd = _melgan_pair(T, 331, 332)
    g_optim = Adam(g.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_optim = Adam(d.parameters(), lr=1e-4, betas=(0.5, 0.9))
    d_tr = DiscriminatorTrainer(g, g_optim, d, d_optim, mel_gan_disc_loss, least_squares_disc_loss)
    g_tr = GeneratorTrainer(g, g_optim, d, d_optim, mel_gan_gen_loss, least_squares_generator_loss)
    g_ref, d_ref, g_state, d_state = dict(g_sd), dict(d_sd), {}, {}
    worst_d = worst_g = 0.0
    traj = []
    for cyc in range(CYCLES):
        samples = synth.randn(340 + cyc, B, 1, 256 * T) * 0.1
        feats = synth.mel_features(370 + cyc, B, T)
        rd = d_tr.train(samples.cuda(), feats.cuda())
        rg = g_tr.train(samples.cuda(), feats.cuda())
        d_loss, _, d_ref = restate.discriminator_train_step(
            g_ref, d_ref, samples, feats, d_state, sub_loss=restate.least_squares_disc_loss)
        g_loss, _, _, g_ref = restate.generator_train_step(
            g_ref, d_ref, samples, feats, g_state, sub_loss=restate.least_squares_generator_loss)
        ed = abs(rd["d_loss"] - d_loss) / max(abs(d_loss), 1e-3)
        eg = abs(rg["g_loss"] - g_loss) / max(abs(g_loss), 1.0)
        worst_d, worst_g = max(worst_d, ed), max(worst_g, eg)
        traj.append((rd["d_loss"], d_loss, rg["g_loss"], g_loss))
    for cyc, t in enumerate(traj):
        print("cycle %2d  d_loss %.6f (oracle %.6f)  g_loss %.6f (oracle %.6f)" % ((cyc,) + t))
    print("20 cycles: worst relative loss deviation D %.3e G %.3e" % (worst_d, worst_g))
    assert traj[0][1] != traj[-1][1]                    # the weights really moved
    assert worst_d < 1e-3 and worst_g < 1e-3
    # both generators after 20 steps on a fresh input: still the same function to a few percent
    probe = synth.mel_features(399, 2, T)
    with torch.no_grad():
        y = g(probe.cuda()).cpu()
    ref = restate.melgan_generator(probe, g_ref)
    drift = rel_l2(y, ref)
    print("generator after 20 independent cycles vs oracle generator: rel_l2 %.3e" % drift)
    assert drift < 5e-2
