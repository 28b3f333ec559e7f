"""Diagnostic (not collected by pytest): per-tensor gradient error of one D step and one G step
vs the oracle.  python -m tests.diag_train"""
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2
from tests.test_gpu_train import _pair, _trainers


def main(B=2, T=8):
    g_sd = restate.randomize_biases(restate.melgan_generator_state(111), 1111)
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(112), 1112)
    g, d = _pair(g_sd, d_sd, T)
    d_tr, g_tr = _trainers(g, d)
    samples = synth.randn(113, B, 1, 256 * T) * 0.1
    features = synth.mel_features(114, B, T)
    rd = d_tr.train(samples.cuda(), features.cuda())
    d_loss, d_grads, d_new = restate.discriminator_train_step(g_sd, d_sd, samples, features, {})
    print("d_loss", rd["d_loss"], d_loss)
    for k, p in d.named_parameters():
        print("  D %-22s |g| %.3e rel %.4f" % (k, float(d_grads[k].norm()), rel_l2(p.grad, d_grads[k])))
    d.load_state_dict(d_new)
    rg = g_tr.train(samples.cuda(), features.cuda())
    g_loss, fake, g_grads, g_new = restate.generator_train_step(g_sd, d_new, samples, features, {})
    print("g_loss", rg["g_loss"], g_loss)
    for k, p in g.named_parameters():
        print("  G %-28s |g| %.3e rel %.4f" % (k, float(g_grads[k].norm()), rel_l2(p.grad, g_grads[k])))


if __name__ == "__main__":
    main()
