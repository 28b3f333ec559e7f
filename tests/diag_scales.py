"""Diagnostic (not collected): 3-scale MelGanDiscriminator features vs oracle, train-mode path."""
import torch

from oracle import restate, synth
from tests.gpu_util import rel_l2


@torch.enable_grad()
def main():
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(112), 1112)
    d = MelGanDiscriminator()
    d.load_state_dict(d_sd)
    d = d.cuda()
    g_sd = restate.randomize_biases(restate.melgan_generator_state(111), 1111)
    with torch.no_grad():
        fake = restate.melgan_generator(synth.mel_features(114, 2, 8), g_sd)
    real = synth.randn(113, 2, 1, 2048) * 0.1
    for name, x in (("fake", fake), ("real", real)):
        fr, jr = restate.melgan_discriminator(x, d_sd)
        f, j = d(x.cuda())
        for s in range(3):
            a, b = f[s][5].detach().cpu(), fr[s][5]
            mism = (torch.sign(a) != torch.sign(b))
            print(name, "scale", s, "shape", tuple(b.shape), "feat5 rel %.2e" % rel_l2(a, b),
                  "feat4 rel %.2e" % rel_l2(f[s][4], fr[s][4]),
                  "judge rel %.2e" % rel_l2(j[s], jr[s]), "sign mismatches", int(mism.sum()),
                  "max |ref| at mismatch %.2e" % (float(b[mism].abs().max()) if mism.any() else 0.0),
                  "frac positive %.3f" % float((b > 0).float().mean()))


if __name__ == "__main__":
    main()


@torch.enable_grad()
def backward_by_scale():
    import torch.nn.functional as F
    from music_synthesis_b200.discriminator.melgan import MelGanDiscriminator
    from music_synthesis_b200.loss.loss import mel_gan_disc_loss
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(112), 1112)
    g_sd = restate.randomize_biases(restate.melgan_generator_state(111), 1111)
    with torch.no_grad():
        fake = restate.melgan_generator(synth.mel_features(114, 2, 8), g_sd)
    real = synth.randn(113, 2, 1, 2048) * 0.1
    for scales in (0, 1, 2):
        for order in ("fake-first", "real-first"):
            d = MelGanDiscriminator()
            d.load_state_dict(d_sd)
            d = d.cuda()
            d.scales = scales
            ref = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
            _, jf = restate.melgan_discriminator(fake, ref, scales=scales)
            _, jr = restate.melgan_discriminator(real, ref, scales=scales)
            lr = restate.mel_gan_disc_loss(jr, jf)
            gref = dict(zip(ref, torch.autograd.grad(lr, list(ref.values()))))
            if order == "fake-first":
                _, j2 = d(fake.cuda())
                _, j1 = d(real.cuda())
            else:
                _, j1 = d(real.cuda())
                _, j2 = d(fake.cuda())
            lo = mel_gan_disc_loss(j1, j2)
            lo.backward()
            print("scales", scales, order, "loss", float(lo.detach()), float(lr.detach()),
                  " ".join("%s %.4f" % (k.replace("disc.", "").replace("main.", "m").replace("weight", "w").replace("bias", "b"),
                                        rel_l2(p.grad, gref[k])) for k, p in d.named_parameters()))


if __name__ == "__main__":
    backward_by_scale()
