"""Diagnostic (not collected): FullDiscriminator forward/backward vs oracle on one scale."""
import torch
import torch.nn.functional as F

from oracle import restate, synth
from tests.gpu_util import rel_l2


@torch.enable_grad()
def main():
    from music_synthesis_b200.discriminator.full import FullDiscriminator
    from music_synthesis_b200.loss.loss import hinge_discriminator_loss
    d_sd = restate.randomize_biases(restate.melgan_discriminator_state(112), 1112)
    sd = {k[len("disc."):]: v for k, v in d_sd.items()}
    d = FullDiscriminator()
    d.load_state_dict(sd)
    d = d.cuda()
    real = synth.randn(113, 2, 1, 2048) * 0.1
    fake = synth.randn(115, 2, 1, 2048) * 0.01
    ref = {k: v.clone().requires_grad_(True) for k, v in d_sd.items()}
    for name, lossfn_ref, lossfn in (
            ("mean(judge(real))", lambda jr, jf: jr.mean(), None),
            ("hinge(real,fake)", lambda jr, jf: (F.relu(1 - jr) + F.relu(1 + jf)).mean(), hinge_discriminator_loss)):
        fr, jr = restate.full_discriminator(real, ref)
        ff, jf = restate.full_discriminator(fake, ref)
        lr = lossfn_ref(jr, jf)
        gref = dict(zip(ref, torch.autograd.grad(lr, list(ref.values()), allow_unused=True)))
        d.zero_grad()
        f1, j1 = d(real.cuda())
        f2, j2 = d(fake.cuda())
        print(name, "feature5 rel", rel_l2(f1[5], fr[5]), rel_l2(f2[5], ff[5]),
              "sign mismatches", int((torch.sign(f1[5].cpu()) != torch.sign(fr[5])).sum()),
              int((torch.sign(f2[5].cpu()) != torch.sign(ff[5])).sum()))
        if lossfn is None:
            from music_synthesis_b200.loss.loss import hinge_generator_loss
            lo = -hinge_generator_loss(j1)
        else:
            lo = lossfn(j1, j2)
        print("  loss", float(lo), float(lr))
        lo.backward()
        for k, p in d.named_parameters():
            g = gref["disc." + k]
            print("   %-16s |g| %.3e rel %.4f" % (k, float(g.norm()), rel_l2(p.grad, g)))


if __name__ == "__main__":
    main()
