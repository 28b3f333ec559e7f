import os, sys
sys.path.insert(0, "/root/repo")
import torch
from music_synthesis_b200.generator.full import MelGanGenerator
from oracle import restate, synth
torch.set_grad_enabled(False)
sd = restate.randomize_biases(restate.melgan_generator_state(3), 1003)
g = MelGanGenerator(64, 128).eval(); g.load_state_dict(sd); g = g.cuda()
def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm(dim=(1,2)) / b.norm(dim=(1,2))).tolist()
for B, T in ((1, 512), (2, 512), (2, 256), (4, 256), (3, 300), (2, 384), (2, 260)):
    x = synth.mel_features(5, B, T)
    y = g(x.cuda())
    print(B, T, ["%.2e" % v for v in rel(y, restate.melgan_generator(x, sd))])
