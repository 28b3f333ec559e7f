"""Debug: which layer's per-clip result depends on the rest of the batch / mismatches the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from music_synthesis_b200 import ops
from music_synthesis_b200.generator.full import MelGanGenerator
from oracle import restate, synth

torch.set_grad_enabled(False)

def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()

def blk32(x):
    B, C, L = x.shape
    return x.view(B, C // 8, 8, L).permute(0, 1, 3, 2).contiguous()

print("== edge shapes: whole generator vs oracle")
sd = restate.randomize_biases(restate.melgan_generator_state(3), 1003)
g = MelGanGenerator(64, 128).eval(); g.load_state_dict(sd); g = g.cuda()
for B, T in ((1, 4), (1, 5), (1, 8), (2, 512), (1, 33)):
    x = synth.mel_features(5, B, T)
    print(B, T, rel(g(x.cuda()), restate.melgan_generator(x, sd)))

print("== batch independence per layer type (clip 0 of B=1 vs B=40)")
T = 256
for name, kind, cin, cout, lin, k, dil, pad, stride in (
        ("first", ops.MS_CONV, 128, 512, T + 6, 7, 1, 0, 1), ("convT1", ops.MS_CONVT, 512, 256, T, 16, 1, 4, 8),
        ("s256", ops.MS_CONV, 256, 256, 8 * T, 3, 9, 9, 1), ("convT2", ops.MS_CONVT, 256, 128, 8 * T, 16, 1, 4, 8),
        ("convT3", ops.MS_CONVT, 128, 64, 64 * T, 4, 1, 1, 2), ("convT4", ops.MS_CONVT, 64, 32, 128 * T, 4, 1, 1, 2)):
    outs = []
    for B in (1, 40):
        torch.manual_seed(0)
        x = torch.randn(40, cin, lin)[:B]
        w = torch.randn((cout, cin, k) if kind == ops.MS_CONV else (cin, cout, k)) * 0.05
        d = ops.conv_desc(kind, B, cin, cout, lin, k, dil, pad, stride, leaky=True)
        _, y32 = ops.conv_fwd(d, ops.pack_ncl(x.cuda()), ops.pack_conv_weight(d, w.cuda()), None, want16=False, want32=True)
        outs.append(y32[0].clone())
    print(name, "identical:", torch.equal(outs[0], outs[1]), "max diff", (outs[0] - outs[1]).abs().max().item())
for C, L in ((128, 16384), (64, 32768), (32, 65536)):
    s = synth.residual_stack_state(1, C)
    params = []
    for a in range(3):
        for c in range(2):
            params += [s[f"s.main.{a}.main.{c}.weight"].cuda(), s[f"s.main.{a}.main.{c}.bias"].cuda()]
    blob = ops.resstack_pack_weights(params, C)
    outs = []
    for B in (1, 40):
        torch.manual_seed(1)
        x = (torch.randn(40, C, L) * 0.1)[:B]
        _, y32 = ops.resstack_fwd(blk32(x).cuda(), blob, [1, 3, 9])
        outs.append(y32[0].clone())
    diff = (outs[0] - outs[1]).abs()
    print("stack", C, "identical:", torch.equal(outs[0], outs[1]), "max diff", diff.max().item(),
          "rows differing:", torch.nonzero(diff.amax(dim=(0, 2)) > 0).flatten()[:12].tolist())
