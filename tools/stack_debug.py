import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops
from oracle import restate, synth
torch.set_grad_enabled(False)
for C, B, L in ((128, 6, 16384), (64, 6, 16384), (32, 6, 32768)):
    sd = synth.residual_stack_state(50 + C, C)
    x = synth.randn(51, B, C, L) * 0.5
    ref = restate.residual_stack(x, sd, "s")
    params = []
    for a in range(3):
        for c in range(2):
            params += [sd[f"s.main.{a}.main.{c}.weight"].cuda(), sd[f"s.main.{a}.main.{c}.bias"].cuda()]
    blob = ops.resstack_pack_weights(params, C)
    x32 = x.view(B, C // 8, 8, L).permute(0, 1, 3, 2).contiguous().cuda()
    _, y32 = ops.resstack_fwd(x32, blob, [1, 3, 9])
    got = ops.unpack_blk32(y32).cpu()
    per_clip = ((got - ref).double().norm(dim=(1, 2)) / ref.double().norm(dim=(1, 2))).tolist()
    print("stack C=%d tiles/clip=%d" % (C, -(-L // (32768 // C - 32))), ["%.1e" % v for v in per_clip])
