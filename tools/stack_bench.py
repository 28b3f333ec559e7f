"""Back-to-back timing of the fused ResidualStack kernel (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops
from oracle import synth

B = int(os.environ.get("B", "256"))
print("MSB_STACK_PAIR=%s" % os.environ.get("MSB_STACK_PAIR", "1 (default)"))
for C, L in ((128, 16384), (64, 32768), (32, 65536)):
    sd = synth.residual_stack_state(1, C)
    params = []
    for a in range(3):
        for c in range(2):
            params += [sd[f"s.main.{a}.main.{c}.weight"].cuda(), sd[f"s.main.{a}.main.{c}.bias"].cuda()]
    blob = ops.resstack_pack_weights(params, C)
    x32 = torch.randn(B, C // 8, L, 8, device="cuda") * 0.1
    for _ in range(2):
        ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 5
    e0.record()
    for _ in range(n):
        ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flop = 6 * 2.0 * 3 * C * C * B * L
    print("C=%3d  %9.1f us  %7.1f TFLOP/s" % (C, us, flop / us / 1e6))
