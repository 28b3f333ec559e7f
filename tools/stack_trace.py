"""Debug aid: per-phase clock64 timeline of CTA 0 of the fused ResidualStack kernel."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops, _lib
from oracle import synth

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, L = 16, {128: 16384, 64: 32768, 32: 65536}[C]
sd = synth.residual_stack_state(1, C)
params = []
for a in range(3):
    for c in range(2):
        params += [sd[f"s.main.{a}.main.{c}.weight"].cuda(), sd[f"s.main.{a}.main.{c}.bias"].cuda()]
blob = ops.resstack_pack_weights(params, C)
x32 = torch.randn(B, C // 8, L, 8, device="cuda") * 0.1
lib = _lib.lib()
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
lib.ms_debug_set_stack_trace.argtypes = [ctypes.c_void_p]
lib.ms_debug_set_stack_trace(ctypes.c_void_p(dbg.data_ptr()))
ops.resstack_fwd(x32, blob, [1, 3, 9], want16=True, want32=False)
torch.cuda.synchronize()
lib.ms_debug_set_stack_trace(ctypes.c_void_p(0))
d = dbg.cpu().tolist()
t0 = d[480]
print("C=%d ; cycles relative to prologue start of tile 0" % C)
print("prologue tile0: %d -> %d ; tile1: %d -> %d" % (0, d[481] - t0, d[482] - t0, d[483] - t0))
for n in range(12):
    m = n * 16
    print("conv %2d MMA A[wait %d->%d first %d done %d]  B[wait %d->%d done %d]" % (
        n, d[m] - t0, d[m + 1] - t0, d[m + 2] - t0, d[m + 3] - t0, d[m + 4] - t0, d[m + 5] - t0, d[m + 7] - t0))
    e = 256 + n * 16
    print("        EPI: A[wait %d->%d done %d] B[wait %d->%d done %d]" % (
        d[e] - t0, d[e + 1] - t0, d[e + 2] - t0, d[e + 4] - t0, d[e + 5] - t0, d[e + 6] - t0))
