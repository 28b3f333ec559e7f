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
MB = 256 // C
t0 = d[480]
print("C=%d MB=%d ; all times in cycles relative to prologue start of tile 0" % (C, MB))
print("prologue tile0: %d -> %d ; tile1: %d -> %d" % (0, d[481] - t0, d[482] - t0, d[483] - t0))
for n in range(12):
    mma = []
    for mb in range(MB if MB <= 4 else 4):
        b = n * 16 + mb * 4
        if mb * 4 + 2 < 15:
            mma.append("mb%d[wait_act %d->%d, w_ready %d]" % (mb, d[b] - t0, d[b + 1] - t0, d[b + 2] - t0))
    print("conv %2d MMA: %s issue_done %d" % (n, " ".join(mma), d[n * 16 + 15] - t0))
    epi = []
    for mb in range(MB if MB <= 4 else 4):
        b = 256 + n * 16 + mb * 4
        if mb * 4 + 2 < 16:
            epi.append("mb%d[wait %d->%d done %d]" % (mb, d[b] - t0, d[b + 1] - t0, d[b + 2] - t0))
    print("        EPI: %s" % " ".join(epi))
