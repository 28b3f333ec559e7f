"""Debug aid: clock64 timeline of CTA 0's third tile in the fused upsampling stage kernel.
Needs a library built with MSB_NVCC_EXTRA=-DMSB_UP_ABLATE."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops, _lib

C = int(sys.argv[1]) if len(sys.argv) > 1 else 32
B, lin = 64, {64: 16384, 32: 32768}[C]
params = [torch.randn(2 * C, C, 4, device="cuda") * 0.02, torch.zeros(C, device="cuda")]
for _ in range(6):
    params += [torch.randn(C, C, 3, device="cuda") * 0.02, torch.zeros(C, device="cuda")]
blob = ops.upstack_pack_weights(params, C)
x16 = ops.pack_ncl(torch.randn(B, 2 * C, lin, device="cuda") * 0.1)
tail = (torch.randn(1, 32, 7, device="cuda") * 0.02, torch.zeros(1, device="cuda")) if C == 32 else None
lib = _lib.lib()
dbg = torch.zeros(256, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.upstack_fwd(x16, blob, [1, 3, 9], tail=tail)
lib.ms_debug_set_upstack_trace.argtypes = [ctypes.c_void_p]
lib.ms_debug_set_upstack_trace(ctypes.c_void_p(dbg.data_ptr()))
ops.upstack_fwd(x16, blob, [1, 3, 9], tail=tail)
torch.cuda.synchronize()
lib.ms_debug_set_upstack_trace(ctypes.c_void_p(0))
d = dbg.cpu().tolist()
t0 = d[243]
r = lambda i: d[i] - t0
print("C=%d ablate=%s; cycles relative to the MMA warp reaching the tile" % (C, os.environ.get("MSB_UP_ABLATE", "0")))
print("producer: in_free wait %d -> %d ; in_full wait A %d -> %d  B %d -> %d" % (r(240), r(241), r(243), r(244), r(245), r(246)))
for s in range(7):
    m = s * 16
    print("stage %d MMA A[wait %d->%d first %d done %d]  B[wait %d->%d done %d]" % (
        s, r(m), r(m + 1), r(m + 2), r(m + 3) if s else r(m + 2), r(m + 4), r(m + 5), r(m + 7) if s else r(m + 6)))
    e = 128 + s * 16
    print("        EPI: p0[wait %d->%d done %d] p1[wait %d->%d done %d]" % (
        r(e), r(e + 1), r(e + 2), r(e + 4), r(e + 5), r(e + 6)))
