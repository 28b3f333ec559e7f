"""Debug aid: clock64 timeline of cluster 0's leader CTA in the CTA-pair conv kernel
(local tiles 4..7).  Needs a library built with MSB_NVCC_EXTRA=-DMSB_CONV_ABLATE."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops, _lib

which = sys.argv[1] if len(sys.argv) > 1 else "convT2"
B, T = 64, 256
cfgs = {"convT2": (ops.MS_CONVT, 256, 128, 8 * T, 16, 1, 4, 8), "s256c1": (ops.MS_CONV, 256, 256, 8 * T, 3, 3, 3, 1),
        "convT1": (ops.MS_CONVT, 512, 256, T, 16, 1, 4, 8)}
kind, cin, cout, lin, k, dil, pad, stride = cfgs[which]
d = ops.conv_desc(kind, B, cin, cout, lin, k, dil, pad, stride, leaky=True)
x16 = torch.zeros((B, cin // 8, lin, 8), dtype=torch.int16, device="cuda")
w = torch.randn((cout, cin, k) if kind == ops.MS_CONV else (cin, cout, k), device="cuda") * 0.02
wp = ops.pack_conv_weight(d, w)
bias = torch.zeros(cout, device="cuda")
lib = _lib.lib()
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
run = lambda: ops.conv_fwd(d, x16, wp, bias, None, want16=(kind == ops.MS_CONV), want32=(kind == ops.MS_CONVT))
for _ in range(2):
    run()
lib.ms_debug_set_conv_trace.argtypes = [ctypes.c_void_p]
lib.ms_debug_set_conv_trace(ctypes.c_void_p(dbg.data_ptr()))
run()
torch.cuda.synchronize()
lib.ms_debug_set_conv_trace(ctypes.c_void_p(0))
v = dbg.cpu().tolist()
t0 = v[64]
r = lambda i: v[i] - t0
print("%s ablate=%s wres=%s; cycles relative to the MMA warp reaching local tile 4" % (
    which, os.environ.get("MSB_CONV_ABLATE", "0"), os.environ.get("MSB_CONV_WRES", "1")))
for ti in range(4):
    o = ti * 16
    print("tile %d PROD " % (ti + 4) + " ".join("kb%d[empty %d->%d issued %d]" % (kb, r(o + kb * 3), r(o + kb * 3 + 1), r(o + kb * 3 + 2)) for kb in range(4)))
    m = 64 + o
    print("       MMA  tempty %d->%d " % (r(m), r(m + 1)) + " ".join("kb%d[full %d pfull %d issued %d]" % (kb, r(m + 2 + kb * 3), r(m + 3 + kb * 3), r(m + 4 + kb * 3)) for kb in range(4)))
    e = 128 + o
    print("       EPI  start %d tables %d tfull %d done %d" % (r(e), r(e + 1), r(e + 2), r(e + 3)))
