"""Microbenchmark: tensor-pipe cycles of 64 M=128 K=16 MMAs vs N, the A operand's start row
(tap shift), the number of accumulators cycled through and the run length per accumulator."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "native", "libmsb200_dbg.so"))
out = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.ms_debug_mma_shift.argtypes = [ctypes.c_void_p] + [ctypes.c_int] * 5 + [ctypes.c_void_p]
def run(N, shift, lbo, nacc, runlen):
    lib.ms_debug_mma_shift(ctypes.c_void_p(out.data_ptr()), N, shift, lbo, nacc, runlen, ctypes.c_void_p(0))
    torch.cuda.synchronize()
    return out.cpu().tolist()[0] / 64
for N in (32, 64, 128, 256):
    print("N=%3d shift 0/1/5/9: %s" % (N, " ".join("%.1f" % run(N, s, 256, 2, 0) for s in (0, 1, 5, 9))))
for N in (32, 64, 128):
    for nacc, runlen in ((1, 0), (2, 0), (2, 1), (2, 2), (2, 3), (2, 5)):
        print("N=%3d  %d accumulator(s), runs of %2d MMAs: %.1f cycles/MMA" % (N, nacc, 1 << runlen, run(N, 1, 256, nacc, runlen)))
