"""Microbenchmark: tensor-pipe cycles of 64 M=128 K=16 MMAs vs N and the A operand's start row."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "native", "libmsb200_dbg.so"))
out = torch.zeros(8, dtype=torch.int64, device="cuda")
lib.ms_debug_mma_shift.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for lbo in (256, 512, 1024):
    for N in (32, 64, 128, 256):
        row = []
        for shift in (0, 1, 2, 3, 4, 5, 8, 9):
            lib.ms_debug_mma_shift(ctypes.c_void_p(out.data_ptr()), N, shift, lbo, ctypes.c_void_p(0))
            torch.cuda.synchronize()
            row.append("%d:%.1f" % (shift, out.cpu().tolist()[0] / 64))
        print("LBO rows %4d N=%3d  cycles/MMA by shift  %s" % (lbo, N, "  ".join(row)))
