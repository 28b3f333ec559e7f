import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "native", "libmsb200_dbg.so"))
out = torch.zeros(64, dtype=torch.int64, device="cuda")
lib.ms_debug_microbench.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
for _ in range(2):
    lib.ms_debug_microbench(ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(0))
torch.cuda.synchronize()
o = out.cpu().tolist()
names = ["clock overhead", "mbar_wait (completed)", "elect+syncwarp", "tc_fence_after", "commit issue (idle)", "commit->wait (idle)"]
for i, n in enumerate(names):
    print("%-28s %6d" % (n, o[i]))
s = 6
for N in (32, 64, 128, 256):
    for v in ("same acc", "2 accs"):
        print("16 MMAs N=%-3d %-8s issue %6d  issue+complete %6d  (ideal exec %d)" % (N, v, o[s], o[s + 1], 16 * N // 2))
        s += 2
print("64 MMAs N=128 issue %d complete %d (ideal %d)" % (o[s], o[s + 1], 64 * 64)); s += 2
print("64 MMAs N=32 (4 accs) issue %d complete %d (ideal %d)" % (o[s], o[s + 1], 64 * 16))

names2 = ["tmem ld16+wait", "4x tmem ld16 + wait", "tmem st16 + wait", "2x st.shared.v4", "fence.proxy.async",
          "tc_fence_before+mbar_arrive", "LDG.128 (L2/cold)", "LDG.128 (L1 hit)", "named bar (1 warp)"]
for i, n in enumerate(names2):
    print("%-30s %6d" % (n, o[26 + i]))

# TMEM read bandwidth (one CTA): W warps x iters loads of 32 lanes x NCOL fp32 columns
lib.ms_debug_tmem_bw.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
for ncol in (16, 32, 64):
    for warps in (1, 4, 8, 16):
        iters = 1024
        for _ in range(2):
            lib.ms_debug_tmem_bw(ctypes.c_void_p(out.data_ptr()), ncol, warps, iters, ctypes.c_void_p(0))
        torch.cuda.synchronize()
        cyc = out.cpu().tolist()[0]
        print("tmem ld x%-2d %2d warps: %7d cycles for %8d B = %.1f B/clk per SM" % (
            ncol, warps, cyc, warps * iters * 128 * ncol, warps * iters * 128 * ncol / cyc))
