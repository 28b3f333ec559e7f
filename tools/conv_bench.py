"""Back-to-back timing of single generator layers through ms_conv_fwd (CUDA events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops

B = int(os.environ.get("B", "256"))
T = 256
layers = [  # name, kind, cin, cout, lin, k, dil, pad, stride, res
    ("first", ops.MS_CONV, 128, 512, T + 6, 7, 1, 0, 1, False),
    ("convT1", ops.MS_CONVT, 512, 256, T, 16, 1, 4, 8, False),
    ("s256c1", ops.MS_CONV, 256, 256, 8 * T, 3, 3, 3, 1, False),
    ("s256c2", ops.MS_CONV, 256, 256, 8 * T, 3, 1, 1, 1, True),
    ("convT2", ops.MS_CONVT, 256, 128, 8 * T, 16, 1, 4, 8, False),
    ("convT3", ops.MS_CONVT, 128, 64, 64 * T, 4, 1, 1, 2, False),
    ("convT4", ops.MS_CONVT, 64, 32, 128 * T, 4, 1, 1, 2, False),
]
print("MSB_CONV_PAIR=%s B=%d" % (os.environ.get("MSB_CONV_PAIR", "1"), B))
for name, kind, cin, cout, lin, k, dil, pad, stride, res in layers:
    d = ops.conv_desc(kind, B, cin, cout, lin, k, dil, pad, stride, leaky=True)
    lout = ops.conv_out_len(d)
    x16 = torch.zeros((B, cin // 8, lin, 8), dtype=torch.int16, device="cuda")
    w = torch.randn((cout, cin, k) if kind == ops.MS_CONV else (cin, cout, k), device="cuda") * 0.02
    wp = ops.pack_conv_weight(d, w)
    bias = torch.zeros(cout, device="cuda")
    r32 = torch.zeros((B, cout // 8, lout, 8), device="cuda") if res else None
    want16 = kind == ops.MS_CONV
    for _ in range(3):
        ops.conv_fwd(d, x16, wp, bias, r32, want16=want16, want32=(res or kind == ops.MS_CONVT))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        ops.conv_fwd(d, x16, wp, bias, r32, want16=want16, want32=(res or kind == ops.MS_CONVT))
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / n
    flop = 2.0 * B * lout * cout * cin * (k if kind == ops.MS_CONV else 2) / (1 if kind == ops.MS_CONV else stride) * (1 if kind == ops.MS_CONV else 1)
    if kind == ops.MS_CONVT:
        flop = 2.0 * B * lin * cin * cout * k
    print("%-8s %9.1f us  %7.1f TFLOP/s" % (name, us, flop / us / 1e6))
