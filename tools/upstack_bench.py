"""Times the fused upsampling stages alone (csrc/upstack.cu) at the bench's per-pass size
(64 clips): C=64 (lin 16384) and C=32 + tail (lin 32768).  With a library built with
MSB_NVCC_EXTRA=-DMSB_UP_ABLATE, MSB_UP_ABLATE=<bits> switches parts of the kernel off (results are
then garbage; only the time is meaningful): 1 no MMAs, 2 empty epilogue, 4 no operand stores,
8 no tensor-memory traffic in the epilogue."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
torch.manual_seed(0)
out = {"ablate": os.environ.get("MSB_UP_ABLATE", "0"), "clips": B}
for C, lin in ((64, 16384), (32, 32768)):
    params = [torch.randn(2 * C, C, 4, device="cuda") * 0.02, torch.zeros(C, device="cuda")]
    for _ in range(6):
        params += [torch.randn(C, C, 3, device="cuda") * 0.02, torch.zeros(C, device="cuda")]
    blob = ops.upstack_pack_weights(params, C)
    x16 = ops.pack_ncl(torch.randn(B, 2 * C, lin, device="cuda") * 0.1)
    tail = (torch.randn(1, 32, 7, device="cuda") * 0.02, torch.zeros(1, device="cuda")) if C == 32 else None
    run = lambda: ops.upstack_fwd(x16, blob, [1, 3, 9], tail=tail)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    flop = (6 * 2 * 3 * C * C + 2 * 2 * 2 * C * C) * B * 2 * lin
    out["c%d" % C] = {"us": round(us, 1), "tflops": round(flop / us / 1e6, 1)}
print(json.dumps(out))
