import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from music_synthesis_b200 import ops
torch.set_grad_enabled(False)
B, C, L = int(sys.argv[1]), 256, int(sys.argv[2])
torch.manual_seed(0)
x = torch.randn(B, C, L); w = torch.randn(C, C, 3) * 0.05
d = ops.conv_desc(ops.MS_CONV, B, C, C, L, 3, 1, 1)
ref = F.conv1d(x.half().float().double(), w.half().float().double(), padding=1).float()
_, y32 = ops.conv_fwd(d, ops.pack_ncl(x.cuda()), ops.pack_conv_weight(d, w.cuda()), None, want16=False, want32=True)
got = ops.unpack_blk32(y32).cpu()
err = (got - ref).abs()
mt = L // 256
print("overall rel", ((got - ref).norm() / ref.norm()).item())
bad_tiles = []
for b in range(B):
    for m in range(mt):
        e = err[b, :, m * 256:(m + 1) * 256]
        q = [e[:128, :128].max().item(), e[128:, :128].max().item(), e[:128, 128:].max().item(), e[128:, 128:].max().item()]
        if max(q) > 1e-3:
            bad_tiles.append((b * mt + m, ["%.1e" % v for v in q]))
print("tiles", B * mt, "bad", len(bad_tiles))
for t in bad_tiles[:12]:
    print("tile", t[0], "[cols<128 rows<128, cols>=128 rows<128, cols<128 rows>=128, cols>=128 rows>=128]", t[1])
