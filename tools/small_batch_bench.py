"""ms per generator forward at a small per-GPU batch (config 3 sharded over 8 GPUs = 32 clips)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200.generator.full import MelGanGenerator
from music_synthesis_b200.experiment.init import weights_init
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
torch.manual_seed(0)
g = MelGanGenerator(256, 128).eval(); g.apply(weights_init); g = g.cuda()
x = torch.randn(B, 128, 256, device="cuda")
with torch.no_grad():
    for _ in range(5): g(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(40): g(x)
    e1.record(); torch.cuda.synchronize()
print("B=%d %s: %.4f ms" % (B, " ".join("%s=%s" % (k, v) for k, v in os.environ.items() if k.startswith("MSB_")), e0.elapsed_time(e1) / 40))
