"""Host-to-host MelGanGenerator.generate() on BASELINE config 3 (256 clips) with different chunk
plans: the first chunk's H2D copy and the last chunk's D2H copy are exposed, every chunk pays a
fixed pass overhead, and a chunk's D2H copy has to fit under the NEXT chunk's kernels."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from music_synthesis_b200.generator.full import MelGanGenerator

torch.set_grad_enabled(False)
B, T = 256, 256
gen = MelGanGenerator(T, 128).cuda()
x = torch.from_numpy(np.random.RandomState(0).standard_normal((B, 128, T)).astype(np.float32)).pin_memory()
y = torch.empty((B, 1, 256 * T), dtype=torch.float32).pin_memory()
plans = [[8, 120, 120, 8], [8, 124, 108, 16], [8, 120, 104, 24], [16, 120, 104, 16], [8, 128, 96, 24],
         [4, 124, 112, 16], [8, 116, 100, 24, 8], [24, 208, 24], [8, 240, 8], [6, 122, 122, 6],
         [12, 116, 116, 12], [8, 80, 80, 80, 8]]
if len(sys.argv) > 1:
    plans = [[int(v) for v in a.split(",")] for a in sys.argv[1:]]
# round-robin over the plans (the box is power-capped: a sustained run drifts by several per cent,
# so every plan is timed once per round and the rounds are summarised by median and minimum)
rounds = 6
times = {tuple(p): [] for p in plans}
for plan in plans:
    gen.generate(x, out=y, chunks=plan)
for r in range(rounds):
    for plan in (plans if r % 2 == 0 else plans[::-1]):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 4
        e0.record()
        for _ in range(n):
            gen.generate(x, out=y, chunks=plan)
        e1.record()
        torch.cuda.synchronize()
        times[tuple(plan)].append(e0.elapsed_time(e1) / n)
for plan, t in times.items():
    t = sorted(t)
    print(json.dumps({"plan": list(plan), "median_ms": round(t[len(t) // 2], 3), "min_ms": round(t[0], 3)}))
