"""CPU baselines of the side rows (not a pytest module): the oracle port of Audio2Mel, the FFT
band split / merge and `batch_stream` timed on the host cores, next to the GPU numbers of
tools/side_bench.py.  Lives under tests/ because it executes oracle/.

    python tests/side_cpu_baseline.py            # prints one JSON line per row
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bases, restate, synth  # noqa: E402


def best_of(fn, n=5):
    fn()
    times = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return min(times)


def main():
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    rows = []
    a = synth.uniform_audio(3, 64, 16384)
    basis = torch.from_numpy(np.asarray(bases.librosa_mel(22050, 1024, 128), dtype=np.float32))
    window = torch.hann_window(1024)
    t = best_of(lambda: restate.audio2mel(a, basis, window))
    rows.append({"row": "a1 Audio2Mel (oracle port)", "workload": "B=64 x 16384 samples", "us": t * 1e6,
                 "samples_per_s": 64 * 16384 / t})
    x = synth.randn(4, 64, 1, 65536) * 0.1
    t = best_of(lambda: restate.fft_frequency_decompose(x, 4096), n=3)
    rows.append({"row": "a8 fft_frequency_decompose (oracle port)", "workload": "B=64 x 65536, 5 bands",
                 "us": t * 1e6})
    bands = restate.fft_frequency_decompose(x, 4096)
    t = best_of(lambda: restate.fft_frequency_recompose(bands, 65536), n=3)
    rows.append({"row": "a8 fft_frequency_recompose (oracle port)", "workload": "B=64 x 65536, 5 bands",
                 "us": t * 1e6})
    rs = np.random.RandomState(0)
    chunks = [(rs.random_sample(661500) * 2 - 1).astype(np.float32) for _ in range(8)]
    specs = [rs.standard_normal(((len(c) - 640) // 256 + 1, 128)).astype(np.float32) for c in chunks]
    spec = {"audio": (8192, 1), "spectrogram": (32, 128)}
    t = best_of(lambda: restate.batch_stream(chunks, specs, 32, spec, "spectrogram", 0, 1))
    rows.append({"row": "f3 batch_stream (oracle port, features already cached in memory)",
                 "workload": "B=32 x (8192 samples + 128x32 log-mel)", "us": t * 1e6, "clips_per_s": 32 / t})
    for r in rows:
        r["cores"] = cores
        r["torch_threads"] = torch.get_num_threads()
        print(json.dumps(r))


if __name__ == "__main__":
    main()
