"""FFT band split / merge with two passes per launch against one pass per launch
(MSB_FFT_FUSE) and with / without the loaders' twiddle table (MSB_FFT_TABLE; both read by the library per call): CUDA events over warm back-to-back calls, one JSON
line per setting.  Algorithmic bytes = the clip in + the five bands out (31/16 of the clip)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
from music_synthesis_b200 import _lib

torch.set_grad_enabled(False)
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    PEAK = float(PEAK.get("hbm_gbps", PEAK.get("hbm_gbs", 6536.7))) * 1e9
except Exception:
    PEAK = 6536.7e9


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / n


for B in ([int(v) for v in sys.argv[1:]] or [64, 512]):
    rs = np.random.RandomState(4)
    x = torch.from_numpy((rs.standard_normal((8, 1, 65536)) * 0.1).astype(np.float32)).repeat(B // 8, 1, 1).cuda()
    nbytes = x.numel() * 4 * (1 + 31 / 16)
    for rep in range(1):
        for fuse, table in (("1", "1"), ("1", "0"), ("0", "0")):
            os.environ["MSB_FFT_FUSE"] = fuse
            os.environ["MSB_FFT_TABLE"] = table
            l0 = _lib.launch_count()
            bands = fft_frequency_decompose(x, 4096)
            l1 = _lib.launch_count()
            fft_frequency_recompose(bands, 65536)
            l2 = _lib.launch_count()
            ts = timeit(lambda: fft_frequency_decompose(x, 4096))
            tm = timeit(lambda: fft_frequency_recompose(bands, 65536))
            print(json.dumps({"clips": B, "fuse": fuse, "table": table, "split_us": round(ts * 1e6, 1), "split_frac": round(nbytes / ts / PEAK, 4),
                              "split_launches": l1 - l0, "merge_us": round(tm * 1e6, 1),
                              "merge_frac": round(nbytes / tm / PEAK, 4), "merge_launches": l2 - l1}))
            del bands
