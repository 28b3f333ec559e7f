"""Timings of the HBM-bound side rows (a1 Audio2Mel, a8 FFT band split / merge): CUDA events over
warm back-to-back launches, one JSON line per row.  Run with the knobs the library reads per call (MSB_A2M_RADIX4, MSB_FFT_LEGACY, MSB_FFT_STAGED) toggled in
this process for the A/B numbers of profiles/."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from music_synthesis_b200.feature.feature import Audio2Mel
from music_synthesis_b200.audio.transform import fft_frequency_decompose, fft_frequency_recompose
import numpy as np

torch.set_grad_enabled(False)


def uniform_audio(seed, batch, samples):
    rs = np.random.RandomState(seed)
    return torch.from_numpy((rs.random_sample((batch, 1, samples)) * 2 - 1).astype(np.float32))


def randn(seed, *shape):
    return torch.from_numpy(np.random.RandomState(seed).standard_normal(shape).astype(np.float32))

try:
    PEAK_HBM = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))
    PEAK_HBM = float(PEAK_HBM.get("hbm_gbps", PEAK_HBM.get("hbm_gbs", 6536.7))) * 1e9
except Exception:
    PEAK_HBM = 6536.7e9
KNOBS = ("MSB_A2M_RADIX4", "MSB_FFT_LEGACY", "MSB_FFT_STAGED", "MSB_FFT_PACKED", "MSB_FFT_MERGE_GATHER",
         "MSB_FFT_FUSE", "MSB_FFT_TABLE")


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




def run(knobs):
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(knobs)
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    for B in (64, 4096):
        a = uniform_audio(3, min(B, 64), 16384).repeat(B // min(B, 64), 1, 1).cuda()
        t = timeit(lambda: a2m(a))
        nbytes = 4 * B * 16384 + 4 * B * 128 * 62
        print(json.dumps({"row": "a1 Audio2Mel", "workload": "B=%d x 16384 samples" % B, "us": round(t * 1e6, 2),
                          "samples_per_s": B * 16384 / t, "bound": "hbm", "achieved_GBs": nbytes / t / 1e9,
                          "frac": nbytes / t / PEAK_HBM, "knobs": knobs}))
    for B in (64, 512):
        x = (randn(4, 8, 1, 65536) * 0.1).repeat(B // 8, 1, 1).cuda()
        t = timeit(lambda: fft_frequency_decompose(x, 4096))
        # algorithmic bytes: the clip in, the five bands (31/16 of the clip) out
        nbytes = x.numel() * 4 * (1 + 31 / 16)
        print(json.dumps({"row": "a8 fft_frequency_decompose", "workload": "B=%d x 65536, 5 bands" % B,
                          "us": round(t * 1e6, 2), "bound": "hbm", "achieved_GBs": nbytes / t / 1e9,
                          "frac": nbytes / t / PEAK_HBM, "knobs": knobs}))
        bands = fft_frequency_decompose(x, 4096)
        t = timeit(lambda: fft_frequency_recompose(bands, 65536))
        print(json.dumps({"row": "a8 fft_frequency_recompose", "workload": "B=%d x 65536, 5 bands" % B,
                          "us": round(t * 1e6, 2), "bound": "hbm", "achieved_GBs": nbytes / t / 1e9,
                          "frac": nbytes / t / PEAK_HBM, "knobs": knobs}))
        del bands, x


def feed_rows():
    """one training batch from the resident stores (host-side crop drawing +
    plan upload + two ms_gather_crops launches), and the gather alone at a bandwidth-sized batch."""
    import numpy as np
    from music_synthesis_b200.data import DeviceAudioStore, batch_stream
    rs = np.random.RandomState(0)
    chunks = [(rs.random_sample(661500) * 2 - 1).astype(np.float32) for _ in range(8)]   # 8 x 30 s
    store = DeviceAudioStore(chunks)
    spec = {"audio": (8192, 1), "spectrogram": (32, 128)}
    for B in (32, 4096):
        stream = batch_stream(store, B, spec, seed=0)
        t = timeit(lambda: next(stream), n=20)
        nbytes = 2 * 4 * B * (8192 + 128 * 32)
        print(json.dumps({"row": "f3 data feed: batch_stream next()", "workload": "B=%d x (8192 samples + 128x32 log-mel)" % B,
                          "us": round(t * 1e6, 2), "clips_per_s": B / t, "bound": "hbm",
                          "achieved_GBs": nbytes / t / 1e9, "frac": nbytes / t / PEAK_HBM,
                          "note": "includes the host-side crop drawing (python) and the plan upload"}))
        picks, starts = np.zeros(B, dtype=np.int64), np.arange(B, dtype=np.int64) % 2000
        t = timeit(lambda: store.gather(picks, starts, spec), n=20)
        print(json.dumps({"row": "f3 data feed: gather only", "workload": "B=%d" % B, "us": round(t * 1e6, 2),
                          "bound": "hbm", "achieved_GBs": nbytes / t / 1e9, "frac": nbytes / t / PEAK_HBM}))


feed_rows()
# the library reads the knobs per call: default (new kernels), then the A/B settings
run({})
run({"MSB_FFT_FUSE": "0"})
run({"MSB_FFT_TABLE": "0"})
run({"MSB_FFT_MERGE_GATHER": "0"})
run({"MSB_FFT_PACKED": "0"})
run({"MSB_FFT_STAGED": "0"})
run({"MSB_A2M_RADIX4": "1", "MSB_FFT_LEGACY": "1"})
