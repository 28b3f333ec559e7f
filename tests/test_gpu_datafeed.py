"""-m gpu: the GPU-resident data feed against three batches of the
unmodified reference's `batch_stream` (tests/golden/batch_stream.npz) -- bit-exact, it is pure
data movement -- and with the log-mel store computed by the fused Audio2Mel kernel."""
import numpy as np
import pytest
import torch

from oracle import restate, synth

pytestmark = pytest.mark.gpu

FEED_SPEC = {"audio": (2048, 1), "spectrogram": (32, 16)}


def test_batch_stream_matches_reference_batches(golden):
    from music_synthesis_b200.data import DeviceAudioStore, batch_stream
    g = golden("batch_stream")
    audio, spec = synth.feed_chunks(int(g["chunk_seed"]))
    store = DeviceAudioStore(audio, spectrograms=spec)
    stream = batch_stream(store, int(g["batch_size"]), FEED_SPEC, "spectrogram", seed=int(g["seed"]))
    for i in range(3):
        a, s = next(stream)
        assert a.is_cuda and a.shape == (6, 1, 2048) and s.shape == (6, 16, 32)
        assert np.array_equal(a.cpu().numpy(), g[f"audio_{i}"])
        assert np.array_equal(s.cpu().numpy(), g[f"spectrogram_{i}"])


def test_store_computes_log_mel_on_the_gpu_and_ranks_draw_different_crops():
    """cfg4 shapes: 8192-sample crops with their 32 aligned log-mel frames, the store's
    spectrograms coming from Audio2Mel over whole chunks (feature/feature.py:79-85)."""
    from music_synthesis_b200.data import DeviceAudioStore, batch_stream
    from music_synthesis_b200.feature.feature import Audio2Mel
    rs = np.random.RandomState(5)
    chunks = [(rs.random_sample(n) * 2 - 1).astype(np.float32) for n in (66150, 30000, 5000)]
    store = DeviceAudioStore(chunks)
    a2m = Audio2Mel(1024, 256, 1024, 22050, 128).cuda()
    full = [a2m(torch.from_numpy(c).view(1, 1, -1).cuda())[0].cpu().numpy().T for c in chunks]
    assert [f.shape[0] for f in full] == list(store.frames)
    spec = {"audio": (8192, 1), "spectrogram": (32, 128)}
    ref = restate.batch_stream(chunks, full, 16, spec, "spectrogram", 3, 2)
    stream = batch_stream(store, 16, spec, "spectrogram", seed=3)
    for i in range(2):
        a, s = next(stream)
        assert np.array_equal(a.cpu().numpy(), ref[i][0])
        assert np.array_equal(s.cpu().numpy(), ref[i][1])
    # 9 frames: an odd crop length takes the one-sample-per-thread kernel, its 2304 audio samples
    # the four-per-thread one
    odd = {"audio": (2304, 1), "spectrogram": (9, 128)}
    ref_odd = restate.batch_stream(chunks, full, 5, odd, "spectrogram", 4, 1)
    a, s = next(batch_stream(store, 5, odd, "spectrogram", seed=4))
    assert np.array_equal(a.cpu().numpy(), ref_odd[0][0])
    assert np.array_equal(s.cpu().numpy(), ref_odd[0][1])
    other = next(batch_stream(store, 16, spec, "spectrogram", seed=3, rank=1))
    assert not np.array_equal(other[0].cpu().numpy(), ref[0][0])
    # peak normalisation of `audio()` (feature/feature.py:67)
    n = DeviceAudioStore([c * 0.3 for c in chunks], normalize=True)
    assert abs(float(n.audio[: len(chunks[0])].abs().max()) - 0.95) < 1e-6
