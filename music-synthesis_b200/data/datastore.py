"""GPU-resident training-data feed.

The reference's `batch_stream` (featuresynth/data/datastore.py:19-80) assembles every batch in
Python: pick a random file chunk, read its cached log-mel spectrogram and audio out of LMDB,
`random_slice` the anchor feature, cut the aligned slice of the other feature, zero-pad, stack.
Here the decoded chunks live in HBM -- the audio as one flat buffer and, next to it, the log-mel
spectrogram of every chunk computed ONCE by the fused Audio2Mel kernel (the reference's cached
`spectrogram()` feature, feature/feature.py:79-85) -- and a batch is two `ms_gather_crops`
launches.  Crop positions are drawn on the host with the reference's own two generators
(`random.choice` for the chunk, `numpy.random.randint` for the start), so a seeded stream
reproduces the reference's batches bit for bit (tests/golden/batch_stream.npz).

File decoding / resampling (zounds, librosa, soundfile) and the LMDB cache are out of scope:
the store is built from decoded arrays.
"""
import random

import numpy as np
import torch

from .. import _lib
from .._lib import check, ptr, stream_ptr

AUDIO, SPECTROGRAM = "audio", "spectrogram"


def draw_crops(py_rng, np_rng, anchor_lengths, anchor_size, batch_size):
    """(picks, starts) of one batch.  Per example, in the reference's order: the chunk by
    `random.choice` (datastore.py:48), then `random_slice` (datastore.py:8-16): start uniform in
    [0, len - size), or 0 when the chunk is not longer than the crop."""
    picks = np.empty(batch_size, dtype=np.int64)
    starts = np.empty(batch_size, dtype=np.int64)
    order = range(len(anchor_lengths))
    for i in range(batch_size):
        picks[i] = py_rng.choice(order)
        room = int(anchor_lengths[picks[i]]) - anchor_size
        starts[i] = np_rng.randint(0, room) if room > 0 else 0
    return picks, starts


class DeviceAudioStore:
    """Decoded audio chunks and their log-mel spectrograms, resident on one GPU."""

    def __init__(self, chunks, audio_to_mel=None, spectrograms=None, device="cuda",
                 normalize=False):
        """chunks: 1-D float arrays (what `audio(file_chunk, samplerate)` yields,
        feature/feature.py:62-69; `normalize=True` applies its peak normalisation to 0.95).
        spectrograms: optional precomputed (frames, channels) arrays, one per chunk; otherwise
        `audio_to_mel` (default `Audio2Mel(1024, 256, 1024, 22050, 128)`, feature.py:74-75)
        runs over every chunk on the GPU."""
        chunks = [np.ascontiguousarray(c, dtype=np.float32).reshape(-1) for c in chunks]
        if not chunks:
            raise ValueError("DeviceAudioStore needs at least one chunk")
        if normalize:
            chunks = [c / max(float(np.abs(c).max()), 1e-12) * 0.95 for c in chunks]
            chunks = [c.astype(np.float32) for c in chunks]
        self.device = torch.device(device)
        self.lengths = np.array([len(c) for c in chunks], dtype=np.int64)
        self.audio_offsets = np.concatenate([[0], np.cumsum(self.lengths)[:-1]]).astype(np.int64)
        self.audio = torch.from_numpy(np.concatenate(chunks)).to(self.device)
        if spectrograms is None:
            specs = self._spectrograms(chunks, audio_to_mel)
        else:
            # (frames, channels) as the reference caches them -> channel-major rows
            specs = [torch.from_numpy(np.ascontiguousarray(
                np.asarray(s, dtype=np.float32).T)).to(self.device) for s in spectrograms]
        if len(specs) != len(chunks):
            raise ValueError("one spectrogram per chunk")
        self.channels = int(specs[0].shape[0])
        self.frames = np.array([int(s.shape[1]) for s in specs], dtype=np.int64)
        sizes = self.frames * self.channels
        self.spec_offsets = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
        self.spec = torch.cat([s.reshape(-1) for s in specs])
        if self.spec.numel() == 0:        # every chunk shorter than one frame: all padding
            self.spec = torch.zeros(1, dtype=torch.float32, device=self.device)

    def _spectrograms(self, chunks, audio_to_mel):
        if audio_to_mel is None:
            from ..feature.feature import Audio2Mel
            audio_to_mel = Audio2Mel(1024, 256, 1024, 22050, 128)
        audio_to_mel = audio_to_mel.to(self.device)
        out = []
        with torch.no_grad():
            for off, n in zip(self.audio_offsets, self.lengths):
                x = self.audio[int(off): int(off + n)].view(1, 1, -1)
                out.append(audio_to_mel(x)[0])                    # (channels, frames)
        return out

    def __len__(self):
        return len(self.lengths)

    # -- one batch -------------------------------------------------------------------------
    def crop_plans(self, picks, starts, feature_spec, anchor_feature=SPECTROGRAM):
        """{feature: int64 (B, 3) rows (origin, pitch, valid)} for ms_gather_crops."""
        anchor_size = feature_spec[anchor_feature][0]
        picks = np.asarray(picks, dtype=np.int64)
        starts = np.asarray(starts, dtype=np.int64)
        plans = {}
        for feat, (size, channels) in feature_spec.items():
            ratio = size // anchor_size                      # datastore.py:33-35
            lo = starts * ratio
            if feat == AUDIO:
                if channels != 1:
                    raise ValueError("audio is mono")
                avail, base, pitch = self.lengths[picks], self.audio_offsets[picks], 0 * picks
            elif feat == SPECTROGRAM:
                if channels != self.channels:
                    raise ValueError("spectrogram store has %d channels" % self.channels)
                avail, base, pitch = self.frames[picks], self.spec_offsets[picks], self.frames[picks]
            else:
                raise KeyError("unknown feature %r" % feat)
            valid = np.clip(avail - lo, 0, size)
            plans[feat] = np.stack([base + np.minimum(lo, avail), pitch, valid], axis=1)
        return plans

    def gather(self, picks, starts, feature_spec, anchor_feature=SPECTROGRAM):
        """Batch tensors in `feature_spec` order: audio (B, 1, size), spectrogram
        (B, channels, size) -- the shapes `conform` gives them in the reference."""
        plans = self.crop_plans(picks, starts, feature_spec, anchor_feature)
        flat = np.concatenate([plans[f] for f in feature_spec], axis=0)
        plan_dev = torch.from_numpy(flat).to(self.device, non_blocking=True)
        out, row, B = [], 0, len(picks)
        for feat, (size, channels) in feature_spec.items():
            y = torch.empty((B, channels, size), dtype=torch.float32, device=self.device)
            store = self.audio if feat == AUDIO else self.spec
            check(_lib.lib().ms_gather_crops(ptr(store), ptr(plan_dev[row: row + B]), ptr(y), B,
                                             channels, size, stream_ptr()), "ms_gather_crops")
            out.append(y)
            row += B
        return tuple(out)


def batch_stream(store, batch_size, feature_spec, anchor_feature=SPECTROGRAM, seed=None,
                 rank=0):
    """Endless stream of batches, `feature_spec` order, as CUDA tensors.  `seed` seeds both
    generators the way `random.seed(seed); numpy.random.seed(seed)` seeds the reference's;
    data-parallel ranks pass their rank so every GPU draws different crops."""
    py_rng = random.Random(None if seed is None else seed + rank)
    np_rng = np.random.RandomState(None if seed is None else seed + rank)
    anchor_size = feature_spec[anchor_feature][0]
    anchor_lengths = store.frames if anchor_feature == SPECTROGRAM else store.lengths
    while True:
        picks, starts = draw_crops(py_rng, np_rng, anchor_lengths, anchor_size, batch_size)
        yield store.gather(picks, starts, feature_spec, anchor_feature)
