"""TEST INFRASTRUCTURE ONLY.  Seeded synthetic inputs shared by the fixture
generator, the oracle tests, the GPU parity tests and bench.py.  Everything is
drawn from numpy RandomState (legacy generator: bit-stable across numpy versions
and machines), never from torch's RNG."""
import numpy as np
import torch


def randn(seed, *shape):
    rs = np.random.RandomState(seed)
    return torch.from_numpy(rs.standard_normal(shape).astype(np.float32))


def mel_features(seed, batch, frames, channels=128):
    """Synthetic conditioning features (B, 128, T) ~ N(0,1)."""
    return randn(seed, batch, channels, frames)


def uniform_audio(seed, batch, samples):
    """Uniform audio in [-1, 1)."""
    rs = np.random.RandomState(seed)
    return torch.from_numpy(
        (rs.random_sample((batch, 1, samples)) * 2 - 1).astype(np.float32))


def residual_stack_state(seed, channels, prefix="s", bias_std=0.01):
    rs = np.random.RandomState(seed)
    sd = {}
    for a in range(3):
        for c in range(2):
            sd[f"{prefix}.main.{a}.main.{c}.weight"] = torch.from_numpy(
                (rs.standard_normal((channels, channels, 3)) * 0.02).astype(np.float32))
            sd[f"{prefix}.main.{a}.main.{c}.bias"] = torch.from_numpy(
                (rs.standard_normal((channels,)) * bias_std).astype(np.float32))
    return sd


def feed_chunks(seed, lengths=(6000, 3000, 1200, 400), channels=16, hop=64, frame_pad=192):
    """Decoded 'file chunks' for the data-feed tests: audio (n,) and a stand-in cached
    spectrogram (frames, channels) per chunk, frames = (n + frame_pad - 4*hop)//hop + 1 like
    Audio2Mel's framing.  The last chunks are shorter than a crop (zero-padding cases)."""
    rs = np.random.RandomState(seed)
    audio, spec = [], []
    for n in lengths:
        audio.append((rs.random_sample(n) * 2 - 1).astype(np.float32))
        frames = max((n + frame_pad - 4 * hop) // hop + 1, 0)
        spec.append(rs.standard_normal((frames, channels)).astype(np.float32))
    return audio, spec
