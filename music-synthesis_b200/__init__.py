"""b200-wavesynth: B200-native (sm_100a) implementation of the stage-two hot path of
JohnVinyard/music-synthesis -- conditional waveform synthesis -- behind the
reference's own nn.Module API.  See DESIGN.md.

    from music_synthesis_b200.generator.full import MelGanGenerator
    from music_synthesis_b200.feature.feature import Audio2Mel

Arithmetic runs in csrc/libmsb200.so (hand-written CUDA: tcgen05/TMEM implicit-GEMM
convolutions fed by the TMA unit, fused epilogues).  There is no CPU fallback.
"""
__version__ = "0.1.0"

from . import _lib  # noqa: F401
