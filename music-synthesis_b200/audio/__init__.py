from .transform import fft_frequency_decompose, fft_frequency_recompose  # noqa: F401
from .representation import BaseAudioRepresentation, MultiScale, RawAudio  # noqa: F401
