from .transform import fft_frequency_decompose, fft_frequency_recompose  # noqa: F401
