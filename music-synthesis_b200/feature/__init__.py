from .feature import Audio2Mel  # noqa: F401
