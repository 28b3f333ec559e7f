from .modules import ResidualAtom, ResidualStack, LearnedUpSample, zero_grad  # noqa: F401
