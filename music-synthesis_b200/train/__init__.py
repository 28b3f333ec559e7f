from .train import GeneratorTrainer, DiscriminatorTrainer  # noqa: F401
from .optim import Adam  # noqa: F401
