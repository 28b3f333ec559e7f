from .train import GeneratorTrainer, DiscriminatorTrainer, training_loop  # noqa: F401
from .optim import Adam  # noqa: F401
