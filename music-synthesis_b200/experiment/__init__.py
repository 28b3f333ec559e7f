from .realmelgan import Generator, Discriminator, NLayerDiscriminator, ResnetBlock  # noqa: F401
