from .experiment import BaseGanExperiment, Experiment  # noqa: F401
from .init import weights_init  # noqa: F401
from .realmelgan import Generator, Discriminator, NLayerDiscriminator, ResnetBlock  # noqa: F401
from .wirings import (FilterBankMultiscaleExperiment, MultiScaleMelGanExperiment,  # noqa: F401
                      MultiScaleNoDeRecompose, MultiScaleNoDeRecomposeUnconditionedShortKernel,
                      RealMelGanExperiment)
