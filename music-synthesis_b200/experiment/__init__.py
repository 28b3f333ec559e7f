from .experiment import BaseGanExperiment, Experiment  # noqa: F401
from .init import weights_init  # noqa: F401
from .melgan import MultiScaleMelGanExperiment  # noqa: F401
from .multiscale import (FilterBankMultiscaleExperiment, MultiScaleNoDeRecompose,  # noqa: F401
                         MultiScaleNoDeRecomposeUnconditionedShortKernel)
from .realmelgan import (Generator, Discriminator, NLayerDiscriminator, ResnetBlock,  # noqa: F401
                         RealMelGanExperiment)
