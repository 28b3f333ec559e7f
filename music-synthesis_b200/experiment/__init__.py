import sys as _sys

from .experiment import BaseGanExperiment, Experiment  # noqa: F401
from .init import weights_init  # noqa: F401
from .realmelgan import Generator, Discriminator, NLayerDiscriminator, ResnetBlock  # noqa: F401
from . import wirings as _wirings
from .wirings import (AlternateFilterBankExperiment, ConditionalFilterBankExperiment,  # noqa: F401
                      FilterBankExperiment,
                      FilterBankMultiscaleExperiment, MultiScaleMelGanExperiment,
                      MultiScaleNoDeRecompose, MultiScaleNoDeRecomposeUnconditionedShortKernel,
                      RealMelGanExperiment)

# the reference spreads its experiment classes over experiment/{melgan,multiscale,filterbank}.py;
# here they are data in wirings.py -- keep the reference's import paths as aliases of that module
for _name in ("melgan", "multiscale", "filterbank"):
    _sys.modules[__name__ + "." + _name] = _wirings
