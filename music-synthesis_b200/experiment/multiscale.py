"""Import path of the reference (featuresynth/experiment/multiscale.py); the wirings this path
covers are in wirings.py."""
from .wirings import (FilterBankMultiscaleExperiment, MultiScaleNoDeRecompose,  # noqa: F401
                      MultiScaleNoDeRecomposeUnconditionedShortKernel)
