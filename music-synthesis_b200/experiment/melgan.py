"""Import path of the reference (featuresynth/experiment/melgan.py); the wiring itself is in
wirings.py."""
from .wirings import MultiScaleMelGanExperiment  # noqa: F401
