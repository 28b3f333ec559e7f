from .full import MelGanGenerator  # noqa: F401
from .multiscale import FilterBankChannelGenerator, FilterBankMultiScaleGenerator  # noqa: F401
from .filterbank import FilterBankGenerator, ResidualStackFilterBankGenerator  # noqa: F401
