from .full import FullDiscriminator  # noqa: F401
from .melgan import MelGanDiscriminator  # noqa: F401
from .multiscale import FilterBankChannelDiscriminator, FilterBankMultiScaleDiscriminator  # noqa: F401
from .filterbank import FilterBankDiscriminator  # noqa: F401
