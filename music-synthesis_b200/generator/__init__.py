from .full import MelGanGenerator  # noqa: F401
