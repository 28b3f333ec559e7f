"""The reference's initialiser contract (featuresynth/experiment/init.py:3-9), the "random-init
weights" of every parity test: modules whose class name contains "Conv" get N(0, 0.02) weights
and, when they have one, a zero bias.  Used through `module.apply(weights_init)` exactly like
the reference's Experiment does (experiment/experiment.py:109-115)."""
import torch


@torch.no_grad()
def weights_init(module, std=0.02):
    if "Conv" not in type(module).__name__:
        return
    torch.nn.init.normal_(module.weight, mean=0.0, std=std)
    bias = getattr(module, "bias", None)
    if bias is not None:
        torch.nn.init.zeros_(bias)
