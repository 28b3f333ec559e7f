"""Drop-in mirror of featuresynth/generator/full.py:16-50 (MelGanGenerator).

Same constructor, `forward(x (B,in_channels,T)) -> (B,1,256T)` and the 60 state-dict
keys (`main.1.weight` ... `main.15.bias`); the forward pass is one
call into the C ABI (`ms_melgan_generator_fwd`): tcgen05 implicit-GEMM convolutions
with fused bias / LeakyReLU / residual epilogues and an fp32 residual stream.
"""
import torch
from torch import nn

from .. import ops
from .._lib import MS_F16, MsbError
from ..util.modules import ResidualStack


class MelGanGenerator(nn.Module):
    #: clips per pass through the layer schedule (bounds the activation workspace)
    clips_per_pass = 256

    def __init__(self, input_size, in_channels, operand=MS_F16):
        super().__init__()
        self.in_channels = in_channels
        self.input_size = input_size
        self.operand = operand
        self.main = nn.Sequential(
            nn.ReflectionPad1d(3),
            nn.Conv1d(in_channels, 512, 7, 1, padding=0),
            nn.LeakyReLU(0.2),

            nn.ConvTranspose1d(512, 256, 16, 8, 4),
            nn.LeakyReLU(0.2),
            ResidualStack(256, [1, 3, 9], operand=operand),

            nn.ConvTranspose1d(256, 128, 16, 8, 4),
            nn.LeakyReLU(0.2),
            ResidualStack(128, [1, 3, 9], operand=operand),

            nn.ConvTranspose1d(128, 64, 4, 2, 1),
            nn.LeakyReLU(0.2),
            ResidualStack(64, [1, 3, 9], operand=operand),

            nn.ConvTranspose1d(64, 32, 4, 2, 1),
            nn.LeakyReLU(0.2),
            ResidualStack(32, [1, 3, 9], operand=operand),

            nn.Conv1d(32, 1, 7, 1, 3),
            nn.Tanh(),
        )
        self._weights = None
        self._weights_key = None
        self._workspace = None

    # ---- packed-weight cache ------------------------------------------------
    def _packed_weights(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params) + (self.operand,)
        if key != self._weights_key:
            self._weights = ops.MelGanWeights(params, self.in_channels, self.operand)
            self._weights_key = key
        return self._weights

    def _get_workspace(self, batch, frames, device):
        need = ops.melgan_workspace_bytes(min(batch, self.clips_per_pass), frames,
                                          self.in_channels)
        ws = self._workspace
        if ws is None or ws.numel() < need or ws.device != device:
            ws = self._workspace = torch.empty(need, dtype=torch.uint8, device=device)
        return ws

    def forward(self, x):
        if torch.is_grad_enabled() and (x.requires_grad or
                                        any(p.requires_grad for p in self.parameters())):
            raise MsbError("MelGanGenerator (sm_100a path) is forward-only in this build: "
                           "call under torch.no_grad()")
        if x.dim() != 3 or x.shape[1] != self.in_channels:
            raise MsbError("expected (B, %d, T) features" % self.in_channels)
        ws = self._get_workspace(x.shape[0], x.shape[2], x.device)
        return ops.melgan_generator_fwd(self._packed_weights(), x, ws)
