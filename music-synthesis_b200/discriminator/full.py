"""Drop-in mirror of featuresynth/discriminator/full.py:10-40 (FullDiscriminator).

Same layer stack, state-dict keys (`main.{0..5}.{weight,bias}`, `judge.{weight,bias}`) and
forward contract: x (B,1,N) -> ([6 feature maps], score).  `weight_norm` is the identity in
the reference (full.py:6-7).  With autograd enabled the same kernels are recorded as
torch.autograd.Functions with C-ABI backward passes (../autograd.py).

Kernels: the 1->16 k15 conv and the four k41 stride-4 grouped convs (4 input channels per
group) run as fp32 CUDA-core direct convolutions (`ms_conv1d_direct_fwd`); the dense
1024->1024 k5 layer (70 % of the FLOPs) is the tcgen05 implicit GEMM (`ms_conv_fwd`) over a
two-term fp16 split of its input (see autograd.dense_split_fwd); the
1024->1 judge is `ms_conv_to_mono`.
"""
import torch
from torch import nn

from .. import autograd as ag
from .. import ops
from .._lib import MS_CONV, MS_F16, MsbError
from ..util.modules import _PackedConv


class FullDiscriminator(nn.Module):
    #: dense layer in split precision (3 fp16 passes, ~1.7e-5 instead of ~3e-4 on its outputs).
    #: The top-layer activations of the real and the fake batch differ by ~1e-4 relative while a
    #: GAN step differentiates exactly that difference (feature matching, hinge cancellation).
    split_precision = True

    def __init__(self, operand=MS_F16):
        super().__init__()
        self.operand = operand
        self.main = nn.Sequential(
            nn.Conv1d(1, 16, 15, 1, padding=7),
            nn.Conv1d(16, 64, 41, 4, padding=20, groups=4),
            nn.Conv1d(64, 256, 41, 4, padding=20, groups=16),
            nn.Conv1d(256, 1024, 41, 4, padding=20, groups=64),
            nn.Conv1d(1024, 1024, 41, 4, padding=20, groups=256),
            nn.Conv1d(1024, 1024, 5, 1, padding=2),
        )
        self.judge = nn.Conv1d(1024, 1, 3, 1, padding=1)
        self._packed = _PackedConv()
        self._cache = ag.WeightCache()

    def _forward_train(self, x):
        features = []
        for layer in list(self.main)[:5]:
            x = ag.DirectConv.apply(x, layer.weight, layer.bias, layer.stride[0], layer.padding[0],
                                    layer.groups, True)
            features.append(x)
        dense = self.main[5]
        y32 = ag.DenseConvNCL.apply(x, dense.weight, dense.bias, self._cache, dense.padding[0], True)
        features.append(ag.UnpackBlk32.apply(y32))
        j = ag.MonoConv.apply(y32, self.judge.weight, self.judge.bias, self.judge.kernel_size[0],
                              self.judge.padding[0], False)
        return features, j

    def forward(self, x):
        if ag.needs_grad(self, x):
            if self.operand != MS_F16:
                raise MsbError("training runs with fp16 forward operands")
            return self._forward_train(x.contiguous())
        features = []
        for layer in list(self.main)[:5]:
            x = ops.conv1d_direct(x, layer.weight, layer.bias, stride=layer.stride[0],
                                  pad=layer.padding[0], groups=layer.groups, leaky=True)
            features.append(x)
        dense = self.main[5]
        if self.split_precision:
            _, y32 = ag.dense_split_fwd(x, dense.weight, dense.bias, self._cache, dense.padding[0],
                                        True, want16=False)
        else:
            B, C, L = x.shape
            d = ops.conv_desc(MS_CONV, B, C, dense.out_channels, L, dense.kernel_size[0], 1,
                              dense.padding[0], leaky=True, operand=self.operand)
            _, y32 = ops.conv_fwd(d, ops.pack_ncl(x, operand=self.operand),
                                  self._packed.get(d, dense.weight), dense.bias,
                                  want16=False, want32=True)
        features.append(ops.unpack_blk32(y32))
        j = ops.conv_to_mono(y32, self.judge.weight, self.judge.bias,
                             self.judge.kernel_size[0], self.judge.padding[0], False)
        return features, j
