"""Drop-in mirrors of the shared blocks in featuresynth/util/modules.py.

Same class names, constructor signatures, forward contracts and state-dict keys as
the reference; parameters live in ordinary nn.Conv1d / nn.ConvTranspose1d containers
(so `module.apply(weights_init)`, `.parameters()`, `.state_dict()` and checkpoints
behave identically) but `forward` dispatches to the tcgen05 kernels through the C ABI.
Under torch.no_grad() the fused inference kernels run; with autograd enabled the same modules
record torch.autograd.Functions whose forward and backward are C-ABI calls (see ../autograd.py).
"""
import torch
from torch import nn

from .. import autograd as ag
from .. import ops
from .._lib import MS_CONV, MS_CONVT, MS_F16, MsbError


def weight_norm(m):
    """torch.nn.utils.weight_norm (the legacy hook form the reference uses: `weight_g` /
    `weight_v` parameters, norm over all dims but 0)"""
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return torch.nn.utils.weight_norm(m)


def wn_weight(m):
    """differentiable weight-norm fold g * v / ||v|| of a weight-normed layer on the device
    (ms_weight_norm_fold / ms_weight_norm_bwd); the folded tensor is new on every call, so the
    packed-image caches key on the underlying parameters instead"""
    w = ag.WeightNorm.apply(m.weight_v, m.weight_g)
    w._msb_key = ("wn", m.weight_v.data_ptr(), m.weight_v._version,
                  m.weight_g.data_ptr(), m.weight_g._version)
    return w


def layer_weight(m):
    """the (possibly weight-normed) conv weight of a layer as an autograd-visible tensor"""
    return wn_weight(m) if hasattr(m, "weight_v") else m.weight


def zero_grad(*optims):
    """featuresynth/util/modules.py:38-40."""
    for o in optims:
        o.zero_grad()


class _PackedConv:
    """Caches the packed 16-bit weight image of one conv layer; repacks when the
    parameter is modified in place (optimizer step, load_state_dict, init) or moved."""

    def __init__(self):
        self.key = None
        self.packed = None

    def get(self, desc, weight):
        key = (weight.data_ptr(), weight._version, desc.kind, desc.operand)
        if key != self.key:
            self.packed = ops.pack_conv_weight(desc, weight.detach())
            self.key = key
        return self.packed


def _no_grad_check(*tensors):
    if torch.is_grad_enabled() and any(t.requires_grad for t in tensors):
        raise MsbError(
            "this block has no backward on the sm_100a path yet: call under torch.no_grad()")


class ResidualAtom(nn.Module):
    """featuresynth/util/modules.py:350-388:
    x + leaky(conv_k3_pad1(leaky(conv_k3_dil_d(x))))."""

    def __init__(self, channels, dilation, add_weight_norm=False, operand=MS_F16):
        super().__init__()
        self.add_weight_norm = add_weight_norm
        self.dilation = dilation
        self.channels = channels
        self.operand = operand
        first = nn.Conv1d(channels, channels, 3, 1, dilation=dilation, padding=dilation)
        second = nn.Conv1d(channels, channels, 3, 1, 1)
        if add_weight_norm:         # util/modules.py:366-368 (generator/filterbank.py stacks)
            first, second = weight_norm(first), weight_norm(second)
        self.main = nn.Sequential(first, second)
        self._packed = (_PackedConv(), _PackedConv())
        self._cache = (ag.WeightCache(), ag.WeightCache())

    def forward_blocked_train(self, x32, x16):
        """autograd-recorded form: (x32, x16) -> (y32, y16)"""
        if self.operand != MS_F16:
            raise MsbError("training runs with fp16 forward operands")
        c1, c2 = self.main[0], self.main[1]
        return ag.ResidualAtomBlk.apply(x32, x16, layer_weight(c1), c1.bias, layer_weight(c2),
                                        c2.bias, self._cache[0], self._cache[1], self.dilation)

    def forward_blocked(self, x16, x32):
        """(x16, x32) channel-blocked in -> channel-blocked out (no layout conversion)."""
        B, _, L, _ = x16.shape
        C = self.channels
        d1 = ops.conv_desc(MS_CONV, B, C, C, L, 3, self.dilation, self.dilation, leaky=True,
                           operand=self.operand)
        d2 = ops.conv_desc(MS_CONV, B, C, C, L, 3, 1, 1, leaky=True, operand=self.operand)
        c1, c2 = self.main[0], self.main[1]
        y16, _ = ops.conv_fwd(d1, x16, self._packed[0].get(d1, c1.weight), c1.bias)
        return ops.conv_fwd(d2, y16, self._packed[1].get(d2, c2.weight), c2.bias, res32=x32,
                            want16=True, want32=True)

    def forward(self, x):
        if ag.needs_grad(self, x) or self.add_weight_norm:
            y32, _ = self.forward_blocked_train(ag.PackBlk32.apply(x), ops.pack_ncl(x.detach()))
            return ag.UnpackBlk32.apply(y32)
        x16 = ops.pack_ncl(x, operand=self.operand)
        x32 = _blk32_from_ncl(x)
        _, y32 = self.forward_blocked(x16, x32)
        return ops.unpack_blk32(y32)


def _blk32_from_ncl(x):
    # (B,C,L) -> (B,C/8,L,8) fp32: pure data movement (torch is plumbing here)
    B, C, L = x.shape
    return x.view(B, C // 8, 8, L).permute(0, 1, 3, 2).contiguous()


class ResidualStack(nn.Module):
    """featuresynth/util/modules.py:391-405."""

    def __init__(self, channels, dilations, add_weight_norm=False, operand=MS_F16):
        super().__init__()
        self.dilations = dilations
        self.channels = channels
        self.operand = operand
        self.add_weight_norm = add_weight_norm
        self.main = nn.Sequential(
            *[ResidualAtom(channels, d, add_weight_norm, operand) for d in dilations])
        self._blob = None
        self._blob_key = None

    def _fused_blob(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if key != self._blob_key:
            self._blob = ops.resstack_pack_weights(params, self.channels, self.operand)
            self._blob_key = key
        return self._blob

    def forward_blocked_train(self, x32, x16):
        for atom in self.main:
            x32, x16 = atom.forward_blocked_train(x32, x16)
        return x32, x16

    def forward(self, x):
        if ag.needs_grad(self, x) or self.add_weight_norm:
            y32, _ = self.forward_blocked_train(ag.PackBlk32.apply(x), ops.pack_ncl(x.detach()))
            return ag.UnpackBlk32.apply(y32)
        if (len(self.dilations) == 3 and ops.resstack_supported(self.channels)
                and sum(self.dilations) + 3 <= 16):
            # one fused kernel: activations in SMEM, fp32 residual stream in TMEM
            _, y32 = ops.resstack_fwd(_blk32_from_ncl(x), self._fused_blob(),
                                      list(self.dilations), self.operand)
            return ops.unpack_blk32(y32)
        x16 = ops.pack_ncl(x, operand=self.operand)
        x32 = _blk32_from_ncl(x)
        for atom in self.main:
            x16, x32 = atom.forward_blocked(x16, x32)
        return ops.unpack_blk32(x32)


class LearnedUpSample(nn.Module):
    """featuresynth/util/modules.py:168-188: ConvTranspose1d(k, stride=s,
    padding=(k-s)//2, bias=False) followed by `activation`.  k = 2*s (multiscale generators)
    and k = 4*s (generator/filterbank.py:108-114) with LeakyReLU(0.2) -- the configurations the
    hot-path experiments use -- run as polyphase implicit GEMMs with the activation fused."""

    def __init__(self, in_channels, out_channels, kernel_size, scale_factor, activation=None,
                 operand=MS_F16):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.scale_factor = scale_factor
        self.activation = activation
        self.operand = operand
        if kernel_size % scale_factor != 0 or kernel_size < 2 * scale_factor or \
                (kernel_size - scale_factor) % 2 != 0:
            raise NotImplementedError("LearnedUpSample: kernel_size must be an even multiple "
                                      "(>= 2) of scale_factor")
        self.conv = nn.ConvTranspose1d(in_channels, out_channels, kernel_size,
                                       stride=scale_factor,
                                       padding=(kernel_size - scale_factor) // 2, bias=False)
        self._packed = _PackedConv()

    def forward(self, x, leaky=True):
        _no_grad_check(x, *self.parameters())
        B, C, L = x.shape
        d = ops.conv_desc(MS_CONVT, B, C, self.out_channels, L, self.kernel_size, 1,
                          (self.kernel_size - self.scale_factor) // 2, self.scale_factor,
                          leaky=leaky, operand=self.operand)
        x16 = ops.pack_ncl(x, operand=self.operand)
        _, y32 = ops.conv_fwd(d, x16, self._packed.get(d, self.conv.weight), None,
                              want16=False, want32=True)
        return ops.unpack_blk32(y32)


class UpsamplingStack(nn.Module):
    """featuresynth/util/modules.py:191-223: log_scale(target / start) layers from `layer_func`."""

    def __init__(self, start_size, target_size, scale_factor, layer_func):
        super().__init__()
        import math
        self.layer_func = layer_func
        self.start_size = start_size
        self.target_size = target_size
        self.scale_factor = scale_factor
        n_layers = int(math.log(target_size, scale_factor) - math.log(start_size, scale_factor))
        layers = []
        curr_size = start_size
        for i in range(n_layers):
            out_size = curr_size * scale_factor
            layers.append(layer_func(i, curr_size, out_size, i == 0, i == n_layers - 1))
            curr_size = out_size
        self.main = nn.Sequential(*layers)

    def __iter__(self):
        yield from self.main

    def forward(self, x):
        for layer in self.main:
            x = layer(x)
        return x


class DownsamplingStack(nn.Module):
    """featuresynth/util/modules.py:226-272 (container; the owning discriminator runs the layers
    on the channel-blocked path)."""

    def __init__(self, start_size, target_size, scale_factor, layer_func, activation=None):
        super().__init__()
        import math
        self.activation = activation
        self.scale_factor = scale_factor
        self.target_size = target_size
        self.start_size = start_size
        n_layers = int(math.log(start_size, scale_factor) - math.log(target_size, scale_factor))
        layers = []
        curr_size = start_size
        for i in range(n_layers):
            out_size = curr_size // scale_factor
            layers.append(layer_func(i, curr_size, out_size, i == 0, i == n_layers - 1))
            curr_size = out_size
        self.main = nn.Sequential(*layers)

    def __len__(self):
        return len(self.main)

    def __iter__(self):
        yield from self.main

    @property
    def out_channels(self):
        return self.main[-1].out_channels


def nearest_upsample(feat, size):
    """F.upsample(feat, size=size) (nearest neighbour, the default mode): an index gather"""
    if feat.shape[-1] == size:
        return feat
    idx = (torch.arange(size, device=feat.device) * feat.shape[-1]) // size
    return feat[..., idx].contiguous()


class LowResSpectrogramDiscriminator(nn.Module):
    """featuresynth/util/modules.py:275-344: relu + mean over (channel window x time window) of the
    filter-bank analysis, then stride-2 convs down to `n_judgements` time steps and a judge."""

    def __init__(self, freq_bins, time_steps, n_judgements, kernel_size, max_channels,
                 conditioning_channels=0, log_scaling=False):
        super().__init__()
        import numpy as np
        if log_scaling:
            raise NotImplementedError("LowResSpectrogramDiscriminator: log_scaling is not on this "
                                      "path (no experiment of experiment/filterbank.py sets it)")
        self.log_scaling = log_scaling
        self.conditioning_channels = conditioning_channels
        self.max_channels = max_channels
        self.kernel_size = kernel_size
        self.n_judgements = n_judgements
        self.time_steps = time_steps
        self.freq_bins = freq_bins
        log_channels = np.log2(freq_bins)

        def build(i, curr_size, out_size, first, last):
            cin = min(max_channels, 2 ** (i + log_channels))
            if first:
                cin += conditioning_channels
            cout = min(max_channels, 2 ** (i + log_channels + 1))
            return nn.Conv1d(int(cin), int(cout), kernel_size, stride=2, padding=kernel_size // 2)

        self.stack = DownsamplingStack(time_steps, n_judgements, 2, build)
        self.judge = nn.Conv1d(self.stack.out_channels, 1, 3, 1, 1)
        self._sc = [ag.StridedCache() for _ in self.stack]

    def forward_blocked(self, a32, feat):
        """a32: BLK f32 (B, C/8, L, 8) filter-bank analysis (autograd-tracked when training);
        feat (B, conditioning_channels, T) NCL or None -> ([NCL feature maps], judgement)"""
        from .. import grad_ops
        B, C8, L, _ = a32.shape
        cw, tw = (C8 * 8) // self.freq_bins, L // self.time_steps
        h32, h16 = ag.ReluAvgPool.apply(a32, cw, tw)
        if self.conditioning_channels > 0:
            T = h16.shape[2]
            if feat.shape[-1] < T:
                feat = nearest_upsample(feat, T)
            elif feat.shape[-1] > T:
                f = feat.shape[-1] // T
                feat = ops.avg_pool1d(feat, f, f, 0)
            h32 = torch.cat([h32, grad_ops.pack_ncl32(feat)], dim=1)
            h16 = torch.cat([h16, ops.pack_ncl(feat)], dim=1)
        features = []
        length = h16.shape[2]
        for conv, sc in zip(self.stack, self._sc):
            h32, h16 = ag.StridedConvBlk.apply(h32, h16, conv.weight, conv.bias, sc, 2, length)
            features.append(ag.UnpackBlk32.apply(h32))
            length = h16.shape[2]
        j = ag.MonoConv.apply(h32, self.judge.weight, self.judge.bias, 3, 1, False)
        return features, j
