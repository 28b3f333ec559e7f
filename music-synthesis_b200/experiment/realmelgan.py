"""Drop-in mirrors of the official-MelGAN pair in featuresynth/experiment/realmelgan.py:
`ResnetBlock` (32-45), `Generator` (48-89), `NLayerDiscriminator` (92-155), `Discriminator`
(158-181), `real_mel_gan_feature_loss` / `mel_gan_gen_loss` (185-217).

Same constructors, forward contracts and the legacy `torch.nn.utils.weight_norm` state-dict
layout (`*.bias`, `*.weight_g`, `*.weight_v`; for ConvTranspose1d the norm runs over dim 0 =
in_channels).  The weight-norm fold g*v/||v|| is done on the device when the 16-bit weight image
is (re)packed.  Generator chain, channel-blocked on the tcgen05 kernel:
  reflect-pad k7 conv -> 4 x [LeakyReLU + ConvTranspose (polyphase GEMM) -> 3 x ResnetBlock
  (1x1 shortcut conv; LeakyReLU + ReflectionPad(d) fused in one layout kernel; dilated k3 conv
  with fused LeakyReLU; 1x1 conv with the shortcut added in its fp32 epilogue)] ->
  LeakyReLU + ReflectionPad(3) -> fp32 32->1 k7 conv + tanh.
Discriminator: reflect-padded direct conv, grouped strided direct convs, tcgen05 dense k5 conv,
fp32 judge, AvgPool1d(4,2,1,count_include_pad=False) between the three independent scales.
With autograd enabled the same kernels are recorded as torch.autograd.Functions with C-ABI backward
passes (../autograd.py): `WeightNorm` (fold + its gradient to weight_g / weight_v), `ActPadBlk`
(LeakyReLU + reflection pad and the scatter of its gradient), `ConvBlk` with a residual input (the
ResnetBlock shortcut), `DenseConvNCL`, `DirectConv`, `MonoConv`, `AvgPool`.
"""
import numpy as np
import torch
import torch.nn as nn
import warnings

import torch.nn.functional as F

from .. import autograd as ag
from .. import ops
from .._lib import MS_CONV, MS_CONVT, MS_F16, MsbError
from ..loss.loss import _Acc, L1, hinge_generator_loss


from ..util.modules import weight_norm as _weight_norm, wn_weight as _wn  # noqa: E402


def WNConv1d(*args, **kwargs):
    return _weight_norm(nn.Conv1d(*args, **kwargs))


def WNConvTranspose1d(*args, **kwargs):
    return _weight_norm(nn.ConvTranspose1d(*args, **kwargs))


class _PackedWN:
    """packed 16-bit image of a weight-normed conv; refolded when g or v change"""

    def __init__(self):
        self.key = None
        self.val = None

    def folded(self, m):
        key = (m.weight_g.data_ptr(), m.weight_g._version, m.weight_v.data_ptr(), m.weight_v._version)
        if key != self.key:
            self.val = [ops.weight_norm_fold(m.weight_v.detach(), m.weight_g.detach()), None, None]
            self.key = key
        return self.val

    def get(self, desc, m):
        val = self.folded(m)
        sig = (desc.kind, desc.ksize, desc.dilation, desc.stride, desc.operand)
        if val[1] is None or val[2] != sig:
            val[1] = ops.pack_conv_weight(desc, val[0])
            val[2] = sig
        return val[1]


def _fwd_only(module, x):
    if torch.is_grad_enabled() and (x.requires_grad or
                                    any(p.requires_grad for p in module.parameters())):
        raise MsbError("sm_100a path is forward-only in this build: use torch.no_grad()")


class ResnetBlock(nn.Module):
    def __init__(self, dim, dilation=1):
        super().__init__()
        self.dim = dim
        self.dilation = dilation
        self.block = nn.Sequential(
            nn.LeakyReLU(0.2),
            nn.ReflectionPad1d(dilation),
            WNConv1d(dim, dim, kernel_size=3, dilation=dilation),
            nn.LeakyReLU(0.2),
            WNConv1d(dim, dim, kernel_size=1),
        )
        self.shortcut = WNConv1d(dim, dim, kernel_size=1)
        self._p = (_PackedWN(), _PackedWN(), _PackedWN())
        self._c = (ag.WeightCache(), ag.WeightCache(), ag.WeightCache())

    def forward_blocked_train(self, x32, x16):
        """autograd-recorded form: (x32, x16) -> (y32, y16)"""
        d = self.dilation
        s32, _ = ag.conv_blk(x32, x16, _wn(self.shortcut), self.shortcut.bias, self._c[2], MS_CONV,
                             1, 0, 1, False)
        a32, a16 = ag.ActPadBlk.apply(x32, x16, d, 1, True)
        c1, c2 = self.block[2], self.block[4]
        h32, h16 = ag.conv_blk(a32, a16, _wn(c1), c1.bias, self._c[0], MS_CONV, d, 0, 1, True)
        return ag.conv_blk(h32, h16, _wn(c2), c2.bias, self._c[1], MS_CONV, 1, 0, 1, False, s32)

    def forward_blocked(self, x16):
        """x16 (B, dim/8, L, 8) un-activated 16-bit operand -> (x16', x32')"""
        B, _, L, _ = x16.shape
        C, d = self.dim, self.dilation
        k1 = ops.conv_desc(MS_CONV, B, C, C, L, 1, 1, 0)
        _, s32 = ops.conv_fwd(k1, x16, self._p[2].get(k1, self.shortcut), self.shortcut.bias,
                              want16=False, want32=True)
        a16 = ops.act_pad(x16, d, 1, leaky=True)                  # LeakyReLU + ReflectionPad1d(d)
        k3 = ops.conv_desc(MS_CONV, B, C, C, L + 2 * d, 3, d, 0, leaky=True)   # + next LeakyReLU
        h16, _ = ops.conv_fwd(k3, a16, self._p[0].get(k3, self.block[2]), self.block[2].bias)
        return ops.conv_fwd(k1, h16, self._p[1].get(k1, self.block[4]), self.block[4].bias,
                            res32=s32, want16=True, want32=True)

    def forward(self, x):
        if ag.needs_grad(self, x):
            y32, _ = self.forward_blocked_train(ag.PackBlk32.apply(x), ops.pack_ncl(x.detach()))
            return ag.UnpackBlk32.apply(y32)
        _, y32 = self.forward_blocked(ops.pack_ncl(x))
        return ops.unpack_blk32(y32)


class Generator(nn.Module):
    def __init__(self, input_size, ngf, n_residual_layers):
        super().__init__()
        ratios = [8, 8, 2, 2]
        self.hop_length = np.prod(ratios)
        mult = int(2 ** len(ratios))
        model = [nn.ReflectionPad1d(3), WNConv1d(input_size, mult * ngf, kernel_size=7, padding=0)]
        for i, r in enumerate(ratios):
            model += [nn.LeakyReLU(0.2),
                      WNConvTranspose1d(mult * ngf, mult * ngf // 2, kernel_size=r * 2, stride=r,
                                        padding=r // 2 + r % 2, output_padding=r % 2)]
            for j in range(n_residual_layers):
                model += [ResnetBlock(mult * ngf // 2, dilation=3 ** j)]
            mult //= 2
        model += [nn.LeakyReLU(0.2), nn.ReflectionPad1d(3), WNConv1d(ngf, 1, kernel_size=7, padding=0),
                  nn.Tanh()]
        self.model = nn.Sequential(*model)
        self._p = {}

    def _packed(self, idx):
        if idx not in self._p:
            self._p[idx] = _PackedWN()
        return self._p[idx]

    def _forward_train(self, x):
        if x.requires_grad:
            raise MsbError("gradients w.r.t. the conditioning features are not on this path")
        layers = list(self.model)
        caches = self.__dict__.setdefault("_c", {})
        x16 = ops.pack_ncl(x, 3, 1)                                   # ReflectionPad1d(3)
        first = layers[1]
        h32, h16 = ag.conv_blk(None, x16, _wn(first), first.bias,
                               caches.setdefault(1, ag.WeightCache()), MS_CONV, 1, 0, 1, False)
        L = x.shape[-1]
        i = 2
        while i < len(layers):
            m = layers[i]
            if isinstance(m, nn.ConvTranspose1d):
                r = m.stride[0]
                if m.output_padding[0] != 0 or m.kernel_size[0] != 2 * r:
                    raise NotImplementedError("odd upsampling ratios are not on this path")
                a32, a16 = ag.ActPadBlk.apply(h32, h16, 0, 0, True)   # the LeakyReLU before it
                h32, h16 = ag.conv_blk(a32, a16, _wn(m), m.bias, caches.setdefault(i, ag.WeightCache()),
                                       MS_CONVT, 1, m.padding[0], r, False)
                L *= r
            elif isinstance(m, ResnetBlock):
                h32, h16 = m.forward_blocked_train(h32, h16)
            elif isinstance(m, nn.Conv1d):                            # final k7 conv + tanh
                a32, _ = ag.ActPadBlk.apply(h32, h16, 3, 1, True)     # LeakyReLU + ReflectionPad1d(3)
                # valid conv over the padded stream: the first L outputs are the layer's output
                return ag.MonoConv.apply(a32, _wn(m), m.bias, 7, 0, True)[:, :, :L].contiguous()
            i += 1
        raise MsbError("malformed generator")

    #: "fast" (fp16 operands) | "exact" (three-term bf16 split, ops.ExactConv) | "auto"; see
    #: generator/full.py::MelGanGenerator.precision.  The weight-normed generator's activations
    #: sit around 3e-6 rms at init -- inside the fp16 subnormals -- and still make the 1e-3 bar in
    #: the fast mode (2.9e-4); a trained checkpoint with other scales can use "exact".
    precision = "fast"
    auto_tolerance = 5e-4

    @torch.no_grad()
    def _forward_exact(self, x):
        from .. import grad_ops
        ex = self.__dict__.setdefault("_exact", {})

        def conv(name):
            return ex.setdefault(name, ops.ExactConv())

        def w(m):
            t = ops.weight_norm_fold(m.weight_v.detach(), m.weight_g.detach())
            t._msb_key = ("wn", m.weight_v.data_ptr(), m.weight_v._version,
                          m.weight_g.data_ptr(), m.weight_g._version)
            return t

        layers = list(self.model)
        h = grad_ops.pack_ncl32(x.contiguous())
        first = layers[1]
        h = conv(1)(h, w(first), first.bias, MS_CONV, 1, 0, 1, False, pad_in=3, pad_mode=1)
        L = x.shape[-1]
        i = 2
        while i < len(layers):
            m = layers[i]
            if isinstance(m, nn.ConvTranspose1d):
                r = m.stride[0]
                h = conv(i)(h, w(m), m.bias, MS_CONVT, 1, m.padding[0], r, False, act_in=True)
                L *= r
            elif isinstance(m, ResnetBlock):
                d = m.dilation
                s32 = conv((i, "s"))(h, w(m.shortcut), m.shortcut.bias, MS_CONV, 1, 0, 1, False)
                t = conv((i, 0))(h, w(m.block[2]), m.block[2].bias, MS_CONV, d, 0, 1, True,
                                 act_in=True, pad_in=d, pad_mode=1)
                h = conv((i, 1))(t, w(m.block[4]), m.block[4].bias, MS_CONV, 1, 0, 1, False, res32=s32)
            elif isinstance(m, nn.Conv1d):
                a32 = ops.act_pad(h, 3, 1, leaky=True)
                y = ops.conv_to_mono(a32, w(m), m.bias, 7, 0, True)
                return y[:, :, :L].contiguous()
            i += 1
        raise MsbError("malformed generator")

    def forward(self, x):
        if ag.needs_grad(self, x):
            return self._forward_train(x.contiguous())
        if self.precision == "exact":
            return self._forward_exact(x)
        if self.precision == "auto":
            from ..generator.full import auto_precision_ok
            if not auto_precision_ok(self, x, self._forward_fast):
                return self._forward_exact(x)
        return self._forward_fast(x)

    def _forward_fast(self, x):
        B, _, T = x.shape
        layers = list(self.model)
        x16 = ops.pack_ncl(x, 3, 1)                                   # ReflectionPad1d(3)
        first = layers[1]
        d = ops.conv_desc(MS_CONV, B, first.in_channels, first.out_channels, T + 6, 7, 1, 0)
        h16, h32 = ops.conv_fwd(d, x16, self._packed(1).get(d, first), first.bias)
        L = T
        i = 2
        while i < len(layers):
            m = layers[i]
            if isinstance(m, nn.ConvTranspose1d):
                r = m.stride[0]
                if m.output_padding[0] != 0 or m.kernel_size[0] != 2 * r:
                    raise NotImplementedError("odd upsampling ratios are not on this path")
                a16 = ops.act_pad(h16, 0, 0, leaky=True)              # the LeakyReLU before it
                d = ops.conv_desc(MS_CONVT, B, m.in_channels, m.out_channels, L, 2 * r, 1,
                                  m.padding[0], r)
                h16, h32 = ops.conv_fwd(d, a16, self._packed(i).get(d, m), m.bias,
                                        want16=True, want32=True)
                L *= r
            elif isinstance(m, ResnetBlock):
                h16, h32 = m.forward_blocked(h16)
            elif isinstance(m, nn.Conv1d):                            # final k7 conv, 1 channel
                a32 = ops.act_pad(h32, 3, 1, leaky=True)              # LeakyReLU + ReflectionPad1d(3)
                w = self._packed(i).folded(m)[0]
                y = ops.conv_to_mono(a32, w, m.bias, 7, 0, True)      # + Tanh
                return y[:, :, :L].contiguous()
            i += 1
        raise MsbError("malformed generator")


class NLayerDiscriminator(nn.Module):
    def __init__(self, ndf, n_layers, downsampling_factor, conditioning_channels=0):
        super().__init__()
        if conditioning_channels != 0:
            raise NotImplementedError("conditioned NLayerDiscriminator is not on this path yet")
        self.conditioning_channels = conditioning_channels
        model = nn.ModuleDict()
        model["layer_0"] = nn.Sequential(nn.ReflectionPad1d(7), WNConv1d(1, ndf, kernel_size=15),
                                         nn.LeakyReLU(0.2, True))
        nf = ndf
        stride = downsampling_factor
        for n in range(1, n_layers + 1):
            nf_prev = nf
            nf = min(nf * stride, 1024)
            model["layer_%d" % n] = nn.Sequential(
                WNConv1d(nf_prev, nf, kernel_size=stride * 10 + 1, stride=stride, padding=stride * 5,
                         groups=nf_prev // 4), nn.LeakyReLU(0.2, True))
        nf = min(nf * 2, 1024)
        model["layer_%d" % (n_layers + 1)] = nn.Sequential(
            WNConv1d(nf_prev, nf, kernel_size=5, stride=1, padding=2), nn.LeakyReLU(0.2, True))
        model["layer_%d" % (n_layers + 2)] = WNConv1d(nf, 1, kernel_size=3, stride=1, padding=1)
        self.model = model
        self.n_layers = n_layers
        self._p = {k: _PackedWN() for k in model}
        self._dense_cache = ag.WeightCache()

    def _forward_train(self, x):
        results = []
        keys = list(self.model.keys())
        conv0 = self.model["layer_0"][1]
        # ReflectionPad1d(7) on the 1-channel input (ms_reflect_pad_ncl + its gather backward)
        h = ag.DirectConv.apply(ag.ReflectPadNCL.apply(x, 7), _wn(conv0), conv0.bias, 1, 0, 1, True)
        results.append(h)
        for n in range(1, self.n_layers + 1):
            c = self.model["layer_%d" % n][0]
            h = ag.DirectConv.apply(h, _wn(c), c.bias, c.stride[0], c.padding[0], c.groups, True)
            results.append(h)
        dense = self.model[keys[-2]][0]
        y32 = ag.DenseConvNCL.apply(h, _wn(dense), dense.bias, self._dense_cache, dense.padding[0],
                                    True)
        results.append(ag.UnpackBlk32.apply(y32))
        judge = self.model[keys[-1]]
        results.append(ag.MonoConv.apply(y32, _wn(judge), judge.bias, judge.kernel_size[0],
                                         judge.padding[0], False))
        return results

    def forward(self, x, feat):
        if ag.needs_grad(self, x):
            return self._forward_train(x.contiguous())
        results = []
        keys = list(self.model.keys())
        conv0 = self.model["layer_0"][1]
        h = ops.conv1d_direct(x, self._p["layer_0"].folded(conv0)[0], conv0.bias, 1, 7, 1,
                              leaky=True, pad_mode=1)
        results.append(h)
        for n in range(1, self.n_layers + 1):
            c = self.model["layer_%d" % n][0]
            h = ops.conv1d_direct(h, self._p["layer_%d" % n].folded(c)[0], c.bias, c.stride[0],
                                  c.padding[0], c.groups, leaky=True)
            results.append(h)
        dense = self.model[keys[-2]][0]
        # split-precision dense layer, like the training path: the real and the fake batch must
        # see the SAME arithmetic (the feature-matching loss differences them)
        wd = self._p[keys[-2]].folded(dense)[0]
        wd._msb_key = ("wn", dense.weight_v.data_ptr(), dense.weight_v._version,
                       dense.weight_g.data_ptr(), dense.weight_g._version)
        _, y32 = ag.dense_split_fwd(h, wd, dense.bias, self._dense_cache, dense.padding[0], True,
                                    want16=False)
        results.append(ops.unpack_blk32(y32))
        judge = self.model[keys[-1]]
        results.append(ops.conv_to_mono(y32, self._p[keys[-1]].folded(judge)[0], judge.bias,
                                        judge.kernel_size[0], judge.padding[0], False))
        return results


class Discriminator(nn.Module):
    def __init__(self, num_D, ndf, n_layers, downsampling_factor, conditioning_channels=0):
        super().__init__()
        self.conditioning_channels = conditioning_channels
        self.model = nn.ModuleDict()
        for i in range(num_D):
            self.model[f"disc_{i}"] = NLayerDiscriminator(ndf, n_layers, downsampling_factor,
                                                          conditioning_channels)
        self.downsample = nn.AvgPool1d(4, stride=2, padding=1, count_include_pad=False)

    def forward(self, x, feat):
        features = []
        judgements = []
        for key, disc in self.model.items():
            z = disc(x, feat)
            features.append(z[:-1])
            judgements.append(z[-1])
            if torch.is_grad_enabled() and x.requires_grad:
                x = ag.AvgPool.apply(x, 4, 2, 1, False)
            else:
                x = ops.avg_pool1d(x, 4, 2, 1, count_include_pad=False)
        return features, judgements


def real_mel_gan_feature_loss(real_features, fake_features):
    """experiment/realmelgan.py:185-202"""
    acc = _Acc(real_features[0][0].device)
    wt = (1 / 3) * (4.0 / 5)
    for r_group, f_group in zip(real_features, fake_features):
        for r_f, f_f in zip(r_group, f_group):
            acc.add(L1, r_f, f_f, wt)
    return acc.value()


def mel_gan_gen_loss(real_features, fake_features, real_judgements, fake_judgements,
                     gan_loss=hinge_generator_loss, feature_loss_weight=10):
    """experiment/realmelgan.py:205-217"""
    acc = _Acc(fake_judgements[0].device)
    for _, f in zip(real_judgements, fake_judgements):
        gan_loss(f, _acc=acc)
    wt = (1 / 3) * (4.0 / 5) * float(feature_loss_weight)
    for r_group, f_group in zip(real_features, fake_features):
        for r_f, f_f in zip(r_group, f_group):
            acc.add(L1, r_f, f_f, wt)
    return acc.value()



def __getattr__(name):
    # `RealMelGanExperiment` lives with the other wirings (wirings.py imports this module)
    if name == "RealMelGanExperiment":
        from .wirings import RealMelGanExperiment
        return RealMelGanExperiment
    raise AttributeError(name)
