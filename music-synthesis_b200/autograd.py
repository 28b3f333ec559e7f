"""torch.autograd.Function wrappers that put the sm_100a kernels on the TRAINING path.

The reference trains through plain autograd (featuresynth/train/train.py:26-74:
`loss.backward()` then `optim.step()`); here every differentiable block of the hot path is a
Function whose forward AND backward are calls into the C ABI (include/msb200.h):

  * activations travel between blocks as a pair (x32, x16): the BLK f32 tensor is the
    autograd-visible value (its gradient is BLK f32), the BLK f16 tensor is the tensor-core
    operand image written by the producing kernel's epilogue (marked non-differentiable);
  * backward of a dense conv = `ms_blk_act_bwd` (LeakyReLU' + bias gradient + bf16 operand)
    -> `ms_wgrad_fwd` (tcgen05 time-reduction GEMM) and `ms_conv_fwd` on the transposed /
    tap-reversed weights (input gradient, residual gradient added in its epilogue);
  * the discriminator's grouped convs, the single-channel convs, pooling and the losses use
    the direct fp32 kernels.

torch is the tape and the allocator; no torch arithmetic runs on the path except autograd's
own accumulation of gradients that fan in from several consumers.
"""
import ctypes

import torch
from torch.autograd import Function

from . import _lib, grad_ops, ops
from ._lib import MS_CONV, MS_CONVT, MS_F16, check, ptr, stream_ptr

GRAD_FMT = grad_ops.GRAD_FMT


class WeightCache:
    """Packed 16-bit images of one layer's weight: the forward image (fp16) and the
    input-gradient image (bf16, transposed / tap-reversed); rebuilt when the parameter changes
    in place (optimizer step, init, load_state_dict) or is moved."""

    def __init__(self):
        self._fwd = (None, None)
        self._bwd = (None, None)

    @staticmethod
    def _key(w, desc):
        # derived weights (weight-norm folds) carry the identity of their parameters
        ident = getattr(w, "_msb_key", None) or (w.data_ptr(), w._version)
        return (ident, desc.kind, desc.cin, desc.cout, desc.ksize, desc.stride, desc.operand)

    def fwd(self, desc, w):
        key = self._key(w, desc)
        if self._fwd[0] != key:
            self._fwd = (key, ops.pack_conv_weight(desc, w.detach()))
        return self._fwd[1]

    def fwd_dup(self, desc, w):
        """forward image of [W_hi, W_hi, W_lo] along the input-channel axis (split precision)"""
        key = self._key(w, desc)
        if self._fwd[0] != key:
            self._fwd = (key, ops.pack_conv_weight(desc, ops.weight_split(w.detach(), scale=SPLIT_SW)))
        return self._fwd[1]

    def fwd_wsplit(self, desc, w, kind):
        """forward image of [W_hi, W_lo] along the input-channel axis (weight-split forward:
        the conv sees the activation operand twice, `desc.cin` = 2 x the layer's)"""
        key = self._key(w, desc)
        if self._fwd[0] != key:
            ws = ops.weight_split(w.detach(), scale=ops.W_SPLIT_SCALE, terms=2, kind=kind)
            self._fwd = (key, ops.pack_conv_weight(desc, ws))
        return self._fwd[1]

    def fwd_split3(self, desc, w, kind):
        """forward image of [W_hi, W_hi, W_lo] for a Conv1d or ConvTranspose1d weight"""
        key = self._key(w, desc)
        if self._fwd[0] != key:
            ws = ops.weight_split(w.detach(), scale=SPLIT_SW, terms=3, kind=kind)
            self._fwd = (key, ops.pack_conv_weight(desc, ws))
        return self._fwd[1]

    def dgrad(self, desc, w, kind, stride, pad):
        key = self._key(w, desc)
        if self._bwd[0] != key:
            wv = grad_ops.weight_dgrad_view(w.detach(), kind, stride, pad)
            self._bwd = (key, ops.pack_conv_weight(desc, wv))
        return self._bwd[1]


def _dgrad_conv(cache, w, dz16, kind, dilation, pad, stride, res32=None, extra_pad=0):
    """input gradient (BLK f32) of a dense conv / transposed conv; `res32` added in the epilogue.
    extra_pad: the forward conv dropped `extra_pad` output rows at the end (crop = a smaller right
    padding); its input gradient needs that much more padding -- the result then starts
    `extra_pad` rows early (row u of the true gradient is row u + extra_pad)."""
    B, _, L, _ = dz16.shape
    if kind == MS_CONV:
        cout, cin, k = w.shape
        d = ops.conv_desc(MS_CONV, B, cout, cin, L, k, dilation,
                          dilation * (k - 1) - pad + extra_pad, operand=GRAD_FMT)
    else:
        cin, cout, k = w.shape
        ntaps, first = grad_ops.convt_dgrad_taps(k, stride, pad)
        d = ops.conv_desc(MS_CONV, B, stride * cout, cin, L, ntaps, 1, -first, operand=GRAD_FMT)
    _, dx32 = ops.conv_fwd(d, dz16, cache.dgrad(d, w, kind, stride, pad), None, res32=res32,
                           want16=False, want32=True)
    return dx32


class ConvBlk(Function):
    """y = act(conv(x) + b) on channel-blocked tensors (stride-1 dilated Conv1d or k = 2*stride
    ConvTranspose1d).  Returns (y32, y16)."""

    @staticmethod
    def forward(ctx, x32, x16, w, b, cache, kind, dilation, pad, stride, leaky, res32=None,
                wsplit=False):
        """res32 (optional, BLK f32): added after the activation, y = act(conv + b) + res32.
        wsplit: 1 = weight-split forward ([x, x] * [W_hi, W_lo]: no weight rounding, 2x the MMA
        work), 2 = full split precision (3x); the backward is the same either way."""
        B, _, L, _ = x16.shape
        if kind == MS_CONV:
            cout, cin, k = w.shape
        else:
            cin, cout, k = w.shape
        if res32 is not None and leaky:
            raise _lib.MsbError("ConvBlk: residual input only with a linear epilogue")
        if wsplit == 2:
            # full split precision: [x_hi, x_lo, x_hi] * [W_hi, W_hi, W_lo] = x*W to ~2^-22
            if x32 is None:
                raise _lib.MsbError("ConvBlk: the full-split forward needs the fp32 input")
            d = ops.conv_desc(kind, B, 3 * cin, cout, L, k, dilation, pad, stride, leaky=leaky,
                              alpha=1.0 / (SPLIT_SX * SPLIT_SW))
            xs = ops.blk32_split(x32, operand=MS_F16, terms=3, scale=SPLIT_SX)
            y16, y32 = ops.conv_fwd(d, xs, cache.fwd_split3(d, w, kind), b, res32=res32,
                                    want16=True, want32=True)
        elif wsplit:
            d = ops.conv_desc(kind, B, 2 * cin, cout, L, k, dilation, pad, stride, leaky=leaky,
                              alpha=1.0 / ops.W_SPLIT_SCALE, x_repeat=2)
            y16, y32 = ops.conv_fwd(d, x16, cache.fwd_wsplit(d, w, kind), b, res32=res32,
                                    want16=True, want32=True)
        else:
            d = ops.conv_desc(kind, B, cin, cout, L, k, dilation, pad, stride, leaky=leaky)
            y16, y32 = ops.conv_fwd(d, x16, cache.fwd(d, w), b, res32=res32, want16=True,
                                    want32=True)
        ctx.save_for_backward(x16, w, y16)
        ctx.cfg = (cache, kind, dilation, pad, stride, leaky, b is not None)
        ctx.wkey = getattr(w, "_msb_key", None)      # identity of a derived (weight-normed) weight
        ctx.has_res = res32 is not None
        ctx.mark_non_differentiable(y16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _unused):
        x16, w, y16 = ctx.saved_tensors
        cache, kind, dilation, pad, stride, leaky, has_bias = ctx.cfg
        if ctx.wkey is not None:
            w._msb_key = ctx.wkey
        dres = dy32 if (ctx.has_res and ctx.needs_input_grad[10]) else None
        need_x, need_w = ctx.needs_input_grad[0], ctx.needs_input_grad[2]
        s2d = stride if kind == MS_CONVT else 1
        dz16, db = grad_ops.act_bwd(dy32, sign16=y16 if leaky else None,
                                    want_bias=has_bias and ctx.needs_input_grad[3], s2d=s2d)
        dw = dx32 = None
        if need_w:
            if kind == MS_CONV:
                dw = grad_ops.conv_wgrad(dz16, x16, tuple(w.shape), dilation, pad)
            else:
                dw = grad_ops.convt_wgrad(x16, dz16, tuple(w.shape), stride, pad)
        if need_x:
            dx32 = _dgrad_conv(cache, w, dz16, kind, dilation, pad, stride)
        return dx32, None, dw, db, None, None, None, None, None, None, dres, None


def conv_blk(x32, x16, w, b, cache, kind, dilation, pad, stride, leaky, res32=None, wsplit=False):
    """ConvBlk.apply with every argument spelled out (autograd wants one gradient per argument)"""
    return ConvBlk.apply(x32, x16, w, b, cache, kind, dilation, pad, stride, leaky, res32, wsplit)


class ResidualAtomBlk(Function):
    """x + leaky(conv_k3_pad1(leaky(conv_k3_dil_d(x)))), featuresynth/util/modules.py:384-388.
    Returns (y32, y16)."""

    @staticmethod
    def forward(ctx, x32, x16, w1, b1, w2, b2, cache1, cache2, dilation):
        B, C8, L, _ = x16.shape
        C = C8 * 8
        d1 = ops.conv_desc(MS_CONV, B, C, C, L, 3, dilation, dilation, leaky=True)
        d2 = ops.conv_desc(MS_CONV, B, C, C, L, 3, 1, 1, leaky=True)
        h16, _ = ops.conv_fwd(d1, x16, cache1.fwd(d1, w1), b1)
        y16, y32 = ops.conv_fwd(d2, h16, cache2.fwd(d2, w2), b2, res32=x32, want16=True,
                                want32=True)
        ctx.save_for_backward(x32, x16, h16, y32, w1, w2)
        ctx.cfg = (cache1, cache2, dilation)
        ctx.wkeys = (getattr(w1, "_msb_key", None), getattr(w2, "_msb_key", None))
        ctx.mark_non_differentiable(y16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _unused):
        x32, x16, h16, y32, w1, w2 = ctx.saved_tensors
        cache1, cache2, dilation = ctx.cfg
        for w_, k_ in zip((w1, w2), ctx.wkeys):
            if k_ is not None:
                w_._msb_key = k_
        need_w = ctx.needs_input_grad[2]
        dy32 = dy32.contiguous()
        # outer LeakyReLU: its output is y - x (sign of the fp32 difference)
        dz2, db2 = grad_ops.act_bwd(dy32, ya32=y32, yb32=x32, want_bias=need_w)
        dw2 = grad_ops.conv_wgrad(dz2, h16, tuple(w2.shape), 1, 1) if need_w else None
        dh32 = _dgrad_conv(cache2, w2, dz2, MS_CONV, 1, 1, 1)
        dz1, db1 = grad_ops.act_bwd(dh32, sign16=h16, want_bias=need_w)
        dw1 = grad_ops.conv_wgrad(dz1, x16, tuple(w1.shape), dilation, dilation) if need_w else None
        dx32 = None
        if ctx.needs_input_grad[0]:
            dx32 = _dgrad_conv(cache1, w1, dz1, MS_CONV, dilation, dilation, 1, res32=dy32)
        return dx32, None, dw1, db1, dw2, db2, None, None, None


class DilatedLayerBlk(Function):
    """one layer of a residual DilatedStack (util/modules.py:120-137 as configured by
    ChannelGenerator): y = LeakyReLU(conv_k3_dil_d(x) + x), bias-free.  -> (y32, y16)"""

    @staticmethod
    def forward(ctx, x32, x16, w, cache, dilation):
        B, C8, L, _ = x16.shape
        C = C8 * 8
        d = ops.conv_desc(MS_CONV, B, C, C, L, 3, dilation, dilation, leaky=2)
        y16, y32 = ops.conv_fwd(d, x16, cache.fwd(d, w), None, res32=x32, want16=True, want32=True)
        ctx.save_for_backward(x16, y16, w)
        ctx.cfg = (cache, dilation)
        ctx.mark_non_differentiable(y16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _unused):
        x16, y16, w = ctx.saved_tensors
        cache, dilation = ctx.cfg
        # dz = dy * LeakyReLU'(y) is the gradient of BOTH the conv output and the skip input
        # (the skip path takes it in fp32, the GEMMs as a 16-bit operand)
        dz16, _, dz32 = grad_ops.act_bwd(dy32, sign16=y16, want_bias=False, want32=True)
        dw = grad_ops.conv_wgrad(dz16, x16, tuple(w.shape), dilation, dilation) \
            if ctx.needs_input_grad[2] else None
        dx32 = None
        if ctx.needs_input_grad[0]:
            dx32 = _dgrad_conv(cache, w, dz16, MS_CONV, dilation, dilation, 1, res32=dz32)
        return dx32, None, dw, None, None


class WeightNorm(Function):
    """torch.nn.utils.weight_norm (dim 0): W = g * v / ||v|| (experiment/realmelgan.py:24-29)"""

    @staticmethod
    def forward(ctx, v, g):
        ctx.save_for_backward(v, g)
        return ops.weight_norm_fold(v, g)

    @staticmethod
    def backward(ctx, dw):
        v, g = ctx.saved_tensors
        return grad_ops.weight_norm_bwd(dw, v, g)


class ActPadBlk(Function):
    """(x32, x16) -> (a32, a16) = LeakyReLU / padding (zero or reflection) of both images: the
    nn.LeakyReLU(0.2) + nn.ReflectionPad1d in front of the convs of the official MelGAN blocks
    (experiment/realmelgan.py:35-37, 60-61, 80-81)"""

    @staticmethod
    def forward(ctx, x32, x16, pad, pad_mode, leaky):
        a16 = ops.act_pad(x16, pad, pad_mode, leaky)
        a32 = ops.act_pad(x32, pad, pad_mode, leaky)
        ctx.save_for_backward(x16)
        ctx.cfg = (x16.shape[2], pad, pad_mode, leaky)
        ctx.mark_non_differentiable(a16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return a32, a16

    @staticmethod
    def backward(ctx, da32, _unused):
        (x16,) = ctx.saved_tensors
        length, pad, pad_mode, leaky = ctx.cfg
        return grad_ops.act_pad_bwd(da32, x16 if leaky else None, length, pad, pad_mode), None, None, None, None


class MonoConv(Function):
    """(B,C,L) BLK f32 -> (B,1,L): single-output-channel conv (+ tanh): the generator's last
    layer (generator/full.py:43-44) and the discriminator's judge (discriminator/full.py:22)."""

    @staticmethod
    def forward(ctx, x32, w, b, ksize, pad, tanh_out):
        y = ops.conv_to_mono(x32, w, b, ksize, pad, tanh_out)
        ctx.save_for_backward(x32, w, y)
        ctx.cfg = (ksize, pad, tanh_out, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x32, w, y = ctx.saved_tensors
        ksize, pad, tanh_out, has_bias = ctx.cfg
        dx32, dw, db = grad_ops.conv_to_mono_bwd(dy, y if tanh_out else None, x32, w, ksize, pad,
                                                 need_dx=ctx.needs_input_grad[0],
                                                 need_dw=ctx.needs_input_grad[1],
                                                 has_bias=has_bias)
        return dx32, dw, (db if ctx.needs_input_grad[2] else None), None, None, None


class PackBlk32(Function):
    """NCL f32 -> BLK f32 (entry of a stand-alone block called with (B,C,L) tensors)."""

    @staticmethod
    def forward(ctx, x):
        return grad_ops.pack_ncl32(x)

    @staticmethod
    def backward(ctx, dy32):
        return ops.unpack_blk32(dy32.contiguous())


class UnpackBlk32(Function):
    """BLK f32 -> NCL f32 (feature maps handed to the losses)."""

    @staticmethod
    def forward(ctx, x32):
        return ops.unpack_blk32(x32)

    @staticmethod
    def backward(ctx, dy):
        return grad_ops.pack_ncl32(dy)


class ReflectPadNCL(Function):
    """nn.ReflectionPad1d on a (B, C, L) fp32 tensor (experiment/realmelgan.py:98-102): gather
    forward, fixed-order gather backward."""

    @staticmethod
    def forward(ctx, x, pad):
        x = x.contiguous()
        B, C, L = x.shape
        y = torch.empty((B, C, L + 2 * pad), dtype=torch.float32, device=x.device)
        check(_lib.lib().ms_reflect_pad_ncl(ptr(x), ptr(y), B * C, L, pad, stream_ptr()),
                  "ms_reflect_pad_ncl")
        ctx.cfg = (B, C, L, pad)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, L, pad = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty((B, C, L), dtype=torch.float32, device=dy.device)
        check(_lib.lib().ms_reflect_pad_ncl_bwd(ptr(dy), ptr(dx), B * C, L, pad, stream_ptr()),
              "ms_reflect_pad_ncl_bwd")
        return dx, None


class DirectConv(Function):
    """grouped / strided fp32 conv1d (+ LeakyReLU), NCL in / out (discriminator/full.py:13-18)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, pad, groups, leaky):
        y = ops.conv1d_direct(x, w, b, stride, pad, groups, leaky)
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (stride, pad, groups, leaky, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        stride, pad, groups, leaky, has_bias = ctx.cfg
        dx, dw, db = grad_ops.conv1d_direct_bwd(dy, y if leaky else None, x, w, stride, pad, groups,
                                                leaky, need_dx=ctx.needs_input_grad[0],
                                                need_dw=ctx.needs_input_grad[1], has_bias=has_bias)
        return dx, dw, (db if ctx.needs_input_grad[2] else None), None, None, None, None


# power-of-two operand scales of the split-precision layer: keep the lo terms of activations
# (|x| >~ 4e-3) and weights (|w| >~ 1e-3) normal in fp16, overflow only beyond |x| ~ 1e3 / |w| ~ 250
SPLIT_SX, SPLIT_SW = 64.0, 256.0


def dense_split_fwd(x, w, b, cache, pad, leaky, want16=True):
    """LeakyReLU(dense Conv1d) in split precision: [x_hi, x_lo, x_hi] * [W_hi, W_hi, W_lo] over 3C
    fp16 channels = x*W to ~2^-22.  The discriminator's top-layer activations are bias-dominated
    and the real / fake gradients of a GAN step cancel to first order: which LeakyReLU masks
    differ between the two batches -- and with them the whole gradient of the top layers --
    hinges on differences of ~1e-4 relative, below a single fp16 rounding of x or W."""
    B, C, L = x.shape
    cout, _, k = w.shape
    x16 = ops.pack_ncl_split(x, terms=3, scale=SPLIT_SX)
    d = ops.conv_desc(MS_CONV, B, 3 * C, cout, L, k, 1, pad, leaky=leaky,
                      alpha=1.0 / (SPLIT_SX * SPLIT_SW))
    return ops.conv_fwd(d, x16, cache.fwd_dup(d, w), b, want16=want16, want32=True)


class DenseConvNCL(Function):
    """NCL f32 in -> BLK f32 out: LeakyReLU(dense Conv1d) on the tcgen05 kernel
    (discriminator/full.py:19, the 1024 -> 1024 k5 layer), split-precision input."""

    @staticmethod
    def forward(ctx, x, w, b, cache, pad, leaky):
        y16, y32 = dense_split_fwd(x, w, b, cache, pad, leaky)
        ctx.save_for_backward(x, w, y16)
        ctx.cfg = (cache, pad, leaky, b is not None)
        return y32

    @staticmethod
    def backward(ctx, dy32):
        x, w, y16 = ctx.saved_tensors
        cache, pad, leaky, has_bias = ctx.cfg
        need_w = ctx.needs_input_grad[1]
        dz16, db = grad_ops.act_bwd(dy32, sign16=y16 if leaky else None,
                                    want_bias=has_bias and ctx.needs_input_grad[2])
        dw = None
        if need_w:
            # (hi, lo) of the layer input in the gradient format, scaled out of the subnormals
            xs = ops.pack_ncl_split(x, operand=GRAD_FMT, scale=SPLIT_SX)
            dw = grad_ops.conv_wgrad(dz16, xs, tuple(w.shape), 1, pad, fmt_x=GRAD_FMT, fold=2,
                                     alpha=1.0 / SPLIT_SX)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.unpack_blk32(_dgrad_conv(cache, w, dz16, MS_CONV, 1, pad, 1))
        return dx, dw, db, None, None, None


class BankSynthesis(Function):
    """zounds FilterBank.transposed_convolve on the last generator activation
    (generator/multiscale.py:90-91): (h32, h16) BLK (B,n/8,L,8) -> (B,1,L).  The bank is fixed:
    backward = ms_diag_sum_bwd -> bf16 operand -> input-gradient conv with the bank weights."""

    @staticmethod
    def forward(ctx, h32, h16, bank, wsplit=False):
        L = h16.shape[2]
        ctx.bank = bank
        ctx.shape = tuple(h16.shape)
        return bank.transposed_convolve_blocked(h16, L, wsplit, h32 if wsplit == 2 else None)

    @staticmethod
    def backward(ctx, dy):
        bank = ctx.bank
        B, _, L, _ = ctx.shape
        dz32 = grad_ops.diag_sum_bwd(dy, bank.syn_ch, bank.synthesis_rows(L), bank.syn_nph, 1)
        dz16, _ = grad_ops.act_bwd(dz32, want_bias=False)
        w = bank.synthesis_weight(dy.device)
        cache = bank.__dict__.setdefault("_synth_dgrad_cache", WeightCache())
        return _dgrad_conv(cache, w, dz16, MS_CONV, bank.syn_nph, bank.syn_pad, 1), None, None, None


class BankAnalysis(Function):
    """zounds FilterBank.convolve (discriminator/multiscale.py:112) with a differentiable input:
    x (B,1,L) -> (a32, a16) BLK (B,n/8,L+1,8).  Backward = input-gradient conv with the bank
    weights over the 16-wide expansion, then ms_expand_mono_bwd."""

    @staticmethod
    def forward(ctx, x, bank):
        a16, a32 = bank._analysis(x, True, True)
        ctx.bank = bank
        ctx.L = x.shape[-1]
        ctx.mark_non_differentiable(a16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return a32, a16

    @staticmethod
    def backward(ctx, da32, _unused):
        bank = ctx.bank
        dz16, _ = grad_ops.act_bwd(da32, want_bias=False)
        w = bank.analysis_weight(da32.device)
        cache = bank.__dict__.setdefault("_analysis_dgrad_cache", WeightCache())
        de32 = _dgrad_conv(cache, w, dz16, MS_CONV, 16, 0, 1)
        return grad_ops.expand_mono_bwd(de32, ctx.L, bank.pad), None


class StridedCache:
    """weights of a stride-s conv rewritten as a stride-1 conv over the space-to-depth input
    (ops.strided_conv_weight) + their packed forward / input-gradient images"""

    def __init__(self):
        self.key = None
        self.w1 = self.taps = self.pad = None
        self.cache = WeightCache()

    def get(self, w, stride):
        key = (w.data_ptr(), w._version, stride)
        if key != self.key:
            self.w1, self.taps, self.pad = ops.strided_conv_weight(w.detach(), stride)
            self.key = key
        return self.w1, self.taps, self.pad


class StridedConvBlk(Function):
    """LeakyReLU(Conv1d(C, Cout, k, stride s, padding k//2)) on channel-blocked tensors
    (discriminator/multiscale.py:83-88, 103-107): space-to-depth + stride-1 tcgen05 conv.
    `length` = valid rows of the input (its tensor may carry one extra row).  -> (y32, y16)"""

    @staticmethod
    def forward(ctx, h32, h16, w, b, sc, stride, length):
        B = h16.shape[0]
        cout, cin, k = w.shape
        xs = ops.space_to_depth(h16, stride, length)
        lx = xs.shape[2]
        w1, taps, pad = sc.get(w, stride)
        crop = lx + 2 * pad - (taps - 1) - lx
        d = ops.conv_desc(MS_CONV, B, stride * cin, cout, lx, taps, 1, pad, leaky=True, crop=crop)
        y16, y32 = ops.conv_fwd(d, xs, sc.cache.fwd(d, w1), b, want16=True, want32=True)
        ctx.save_for_backward(xs, w, y16)
        ctx.cfg = (sc, stride, length, h16.shape[2], crop, b is not None)
        ctx.mark_non_differentiable(y16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return y32, y16

    @staticmethod
    def backward(ctx, dy32, _unused):
        xs, w, y16 = ctx.saved_tensors
        sc, stride, length, in_rows, crop, has_bias = ctx.cfg
        cout, cin, k = w.shape
        w1, taps, pad = sc.get(w, stride)
        dz16, db = grad_ops.act_bwd(dy32, sign16=y16, want_bias=has_bias and ctx.needs_input_grad[3])
        dw = None
        if ctx.needs_input_grad[2]:
            dw1 = grad_ops.conv_wgrad(dz16, xs, tuple(w1.shape), 1, pad)
            # back to the (Cout, C, k) layout: tap kk = (j, i) with kk - k//2 = s*j + i
            dw = ops.strided_conv_weight_grad(dw1, tuple(w.shape), stride)
        dx32 = None
        if ctx.needs_input_grad[0]:
            dxs = _dgrad_conv(sc.cache, w1, dz16, MS_CONV, 1, pad, 1, extra_pad=crop)
            lx = xs.shape[2]                       # dxs: lx + crop rows, true row u at u + crop
            dx32 = grad_ops.depth_to_space32(dxs, cin, stride, in_rows, length, rows_valid=lx,
                                             row_offset=crop)
        return dx32, None, dw, db, None, None, None


class NoiseMix(Function):
    """y = add + sum_c a[:, c] * n[c] (generator/filterbank.py:83-91): a32 BLK f32 (B,C/8,L,8),
    n32 BLK f32 (1,C/8,L,8) fixed noise (no gradient), add (B,1,L)"""

    @staticmethod
    def forward(ctx, a32, n32, add):
        B, C8, L, _ = a32.shape
        y = torch.empty((B, 1, L), dtype=torch.float32, device=a32.device)
        check(_lib.lib().ms_noise_mix_fwd(ptr(a32.contiguous()), ptr(n32.contiguous()),
                                          ptr(add.contiguous()), ptr(y), B, C8 * 8, L,
                                          stream_ptr()), "ms_noise_mix_fwd")
        ctx.save_for_backward(n32)
        ctx.shape = (B, C8, L)
        return y

    @staticmethod
    def backward(ctx, dy):
        (n32,) = ctx.saved_tensors
        B, C8, L = ctx.shape
        dy = dy.contiguous()
        da = None
        if ctx.needs_input_grad[0]:
            da = torch.empty((B, C8, L, 8), dtype=torch.float32, device=dy.device)
            check(_lib.lib().ms_noise_mix_bwd(ptr(dy), ptr(n32), ptr(da), B, C8 * 8, L,
                                              stream_ptr()), "ms_noise_mix_bwd")
        return da, None, (dy if ctx.needs_input_grad[2] else None)


class ReluAvgPool(Function):
    """front end of LowResSpectrogramDiscriminator (util/modules.py:315-325): relu, then the mean
    over (channel_window x time_window) windows of a BLK f32 tensor -> (p32, p16)"""

    @staticmethod
    def forward(ctx, a32, cw, tw):
        B, C8, L, _ = a32.shape
        shape = (B, C8 // cw, L // tw, 8)
        p32 = torch.empty(shape, dtype=torch.float32, device=a32.device)
        p16 = torch.empty(shape, dtype=torch.int16, device=a32.device)
        check(_lib.lib().ms_relu_avgpool2d_fwd(ptr(a32.contiguous()), ptr(p16), ptr(p32), B, C8 * 8,
                                               L, cw, tw, MS_F16, stream_ptr()),
              "ms_relu_avgpool2d_fwd")
        ctx.save_for_backward(a32)
        ctx.cfg = (cw, tw)
        ctx.mark_non_differentiable(p16)
        ctx.set_materialize_grads(False)   # no zero-filled "gradient" for the 16-bit image
        return p32, p16

    @staticmethod
    def backward(ctx, dp32, _unused):
        (a32,) = ctx.saved_tensors
        cw, tw = ctx.cfg
        B, C8, L, _ = a32.shape
        dx = torch.empty_like(a32)
        check(_lib.lib().ms_relu_avgpool2d_bwd(ptr(dp32.contiguous()), ptr(a32), ptr(dx), B, C8 * 8,
                                               L, cw, tw, stream_ptr()), "ms_relu_avgpool2d_bwd")
        return dx, None, None


class AvgPool(Function):
    @staticmethod
    def forward(ctx, x, ksize, stride, pad, include_pad):
        ctx.cfg = (x.shape[-1], ksize, stride, pad, include_pad)
        return ops.avg_pool1d(x, ksize, stride, pad, include_pad)

    @staticmethod
    def backward(ctx, dy):
        lin, ksize, stride, pad, include_pad = ctx.cfg
        return grad_ops.avg_pool1d_bwd(dy, lin, ksize, stride, pad, include_pad), None, None, None, None


class WeightedLoss(Function):
    """sum_i weight_i * L_{mode_i}(a_i, b_i) as ONE device scalar (featuresynth/loss/loss.py);
    spec: list of (mode, index of a, index of b or -1, weight)."""

    @staticmethod
    def forward(ctx, spec, *tensors):
        dev = tensors[0].device
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        ws = torch.empty(_lib.lib().ms_reduce_workspace_bytes(), dtype=torch.uint8, device=dev)
        keep = [t.contiguous() for t in tensors]
        for mode, ia, ib, weight in spec:
            a = keep[ia]
            b = keep[ib] if ib >= 0 else None
            check(_lib.lib().ms_reduce_fwd(mode, ptr(a), ptr(b), a.numel(), float(weight), ptr(out),
                                           1, ptr(ws), stream_ptr()), "ms_reduce_fwd")
        ctx.spec = spec
        ctx.save_for_backward(*keep)
        return out.reshape(())

    @staticmethod
    def backward(ctx, g):
        tensors = ctx.saved_tensors
        g = g.contiguous().reshape(1)
        grads = [None] * len(tensors)

        def acc(i, t):
            grads[i] = t if grads[i] is None else grads[i] + t

        for mode, ia, ib, weight in ctx.spec:
            need_a = ctx.needs_input_grad[1 + ia]
            need_b = ib >= 0 and ctx.needs_input_grad[1 + ib]
            if not (need_a or need_b):
                continue
            da, db = grad_ops.reduce_bwd(mode, tensors[ia], tensors[ib] if ib >= 0 else None, weight,
                                         g, need_a, need_b)
            if need_a:
                acc(ia, da)
            if need_b:
                acc(ib, db)
        return (None, *grads)


def needs_grad(module, *inputs):
    """True when the call must be recorded on the autograd tape."""
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in inputs) or \
        any(p.requires_grad for p in module.parameters())
