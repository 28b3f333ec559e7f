"""Tensor-level wrappers over the backward-pass / optimiser entry points of the C ABI
(include/msb200.h, "Training step").  torch supplies device memory and the stream only."""
import ctypes

import torch

from . import _lib, ops
from ._lib import MS_CONV, MS_CONVT, MS_F16, MS_BF16, check, ptr, stream_ptr

# 16-bit format of the backward GEMMs.  Default bf16: the range of fp32, so the GAN gradients
# (1e-2 at the discriminator's input down to 1e-9 deep in the generator within ONE backward pass --
# wider than fp16's normal range) need no loss scaling.  tcgen05 kind::f16 cannot mix fp16 and
# bf16 operands, so dgrad weights are packed in bf16 and the saved fp16 activations are converted
# for the wgrad.  MSB_GRAD_FMT=f16 selects fp16 operands under the dynamic loss scaler of
# train/train.py (measured: discriminator gradients 9x closer to the oracle, 3e-4; generator
# gradients unchanged at 4e-2 -- their error is LeakyReLU mask flips caused by the fp16 FORWARD,
# not backward rounding -- and NaN / underflow at either end of the generator step's range).
GRAD_FMT = MS_F16 if __import__('os').environ.get('MSB_GRAD_FMT') == 'f16' else MS_BF16
NEEDS_LOSS_SCALE = GRAD_FMT == MS_F16


TICKET_BYTES = 65536      # MS_TICKET_BYTES (include/msb200.h)
_ws = {}


def _workspace(nbytes, device):
    """Scratch buffer of the current (device, stream) for kernels that reduce across thread
    blocks.  Zero-filled at allocation: its first MS_TICKET_BYTES are the ticket counters of the
    deterministic two-stage reductions (include/msb200.h), which every launch leaves zero."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _ws.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.zeros(max(nbytes, 4 << 20), dtype=torch.uint8, device=device)
        _ws[key] = buf
    return buf


def convert16(x16, src_fmt, dst_fmt):
    if src_fmt == dst_fmt:
        return x16
    out = torch.empty_like(x16)
    check(_lib.lib().ms_blk16_convert(ptr(x16), ptr(out), x16.numel(), src_fmt, dst_fmt,
                                      stream_ptr()), "ms_blk16_convert")
    return out


def act_bwd(dy32, sign16=None, ya32=None, yb32=None, want_bias=True, fmt=GRAD_FMT, s2d=1,
            want32=False):
    """dz16 = to16(dy32 * LeakyReLU'(.)) (+ bias gradient).  Returns (dz16, dbias|None), or
    (dz16, dbias|None, dz32) with want32 (the masked gradient also in fp32)."""
    dy32 = dy32.contiguous()
    B, C8, L, _ = dy32.shape
    if s2d > 1:
        dz = torch.empty((B, s2d * C8, L // s2d, 8), dtype=torch.int16, device=dy32.device)
    else:
        dz = torch.empty((B, C8, L, 8), dtype=torch.int16, device=dy32.device)
    db = ws = None
    if want_bias:
        db = torch.empty(C8 * 8, dtype=torch.float32, device=dy32.device)
        ws = _workspace(_lib.lib().ms_blk_act_bwd_workspace_bytes(B, C8 * 8, L), dy32.device)
    dz32 = torch.empty_like(dy32) if want32 else None
    check(_lib.lib().ms_blk_act_bwd(ptr(dy32), ptr(sign16), ptr(ya32), ptr(yb32), ptr(dz),
                                    ptr(dz32), ptr(db), B, C8 * 8, L, fmt, s2d, ptr(ws),
                                    0 if ws is None else ws.numel(), stream_ptr()),
          "ms_blk_act_bwd")
    return (dz, db, dz32) if want32 else (dz, db)


def convt_dgrad_taps(ksize, stride, pad):
    """(ntaps, first_shift) of the stride-1 conv over the space-to-depth gradient that computes
    the input gradient of a ConvTranspose1d (ms_convt_dgrad_taps)"""
    first = ctypes.c_int(0)
    n = _lib.lib().ms_convt_dgrad_taps(ksize, stride, pad, ctypes.byref(first))
    if n < 1:
        raise _lib.MsbError("unsupported ConvTranspose1d geometry for the backward pass")
    if first.value != -(n - 1) // 2 or n % 2 == 0:
        raise _lib.MsbError("ConvTranspose1d backward: asymmetric tap range (padding must be "
                            "(kernel_size - stride) / 2)")
    return n, first.value


def weight_dgrad_view(w, kind, stride=1, pad=0):
    """reference-layout weight of the conv that computes the input gradient (see msb200.h)"""
    w = w.contiguous()
    if kind == MS_CONV:
        cout, cin, k = w.shape
        out = torch.empty((cin, cout, k), dtype=torch.float32, device=w.device)
    else:
        cin, cout, k = w.shape
        ntaps, _ = convt_dgrad_taps(k, stride, pad)
        out = torch.empty((cin, stride * cout, ntaps), dtype=torch.float32, device=w.device)
    check(_lib.lib().ms_weight_dgrad_view(ptr(w), ptr(out), kind, cout, cin, k, stride, pad,
                                          stream_ptr()), "ms_weight_dgrad_view")
    return out


def conv_dgrad(w, dz16, kind, dilation=1, pad=0, stride=1, res32=None, operand=GRAD_FMT):
    """Input gradient (BLK f32) of a dense conv / transposed conv from its 16-bit output
    gradient `dz16` (space-to-depth layout for MS_CONVT).  res32 is added in the epilogue."""
    B, C8, L, _ = dz16.shape
    wv = weight_dgrad_view(w, kind, stride, pad)
    if kind == MS_CONV:
        cout, cin, k = w.shape
        d = ops.conv_desc(MS_CONV, B, cout, cin, L, k, dilation, dilation * (k - 1) - pad,
                          operand=operand)
    else:
        cin, cout, k = w.shape
        ntaps, first = convt_dgrad_taps(k, stride, pad)
        d = ops.conv_desc(MS_CONV, B, stride * cout, cin, L, ntaps, 1, -first, operand=operand)
    _, dx32 = ops.conv_fwd(d, dz16, ops.pack_conv_weight(d, wv), None, res32=res32,
                           want16=False, want32=True)
    return dx32


def wgrad(a16, x16, shifts, mode, w_shape, fmt, stride=1, pad=0, fold=1, alpha=1.0):
    """Weight gradient (reference layout `w_shape`) on the tcgen05 time-reduction GEMM."""
    B, Cm8, La, _ = a16.shape
    _, Cn8, Lx, _ = x16.shape
    taps = len(shifts)
    sh = (ctypes.c_int * taps)(*shifts)
    L = _lib.lib()
    n = L.ms_wgrad_workspace_bytes(B, Cm8 * 8, Cn8 * 8, La, Lx, taps, sh)
    if n == 0:
        raise _lib.MsbError("unsupported weight-gradient geometry")
    # split-K partial tiles live behind the ticket counters of the shared scratch buffer
    ws = _workspace(n + TICKET_BYTES, a16.device)[TICKET_BYTES:]
    dw = torch.empty(w_shape, dtype=torch.float32, device=a16.device)
    cout = w_shape[1] if mode == MS_CONVT else 0
    check(L.ms_wgrad_fwd(ptr(a16), ptr(x16), B, Cm8 * 8, Cn8 * 8, La, Lx, taps, sh, fmt,
                         mode, stride, pad, cout, w_shape[2], fold, float(alpha), 0.0, ptr(dw),
                         ptr(ws), ws.numel(), stream_ptr()),
          "ms_wgrad_fwd")
    return dw


def conv_wgrad(dz16, x16, w_shape, dilation=1, pad=0, fmt_dz=GRAD_FMT, fmt_x=MS_F16, fold=1,
               alpha=1.0):
    """dW of a stride-1 Conv1d: dz16 (B,Cout/8,Lout,8), x16 (B,fold*Cin/8,Lin,8); fold = 2:
    x16 is a two-term split (ops.pack_ncl_split) of the layer input."""
    k = w_shape[2]
    return wgrad(dz16, convert16(x16, fmt_x, fmt_dz), [t * dilation - pad for t in range(k)],
                 MS_CONV, w_shape, fmt_dz, fold=fold, alpha=alpha)


def convt_wgrad(x16, dzs16, w_shape, stride, pad, fmt_dz=GRAD_FMT, fmt_x=MS_F16):
    """dW of a ConvTranspose1d (k = J*stride): x16 layer input, dzs16 space-to-depth dz."""
    ntaps, first = convt_dgrad_taps(w_shape[2], stride, pad)
    return wgrad(convert16(x16, fmt_x, fmt_dz), dzs16, list(range(first, first + ntaps)),
                 MS_CONVT, w_shape, fmt_dz, stride, pad)


def pack_ncl32(x):
    x = x.contiguous()
    B, C, L = x.shape
    y = torch.empty((B, C // 8, L, 8), dtype=torch.float32, device=x.device)
    check(_lib.lib().ms_pack_ncl_to_blk32(ptr(x), ptr(y), B, C, L, stream_ptr()),
          "ms_pack_ncl_to_blk32")
    return y


def conv1d_direct_bwd(dy, y, x, w, stride, pad, groups, leaky, need_dx=True, need_dw=True,
                      has_bias=True):
    """backward of ops.conv1d_direct (zero padding) -> (dx, dw, dbias)"""
    dy = dy.contiguous()
    B, cin, lin = x.shape
    cout, _, k = w.shape
    L = _lib.lib()
    dx = dw = db = None
    if need_dx:
        dx = torch.empty_like(x)
        check(L.ms_conv1d_direct_dgrad(ptr(dy), ptr(y), ptr(w.contiguous()), ptr(dx), B, cin, cout,
                                       lin, k, stride, pad, groups, int(leaky), stream_ptr()),
              "ms_conv1d_direct_dgrad")
    if need_dw:
        dw = torch.empty_like(w)
        db = torch.empty(cout, dtype=torch.float32, device=x.device) if has_bias else None
        ws = _workspace(L.ms_conv1d_direct_wgrad_workspace_bytes(B, cin, cout, lin, k, stride, pad,
                                                                 groups), x.device)
        check(L.ms_conv1d_direct_wgrad(ptr(dy), ptr(y), ptr(x.contiguous()), ptr(dw), ptr(db), B,
                                       cin, cout, lin, k, stride, pad, groups, int(leaky),
                                       ptr(ws), ws.numel(), stream_ptr()), "ms_conv1d_direct_wgrad")
    return dx, dw, db


def conv_to_mono_bwd(dy, y_tanh, x32, w, ksize, pad, need_dx=True, need_dw=True, has_bias=True):
    """backward of ops.conv_to_mono -> (dx32, dw, dbias)"""
    dy = dy.contiguous()
    B, C8, L, _ = x32.shape
    dzm = torch.empty((B, 1, L), dtype=torch.float32, device=x32.device)
    dx = torch.empty_like(x32) if need_dx else None
    dw = torch.empty_like(w) if need_dw else None
    db = torch.empty(1, dtype=torch.float32, device=x32.device) if (need_dw and has_bias) else None
    ws = _workspace(_lib.lib().ms_conv_to_mono_bwd_workspace_bytes(B, C8 * 8, L, ksize),
                    x32.device) if need_dw else None
    check(_lib.lib().ms_conv_to_mono_bwd(ptr(dy), ptr(y_tanh), ptr(x32), ptr(w.contiguous()),
                                         ptr(dzm), ptr(dx), ptr(dw), ptr(db), B, C8 * 8, L, ksize,
                                         pad, ptr(ws), 0 if ws is None else ws.numel(),
                                         stream_ptr()), "ms_conv_to_mono_bwd")
    return dx, dw, db


def avg_pool1d_bwd(dy, lin, ksize, stride, pad, count_include_pad=True):
    dy = dy.contiguous()
    B, C, _ = dy.shape
    dx = torch.empty((B, C, lin), dtype=torch.float32, device=dy.device)
    check(_lib.lib().ms_avg_pool1d_bwd(ptr(dy), ptr(dx), B * C, lin, ksize, stride, pad,
                                       int(count_include_pad), stream_ptr()), "ms_avg_pool1d_bwd")
    return dx


def reduce_bwd(mode, a, b, weight, grad_out=None, need_a=True, need_b=False):
    a = a.contiguous()
    b = b.contiguous() if b is not None else None
    da = torch.empty_like(a) if need_a else None
    db = torch.empty_like(b) if (need_b and b is not None) else None
    check(_lib.lib().ms_reduce_bwd(mode, ptr(a), ptr(b), a.numel(), float(weight), ptr(grad_out),
                                   ptr(da), ptr(db), stream_ptr()), "ms_reduce_bwd")
    return da, db


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, grad_scale=1.0):
    check(_lib.lib().ms_adam_step(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq),
                                  param.numel(), lr, beta1, beta2, eps, step, grad_scale,
                                  stream_ptr()), "ms_adam_step")


def adam_step_dev(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step_dev,
                  grad_scale=1.0, skip_flag=None):
    """Adam step with the step counter in device memory (int32 tensor): graph-capturable;
    skip_flag (int32 device tensor): non-zero = leave everything untouched"""
    check(_lib.lib().ms_adam_step_dev(ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq),
                                      param.numel(), lr, beta1, beta2, eps, ptr(step_dev),
                                      grad_scale, ptr(skip_flag), stream_ptr()), "ms_adam_step_dev")


def grad_unscale_check(grad, inv_scale_dev, flag_dev):
    """grad *= inv_scale (device scalar) in place; flag_dev <- any non-finite element"""
    check(_lib.lib().ms_grad_unscale_check(ptr(grad), grad.numel(), ptr(inv_scale_dev),
                                           ptr(flag_dev), stream_ptr()), "ms_grad_unscale_check")


def diag_sum_bwd(dy, channels, z_len, nphase, skew):
    dy = dy.contiguous()
    B, _, L = dy.shape
    dz = torch.empty((B, channels // 8, z_len, 8), dtype=torch.float32, device=dy.device)
    check(_lib.lib().ms_diag_sum_bwd(ptr(dy), ptr(dz), B, channels, z_len, L, nphase, skew,
                                     stream_ptr()), "ms_diag_sum_bwd")
    return dz


def expand_mono_bwd(de32, length, shift):
    de32 = de32.contiguous()
    B, _, Lx, _ = de32.shape
    dx = torch.empty((B, 1, length), dtype=torch.float32, device=de32.device)
    check(_lib.lib().ms_expand_mono_bwd(ptr(de32), ptr(dx), B, length, Lx, shift, stream_ptr()),
          "ms_expand_mono_bwd")
    return dx


def depth_to_space32(dys32, channels, stride, out_rows, length, rows_valid=None, row_offset=0):
    """gradient of ops.space_to_depth: (B, s*C/8, rows, 8) f32 -> (B, C/8, out_rows, 8) f32;
    space-to-depth row u is read from row u + row_offset, rows u >= rows_valid count as zero"""
    dys32 = dys32.contiguous()
    B, _, lx, _ = dys32.shape
    dx = torch.empty((B, channels // 8, out_rows, 8), dtype=torch.float32, device=dys32.device)
    check(_lib.lib().ms_depth_to_space_blk32(ptr(dys32), ptr(dx), B, channels, lx,
                                             (lx - row_offset) if rows_valid is None else rows_valid,
                                             row_offset, out_rows, length, stride, stream_ptr()),
          "ms_depth_to_space_blk32")
    return dx


def act_pad_bwd(dy32, sign16, length, pad, pad_mode):
    """gradient of ops.act_pad on the fp32 stream (LeakyReLU mask from the 16-bit image of x)"""
    dy32 = dy32.contiguous()
    B, C8, _, _ = dy32.shape
    dx = torch.empty((B, C8, length, 8), dtype=torch.float32, device=dy32.device)
    check(_lib.lib().ms_blk_act_pad_bwd(ptr(dy32), ptr(sign16), ptr(dx), B, C8 * 8, length, pad,
                                        pad_mode, stream_ptr()), "ms_blk_act_pad_bwd")
    return dx


def weight_norm_bwd(dw, v, g):
    """gradient of ops.weight_norm_fold -> (dv, dg with the shape of g)"""
    dw, v = dw.contiguous(), v.contiguous()
    rows = v.shape[0]
    dv = torch.empty_like(v)
    dg = torch.empty(rows, dtype=torch.float32, device=v.device)
    check(_lib.lib().ms_weight_norm_bwd(ptr(dw), ptr(v), ptr(g.contiguous()), ptr(dv), ptr(dg), rows,
                                        v.numel() // rows, stream_ptr()), "ms_weight_norm_bwd")
    return dv, dg.reshape(g.shape)
