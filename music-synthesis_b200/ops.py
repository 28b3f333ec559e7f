"""Tensor-level wrappers over the C ABI (device pointers in, device pointers out).

torch is used for device memory and streams only; all arithmetic happens in
csrc/libmsb200.so.  16-bit channel-blocked tensors are carried as torch.int16 with
shape (B, C/8, L, 8); the fp32 residual stream as float32 of the same shape.
"""
import ctypes

import torch

from . import _lib
from ._lib import ConvDesc, MS_CONV, MS_CONVT, MS_F16, MS_BF16, check, ptr, stream_ptr

OPERANDS = {"f16": MS_F16, "fp16": MS_F16, "bf16": MS_BF16}


def conv_desc(kind, batch, cin, cout, lin, ksize, dilation=1, pad=0, stride=1, leaky=False,
              operand=MS_F16, alpha=1.0, crop=0, x_repeat=1):
    return ConvDesc(kind, batch, cin, cout, lin, ksize, dilation, pad, stride, int(leaky),
                    operand, alpha, crop, x_repeat)


def conv_out_len(desc):
    n = _lib.lib().ms_conv_out_len(ctypes.byref(desc))
    if n < 0:
        raise _lib.MsbError("unsupported convolution descriptor")
    return n


def pack_ncl(x, pad=0, pad_mode=0, operand=MS_F16):
    """(B,C,L) f32 -> (B,C/8,L+2*pad,8) 16-bit, zero (0) or reflection (1) padded."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, C, L = x.shape
    y = torch.empty((B, C // 8, L + 2 * pad, 8), dtype=torch.int16, device=x.device)
    check(_lib.lib().ms_pack_ncl_to_blk16(ptr(x), ptr(y), B, C, L, pad, pad_mode, operand,
                                          stream_ptr()), "ms_pack_ncl_to_blk16")
    return y


def pack_ncl_split(x, operand=MS_F16, terms=2, scale=1.0):
    """(B,C,L) f32 -> (B,terms*C/8,L,8) 16-bit split: channels [0,C) hi, [C,2C) lo, [2C,3C) hi."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, C, L = x.shape
    y = torch.empty((B, terms * C // 8, L, 8), dtype=torch.int16, device=x.device)
    check(_lib.lib().ms_pack_ncl_split_blk16(ptr(x), ptr(y), B, C, L, operand, terms,
                                             float(scale), stream_ptr()), "ms_pack_ncl_split_blk16")
    return y


def weight_split(w, operand=MS_F16, scale=1.0, terms=3, kind=MS_CONV):
    """Split-precision weight along the INPUT-channel axis (see ms_weight_split): a Conv1d
    weight (Cout,Cin,K) -> (Cout,terms*Cin,K), a ConvTranspose1d weight (Cin,Cout,K) ->
    (terms*Cin,Cout,K); terms 3 = [w, w, w - to16(w)], terms 2 = [w, w - to16(w)]."""
    w = w.contiguous()
    if kind == MS_CONV:
        cout, cin, k = w.shape
        out = torch.empty((cout, terms * cin, k), dtype=torch.float32, device=w.device)
        args = (cout, cin, k)
    else:
        cin, cout, k = w.shape
        out = torch.empty((terms * cin, cout, k), dtype=torch.float32, device=w.device)
        args = (1, cin, cout * k)
    check(_lib.lib().ms_weight_split(ptr(w), ptr(out), *args, operand, float(scale), terms,
                                     stream_ptr()),
          "ms_weight_split")
    return out


#: power-of-two weight scale of the weight-split forward: N(0, 0.02)-sized weights give lo terms
#: of ~1e-5, inside the fp16 subnormals; x256 moves them to ~2.5e-3 (overflow beyond |w| ~ 250)
W_SPLIT_SCALE = 256.0


_relaxed = [False]


class relaxed_forward:
    """Context of DiscriminatorTrainer's generator call: that fake batch is not returned to the
    caller, it only feeds the discriminator's loss (asserted to 2e-3), so modules may skip the
    precision extras of their inference path (the weight-split forward of the filter-bank
    generators: 2x the MMA work)."""

    def __enter__(self):
        self.prev = _relaxed[0]
        _relaxed[0] = True

    def __exit__(self, *exc):
        _relaxed[0] = self.prev
        return False


def relaxed():
    return _relaxed[0]


def dup_channels(x16, times=2):
    """BLK 16-bit (B,C/8,L,8) -> (B,times*C/8,L,8): the operand of a weight-split conv (data
    movement only)"""
    return torch.cat([x16] * times, dim=1)


def blk32_split(x32, pad=0, pad_mode=0, leaky=False, operand=MS_BF16, terms=3, scale=1.0):
    """BLK f32 (B,C/8,L,8) -> split BLK 16-bit (B,terms*C/8,L+2*pad,8) = [hi, lo(, hi)] of
    act(pad(x)) (see ms_blk32_split_blk16)"""
    B, C8, L, _ = x32.shape
    y = torch.empty((B, terms * C8, L + 2 * pad, 8), dtype=torch.int16, device=x32.device)
    check(_lib.lib().ms_blk32_split_blk16(ptr(x32.contiguous()), ptr(y), B, C8 * 8, L, pad,
                                          pad_mode, int(leaky), operand, terms, float(scale),
                                          stream_ptr()), "ms_blk32_split_blk16")
    return y


def unpack_blk32(x32):
    B, C8, L, _ = x32.shape
    y = torch.empty((B, C8 * 8, L), dtype=torch.float32, device=x32.device)
    check(_lib.lib().ms_unpack_blk32_to_ncl(ptr(x32), ptr(y), B, C8 * 8, L, stream_ptr()),
          "ms_unpack_blk32_to_ncl")
    return y


def unpack_blk16(x16, operand=MS_F16):
    B, C8, L, _ = x16.shape
    y = torch.empty((B, C8 * 8, L), dtype=torch.float32, device=x16.device)
    check(_lib.lib().ms_unpack_blk16_to_ncl(ptr(x16), ptr(y), B, C8 * 8, L, operand,
                                            stream_ptr()), "ms_unpack_blk16_to_ncl")
    return y


def pack_conv_weight(desc, w):
    """fp32 reference-layout weight -> packed 16-bit tile image (uint8 tensor)."""
    _lib.require_cuda(w, "weight")
    n = _lib.lib().ms_conv_packed_weight_bytes(ctypes.byref(desc))
    if n == 0:
        raise _lib.MsbError("unsupported convolution descriptor")
    out = torch.empty(n, dtype=torch.uint8, device=w.device)
    check(_lib.lib().ms_conv_pack_weight(ctypes.byref(desc), ptr(w.contiguous()), ptr(out),
                                         stream_ptr()), "ms_conv_pack_weight")
    return out


def conv_fwd(desc, x16, w_packed, bias=None, res32=None, want16=True, want32=False):
    """One tcgen05 implicit-GEMM conv / transposed conv.  Returns (y16, y32)."""
    lout = conv_out_len(desc)
    shape = (desc.batch, desc.cout // 8, lout, 8)
    y16 = torch.empty(shape, dtype=torch.int16, device=x16.device) if want16 else None
    y32 = torch.empty(shape, dtype=torch.float32, device=x16.device) if want32 else None
    check(_lib.lib().ms_conv_fwd(ctypes.byref(desc), ptr(x16), ptr(w_packed), ptr(bias),
                                 ptr(res32), ptr(y16), ptr(y32), stream_ptr()), "ms_conv_fwd")
    return y16, y32


def space_to_depth(x16, stride, length=None):
    """BLK 16-bit (B,C/8,L,8) -> (B, stride*C/8, ceil(length/stride), 8)."""
    B, C8, rows, _ = x16.shape
    length = rows if length is None else length
    lx = (length + stride - 1) // stride
    y = torch.empty((B, stride * C8, lx, 8), dtype=torch.int16, device=x16.device)
    check(_lib.lib().ms_space_to_depth_blk16(ptr(x16), ptr(y), B, C8 * 8, rows, length, stride,
                                             stream_ptr()), "ms_space_to_depth_blk16")
    return y


def strided_conv_geometry(k, stride):
    half = k // 2
    j_min = -((half + stride - 1) // stride)
    return half // stride - j_min + 1, -j_min          # (taps, pad)


def strided_conv_weight(w, stride):
    """(Cout, C, k) weight of a stride-s conv with padding k//2 -> the equivalent stride-1
    weight (Cout, s*C, taps) over the space-to-depth input, and (taps, pad) -- one kernel launch
    (ms_strided_weight_view)."""
    w = w.contiguous()
    cout, c, k = w.shape
    taps, pad = strided_conv_geometry(k, stride)
    out = torch.empty((cout, stride * c, taps), dtype=w.dtype, device=w.device)
    check(_lib.lib().ms_strided_weight_view(ptr(w), ptr(out), cout, c, k, stride, 0, stream_ptr()),
          "ms_strided_weight_view")
    return out, taps, pad


def strided_conv_weight_grad(dw1, w_shape, stride):
    """gradient of strided_conv_weight: dw1 (Cout, s*C, taps) -> dw (Cout, C, k)"""
    cout, c, k = w_shape
    dw = torch.empty((cout, c, k), dtype=dw1.dtype, device=dw1.device)
    check(_lib.lib().ms_strided_weight_view(ptr(dw1.contiguous()), ptr(dw), cout, c, k, stride, 1,
                                            stream_ptr()), "ms_strided_weight_view")
    return dw


def conv_to_mono(x32, w, bias, ksize, pad, tanh_out):
    B, C8, L, _ = x32.shape
    y = torch.empty((B, 1, L), dtype=torch.float32, device=x32.device)
    check(_lib.lib().ms_conv_to_mono(ptr(x32), ptr(w.contiguous()), ptr(bias), ptr(y), B,
                                     C8 * 8, L, ksize, pad, int(tanh_out), stream_ptr()),
          "ms_conv_to_mono")
    return y


def mel_row_ranges(mel_basis):
    """int32 (n_mels, 2): [first, last+1) non-zero column of each basis row (host-side scan)."""
    nz = (mel_basis != 0).cpu()
    n_mels, bins = nz.shape
    any_ = nz.any(dim=1)
    first = torch.where(any_, nz.float().argmax(dim=1), torch.zeros(n_mels, dtype=torch.long))
    last = torch.where(any_, bins - nz.flip(1).float().argmax(dim=1), torch.zeros(n_mels, dtype=torch.long))
    return torch.stack([first, last], dim=1).to(torch.int32).contiguous()


def audio2mel(audio, window, mel_basis, n_fft, hop, row_ranges=None):
    _lib.require_cuda(audio, "audio")
    audio = audio.contiguous()
    B, _, N = audio.shape
    n_mels = mel_basis.shape[0]
    F = _lib.lib().ms_audio2mel_frames(N, n_fft, hop)
    if F < 0:
        raise _lib.MsbError("invalid Audio2Mel geometry")
    out = torch.empty((B, n_mels, F), dtype=torch.float32, device=audio.device)
    check(_lib.lib().ms_audio2mel_fwd(ptr(audio), ptr(window.contiguous()),
                                      ptr(mel_basis.contiguous()), ptr(row_ranges), ptr(out), B, N,
                                      n_fft, hop, n_mels, stream_ptr()), "ms_audio2mel_fwd")
    return out


def act_pad(x, pad=0, pad_mode=0, leaky=False, operand=MS_F16):
    """Channel-blocked tensor (int16 = 16-bit operand, float32 = fp32 stream) ->
    act(pad(x)): zero (0) / reflection (1) padding and optional LeakyReLU(0.2)."""
    B, C8, L, _ = x.shape
    bits = 32 if x.dtype == torch.float32 else 16
    y = torch.empty((B, C8, L + 2 * pad, 8), dtype=x.dtype, device=x.device)
    check(_lib.lib().ms_blk_act_pad(ptr(x), ptr(y), bits, B, C8 * 8, L, pad, pad_mode, int(leaky),
                                    operand, stream_ptr()), "ms_blk_act_pad")
    return y


def weight_norm_fold(v, g):
    """weight_g * weight_v / ||weight_v|| (norm over all dims but 0), on the device."""
    _lib.require_cuda(v, "weight_v")
    v = v.contiguous()
    rows = v.shape[0]
    out = torch.empty_like(v)
    check(_lib.lib().ms_weight_norm_fold(ptr(v), ptr(g.contiguous()), ptr(out), rows,
                                         v.numel() // rows, stream_ptr()), "ms_weight_norm_fold")
    return out


def conv1d_direct(x, w, bias, stride=1, pad=0, groups=1, leaky=False, pad_mode=0):
    """fp32 CUDA-core conv1d (grouped / strided), NCL in -> NCL out."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, cin, lin = x.shape
    cout, cin_g, k = w.shape
    if cin_g * groups != cin:
        raise _lib.MsbError("conv1d_direct: weight / groups mismatch")
    lout = _lib.lib().ms_conv1d_out_len(lin, k, stride, pad)
    if lout < 0:
        raise _lib.MsbError("conv1d_direct: invalid geometry")
    y = torch.empty((B, cout, lout), dtype=torch.float32, device=x.device)
    check(_lib.lib().ms_conv1d_direct_fwd(ptr(x), ptr(w.contiguous()), ptr(bias), ptr(y), B, cin,
                                          cout, lin, k, stride, pad, groups, int(leaky), pad_mode,
                                          stream_ptr()), "ms_conv1d_direct_fwd")
    return y


def avg_pool1d(x, ksize, stride, pad, count_include_pad=True):
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, C, lin = x.shape
    lout = _lib.lib().ms_conv1d_out_len(lin, ksize, stride, pad)
    y = torch.empty((B, C, lout), dtype=torch.float32, device=x.device)
    check(_lib.lib().ms_avg_pool1d_fwd(ptr(x), ptr(y), B * C, lin, ksize, stride, pad,
                                       int(count_include_pad), stream_ptr()), "ms_avg_pool1d_fwd")
    return y


def resstack_supported(channels):
    return bool(_lib.lib().ms_resstack_supported(channels))


def resstack_pack_weights(params, channels, operand=MS_F16):
    """12 tensors (w,b of main.{a}.main.{0,1}) -> packed blob for the fused stack."""
    L = _lib.lib()
    n = L.ms_resstack_packed_weight_bytes(channels)
    if n == 0 or len(params) != 12:
        raise _lib.MsbError("fused ResidualStack: unsupported configuration")
    keep = [p.detach().contiguous() for p in params]
    for p in keep:
        _lib.require_cuda(p, "parameter")
    arr = (ctypes.c_void_p * 12)(*[p.data_ptr() for p in keep])
    blob = torch.empty(n, dtype=torch.uint8, device=keep[0].device)
    check(L.ms_resstack_pack_weights(arr, channels, operand, ptr(blob), stream_ptr()),
          "ms_resstack_pack_weights")
    return blob


def resstack_fwd(x32, blob, dilations, operand=MS_F16, want16=False, want32=True):
    """Fused ResidualStack on a BLK f32 tensor (B,C/8,L,8).  Returns (y16, y32)."""
    B, C8, L, _ = x32.shape
    y16 = torch.empty((B, C8, L, 8), dtype=torch.int16, device=x32.device) if want16 else None
    y32 = torch.empty_like(x32) if want32 else None
    dil = (ctypes.c_int * 3)(*dilations)
    check(_lib.lib().ms_resstack_fwd(C8 * 8, B, L, dil, operand, ptr(x32), ptr(blob), ptr(y16),
                                     ptr(y32), stream_ptr()), "ms_resstack_fwd")
    return y16, y32


def upstack_pack_weights(params, channels, operand=MS_F16):
    """14 tensors (ConvTranspose1d w, b, then the stack's 12) -> packed blob of the fused stage."""
    L = _lib.lib()
    n = L.ms_upstack_packed_weight_bytes(channels)
    if n == 0 or len(params) != 14:
        raise _lib.MsbError("fused upsampling stage: unsupported configuration")
    keep = [p.detach().contiguous() for p in params]
    for p in keep:
        _lib.require_cuda(p, "parameter")
    arr = (ctypes.c_void_p * 14)(*[p.data_ptr() for p in keep])
    blob = torch.empty(n, dtype=torch.uint8, device=keep[0].device)
    check(L.ms_upstack_pack_weights(arr, channels, operand, ptr(blob), stream_ptr()),
          "ms_upstack_pack_weights")
    return blob


def upstack_fwd(x16, blob, dilations, operand=MS_F16, want16=True, want32=False, tail=None):
    """ConvTranspose1d(2C->C, 4, 2, 1) + LeakyReLU + ResidualStack(C) on a BLK 16-bit tensor
    (B,2C/8,lin,8).  Returns (y16, y32), or the (B,1,2*lin) waveform when tail=(w, b) (C = 32)."""
    B, C8in, lin, _ = x16.shape
    C8 = C8in // 2
    dil = (ctypes.c_int * 3)(*dilations)
    if tail is not None:
        tw, tb = tail
        y = torch.empty((B, 1, 2 * lin), dtype=torch.float32, device=x16.device)
        check(_lib.lib().ms_upstack_fwd(C8 * 8, B, lin, dil, operand, ptr(x16), ptr(blob), None,
                                        None, ptr(tw), ptr(tb), ptr(y), stream_ptr()),
              "ms_upstack_fwd")
        return y
    y16 = torch.empty((B, C8, 2 * lin, 8), dtype=torch.int16, device=x16.device) if want16 else None
    y32 = torch.empty((B, C8, 2 * lin, 8), dtype=torch.float32, device=x16.device) if want32 else None
    check(_lib.lib().ms_upstack_fwd(C8 * 8, B, lin, dil, operand, ptr(x16), ptr(blob), ptr(y16),
                                    ptr(y32), None, None, None, stream_ptr()), "ms_upstack_fwd")
    return y16, y32


class MelGanWeights:
    """Packed parameter blob of a MelGanGenerator (60 state-dict tensors, in order)."""

    def __init__(self, params, in_channels, operand=MS_F16):
        L = _lib.lib()
        if len(params) != _lib.MELGAN_NUM_PARAMS:
            raise _lib.MsbError("MelGanGenerator has 60 parameter tensors, got %d" % len(params))
        n = L.ms_melgan_packed_weight_bytes(in_channels, operand)
        if n == 0:
            raise _lib.MsbError("unsupported MelGanGenerator configuration")
        for p in params:
            _lib.require_cuda(p, "parameter")
        keep = [p.detach().contiguous() for p in params]
        arr = (ctypes.c_void_p * len(keep))(*[p.data_ptr() for p in keep])
        self.blob = torch.empty(n, dtype=torch.uint8, device=keep[0].device)
        self.in_channels = in_channels
        self.operand = operand
        check(L.ms_melgan_pack_weights(arr, in_channels, operand, ptr(self.blob), stream_ptr()),
              "ms_melgan_pack_weights")


def melgan_workspace_bytes(batch, frames, in_channels):
    return int(_lib.lib().ms_melgan_workspace_bytes(batch, frames, in_channels))


def melgan_generator_fwd(weights, x, workspace, out=None):
    """x (B,C,T) f32 cuda -> (B,1,256T) f32.  `workspace`: uint8 cuda tensor; the batch
    is processed in passes of as many clips as the workspace holds."""
    _lib.require_cuda(x, "x")
    x = x.contiguous()
    B, C, T = x.shape
    if C != weights.in_channels:
        raise _lib.MsbError("expected %d input channels, got %d" % (weights.in_channels, C))
    if out is None:
        out = torch.empty((B, 1, 256 * T), dtype=torch.float32, device=x.device)
    check(_lib.lib().ms_melgan_generator_fwd(ptr(weights.blob), C, weights.operand, ptr(x),
                                             ptr(out), B, T, ptr(workspace),
                                             workspace.numel(), stream_ptr()),
          "ms_melgan_generator_fwd")
    return out


# ------------------------------------------------------------------ "exact" operand mode
class ExactConv:
    """One dense conv / transposed conv in the EXACT operand mode: both operands as three-term
    bf16 splits, [x_hi, x_lo, x_hi] * [W_hi, W_hi, W_lo] over 3C channels on the same tcgen05
    kernel = x*W to ~2^-16 with the RANGE of fp32 (no operand scales, nothing can overflow or
    fall into subnormals the way fp16 operands can with weights far from the N(0, 0.02) init).
    3x the MMA work plus one split pass per layer: the parity / dynamic-range fallback of the
    fast fp16 path, not the production path.  Caches the packed weight image per weight version."""

    def __init__(self):
        self.key = None
        self.packed = None

    def __call__(self, x32, w, bias, kind=MS_CONV, dilation=1, pad=0, stride=1, leaky=False,
                 res32=None, act_in=False, pad_in=0, pad_mode=0):
        """x32 BLK f32 (B,C/8,L,8) -> y32 BLK f32.  act_in / pad_in / pad_mode: LeakyReLU and
        zero (0) / reflection (1) padding applied to the input first (fused into the split pass)."""
        B, C8, L, _ = x32.shape
        xs = blk32_split(x32, pad_in, pad_mode, act_in, MS_BF16, 3, 1.0)
        if kind == MS_CONV:
            cout, cin, k = w.shape
        else:
            cin, cout, k = w.shape
        d = conv_desc(kind, B, 3 * cin, cout, L + 2 * pad_in, k, dilation, pad, stride,
                      leaky=leaky, operand=MS_BF16)
        key = (getattr(w, "_msb_key", None) or (w.data_ptr(), w._version), kind, k, stride)
        if key != self.key:
            self.packed = pack_conv_weight(d, weight_split(w.detach(), MS_BF16, 1.0, 3, kind))
            self.key = key
        _, y32 = conv_fwd(d, xs, self.packed, bias, res32=res32, want16=False, want32=True)
        return y32
