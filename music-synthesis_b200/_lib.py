"""ctypes binding of csrc/libmsb200.so (the C ABI declared in include/msb200.h).

There is no CPU or eager-PyTorch fallback behind these calls: if the shared object is
missing, or a call returns a non-zero status, this module raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
# MSB_LIB_PATH: load another build of the same ABI (e.g. the clock-traced debug build of
# tools/stack_trace.py); there is still no fallback -- a missing file raises
LIB_PATH = os.environ.get("MSB_LIB_PATH") or os.path.join(_HERE, "csrc", "libmsb200.so")

MS_CONV, MS_CONVT = 0, 1
MS_F16, MS_BF16 = 0, 1
MELGAN_NUM_PARAMS = 60


class ConvDesc(Structure):
    """mirrors `ms_conv_desc` (include/msb200.h)"""
    _fields_ = [("kind", c_int), ("batch", c_int), ("cin", c_int), ("cout", c_int),
                ("lin", c_int), ("ksize", c_int), ("dilation", c_int), ("pad", c_int),
                ("stride", c_int), ("leaky", c_int), ("operand", c_int),
                ("alpha", c_float), ("crop", c_int), ("x_repeat", c_int)]


# name -> (restype, argtypes); every symbol include/msb200.h declares
SIGNATURES = {
    "ms_version": (c_int, []),
    "ms_strerror": (c_char_p, [c_int]),
    "ms_last_cuda_error": (c_char_p, []),
    "ms_launch_count": (c_uint64, []),
    "ms_conv_out_len": (c_int, [POINTER(ConvDesc)]),
    "ms_conv_packed_weight_bytes": (c_size_t, [POINTER(ConvDesc)]),
    "ms_conv_pack_weight": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p]),
    "ms_conv_fwd": (c_int, [POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p]),
    "ms_pack_ncl_to_blk16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p]),
    "ms_unpack_blk32_to_ncl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ms_unpack_blk16_to_ncl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_void_p]),
    "ms_expand_mono_to_blk16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p]),
    "ms_diag_sum": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                            c_void_p]),
    "ms_strided_weight_view": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p]),
    "ms_noise_mix_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                 c_void_p]),
    "ms_noise_mix_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ms_relu_avgpool2d_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "ms_relu_avgpool2d_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_void_p]),
    "ms_space_to_depth_blk16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_void_p]),
    "ms_conv_to_mono": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                c_int, c_int, c_int, c_void_p]),
    "ms_conv1d_out_len": (c_int, [c_int, c_int, c_int, c_int]),
    "ms_conv1d_direct_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ms_avg_pool1d_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "ms_blk_act_pad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_int, c_void_p]),
    "ms_weight_norm_fold": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "ms_resstack_supported": (c_int, [c_int]),
    "ms_resstack_packed_weight_bytes": (c_size_t, [c_int]),
    "ms_resstack_pack_weights": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p]),
    "ms_resstack_fwd": (c_int, [c_int, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "ms_resstack_tail_fwd": (c_int, [c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "ms_reflect_pad_ncl": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ms_reflect_pad_ncl_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ms_upstack_supported": (c_int, [c_int]),
    "ms_upstack_packed_weight_bytes": (c_size_t, [c_int]),
    "ms_upstack_pack_weights": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p]),
    "ms_upstack_fwd": (c_int, [c_int, c_int, c_int, POINTER(c_int), c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "ms_fft_bands_workspace_bytes": (c_size_t, [c_int, c_int]),
    "ms_fft_frequency_decompose": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p), c_int,
                                           c_void_p, c_size_t, c_void_p]),
    "ms_fft_frequency_recompose": (c_int, [POINTER(c_void_p), POINTER(c_int), c_int, c_int, c_int,
                                           c_void_p, c_void_p, c_size_t, c_void_p]),
    "ms_fft_decompose_adjoint_fix": (c_int, [POINTER(c_void_p), POINTER(c_int), c_int, c_int, c_int,
                                             c_void_p, c_void_p]),
    "ms_fft_recompose_adjoint_fix": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p),
                                             POINTER(c_int), c_int, c_void_p]),
    "ms_reduce_workspace_bytes": (c_size_t, []),
    "ms_reduce_fwd": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_float, c_void_p, c_int,
                              c_void_p, c_void_p]),
    "ms_melgan_packed_weight_bytes": (c_size_t, [c_int, c_int]),
    "ms_melgan_pack_weights": (c_int, [POINTER(c_void_p), c_int, c_int, c_void_p, c_void_p]),
    "ms_melgan_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ms_melgan_generator_fwd": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int,
                                        c_int, c_void_p, c_size_t, c_void_p]),
    "ms_audio2mel_frames": (c_int, [c_int, c_int, c_int]),
    "ms_audio2mel_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                 c_int, c_int, c_int, c_void_p]),
    "ms_blk_act_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "ms_blk_act_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                               c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "ms_convt_dgrad_taps": (c_int, [c_int, c_int, c_int, POINTER(c_int)]),
    "ms_weight_dgrad_view": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_void_p]),
    "ms_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int,
                                            POINTER(c_int)]),
    "ms_wgrad_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                             POINTER(c_int), c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                             c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "ms_pack_ncl_split_blk16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_float, c_void_p]),
    "ms_weight_split": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_int,
                                c_void_p]),
    "ms_blk32_split_blk16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_float, c_void_p]),
    "ms_blk16_convert": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p]),
    "ms_diag_sum_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                c_void_p]),
    "ms_expand_mono_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "ms_depth_to_space_blk32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                        c_int, c_int, c_int, c_void_p]),
    "ms_blk_act_pad_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_void_p]),
    "ms_weight_norm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                   c_void_p]),
    "ms_pack_ncl_to_blk32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "ms_conv1d_direct_dgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "ms_conv1d_direct_wgrad_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int,
                                                          c_int, c_int]),
    "ms_conv1d_direct_wgrad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                       c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p, c_size_t, c_void_p]),
    "ms_conv_to_mono_bwd_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "ms_conv_to_mono_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p, c_size_t, c_void_p]),
    "ms_avg_pool1d_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "ms_reduce_bwd": (c_int, [c_int, c_void_p, c_void_p, c_size_t, c_float, c_void_p, c_void_p,
                              c_void_p, c_void_p]),
    "ms_adam_step_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float,
                                 c_float, c_float, c_float, c_void_p, c_float, c_void_p, c_void_p]),
    "ms_grad_unscale_check": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "ms_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float,
                             c_float, c_float, c_int, c_float, c_void_p]),
    "ms_gather_crops": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


class MsbError(RuntimeError):
    pass


def lib():
    """The loaded shared library (loads on first use; raises if it was not built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s not found: the CUDA extension is not built. Run "
                "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc). "
                "There is no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if the ABI drifted
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status, what):
    if status != 0:
        L = lib()
        msg = L.ms_strerror(status).decode()
        cuda = L.ms_last_cuda_error().decode()
        raise MsbError("%s failed: %s%s" % (what, msg, (" -- " + cuda) if cuda else ""))


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return c_void_p(0 if t is None else t.data_ptr())


def launch_count():
    return int(lib().ms_launch_count())


def require_cuda(t, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise MsbError(
            "%s must be a CUDA tensor: this path has no CPU implementation" % name)
    if t.dtype != torch.float32:
        raise MsbError("%s must be float32 (got %s)" % (name, t.dtype))
    require_current_device(t, name)


def require_current_device(t, name="tensor"):
    """Every C-ABI launch goes to the CURRENT device's current stream (stream_ptr()); a tensor
    that lives on another GPU would be read by kernels running on the wrong device.  Raise
    instead (callers select the device with torch.cuda.set_device / torch.cuda.device)."""
    import torch
    if t.is_cuda and t.device.index != torch.cuda.current_device():
        raise MsbError("%s is on cuda:%d but the current device is cuda:%d: wrap the call in "
                       "torch.cuda.device(%s.device)" % (name, t.device.index,
                                                          torch.cuda.current_device(), name))
