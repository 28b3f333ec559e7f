/* msb200 -- C ABI of the B200-native stage-two waveform-synthesis hot path.
 *
 * The reference (JohnVinyard/music-synthesis) is pure Python/PyTorch: it has NO FFI
 * layer for this path; its "plugin boundary" is torch.nn.Module.forward.  The entry
 * points below are therefore what a binding for each hot-path module would call;
 * every function names the reference interface (file:line under /root/reference)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stubs.
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures: device pointers are `void*` /
 *     `float*`, the CUDA stream is passed as `void*` (a cudaStream_t / CUstream).
 *   - the caller owns all memory and the stream; the library never allocates device
 *     memory, never synchronises, and keeps no global state (thread-safe).
 *   - every function returns MS_OK (0) or a negative ms_status; ms_strerror()
 *     describes it.  There is NO CPU fallback: without a sm_100 device every compute
 *     entry point returns MS_ERR_CUDA.
 *
 * Device tensor layouts
 *   NCL f32   : (B, C, L) contiguous fp32 -- the reference's tensor layout.
 *   BLK f16   : (B, C/8, L, 8) fp16  -- "channel-blocked": 16-byte vectors of 8
 *               consecutive channels, time-major inside a block.  Operand layout of
 *               the tcgen05 implicit GEMM (any row shift = +16 bytes, no swizzle).
 *   BLK f32   : (B, C/8, L, 8) fp32  -- the fp32 residual stream, same blocking.
 */
#ifndef MSB200_H
#define MSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int ms_status;
enum {
  MS_OK = 0,
  MS_ERR_INVALID = -1,   /* bad descriptor / unsupported shape */
  MS_ERR_CUDA = -2,      /* CUDA runtime error (see ms_last_cuda_error) */
  MS_ERR_WORKSPACE = -3  /* workspace too small */
};

int ms_version(void);
const char* ms_strerror(ms_status s);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* ms_last_cuda_error(void);
/* number of kernels this library has launched in this process (all threads);
 * used by bench.py for its gpu_launches claim */
uint64_t ms_launch_count(void);

/* ---------------------------------------------------------------------------
 * Dense 1-D convolution / transposed convolution as a tcgen05 implicit GEMM.
 *   replaces F.conv1d / F.conv_transpose1d as dispatched from
 *     featuresynth/generator/full.py:24,27,31,35,39        (MelGanGenerator trunk)
 *     featuresynth/util/modules.py:358-365,384-388         (ResidualAtom)
 *     featuresynth/util/modules.py:178-184                 (LearnedUpSample)
 *     featuresynth/discriminator/full.py:19 (1024->1024 k5)
 *   kind MS_CONV     : stride-1 conv, `ksize` taps, `dilation`, zero padding `pad`
 *                      (pad may be 0 for a pre-padded input): Lout = Lin + 2*pad
 *                      - dilation*(ksize-1).
 *   kind MS_CONVT    : ConvTranspose1d with ksize == 2*stride, padding `pad`
 *                      (1 <= pad <= stride): Lout = stride*Lin + ksize - stride
 *                      - 2*pad ... for the reference's (16,8,4),(4,2,1),(8,4,2)
 *                      this is stride*Lin.  Computed in polyphase form (2 taps at the
 *                      input rate, N = stride*Cout) so there is no overlap-add.
 *   epilogue         : y = act(alpha * acc + bias) [+ res32] (`leaky` = 1) or
 *                      y = act(alpha * acc + bias + res32) (`leaky` = 2); act = LeakyReLU(0.2).
 *                      Writes BLK f16 (y16) and/or BLK f32 (y32).
 *   operands         : fp16 (default) or bf16, fp32 accumulate in tensor memory.
 * ------------------------------------------------------------------------- */
enum { MS_CONV = 0, MS_CONVT = 1 };
enum { MS_F16 = 0, MS_BF16 = 1 };

typedef struct {
  int kind;      /* MS_CONV | MS_CONVT */
  int batch;     /* B */
  int cin;       /* multiple of 16 */
  int cout;      /* multiple of 8 (MS_CONV: multiple of 16) */
  int lin;       /* input length */
  int ksize;     /* taps, 1..32 (MS_CONVT: a multiple >= 2 of stride) */
  int dilation;  /* MS_CONV only */
  int pad;       /* zero padding (both sides) */
  int stride;    /* MS_CONVT only (MS_CONV: must be 1) */
  int leaky;     /* 1: y = LeakyReLU(acc+bias) [+ res32]; 2: y = LeakyReLU(acc+bias+res32)
                    (DilatedStack, util/modules.py:130-135); 0: no activation */
  int operand;   /* MS_F16 | MS_BF16 */
  float alpha;   /* accumulator scale (1.0f normally) */
  int crop;      /* MS_CONV: output rows dropped at the end (asymmetric padding) */
  int x_repeat;  /* 0 / 1: x16 has cin channels.  r > 1: x16 has cin / r channels and the K loop
                    wraps around them r times -- the operand of a weight-split conv
                    [x, x] * [W_hi, W_lo] without materialising the duplicated tensor */
} ms_conv_desc;

/* output length for a descriptor (or <0 on invalid) */
int ms_conv_out_len(const ms_conv_desc* d);
/* bytes of the packed (tile-ordered, 16-bit) weight image for this descriptor */
size_t ms_conv_packed_weight_bytes(const ms_conv_desc* d);
/* pack fp32 weights in the reference layout -- (Cout,Cin,K) for MS_CONV,
 * (Cin,Cout,K) for MS_CONVT, i.e. nn.Conv1d.weight / nn.ConvTranspose1d.weight --
 * into the packed 16-bit image (device to device). */
ms_status ms_conv_pack_weight(const ms_conv_desc* d, const float* w_f32, void* w_packed,
                              void* stream);
/* x16: BLK f16 (B,cin/8,lin,8); bias: fp32 [cout] or NULL; res32: BLK f32
 * (B,cout/8,Lout,8) or NULL; y16 / y32: outputs, either may be NULL. */
ms_status ms_conv_fwd(const ms_conv_desc* d, const void* x16, const void* w_packed,
                      const float* bias, const float* res32, void* y16, float* y32,
                      void* stream);

/* ---------------------------------------------------------------------------
 * Layout conversion at the path boundary.
 *   ms_pack_ncl_to_blk16: NCL f32 (B,C,L) -> BLK 16-bit (B,C/8,L+2*pad,8) with
 *     reflection (pad_mode 1: nn.ReflectionPad1d, generator/full.py:23) or zero
 *     (pad_mode 0) padding baked in.  C multiple of 8.
 *   ms_unpack_blk32_to_ncl: BLK f32 -> NCL f32 (feature maps handed back to torch).
 * ------------------------------------------------------------------------- */
ms_status ms_pack_ncl_to_blk16(const float* x, void* y16, int batch, int channels,
                               int len, int pad, int pad_mode, int operand, void* stream);
ms_status ms_unpack_blk32_to_ncl(const float* x32, float* y, int batch, int channels,
                                 int len, void* stream);
/* split-precision operands.  ms_pack_ncl_split_blk16: NCL f32 (B,C,L) -> BLK 16-bit
 * (B, terms*C/8, L, 8): channels [0,C) = hi = to16(s*x), [C,2C) = lo = to16(s*x - hi) and, for
 * terms = 3, [2C,3C) = hi again.  ms_blk32_split_blk16: the same split of a BLK f32 tensor
 * (optionally LeakyReLU and zero / reflection padding first).  ms_weight_split: w (cout,cin,k) ->
 * (cout,terms*cin,k) = [s*w, s*w, s*w - to16(s*w)] (terms = 3) or [s*w, s*w - to16(s*w)]
 * (terms = 2; a ConvTranspose1d weight (cin,cout,k) is split along cin by passing cout = 1,
 * cin = cin, ksize = cout*k).  The 3C-channel conv [x_hi,x_lo,x_hi]*[W_hi,W_hi,W_lo]
 * (alpha = 1/(s_x*s_w)) equals x*W to ~2^-22 (fp16) / ~2^-16 (bf16: the "exact" operand mode,
 * range of fp32, no scales needed); the 2C-channel conv [x,x]*[W_hi,W_lo] removes the weight
 * rounding only ("weight-split" mode: 2x the MMA work, ~30 % less forward error).
 * The power-of-two scales keep the fp16 lo terms out of the subnormals.
 * Used for the dense 1024 -> 1024 layer of the discriminator (discriminator/full.py:19): its
 * activations are bias-dominated and the real/fake gradients of a GAN step cancel to first
 * order, so LeakyReLU masks and weight gradients hinge on differences of ~1e-4 relative. */
ms_status ms_pack_ncl_split_blk16(const float* x, void* y16, int batch, int channels, int len,
                                  int operand, int terms, float scale, void* stream);
ms_status ms_weight_split(const float* w, float* out, int cout, int cin, int ksize, int operand,
                          float scale, int terms, void* stream);
ms_status ms_blk32_split_blk16(const float* x32, void* y16, int batch, int channels, int len,
                               int pad, int pad_mode, int leaky, int operand, int terms,
                               float scale, void* stream);
/* weight of a stride-s conv with padding k/2 as the stride-1 conv over the space-to-depth input
 * (ms_space_to_depth_blk16): w (cout, C, k) -> w1 (cout, s*C, taps) with taps = (k/2)/s +
 * ceil((k/2)/s) + 1 (backward = 0), or the inverse gather of its gradient dw1 -> dw (backward = 1) */
ms_status ms_strided_weight_view(const float* src, float* dst, int cout, int channels, int ksize,
                                 int stride, int backward, void* stream);
/* noise head of ResidualStackFilterBankGenerator (generator/filterbank.py:76-86):
 * y[b,t] = add[b,t] + sum_c a32[b,c,t] * n32[c,t]; a32 BLK f32 (B,C/8,L,8), n32 BLK f32
 * (1,C/8,L,8) shared by the batch, add (B,1,L) or NULL.  _bwd: da32 = dy[b,t] * n32[c,t]. */
ms_status ms_noise_mix_fwd(const float* a32, const float* n32, const float* add, float* y,
                           int batch, int channels, int len, void* stream);
ms_status ms_noise_mix_bwd(const float* dy, const float* n32, float* da32, int batch, int channels,
                           int len, void* stream);
/* front end of LowResSpectrogramDiscriminator (util/modules.py:315-325): relu, then the mean
 * over (channel_window x time_window) windows: BLK f32 (B,C/8,L,8) -> BLK f32 and/or 16-bit
 * (B,(C/cw)/8,L/tw,8).  cw must divide 8 or be a multiple of 8; tw must divide L. */
ms_status ms_relu_avgpool2d_fwd(const float* x32, void* y16, float* y32, int batch, int channels,
                                int len, int channel_window, int time_window, int operand,
                                void* stream);
ms_status ms_relu_avgpool2d_bwd(const float* dy32, const float* x32, float* dx32, int batch,
                                int channels, int len, int channel_window, int time_window,
                                void* stream);
/* space-to-depth along time: BLK 16-bit (B,C/8,src_rows,8) -> (B, s*C/8, ceil(len/s), 8) with
 * Y[u, i*C + c] = X[s*u + i, c] for s*u + i < len (0 beyond).  Turns the stride-s k7 convs of
 * featuresynth/discriminator/multiscale.py:83-88 into stride-1 convs over s*C channels. */
ms_status ms_space_to_depth_blk16(const void* x16, void* y16, int batch, int channels,
                                  int src_rows, int len, int stride, void* stream);
/* y[t'] = act(x[map(t'-pad)]) on a channel-blocked tensor (elem_bits 16 or 32): zero / reflection
 * padding (+ LeakyReLU 0.2) -- the pre-activation + ReflectionPad1d in front of the convs of the
 * official MelGAN blocks (experiment/realmelgan.py:35-37,80-81). */
ms_status ms_blk_act_pad(const void* x, void* y, int elem_bits, int batch, int channels, int len,
                         int pad, int pad_mode, int leaky, int operand, void* stream);
/* weight-norm fold out[r,:] = g[r] * v[r,:] / ||v[r,:]|| (torch weight_norm, dim=0;
 * experiment/realmelgan.py:24-29). */
ms_status ms_weight_norm_fold(const float* v, const float* g, float* out, int rows, int cols,
                              void* stream);
ms_status ms_unpack_blk16_to_ncl(const void* x16, float* y, int batch, int channels,
                                 int len, int operand, void* stream);

/* ---------------------------------------------------------------------------
 * Building blocks of the fixed Morlet filter bank (zounds.learn.FilterBank as used at
 * featuresynth/discriminator/multiscale.py:112 and generator/multiscale.py:91).  Both bank
 * operations are re-shaped so that their 2*n*k FLOP/sample run on the tcgen05 conv kernel:
 *   analysis  conv1d(x(B,1,L), bank(n,1,k), padding=k/2):
 *       ms_expand_mono_to_blk16 (Y[u,i] = x[u+i-shift], 16 channels) then ms_conv_fwd as a
 *       16 -> n channel conv with k/16 taps of dilation 16;
 *   synthesis conv_transpose1d(x(B,n,L+1), bank, padding=k/2) -> (B,1,L):
 *       ms_conv_fwd as an n -> 16 channel conv (8 phase channels used) with k/8 taps of
 *       dilation 8, then ms_diag_sum (y[t] = sum_i z[t+i+skew, i]).
 * ------------------------------------------------------------------------- */
ms_status ms_expand_mono_to_blk16(const float* x, void* y16, int batch, int len, int out_len,
                                  int shift, int operand, void* stream);
ms_status ms_diag_sum(const float* z32, float* y, int batch, int channels, int z_len,
                      int out_len, int nphase, int skew, void* stream);

/* ---------------------------------------------------------------------------
 * Single-output-channel conv + optional tanh, fp32 CUDA-core path.
 *   replaces Conv1d(32,1,7,1,3) + Tanh at generator/full.py:43-44 and the judge
 *   convs (1024->1 k3) at discriminator/full.py:22.
 *   x32: BLK f32 (B,cin/8,L,8); w: fp32 (1,cin,ksize) reference layout; y: (B,1,L).
 * ------------------------------------------------------------------------- */
ms_status ms_conv_to_mono(const float* x32, const float* w, const float* bias, float* y,
                          int batch, int cin, int len, int ksize, int pad, int tanh_out,
                          void* stream);

/* ---------------------------------------------------------------------------
 * Grouped / strided conv1d on CUDA cores (fp32, NCL in / NCL out, fused bias + LeakyReLU)
 * and average pooling.
 *   replaces F.conv1d(..., stride, padding, groups) at featuresynth/discriminator/full.py:
 *   13-18 (Conv1d(1,16,15,1,7) and the k41 s4 grouped convs, 4 input channels per group)
 *   and F.avg_pool1d(x, 4, 2, 2) (count_include_pad=True) at discriminator/melgan.py:22.
 *   w: (cout, cin/groups, k) reference layout.  Lout = (lin + 2*pad - k)/stride + 1.
 * ------------------------------------------------------------------------- */
int ms_conv1d_out_len(int lin, int ksize, int stride, int pad);
/* pad_mode: 0 zero padding, 1 reflection (nn.ReflectionPad1d(7) in front of the first conv of
 * NLayerDiscriminator, experiment/realmelgan.py:98-102).  count_include_pad: 1 = F.avg_pool1d
 * default (discriminator/melgan.py:22), 0 = nn.AvgPool1d(4,2,1,count_include_pad=False)
 * (experiment/realmelgan.py:166-167). */
ms_status ms_conv1d_direct_fwd(const float* x, const float* w, const float* bias, float* y,
                               int batch, int cin, int cout, int lin, int ksize, int stride,
                               int pad, int groups, int leaky, int pad_mode, void* stream);
ms_status ms_avg_pool1d_fwd(const float* x, float* y, int batch_channels, int lin, int ksize,
                            int stride, int pad, int count_include_pad, void* stream);

/* ---------------------------------------------------------------------------
 * Fused ResidualStack: 3 ResidualAtoms = 6 k3 convolutions (dilations d0,1,d1,1,d2,1)
 * in one kernel; activations stay in shared memory, the fp32 residual stream in
 * tensor memory.  channels in {32, 64, 128}; d0+d1+d2+3 <= 16.
 *   replaces ResidualStack.forward, featuresynth/util/modules.py:391-405
 *   (atoms: 350-388) as used at generator/full.py:29,33,37,41.
 *   params: 12 device pointers in state-dict order (w,b of main.{a}.main.{0,1}).
 *   x32: BLK f32 (B,C/8,L,8) in; y16 (BLK 16-bit) and/or y32 (BLK f32) out.
 * ------------------------------------------------------------------------- */
int ms_resstack_supported(int channels);
size_t ms_resstack_packed_weight_bytes(int channels);
ms_status ms_resstack_pack_weights(const float* const* params /* 12 device ptrs */,
                                   int channels, int operand, void* packed, void* stream);
ms_status ms_resstack_fwd(int channels, int batch, int len, const int* dilations /* [3] */,
                          int operand, const float* x32, const void* packed, void* y16,
                          float* y32, void* stream);

/* Last generator stage in one kernel: ResidualStack(32) followed by the 32 -> 1 k7 conv
 * (zero pad 3) + tanh of generator/full.py:43-44.  y: (B,1,L) fp32. */
ms_status ms_resstack_tail_fwd(int batch, int len, const int* dilations /* [3] */,
                               int operand, const float* x32, const void* packed,
                               const float* tail_w /* (1,32,7) */, const float* tail_b,
                               float* y, void* stream);

/* One generator stage in one kernel: ConvTranspose1d(2C -> C, k 4, stride 2, pad 1) +
 * LeakyReLU(0.2) + ResidualStack(C) [+ the 32 -> 1 k7 conv + tanh when tail_y != NULL, C = 32].
 *   replaces generator/full.py:35-37 (C = 64) and 39-44 (C = 32); atoms util/modules.py:350-405.
 *   params: 14 device pointers in state-dict order (ConvTranspose w (2C,C,4), b, then the
 *   stack's 12).  x16: BLK 16-bit (B,2C/8,lin,8), the previous stage's operand image;
 *   y16 / y32: BLK (B,C/8,2*lin,8) or NULL; tail_y: (B,1,2*lin) f32.  Dilations must be odd
 *   (phase-split tiles), d0+d1+d2+3 <= 16. */
int ms_upstack_supported(int channels);
size_t ms_upstack_packed_weight_bytes(int channels);
ms_status ms_upstack_pack_weights(const float* const* params /* 14 device ptrs */, int channels,
                                  int operand, void* packed, void* stream);
ms_status ms_upstack_fwd(int channels, int batch, int lin, const int* dilations /* [3] */,
                         int operand, const void* x16, const void* packed, void* y16, float* y32,
                         const float* tail_w, const float* tail_b, float* tail_y, void* stream);

/* ---------------------------------------------------------------------------
 * FFT octave-band split / merge (the fixed multiscale FFT filterbank).
 *   replaces fft_frequency_decompose / fft_resample / fft_frequency_recompose,
 *   featuresynth/audio/transform.py:50-115 (MultiScale.from_audio / to_audio,
 *   audio/representation.py:82-103).  x: (B,1,n) f32, n and min_size powers of two.
 *   decompose: band i has size min_size << i (i = 0 .. nbands-1, last = n); bands_out[i] is
 *   a device pointer to (B,1,size_i) f32.  recompose: any subset of bands -> (B,1,desired).
 *   `bands_out` / `bands` / `sizes` are HOST arrays (of device pointers / ints).
 * ------------------------------------------------------------------------- */
size_t ms_fft_bands_workspace_bytes(int batch, int n);
ms_status ms_fft_frequency_decompose(const float* x, int batch, int n, int min_size,
                                     float* const* bands_out, int nbands, void* workspace,
                                     size_t workspace_bytes, void* stream);
ms_status ms_fft_frequency_recompose(const float* const* bands, const int* sizes, int nbands,
                                     int batch, int desired_size, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream);

/* Gradients of the band split / merge (training with decompose=True / recompose=True,
 * generator/multiscale.py:166-178, discriminator/multiscale.py:212-252).  The merge is the adjoint
 * of the split and vice versa up to one bin per band (bin S/2: Nyquist of the S-point transform,
 * interior of the n-point one), a rank-1 term these calls apply IN PLACE:
 *   d(split)/dx^T g  = ms_fft_frequency_recompose(g)   then ms_fft_decompose_adjoint_fix(g, .., dx)
 *   d(merge)/db^T g  = ms_fft_frequency_decompose(g)   then ms_fft_recompose_adjoint_fix(g, .., db)
 * `dbands` / `sizes` are HOST arrays; rows = batch * channels; reductions in a fixed order. */
ms_status ms_fft_decompose_adjoint_fix(const float* const* dbands, const int* sizes, int nbands,
                                       int batch, int n, float* dx, void* stream);
ms_status ms_fft_recompose_adjoint_fix(const float* dy, int batch, int n, float* const* dbands,
                                       const int* sizes, int nbands, void* stream);

/* ---------------------------------------------------------------------------
 * GAN loss reductions (forward), deterministic (no atomics).
 *   replaces featuresynth/loss/loss.py:5-79.  out[0] (+)= weight * L(a, b) with
 *   MS_RED_L1      mean|a-b|                     F.l1_loss in mel_gan_feature_loss (28-65)
 *   MS_RED_HINGE_D mean(relu(1-a)+relu(1+b))     hinge_discriminator_loss(real, fake) (17-18)
 *   MS_RED_HINGE_G mean(-a)                      hinge_generator_loss(fake)           (9-10)
 *   MS_RED_LSQ_D   0.5(mean((a-1)^2)+mean(b^2))  least_squares_disc_loss(real, fake)  (13-14)
 *   MS_RED_LSQ_G   0.5 mean((a-1)^2)             least_squares_generator_loss(fake)   (5-6)
 *   workspace: ms_reduce_workspace_bytes() bytes of device memory.
 * ------------------------------------------------------------------------- */
enum { MS_RED_L1 = 0, MS_RED_HINGE_D = 1, MS_RED_HINGE_G = 2, MS_RED_LSQ_D = 3, MS_RED_LSQ_G = 4 };
size_t ms_reduce_workspace_bytes(void);
ms_status ms_reduce_fwd(int mode, const float* a, const float* b, size_t n, float weight,
                        float* out, int accumulate, void* workspace, void* stream);

/* ---------------------------------------------------------------------------
 * Whole MelGanGenerator forward (inference), features -> waveform.
 *   replaces MelGanGenerator.forward, featuresynth/generator/full.py:47-50
 *   (layer list 22-45; ResidualStack util/modules.py:391-405).
 *   `weights` is the packed parameter blob produced by ms_melgan_pack_weights from
 *   the 60 state-dict tensors in state-dict order.
 *   x: NCL f32 (B,128,T) ; y: NCL f32 (B,1,256*T).
 * ------------------------------------------------------------------------- */
#define MS_MELGAN_NUM_PARAMS 60
size_t ms_melgan_packed_weight_bytes(int in_channels, int operand);
ms_status ms_melgan_pack_weights(const float* const* params /* 60 device ptrs */,
                                 int in_channels, int operand, void* packed, void* stream);
size_t ms_melgan_workspace_bytes(int batch, int frames, int in_channels);
ms_status ms_melgan_generator_fwd(const void* packed_weights, int in_channels, int operand,
                                  const float* x, float* y, int batch, int frames,
                                  void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------
 * Audio2Mel: right zero-pad (n_fft-hop)/2, frame (n_fft=1024, hop), Hann window,
 * real FFT, magnitude, mel projection, log10(clamp(.,1e-5)) -- one fused kernel.
 *   replaces Audio2Mel.forward, featuresynth/feature/feature.py:39-59.
 *   audio: (B,1,N) f32; window: (1024) f32; mel_basis: (n_mels,513) f32;
 *   row_ranges: optional device int[n_mels][2] = [first, last+1) non-zero bin of each mel row
 *   (the Slaney filters are banded; NULL = treat the basis as dense -- same result);
 *   out: (B,n_mels,F) f32 with F = (N + 384 - 1024)/hop + 1.
 * ------------------------------------------------------------------------- */
int ms_audio2mel_frames(int samples, int n_fft, int hop);
ms_status ms_audio2mel_fwd(const float* audio, const float* window, const float* mel_basis,
                           const int* row_ranges, float* out, int batch, int samples, int n_fft,
                           int hop, int n_mels, void* stream);

/* ---------------------------------------------------------------------------
 * Training step (backward pass + optimiser).  The reference gets these from autograd
 * (featuresynth/train/train.py:36,70: loss.backward()) and torch.optim.Adam
 * (experiment/experiment.py:111-117); here they are explicit kernels called from
 * torch.autograd.Function.backward through this ABI.
 *   Gradients travel as BLK f32 between layers and are converted to a 16-bit GEMM operand
 *   (bf16 -- the default: the range of fp32, no loss scaling -- or fp16 under the dynamic loss
 *   scaler of train/train.py) by ms_blk_act_bwd, which also applies LeakyReLU' and
 *   forms the bias gradient.
 *   Determinism: no floating-point number is added in arrival order anywhere in the backward
 *   pass.  The kernels that reduce across thread blocks (bias / direct-conv / single-channel
 *   weight gradients) write per-block partial sums into `workspace` and the block that draws
 *   the last integer ticket adds them in block-index order (csrc/det_reduce.cuh).  Workspace
 *   contract of those entry points: the first MS_TICKET_BYTES bytes are counters that must be
 *   ZERO when the buffer is first used; every launch leaves them zero again, so one zero-fill
 *   at allocation serves all later launches on the same stream (CUDA-graph replays included).
 *   The results are ASSIGNED (dw = ..., dbias = ...), no zero-initialisation is needed.
 * ------------------------------------------------------------------------- */
#define MS_TICKET_BYTES 65536
/* dz16 = to16(dy32 * LeakyReLU'(.)), dbias[c] += sum_{b,l} dz.  Sign source: `sign16` (BLK
 * 16-bit saved activation) or the fp32 pair (ya32 - yb32 > 0; the branch of a ResidualAtom,
 * util/modules.py:384-388); none = no activation.  s2d_stride > 1 writes dz16 in the
 * space-to-depth layout of ms_space_to_depth_blk16 (input of the ConvTranspose1d dgrad).
 * dz32 (optional, NULL to skip): the same masked gradient in fp32, BLK f32 like dy32 -- the skip
 * path of a residual DilatedStack layer.  dbias may be NULL; otherwise `workspace` (see the
 * contract above, ms_blk_act_bwd_workspace_bytes) is required and dbias[c] = sum_{b,l} dz. */
size_t ms_blk_act_bwd_workspace_bytes(int batch, int channels, int len);
ms_status ms_blk_act_bwd(const float* dy32, const void* sign16, const float* ya32,
                         const float* yb32, void* dz16, float* dz32, float* dbias, int batch,
                         int channels, int len, int fmt, int s2d_stride, void* workspace,
                         size_t workspace_bytes, void* stream);
/* reference-layout fp32 weights of the convolution that computes the INPUT gradient, to be
 * packed with ms_conv_pack_weight and run with ms_conv_fwd:
 *   MS_CONV  w (cout,cin,k)  -> out (cin,cout,k) tap-reversed; run as MS_CONV cin'=cout,
 *            cout'=cin, same k / dilation, pad' = dilation*(k-1) - pad;
 *   MS_CONVT w (cin,cout,k), k = J*s -> out (cin, s*cout, ntaps); run as MS_CONV over the
 *            space-to-depth gradient (s*cout channels), ntaps taps, pad = -first_shift, with
 *            ntaps / first_shift from ms_convt_dgrad_taps (k = 2s, pad = s/2: 3 taps, shift -1;
 *            k = 4s, pad = 3s/2: 5 taps, shift -2). */
int ms_convt_dgrad_taps(int ksize, int stride, int pad, int* first_shift);
ms_status ms_weight_dgrad_view(const float* w, float* out, int kind, int cout, int cin, int ksize,
                               int stride, int pad, void* stream);
/* Weight gradient as a tcgen05 GEMM reduced over time (MN-major operands straight from the
 * channel-blocked layout):  G[t][m][n] = sum_b sum_l a16[b,m,l] * x16[b,n,l+shifts[t]]
 * (rows outside [0,lx) are zero), scattered into dw:
 *   mode MS_CONV : Conv1d weight (cout=cm, cin=cn, k=taps); a16 = dz, x16 = layer input,
 *                  shifts[t] = t*dilation - pad;
 *   mode MS_CONVT: ConvTranspose1d weight (cin=cm, cout, ksize); a16 = layer input,
 *                  x16 = space-to-depth dz (cn = stride*cout), shifts = first_shift ..
 *                  first_shift + ntaps - 1 of ms_convt_dgrad_taps ({-1,0,1} for k = 2*stride).
 * fmt: MS_F16 | MS_BF16 of BOTH operands (tcgen05 kind::f16 does not mix them: a mixed
 * instruction descriptor raises an illegal-instruction fault on sm_100a).  dw = beta*dw + alpha*G.
 * fold = 2 (MS_CONV): x16 holds a two-term split (ms_pack_ncl_split_blk16) of the layer input in
 * its two channel halves; the halves of G are summed into the cn/2-channel weight gradient. */
size_t ms_wgrad_workspace_bytes(int batch, int cm, int cn, int la, int lx, int taps,
                                const int* shifts);
ms_status ms_wgrad_fwd(const void* a16, const void* x16, int batch, int cm, int cn, int la,
                       int lx, int taps, const int* shifts, int fmt, int mode, int stride,
                       int pad, int cout, int ksize, int fold, float alpha, float beta, float* dw,
                       void* workspace, size_t workspace_bytes, void* stream);
/* 16-bit operand format conversion (fp16 forward activations -> bf16 for the weight-gradient
 * GEMM, whose other operand is a bf16 gradient) */
ms_status ms_blk16_convert(const void* src, void* dst, size_t elems, int src_fmt, int dst_fmt,
                           void* stream);
/* gradients of the layout kernels of the filter-bank / strided-conv paths:
 *   ms_diag_sum_bwd          dz32[b,u,i] = dy[b, u-i-skew] (i < nphase), BLK f32 (B,channels/8,z_len,8)
 *   ms_expand_mono_bwd       dx[b,t] = sum_{i<16} de32[b, t-i+shift, i], de32 BLK f32 (B,2,exp_len,8)
 *   ms_depth_to_space_blk32  dx32[b,c,t] = dys32[b,(t%s)*C/8+c8,t/s+row_offset] for t < len and
 *                            t/s < rows_valid, else 0; dys32 BLK f32 (B, s*C/8, src_rows, 8), dx32 (B, C/8, out_rows, 8) */
ms_status ms_diag_sum_bwd(const float* dy, float* dz32, int batch, int channels, int z_len,
                          int out_len, int nphase, int skew, void* stream);
ms_status ms_expand_mono_bwd(const float* de32, float* dx, int batch, int len, int exp_len,
                             int shift, void* stream);
ms_status ms_depth_to_space_blk32(const float* dys32, float* dx32, int batch, int channels,
                                  int src_rows, int rows_valid, int row_offset, int out_rows,
                                  int len, int stride, void* stream);
/* ReflectionPad1d on a plain (rows, len) fp32 tensor, y: (rows, len + 2*pad), and its gradient
 * (gathers in a fixed order: deterministic).
 *   replaces nn.ReflectionPad1d(7) of NLayerDiscriminator, experiment/realmelgan.py:98-102. */
ms_status ms_reflect_pad_ncl(const float* x, float* y, int rows, int len, int pad, void* stream);
ms_status ms_reflect_pad_ncl_bwd(const float* dy, float* dx, int rows, int len, int pad,
                                 void* stream);

/* gradient of ms_blk_act_pad on the fp32 stream: dx32[t] = act'(x[t]) * (dy32[t+pad] + the rows that
 * reflect onto t); sign16 = BLK 16-bit image of x (LeakyReLU mask) or NULL (no activation).
 * dy32 BLK f32 (B,C/8,len+2*pad,8), dx32 (B,C/8,len,8). */
ms_status ms_blk_act_pad_bwd(const float* dy32, const void* sign16, float* dx32, int batch,
                             int channels, int len, int pad, int pad_mode, void* stream);
/* gradient of ms_weight_norm_fold: dW (rows,cols) -> dv (rows,cols), dg (rows) */
ms_status ms_weight_norm_bwd(const float* dw, const float* v, const float* g, float* dv, float* dg,
                             int rows, int cols, void* stream);
/* NCL f32 -> BLK f32 (gradient of ms_unpack_blk32_to_ncl) */
ms_status ms_pack_ncl_to_blk32(const float* x, float* y32, int batch, int channels, int len,
                               void* stream);
/* backward of ms_conv1d_direct_fwd (zero padding).  `y` = forward output (LeakyReLU mask) when
 * leaky.  dw / dbias are assigned (deterministic two-stage sums through `workspace`, contract
 * above); dbias may be NULL. */
ms_status ms_conv1d_direct_dgrad(const float* dy, const float* y, const float* w, float* dx,
                                 int batch, int cin, int cout, int lin, int ksize, int stride,
                                 int pad, int groups, int leaky, void* stream);
size_t ms_conv1d_direct_wgrad_workspace_bytes(int batch, int cin, int cout, int lin, int ksize,
                                              int stride, int pad, int groups);
ms_status ms_conv1d_direct_wgrad(const float* dy, const float* y, const float* x, float* dw,
                                 float* dbias, int batch, int cin, int cout, int lin, int ksize,
                                 int stride, int pad, int groups, int leaky, void* workspace,
                                 size_t workspace_bytes, void* stream);
/* backward of ms_conv_to_mono: dzm (B,1,L) scratch/out = dy * (1 - y_tanh^2) (y_tanh NULL: no
 * tanh); dx32 BLK f32 (may be NULL); dw (cin,k) / dbias (1) assigned, may be NULL (`workspace`,
 * contract above, is needed with dw). */
size_t ms_conv_to_mono_bwd_workspace_bytes(int batch, int cin, int len, int ksize);
ms_status ms_conv_to_mono_bwd(const float* dy, const float* y_tanh, const float* x32,
                              const float* w, float* dzm, float* dx32, float* dw, float* dbias,
                              int batch, int cin, int len, int ksize, int pad, void* workspace,
                              size_t workspace_bytes, void* stream);
ms_status ms_avg_pool1d_bwd(const float* dy, float* dx, int batch_channels, int lin, int ksize,
                            int stride, int pad, int count_include_pad, void* stream);
/* gradients of the ms_reduce_fwd terms wrt a (da) and b (db), either may be NULL;
 * grad_out: device scalar = upstream gradient of the loss (NULL = 1). */
ms_status ms_reduce_bwd(int mode, const float* a, const float* b, size_t n, float weight,
                        const float* grad_out, float* da, float* db, void* stream);
/* torch.optim.Adam step (no amsgrad / weight decay) on a flat fp32 buffer; `step` counts from
 * 1; grad_scale multiplies the gradient first (1/world_size after a summing all-reduce). */
ms_status ms_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       size_t n, float lr, float beta1, float beta2, float eps, int step,
                       float grad_scale, void* stream);
/* same with the step counter in device memory: *step_dev is incremented, then used -- a CUDA
 * graph of the whole training step can be replayed without host-side state */
ms_status ms_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           size_t n, float lr, float beta1, float beta2, float eps, int* step_dev,
                           float grad_scale, const int* skip_flag, void* stream);
/* dynamic loss scaling of the fp16 backward: grad *= *inv_scale_dev in place; *flag_dev = 1 when any
 * element is not finite (an fp16 gradient operand overflowed), else 0.  ms_adam_step_dev with
 * skip_flag = flag_dev then leaves parameters, moments and the step counter untouched. */
ms_status ms_grad_unscale_check(float* grad, size_t n, const float* inv_scale_dev, int* flag_dev,
                                void* stream);

/* ---------------------------------------------------------------------------
 * Data feed: aligned random crops of the resident audio / log-mel stores into a training batch.
 *   replaces the per-example slicing of batch_stream / random_slice,
 *   featuresynth/data/datastore.py:8-80 (positions are drawn by the caller).
 *   store: flat f32 buffer holding every chunk; plan: device int64 (B,3) rows
 *   {origin, pitch, valid}: out[b,c,t] = store[origin + c*pitch + t] for t < valid, else 0
 *   (the reference's zero padding of short chunks); out: (B, channels, len) f32.
 * ------------------------------------------------------------------------- */
ms_status ms_gather_crops(const float* store, const long long* plan, float* out, int batch,
                          int channels, int len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MSB200_H */
