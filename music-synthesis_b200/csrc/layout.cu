// Boundary kernels: NCL f32 <-> channel-blocked layouts, and the single-output-channel
// convolution (+tanh) that ends the generator.  All HBM-bound, vectorised 16/32-byte
// accesses, coalesced along time.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

__device__ __forceinline__ uint32_t pack2op(float a, float b, int operand) {
  if (operand == MS_BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

// (B,C,L) f32 -> (B,C/8,L+2*pad,8) 16-bit; one thread per output 16-byte vector.
__global__ void pack_ncl_to_blk16_kernel(const float* __restrict__ x, uint4* __restrict__ y,
                                         int C, int L, int pad, int pad_mode, int operand,
                                         size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Lp = L + 2 * pad;
  const int tp = static_cast<int>(i % Lp);
  const size_t bc = i / Lp;  // b*(C/8) + chunk
  int t = tp - pad;
  bool zero = false;
  if (t < 0) {
    if (pad_mode == 1) t = -t; else zero = true;
  } else if (t >= L) {
    if (pad_mode == 1) t = 2 * (L - 1) - t; else zero = true;
  }
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j)
    f[j] = zero ? 0.f : __ldg(x + (bc * 8 + j) * L + t);
  uint4 o;
  o.x = pack2op(f[0], f[1], operand);
  o.y = pack2op(f[2], f[3], operand);
  o.z = pack2op(f[4], f[5], operand);
  o.w = pack2op(f[6], f[7], operand);
  y[i] = o;
}

// two-term split (hi, lo) of an NCL f32 tensor into the two channel halves of a BLK 16-bit tensor
__global__ void pack_ncl_split_blk16_kernel(const float* __restrict__ x, uint4* __restrict__ y,
                                            int C8, int L, int operand, int terms,
                                            float scale, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % L);
  const size_t bc = i / L;            // b * C8 + c8
  const size_t b = bc / C8;
  const int c8 = static_cast<int>(bc - b * C8);
  float f[8], lo[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = scale * __ldg(x + (bc * 8 + j) * L + t);
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack2op(f[2 * j], f[2 * j + 1], operand);
    float2 back;
    if (operand == MS_BF16) back = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[j]));
    else back = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
    lo[2 * j] = f[2 * j] - back.x;
    lo[2 * j + 1] = f[2 * j + 1] - back.y;
  }
  const size_t ohi = (b * terms * C8 + c8) * L + t;
  const size_t olo = (b * terms * C8 + C8 + c8) * L + t;
  y[ohi] = make_uint4(h[0], h[1], h[2], h[3]);
  if (terms == 3) y[(b * terms * C8 + 2 * C8 + c8) * L + t] = make_uint4(h[0], h[1], h[2], h[3]);
  y[olo] = make_uint4(pack2op(lo[0], lo[1], operand), pack2op(lo[2], lo[3], operand),
                      pack2op(lo[4], lo[5], operand), pack2op(lo[6], lo[7], operand));
}

// weight of the three-term split conv: out (Co, 3*Ci, K) = [w, w, w - to16(w)] along Ci, so that
// [x_hi, x_lo, x_hi] * [W_hi, W_hi, W_lo] = x*W to ~2^-22 after the 16-bit pack
__global__ void weight_split_kernel(const float* __restrict__ w, float* __restrict__ out, int Ci,
                                    int K, int operand, float scale, int terms, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int k = static_cast<int>(i % K);
  const int ci = static_cast<int>((i / K) % Ci);
  const size_t co = i / (static_cast<size_t>(K) * Ci);
  const float v = scale * __ldg(w + i);
  const float hi = operand == MS_BF16 ? __bfloat162float(__float2bfloat16_rn(v))
                                      : __half2float(__float2half_rn(v));
  float* o = out + (co * terms * Ci) * K;
  o[(static_cast<size_t>(ci)) * K + k] = v;
  if (terms == 3) o[(static_cast<size_t>(Ci + ci)) * K + k] = v;
  o[(static_cast<size_t>((terms - 1) * Ci + ci)) * K + k] = v - hi;
}

// (hi, lo[, hi]) split of a BLK f32 tensor (B, C/8, L, 8) into the channel thirds / halves of a
// BLK 16-bit tensor (B, terms*C/8, L, 8); optional LeakyReLU / zero or reflection padding first
// (the activation + padding in front of the convs of a layer-wise chain)
__global__ void blk32_split_blk16_kernel(const float* __restrict__ x, uint4* __restrict__ y,
                                         int C8, int L, int pad, int pad_mode, int leaky,
                                         int operand, int terms, float scale, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int Lp = L + 2 * pad;
  const int tp = static_cast<int>(i % Lp);
  const size_t bc = i / Lp;            // b * C8 + c8
  const size_t b = bc / C8;
  const int c8 = static_cast<int>(bc - b * C8);
  int t = tp - pad;
  bool zero = false;
  if (t < 0 || t >= L) {
    if (pad_mode == 1) t = t < 0 ? -t : 2 * (L - 1) - t;
    else zero = true;
  }
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, lo[8];
  if (!zero) ld_global_nc_v8(x + (bc * L + t) * 8, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = scale * (leaky ? leaky02(f[j]) : f[j]);
  uint32_t h[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = pack2op(f[2 * j], f[2 * j + 1], operand);
    float2 back;
    if (operand == MS_BF16) back = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&h[j]));
    else back = __half22float2(*reinterpret_cast<const __half2*>(&h[j]));
    lo[2 * j] = f[2 * j] - back.x;
    lo[2 * j + 1] = f[2 * j + 1] - back.y;
  }
  y[(b * terms * C8 + c8) * Lp + tp] = make_uint4(h[0], h[1], h[2], h[3]);
  if (terms == 3) y[(b * terms * C8 + 2 * C8 + c8) * Lp + tp] = make_uint4(h[0], h[1], h[2], h[3]);
  y[(b * terms * C8 + C8 + c8) * Lp + tp] =
      make_uint4(pack2op(lo[0], lo[1], operand), pack2op(lo[2], lo[3], operand),
                 pack2op(lo[4], lo[5], operand), pack2op(lo[6], lo[7], operand));
}

__global__ void unpack_blk32_to_ncl_kernel(const float4* __restrict__ x, float* __restrict__ y,
                                           int L, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % L);
  const size_t bc = i / L;
  const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
  float* d = y + bc * 8 * L + t;
  d[0 * (size_t)L] = a.x; d[1 * (size_t)L] = a.y; d[2 * (size_t)L] = a.z; d[3 * (size_t)L] = a.w;
  d[4 * (size_t)L] = b.x; d[5 * (size_t)L] = b.y; d[6 * (size_t)L] = b.z; d[7 * (size_t)L] = b.w;
}

__global__ void unpack_blk16_to_ncl_kernel(const uint4* __restrict__ x, float* __restrict__ y,
                                           int L, int operand, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % L);
  const size_t bc = i / L;
  const uint4 v = __ldg(x + i);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float* d = y + bc * 8 * L + t;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f;
    if (operand == MS_BF16)
      f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    else
      f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    d[(2 * j) * (size_t)L] = f.x;
    d[(2 * j + 1) * (size_t)L] = f.y;
  }
}

// y[b,0,t] = act(bias + sum_c sum_k w[c,k] * x[b,c,t+k-pad]); x is BLK f32.
// One thread per output sample; weights staged in shared memory (broadcast reads).
__global__ void conv_to_mono_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                    const float* __restrict__ bias, float* __restrict__ y,
                                    int cin, int L, int ksize, int pad, int tanh_out) {
  extern __shared__ float sw[];  // [cin/8][ksize][8]
  const int chunks = cin >> 3;
  for (int i = threadIdx.x; i < cin * ksize; i += blockDim.x) {
    const int e = i % 8;
    const int k = (i / 8) % ksize;
    const int c = i / (8 * ksize);
    sw[i] = w[(c * 8 + e) * ksize + k];
  }
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= L) return;
  float acc0 = 0.f, acc1 = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float4* xr = reinterpret_cast<const float4*>(
        x + (static_cast<size_t>(b) * chunks + c) * static_cast<size_t>(L) * 8);
    const float* wc = sw + c * ksize * 8;
    for (int k = 0; k < ksize; ++k) {
      const int ti = t + k - pad;
      if (ti < 0 || ti >= L) continue;
      const float4 a = __ldg(xr + 2 * static_cast<size_t>(ti));
      const float4 bq = __ldg(xr + 2 * static_cast<size_t>(ti) + 1);
      const float* wk = wc + k * 8;
      acc0 = fmaf(a.x, wk[0], acc0); acc1 = fmaf(a.y, wk[1], acc1);
      acc0 = fmaf(a.z, wk[2], acc0); acc1 = fmaf(a.w, wk[3], acc1);
      acc0 = fmaf(bq.x, wk[4], acc0); acc1 = fmaf(bq.y, wk[5], acc1);
      acc0 = fmaf(bq.z, wk[6], acc0); acc1 = fmaf(bq.w, wk[7], acc1);
    }
  }
  float v = acc0 + acc1 + (bias != nullptr ? bias[0] : 0.f);
  if (tanh_out) v = tanhf(v);
  y[static_cast<size_t>(b) * L + t] = v;
}

// Same for wide inputs / short sequences (the discriminators' 1024 -> 1 judges run on 3..64 time
// steps): 32 outputs x 8 channel lanes per block, each lane sums every 8th channel block, the
// lanes combine in shared memory -- the serial 128-block channel loop becomes 16 steps.
__global__ void __launch_bounds__(256)
conv_to_mono_wide_kernel(const float* __restrict__ x, const float* __restrict__ w,
                         const float* __restrict__ bias, float* __restrict__ y, int cin, int L,
                         int ksize, int pad, int tanh_out) {
  __shared__ float sh[8][33];
  const int chunks = cin >> 3;
  const int tx = threadIdx.x & 31, lane = threadIdx.x >> 5;
  const int t = blockIdx.x * 32 + tx;
  const int b = blockIdx.y;
  float acc = 0.f;
  if (t < L) {
    for (int c = lane; c < chunks; c += 8) {
      const float4* xr = reinterpret_cast<const float4*>(
          x + (static_cast<size_t>(b) * chunks + c) * static_cast<size_t>(L) * 8);
      for (int k = 0; k < ksize; ++k) {
        const int ti = t + k - pad;
        if (ti < 0 || ti >= L) continue;
        const float4 a = __ldg(xr + 2 * static_cast<size_t>(ti));
        const float4 bq = __ldg(xr + 2 * static_cast<size_t>(ti) + 1);
        const float* wk = w + static_cast<size_t>(c) * 8 * ksize + k;     // w[(c*8+e)*ksize + k]
        acc = fmaf(a.x, __ldg(wk), acc);
        acc = fmaf(a.y, __ldg(wk + ksize), acc);
        acc = fmaf(a.z, __ldg(wk + 2 * ksize), acc);
        acc = fmaf(a.w, __ldg(wk + 3 * ksize), acc);
        acc = fmaf(bq.x, __ldg(wk + 4 * ksize), acc);
        acc = fmaf(bq.y, __ldg(wk + 5 * ksize), acc);
        acc = fmaf(bq.z, __ldg(wk + 6 * ksize), acc);
        acc = fmaf(bq.w, __ldg(wk + 7 * ksize), acc);
      }
    }
  }
  sh[lane][tx] = acc;
  __syncthreads();
  if (lane != 0 || t >= L) return;
  float v = acc;
#pragma unroll
  for (int l = 1; l < 8; ++l) v += sh[l][tx];
  v += bias != nullptr ? bias[0] : 0.f;
  if (tanh_out) v = tanhf(v);
  y[static_cast<size_t>(b) * L + t] = v;
}

// Filter-bank analysis, step 1: sliding-window expansion of a mono signal into 16 channels,
//   Y[b, u, i] = x[b, u + i - shift]   (0 outside the clip),  i = 0..15,  BLK 16-bit out.
// A 1 -> n-channel k-tap convolution then becomes a 16 -> n-channel conv with k/16 taps of
// dilation 16 -- tensor-core shaped (see ms_filterbank_* in include/msb200.h).
__global__ void expand_mono_kernel(const float* __restrict__ x, uint4* __restrict__ y, int L,
                                   int Lx, int shift, int operand, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int u = static_cast<int>(gid % Lx);
  const size_t bc = gid / Lx;          // b * 2 + chunk
  const int chunk = static_cast<int>(bc & 1);
  const size_t b = bc >> 1;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int t = u + chunk * 8 + j - shift;
    f[j] = (t >= 0 && t < L) ? __ldg(x + b * L + t) : 0.f;
  }
  uint4 o;
  o.x = pack2op(f[0], f[1], operand);
  o.y = pack2op(f[2], f[3], operand);
  o.z = pack2op(f[4], f[5], operand);
  o.w = pack2op(f[6], f[7], operand);
  y[gid] = o;
}

// Filter-bank synthesis, last step: y[b, t] = sum_{i < nphase} z[b, t + i + skew, channel i]
// over a BLK f32 tensor z (B, C/8, Lz, 8) (anti-diagonal sum of the phase channels).
__global__ void diag_sum_kernel(const float* __restrict__ z, float* __restrict__ y, int C8,
                                int Lz, int L, int nphase, int skew, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % L);
  const size_t b = gid / L;
  float acc = 0.f;
  for (int i = 0; i < nphase; ++i) {
    const int u = t + i + skew;
    if (u >= 0 && u < Lz)
      acc += __ldg(z + ((b * C8 + (i >> 3)) * Lz + u) * 8 + (i & 7));
  }
  y[gid] = acc;
}

// Weight of a stride-s conv (padding k/2) as the stride-1 conv over the space-to-depth input
// (discriminator/multiscale.py:83-88 rewritten, see ms_space_to_depth_blk16):
//   dir 0: w (Cout, C, k) -> w1 (Cout, s*C, taps), w1[co][i*C + c][t] = w[co][c][(t + jmin)*s + i + k/2]
//          (0 where that tap does not exist); dir 1: the inverse gather, dw1 -> dw (every (c, kk) has
//          exactly one image).  taps = jmax - jmin + 1, jmin = -ceil((k/2)/s), jmax = (k/2)/s.
__global__ void strided_weight_view_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                           int C, int k, int s, int taps, int jmin, int dir,
                                           size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int half = k / 2;
  if (dir == 0) {
    const int t = static_cast<int>(gid % taps);
    const int ic = static_cast<int>((gid / taps) % (s * C));
    const size_t co = gid / (static_cast<size_t>(taps) * s * C);
    const int i = ic / C, c = ic - i * C;
    const int kk = (t + jmin) * s + i + half;
    dst[gid] = (kk >= 0 && kk < k) ? __ldg(src + (co * C + c) * k + kk) : 0.f;
  } else {
    const int kk = static_cast<int>(gid % k);
    const int c = static_cast<int>((gid / k) % C);
    const size_t co = gid / (static_cast<size_t>(k) * C);
    const int m = kk - half;
    const int j = (m >= 0) ? m / s : -((-m + s - 1) / s);
    const int i = m - j * s;
    dst[gid] = __ldg(src + (co * s * C + static_cast<size_t>(i) * C + c) * taps + (j - jmin));
  }
}

// Noise head of ResidualStackFilterBankGenerator (featuresynth/generator/filterbank.py:76-86):
//   y[b, t] = add[b, t] + sum_c a[b, c, t] * n[c, t]
// a: BLK f32 (B, C/8, L, 8) (`to_noise(x)`), n: BLK f32 (1, C/8, L, 8) (the filter-bank analysis
// of one white-noise row, shared by the batch), add: (B, 1, L) harmonic part or null.
__global__ void noise_mix_kernel(const float* __restrict__ a, const float* __restrict__ n,
                                 const float* __restrict__ add, float* __restrict__ y, int C8,
                                 int L, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % L);
  const size_t b = gid / L;
  float acc = add != nullptr ? __ldg(add + gid) : 0.f;
  for (int c = 0; c < C8; ++c) {
    float fa[8], fn[8];
    ld_global_nc_v8(a + ((b * C8 + c) * L + t) * 8, fa);
    ld_global_nc_v8(n + (static_cast<size_t>(c) * L + t) * 8, fn);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc = fmaf(fa[j], fn[j], acc);
  }
  y[gid] = acc;
}

// LowResSpectrogramDiscriminator front end (featuresynth/util/modules.py:315-325):
//   y[b, f, s] = mean over channels [f*cw, (f+1)*cw) x time [s*tw, (s+1)*tw) of relu(x[b, c, t])
// x: BLK f32 (B, C/8, L, 8) (filter-bank analysis), y: BLK f32 / 16-bit (B, (C/cw)/8, L/tw, 8).
// One block = one (b, output channel block, s): 8 outputs from 8*cw channels x tw time steps;
// fixed-order block reduction (warp shuffles, then the warps in order).
__global__ void __launch_bounds__(128)
relu_avgpool2d_kernel(const float* __restrict__ x, float* __restrict__ y32,
                      uint4* __restrict__ y16, int C8, int L, int cw, int tw, int operand) {
  const int s = blockIdx.x, f8 = blockIdx.y, b = blockIdx.z;
  const int F8 = gridDim.y, S = gridDim.x;
  const int nin = cw;                          // input channel blocks feeding this output block
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int i = threadIdx.x; i < nin * tw; i += blockDim.x) {
    const int ic = i / tw, t = s * tw + (i - ic * tw);
    const int c8 = f8 * cw + ic;               // input channel block: channels c8*8 .. c8*8+7
    float f[8];
    ld_global_nc_v8(x + ((static_cast<size_t>(b) * C8 + c8) * L + t) * 8, f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int o = (ic * 8 + j) / cw;         // output bin within this block (0..7)
      const float v = fmaxf(f[j], 0.f);
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] += (q == o) ? v : 0.f;
    }
  }
  __shared__ float sh[4][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = acc[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const float inv = 1.f / static_cast<float>(cw * tw);
    float o8[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o8[j] = (((sh[0][j] + sh[1][j]) + sh[2][j]) + sh[3][j]) * inv;
    const size_t oi = (static_cast<size_t>(b) * F8 + f8) * S + s;
    if (y32 != nullptr) st_global_v8(y32 + oi * 8, o8);
    if (y16 != nullptr)
      y16[oi] = make_uint4(pack2op(o8[0], o8[1], operand), pack2op(o8[2], o8[3], operand),
                           pack2op(o8[4], o8[5], operand), pack2op(o8[6], o8[7], operand));
  }
}

// y[b, c, t'] = act(x[b, c, map(t' - pad)]) on channel-blocked tensors: zero (mode 0) or
// reflection (mode 1) padding with an optional LeakyReLU(0.2) -- the pre-activation +
// ReflectionPad1d that precede the convs of the official MelGAN blocks
// (experiment/realmelgan.py:35-37, 80-81).  ELEM = 16 (16-bit operand) or 32 (fp32 stream).
template <int ELEM>
__global__ void act_pad_kernel(const void* __restrict__ xin, void* __restrict__ yout, int L, int pad,
                               int mode, int leaky, int operand, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int Lp = L + 2 * pad;
  const int tp = static_cast<int>(gid % Lp);
  const size_t bc = gid / Lp;
  int t = tp - pad;
  bool zero = false;
  if (t < 0) { if (mode == 1) t = -t; else zero = true; }
  else if (t >= L) { if (mode == 1) t = 2 * (L - 1) - t; else zero = true; }
  float f[8];
  if (zero) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = 0.f;
  } else if (ELEM == 32) {
    ld_global_nc_v8(static_cast<const float*>(xin) + (bc * L + t) * 8, f);
  } else {
    const uint4 v = __ldg(static_cast<const uint4*>(xin) + bc * L + t);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float2 q;
      if (operand == MS_BF16) q = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
      else q = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
      f[2 * j] = q.x; f[2 * j + 1] = q.y;
    }
  }
  if (leaky) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = leaky02(f[j]);
  }
  if (ELEM == 32) {
    st_global_v8(static_cast<float*>(yout) + gid * 8, f);
  } else {
    uint4 o;
    o.x = pack2op(f[0], f[1], operand); o.y = pack2op(f[2], f[3], operand);
    o.z = pack2op(f[4], f[5], operand); o.w = pack2op(f[6], f[7], operand);
    static_cast<uint4*>(yout)[gid] = o;
  }
}

// weight normalisation fold: out[r, :] = g[r] * v[r, :] / ||v[r, :]||_2 (norm over all dims but 0,
// torch.nn.utils.weight_norm with dim=0); one block per row
__global__ void weight_norm_fold_kernel(const float* __restrict__ v, const float* __restrict__ g,
                                        float* __restrict__ out, int cols) {
  __shared__ float sh[256];
  const float* vr = v + static_cast<size_t>(blockIdx.x) * cols;
  float acc = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) acc += vr[i] * vr[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  const float scale = g[blockIdx.x] / sqrtf(sh[0]);
  for (int i = threadIdx.x; i < cols; i += blockDim.x)
    out[static_cast<size_t>(blockIdx.x) * cols + i] = vr[i] * scale;
}

// space-to-depth along time, 16-byte vectors: Y[b, i*C8 + c, u] = X[b, c, s*u + i]
__global__ void space_to_depth_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int C8,
                                      int src_rows, int len, int stride, int lx, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int u = static_cast<int>(gid % lx);
  size_t r = gid / lx;
  const int cp = static_cast<int>(r % (stride * C8));
  const size_t b = r / (stride * C8);
  const int i = cp / C8, c = cp - i * C8;
  const int t = stride * u + i;
  uint4 v = make_uint4(0u, 0u, 0u, 0u);
  if (t < len) v = __ldg(x + (b * C8 + c) * static_cast<size_t>(src_rows) + t);
  y[gid] = v;
}

ms_status conv_to_mono(const float* x32, const float* w, const float* bias, float* y, int batch,
                       int cin, int len, int ksize, int pad, int tanh_out,
                       cudaStream_t stream) {
  if (batch <= 0 || cin <= 0 || cin % 8 != 0 || len <= 0 || ksize <= 0) return MS_ERR_INVALID;
  if (cin >= 256 && batch <= 65535) {
    dim3 wgrid(ceil_div(len, 32), batch);
    conv_to_mono_wide_kernel<<<wgrid, 256, 0, stream>>>(x32, w, bias, y, cin, len, ksize, pad,
                                                        tanh_out);
    return after_launch("conv_to_mono_wide_kernel");
  }
  const size_t smem = static_cast<size_t>(cin) * ksize * sizeof(float);
  if (smem > 48 * 1024) return MS_ERR_INVALID;
  dim3 grid(ceil_div(len, 256), batch);
  conv_to_mono_kernel<<<grid, 256, smem, stream>>>(x32, w, bias, y, cin, len, ksize, pad,
                                                   tanh_out);
  return after_launch("conv_to_mono_kernel");
}

ms_status pack_ncl_to_blk16(const float* x, void* y16, int batch, int channels, int len,
                            int pad, int pad_mode, int operand, cudaStream_t stream) {
  if (batch <= 0 || channels <= 0 || channels % 8 != 0 || len <= 0 || pad < 0)
    return MS_ERR_INVALID;
  if (pad_mode == 1 && pad >= len) return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * (len + 2 * pad);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  pack_ncl_to_blk16_kernel<<<blocks, 256, 0, stream>>>(x, static_cast<uint4*>(y16), channels,
                                                       len, pad, pad_mode, operand, total);
  return after_launch("pack_ncl_to_blk16_kernel");
}

}  // namespace msb

namespace msb {

__global__ void reflect_pad_ncl_kernel(const float* __restrict__ x, float* __restrict__ y, int len,
                                       int pad, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int lp = len + 2 * pad;
  const size_t r = i / lp;
  int t = static_cast<int>(i - r * lp) - pad;
  t = t < 0 ? -t : (t >= len ? 2 * (len - 1) - t : t);
  y[i] = __ldg(x + r * len + t);
}

__global__ void reflect_pad_ncl_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx,
                                           int len, int pad, size_t total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t r = i / len;
  const int t = static_cast<int>(i - r * len);
  const float* g = dy + r * (len + 2 * pad);
  float v = __ldg(g + t + pad);
  if (t >= 1 && t <= pad) v += __ldg(g + pad - t);                              // left mirror
  if (t >= len - 1 - pad && t <= len - 2) v += __ldg(g + pad + 2 * (len - 1) - t);   // right mirror
  dx[i] = v;
}

}  // namespace msb

using namespace msb;

extern "C" {

ms_status ms_weight_split(const float* w, float* out, int cout, int cin, int ksize, int operand,
                          float scale, int terms, void* stream) {
  if (w == nullptr || out == nullptr || cout <= 0 || cin <= 0 || ksize <= 0 ||
      (terms != 2 && terms != 3))
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(cout) * cin * ksize;
  weight_split_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(w, out, cin, ksize, operand, scale,
                                                             terms, total);
  return after_launch("weight_split_kernel");
}

ms_status ms_blk32_split_blk16(const float* x32, void* y16, int batch, int channels, int len,
                               int pad, int pad_mode, int leaky, int operand, int terms,
                               float scale, void* stream) {
  if (x32 == nullptr || y16 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      len <= 0 || pad < 0 || (pad_mode == 1 && pad >= len) || (terms != 2 && terms != 3))
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * (len + 2 * pad);
  blk32_split_blk16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(
      x32, static_cast<uint4*>(y16), channels / 8, len, pad, pad_mode, leaky, operand, terms,
      scale, total);
  return after_launch("blk32_split_blk16_kernel");
}

ms_status ms_pack_ncl_split_blk16(const float* x, void* y16, int batch, int channels, int len,
                                  int operand, int terms, float scale, void* stream) {
  if (x == nullptr || y16 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      len <= 0 || (terms != 2 && terms != 3))
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  pack_ncl_split_blk16_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(
      x, static_cast<uint4*>(y16), channels / 8, len, operand, terms, scale, total);
  return after_launch("pack_ncl_split_blk16_kernel");
}

ms_status ms_pack_ncl_to_blk16(const float* x, void* y16, int batch, int channels, int len,
                               int pad, int pad_mode, int operand, void* stream) {
  if (x == nullptr || y16 == nullptr) return MS_ERR_INVALID;
  return pack_ncl_to_blk16(x, y16, batch, channels, len, pad, pad_mode, operand,
                           static_cast<cudaStream_t>(stream));
}

ms_status ms_unpack_blk32_to_ncl(const float* x32, float* y, int batch, int channels, int len,
                                 void* stream) {
  if (x32 == nullptr || y == nullptr || batch <= 0 || channels % 8 != 0 || len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  unpack_blk32_to_ncl_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(x32), y, len, total);
  return after_launch("unpack_blk32_to_ncl_kernel");
}

ms_status ms_unpack_blk16_to_ncl(const void* x16, float* y, int batch, int channels, int len,
                                 int operand, void* stream) {
  if (x16 == nullptr || y == nullptr || batch <= 0 || channels % 8 != 0 || len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  unpack_blk16_to_ncl_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x16), y, len, operand, total);
  return after_launch("unpack_blk16_to_ncl_kernel");
}

ms_status ms_blk_act_pad(const void* x, void* y, int elem_bits, int batch, int channels, int len,
                         int pad, int pad_mode, int leaky, int operand, void* stream) {
  if (x == nullptr || y == nullptr || batch <= 0 || channels % 8 != 0 || len <= 0 || pad < 0 ||
      (elem_bits != 16 && elem_bits != 32) || (pad_mode == 1 && pad >= len))
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * (len + 2 * pad);
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bits == 16)
    act_pad_kernel<16><<<blocks, 256, 0, st>>>(x, y, len, pad, pad_mode, leaky, operand, total);
  else
    act_pad_kernel<32><<<blocks, 256, 0, st>>>(x, y, len, pad, pad_mode, leaky, operand, total);
  return after_launch("act_pad_kernel");
}

ms_status ms_weight_norm_fold(const float* v, const float* g, float* out, int rows, int cols,
                              void* stream) {
  if (v == nullptr || g == nullptr || out == nullptr || rows <= 0 || cols <= 0) return MS_ERR_INVALID;
  weight_norm_fold_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(v, g, out, cols);
  return after_launch("weight_norm_fold_kernel");
}

ms_status ms_space_to_depth_blk16(const void* x16, void* y16, int batch, int channels,
                                  int src_rows, int len, int stride, void* stream) {
  if (x16 == nullptr || y16 == nullptr || batch <= 0 || channels % 8 != 0 || stride < 1 ||
      len <= 0 || len > src_rows)
    return MS_ERR_INVALID;
  const int lx = (len + stride - 1) / stride;
  const size_t total = static_cast<size_t>(batch) * stride * (channels / 8) * lx;
  space_to_depth_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                          static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(x16), static_cast<uint4*>(y16), channels / 8, src_rows, len,
      stride, lx, total);
  return after_launch("space_to_depth_kernel");
}

ms_status ms_expand_mono_to_blk16(const float* x, void* y16, int batch, int len, int out_len,
                                  int shift, int operand, void* stream) {
  if (x == nullptr || y16 == nullptr || batch <= 0 || len <= 0 || out_len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * 2 * out_len;
  expand_mono_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(x, static_cast<uint4*>(y16), len,
                                                            out_len, shift, operand, total);
  return after_launch("expand_mono_kernel");
}

ms_status ms_diag_sum(const float* z32, float* y, int batch, int channels, int z_len, int out_len,
                      int nphase, int skew, void* stream) {
  if (z32 == nullptr || y == nullptr || batch <= 0 || channels % 8 != 0 || z_len <= 0 ||
      out_len <= 0 || nphase <= 0 || nphase > channels)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * out_len;
  diag_sum_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                    static_cast<cudaStream_t>(stream)>>>(z32, y, channels / 8, z_len, out_len,
                                                         nphase, skew, total);
  return after_launch("diag_sum_kernel");
}

ms_status ms_strided_weight_view(const float* src, float* dst, int cout, int channels, int ksize,
                                 int stride, int backward, void* stream) {
  if (src == nullptr || dst == nullptr || cout <= 0 || channels <= 0 || ksize <= 0 || stride < 1)
    return MS_ERR_INVALID;
  const int half = ksize / 2;
  const int jmin = -((half + stride - 1) / stride), jmax = half / stride;
  const int taps = jmax - jmin + 1;
  const size_t total = backward ? static_cast<size_t>(cout) * channels * ksize
                                : static_cast<size_t>(cout) * stride * channels * taps;
  strided_weight_view_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                               static_cast<cudaStream_t>(stream)>>>(
      src, dst, channels, ksize, stride, taps, jmin, backward ? 1 : 0, total);
  return after_launch("strided_weight_view_kernel");
}

ms_status ms_noise_mix_fwd(const float* a32, const float* n32, const float* add, float* y,
                           int batch, int channels, int len, void* stream) {
  if (a32 == nullptr || n32 == nullptr || y == nullptr || batch <= 0 || channels <= 0 ||
      channels % 8 != 0 || len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * len;
  noise_mix_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                     static_cast<cudaStream_t>(stream)>>>(a32, n32, add, y, channels / 8, len, total);
  return after_launch("noise_mix_kernel");
}

ms_status ms_relu_avgpool2d_fwd(const float* x32, void* y16, float* y32, int batch, int channels,
                                int len, int channel_window, int time_window, int operand,
                                void* stream) {
  if (x32 == nullptr || (y16 == nullptr && y32 == nullptr) || batch <= 0 || channels <= 0 ||
      len <= 0 || channel_window < 1 || time_window < 1 || channels % (8 * channel_window) != 0 ||
      len % time_window != 0 || (8 % channel_window != 0 && channel_window % 8 != 0))
    return MS_ERR_INVALID;
  const int F8 = channels / channel_window / 8, S = len / time_window;
  if (F8 > 65535 || batch > 65535) return MS_ERR_INVALID;
  dim3 grid(S, F8, batch);
  relu_avgpool2d_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      x32, y32, static_cast<uint4*>(y16), channels / 8, len, channel_window, time_window, operand);
  return after_launch("relu_avgpool2d_kernel");
}

ms_status ms_conv_to_mono(const float* x32, const float* w, const float* bias, float* y,
                          int batch, int cin, int len, int ksize, int pad, int tanh_out,
                          void* stream) {
  if (x32 == nullptr || w == nullptr || y == nullptr) return MS_ERR_INVALID;
  return conv_to_mono(x32, w, bias, y, batch, cin, len, ksize, pad, tanh_out,
                      static_cast<cudaStream_t>(stream));
}

/* ReflectionPad1d on a plain (rows, len) fp32 tensor and its gradient (both gathers: the
 * backward sums, in a fixed order, the at most three padded positions that read input sample i).
 *   replaces nn.ReflectionPad1d(7) in front of NLayerDiscriminator's first conv,
 *   experiment/realmelgan.py:98-102. */
ms_status ms_reflect_pad_ncl(const float* x, float* y, int rows, int len, int pad, void* stream) {
  if (x == nullptr || y == nullptr || rows <= 0 || len <= 1 || pad < 0 || pad >= len)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(rows) * (len + 2 * pad);
  msb::reflect_pad_ncl_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(x, y, len, pad, total);
  return msb::after_launch("reflect_pad_ncl_kernel");
}

ms_status ms_reflect_pad_ncl_bwd(const float* dy, float* dx, int rows, int len, int pad,
                                 void* stream) {
  if (dy == nullptr || dx == nullptr || rows <= 0 || len <= 1 || pad < 0 || pad >= len)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(rows) * len;
  msb::reflect_pad_ncl_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                                    static_cast<cudaStream_t>(stream)>>>(dy, dx, len, pad, total);
  return msb::after_launch("reflect_pad_ncl_bwd_kernel");
}

}  // extern "C"
