// Backward-pass kernels of the GAN training step that are not GEMM-shaped: activation
// gradients + operand conversion, bias gradients, the grouped / strided direct convolutions
// of the discriminator, the single-channel convs, pooling, loss gradients and Adam.
// The reference gets all of these from autograd (featuresynth/train/train.py:36,70 ->
// loss.backward()) and torch.optim.Adam (experiment/experiment.py:111-117).
// All HBM/latency-bound CUDA-core work: coalesced along time, 16/32-byte vectors on the
// channel-blocked tensors.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "det_reduce.cuh"
#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

__device__ __forceinline__ uint32_t bw_pack2(float a, float b, int fmt) {
  if (fmt == MS_BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

// dz = dy * LeakyReLU'(.) on BLK tensors -> 16-bit GEMM operand (optionally written in the
// space-to-depth layout the transposed-conv backward consumes) + per-channel bias gradient.
//   sign source: s16 (sign bit of the saved 16-bit activation) or the fp32 pair (ya - yb > 0,
//   the residual branch y - x of a ResidualAtom); neither -> no activation.
struct ActBwdParams {
  const float* dy;      // BLK f32 (B, C/8, L, 8)
  const uint16_t* s16;  // BLK 16-bit or null
  const float* ya;      // BLK f32 or null
  const float* yb;      // BLK f32 or null
  uint16_t* dz;         // BLK 16-bit (B, C/8, L, 8) or s2d (B, s*C/8, L/s, 8)
  float* dbias;         // [C] = sum over batch and time (fixed order), or null
  float* dz32;          // optional fp32 copy of dz (BLK f32, same layout as dy), or null
  int C8, L, fmt, s2d;
  int rows_per_block;   // multiple of 256; gridDim.x = ceil(L / rows_per_block) <= kActBwdMaxXT
  float* part;          // [B * gridDim.x][C] partial bias sums (workspace), with dbias
  unsigned int* ticket; // one counter for the launch (workspace)
};

constexpr int kActBwdRows = 1024;   // minimum rows per block (256 threads x 4)
constexpr int kActBwdMaxXT = 8;     // at most 8 time tiles per (clip, channel group)

static int act_bwd_rows_per_block(int len) {
  int rows = kActBwdRows;
  const int need = ceil_div(len, kActBwdMaxXT);
  if (need > rows) rows = ceil_div(need, 256) * 256;
  return rows;
}

__global__ void __launch_bounds__(256)
act_bwd_kernel(const ActBwdParams p) {
  const size_t bc = blockIdx.y;                 // b * C8 + c8
  const int c8 = static_cast<int>(bc % p.C8);
  const size_t b = bc / p.C8;
  float bsum[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) bsum[j] = 0.f;
  const int t_end = min(p.L, (static_cast<int>(blockIdx.x) + 1) * p.rows_per_block);
  for (int t = blockIdx.x * p.rows_per_block + threadIdx.x; t < t_end; t += 256) {
    const size_t idx = bc * p.L + t;
    float g[8];
    ld_global_nc_v8(p.dy + idx * 8, g);
    if (p.s16 != nullptr) {
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(p.s16) + idx);
      const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (w[j] & 0x8000u) g[2 * j] *= 0.2f;
        if (w[j] & 0x80000000u) g[2 * j + 1] *= 0.2f;
      }
    } else if (p.ya != nullptr) {
      float a[8], x[8];
      ld_global_nc_v8(p.ya + idx * 8, a);
      ld_global_nc_v8(p.yb + idx * 8, x);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(a[j] - x[j] > 0.f)) g[j] *= 0.2f;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) bsum[j] += g[j];
    if (p.dz32 != nullptr) st_global_v8(p.dz32 + idx * 8, g);
    uint4 o;
    o.x = bw_pack2(g[0], g[1], p.fmt); o.y = bw_pack2(g[2], g[3], p.fmt);
    o.z = bw_pack2(g[4], g[5], p.fmt); o.w = bw_pack2(g[6], g[7], p.fmt);
    size_t oidx = idx;
    if (p.s2d > 1) {
      const int u = t / p.s2d, i = t - u * p.s2d;
      oidx = ((b * p.s2d + i) * p.C8 + c8) * static_cast<size_t>(p.L / p.s2d) + u;
    }
    reinterpret_cast<uint4*>(p.dz)[oidx] = o;
  }
  if (p.dbias == nullptr) return;
  __shared__ float sh[8][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float v = bsum[j];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][j] = v;
  }
  __syncthreads();
  const int C = p.C8 * 8;
  if (threadIdx.x < 8) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    p.part[(b * gridDim.x + blockIdx.x) * C + c8 * 8 + threadIdx.x] = v;
  }
  // the last block of the launch adds the partials in (clip, time tile) order
  if (!det_last_block(p.ticket, gridDim.x * gridDim.y)) return;
  const int np = static_cast<int>(gridDim.y / p.C8) * gridDim.x;
  for (int c = threadIdx.x; c < C; c += 256) p.dbias[c] = det_sum_strided(p.part, np, C, c);
}

// reference-layout weights of the convolution that computes the INPUT gradient:
//   MS_CONV : w (Co,Ci,K) -> out (Ci,Co,K), out[ci][co][k] = w[co][ci][K-1-k]
//   MS_CONVT: w (Ci,Co,K), stride s, padding p -> out (Ci, s*Co, ntaps) over the space-to-depth
//             gradient, out[ci][r*Co+co][t] = w[ci][co][s*(t + dmin) + r + p]  (0 where that tap
//             does not exist); ntaps = dmax - dmin + 1, see ms_convt_dgrad_taps
__global__ void weight_dgrad_view_kernel(const float* __restrict__ w, float* __restrict__ out,
                                         int kind, int co_n, int ci_n, int K, int stride, int pad,
                                         int ntaps, int dmin, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  if (kind == MS_CONV) {
    const int k = static_cast<int>(i % K);
    const int co = static_cast<int>((i / K) % co_n);
    const int ci = static_cast<int>(i / (static_cast<size_t>(K) * co_n));
    out[i] = w[(static_cast<size_t>(co) * ci_n + ci) * K + (K - 1 - k)];
  } else {
    const int t = static_cast<int>(i % ntaps);
    const int n = static_cast<int>((i / ntaps) % (stride * co_n));
    const int ci = static_cast<int>(i / (static_cast<size_t>(ntaps) * stride * co_n));
    const int r = n / co_n, co = n - r * co_n;
    const int k = stride * (t + dmin) + r + pad;
    out[i] = (k >= 0 && k < K) ? w[(static_cast<size_t>(ci) * co_n + co) * K + k] : 0.f;
  }
}

// fp16 <-> bf16 on 16-byte vectors
__global__ void blk16_convert_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst,
                                     size_t nvec, int src_fmt, int dst_fmt) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= nvec) return;
  const uint4 v = __ldg(src + i);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 f;
    if (src_fmt == MS_BF16) f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
    else f = __half22float2(*reinterpret_cast<const __half2*>(&w[j]));
    o[j] = bw_pack2(f.x, f.y, dst_fmt);
  }
  dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
}

// NCL f32 -> BLK f32 (gradient of ms_unpack_blk32_to_ncl)
__global__ void pack_ncl_to_blk32_kernel(const float* __restrict__ x, float* __restrict__ y, int L,
                                         size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % L);
  const size_t bc = i / L;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = __ldg(x + (bc * 8 + j) * L + t);
  st_global_v8(y + i * 8, f);
}

// ---- backward of the layout kernels of the filter-bank / strided-conv paths (csrc/layout.cu) ----
// ms_diag_sum: y[b,t] = sum_{i<nphase} z[b, t+i+skew, i]  ->  dz[b,u,i] = dy[b, u-i-skew]
__global__ void diag_sum_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dz, int C8,
                                    int Lz, int L, int nphase, int skew, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;                       // one thread per (b, c8, u): 8 channels
  const int u = static_cast<int>(gid % Lz);
  const size_t bc = gid / Lz;
  const int c8 = static_cast<int>(bc % C8);
  const size_t b = bc / C8;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = c8 * 8 + j;
    const int t = u - i - skew;
    f[j] = (i < nphase && t >= 0 && t < L) ? __ldg(dy + b * L + t) : 0.f;
  }
  st_global_v8(dz + gid * 8, f);
}

// ms_expand_mono_to_blk16: Y[b,u,i] = x[b, u+i-shift]  ->  dx[b,t] = sum_i dY[b, t-i+shift, i]
__global__ void expand_mono_bwd_kernel(const float* __restrict__ dE, float* __restrict__ dx, int L,
                                       int Lx, int shift, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % L);
  const size_t b = gid / L;
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int u = t - i + shift;
    if (u >= 0 && u < Lx) acc += __ldg(dE + ((b * 2 + (i >> 3)) * Lx + u) * 8 + (i & 7));
  }
  dx[gid] = acc;
}

// ms_space_to_depth_blk16: Y[b, i*C8+c, u] = X[b, c, s*u+i]  ->  dX[b,c,t] = dY[b,(t%s)*C8+c,t/s]
// (fp32; dY row u is stored at row u + row_off; rows u >= rows_valid do not exist = zero; rows
// t >= len of dX are zero)
__global__ void depth_to_space_blk32_kernel(const float* __restrict__ dys, float* __restrict__ dx,
                                            int C8, int lx_rows, int rows_valid, int row_off,
                                            int out_rows, int len, int stride, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % out_rows);
  const size_t bc = gid / out_rows;
  const int c = static_cast<int>(bc % C8);
  const size_t b = bc / C8;
  float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int u = t / stride, i = t - u * stride;
  if (t < len && u < rows_valid)
    ld_global_nc_v8(dys + ((b * stride * C8 + static_cast<size_t>(i) * C8 + c) * lx_rows + u + row_off) * 8, f);
  st_global_v8(dx + gid * 8, f);
}

// gradient of ms_blk_act_pad (fp32 stream): y[tp] = act(x[map(tp - pad)]), zero (mode 0) or
// reflection (mode 1) padding: dx[t] = act'(x[t]) * (dy[t+pad] + reflected copies of row t).
// sign16: BLK 16-bit image of x (LeakyReLU mask) or null (no activation).
__global__ void act_pad_bwd_kernel(const float* __restrict__ dy, const uint4* __restrict__ sign16,
                                   float* __restrict__ dx, int L, int pad, int mode,
                                   size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int Lp = L + 2 * pad;
  const int t = static_cast<int>(gid % L);
  const size_t bc = gid / L;
  const float* row = dy + bc * static_cast<size_t>(Lp) * 8;
  float g[8];
  ld_global_nc_v8(row + static_cast<size_t>(t + pad) * 8, g);
  if (mode == 1) {
    float r[8];
    if (t >= 1 && t <= pad) {                     // left reflection: padded row pad - t
      ld_global_nc_v8(row + static_cast<size_t>(pad - t) * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += r[j];
    }
    if (t <= L - 2 && t >= L - 1 - pad) {         // right reflection: padded row pad + 2(L-1) - t
      ld_global_nc_v8(row + static_cast<size_t>(pad + 2 * (L - 1) - t) * 8, r);
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] += r[j];
    }
  }
  if (sign16 != nullptr) {
    const uint4 v = __ldg(sign16 + gid);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (w[j] & 0x8000u) g[2 * j] *= 0.2f;
      if (w[j] & 0x80000000u) g[2 * j + 1] *= 0.2f;
    }
  }
  st_global_v8(dx + gid * 8, g);
}

// gradient of ms_weight_norm_fold (W = g * v / ||v||, rows = dim 0): one block per row
//   dg[r] = <dW_r, v_r> / ||v_r||,   dv_r = (g_r/||v_r||) * (dW_r - (dg[r]/||v_r||) * v_r)
__global__ void weight_norm_bwd_kernel(const float* __restrict__ dw, const float* __restrict__ v,
                                       const float* __restrict__ g, float* __restrict__ dv,
                                       float* __restrict__ dg, int cols) {
  __shared__ float sh[2][256];
  const size_t base = static_cast<size_t>(blockIdx.x) * cols;
  float nn = 0.f, dot = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) {
    const float vi = v[base + i];
    nn += vi * vi;
    dot += dw[base + i] * vi;
  }
  sh[0][threadIdx.x] = nn;
  sh[1][threadIdx.x] = dot;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) {
      sh[0][threadIdx.x] += sh[0][threadIdx.x + s];
      sh[1][threadIdx.x] += sh[1][threadIdx.x + s];
    }
    __syncthreads();
  }
  const float n = sqrtf(sh[0][0]);
  const float dgr = sh[1][0] / n;
  const float gr = g[blockIdx.x];
  if (threadIdx.x == 0) dg[blockIdx.x] = dgr;
  for (int i = threadIdx.x; i < cols; i += blockDim.x)
    dv[base + i] = (gr / n) * (dw[base + i] - (dgr / n) * v[base + i]);
}

// gradient of ms_noise_mix_fwd wrt a: da[b, c, t] = dy[b, t] * n[c, t]  (d add = dy)
__global__ void noise_mix_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ n,
                                     float* __restrict__ da, int C8, int L, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % L);
  const size_t bc = gid / L;
  const int c = static_cast<int>(bc % C8);
  const size_t b = bc / C8;
  const float g = __ldg(dy + b * L + t);
  float fn[8], o[8];
  ld_global_nc_v8(n + (static_cast<size_t>(c) * L + t) * 8, fn);
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = g * fn[j];
  st_global_v8(da + gid * 8, o);
}

// gradient of ms_relu_avgpool2d_fwd: dx[b,c,t] = (x > 0) * dy[b, c/cw, t/tw] / (cw*tw)
__global__ void relu_avgpool2d_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                          float* __restrict__ dx, int C8, int L, int cw, int tw,
                                          size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = static_cast<int>(gid % L);
  const size_t bc = gid / L;
  const int c8 = static_cast<int>(bc % C8);
  const size_t b = bc / C8;
  const int F = C8 * 8 / cw, S = L / tw;
  const float inv = 1.f / static_cast<float>(cw * tw);
  float f[8], g[8];
  ld_global_nc_v8(x + gid * 8, f);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int fo = (c8 * 8 + j) / cw;
    const float d = __ldg(dy + ((b * (F / 8) + (fo >> 3)) * S + t / tw) * 8 + (fo & 7));
    g[j] = f[j] > 0.f ? d * inv : 0.f;
  }
  st_global_v8(dx + gid * 8, g);
}

// ------------------------------------------------------------ direct conv backward (NCL f32)
struct DirectBwdParams {
  const float* dy;   // (B, cout, lout)
  const float* y;    // forward output (LeakyReLU mask source) or null
  const float* x;    // (B, cin, lin)          [wgrad]
  const float* w;    // (cout, cin/groups, k)  [dgrad]
  float* dx;         // (B, cin, lin)          [dgrad]
  float* dw;         // (cout, cin/groups, k)  [wgrad] = sum over clips and time, fixed order
  float* dbias;      // (cout)                 [wgrad] or null
  int B, cin, cout, lin, lout, k, stride, pad, groups, leaky;
  float* part;          // [wgrad] per-block partial sums (workspace)
  unsigned int* ticket; // [wgrad] one counter per reduction group (workspace)
};

__device__ __forceinline__ float masked(float g, float y, int leaky) {
  return (leaky && !(y > 0.f)) ? 0.2f * g : g;
}

constexpr int kDgTile = 128;

// dx[b, ci, i] = sum_{co in group} sum_k w[co, cil, k] * dz[b, co, (i + pad - k)/stride]
__global__ void __launch_bounds__(kDgTile)
direct_dgrad_kernel(const DirectBwdParams p) {
  extern __shared__ float sm[];
  const int cin_g = p.cin / p.groups, cout_g = p.cout / p.groups;
  const int g = blockIdx.y, b = blockIdx.z;
  const int i0 = blockIdx.x * kDgTile;
  // dz rows l in [l0, l0 + nl): l = (i + pad - k)/stride for i in the tile
  int l0 = (i0 + p.pad - (p.k - 1));
  l0 = l0 < 0 ? -((-l0 + p.stride - 1) / p.stride) : l0 / p.stride;
  const int l1 = (i0 + kDgTile - 1 + p.pad) / p.stride;
  const int nl = l1 - l0 + 1;
  float* sw = sm;                               // [cout_g][cin_g][k]
  float* sz = sm + cout_g * cin_g * p.k;        // [cout_g][nl]
  for (int i = threadIdx.x; i < cout_g * cin_g * p.k; i += kDgTile)
    sw[i] = __ldg(p.w + static_cast<size_t>(g) * cout_g * cin_g * p.k + i);
  for (int i = threadIdx.x; i < cout_g * nl; i += kDgTile) {
    const int oc = i / nl, l = l0 + (i - oc * nl);
    float v = 0.f;
    if (l >= 0 && l < p.lout) {
      const size_t idx = (static_cast<size_t>(b) * p.cout + g * cout_g + oc) * p.lout + l;
      v = __ldg(p.dy + idx);
      if (p.y != nullptr) v = masked(v, __ldg(p.y + idx), p.leaky);
    }
    sz[i] = v;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= p.lin) return;
  const int ip = i + p.pad;
  const int kfirst = ip % p.stride;            // taps with (ip - k) % stride == 0
  for (int c = 0; c < cin_g; ++c) {
    float acc = 0.f;
    for (int oc = 0; oc < cout_g; ++oc) {
      const float* wr = sw + (oc * cin_g + c) * p.k;
      const float* zr = sz + oc * nl - l0;
      for (int k = kfirst; k < p.k; k += p.stride) {
        const int l = (ip - k) / p.stride;
        if (ip - k >= 0) acc = fmaf(wr[k], zr[l], acc);
      }
    }
    p.dx[(static_cast<size_t>(b) * p.cin + g * cin_g + c) * p.lin + i] = acc;
  }
}

constexpr int kWgTileL = 256;
constexpr int kWgDirectThreads = 256;

// dw[co, cil, k] += sum_b sum_l dz[b, co, l] * x[b, g*cin_g + cil, l*stride + k - pad]
// one CTA = one output channel x one tile of l x a set of clips; thread = (output o, lane group)
__global__ void __launch_bounds__(kWgDirectThreads)
direct_wgrad_kernel(const DirectBwdParams p) {
  extern __shared__ float sm[];
  const int cin_g = p.cin / p.groups, cout_g = p.cout / p.groups;
  const int co = blockIdx.x;
  const int g = co / cout_g;
  const int l0 = blockIdx.y * kWgTileL;
  const int nl = min(kWgTileL, p.lout - l0);
  const int win = (kWgTileL - 1) * p.stride + p.k;
  float* sz = sm;                 // [kWgTileL]
  float* sx = sm + kWgTileL;      // [cin_g][win]
  const int nout = cin_g * p.k;
  const int ngroups = max(1, kWgDirectThreads / nout);
  const int o = threadIdx.x % nout, lg = threadIdx.x / nout;
  const bool active = lg < ngroups;
  const int cil = o / p.k, k = o - cil * p.k;
  float acc = 0.f, bacc = 0.f;
  for (int b = blockIdx.z; b < p.B; b += gridDim.z) {
    __syncthreads();
    for (int i = threadIdx.x; i < nl; i += kWgDirectThreads) {
      const size_t idx = (static_cast<size_t>(b) * p.cout + co) * p.lout + l0 + i;
      float v = __ldg(p.dy + idx);
      if (p.y != nullptr) v = masked(v, __ldg(p.y + idx), p.leaky);
      sz[i] = v;
    }
    const int in0 = l0 * p.stride - p.pad;
    const int nwin = (nl - 1) * p.stride + p.k;
    for (int i = threadIdx.x; i < cin_g * nwin; i += kWgDirectThreads) {
      const int c = i / nwin, j = i - c * nwin;
      const int ti = in0 + j;
      sx[c * win + j] = (ti >= 0 && ti < p.lin)
          ? __ldg(p.x + (static_cast<size_t>(b) * p.cin + g * cin_g + c) * p.lin + ti) : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* xr = sx + cil * win + k;
      for (int l = lg; l < nl; l += ngroups) acc = fmaf(sz[l], xr[l * p.stride], acc);
      if (o == 0)
        for (int l = lg; l < nl; l += ngroups) bacc += sz[l];
    }
  }
  // combine the lane groups
  __syncthreads();
  float* red = sm;                // [ngroups][nout] <= 256 floats
  if (active) red[lg * nout + o] = acc;
  __syncthreads();
  // partial slot of this block: [co][slot][nout + 1]
  const int nslots = gridDim.y * gridDim.z;
  const int slot = blockIdx.y * gridDim.z + blockIdx.z;
  float* mine = p.part + (static_cast<size_t>(co) * nslots + slot) * (nout + 1);
  if (threadIdx.x < nout) {
    float v = 0.f;
    for (int q = 0; q < ngroups; ++q) v += red[q * nout + threadIdx.x];
    mine[threadIdx.x] = v;
  }
  __syncthreads();
  if (active && o == 0) red[lg] = bacc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int q = 0; q < ngroups; ++q) v += red[q];
    mine[nout] = v;
  }
  // the last block of this output channel adds the partials in slot order
  if (!det_last_block(p.ticket + co, nslots)) return;
  const float* grp = p.part + static_cast<size_t>(co) * nslots * (nout + 1);
  for (int i = threadIdx.x; i < nout; i += kWgDirectThreads)
    p.dw[static_cast<size_t>(co) * nout + i] = det_sum_strided(grp, nslots, nout + 1, i);
  if (p.dbias != nullptr && threadIdx.x == 0)
    p.dbias[co] = det_sum_strided(grp, nslots, nout + 1, nout);
}


// ---- register-tiled variants for the discriminators' shapes (COG in {4,16}, CIG in {1,4}) ----
// dgrad: one thread = one input time step x all CIG input channels of the group.  Taps visited
// are k = kfirst + s*j with output row l = lq - j (no integer division in the loop); the tap
// loop is the OUTER loop and the COG output channels are unrolled inside it: dz is staged as
// [l][oc] and the weights as [oc][k][c] (threads of a warp differ in k by < s: consecutive
// 16-byte words, no bank conflict), so one step reads COG/4 + COG 16-byte vectors for 4*COG
// FMAs (1.3 instructions per FMA instead of the 6 of the (oc, tap) loop order).
template <int COG, int CIG>
__global__ void __launch_bounds__(kDgTile)
direct_dgrad_tiled_kernel(const DirectBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y, b = blockIdx.z;
  const int i0 = blockIdx.x * kDgTile;
  const int s = p.stride;
  int l0 = (i0 + p.pad - (p.k - 1));
  l0 = l0 < 0 ? -((-l0 + s - 1) / s) : l0 / s;
  const int l1 = (i0 + kDgTile - 1 + p.pad) / s;
  const int nl = l1 - l0 + 1;
  float* sw = sm;                               // [COG][k][4]  (c padded to 4)
  float* sz = sm + p.k * COG * 4;               // [nl][COG]
  for (int i = threadIdx.x; i < p.k * COG * 4; i += kDgTile) {
    const int c = i & 3, k = (i >> 2) % p.k, oc = (i >> 2) / p.k;
    sw[i] = c < CIG ? __ldg(p.w + (static_cast<size_t>(g * COG + oc) * CIG + c) * p.k + k) : 0.f;
  }
  for (int i = threadIdx.x; i < COG * nl; i += kDgTile) {
    const int oc = i / nl, li = i - oc * nl, l = l0 + li;     // coalesced along l
    float v = 0.f;
    if (l >= 0 && l < p.lout) {
      const size_t idx = (static_cast<size_t>(b) * p.cout + g * COG + oc) * p.lout + l;
      v = __ldg(p.dy + idx);
      if (p.y != nullptr) v = masked(v, __ldg(p.y + idx), p.leaky);
    }
    sz[li * COG + oc] = v;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= p.lin) return;
  const int ip = i + p.pad;
  const int kfirst = ip % s;
  const int lq = (ip - kfirst) / s;             // output row of tap kfirst
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const float4* zrow = reinterpret_cast<const float4*>(sz + (lq - l0) * COG);
  const float4* wrow = reinterpret_cast<const float4*>(sw) + kfirst;
  int j = 0;
  for (int k = kfirst; k < p.k && j <= lq; k += s, ++j) {
    const float4* z4 = zrow - j * (COG / 4);            // dz[lq - j][0..COG)
    const float4* w4 = wrow + j * s;                    // w[oc][k][0..4) at w4[oc * K]
#pragma unroll
    for (int q = 0; q < COG / 4; ++q) {
      const float4 z = z4[q];
      const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 w = w4[(q * 4 + e) * p.k];
        acc[0] = fmaf(w.x, zz[e], acc[0]);
        acc[1] = fmaf(w.y, zz[e], acc[1]);
        acc[2] = fmaf(w.z, zz[e], acc[2]);
        acc[3] = fmaf(w.w, zz[e], acc[3]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CIG; ++c)
    p.dx[(static_cast<size_t>(b) * p.cin + g * CIG + c) * p.lin + i] = acc[c];
}

// wgrad: one CTA = one group x one tile of kWgTileL output rows x a set of clips; one thread =
// (input channel, 4 consecutive taps) x ALL COG output channels x every `lanes`-th row:
// 4*COG FMAs per 4 input reads + COG/4 broadcast 16-byte reads of dz ([l][oc] in smem).
template <int COG, int CIG>
__global__ void __launch_bounds__(kWgDirectThreads)
direct_wgrad_tiled_kernel(const DirectBwdParams p) {
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.x;
  const int s = p.stride;
  const int l0 = blockIdx.y * kWgTileL;
  const int nl = min(kWgTileL, p.lout - l0);
  const int kq = (p.k + 3) / 4;                          // tap quads
  const int win = (kWgTileL - 1) * s + 4 * kq;
  float* sz = sm;                                        // [kWgTileL][COG]
  float* sx = sm + kWgTileL * COG;                       // [CIG][win]
  const int nthr = CIG * kq;                             // threads per lane group
  const int lanes = kWgDirectThreads / nthr;             // row-interleaved lane groups
  const int o = threadIdx.x % nthr, lg = threadIdx.x / nthr;
  const bool active = lg < lanes;
  const int c = o / kq, k0 = (o - c * kq) * 4;
  float acc[4][COG];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int oc = 0; oc < COG; ++oc) acc[q][oc] = 0.f;
  float bacc[COG];
#pragma unroll
  for (int oc = 0; oc < COG; ++oc) bacc[oc] = 0.f;
  for (int b = blockIdx.z; b < p.B; b += gridDim.z) {
    __syncthreads();
    for (int i = threadIdx.x; i < nl * COG; i += kWgDirectThreads) {
      const int l = i % nl, oc = i / nl;
      const size_t idx = (static_cast<size_t>(b) * p.cout + g * COG + oc) * p.lout + l0 + l;
      float v = __ldg(p.dy + idx);
      if (p.y != nullptr) v = masked(v, __ldg(p.y + idx), p.leaky);
      sz[l * COG + oc] = v;
    }
    const int in0 = l0 * s - p.pad;
    const int nwin = (nl - 1) * s + 4 * kq;
    for (int i = threadIdx.x; i < CIG * nwin; i += kWgDirectThreads) {
      const int cc = i / nwin, j = i - cc * nwin;
      const int ti = in0 + j;
      sx[cc * win + j] = (ti >= 0 && ti < p.lin)
          ? __ldg(p.x + (static_cast<size_t>(b) * p.cin + g * CIG + cc) * p.lin + ti) : 0.f;
    }
    __syncthreads();
    if (active) {
      const float* xr = sx + c * win + k0;
      for (int l = lg; l < nl; l += lanes) {
        const float x0 = xr[l * s], x1 = xr[l * s + 1], x2 = xr[l * s + 2], x3 = xr[l * s + 3];
        const float4* z4 = reinterpret_cast<const float4*>(sz + l * COG);
#pragma unroll
        for (int q4 = 0; q4 < COG / 4; ++q4) {
          const float4 z = z4[q4];
          const float zz[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int oc = q4 * 4 + e;
            acc[0][oc] = fmaf(zz[e], x0, acc[0][oc]);
            acc[1][oc] = fmaf(zz[e], x1, acc[1][oc]);
            acc[2][oc] = fmaf(zz[e], x2, acc[2][oc]);
            acc[3][oc] = fmaf(zz[e], x3, acc[3][oc]);
            if (o == 0) bacc[oc] += zz[e];
          }
        }
      }
    }
  }
  // lane groups combine in shared memory, one after the other (fixed order):
  // result [COG][CIG][4*kq] (+ COG bias sums)
  __syncthreads();
  float* red = sm;                                       // COG * CIG * 4*kq floats (<= sz + sx)
  float* bred = red + COG * CIG * 4 * kq;
  for (int i = threadIdx.x; i < COG * CIG * 4 * kq + COG; i += kWgDirectThreads) red[i] = 0.f;
  __syncthreads();
  for (int turn = 0; turn < lanes; ++turn) {
    if (active && lg == turn) {
#pragma unroll
      for (int oc = 0; oc < COG; ++oc) {
#pragma unroll
        for (int q = 0; q < 4; ++q) red[(oc * CIG + c) * 4 * kq + k0 + q] += acc[q][oc];
        if (o == 0) bred[oc] += bacc[oc];
      }
    }
    __syncthreads();
  }
  // partial slot of this block: [g][slot][COG*CIG*k + COG]
  const int nel = COG * CIG * p.k + COG;
  const int nslots = gridDim.y * gridDim.z;
  const int slot = blockIdx.y * gridDim.z + blockIdx.z;
  float* mine = p.part + (static_cast<size_t>(g) * nslots + slot) * nel;
  for (int i = threadIdx.x; i < COG * CIG * p.k; i += kWgDirectThreads) {
    const int k = i % p.k, oc_c = i / p.k;
    mine[i] = red[oc_c * 4 * kq + k];
  }
  if (threadIdx.x < COG) mine[COG * CIG * p.k + threadIdx.x] = bred[threadIdx.x];
  // the last block of this group adds the partials in slot order
  if (!det_last_block(p.ticket + g, nslots)) return;
  const float* grp = p.part + static_cast<size_t>(g) * nslots * nel;
  for (int i = threadIdx.x; i < COG * CIG * p.k; i += kWgDirectThreads)
    p.dw[static_cast<size_t>(g) * COG * CIG * p.k + i] = det_sum_strided(grp, nslots, nel, i);
  if (p.dbias != nullptr && threadIdx.x < COG)
    p.dbias[g * COG + threadIdx.x] = det_sum_strided(grp, nslots, nel, COG * CIG * p.k + threadIdx.x);
}

template <int COG, int CIG>
static ms_status launch_dgrad_tiled(const DirectBwdParams& p, cudaStream_t st) {
  const int nl = (kDgTile + p.k) / p.stride + 3;
  const size_t smem = sizeof(float) * (static_cast<size_t>(COG) * p.k * 4 + static_cast<size_t>(COG) * nl);
  if (smem > 48 * 1024) return MS_ERR_INVALID;
  dim3 grid(ceil_div(p.lin, kDgTile), p.groups, p.B);
  direct_dgrad_tiled_kernel<COG, CIG><<<grid, kDgTile, smem, st>>>(p);
  return after_launch("direct_dgrad_tiled_kernel");
}

// grid of the direct weight-gradient kernels: x = reduction groups (tiled: conv groups, generic:
// output channels), y = time tiles, z = clip slices; `budget` caps the total block count
static dim3 direct_wgrad_grid(int ngroups, int lout, int batch, long long budget) {
  const int ltiles = ceil_div(lout, kWgTileL);
  long long z = budget / (static_cast<long long>(ngroups) * ltiles);
  if (z < 1) z = 1;
  if (z > batch) z = batch;
  return dim3(ngroups, ltiles, static_cast<unsigned>(z));
}

static bool direct_wgrad_is_tiled(int cout_g, int cin_g, int ksize, int stride, int groups) {
  if (groups > 65535) return false;
  if (!((cout_g == 16 && (cin_g == 4 || cin_g == 1)) || (cout_g == 4 && cin_g == 4))) return false;
  const int kq = (ksize + 3) / 4;
  if (cin_g * kq > kWgDirectThreads) return false;
  const int win = (kWgTileL - 1) * stride + 4 * kq;
  const size_t smem = sizeof(float) * (static_cast<size_t>(kWgTileL) * cout_g + static_cast<size_t>(cin_g) * win);
  return smem <= 48 * 1024 && cout_g * cin_g * 4 * kq + cout_g <= kWgTileL * cout_g + cin_g * win;
}

template <int COG, int CIG>
static ms_status launch_wgrad_tiled(const DirectBwdParams& p, cudaStream_t st) {
  const int kq = (p.k + 3) / 4;
  const int win = (kWgTileL - 1) * p.stride + 4 * kq;
  const size_t smem = sizeof(float) * (static_cast<size_t>(kWgTileL) * COG + static_cast<size_t>(CIG) * win);
  const dim3 grid = direct_wgrad_grid(p.groups, p.lout, p.B, 2048);
  if (grid.y > 65535) return MS_ERR_INVALID;
  direct_wgrad_tiled_kernel<COG, CIG><<<grid, kWgDirectThreads, smem, st>>>(p);
  return after_launch("direct_wgrad_tiled_kernel");
}

// --------------------------------------------------------- single-output-channel conv backward
// forward: y[b,l] = act(bias + sum_c sum_k w[c,k] x[b,c,l+k-pad]),  x BLK f32.
// dzm[b,l] = dy[b,l] * (tanh ? 1 - y^2 : 1)
__global__ void mono_dz_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                               float* __restrict__ dzm, size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  float g = __ldg(dy + i);
  if (y != nullptr) {
    const float v = __ldg(y + i);
    g *= (1.f - v * v);
  }
  dzm[i] = g;
}

// dx[b,c,i] = sum_k w[c,k] * dzm[b, i - k + pad]   -> BLK f32
__global__ void mono_dgrad_kernel(const float* __restrict__ dzm, const float* __restrict__ w,
                                  float* __restrict__ dx, int C8, int L, int ksize, int pad) {
  extern __shared__ float sw[];    // [ksize][8] of this channel group
  const int c8 = blockIdx.y, b = blockIdx.z;
  for (int i = threadIdx.x; i < 8 * ksize; i += blockDim.x) {
    const int k = i / 8, e = i - k * 8;
    sw[i] = __ldg(w + (c8 * 8 + e) * ksize + k);
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= L) return;
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = 0.f;
  const float* z = dzm + static_cast<size_t>(b) * L;
  for (int k = 0; k < ksize; ++k) {
    const int l = i - k + pad;
    if (l < 0 || l >= L) continue;
    const float g = __ldg(z + l);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(sw[k * 8 + j], g, f[j]);
  }
  st_global_v8(dx + ((static_cast<size_t>(b) * C8 + c8) * L + i) * 8, f);
}

constexpr int kMonoMaxK = 8;
// dw[c,k] += sum_b sum_l dzm[b,l] * x[b,c,l+k-pad];  dbias += sum dzm
__global__ void __launch_bounds__(256)
mono_wgrad_kernel(const float* __restrict__ dzm, const float* __restrict__ x,
                  float* __restrict__ dw, float* __restrict__ dbias, int B, int C8, int L,
                  int ksize, int pad, float* __restrict__ part, unsigned int* ticket) {
  const int c8 = blockIdx.y;
  float acc[kMonoMaxK][8];
#pragma unroll
  for (int k = 0; k < kMonoMaxK; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  float bacc = 0.f;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const float* z = dzm + static_cast<size_t>(b) * L;
    const float* xr = x + (static_cast<size_t>(b) * C8 + c8) * static_cast<size_t>(L) * 8;
    // thread owns input row i; it contributes to tap k with dzm[i - k + pad]
    for (int i = blockIdx.x * 256 + threadIdx.x; i < L; i += gridDim.x * 256) {
      float f[8];
      ld_global_nc_v8(xr + static_cast<size_t>(i) * 8, f);
#pragma unroll
      for (int k = 0; k < kMonoMaxK; ++k) {
        if (k >= ksize) break;
        const int l = i - k + pad;
        if (l < 0 || l >= L) continue;
        const float g = __ldg(z + l);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[k][j] = fmaf(g, f[j], acc[k][j]);
      }
      if (c8 == 0) bacc += __ldg(z + i);
    }
  }
  __shared__ float sh[8][kMonoMaxK * 8 + 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMonoMaxK; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float v = acc[k][j];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) sh[warp][k * 8 + j] = v;
    }
  }
  {
    float v = bacc;
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[warp][kMonoMaxK * 8] = v;
  }
  __syncthreads();
  // partial slot of this block: [c8][slot][ksize*8 + 1]
  const int nel = ksize * 8 + 1;
  const int nslots = gridDim.x * gridDim.z;
  const int slot = blockIdx.x * gridDim.z + blockIdx.z;
  float* mine = part + (static_cast<size_t>(c8) * nslots + slot) * nel;
  if (threadIdx.x < ksize * 8) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += sh[w][threadIdx.x];
    mine[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) {
    float v = 0.f;
    for (int w = 0; w < 8; ++w) v += sh[w][kMonoMaxK * 8];
    mine[ksize * 8] = v;
  }
  // the last block of this channel group adds the partials in slot order
  if (!det_last_block(ticket + c8, nslots)) return;
  const float* grp = part + static_cast<size_t>(c8) * nslots * nel;
  if (threadIdx.x < ksize * 8) {
    const int k = threadIdx.x / 8, j = threadIdx.x - k * 8;
    dw[(c8 * 8 + j) * ksize + k] = det_sum_strided(grp, nslots, nel, threadIdx.x);
  }
  if (c8 == 0 && dbias != nullptr && threadIdx.x == 0)
    dbias[0] = det_sum_strided(grp, nslots, nel, ksize * 8);
}

// ---------------------------------------------------------------- pooling / losses / Adam
__global__ void avg_pool_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, int lin,
                                    int lout, int k, int stride, int pad, int include_pad,
                                    size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ti = static_cast<int>(i % lin);
  const size_t bc = i / lin;
  float acc = 0.f;
  // windows t with t*stride - pad <= ti < t*stride - pad + k
  int t_hi = (ti + pad) / stride;
  if (t_hi > lout - 1) t_hi = lout - 1;
  for (int t = t_hi; t >= 0 && t * stride - pad + k > ti; --t) {
    float div = static_cast<float>(k);
    if (!include_pad) {
      const int a = max(t * stride - pad, 0), e = min(t * stride - pad + k, lin);
      div = static_cast<float>(max(e - a, 1));
    }
    acc += __ldg(dy + bc * lout + t) / div;
  }
  dx[i] = acc;
}

// gradients of the ms_reduce_fwd terms; *gscale (device scalar, may be null = 1) is the
// upstream gradient of the scalar loss
__global__ void reduce_bwd_kernel(int mode, const float* __restrict__ a,
                                  const float* __restrict__ b, size_t n, float scale,
                                  const float* __restrict__ gscale, float* __restrict__ da,
                                  float* __restrict__ db) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float s = scale * (gscale != nullptr ? __ldg(gscale) : 1.f);
  const float av = __ldg(a + i);
  const float bv = b != nullptr ? __ldg(b + i) : 0.f;
  float ga = 0.f, gb = 0.f;
  switch (mode) {
    case MS_RED_L1: {
      const float d = av - bv;
      ga = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
      gb = -ga;
      break;
    }
    case MS_RED_HINGE_D:
      ga = (1.f - av > 0.f) ? -1.f : 0.f;
      gb = (1.f + bv > 0.f) ? 1.f : 0.f;
      break;
    case MS_RED_HINGE_G: ga = -1.f; break;
    case MS_RED_LSQ_D: ga = av - 1.f; gb = bv; break;
    default: ga = av - 1.f; break;
  }
  if (da != nullptr) da[i] = ga * s;
  if (db != nullptr) db[i] = gb * s;
}

// torch.optim.Adam (no amsgrad, no weight decay) on a flat parameter buffer
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                            float* __restrict__ m, float* __restrict__ v, size_t n, float lr,
                            float beta1, float beta2, float eps, float bc1, float bc2_sqrt,
                            float grad_scale) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

// loss-scaled fp16 backward: g *= inv_scale in place, *flag |= 1 if any element is not finite
// (an overflowed fp16 gradient operand) -- the step is then skipped and the scale lowered
__global__ void grad_unscale_check_kernel(float* __restrict__ g, size_t n,
                                          const float* __restrict__ inv_scale, int* flag) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float v = g[i] * __ldg(inv_scale);
  g[i] = v;
  if (!isfinite(v)) *flag = 1;
}

// same, with the step count in device memory (CUDA-graph replays advance it on the device)
__global__ void adam_tick_kernel(int* step, const int* __restrict__ skip) {
  if (skip == nullptr || *skip == 0) *step += 1;
}

__global__ void adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g,
                                float* __restrict__ m, float* __restrict__ v, size_t n, float lr,
                                float beta1, float beta2, float eps, const int* __restrict__ step,
                                float grad_scale, const int* __restrict__ skip) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  if (skip != nullptr && *skip != 0) return;          // non-finite gradients: no update
  const float t = static_cast<float>(*step);
  const float bc1 = 1.f - powf(beta1, t);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, t));
  const float gi = g[i] * grad_scale;
  const float mi = beta1 * m[i] + (1.f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

}  // namespace msb

using namespace msb;

extern "C" {

size_t ms_blk_act_bwd_workspace_bytes(int batch, int channels, int len) {
  if (batch <= 0 || channels <= 0 || len <= 0) return 0;
  const int xt = ceil_div(len, act_bwd_rows_per_block(len));
  return kTicketBytes + sizeof(float) * static_cast<size_t>(batch) * xt * channels;
}

ms_status ms_blk_act_bwd(const float* dy32, const void* sign16, const float* ya32,
                         const float* yb32, void* dz16, float* dz32, float* dbias, int batch,
                         int channels, int len, int fmt, int s2d_stride, void* workspace,
                         size_t workspace_bytes, void* stream) {
  if (dy32 == nullptr || dz16 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      len <= 0)
    return MS_ERR_INVALID;
  if ((ya32 == nullptr) != (yb32 == nullptr)) return MS_ERR_INVALID;
  if (s2d_stride < 1 || len % s2d_stride != 0) return MS_ERR_INVALID;
  if (fmt != MS_F16 && fmt != MS_BF16) return MS_ERR_INVALID;
  const long long rows = static_cast<long long>(batch) * (channels / 8);
  if (rows > 65535) return MS_ERR_INVALID;
  if (dbias != nullptr && (workspace == nullptr ||
                           workspace_bytes < ms_blk_act_bwd_workspace_bytes(batch, channels, len)))
    return MS_ERR_WORKSPACE;
  const int rpb = act_bwd_rows_per_block(len);
  ActBwdParams p{dy32, static_cast<const uint16_t*>(sign16), ya32, yb32,
                 static_cast<uint16_t*>(dz16), dbias, dz32, channels / 8, len, fmt, s2d_stride,
                 rpb,
                 reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kTicketBytes),
                 static_cast<unsigned int*>(workspace)};
  dim3 grid(ceil_div(len, rpb), static_cast<unsigned>(rows));
  act_bwd_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("act_bwd_kernel");
}

int ms_convt_dgrad_taps(int ksize, int stride, int pad, int* first_shift) {
  if (stride < 1 || ksize < stride || ksize % stride != 0 || pad < 0) return -1;
  const int dmin = -((pad + stride - 1) / stride);
  const int dmax = (ksize - 1 - pad) / stride;
  if (dmax < dmin) return -1;
  if (first_shift != nullptr) *first_shift = dmin;
  return dmax - dmin + 1;
}

ms_status ms_weight_dgrad_view(const float* w, float* out, int kind, int cout, int cin, int ksize,
                               int stride, int pad, void* stream) {
  if (w == nullptr || out == nullptr || cout <= 0 || cin <= 0 || ksize <= 0) return MS_ERR_INVALID;
  size_t total;
  int ntaps = 0, dmin = 0;
  if (kind == MS_CONV) {
    total = static_cast<size_t>(cin) * cout * ksize;
  } else if (kind == MS_CONVT) {
    ntaps = ms_convt_dgrad_taps(ksize, stride, pad, &dmin);
    if (ntaps < 1) return MS_ERR_INVALID;
    total = static_cast<size_t>(cin) * stride * cout * ntaps;
  } else {
    return MS_ERR_INVALID;
  }
  weight_dgrad_view_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(w, out, kind, cout, cin, ksize,
                                                                  stride, pad, ntaps, dmin, total);
  return after_launch("weight_dgrad_view_kernel");
}

ms_status ms_blk16_convert(const void* src, void* dst, size_t elems, int src_fmt, int dst_fmt,
                           void* stream) {
  if (src == nullptr || dst == nullptr || elems % 8 != 0) return MS_ERR_INVALID;
  if (elems == 0) return MS_OK;
  const size_t nvec = elems / 8;
  blk16_convert_kernel<<<static_cast<unsigned>((nvec + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint4*>(src), static_cast<uint4*>(dst), nvec, src_fmt, dst_fmt);
  return after_launch("blk16_convert_kernel");
}

ms_status ms_pack_ncl_to_blk32(const float* x, float* y32, int batch, int channels, int len,
                               void* stream) {
  if (x == nullptr || y32 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  pack_ncl_to_blk32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                             static_cast<cudaStream_t>(stream)>>>(x, y32, len, total);
  return after_launch("pack_ncl_to_blk32_kernel");
}

ms_status ms_conv1d_direct_dgrad(const float* dy, const float* y, const float* w, float* dx,
                                 int batch, int cin, int cout, int lin, int ksize, int stride,
                                 int pad, int groups, int leaky, void* stream) {
  if (dy == nullptr || w == nullptr || dx == nullptr || batch <= 0 || cin <= 0 || cout <= 0 ||
      groups <= 0 || cin % groups != 0 || cout % groups != 0 || (leaky && y == nullptr))
    return MS_ERR_INVALID;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout <= 0) return MS_ERR_INVALID;
  DirectBwdParams p{dy, y, nullptr, w, dx, nullptr, nullptr, batch, cin, cout, lin, lout,
                    ksize, stride, pad, groups, leaky, nullptr, nullptr};
  const int cin_g = cin / groups, cout_g = cout / groups;
  if (groups > 65535 || batch > 65535) return MS_ERR_INVALID;
  {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ms_status ts = MS_ERR_INVALID;
    if (cout_g == 16 && cin_g == 4) ts = launch_dgrad_tiled<16, 4>(p, st);
    else if (cout_g == 16 && cin_g == 1) ts = launch_dgrad_tiled<16, 1>(p, st);
    else if (cout_g == 4 && cin_g == 4) ts = launch_dgrad_tiled<4, 4>(p, st);
    if (ts != MS_ERR_INVALID) return ts;
  }
  const int nl = (kDgTile + ksize) / stride + 3;
  const size_t smem = sizeof(float) * (static_cast<size_t>(cout_g) * cin_g * ksize +
                                       static_cast<size_t>(cout_g) * nl);
  if (smem > 200 * 1024 || groups > 65535 || batch > 65535) return MS_ERR_INVALID;
  if (smem > 48 * 1024) {
    static thread_local size_t attr_set = 0;
    if (smem > attr_set) {
      cudaError_t e = cudaFuncSetAttribute(direct_dgrad_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(direct_dgrad_kernel)");
      attr_set = smem;
    }
  }
  dim3 grid(ceil_div(lin, kDgTile), groups, batch);
  direct_dgrad_kernel<<<grid, kDgTile, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("direct_dgrad_kernel");
}

size_t ms_conv1d_direct_wgrad_workspace_bytes(int batch, int cin, int cout, int lin, int ksize,
                                              int stride, int pad, int groups) {
  if (batch <= 0 || cin <= 0 || cout <= 0 || groups <= 0 || cin % groups != 0 ||
      cout % groups != 0)
    return 0;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout <= 0) return 0;
  const int cin_g = cin / groups, cout_g = cout / groups;
  size_t blocks, nel;
  if (direct_wgrad_is_tiled(cout_g, cin_g, ksize, stride, groups)) {
    const dim3 g = direct_wgrad_grid(groups, lout, batch, 2048);
    blocks = static_cast<size_t>(g.x) * g.y * g.z;
    nel = static_cast<size_t>(cout_g) * cin_g * ksize + cout_g;
  } else {
    const dim3 g = direct_wgrad_grid(cout, lout, batch, 4096);
    blocks = static_cast<size_t>(g.x) * g.y * g.z;
    nel = static_cast<size_t>(cin_g) * ksize + 1;
  }
  return kTicketBytes + sizeof(float) * blocks * nel;
}

ms_status ms_conv1d_direct_wgrad(const float* dy, const float* y, const float* x, float* dw,
                                 float* dbias, int batch, int cin, int cout, int lin, int ksize,
                                 int stride, int pad, int groups, int leaky, void* workspace,
                                 size_t workspace_bytes, void* stream) {
  if (dy == nullptr || x == nullptr || dw == nullptr || batch <= 0 || cin <= 0 || cout <= 0 ||
      groups <= 0 || cin % groups != 0 || cout % groups != 0 || (leaky && y == nullptr))
    return MS_ERR_INVALID;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout <= 0) return MS_ERR_INVALID;
  const int cin_g = cin / groups, cout_g = cout / groups;
  if (cin_g * ksize > kWgDirectThreads) return MS_ERR_INVALID;
  const size_t need = ms_conv1d_direct_wgrad_workspace_bytes(batch, cin, cout, lin, ksize, stride,
                                                             pad, groups);
  if (workspace == nullptr || need == 0 || workspace_bytes < need) return MS_ERR_WORKSPACE;
  DirectBwdParams p{dy, y, x, nullptr, nullptr, dw, dbias, batch, cin, cout, lin, lout,
                    ksize, stride, pad, groups, leaky,
                    reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kTicketBytes),
                    static_cast<unsigned int*>(workspace)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (direct_wgrad_is_tiled(cout_g, cin_g, ksize, stride, groups)) {
    if (groups > kMaxTickets) return MS_ERR_INVALID;
    if (cout_g == 16 && cin_g == 4) return launch_wgrad_tiled<16, 4>(p, st);
    if (cout_g == 16 && cin_g == 1) return launch_wgrad_tiled<16, 1>(p, st);
    return launch_wgrad_tiled<4, 4>(p, st);
  }
  const int win = (kWgTileL - 1) * stride + ksize;
  const size_t smem = sizeof(float) * (kWgTileL + static_cast<size_t>(cin_g) * win);
  if (smem > 200 * 1024 || cout > kMaxTickets) return MS_ERR_INVALID;
  if (smem > 48 * 1024) {
    static thread_local size_t attr_set = 0;
    if (smem > attr_set) {
      cudaError_t e = cudaFuncSetAttribute(direct_wgrad_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(direct_wgrad_kernel)");
      attr_set = smem;
    }
  }
  const dim3 grid = direct_wgrad_grid(cout, lout, batch, 4096);
  if (grid.y > 65535) return MS_ERR_INVALID;
  direct_wgrad_kernel<<<grid, kWgDirectThreads, smem, st>>>(p);
  return after_launch("direct_wgrad_kernel");
}

static dim3 mono_wgrad_grid(int batch, int cin, int len) {
  int xb = ceil_div(len, 256 * 8);
  if (xb > 64) xb = 64;
  int zb = 2048 / (xb * (cin / 8));
  if (zb < 1) zb = 1;
  if (zb > batch) zb = batch;
  return dim3(xb, cin / 8, zb);
}

size_t ms_conv_to_mono_bwd_workspace_bytes(int batch, int cin, int len, int ksize) {
  if (batch <= 0 || cin <= 0 || cin % 8 != 0 || len <= 0 || ksize <= 0) return 0;
  const dim3 g = mono_wgrad_grid(batch, cin, len);
  return kTicketBytes + sizeof(float) * static_cast<size_t>(g.x) * g.y * g.z * (ksize * 8 + 1);
}

ms_status ms_conv_to_mono_bwd(const float* dy, const float* y_tanh, const float* x32,
                              const float* w, float* dzm, float* dx32, float* dw, float* dbias,
                              int batch, int cin, int len, int ksize, int pad, void* workspace,
                              size_t workspace_bytes, void* stream) {
  if (dy == nullptr || dzm == nullptr || w == nullptr || batch <= 0 || cin <= 0 ||
      cin % 8 != 0 || len <= 0 || ksize <= 0 || ksize > kMonoMaxK || batch > 65535)
    return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t total = static_cast<size_t>(batch) * len;
  mono_dz_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(dy, y_tanh, dzm, total);
  ms_status s = after_launch("mono_dz_kernel");
  if (s != MS_OK) return s;
  if (dx32 != nullptr) {
    dim3 grid(ceil_div(len, 256), cin / 8, batch);
    mono_dgrad_kernel<<<grid, 256, sizeof(float) * 8 * ksize, st>>>(dzm, w, dx32, cin / 8, len,
                                                                   ksize, pad);
    s = after_launch("mono_dgrad_kernel");
    if (s != MS_OK) return s;
  }
  if (dw != nullptr) {
    if (x32 == nullptr) return MS_ERR_INVALID;
    if (workspace == nullptr ||
        workspace_bytes < ms_conv_to_mono_bwd_workspace_bytes(batch, cin, len, ksize))
      return MS_ERR_WORKSPACE;
    const dim3 grid = mono_wgrad_grid(batch, cin, len);
    mono_wgrad_kernel<<<grid, 256, 0, st>>>(
        dzm, x32, dw, dbias, batch, cin / 8, len, ksize, pad,
        reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + kTicketBytes),
        static_cast<unsigned int*>(workspace));
    s = after_launch("mono_wgrad_kernel");
  }
  return s;
}

ms_status ms_avg_pool1d_bwd(const float* dy, float* dx, int batch_channels, int lin, int ksize,
                            int stride, int pad, int count_include_pad, void* stream) {
  if (dy == nullptr || dx == nullptr || batch_channels <= 0) return MS_ERR_INVALID;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout <= 0) return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch_channels) * lin;
  avg_pool_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(dy, dx, lin, lout, ksize, stride, pad,
                                                             count_include_pad, total);
  return after_launch("avg_pool_bwd_kernel");
}

ms_status ms_reduce_bwd(int mode, const float* a, const float* b, size_t n, float weight,
                        const float* grad_out, float* da, float* db, void* stream) {
  if (a == nullptr || n == 0 || mode < 0 || mode > 4 || (da == nullptr && db == nullptr))
    return MS_ERR_INVALID;
  if ((mode == MS_RED_L1 || mode == MS_RED_HINGE_D || mode == MS_RED_LSQ_D) && b == nullptr)
    return MS_ERR_INVALID;
  float scale = weight / static_cast<float>(n);
  if (mode == MS_RED_LSQ_D || mode == MS_RED_LSQ_G) scale = weight / static_cast<float>(n);
  reduce_bwd_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                      static_cast<cudaStream_t>(stream)>>>(mode, a, b, n, scale, grad_out, da, db);
  return after_launch("reduce_bwd_kernel");
}

ms_status ms_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                       size_t n, float lr, float beta1, float beta2, float eps, int step,
                       float grad_scale, void* stream) {
  if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr ||
      step < 1)
    return MS_ERR_INVALID;
  if (n == 0) return MS_OK;
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0,
                static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                     beta2, eps, static_cast<float>(bc1),
                                                     static_cast<float>(sqrt(bc2)), grad_scale);
  return after_launch("adam_kernel");
}

ms_status ms_adam_step_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                           size_t n, float lr, float beta1, float beta2, float eps, int* step_dev,
                           float grad_scale, const int* skip_flag, void* stream) {
  if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr ||
      step_dev == nullptr)
    return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  adam_tick_kernel<<<1, 1, 0, st>>>(step_dev, skip_flag);
  ms_status s = after_launch("adam_tick_kernel");
  if (s != MS_OK || n == 0) return s;
  adam_dev_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, step_dev, grad_scale, skip_flag);
  return after_launch("adam_dev_kernel");
}

ms_status ms_grad_unscale_check(float* grad, size_t n, const float* inv_scale_dev, int* flag_dev,
                                void* stream) {
  if (grad == nullptr || inv_scale_dev == nullptr || flag_dev == nullptr) return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(flag_dev, 0, sizeof(int), st);
  if (e != cudaSuccess) return check_cuda(e, "cudaMemsetAsync(flag)");
  if (n == 0) return MS_OK;
  grad_unscale_check_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(
      grad, n, inv_scale_dev, flag_dev);
  return after_launch("grad_unscale_check_kernel");
}

ms_status ms_diag_sum_bwd(const float* dy, float* dz32, int batch, int channels, int z_len,
                          int out_len, int nphase, int skew, void* stream) {
  if (dy == nullptr || dz32 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      z_len <= 0 || out_len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * z_len;
  diag_sum_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                        static_cast<cudaStream_t>(stream)>>>(dy, dz32, channels / 8, z_len, out_len,
                                                             nphase, skew, total);
  return after_launch("diag_sum_bwd_kernel");
}

ms_status ms_expand_mono_bwd(const float* de32, float* dx, int batch, int len, int exp_len,
                             int shift, void* stream) {
  if (de32 == nullptr || dx == nullptr || batch <= 0 || len <= 0 || exp_len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * len;
  expand_mono_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                           static_cast<cudaStream_t>(stream)>>>(de32, dx, len, exp_len, shift, total);
  return after_launch("expand_mono_bwd_kernel");
}

ms_status ms_depth_to_space_blk32(const float* dys32, float* dx32, int batch, int channels,
                                  int src_rows, int rows_valid, int row_offset, int out_rows,
                                  int len, int stride, void* stream) {
  if (dys32 == nullptr || dx32 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      src_rows <= 0 || out_rows <= 0 || stride < 1 || row_offset < 0 ||
      rows_valid + row_offset > src_rows)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * out_rows;
  depth_to_space_blk32_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                                static_cast<cudaStream_t>(stream)>>>(
      dys32, dx32, channels / 8, src_rows, rows_valid, row_offset, out_rows, len, stride, total);
  return after_launch("depth_to_space_blk32_kernel");
}

ms_status ms_blk_act_pad_bwd(const float* dy32, const void* sign16, float* dx32, int batch,
                             int channels, int len, int pad, int pad_mode, void* stream) {
  if (dy32 == nullptr || dx32 == nullptr || batch <= 0 || channels <= 0 || channels % 8 != 0 ||
      len <= 0 || pad < 0 || (pad_mode == 1 && pad >= len))
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  act_pad_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                       static_cast<cudaStream_t>(stream)>>>(
      dy32, static_cast<const uint4*>(sign16), dx32, len, pad, pad_mode, total);
  return after_launch("act_pad_bwd_kernel");
}

ms_status ms_noise_mix_bwd(const float* dy, const float* n32, float* da32, int batch, int channels,
                           int len, void* stream) {
  if (dy == nullptr || n32 == nullptr || da32 == nullptr || batch <= 0 || channels <= 0 ||
      channels % 8 != 0 || len <= 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  noise_mix_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                         static_cast<cudaStream_t>(stream)>>>(dy, n32, da32, channels / 8, len,
                                                              total);
  return after_launch("noise_mix_bwd_kernel");
}

ms_status ms_relu_avgpool2d_bwd(const float* dy32, const float* x32, float* dx32, int batch,
                                int channels, int len, int channel_window, int time_window,
                                void* stream) {
  if (dy32 == nullptr || x32 == nullptr || dx32 == nullptr || batch <= 0 || channels <= 0 ||
      len <= 0 || channel_window < 1 || time_window < 1 || channels % (8 * channel_window) != 0 ||
      len % time_window != 0)
    return MS_ERR_INVALID;
  const size_t total = static_cast<size_t>(batch) * (channels / 8) * len;
  relu_avgpool2d_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                              static_cast<cudaStream_t>(stream)>>>(
      dy32, x32, dx32, channels / 8, len, channel_window, time_window, total);
  return after_launch("relu_avgpool2d_bwd_kernel");
}

ms_status ms_weight_norm_bwd(const float* dw, const float* v, const float* g, float* dv, float* dg,
                             int rows, int cols, void* stream) {
  if (dw == nullptr || v == nullptr || g == nullptr || dv == nullptr || dg == nullptr ||
      rows <= 0 || cols <= 0)
    return MS_ERR_INVALID;
  weight_norm_bwd_kernel<<<rows, 256, 0, static_cast<cudaStream_t>(stream)>>>(dw, v, g, dv, dg, cols);
  return after_launch("weight_norm_bwd_kernel");
}

}  // extern "C"
