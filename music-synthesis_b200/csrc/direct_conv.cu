// Grouped / strided 1-D convolution on CUDA cores (fp32), NCL layout, fused bias +
// LeakyReLU, and the discriminator's average pooling.
//   replaces F.conv1d(groups=g, stride=s) at featuresynth/discriminator/full.py:14-18
//   (Conv1d(1,16,15,1,7) and the four k=41 stride-4 grouped convs with 4 input channels
//   per group: K per output = 164 MACs, 4-16 outputs per group -- not tensor-core shaped)
//   and F.avg_pool1d(x, 4, 2, 2) at featuresynth/discriminator/melgan.py:22.
// One CTA = one (batch, group, 128-output tile): the group's input window is staged in
// shared memory once and reused by every output channel of the group; weights are read
// through the read-only path (warp-uniform addresses -> broadcast).
#include "runtime.cuh"

namespace msb {

constexpr int kDcTile = 128;

struct DirectConvParams {
  const float* x;      // (B, cin, lin)
  const float* w;      // (cout, cin/groups, k)
  const float* bias;   // (cout) or null
  float* y;            // (B, cout, lout)
  int B, cin, cout, lin, lout, k, stride, pad, groups, leaky, pad_mode;
};

__global__ void __launch_bounds__(kDcTile)
direct_conv_kernel(const DirectConvParams p) {
  extern __shared__ float sx[];   // [cin_g][win]
  const int cin_g = p.cin / p.groups, cout_g = p.cout / p.groups;
  const int g = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * kDcTile;
  const int win = (kDcTile - 1) * p.stride + p.k;
  const int in0 = t0 * p.stride - p.pad;
  for (int i = threadIdx.x; i < cin_g * win; i += kDcTile) {
    const int c = i / win, j = i - c * win;
    int ti = in0 + j;
    if (p.pad_mode == 1) {            // reflection padding (nn.ReflectionPad1d)
      if (ti < 0) ti = -ti;
      else if (ti >= p.lin) ti = 2 * (p.lin - 1) - ti;
    }
    sx[i] = (ti >= 0 && ti < p.lin)
                ? __ldg(p.x + (static_cast<size_t>(b) * p.cin + g * cin_g + c) * p.lin + ti)
                : 0.f;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= p.lout) return;
  const float* xw = sx + threadIdx.x * p.stride;
  for (int oc = 0; oc < cout_g; ++oc) {
    const int co = g * cout_g + oc;
    const float* wr = p.w + static_cast<size_t>(co) * cin_g * p.k;
    float a0 = 0.f, a1 = 0.f;
    for (int c = 0; c < cin_g; ++c) {
      const float* xc = xw + c * win;
      const float* wc = wr + c * p.k;
      int k = 0;
      for (; k + 1 < p.k; k += 2) {
        a0 = fmaf(xc[k], __ldg(wc + k), a0);
        a1 = fmaf(xc[k + 1], __ldg(wc + k + 1), a1);
      }
      if (k < p.k) a0 = fmaf(xc[k], __ldg(wc + k), a0);
    }
    float v = a0 + a1 + (p.bias != nullptr ? __ldg(p.bias + co) : 0.f);
    if (p.leaky) v = fmaxf(v, 0.2f * v);
    p.y[(static_cast<size_t>(b) * p.cout + co) * p.lout + t] = v;
  }
}


// Register-tiled variant for the shapes the discriminators actually use (COG = output channels
// per group in {4, 16}, CIG = input channels per group in {1, 4}): one thread = one output time
// step x ALL COG output channels of the group, so every staged input sample feeds COG FMAs and
// the weights are read as broadcast 16-byte vectors ([c][k][oc] in shared memory).  The input
// window is stored de-interleaved by stride phase so that the stride-s reads of neighbouring
// threads hit consecutive banks.
template <int COG, int CIG>
__global__ void __launch_bounds__(kDcTile)
direct_conv_tiled_kernel(const DirectConvParams p) {
  extern __shared__ __align__(16) float sm[];
  const int g = blockIdx.y, b = blockIdx.z;
  const int t0 = blockIdx.x * kDcTile;
  const int s = p.stride;
  const int win = (kDcTile - 1) * s + p.k;
  const int plen = (win + s - 1) / s;            // samples per phase
  float* sw = sm;                                // [CIG][k][COG]
  float* sx = sm + CIG * p.k * COG;              // [CIG][s][plen]
  for (int i = threadIdx.x; i < CIG * p.k * COG; i += kDcTile) {
    const int oc = i % COG, k = (i / COG) % p.k, c = i / (COG * p.k);
    sw[i] = __ldg(p.w + (static_cast<size_t>(g * COG + oc) * CIG + c) * p.k + k);
  }
  const int in0 = t0 * s - p.pad;
  for (int i = threadIdx.x; i < CIG * win; i += kDcTile) {
    const int c = i / win, j = i - c * win;
    int ti = in0 + j;
    if (p.pad_mode == 1) {
      if (ti < 0) ti = -ti;
      else if (ti >= p.lin) ti = 2 * (p.lin - 1) - ti;
    }
    const float v = (ti >= 0 && ti < p.lin)
        ? __ldg(p.x + (static_cast<size_t>(b) * p.cin + g * CIG + c) * p.lin + ti) : 0.f;
    sx[(c * s + (j % s)) * plen + j / s] = v;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= p.lout) return;
  float acc[COG];
#pragma unroll
  for (int o = 0; o < COG; ++o) acc[o] = 0.f;
  for (int c = 0; c < CIG; ++c) {
    for (int r = 0; r < s; ++r) {                 // taps k = r, r+s, ...: phase r of the window
      const float* xp = sx + (c * s + r) * plen + threadIdx.x;
      for (int k = r; k < p.k; k += s, ++xp) {
        const float xv = *xp;
        const float4* w4 = reinterpret_cast<const float4*>(sw + (c * p.k + k) * COG);
#pragma unroll
        for (int o = 0; o < COG / 4; ++o) {
          const float4 w = w4[o];
          acc[4 * o + 0] = fmaf(xv, w.x, acc[4 * o + 0]);
          acc[4 * o + 1] = fmaf(xv, w.y, acc[4 * o + 1]);
          acc[4 * o + 2] = fmaf(xv, w.z, acc[4 * o + 2]);
          acc[4 * o + 3] = fmaf(xv, w.w, acc[4 * o + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < COG; ++o) {
    const int co = g * COG + o;
    float v = acc[o] + (p.bias != nullptr ? __ldg(p.bias + co) : 0.f);
    if (p.leaky) v = fmaxf(v, 0.2f * v);
    p.y[(static_cast<size_t>(b) * p.cout + co) * p.lout + t] = v;
  }
}

template <int COG, int CIG>
static ms_status launch_direct_tiled(const DirectConvParams& p, cudaStream_t st) {
  const int win = (kDcTile - 1) * p.stride + p.k;
  const int plen = (win + p.stride - 1) / p.stride;
  const size_t smem = sizeof(float) * (static_cast<size_t>(CIG) * p.k * COG +
                                       static_cast<size_t>(CIG) * p.stride * plen);
  if (smem > 48 * 1024) return MS_ERR_INVALID;
  dim3 grid(ceil_div(p.lout, kDcTile), p.groups, p.B);
  direct_conv_tiled_kernel<COG, CIG><<<grid, kDcTile, smem, st>>>(p);
  return after_launch("direct_conv_tiled_kernel");
}

// y[b,c,t] = (1/k) * sum_j x[b,c,t*stride - pad + j]   (zero padding counted: the
// reference's count_include_pad=True default)
__global__ void avg_pool_kernel(const float* __restrict__ x, float* __restrict__ y, int lin,
                                int lout, int k, int stride, int pad, int include_pad,
                                size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int t = static_cast<int>(i % lout);
  const size_t bc = i / lout;
  const float* xr = x + bc * lin;
  float acc = 0.f;
  int cnt = 0;
  for (int j = 0; j < k; ++j) {
    const int ti = t * stride - pad + j;
    if (ti >= 0 && ti < lin) { acc += __ldg(xr + ti); ++cnt; }
  }
  y[i] = acc / static_cast<float>(include_pad ? k : (cnt > 0 ? cnt : 1));
}

}  // namespace msb

using namespace msb;

extern "C" {

int ms_conv1d_out_len(int lin, int ksize, int stride, int pad) {
  if (lin <= 0 || ksize <= 0 || stride <= 0 || pad < 0) return MS_ERR_INVALID;
  const int n = lin + 2 * pad - ksize;
  return n < 0 ? 0 : n / stride + 1;
}

ms_status ms_conv1d_direct_fwd(const float* x, const float* w, const float* bias, float* y,
                               int batch, int cin, int cout, int lin, int ksize, int stride,
                               int pad, int groups, int leaky, int pad_mode, void* stream) {
  if (x == nullptr || w == nullptr || y == nullptr || batch <= 0 || cin <= 0 || cout <= 0 ||
      groups <= 0 || cin % groups != 0 || cout % groups != 0)
    return MS_ERR_INVALID;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout < 0) return MS_ERR_INVALID;
  if (lout == 0) return MS_OK;
  if (pad_mode == 1 && pad >= lin) return MS_ERR_INVALID;
  DirectConvParams p{x, w, bias, y, batch, cin, cout, lin, lout, ksize, stride, pad, groups, leaky,
                     pad_mode};
  const int cin_g = cin / groups, cout_g = cout / groups;
  if (groups > 65535 || batch > 65535) return MS_ERR_INVALID;
  {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ms_status ts = MS_ERR_INVALID;
    if (cout_g == 16 && cin_g == 4) ts = launch_direct_tiled<16, 4>(p, st);
    else if (cout_g == 16 && cin_g == 1) ts = launch_direct_tiled<16, 1>(p, st);
    else if (cout_g == 4 && cin_g == 4) ts = launch_direct_tiled<4, 4>(p, st);
    if (ts != MS_ERR_INVALID) return ts;        // otherwise: generic kernel below
  }
  const size_t smem = sizeof(float) * cin_g * ((kDcTile - 1) * stride + ksize);
  if (smem > 200 * 1024) return MS_ERR_INVALID;
  if (smem > 48 * 1024) {
    static thread_local size_t attr_set = 0;
    if (smem > attr_set) {
      cudaError_t e = cudaFuncSetAttribute(direct_conv_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(direct_conv_kernel)");
      attr_set = smem;
    }
  }
  dim3 grid(ceil_div(lout, kDcTile), groups, batch);
  direct_conv_kernel<<<grid, kDcTile, smem, static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("direct_conv_kernel");
}

ms_status ms_avg_pool1d_fwd(const float* x, float* y, int batch_channels, int lin, int ksize,
                            int stride, int pad, int count_include_pad, void* stream) {
  if (x == nullptr || y == nullptr || batch_channels <= 0) return MS_ERR_INVALID;
  const int lout = ms_conv1d_out_len(lin, ksize, stride, pad);
  if (lout < 0) return MS_ERR_INVALID;
  if (lout == 0) return MS_OK;
  const size_t total = static_cast<size_t>(batch_channels) * lout;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  avg_pool_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, y, lin, lout, ksize, stride, pad, count_include_pad, total);
  return after_launch("avg_pool_kernel");
}

}  // extern "C"
