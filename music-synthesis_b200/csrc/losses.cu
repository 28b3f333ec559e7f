// GAN loss reductions (forward): deterministic two-stage sums (per-block partials in a
// fixed order, then one block) -- no atomics, bit-reproducible.
//   replaces featuresynth/loss/loss.py:5-79:
//     MS_RED_L1       mean |a - b|                      (F.l1_loss, feature matching)
//     MS_RED_HINGE_D  mean(relu(1 - a) + relu(1 + b))   hinge_discriminator_loss(r, f)
//     MS_RED_HINGE_G  mean(-a)                          hinge_generator_loss(f)
//     MS_RED_LSQ_D    0.5 * (mean((a-1)^2) + mean(b^2)) least_squares_disc_loss(r, f)
//     MS_RED_LSQ_G    0.5 * mean((a-1)^2)               least_squares_generator_loss(f)
// out[0] (+)= weight * value, so a whole mel_gan_*_loss sum accumulates in one scalar.
#include "runtime.cuh"

namespace msb {

constexpr int kRedThreads = 256;
constexpr int kRedMaxBlocks = 512;

__device__ __forceinline__ double red_term(int mode, float a, float b) {
  switch (mode) {
    case 0: return fabsf(a - b);
    case 1: return fmaxf(1.f - a, 0.f) + fmaxf(1.f + b, 0.f);
    case 2: return -a;
    case 3: return 0.5 * (static_cast<double>(a - 1.f) * (a - 1.f) + static_cast<double>(b) * b);
    default: return 0.5 * static_cast<double>(a - 1.f) * (a - 1.f);
  }
}

__global__ void __launch_bounds__(kRedThreads)
reduce_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n, int mode,
                      double* __restrict__ partial) {
  __shared__ double sh[kRedThreads];
  double acc = 0.0;
  for (size_t i = blockIdx.x * static_cast<size_t>(kRedThreads) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kRedThreads)
    acc += red_term(mode, __ldg(a + i), b != nullptr ? __ldg(b + i) : 0.f);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(kRedThreads)
reduce_final_kernel(const double* __restrict__ partial, int nblocks, double scale, float* out,
                    int accumulate) {
  __shared__ double sh[kRedThreads];
  double acc = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += kRedThreads) acc += partial[i];
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int s = kRedThreads / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float v = static_cast<float>(sh[0] * scale);
    out[0] = accumulate ? out[0] + v : v;
  }
}

}  // namespace msb

using namespace msb;

extern "C" {

size_t ms_reduce_workspace_bytes(void) { return sizeof(double) * kRedMaxBlocks; }

ms_status ms_reduce_fwd(int mode, const float* a, const float* b, size_t n, float weight,
                        float* out, int accumulate, void* workspace, void* stream) {
  if (a == nullptr || out == nullptr || workspace == nullptr || n == 0 || mode < 0 || mode > 4)
    return MS_ERR_INVALID;
  if ((mode == 0 || mode == 1 || mode == 3) && b == nullptr) return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t want = (n + kRedThreads * 8 - 1) / (kRedThreads * 8);
  const int blocks = static_cast<int>(want < 1 ? 1 : (want > kRedMaxBlocks ? kRedMaxBlocks : want));
  double* partial = static_cast<double*>(workspace);
  reduce_partial_kernel<<<blocks, kRedThreads, 0, st>>>(a, b, n, mode, partial);
  ms_status s = after_launch("reduce_partial_kernel");
  if (s != MS_OK) return s;
  reduce_final_kernel<<<1, kRedThreads, 0, st>>>(partial, blocks,
                                                 static_cast<double>(weight) / static_cast<double>(n),
                                                 out, accumulate);
  return after_launch("reduce_final_kernel");
}

}  // extern "C"
