// Training-batch assembly on the GPU: aligned random crops of the resident audio and log-mel
// stores.  Replaces the per-example Python loop of batch_stream / random_slice
// (featuresynth/data/datastore.py:8-80: slice [start, start+size) of the anchor feature, the
// aligned slice [start*ratio, end*ratio) of every other feature, zero padding when the chunk is
// shorter) -- the crop positions are still drawn on the host (same distribution as the
// reference), only the data movement happens here: one launch per feature, coalesced reads of
// the store, coalesced writes of the (B, channels, len) batch.  HBM-bound:
// 2 * 4 * B * channels * len bytes.
#include <cstdint>

#include "runtime.cuh"

namespace msb {

struct CropArgs {
  const float* store;
  const long long* plan;   // (B, 3): origin, pitch, valid
  float* out;
  int channels, len;
};

// grid (ceil(channels * len / 256), B): (channel, time) flattened so that short crops of many
// channels (32 log-mel frames x 128 bins) still fill their CTAs
__global__ void __launch_bounds__(256) gather_crops_kernel(const CropArgs a) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.channels * a.len) return;
  const int c = i / a.len;
  const int t = i - c * a.len;
  const long long origin = __ldg(a.plan + 3 * b);
  const long long pitch = __ldg(a.plan + 3 * b + 1);
  const long long valid = __ldg(a.plan + 3 * b + 2);
  const float v = t < valid ? __ldg(a.store + origin + c * pitch + t) : 0.f;
  a.out[(static_cast<size_t>(b) * a.channels + c) * a.len + t] = v;
}

// len % 4 == 0: four consecutive time steps per thread -- four independent loads in flight (the
// crop origin is arbitrary, so the reads stay scalar) and one 16-byte store; the index split is
// done in units of 4 samples.  (One element per thread reached 33 % of the HBM copy rate.)
__global__ void __launch_bounds__(256) gather_crops4_kernel(const CropArgs a) {
  const int b = blockIdx.y;
  const int len4 = a.len >> 2;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= a.channels * len4) return;
  const int c = i / len4;
  const int t = (i - c * len4) << 2;
  const long long origin = __ldg(a.plan + 3 * b);
  const long long pitch = __ldg(a.plan + 3 * b + 1);
  const long long valid = __ldg(a.plan + 3 * b + 2);
  const float* src = a.store + origin + c * pitch + t;
  float4 v;
  v.x = t + 0 < valid ? __ldg(src + 0) : 0.f;
  v.y = t + 1 < valid ? __ldg(src + 1) : 0.f;
  v.z = t + 2 < valid ? __ldg(src + 2) : 0.f;
  v.w = t + 3 < valid ? __ldg(src + 3) : 0.f;
  *reinterpret_cast<float4*>(a.out + (static_cast<size_t>(b) * a.channels + c) * a.len + t) = v;
}

}  // namespace msb

using namespace msb;

extern "C" {

ms_status ms_gather_crops(const float* store, const long long* plan, float* out, int batch,
                          int channels, int len, void* stream) {
  if (store == nullptr || plan == nullptr || out == nullptr || batch <= 0 || channels <= 0 ||
      len <= 0 || batch > 65535 || static_cast<long long>(channels) * len > 0x7fffffffLL)
    return MS_ERR_INVALID;
  CropArgs a;
  a.store = store; a.plan = plan; a.out = out; a.channels = channels; a.len = len;
  if ((len & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0) {
    const dim3 grid4((channels * (len >> 2) + 255) / 256, batch);
    gather_crops4_kernel<<<grid4, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return after_launch("gather_crops4_kernel");
  }
  const dim3 grid((channels * len + 255) / 256, batch);
  gather_crops_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  return after_launch("gather_crops_kernel");
}

}  // extern "C"
