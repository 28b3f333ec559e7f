#include "runtime.cuh"

#include <atomic>
#include <cstdio>
#include <cstring>

namespace msb {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_cuda_error(cudaError_t e, const char* where) {
  snprintf(g_err, sizeof(g_err), "%s: %s (%s)", where, cudaGetErrorString(e),
           cudaGetErrorName(e));
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int sm_count() {
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  int v = cached[dev].load(std::memory_order_relaxed);
  if (v > 0) return v;
  if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
    return -1;
  cached[dev].store(v, std::memory_order_relaxed);
  return v;
}

}  // namespace msb

extern "C" {

int ms_version(void) { return 100; }

const char* ms_strerror(ms_status s) {
  switch (s) {
    case MS_OK: return "ok";
    case MS_ERR_INVALID: return "invalid descriptor or unsupported shape";
    case MS_ERR_CUDA: return "CUDA error (see ms_last_cuda_error)";
    case MS_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

const char* ms_last_cuda_error(void) { return msb::g_err; }

uint64_t ms_launch_count(void) { return msb::g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
