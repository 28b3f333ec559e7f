// Library-wide helpers: status codes, last-error text, launch counter.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/msb200.h"

namespace msb {

void set_cuda_error(cudaError_t e, const char* where);
void count_launch();

// Call right after a kernel launch.
inline ms_status after_launch(const char* where) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_cuda_error(e, where);
    return MS_ERR_CUDA;
  }
  return MS_OK;
}

inline ms_status check_cuda(cudaError_t e, const char* where) {
  if (e != cudaSuccess) {
    set_cuda_error(e, where);
    return MS_ERR_CUDA;
  }
  return MS_OK;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();  // SMs of the current device (cached per device), <=0 on error

}  // namespace msb
