// Deterministic cross-block reductions for the backward pass (bias / weight gradients).
//
// The reference's gradients come from torch autograd (featuresynth/train/train.py:36,70), whose
// CUDA reductions are not bit-reproducible either -- but a data-parallel step must equal the
// whole-batch step to fp32 reduction-order tolerance and a graph replay must equal the eager
// step bit for bit, so nothing here adds floating-point numbers in arrival order:
//
//   1. every block writes its partial sums to a workspace slot that depends only on its block
//      index;
//   2. the blocks of a reduction group take a ticket (integer counter); the block that draws
//      the last ticket sums the group's partials IN INDEX ORDER and writes the result.
//
// The ticket is the only atomic and it orders nothing numerically.  It wraps back to zero on
// the last arrival (atomicInc), so the counters need zeroing once, when the workspace is
// allocated; launches on one stream reuse them (CUDA-graph replays included).
//
// Workspace layout (every entry point that takes a `workspace`): [kTicketBytes of counters]
// [partials].  The caller zero-fills the first kTicketBytes ONCE.
#pragma once
#include <stdint.h>

#include "../../include/msb200.h"

namespace msb {

constexpr size_t kTicketBytes = MS_TICKET_BYTES;       // 16384 reduction groups per launch
constexpr int kMaxTickets = static_cast<int>(kTicketBytes / sizeof(unsigned int));

// Called by ALL threads of a block after they wrote the block's partials.  Block-uniform
// result: true in the block that arrived last among the `nblocks` blocks sharing `ticket`;
// that block's subsequent loads see every other block's partials.
__device__ __forceinline__ bool det_last_block(unsigned int* ticket, unsigned int nblocks) {
  __shared__ unsigned int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(ticket, nblocks - 1u) == nblocks - 1u) ? 1u : 0u;
  __syncthreads();
  const bool last = s_last != 0u;
  if (last) __threadfence();
  return last;
}

// sum_{p < np} part[p * stride + i] in a fixed order (four interleaved chains for memory-level
// parallelism, combined as (a0 + a1) + (a2 + a3)); volatile-free plain loads: the caller is the
// last block (see above)
__device__ __forceinline__ float det_sum_strided(const float* part, int np, size_t stride,
                                                 size_t i) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int p = 0;
  for (; p + 4 <= np; p += 4) {
    a0 += __ldcg(part + static_cast<size_t>(p) * stride + i);
    a1 += __ldcg(part + static_cast<size_t>(p + 1) * stride + i);
    a2 += __ldcg(part + static_cast<size_t>(p + 2) * stride + i);
    a3 += __ldcg(part + static_cast<size_t>(p + 3) * stride + i);
  }
  for (; p < np; ++p) a0 += __ldcg(part + static_cast<size_t>(p) * stride + i);
  return (a0 + a1) + (a2 + a3);
}

}  // namespace msb
