// Debug-only microbenchmarks of the synchronisation / issue primitives the kernels are
// built from (not part of the public ABI; used by tools/microbench.py).
#include <cuda_runtime.h>

#include "../../music-synthesis_b200/csrc/ptx.cuh"
#include "../../music-synthesis_b200/csrc/runtime.cuh"

namespace msb {

// out[i] = cycles for primitive i (single warp, CTA 0)
__global__ void __launch_bounds__(128, 1) microbench_kernel(long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t bar0 = smem_u32(bars), bar1 = bar0 + 8, bar2 = bar0 + 16;
  const uint32_t sA = smem_u32(smem + 1024);          // 64 KB of A
  const uint32_t sB = sA + 65536;                     // 64 KB of B
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar1, 1); mbar_init(bar2, 1);
    fence_mbar_init();
  }
  // zero operands
  for (int i = threadIdx.x; i < 131072 / 16; i += blockDim.x)
    st_shared_v4(sA + i * 16, 0, 0, 0, 0);
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    int slot = 0;
    long long t0, t1;
    // (0) clock overhead
    t0 = clock64(); t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0; slot++;
    // (1) mbarrier arrive + try_wait on completed phase
    if (threadIdx.x == 0) mbar_arrive(bar0);
    __syncwarp();
    t0 = clock64();
    mbar_wait(bar0, 0);
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0; slot++;
    // (2) elect_one + syncwarp x4
    t0 = clock64();
    int acc = 0;
    for (int i = 0; i < 4; ++i) { if (elect_one()) acc++; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = (t1 - t0) / 4 + (acc > 100); slot++;
    // (3) tc_fence_after x4
    t0 = clock64();
    for (int i = 0; i < 4; ++i) tc_fence_after();
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = (t1 - t0) / 4; slot++;
    // (4) commit with nothing outstanding + wait
    t0 = clock64();
    if (elect_one()) umma_commit(bar1);
    __syncwarp();
    t1 = clock64();
    mbar_wait(bar1, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[slot] = t1 - t0; out[slot + 1] = t2 - t1; } slot += 2;
    // (6..) MMA issue / completion for N in {32,64,128,256}, 16 MMAs each, same accumulator
    uint32_t par1 = 1, par2 = 0;
    const int Ns[4] = {32, 64, 128, 256};
    for (int ni = 0; ni < 4; ++ni) {
      const int N = Ns[ni];
      const uint32_t idesc = umma_idesc_f16(N, 0);
      const uint64_t ad = umma_desc_base_nosw(256 * 16, 128) + (sA >> 4);
      const uint64_t bd = umma_desc_base_nosw(N * 16, 128) + (sB >> 4);
      for (int variant = 0; variant < 2; ++variant) {   // 0: same accumulator, 1: alternate 2
        __syncwarp();
        t0 = clock64();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            umma_f16_ss(tmem + (variant ? (k & 1) * 256 : 0), ad + (k & 7) * 512, bd + (k & 7) * 2 * N, idesc, 1u);
        }
        __syncwarp();
        t1 = clock64();
        if (elect_one()) umma_commit(bar1);
        __syncwarp();
        mbar_wait(bar1, par1); par1 ^= 1;
        t2 = clock64();
        if (threadIdx.x == 0) { out[slot] = (t1 - t0); out[slot + 1] = (t2 - t0); }
        slot += 2;
      }
    }
    // (22) 64 MMAs N=128 back-to-back, to see the steady state rate
    {
      const uint32_t idesc = umma_idesc_f16(128, 0);
      const uint64_t ad = umma_desc_base_nosw(256 * 16, 128) + (sA >> 4);
      const uint64_t bd = umma_desc_base_nosw(128 * 16, 128) + (sB >> 4);
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k)
          umma_f16_ss(tmem, ad + (k & 7) * 512, bd + (k & 7) * 256, idesc, 1u);
      }
      __syncwarp();
      t1 = clock64();
      if (elect_one()) umma_commit(bar1);
      __syncwarp();
      mbar_wait(bar1, par1); par1 ^= 1;
      t2 = clock64();
      if (threadIdx.x == 0) { out[slot] = (t1 - t0); out[slot + 1] = (t2 - t0); }
      slot += 2;
    }
    // (24) 64 MMAs N=32
    {
      const uint32_t idesc = umma_idesc_f16(32, 0);
      const uint64_t ad = umma_desc_base_nosw(256 * 16, 128) + (sA >> 4);
      const uint64_t bd = umma_desc_base_nosw(32 * 16, 128) + (sB >> 4);
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k)
          umma_f16_ss(tmem + (k & 3) * 64, ad + (k & 7) * 512, bd + (k & 7) * 64, idesc, 1u);
      }
      __syncwarp();
      t1 = clock64();
      if (elect_one()) umma_commit(bar1);
      __syncwarp();
      mbar_wait(bar1, par1); par1 ^= 1;
      t2 = clock64();
      if (threadIdx.x == 0) { out[slot] = (t1 - t0); out[slot + 1] = (t2 - t0); }
      slot += 2;
    }
    (void)par2; (void)bar2;
    // ---- epilogue primitives (slot 26..)
    slot = 26;
    uint32_t v[16];
    const uint32_t taddr = tmem;   // warp 0 -> lanes 0..31
    t0 = clock64();
    tmem_ld16(taddr, v);
    tmem_ld_wait();
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0 + (v[0] == 0x12345678u); slot++;     // 26: ld16+wait
    t0 = clock64();
    uint32_t v2[16], v3[16], v4[16];
    tmem_ld16(taddr, v); tmem_ld16(taddr + 16, v2); tmem_ld16(taddr + 32, v3); tmem_ld16(taddr + 48, v4);
    tmem_ld_wait();
    v[0] ^= v2[0] ^ v3[0] ^ v4[0];
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0 + (v[0] == 0x12345678u); slot++;     // 27: 4x ld16 + wait
    t0 = clock64();
    tmem_st16(taddr, v);
    tmem_st_wait();
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0; slot++;                              // 28: st16 + wait
    t0 = clock64();
    st_shared_v4(sA + threadIdx.x * 16, v[0], v[1], v[2], v[3]);
    st_shared_v4(sA + 4096 + threadIdx.x * 16, v[0], v[1], v[2], v[3]);
    t1 = clock64();
    fence_proxy_async_smem();
    t2 = clock64();
    if (threadIdx.x == 0) { out[slot] = t1 - t0; out[slot + 1] = t2 - t1; } slot += 2;  // 29: 2 STS, 30: fence.proxy.async
    slot += 3;
    t0 = clock64();
    named_bar_sync(2, 32);
    t1 = clock64();
    if (threadIdx.x == 0) out[slot] = t1 - t0; slot++;                              // 34: named barrier (1 warp)
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}


// TMEM read bandwidth: W warps (W in {4,8,16}; warp w reads lane quarter w%4) each issue `iters`
// tcgen05.ld of 32 lanes x NCOL fp32 columns back to back (one wait per 4 loads).
// out[0] = cycles from the CTA barrier before to the barrier after; bytes = W*iters*128*NCOL.
template <int NCOL>
__global__ void __launch_bounds__(512, 1) tmem_bw_kernel(long long* out, int warps, int iters) {
  __shared__ uint32_t tmem_slot;
  __shared__ long long tstart;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) { tmem_alloc(smem_u32(&tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  if (threadIdx.x == 0) tstart = clock64();
  __syncthreads();
  if (warp < warps) {
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t col = static_cast<uint32_t>(((i + u) * NCOL + (warp >> 2) * 64) & (511 & ~(NCOL - 1)));
        if (NCOL == 16) { uint32_t v[16]; tmem_ld16(base + col, v); acc ^= v[0] ^ v[15]; }
        else if (NCOL == 32) { uint32_t v[32]; tmem_ld32(base + col, v); acc ^= v[0] ^ v[31]; }
        else { uint32_t v[64]; tmem_ld64(base + col, v); acc ^= v[0] ^ v[63]; }
      }
      tmem_ld_wait();
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) { out[0] = clock64() - tstart; out[1] = acc == 0x12345u; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

}  // namespace msb

// Tensor-pipe time of 64 back-to-back M=128 MMAs as a function of N and of the A operand's start
// row (tap shift = descriptor start address + 16*shift bytes: is a core matrix that straddles two
// 128-byte lines fetched at full rate?).  out[0] = cycles issue -> completion.
namespace msb {
__global__ void __launch_bounds__(128, 1) mma_shift_kernel(long long* out, int N, int shift, int lbo_rows, int nacc, int run) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t bar1 = smem_u32(bars);
  const uint32_t sA = smem_u32(smem + 1024);
  const uint32_t sB = sA + 128 * 1024;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bar1, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  for (int i = threadIdx.x; i < (128 + 64) * 1024 / 16; i += blockDim.x)
    st_shared_v4(sA + i * 16, 0u, 0u, 0u, 0u);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (warp == 0) {
    const uint32_t idesc = umma_idesc_f16(N, 0);
    const uint64_t ad = umma_desc_base_nosw(lbo_rows * 16, 128) + ((sA >> 4) + shift);
    const uint64_t bd = umma_desc_base_nosw(N * 16, 128) + (sB >> 4);
    uint32_t par = 0;
    for (int rep = 0; rep < 2; ++rep) {
      __syncwarp();
      const long long t0 = clock64();
      if (elect_one()) {
#pragma unroll 8
        for (int k = 0; k < 64; ++k)
          umma_f16_ss(tmem + ((k >> run) & (nacc - 1)) * 256, ad + (k & 3) * 2 * lbo_rows, bd + (k & 3) * 2 * N, idesc, 1u);
        umma_commit(bar1);
      }
      __syncwarp();
      mbar_wait(bar1, par); par ^= 1;
      const long long t1 = clock64();
      if (threadIdx.x == 0) out[0] = t1 - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}
}  // namespace msb

extern "C" int ms_debug_mma_shift(long long* dev_out, int N, int shift, int lbo_rows, int nacc, int run, void* stream) {
  const int smem = 1024 + (128 + 64) * 1024;
  cudaFuncSetAttribute(msb::mma_shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  msb::mma_shift_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(dev_out, N, shift, lbo_rows, nacc, run);
  return cudaGetLastError() == cudaSuccess ? 0 : 1;
}

extern "C" int ms_debug_microbench(long long* dev_out, void* stream) {
  const int smem = 1024 + 131072;
  cudaFuncSetAttribute(msb::microbench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  msb::microbench_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(dev_out);
  return msb::after_launch("microbench_kernel");
}

extern "C" int ms_debug_tmem_bw(long long* dev_out, int ncol, int warps, int iters, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ncol == 16) msb::tmem_bw_kernel<16><<<1, 512, 0, st>>>(dev_out, warps, iters);
  else if (ncol == 32) msb::tmem_bw_kernel<32><<<1, 512, 0, st>>>(dev_out, warps, iters);
  else msb::tmem_bw_kernel<64><<<1, 512, 0, st>>>(dev_out, warps, iters);
  return msb::after_launch("tmem_bw_kernel");
}
