// CTA-pair variant of the implicit-GEMM convolution (see conv_gemm.cu for the math).
//
// A cluster of two CTAs (one SM pair) computes a 256-row x NT-column tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256):
//   - CTA rank r stages ITS OWN 128 rows of activations (A) and HALF of every weight
//     stage (rows [r*NT/2, (r+1)*NT/2) of the B tile): weight bytes moved L2->SMEM per
//     output row halve, and each tensor core reads only 4 KB (A) + NT/2*32 B (B) per MMA;
//   - the leader (rank 0) issues the MMAs for both CTAs; D rows 0-127 land in rank 0's
//     tensor memory, rows 128-255 in rank 1's; each CTA drains its own accumulator;
//   - accumulators are double-buffered (2*NT <= 512 columns) so the epilogue of tile i
//     overlaps the MMAs of tile i+1 even at NT = 256.
// Synchronisation across the pair:
//   full[s]      local: this CTA's bulk copies for stage s have landed
//   peer_full[s] leader only: rank 1's stage s has landed (rank 1's warp 1 relays it with a
//                remote mbarrier arrive)
//   empty[s], tmem_full[a]   tcgen05.commit ... multicast::cluster to both CTAs
//   tmem_empty[a] leader only: both epilogues drained accumulator a (rank 1 arrives remotely)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

#ifdef MSB_CONV_ABLATE
#define MSB_CABL(bit) ((p.debug & (bit)) != 0)
// trace of cluster 0's leader, local tile indices 4..7: slot = role base + (ti - 4) * 16 + k
#define MSB_CTRACE(base, k)                                                                  \
  do {                                                                                       \
    if (p.dbg != nullptr && blockIdx.x == 0 && ti >= 4 && ti < 8 && (threadIdx.x & 31) == 0) \
      p.dbg[(base) + (ti - 4) * 16 + (k)] = clock64();                                       \
  } while (0)
#else
#define MSB_CABL(bit) false
#define MSB_CTRACE(base, k) do { } while (0)
#endif

// (16 epilogue warps were tried: no gain -- the drain is bound by the store path, not by issue)
constexpr int kPairEpiSlices = kPairEpiWarps / 4;
constexpr int kPairThreads = 64 + 32 * kPairEpiWarps;

__device__ __forceinline__ uint32_t pack2p(float a, float b, int operand) {
  if (operand == MS_BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
conv_gemm_pair_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // [0..7] full, [8..15] empty, [16..23] peer_full, [24..25] tmem_full, [26..27] tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  float* s_bias = reinterpret_cast<float*>(smem + 512);             // [2][256]
  int2* s_tab = reinterpret_cast<int2*>(smem + 512 + 2048);   // [2][32]
  const uint32_t bar_base = smem_u32(bars);
  // [header][epilogue staging (ConvTranspose)][resident weights (wres)][stage ring]
  const uint32_t stage_out_base = smem_u32(smem + kSmemHeader);
  const uint32_t data_base = stage_out_base + static_cast<uint32_t>(p.out_stage);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto pfull_bar = [&](int s) { return bar_base + 8u * (16 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (24 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (26 + a); };
  // weight-resident mode: this CTA's resident slice landed / (leader) the peer's landed
  const uint32_t wres_full = bar_base + 8u * 28;
  const uint32_t wres_pfull = bar_base + 8u * 29;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = (rank == 0);
  const int cluster_id = blockIdx.x >> 1;
  const int nclusters = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(pfull_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      // one arrival per epilogue warp of both CTAs (per-thread arrivals meant 256 remote
      // mbarrier operations per tile from rank 1, serialised on the leader's barrier)
      mbar_init(tempty_bar(a), 2 * kPairEpiWarps);
    }
    mbar_init(wres_full, 1);
    mbar_init(wres_pfull, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(smem_u32(tmem_slot), p.tmem_cols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // both CTAs' barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks = p.KB >> 3;
  const int nth = p.NT >> 1;   // B rows staged by each CTA
  // Tile walk.  Streaming mode: tile = (b, mt, nt) with nt fastest, clusters interleaved.
  // Weight-resident mode: nt slowest and each cluster takes a CONTIGUOUS run of tiles, so a
  // cluster changes its n-tile (reloads its resident weight slice) at most nnt - 1 times.
  const int per_nt = p.total_tiles / p.nnt;
  const int run = p.wres ? (p.total_tiles + nclusters - 1) / nclusters : 0;
  const int tile_begin = p.wres ? cluster_id * run : cluster_id;
  const int tile_end = p.wres ? min(p.total_tiles, tile_begin + run) : p.total_tiles;
  const int tile_step = p.wres ? 1 : nclusters;
  auto decode = [&](int tile, int& nt_idx, int& mt, int& b) {
    int rest;
    if (p.wres) {
      nt_idx = tile / per_nt;
      rest = tile - nt_idx * per_nt;
    } else {
      nt_idx = tile % p.nnt;
      rest = tile / p.nnt;
    }
    mt = rest % p.mtiles;
    b = rest / p.mtiles;
  };
  const uint32_t wslice = p.wres ? static_cast<uint32_t>(p.nkb) * p.w_stage_bytes : 0u;
  const uint32_t ring_base = data_base + wslice;   // resident weights first, then the stage ring

  if (warp == 0) {
    // =============================== producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    int cur_nt = -1;
    int ti = -1;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      ++ti;
      int nt_idx, mt, b;
      decode(tile, nt_idx, mt, b);
      if (p.wres && nt_idx != cur_nt) {
        if (cur_nt >= 0) {
          // every MMA that reads the old slice has completed once the most recently filled
          // stage has been released (tcgen05.commit covers all earlier MMAs)
          const int last = stage == 0 ? p.stages - 1 : stage - 1;
          const uint32_t last_phase = stage == 0 ? phase ^ 1u : phase;
          mbar_wait(empty_bar(last), last_phase);
        }
        cur_nt = nt_idx;
        if (elect_one()) {
          mbar_arrive_expect_tx(wres_full, wslice);
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) +
                                (static_cast<size_t>(nt_idx) * 2 + rank) * wslice;
          for (int kb = 0; kb < p.nkb; ++kb)
            bulk_g2s(data_base + static_cast<uint32_t>(kb) * p.w_stage_bytes,
                     wsrc + static_cast<size_t>(kb) * p.w_stage_bytes, p.w_stage_bytes, wres_full);
        }
        __syncwarp();
      }
      const int r0 = mt * 256 + static_cast<int>(rank) * 128 + p.min_off;
      const int lo = r0 < 0 ? 0 : r0;
      const int hi = (r0 + p.RA) > p.lin ? p.lin : (r0 + p.RA);
      const int nrows = hi > lo ? hi - lo : 0;
      const bool ragged = (nrows != p.RA);
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (kb < 4) MSB_CTRACE(0, kb * 3);
        mbar_wait(empty_bar(stage), phase ^ 1u);
        if (kb < 4) MSB_CTRACE(0, kb * 3 + 1);
        const uint32_t sA = ring_base + static_cast<uint32_t>(stage) * p.stage_bytes;
        const uint32_t sW = sA + p.a_stage_bytes;
        if (ragged) {
          const int head = lo - r0;
          const int tail0 = head + nrows;
          const int nz = head + (p.RA - tail0);
          for (int i = lane; i < nz * chunks; i += 32) {
            const int c = i / nz;
            int r = i - c * nz;
            r = r < head ? r : tail0 + (r - head);
            st_shared_v4(sA + static_cast<uint32_t>(c * p.RA + r) * 16u, 0u, 0u, 0u, 0u);
          }
          fence_proxy_async_smem();
          __syncwarp();
        }
        // lane 0 announces the bytes (and streams the weight stage); the activation copies, one
        // per 16-byte channel chunk, are issued by as many lanes in parallel (one lane issuing
        // all of them took ~1000 cycles per stage: tools/pair_trace.py)
        if (lane == 0) {
          const uint32_t bytes_a = static_cast<uint32_t>(nrows) * 16u * chunks;
          mbar_arrive_expect_tx(full_bar(stage), bytes_a + (p.wres ? 0u : p.w_stage_bytes));
          if (!p.wres) {
            const uint8_t* wsrc =
                reinterpret_cast<const uint8_t*>(p.w) +
                ((static_cast<size_t>(nt_idx) * 2 + rank) * p.nkb + kb) * p.w_stage_bytes;
            bulk_g2s(sW, wsrc, p.w_stage_bytes, full_bar(stage));
          }
        }
        __syncwarp();
        if (nrows > 0) {
          const size_t cbase = static_cast<size_t>(b) * (p.xcin >> 3) + (kb % p.xnkb) * chunks;
          for (int c = lane; c < chunks; c += 32) {
            const uint16_t* src = p.x + ((cbase + c) * p.lin + lo) * 8;
            bulk_g2s(sA + static_cast<uint32_t>(c * p.RA + (lo - r0)) * 16u, src,
                     static_cast<uint32_t>(nrows) * 16u, full_bar(stage));
          }
        }
        __syncwarp();
        if (kb < 4) MSB_CTRACE(0, kb * 3 + 2);
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ============================ MMA issuer (rank 0) ============================
      const uint32_t idesc = umma_idesc_f16_m256(p.NT, p.operand);
      const uint64_t adesc0 = umma_desc_base_nosw(static_cast<uint32_t>(p.RA) * 16u, 128);
      const uint64_t bdesc0 = umma_desc_base_nosw(static_cast<uint32_t>(nth) * 16u, 128);
      const uint32_t a_kstep = static_cast<uint32_t>(2 * p.RA);
      const uint32_t b_kstep = static_cast<uint32_t>(2 * nth);
      const int nk16 = p.KB >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int cur_nt = -1;
      uint32_t wres_phase = 0;
      int ti = -1;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        ++ti;
        if (p.wres) {
          int nt_idx, mt, b;
          decode(tile, nt_idx, mt, b);
          if (nt_idx != cur_nt) {
            cur_nt = nt_idx;
            mbar_wait(wres_full, wres_phase);
            mbar_wait_cluster(wres_pfull, wres_phase);
            wres_phase ^= 1u;
          }
        }
        MSB_CTRACE(64, 0);
        mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        MSB_CTRACE(64, 1);
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.NT);
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (kb < 4) MSB_CTRACE(64, 2 + kb * 3);
          mbar_wait_cluster(pfull_bar(stage), phase);
          tc_fence_after();
          if (kb < 4) MSB_CTRACE(64, 3 + kb * 3);
          const uint32_t sA = ring_base + static_cast<uint32_t>(stage) * p.stage_bytes;
          const uint32_t sW = p.wres ? data_base + static_cast<uint32_t>(kb) * p.w_stage_bytes
                                     : sA + p.a_stage_bytes;
          if (elect_one()) {
            for (int t = 0; t < p.taps && !MSB_CABL(2); ++t) {
              uint64_t ad = adesc0 + ((sA >> 4) + static_cast<uint32_t>(p.off[t] - p.min_off));
              uint64_t bd = bdesc0 + ((sW >> 4) + static_cast<uint32_t>(t * chunks * nth));
              for (int k16 = 0; k16 < nk16; ++k16) {
                umma2_f16_ss(d_tmem, ad, bd, idesc, (kb | t | k16) != 0 ? 1u : 0u);
                ad += a_kstep;
                bd += b_kstep;
              }
            }
            umma2_commit_mc(empty_bar(stage));
            if (kb == p.nkb - 1) umma2_commit_mc(tfull_bar(acc));
          }
          __syncwarp();
          if (kb < 4) MSB_CTRACE(64, 4 + kb * 3);
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1u;
      }
    } else {
      // ====================== relay (rank 1): stage landed -> leader ======================
      int stage = 0;
      uint32_t phase = 0;
      int cur_nt = -1;
      uint32_t wres_phase = 0;
      for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
        if (p.wres) {
          int nt_idx, mt, b;
          decode(tile, nt_idx, mt, b);
          if (nt_idx != cur_nt) {
            cur_nt = nt_idx;
            mbar_wait(wres_full, wres_phase);
            wres_phase ^= 1u;
            if (elect_one()) mbar_arrive_remote(wres_pfull, 0);
            __syncwarp();
          }
        }
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(full_bar(stage), phase);
          if (elect_one()) mbar_arrive_remote(pfull_bar(stage), 0);
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // column slice of this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    const int cout8 = p.cout >> 3;
    const int ngroups = p.NT >> 4;
    int ti = -1;
    int tab_nt0 = -1, tab_nt1 = -1;     // n-tile whose tables sit in buffer 0 / 1
    // warp-uniform: the specialised epilogue below applies (bias is applied as v + b: alpha = 1)
    const bool fast = p.kind == MS_CONV && p.leaky == 1 && p.alpha == 1.0f && p.operand == MS_F16 &&
                      (p.NT & 31) == 0 && (p.debug & 512) == 0;
    for (int tile = tile_begin; tile < tile_end; tile += tile_step) {
      ++ti;
      int nt_idx, mt, b;
      decode(tile, nt_idx, mt, b);
      const int n0 = nt_idx * p.NT;
      if (warp == 2) MSB_CTRACE(128, 0);
      // bias / column tables of this tile's n-tile: rebuilt only when the n-tile of this
      // accumulator stage changes (for nnt = 1 layers once per CTA; the fill -- a global load
      // and a barrier of all epilogue warps -- cost 1.1-2.3 k cycles per tile on the epilogue's
      // critical path, tools/pair_trace.py)
      if (((acc & 1) ? tab_nt1 : tab_nt0) != nt_idx || (p.debug & 256) != 0) {
        // every epilogue warp is done with the tiles that read this buffer
        named_bar_sync(1, 32 * kPairEpiWarps);
        const int et = threadIdx.x - 64;
        float* tb = s_bias + (acc & 1) * 256;
        int2* tt = s_tab + (acc & 1) * 32;
        if (et < p.NT) {
          const int n = n0 + et;
          const int ch = (p.kind == MS_CONVT) ? convt_col_channel(n, p.stride) : n;
          tb[et] = p.bias != nullptr ? __ldg(p.bias + ch) : 0.f;
        }
        if (et < (p.NT >> 3)) {
          const int n = n0 + et * 8;
          if (p.kind == MS_CONVT) {
            tt[et] = make_int2(convt_col_phase(n, p.stride), convt_col_channel(n, p.stride));
          } else {
            tt[et] = make_int2(0, n);
          }
        }
        named_bar_sync(1, 32 * kPairEpiWarps);
        if (acc & 1) tab_nt1 = nt_idx; else tab_nt0 = nt_idx;
      }
      const float* tbias = s_bias + (acc & 1) * 256;
      const int2* ttab = s_tab + (acc & 1) * 32;
      const int m = mt * 256 + static_cast<int>(rank) * 128 + q * 32 + lane;
      if (p.res32 != nullptr && p.kind == MS_CONV && m < p.Lm) {
        // residual rows of this tile into L2 while the MMAs still run: the epilogue is
        // latency-bound on these loads (8 warps, one 32-byte vector per thread per chunk)
        for (int g = 2 * half; g < ngroups; g += 2 * kPairEpiSlices) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int cidx = g * 2 + h;
            if (cidx * 8 >= p.NT) break;
            const int ch = n0 + cidx * 8;
            prefetch_l2(p.res32 + ((static_cast<size_t>(b) * cout8 + (ch >> 3)) * p.Lout + m) * 8);
          }
        }
      }
      if (warp == 2) MSB_CTRACE(128, 1);
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      if (warp == 2) MSB_CTRACE(128, 2);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(acc * p.NT);
      bool arrived = false;
      for (int g = 2 * half; g < ngroups; g += 2 * kPairEpiSlices) {
        const bool two = (g + 1) < ngroups;
        // the residual vectors of this group's (up to four) chunks are requested BEFORE the
        // accumulator is read: four independent loads in flight per thread instead of one L2
        // round trip per chunk (the loads used to sit between the stores of consecutive chunks)
        float r8[4][8];
        if (p.res32 != nullptr && !MSB_CABL(4)) {
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h >= 2 && !two) break;
            const int2 rc = ttab[g * 2 + h];
            const int orow = (p.kind == MS_CONVT) ? p.stride * m + rc.x - p.pad : m;
            if ((m < p.Lm) && orow >= 0 && orow < p.Lout)
              ld_global_nc_v8(p.res32 + ((static_cast<size_t>(b) * cout8 + (rc.y >> 3)) * p.Lout + orow) * 8,
                              r8[h]);
          }
        }
        uint32_t v[32];
        if (MSB_CABL(8)) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = static_cast<uint32_t>(m + j);
        } else {
          tmem_ld16p(taddr + g * 16, &v[0]);
          if (two) tmem_ld16p(taddr + (g + 1) * 16, &v[16]);
          tmem_ld_wait();
        }
        if (g + 2 * kPairEpiSlices >= ngroups) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote_relaxed(tempty_bar(acc), 0);
          }
          arrived = true;
        }
        if (fast && two) {
          // ---- plain conv, fp16 operands, alpha = 1, LeakyReLU before the residual (every conv
          //      of the generator's C = 256 stage and its first conv): no column tables, packed
          //      fp32 math, one address per chunk.  The generic path below spends ~9 instructions
          //      per element on run-time flags; this one ~4.
          if (m < p.Lm && !MSB_CABL(16)) {
            const size_t row_base = static_cast<size_t>(b) * cout8 * p.Lout + m;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int cidx = g * 2 + h;
              const float4 b0 = *reinterpret_cast<const float4*>(tbias + cidx * 8);
              const float4 b1 = *reinterpret_cast<const float4*>(tbias + cidx * 8 + 4);
              float f[8];
              add_x2(__uint_as_float(v[h * 8 + 0]), __uint_as_float(v[h * 8 + 1]), b0.x, b0.y, f[0], f[1]);
              add_x2(__uint_as_float(v[h * 8 + 2]), __uint_as_float(v[h * 8 + 3]), b0.z, b0.w, f[2], f[3]);
              add_x2(__uint_as_float(v[h * 8 + 4]), __uint_as_float(v[h * 8 + 5]), b1.x, b1.y, f[4], f[5]);
              add_x2(__uint_as_float(v[h * 8 + 6]), __uint_as_float(v[h * 8 + 7]), b1.z, b1.w, f[6], f[7]);
#pragma unroll
              for (int j = 0; j < 8; j += 2) leaky02x2(f[j], f[j + 1], f[j], f[j + 1]);
              if (p.res32 != nullptr) {
#pragma unroll
                for (int j = 0; j < 8; j += 2) add_x2(f[j], f[j + 1], r8[h][j], r8[h][j + 1], f[j], f[j + 1]);
              }
              if (MSB_CABL(1)) continue;
              const size_t idx = row_base + static_cast<size_t>((n0 >> 3) + cidx) * p.Lout;
              if (p.y32 != nullptr) st_global_v8(p.y32 + idx * 8, f);
              if (p.y16 != nullptr)
                *reinterpret_cast<uint4*>(p.y16 + idx * 8) =
                    make_uint4(pack_h2(f[0], f[1]), pack_h2(f[2], f[3]), pack_h2(f[4], f[5]),
                               pack_h2(f[6], f[7]));
            }
          }
          continue;
        }
        if (p.out_stage != 0 && two) {
          // ---- ConvTranspose, staged: the four chunks are phases ph0 .. ph0+3 of one channel
          //      block = 128 contiguous fp32 output bytes of this thread's input-rate row
          const int2 rc0 = ttab[g * 2];
          const int cb = rc0.y >> 3, ph0 = rc0.x;
          float f[4][8];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const int cidx = g * 2 + h;
            const float4 b0 = *reinterpret_cast<const float4*>(tbias + cidx * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(tbias + cidx * 8 + 4);
            f[h][0] = fmaf(__uint_as_float(v[h * 8 + 0]), p.alpha, b0.x);
            f[h][1] = fmaf(__uint_as_float(v[h * 8 + 1]), p.alpha, b0.y);
            f[h][2] = fmaf(__uint_as_float(v[h * 8 + 2]), p.alpha, b0.z);
            f[h][3] = fmaf(__uint_as_float(v[h * 8 + 3]), p.alpha, b0.w);
            f[h][4] = fmaf(__uint_as_float(v[h * 8 + 4]), p.alpha, b1.x);
            f[h][5] = fmaf(__uint_as_float(v[h * 8 + 5]), p.alpha, b1.y);
            f[h][6] = fmaf(__uint_as_float(v[h * 8 + 6]), p.alpha, b1.z);
            f[h][7] = fmaf(__uint_as_float(v[h * 8 + 7]), p.alpha, b1.w);
            if (p.leaky == 1) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[h][j] = leaky02(f[h][j]);
            }
            if (p.res32 != nullptr) {
              const int orow = p.stride * m + ph0 + h - p.pad;
              if ((m < p.Lm) && orow >= 0 && orow < p.Lout) {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[h][j] += r8[h][j];
              }
            }
            if (p.leaky == 2) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[h][j] = leaky02(f[h][j]);
            }
          }
          if (MSB_CABL(1)) continue;
          const uint32_t sst = stage_out_base + static_cast<uint32_t>(warp - 2) * kPairStageWarp;
          const int m0 = m - lane;                       // row of lane 0
          const size_t plane = (static_cast<size_t>(b) * cout8 + cb) * p.Lout;
          if (p.y32 != nullptr) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const uint32_t a = sst + static_cast<uint32_t>(lane * kPairStageRow + h * 32);
              st_shared_v4(a, __float_as_uint(f[h][0]), __float_as_uint(f[h][1]),
                           __float_as_uint(f[h][2]), __float_as_uint(f[h][3]));
              st_shared_v4(a + 16, __float_as_uint(f[h][4]), __float_as_uint(f[h][5]),
                           __float_as_uint(f[h][6]), __float_as_uint(f[h][7]));
            }
            __syncwarp();
            // read back transposed: eight lanes cover one row's 128 bytes, one instruction
            // writes four whole lines
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = 4 * i + (lane >> 3), p8 = lane & 7;
              const int orow = p.stride * (m0 + r) + ph0 + (p8 >> 1) - p.pad;
              if ((m0 + r) < p.Lm && orow >= 0 && orow < p.Lout) {
                const uint4 qv = ld_shared_v4(sst + static_cast<uint32_t>(r * kPairStageRow + p8 * 16));
                *reinterpret_cast<uint4*>(p.y32 + (plane + orow) * 8 + (p8 & 1) * 4) = qv;
              }
            }
            __syncwarp();
          }
          if (p.y16 != nullptr) {
            // 16-bit image: 64 bytes per row, rows of 80 bytes in the staging tile
#pragma unroll
            for (int h = 0; h < 4; ++h)
              st_shared_v4(sst + static_cast<uint32_t>(lane * 80 + h * 16),
                           pack2p(f[h][0], f[h][1], p.operand), pack2p(f[h][2], f[h][3], p.operand),
                           pack2p(f[h][4], f[h][5], p.operand), pack2p(f[h][6], f[h][7], p.operand));
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int r = 8 * i + (lane >> 2), p4 = lane & 3;
              const int orow = p.stride * (m0 + r) + ph0 + p4 - p.pad;
              if ((m0 + r) < p.Lm && orow >= 0 && orow < p.Lout) {
                const uint4 qv = ld_shared_v4(sst + static_cast<uint32_t>(r * 80 + p4 * 16));
                *reinterpret_cast<uint4*>(p.y16 + (plane + orow) * 8) = qv;
              }
            }
            __syncwarp();
          }
          continue;
        }
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          if (h >= 2 && !two) break;
          const int cidx = g * 2 + h;
          const int2 rc = ttab[cidx];
          const int ch = rc.y;
          const int orow = (p.kind == MS_CONVT) ? p.stride * m + rc.x - p.pad : m;
          const bool valid = (m < p.Lm) && orow >= 0 && orow < p.Lout;
          if (!valid || MSB_CABL(16)) continue;
          float f[8];
          {
            const float4 b0 = *reinterpret_cast<const float4*>(tbias + cidx * 8);
            const float4 b1 = *reinterpret_cast<const float4*>(tbias + cidx * 8 + 4);
            f[0] = fmaf(__uint_as_float(v[h * 8 + 0]), p.alpha, b0.x);
            f[1] = fmaf(__uint_as_float(v[h * 8 + 1]), p.alpha, b0.y);
            f[2] = fmaf(__uint_as_float(v[h * 8 + 2]), p.alpha, b0.z);
            f[3] = fmaf(__uint_as_float(v[h * 8 + 3]), p.alpha, b0.w);
            f[4] = fmaf(__uint_as_float(v[h * 8 + 4]), p.alpha, b1.x);
            f[5] = fmaf(__uint_as_float(v[h * 8 + 5]), p.alpha, b1.y);
            f[6] = fmaf(__uint_as_float(v[h * 8 + 6]), p.alpha, b1.z);
            f[7] = fmaf(__uint_as_float(v[h * 8 + 7]), p.alpha, b1.w);
          }
          if (p.leaky == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = leaky02(f[j]);
          }
          const size_t idx = (static_cast<size_t>(b) * cout8 + (ch >> 3)) * p.Lout + orow;
          if (p.res32 != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += r8[h][j];
          }
          if (p.leaky == 2) {                         // activation AFTER the residual add
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = leaky02(f[j]);
          }
          if (MSB_CABL(1)) continue;
          if (p.y32 != nullptr) st_global_v8(p.y32 + idx * 8, f);
          if (p.y16 != nullptr) {
            uint4 o;
            o.x = pack2p(f[0], f[1], p.operand);
            o.y = pack2p(f[2], f[3], p.operand);
            o.z = pack2p(f[4], f[5], p.operand);
            o.w = pack2p(f[6], f[7], p.operand);
            *reinterpret_cast<uint4*>(p.y16 + idx * 8) = o;
          }
        }
      }
      if (!arrived) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(acc)); else mbar_arrive_remote_relaxed(tempty_bar(acc), 0);
        }
      }
      if (warp == 2) MSB_CTRACE(128, 3);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  // nobody leaves while the peer may still touch this CTA's SMEM / barriers / TMEM
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, p.tmem_cols);
  }
}

ms_status launch_conv_pair(const ConvGemmParams& p, size_t smem_bytes, cudaStream_t stream) {
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_pair_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_gemm_pair_kernel)");
    attr_set = true;
  }
  const int sms = sm_count();
  if (sms <= 1) return check_cuda(cudaGetLastError(), "sm_count");
  int clusters = sms / 2;
  if (p.total_tiles < clusters) clusters = p.total_tiles;
  conv_gemm_pair_kernel<<<2 * clusters, kPairThreads, smem_bytes, stream>>>(p);
  return after_launch("conv_gemm_pair_kernel");
}

}  // namespace msb
