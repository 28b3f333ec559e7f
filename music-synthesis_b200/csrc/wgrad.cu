// Weight gradients of the dense 1-D (dilated / transposed) convolutions as a tcgen05 GEMM
// whose reduction dimension is TIME:
//
//   G[t][m][n] = sum_b sum_l  A[b, m, l] * X[b, n, l + shift_t]
//
// A (BLK 16-bit, B x Cm/8 x La x 8) is the output-side operand (dz for a Conv1d, x for a
// ConvTranspose1d), X (BLK 16-bit, B x Cn/8 x Lx x 8) the input-side one.  Both stay in the
// channel-blocked layout of the forward path: 8 consecutive time steps x 8 channels are 128
// contiguous bytes, which is exactly a SWIZZLE_NONE core matrix of an **MN-major** UMMA
// operand (LBO = 128 B between 8-row K groups, SBO = rows*16 B between 8-channel groups).
// A tap shift is again just +16 bytes per row on the descriptor start address, so one
// staged X tile serves every tap; each tap accumulates into its own TMEM column range.
//
// The K = B*La reduction is split across CTAs (one CTA per SM-sized slice); partial tiles
// go to a workspace and wgrad_reduce_kernel sums them in a fixed order (deterministic) and
// scatters into the reference weight layout.
//
// Backward of F.conv1d / F.conv_transpose1d at featuresynth/generator/full.py:24-43,
// util/modules.py:358-388 and discriminator/full.py:19 (autograd does this in the reference:
// train/train.py:36,70).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "conv_gemm.cuh"
#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

constexpr int kWgRK = 128;            // time steps per pipeline stage
constexpr int kWgThreads = 64 + 128;  // producer, MMA, 4 epilogue warps
constexpr int kWgHeader = 1024;

struct WgradParams {
  const uint16_t* a;
  const uint16_t* x;
  float* part;          // [ksplit][taps][Cm][Cn]
  int B, Cm, Cn, La, Lx;
  int taps;
  int shift[kMaxTaps];
  int min_shift, RX;    // RX = kWgRK + max_shift - min_shift rows of X per stage
  int NT, nnt, mblks;
  int chunks_per_clip, total_chunks, chunks_per_split;
  int fold_slots, fold_stride;   // >0: short clips, several per K chunk (slot = La + span rows)
  int stages, a_stage_bytes, x_stage_bytes, stage_bytes, tmem_cols;
  uint32_t idesc;
};

// MN-major SWIZZLE_NONE descriptor: LBO = byte distance between 8-row K groups (128: rows are
// contiguous), SBO = byte distance between 8-channel groups (rows * 16)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t rows) {
  return umma_desc_nosw(0, 128, rows * 16u);
}

__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);   // [0..7] full, [8..15] empty, [16] done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t data_base = smem_u32(smem + kWgHeader);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  const uint32_t done_bar = bar_base + 8u * 16;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nt_idx = blockIdx.x % p.nnt;
  const int mb = blockIdx.x / p.nnt;
  const int split = blockIdx.y;
  const int c_begin = split * p.chunks_per_split;
  int c_end = c_begin + p.chunks_per_split;
  if (c_end > p.total_chunks) c_end = p.total_chunks;
  const int a_groups = min(16, (p.Cm >> 3) - mb * 16);   // real channel groups in this M block
  const int x_groups = p.NT >> 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  // M rows beyond Cm (Cm < 128): those A channel groups are never loaded -- keep them zero
  if (a_groups < 16) {
    for (int s = 0; s < p.stages; ++s) {
      const uint32_t sA = data_base + static_cast<uint32_t>(s) * p.stage_bytes;
      const int n16 = (16 - a_groups) * kWgRK;
      for (int i = threadIdx.x; i < n16; i += kWgThreads)
        st_shared_v4(sA + static_cast<uint32_t>(a_groups * kWgRK + i) * 16u, 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  if (p.fold_slots > 0) {
    for (int s = 0; s < p.stages; ++s) {
      const uint32_t sA = data_base + static_cast<uint32_t>(s) * p.stage_bytes;
      for (int i = threadIdx.x; i < p.stage_bytes / 16; i += kWgThreads)
        st_shared_v4(sA + static_cast<uint32_t>(i) * 16u, 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && p.fold_slots > 0) {
    // ====================== producer, several short clips per chunk ======================
    // slot j of a chunk holds clip ch*slots + j: A row j*S + l <-> dz row l, X row j*S + a <->
    // input row min_shift + a; everything else in the stage stays zero (cleared once above;
    // the copy pattern is the same for every chunk except a last one with fewer clips)
    int stage = 0;
    uint32_t phase = 0;
    const int S = p.fold_stride;
    const int x_lo = p.min_shift < 0 ? -p.min_shift : 0;
    int x_hi = p.Lx - p.min_shift;
    if (x_hi > S) x_hi = S;
    const int x_rows = x_hi > x_lo ? x_hi - x_lo : 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      const int b0 = ch * p.fold_slots;
      const int nclips = min(p.fold_slots, p.B - b0);
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sA = data_base + static_cast<uint32_t>(stage) * p.stage_bytes;
      const uint32_t sX = sA + p.a_stage_bytes;
      if (nclips < p.fold_slots) {
        // fewer clips than slots: the unused slots may hold rows of an earlier chunk
        const int r_first = nclips * S;
        const int nz = kWgRK - r_first;
        for (int i = lane; i < nz * a_groups; i += 32) {
          const int g = i / nz, r = r_first + (i - g * nz);
          st_shared_v4(sA + static_cast<uint32_t>(g * kWgRK + r) * 16u, 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0) {
        const uint32_t bytes = (static_cast<uint32_t>(p.La) * 16u * a_groups +
                                static_cast<uint32_t>(x_rows) * 16u * x_groups) * nclips;
        mbar_arrive_expect_tx(full_bar(stage), bytes);
      }
      __syncwarp();
      // many small copies (one per clip and channel group): all 32 lanes issue them
      for (int i = lane; i < nclips * a_groups; i += 32) {
        const int j = i / a_groups, g = i - j * a_groups;
        const size_t abase = static_cast<size_t>(b0 + j) * (p.Cm >> 3) + mb * 16;
        bulk_g2s(sA + static_cast<uint32_t>(g * kWgRK + j * S) * 16u,
                 p.a + ((abase + g) * p.La) * 8, static_cast<uint32_t>(p.La) * 16u,
                 full_bar(stage));
      }
      if (x_rows > 0) {
        for (int i = lane; i < nclips * x_groups; i += 32) {
          const int j = i / x_groups, g = i - j * x_groups;
          const size_t xbase = static_cast<size_t>(b0 + j) * (p.Cn >> 3) + nt_idx * x_groups;
          bulk_g2s(sX + static_cast<uint32_t>(g * p.RX + j * S + x_lo) * 16u,
                   p.x + ((xbase + g) * p.Lx + (p.min_shift + x_lo)) * 8,
                   static_cast<uint32_t>(x_rows) * 16u, full_bar(stage));
        }
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 0) {
    // =============================== producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      const int b = ch / p.chunks_per_clip;
      const int r0 = (ch - b * p.chunks_per_clip) * kWgRK;
      const int a_rows = min(kWgRK, p.La - r0);
      const int x0 = r0 + p.min_shift;
      const int xlo = x0 < 0 ? 0 : x0;
      const int xhi = (x0 + p.RX) > p.Lx ? p.Lx : (x0 + p.RX);
      const int x_rows = xhi > xlo ? xhi - xlo : 0;
      mbar_wait(empty_bar(stage), phase ^ 1u);
      const uint32_t sA = data_base + static_cast<uint32_t>(stage) * p.stage_bytes;
      const uint32_t sX = sA + p.a_stage_bytes;
      bool filled = false;
      if (a_rows < kWgRK) {
        const int nz = kWgRK - a_rows;
        for (int i = lane; i < nz * a_groups; i += 32) {
          const int g = i / nz, r = a_rows + (i - g * nz);
          st_shared_v4(sA + static_cast<uint32_t>(g * kWgRK + r) * 16u, 0u, 0u, 0u, 0u);
        }
        filled = true;
      }
      if (x_rows < p.RX) {
        const int head = xlo - x0;
        const int tail0 = head + x_rows;
        const int nz = head + (p.RX - tail0);
        for (int i = lane; i < nz * x_groups; i += 32) {
          const int g = i / nz;
          int r = i - g * nz;
          r = r < head ? r : tail0 + (r - head);
          st_shared_v4(sX + static_cast<uint32_t>(g * p.RX + r) * 16u, 0u, 0u, 0u, 0u);
        }
        filled = true;
      }
      if (filled) {
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (elect_one()) {
        const uint32_t bytes = static_cast<uint32_t>(a_rows) * 16u * a_groups +
                               static_cast<uint32_t>(x_rows) * 16u * x_groups;
        mbar_arrive_expect_tx(full_bar(stage), bytes);
        const size_t abase = static_cast<size_t>(b) * (p.Cm >> 3) + mb * 16;
        for (int g = 0; g < a_groups; ++g)
          bulk_g2s(sA + static_cast<uint32_t>(g * kWgRK) * 16u,
                   p.a + ((abase + g) * p.La + r0) * 8, static_cast<uint32_t>(a_rows) * 16u,
                   full_bar(stage));
        if (x_rows > 0) {
          const size_t xbase = static_cast<size_t>(b) * (p.Cn >> 3) + nt_idx * x_groups;
          for (int g = 0; g < x_groups; ++g)
            bulk_g2s(sX + static_cast<uint32_t>(g * p.RX + (xlo - x0)) * 16u,
                     p.x + ((xbase + g) * p.Lx + xlo) * 8, static_cast<uint32_t>(x_rows) * 16u,
                     full_bar(stage));
        }
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint64_t adesc0 = umma_desc_mn(kWgRK);
    const uint64_t xdesc0 = umma_desc_mn(static_cast<uint32_t>(p.RX));
    int stage = 0;
    uint32_t phase = 0;
    for (int ch = c_begin; ch < c_end; ++ch) {
      mbar_wait(full_bar(stage), phase);
      tc_fence_after();
      const uint32_t sA = data_base + static_cast<uint32_t>(stage) * p.stage_bytes;
      const uint32_t sX = sA + p.a_stage_bytes;
      if (elect_one()) {
        for (int t = 0; t < p.taps; ++t) {
          uint64_t ad = adesc0 + (sA >> 4);
          uint64_t xd = xdesc0 + ((sX >> 4) + static_cast<uint32_t>(p.shift[t] - p.min_shift));
          const uint32_t dst = tmem_base + static_cast<uint32_t>(t * p.NT);
          for (int k16 = 0; k16 < kWgRK / 16; ++k16) {
            umma_f16_ss(dst, ad, xd, p.idesc, (ch != c_begin || k16 != 0) ? 1u : 0u);
            ad += 16;   // 16 rows x 16 bytes, in 16-byte units
            xd += 16;
          }
        }
        umma_commit(empty_bar(stage));
        if (ch == c_end - 1) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == p.stages) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (c_end > c_begin) {
    // =============================== epilogue ===============================
    const int q = warp & 3;
    const int m = mb * 128 + q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    const int n0 = nt_idx * p.NT;
    for (int t = 0; t < p.taps; ++t) {
      float* dst = p.part + ((static_cast<size_t>(split) * p.taps + t) * p.Cm + m) * p.Cn + n0;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                             static_cast<uint32_t>(t * p.NT);
      for (int c = 0; c < p.NT; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (m < p.Cm) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(dst + c + j) =
                make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                            __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// out[map(t, m, n)] = beta * out[...] + sum_split sum_f part[split][t][m][n + f*Cn/fold]
// (fixed order).  fold = 2: the X operand carried a two-term (hi, lo) split of the layer input in
// its two channel halves -- both halves are gradients of the same weight.
//   mode MS_CONV : Conv1d weight (Cout = Cm, Cin = Cn/fold, K = taps):   (m*Cin + n)*K + t
//   mode MS_CONVT: ConvTranspose1d weight (Cin = Cm, Cout, K = ksize); n = r*Cout + co,
//                  k = s*shift_t + r + pad (taps whose k falls outside [0, K) do not exist)
// Block = one output row m x 32 consecutive n x 8 split-lanes.  Each thread owns (m, n), ALL
// taps and every 8th split: reads are 128-byte rows, the (up to 148-long) split loop becomes 8
// independent chains, the lanes combine in shared memory.  MS_CONV: the block's 32*taps outputs
// are contiguous in the (Cout, Cin, K) weight -> coalesced writes.  MS_CONVT: scattered element
// writes (small tensors).
constexpr int kRedLanes = 8;
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, float* __restrict__ out, int nsplit, int taps,
                    int Cm, int Cn, int fold, int mode, int stride, int pad, int cout, int ksize,
                    int shift0, float beta, float alpha,
                    int sl /* split lanes per row: 1, 2, 4 or 8 */) {
  __shared__ float sh[kRedLanes][32 * kMaxTaps + 1];
  const int cn_out = Cn / fold;
  const int nx = threadIdx.x & 31, wl = threadIdx.x >> 5;
  const int rows = kRedLanes / sl;              // output rows m per block
  const int row = wl / sl, lane = wl - row * sl;
  const int n0 = blockIdx.x * 32;
  const int n = n0 + nx;
  const int m = blockIdx.y * rows + row;
  float acc[kMaxTaps];
#pragma unroll
  for (int t = 0; t < kMaxTaps; ++t) acc[t] = 0.f;
  if (n < cn_out && m < Cm) {
    const size_t plane = static_cast<size_t>(taps) * Cm * Cn;
    for (int s = lane; s < nsplit; s += sl) {
      const float* base = part + s * plane + static_cast<size_t>(m) * Cn + n;
#pragma unroll
      for (int t = 0; t < kMaxTaps; ++t) {
        if (t >= taps) break;
        for (int f = 0; f < fold; ++f)
          acc[t] += __ldg(base + static_cast<size_t>(t) * Cm * Cn + f * cn_out);
      }
    }
  }
#pragma unroll
  for (int t = 0; t < kMaxTaps; ++t)
    if (t < taps) sh[wl][nx * taps + t] = acc[t];
  __syncthreads();
  const int cnt = min(32, cn_out - n0) * taps;
  for (int ii = threadIdx.x; ii < cnt * rows; ii += 256) {
    const int rr = ii / cnt, i = ii - rr * cnt;
    const int m = blockIdx.y * rows + rr;
    if (m >= Cm) break;
    float v = 0.f;
    for (int l = 0; l < sl; ++l) v += sh[rr * sl + l][i];
    size_t o;
    if (mode == MS_CONV) {
      o = (static_cast<size_t>(m) * cn_out + n0) * taps + i;
    } else {
      const int nn = n0 + i / taps, t = i % taps;
      const int r = nn / cout, co = nn - r * cout;
      const int k = stride * (shift0 + t) + r + pad;
      if (k < 0 || k >= ksize) continue;
      o = (static_cast<size_t>(m) * cout + co) * ksize + k;
    }
    out[o] = (beta != 0.f ? beta * out[o] : 0.f) + alpha * v;
  }
}

struct WgradCfg {
  int NT, nnt, mblks, RX, min_shift, chunks_per_clip, total_chunks, ksplit, chunks_per_split;
  int fold_slots, fold_stride;
  int stages, a_stage_bytes, x_stage_bytes, stage_bytes, tmem_cols;
  size_t smem_bytes, workspace_bytes;
};

static bool make_wgrad_cfg(int B, int Cm, int Cn, int La, int Lx, int taps, const int* shifts,
                           WgradCfg* c) {
  if (B <= 0 || Cm <= 0 || Cn <= 0 || La <= 0 || Lx <= 0) return false;
  if (Cm % 8 != 0 || Cn % 16 != 0 || taps < 1 || taps > kMaxTaps) return false;
  int mn = shifts[0], mx = shifts[0];
  for (int t = 1; t < taps; ++t) {
    mn = shifts[t] < mn ? shifts[t] : mn;
    mx = shifts[t] > mx ? shifts[t] : mx;
  }
  c->min_shift = mn;
  c->RX = kWgRK + mx - mn;
  int nt = (512 / taps) / 16 * 16;
  if (nt > 256) nt = 256;
  while (nt >= 16 && Cn % nt != 0) nt -= 16;
  if (nt < 16) return false;
  const int budget = kSmemBudget - kWgHeader;
  c->a_stage_bytes = 16 * kWgRK * 16;
  auto xbytes = [&](int n) { return (n / 8) * c->RX * 16; };
  while (nt > 16 && (c->a_stage_bytes + xbytes(nt)) * 2 > budget) {
    nt -= 16;
    while (nt >= 16 && Cn % nt != 0) nt -= 16;
  }
  if (nt < 16 || (c->a_stage_bytes + xbytes(nt)) * 2 > budget) return false;
  c->NT = nt;
  c->nnt = Cn / nt;
  c->mblks = (Cm + 127) / 128;
  c->x_stage_bytes = xbytes(nt);
  c->stage_bytes = (c->a_stage_bytes + c->x_stage_bytes + 127) / 128 * 128;
  int s = budget / c->stage_bytes;
  if (s > 4) s = 4;
  c->stages = s;
  int cols = 32;
  while (cols < taps * nt) cols *= 2;
  if (cols > 512) return false;
  c->tmem_cols = cols;
  c->chunks_per_clip = (La + kWgRK - 1) / kWgRK;
  c->fold_slots = c->fold_stride = 0;
  long long total = static_cast<long long>(B) * c->chunks_per_clip;
  if (B > 1 && 2 * La + (mx - mn) <= kWgRK) {
    c->fold_stride = La + (mx - mn);
    c->fold_slots = (kWgRK - La) / c->fold_stride + 1;
    if (c->fold_slots > B) c->fold_slots = B;
    total = (B + c->fold_slots - 1) / c->fold_slots;
  }
  if (total > 0x7fffffffLL) return false;
  c->total_chunks = static_cast<int>(total);
  int sms = sm_count();
  if (sms <= 0) sms = 148;
  int ks = sms / (c->nnt * c->mblks);
  if (ks < 1) ks = 1;
  // every split writes a full (taps x 128 x NT) fp32 partial tile and the reduce kernel reads
  // them all back: with 2 chunks per split (the cfg4 layers: K = B*L of a few thousand rows over
  // 148 CTAs) that traffic cost 3x the GEMM itself (wgrad_reduce_kernel: 15.7 % of the training
  // cycle).  At least kMinChunksPerSplit chunks of MMA work per partial tile.
  constexpr int kMinChunksPerSplit = 8;
  const int ks_work = (c->total_chunks + kMinChunksPerSplit - 1) / kMinChunksPerSplit;
  if (ks > ks_work) ks = ks_work;
  if (ks > c->total_chunks) ks = c->total_chunks;
  c->chunks_per_split = (c->total_chunks + ks - 1) / ks;
  c->ksplit = (c->total_chunks + c->chunks_per_split - 1) / c->chunks_per_split;
  c->smem_bytes = kWgHeader + static_cast<size_t>(c->stages) * c->stage_bytes;
  if (c->smem_bytes < 120 * 1024) c->smem_bytes = 120 * 1024;   // one CTA per SM (TMEM)
  c->workspace_bytes = static_cast<size_t>(c->ksplit) * taps * Cm * Cn * sizeof(float);
  return true;
}

}  // namespace msb

using namespace msb;

extern "C" {

size_t ms_wgrad_workspace_bytes(int batch, int cm, int cn, int la, int lx, int taps,
                                const int* shifts) {
  WgradCfg c;
  if (shifts == nullptr || !make_wgrad_cfg(batch, cm, cn, la, lx, taps, shifts, &c)) return 0;
  return c.workspace_bytes;
}

ms_status ms_wgrad_fwd(const void* a16, const void* x16, int batch, int cm, int cn, int la,
                       int lx, int taps, const int* shifts, int fmt, int mode, int stride,
                       int pad, int cout, int ksize, int fold, float alpha, float beta, float* dw,
                       void* workspace, size_t workspace_bytes, void* stream) {
  WgradCfg c;
  if (a16 == nullptr || x16 == nullptr || dw == nullptr || shifts == nullptr ||
      !make_wgrad_cfg(batch, cm, cn, la, lx, taps, shifts, &c))
    return MS_ERR_INVALID;
  if (fmt != MS_F16 && fmt != MS_BF16) return MS_ERR_INVALID;
  if (mode != MS_CONV && mode != MS_CONVT) return MS_ERR_INVALID;
  if (fold != 1 && !(fold == 2 && mode == MS_CONV && cn % 2 == 0)) return MS_ERR_INVALID;
  if (mode == MS_CONVT) {
    if (stride < 1 || cout < 1 || cn != stride * cout || ksize < stride) return MS_ERR_INVALID;
    for (int t = 1; t < taps; ++t)
      if (shifts[t] != shifts[0] + t) return MS_ERR_INVALID;
  }
  if (workspace == nullptr || workspace_bytes < c.workspace_bytes) return MS_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p;
  p.a = static_cast<const uint16_t*>(a16);
  p.x = static_cast<const uint16_t*>(x16);
  p.part = static_cast<float*>(workspace);
  p.B = batch; p.Cm = cm; p.Cn = cn; p.La = la; p.Lx = lx;
  p.taps = taps;
  for (int t = 0; t < kMaxTaps; ++t) p.shift[t] = t < taps ? shifts[t] : 0;
  p.min_shift = c.min_shift; p.RX = c.RX;
  p.NT = c.NT; p.nnt = c.nnt; p.mblks = c.mblks;
  p.chunks_per_clip = c.chunks_per_clip; p.total_chunks = c.total_chunks;
  p.chunks_per_split = c.chunks_per_split;
  p.fold_slots = c.fold_slots; p.fold_stride = c.fold_stride;
  p.stages = c.stages; p.a_stage_bytes = c.a_stage_bytes; p.x_stage_bytes = c.x_stage_bytes;
  p.stage_bytes = c.stage_bytes; p.tmem_cols = c.tmem_cols;
  // D = F32, A/B format (must be the same), both operands MN-major (bits 15, 16), N, M = 128
  p.idesc = (1u << 4) | (static_cast<uint32_t>(fmt) << 7) | (static_cast<uint32_t>(fmt) << 10) |
            (1u << 15) | (1u << 16) | (static_cast<uint32_t>(c.NT >> 3) << 17) |
            (static_cast<uint32_t>(128 >> 4) << 24);
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(wgrad_kernel)");
    attr_set = true;
  }
  dim3 grid(c.nnt * c.mblks, c.ksplit);
  wgrad_kernel<<<grid, kWgThreads, c.smem_bytes, st>>>(p);
  ms_status s = after_launch("wgrad_kernel");
  if (s != MS_OK) return s;
  int sl = 1;
  while (sl < kRedLanes && sl * 2 <= c.ksplit) sl *= 2;
  dim3 rgrid(ceil_div(cn / fold, 32), ceil_div(cm, kRedLanes / sl));
  wgrad_reduce_kernel<<<rgrid, 256, 0, st>>>(p.part, dw, c.ksplit, taps, cm, cn, fold, mode, stride,
                                            pad, cout, ksize, shifts[0], beta, alpha, sl);
  return after_launch("wgrad_reduce_kernel");
}

}  // extern "C"
