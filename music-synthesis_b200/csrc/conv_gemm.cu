// Dense 1-D (dilated / transposed) convolution as a tcgen05 implicit GEMM.
//
//   D[m, n] = sum_t sum_ci  X[m + off_t, ci] * W_t[n, ci]        (per clip)
//
// GEMM rows m are time steps (UMMA M = 128, TMEM lanes), columns n are output
// channels (MS_CONV) or (phase, channel) pairs (MS_CONVT in polyphase form), K runs
// over input channels, one pass per tap.  The A operand of tap t is the SAME staged
// activation tile viewed with its start address advanced by (off_t - min_off) rows:
// activations are kept channel-blocked (B, C/8, L, 8) so that a row is one 16-byte
// vector, core matrices are 8 consecutive rows (128 contiguous bytes, SWIZZLE_NONE)
// and any row shift is a +16-byte descriptor offset -- no im2col, no re-load per tap.
//
// Persistent CTAs (one per SM), warp-specialised:
//   warp 0      producer: bulk async copies (TMA unit) global -> smem ring, zero-fills
//               out-of-range rows (= the conv's zero padding) at clip edges
//   warp 1      MMA issuer (one thread), owns the TMEM allocation
//   warps 2-5   epilogue: tcgen05.ld -> alpha/bias/LeakyReLU/residual -> global
// Accumulators are double-buffered in tensor memory so the epilogue of tile i
// overlaps the MMAs of tile i+1.
//
// Replaces F.conv1d / F.conv_transpose1d at featuresynth/generator/full.py:24-43 and
// featuresynth/util/modules.py:358-388 (see include/msb200.h).
#include "conv_gemm.cuh"

#include <cstdlib>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

// ---------------------------------------------------------------- configuration
// MSB_CONV_PAIR=0 disables the CTA-pair kernel (process-wide, read once)
static bool pair_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_CONV_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static int pick_nt(int ntot) {
  for (int nt = 256; nt >= 16; nt -= 16)
    if (ntot % nt == 0) return nt;
  return 0;
}

long long* g_conv_dbg = nullptr;   // debugging aid (MSB_CONV_ABLATE builds): see ms_debug_set_conv_trace

// MSB_CONV_WRES=1 enables the weight-resident mode of the pair kernel.  Off by default: measured
// on ConvTranspose 256 -> 128 (64 clips) 195.5 vs 199.8 us -- the layer is bound by its output
// stores, and walking the n-tiles outermost re-reads the input from HBM once per n-tile
// (262 vs 67 MB, ncu) because 600 MB of output stream through L2 in between.
static bool wres_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_CONV_WRES");
    v = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

bool make_conv_cfg(const ms_conv_desc& d, ConvCfg* c) {
  c->wres = 0;
  c->out_stage = 0;
  if (d.batch <= 0 || d.lin <= 0 || d.cin <= 0 || d.cout <= 0) return false;
  if (d.cin % 16 != 0 || d.cout % 8 != 0) return false;
  if (d.operand != MS_F16 && d.operand != MS_BF16) return false;
  const int xrep = d.x_repeat > 1 ? d.x_repeat : 1;
  if (d.cin % xrep != 0 || (d.cin / xrep) % 16 != 0) return false;
  const int xcin = d.cin / xrep;     // k-blocks must not straddle two copies of x
  if (d.kind == MS_CONV) {
    if (d.stride != 1 && d.stride != 0) return false;
    if (d.ksize < 1 || d.ksize > kMaxTaps || d.dilation < 1 || d.pad < 0) return false;
    if (d.cout % 16 != 0) return false;
    c->taps = d.ksize;
    for (int t = 0; t < d.ksize; ++t) c->off[t] = t * d.dilation - d.pad;
    c->Ntot = d.cout;
    if (d.crop < 0) return false;
    c->Lout = d.lin + 2 * d.pad - d.dilation * (d.ksize - 1) - d.crop;
    c->Lm = c->Lout;
  } else if (d.kind == MS_CONVT) {
    // polyphase form, J = ksize / stride taps at the input rate:
    //   out[s*q + r - pad] = sum_{j<J} x[q - j] * W[:, :, r + s*j]
    // GEMM tap t reads x[q + off[t]], off[t] = t - (J - 1), with weight tap r + s*(J - 1 - t)
    // (k = 2s: the upsamplers of generator/full.py; k = 4s: LearnedUpSample(.., 8, 2) of
    // generator/filterbank.py:108-114)
    if (d.stride < 1 || d.ksize % d.stride != 0) return false;
    const int J = d.ksize / d.stride;
    if (J < 2 || J > kMaxTaps) return false;
    if (d.pad < 0 || d.pad > (J - 1) * d.stride) return false;
    c->taps = J;
    for (int t = 0; t < J; ++t) c->off[t] = t - (J - 1);
    c->Ntot = d.stride * d.cout;
    if (c->Ntot % 16 != 0) return false;
    c->Lout = (d.lin - 1) * d.stride - 2 * d.pad + d.ksize;
    // GEMM rows q = 0 .. lin-1; the J - 1 extra rows q = lin .. lin+J-2 (they only meet the
    // last input rows; the last output rows of each clip) are computed by convt_tail_kernel so
    // that lin = 256 is ONE tile
    c->Lm = d.lin;
  } else {
    return false;
  }
  if (c->Lout <= 0) return false;
  int mn = c->off[0], mx = c->off[0];
  for (int t = 1; t < c->taps; ++t) {
    mn = c->off[t] < mn ? c->off[t] : mn;
    mx = c->off[t] > mx ? c->off[t] : mx;
  }
  c->min_off = mn;
  c->fold_slots = 0;
  c->fold_stride = 0;
  // short sequences (the discriminators' top layers run on 3..64 time steps): several clips
  // share one 128-row tile, each in a slot of Lm + span rows so that no tap crosses clips
  const bool fold = d.kind == MS_CONV && d.batch > 1 && 2 * c->Lm + (mx - mn) <= 128;
  c->NT = pick_nt(c->Ntot);
  if (c->NT == 0) return false;
  // geometry of a folded launch; NT / KB / nnt (the packed-weight layout) are already final
  auto apply_fold = [&]() {
    c->nkb = d.cin / c->KB;
    c->MBLK = 1;
    c->RA = 128 + mx - mn;
    c->fold_stride = c->Lm + (mx - mn);
    c->fold_slots = (128 - c->Lm) / c->fold_stride + 1;
    if (c->fold_slots > d.batch) c->fold_slots = d.batch;
    c->mtiles = 1;
    c->a_stage_bytes = (c->KB / 8) * c->RA * 16;
    c->w_stage_bytes = c->taps * (c->KB / 8) * c->NT * 16;
    c->stage_bytes = (c->a_stage_bytes + c->w_stage_bytes + 127) / 128 * 128;
    int fs = (kSmemBudget - kSmemHeader) / c->stage_bytes;
    if (fs > kMaxStages) fs = kMaxStages;
    if (fs < 2) return false;
    c->stages = fs;
    c->acc_stages = 2;
    int fcols = 32;
    while (fcols < 2 * c->NT) fcols *= 2;
    c->tmem_cols = fcols;
    size_t fsmem = kSmemHeader + static_cast<size_t>(c->stages) * c->stage_bytes;
    if (fsmem < 120 * 1024) fsmem = 120 * 1024;
    c->smem_bytes = fsmem;
    c->packed_weight_bytes = static_cast<size_t>(c->nnt) * c->nkb * c->w_stage_bytes;
    return true;
  };
  // the pair kernel pays off when a tile carries enough K work to amortise the cross-CTA
  // hand-offs (measured: wins for K*taps >= 512, loses for the small s=2 upsamplers)
  c->pair = (pair_enabled() && c->NT % 32 == 0 && d.cin * c->taps >= 512) ? 1 : 0;
  if (c->pair) {
    // CTA pair: 256-row cluster tile, one M-block + half of every weight stage per CTA
    c->MBLK = 1;
    c->RA = 128 + mx - mn;
    const int nth = c->NT / 2;
    c->KB = xcin % 64 == 0 ? 64 : (xcin % 32 == 0 ? 32 : 16);
    auto pstage = [&](int kb) { return (kb / 8) * c->RA * 16 + c->taps * (kb / 8) * nth * 16; };
    // ConvTranspose with >= 4 phases: a thread's four phases of a channel block are 128
    // contiguous output bytes, but neighbouring threads are stride*32 bytes apart -- direct
    // stores touch 32 lines per instruction (measured: the k16 s8 upsamplers drained their fp32
    // output at ~12 B/clk per SM).  Their epilogue goes through a per-warp staging tile.
    c->out_stage = (d.kind == MS_CONVT && d.stride % 4 == 0 && c->NT % 32 == 0 && !fold)
                       ? kPairEpiWarps * kPairStageWarp : 0;
    const int pbudget = kSmemBudget - kSmemHeader - c->out_stage;
    while (pstage(c->KB) * 3 > pbudget && c->KB > 16) c->KB /= 2;
    if (pstage(c->KB) * 2 > pbudget) return false;
    c->nnt = c->Ntot / c->NT;
    c->nkb = d.cin / c->KB;
    c->mtiles = (c->Lm + 255) / 256;
    c->a_stage_bytes = (c->KB / 8) * c->RA * 16;
    c->w_stage_bytes = c->taps * (c->KB / 8) * nth * 16;   // per CTA (half of the n-tile)
    c->stage_bytes = (c->a_stage_bytes + c->w_stage_bytes + 127) / 128 * 128;
    int ps = pbudget / c->stage_bytes;
    if (ps > kMaxStages) ps = kMaxStages;
    c->stages = ps;
    c->acc_stages = 2;
    int pcols = 32;
    while (pcols < 2 * c->NT) pcols *= 2;
    c->tmem_cols = pcols;
    size_t psmem = kSmemHeader + c->out_stage + static_cast<size_t>(c->stages) * c->stage_bytes;
    if (psmem < 120 * 1024) psmem = 120 * 1024;
    c->smem_bytes = psmem;
    c->packed_weight_bytes = static_cast<size_t>(c->nnt) * 2 * c->nkb * c->w_stage_bytes;
    // Weight-resident mode: when this CTA's half of an n-tile's weights (all k-blocks, all taps)
    // fits beside three activation stages, it is loaded once per n-tile and the cluster walks
    // the m-tiles of that n-tile (ConvTranspose 256 -> 128, k16 s8: 128 KB per CTA; the layer
    // was bound by re-streaming 256 KB of weights per 256 x 256 tile through the SM<->L2 fabric)
    c->wres = 0;
    if (!fold && wres_enabled()) {
      const int slice = c->nkb * c->w_stage_bytes;
      const int a_stage = (c->a_stage_bytes + 127) / 128 * 128;
      const int as = slice <= 128 * 1024 ? (pbudget - slice) / a_stage : 0;
      if (as >= 3) {
        c->wres = 1;
        c->stage_bytes = a_stage;
        c->stages = as > kMaxStages ? kMaxStages : as;
        c->smem_bytes = kSmemHeader + c->out_stage + static_cast<size_t>(slice) +
                        static_cast<size_t>(c->stages) * c->stage_bytes;
      }
    }
    if (!fold) return true;
    // folded tiles run on the single-CTA kernel: the pair layout IS a single-CTA layout with
    // n-tiles of half the width, so the packed image stays the same for every length
    c->pair = 0;
    c->out_stage = 0;
    c->NT = nth;
    c->nnt = 2 * c->nnt;
    return apply_fold();
  }
  c->MBLK = c->Lm > 128 ? 2 : 1;
  c->RA = 128 * c->MBLK + mx - mn;
  c->KB = xcin % 64 == 0 ? 64 : (xcin % 32 == 0 ? 32 : 16);
  // NT / KB define the packed weight layout, so they must not depend on the input
  // length: size the stages for the largest tile (two M-blocks) regardless of MBLK
  const int ra_max = 256 + mx - mn;
  auto stage = [&](int kb, int nt) {
    return (kb / 8) * ra_max * 16 + c->taps * (kb / 8) * nt * 16;
  };
  const int budget = kSmemBudget - kSmemHeader;
  // want >= 3 stages when possible
  while (stage(c->KB, c->NT) * 3 > budget && c->KB > 16) c->KB /= 2;
  while (stage(c->KB, c->NT) * 2 > budget && c->NT > 16) {
    int nt = c->NT - 16;
    while (nt >= 16 && c->Ntot % nt != 0) nt -= 16;
    if (nt < 16) return false;
    c->NT = nt;
  }
  if (stage(c->KB, c->NT) * 2 > budget) return false;
  c->nnt = c->Ntot / c->NT;
  if (fold) return apply_fold();
  c->nkb = d.cin / c->KB;
  c->mtiles = (c->Lm + 128 * c->MBLK - 1) / (128 * c->MBLK);
  c->a_stage_bytes = (c->KB / 8) * c->RA * 16;
  c->w_stage_bytes = c->taps * (c->KB / 8) * c->NT * 16;
  c->stage_bytes = (c->a_stage_bytes + c->w_stage_bytes + 127) / 128 * 128;
  int s = budget / c->stage_bytes;
  if (s > kMaxStages) s = kMaxStages;
  int need = c->nkb * 2 > 2 ? c->nkb * 2 : 2;  // no point in more than 2 tiles in flight
  if (s > need) s = need;
  if (s < 2) return false;
  c->stages = s;
  c->acc_stages = (2 * c->MBLK * c->NT <= 512) ? 2 : 1;
  int cols = 32;
  while (cols < c->acc_stages * c->MBLK * c->NT) cols *= 2;
  c->tmem_cols = cols;
  size_t smem = kSmemHeader + static_cast<size_t>(c->stages) * c->stage_bytes;
  // always ask for more than half an SM's shared memory: one CTA per SM, so the
  // 2*NT-column TMEM allocation can never contend with a co-resident CTA
  if (smem < 120 * 1024) smem = 120 * 1024;
  c->smem_bytes = smem;
  c->packed_weight_bytes = static_cast<size_t>(c->nnt) * c->nkb * c->w_stage_bytes;
  return true;
}

// ------------------------------------------------------------------ the kernel
__device__ __forceinline__ uint32_t pack2(float a, float b, int operand) {
  if (operand == MS_BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

// 16 epilogue warps = four column slices per TMEM lane quarter.  With 8 (two per scheduler) the
// drain of a 128 x 256 accumulator took ~7.6 k cycles against 4.6 k cycles of MMAs per tile: the
// epilogue is a dependent FFMA/FMNMX chain per element, and two warps cannot fill a scheduler
// (tools/pair_trace.py; the stack kernels run 16 as well)
constexpr int kConvEpiWarps = 8;
constexpr int kConvEpiSlices = kConvEpiWarps / 4;
constexpr int kConvThreads = 64 + 32 * kConvEpiWarps;

__global__ void __launch_bounds__(kConvThreads, 1)
conv_gemm_kernel(const __grid_constant__ ConvGemmParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // [0..7] full, [8..15] empty, [16..17] tmem_full, [18..19] tmem_empty
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 256);
  // per-tile epilogue tables, double-buffered by accumulator stage:
  //   s_bias[2][256] fp32 (bias of each tile column), s_tab[2][32] (phase r, channel of each
  //   8-column chunk) -- so the epilogue has no dependent global loads / integer divisions
  float* s_bias = reinterpret_cast<float*>(smem + 512);
  int2* s_tab = reinterpret_cast<int2*>(smem + 512 + 2048);
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t data_base = smem_u32(smem + kSmemHeader);
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (8 + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (16 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (18 + a); };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kConvEpiWarps);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), p.tmem_cols);
    tmem_relinquish();
  }
  if (p.fold_slots > 0) {
    // folded tiles: the rows between clip slots are never written by the copies -- they are the
    // convolution's zero padding, cleared once for every stage of the ring
    for (int st = 0; st < p.stages; ++st) {
      const uint32_t sA = data_base + static_cast<uint32_t>(st) * p.stage_bytes;
      for (int i = threadIdx.x; i < p.a_stage_bytes / 16; i += kConvThreads)
        st_shared_v4(sA + static_cast<uint32_t>(i) * 16u, 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int chunks = p.KB >> 3;  // 16-byte channel chunks per k-block
  const int rows_per_tile = 128 * p.MBLK;
  const int acc_cols = p.MBLK * p.NT;

  if (warp == 0) {
    // =============================== producer ===============================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt_idx = tile % p.nnt;
      const int rest = tile / p.nnt;
      const int mt = rest % p.mtiles;
      const int b = rest / p.mtiles;
      const int r0 = mt * rows_per_tile + p.min_off;
      const int lo = r0 < 0 ? 0 : r0;
      const int hi = (r0 + p.RA) > p.lin ? p.lin : (r0 + p.RA);
      const int nrows = hi > lo ? hi - lo : 0;
      const bool ragged = (nrows != p.RA) && p.fold_slots == 0;
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(empty_bar(stage), phase ^ 1u);
        const uint32_t sA = data_base + static_cast<uint32_t>(stage) * p.stage_bytes;
        const uint32_t sW = sA + p.a_stage_bytes;
        if (p.fold_slots > 0) {
          // slot j of the tile holds clip b*slots + j: A row j*stride + a <-> input row min_off + a
          const int b0 = b * p.fold_slots;
          const int nclips = min(p.fold_slots, p.B - b0);
          const int a_lo = p.min_off < 0 ? -p.min_off : 0;                 // first valid A row
          int a_hi = p.lin - p.min_off;                                    // one past the last
          if (a_hi > p.fold_stride) a_hi = p.fold_stride;
          const int frows = a_hi > a_lo ? a_hi - a_lo : 0;
          if (lane == 0) {
            mbar_arrive_expect_tx(full_bar(stage),
                                  static_cast<uint32_t>(frows) * 16u * chunks * nclips +
                                      p.w_stage_bytes);
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) +
                                  (static_cast<size_t>(nt_idx) * p.nkb + kb) * p.w_stage_bytes;
            bulk_g2s(sW, wsrc, p.w_stage_bytes, full_bar(stage));
          }
          __syncwarp();
          if (frows > 0) {
            // one small copy per (clip, channel chunk): all 32 lanes issue them
            for (int i = lane; i < nclips * chunks; i += 32) {
              const int j = i / chunks, c = i - j * chunks;
              const size_t cbase = static_cast<size_t>(b0 + j) * (p.xcin >> 3) + (kb % p.xnkb) * chunks;
              bulk_g2s(sA + static_cast<uint32_t>(c * p.RA + j * p.fold_stride + a_lo) * 16u,
                       p.x + ((cbase + c) * p.lin + (p.min_off + a_lo)) * 8,
                       static_cast<uint32_t>(frows) * 16u, full_bar(stage));
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1u;
          }
          continue;
        }
        if (ragged) {
          // rows outside [0, lin) are the convolution's zero padding
          const int head = lo - r0;                    // rows [0, head)
          const int tail0 = head + nrows;              // rows [tail0, RA)
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
        // lane 0 announces the bytes and streams the weight stage; the activation copies (one
        // per 16-byte channel chunk) are issued by as many lanes in parallel -- a single lane
        // needed ~1000 cycles per stage for eight of them (tools/pair_trace.py)
        if (lane == 0) {
          const uint32_t bytes_a = static_cast<uint32_t>(nrows) * 16u * chunks;
          mbar_arrive_expect_tx(full_bar(stage), bytes_a + p.w_stage_bytes);
          const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w) +
                                (static_cast<size_t>(nt_idx) * p.nkb + kb) * p.w_stage_bytes;
          bulk_g2s(sW, wsrc, p.w_stage_bytes, full_bar(stage));
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
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    // Warp-uniform loop; only the tcgen05 instructions are predicated on one elected
    // lane so the descriptors live in uniform registers.
    const uint32_t idesc = umma_idesc_f16(p.NT, p.operand);
    const uint64_t adesc0 = umma_desc_base_nosw(static_cast<uint32_t>(p.RA) * 16u, 128);
    const uint64_t bdesc0 = umma_desc_base_nosw(static_cast<uint32_t>(p.NT) * 16u, 128);
    const uint32_t a_kstep = static_cast<uint32_t>(2 * p.RA);   // 16-byte units per k16
    const uint32_t b_kstep = static_cast<uint32_t>(2 * p.NT);
    const int nk16 = p.KB >> 4;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_cols);
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(full_bar(stage), phase);
        tc_fence_after();
        const uint32_t sA = data_base + static_cast<uint32_t>(stage) * p.stage_bytes;
        const uint32_t sW = sA + p.a_stage_bytes;
        if (elect_one()) {
          for (int t = 0; t < p.taps; ++t) {
            const uint64_t bd0 = bdesc0 + ((sW >> 4) + static_cast<uint32_t>(t * chunks * p.NT));
            for (int mb = 0; mb < p.MBLK; ++mb) {
              uint64_t ad = adesc0 + ((sA >> 4) + static_cast<uint32_t>(p.off[t] - p.min_off + mb * 128));
              uint64_t bd = bd0;
              const uint32_t dst = d_tmem + static_cast<uint32_t>(mb * p.NT);
              for (int k16 = 0; k16 < nk16; ++k16) {
                umma_f16_ss(dst, ad, bd, idesc, (kb | t | k16) != 0 ? 1u : 0u);
                ad += a_kstep;
                bd += b_kstep;
              }
            }
          }
          umma_commit(empty_bar(stage));
          if (kb == p.nkb - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1u;
      }
    }
  } else {
    // =============================== epilogue ===============================
    // 8 warps: lane quarter q = warp % 4 (hardware rule), two warps per quarter take
    // alternate 32-column groups.  TMEM loads are batched (32 columns per wait).
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;   // column slice of this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    const int cout8 = p.cout >> 3;
    const int ngroups = p.NT >> 4;            // 16-column groups per M-block
    int tab_nt0 = -1, tab_nt1 = -1;     // n-tile whose tables sit in buffer 0 / 1
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt_idx = tile % p.nnt;
      const int rest = tile / p.nnt;
      const int mt = rest % p.mtiles;
      const int b = rest / p.mtiles;
      const int n0 = nt_idx * p.NT;
      // bias / column tables of this tile's n-tile: rebuilt only when the n-tile of this
      // accumulator stage changes (for nnt = 1 layers once per CTA; the fill -- a global load
      // and a barrier of all epilogue warps -- cost 1.1-2.3 k cycles per tile on the epilogue's
      // critical path, tools/pair_trace.py)
      if (((acc & 1) ? tab_nt1 : tab_nt0) != nt_idx || (p.debug & 256) != 0) {
        // every epilogue warp is done with the tiles that read this buffer
        named_bar_sync(1, 32 * kConvEpiWarps);
        const int et = threadIdx.x - 64;              // 0 .. 255
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
        named_bar_sync(1, 32 * kConvEpiWarps);
        if (acc & 1) tab_nt1 = nt_idx; else tab_nt0 = nt_idx;
      }
      const float* tbias = s_bias + (acc & 1) * 256;
      const int2* ttab = s_tab + (acc & 1) * 32;
      if (p.res32 != nullptr && p.kind == MS_CONV && p.fold_slots == 0) {
        // residual rows of this tile into L2 while the MMAs still run (the epilogue is
        // latency-bound on these loads)
        for (int mb = 0; mb < p.MBLK; ++mb) {
          const int pm = mt * rows_per_tile + mb * 128 + q * 32 + lane;
          if (pm >= p.Lm) continue;
          for (int g = 2 * half; g < ngroups; g += 2 * kConvEpiSlices) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const int cidx = g * 2 + h;
              if (cidx * 8 >= p.NT) break;
              const int ch = n0 + cidx * 8;
              prefetch_l2(p.res32 + ((static_cast<size_t>(b) * cout8 + (ch >> 3)) * p.Lout + pm) * 8);
            }
          }
        }
      }
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      for (int mb = 0; mb < p.MBLK; ++mb) {
        int m = mt * rows_per_tile + mb * 128 + q * 32 + lane;
        int bb = b;
        bool slot_ok = true;
        if (p.fold_slots > 0) {
          const int j = m / p.fold_stride;
          m -= j * p.fold_stride;                       // row inside the clip
          bb = b * p.fold_slots + j;
          slot_ok = j < p.fold_slots && bb < p.B;
        }
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_cols + mb * p.NT);
        // this warp handles group pairs (2 x 16 columns) g = 2*half, 2*half+4, ...
        for (int g = 2 * half; g < ngroups; g += 2 * kConvEpiSlices) {
          const bool two = (g + 1) < ngroups;
          // the residual vectors of this group's (up to four) chunks are requested BEFORE the
          // accumulator is read: four independent loads in flight per thread instead of one L2
          // round trip per chunk (the loads used to sit between the stores of consecutive chunks)
          float r8[4][8];
          if (p.res32 != nullptr) {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              if (h >= 2 && !two) break;
              const int2 rc = ttab[g * 2 + h];
              const int orow = (p.kind == MS_CONVT) ? p.stride * m + rc.x - p.pad : m;
              if (slot_ok && (m < p.Lm) && orow >= 0 && orow < p.Lout)
                ld_global_nc_v8(p.res32 + ((static_cast<size_t>(bb) * cout8 + (rc.y >> 3)) * p.Lout + orow) * 8,
                                r8[h]);
            }
          }
          uint32_t v[32];
          tmem_ld16p(taddr + g * 16, &v[0]);
          if (two) tmem_ld16p(taddr + (g + 1) * 16, &v[16]);
          tmem_ld_wait();
          if (mb == p.MBLK - 1 && g + 2 * kConvEpiSlices >= ngroups) {
            // last TMEM read of this tile by this thread: hand the accumulator back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            if (h >= 2 && !two) break;
            const int cidx = g * 2 + h;               // 8-column chunk within the tile
            const int2 rc = ttab[cidx];
            const int ch = rc.y;
            const int orow = (p.kind == MS_CONVT) ? p.stride * m + rc.x - p.pad : m;
            const bool valid = slot_ok && (m < p.Lm) && orow >= 0 && orow < p.Lout;
            if (!valid) continue;
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
            const size_t idx = (static_cast<size_t>(bb) * cout8 + (ch >> 3)) * p.Lout + orow;
            if (p.res32 != nullptr) {
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] += r8[h][j];
            }
            if (p.leaky == 2) {                       // activation AFTER the residual add
#pragma unroll
              for (int j = 0; j < 8; ++j) f[j] = leaky02(f[j]);
            }
            if (p.y32 != nullptr) st_global_v8(p.y32 + idx * 8, f);
            if (p.y16 != nullptr) {
              uint4 o;
              o.x = pack2(f[0], f[1], p.operand);
              o.y = pack2(f[2], f[3], p.operand);
              o.z = pack2(f[4], f[5], p.operand);
              o.w = pack2(f[6], f[7], p.operand);
              *reinterpret_cast<uint4*>(p.y16 + idx * 8) = o;
            }
          }
        }
      }
      // warps with no column group in this tile (tiny NT) still have to arrive
      if (2 * half >= ngroups) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      if (++acc == p.acc_stages) {
        acc = 0;
        acc_phase ^= 1u;
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

// ------------------------------------- ConvTranspose tail (GEMM rows q = lin .. lin+J-2)
// out[b, co, s*(lin+e) + r - pad] = act(bias[co] + sum_{t <= J-2-e} sum_ci
//                                       x[b, ci, lin + e + t - (J-1)] * W[ci, co, r + s*(J-1-t)])
// Reads the SAME packed 16-bit weights / 16-bit activations as the main GEMM, accumulates in
// fp32.  Tiny: at most (J-1)*stride*cout outputs per clip.  blockIdx.y = e * stride + r.
constexpr int kTailClips = 8;     // clips per block: a weight vector is loaded once for all of them
constexpr int kTailMaxSmem = 40 * 1024;
// 16 GEMM columns x 8 input-channel slices per block; the block's clips' input rows sit in shared
// memory and every weight vector (packed[nt][kb][tap][c][nn][0..7]) is used for all of them; the
// eight slices of a column are combined in a fixed shuffle order.  (Earlier versions streamed each
// output's weights once per clip through L2: 25-37 us for the 134 MFLOP of the 512 -> 256
// upsampler's tail row; one column per thread without slices was latency-bound at small batches.)
// grid: (ceil(Ntot / 16), ceil(B / clips), taps - 1); dynamic smem = clips * rows * cin * 2 bytes.
constexpr int kTailSlices = 8;
__global__ void __launch_bounds__(128)
convt_tail_kernel(const ConvGemmParams p, int ntp /* packed n-tile */, int clips) {
  extern __shared__ __align__(16) uint8_t tail_smem[];
  uint4* sx = reinterpret_cast<uint4*>(tail_smem);      // [clip][tap][chunk]
  const int e = blockIdx.z;                              // tail row
  const int b0 = blockIdx.y * clips;
  const int slice = threadIdx.x & (kTailSlices - 1);
  const int n = blockIdx.x * (128 / kTailSlices) + (threadIdx.x >> 3);
  const int ntaps = p.taps - 1 - e;                      // taps that meet real input rows
  const int nch = p.cin >> 3;
  const int xc8 = p.xcin >> 3;
  // stage x[b, :, lin + e + t - (taps-1)] for t < ntaps (zero rows before the clip start)
  for (int i = threadIdx.x; i < clips * ntaps * nch; i += blockDim.x) {
    const int c8 = i % nch;
    const int t = (i / nch) % ntaps;
    const int cl = i / (nch * ntaps);
    const int xrow = p.lin + e + t - (p.taps - 1);
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (b0 + cl < p.B && xrow >= 0)
      v = __ldg(reinterpret_cast<const uint4*>(
          p.x + ((static_cast<size_t>(b0 + cl) * xc8 + c8 % xc8) * p.lin + xrow) * 8));
    sx[i] = v;
  }
  __syncthreads();
  const int nq = n < p.Ntot ? n : 0;
  const int r = convt_col_phase(nq, p.stride), co = convt_col_channel(nq, p.stride);
  const int orow = p.stride * (p.lin + e) + r - p.pad;
  const bool live = n < p.Ntot && orow >= 0 && orow < p.Lout;
  const int nt = nq / ntp, nn = nq - nt * ntp;
  const int chunks = p.KB >> 3;
  float acc[kTailClips];
#pragma unroll
  for (int cl = 0; cl < kTailClips; ++cl) acc[cl] = 0.f;
  for (int t = 0; live && t < ntaps; ++t) {
#pragma unroll 4
    for (int c8 = slice; c8 < nch; c8 += kTailSlices) {
      const int ci = c8 * 8;
      const int kb = ci / p.KB, c = (ci % p.KB) >> 3;
      const size_t wi = ((((static_cast<size_t>(nt) * p.nkb + kb) * p.taps + t) * chunks + c) * ntp + nn) * 8;
      const uint4 wq = __ldg(reinterpret_cast<const uint4*>(p.w + wi));
      const uint32_t ww[4] = {wq.x, wq.y, wq.z, wq.w};
      float wf[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float2 f2;
        if (p.operand == MS_BF16) f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&ww[j]));
        else f2 = __half22float2(*reinterpret_cast<const __half2*>(&ww[j]));
        wf[2 * j] = f2.x; wf[2 * j + 1] = f2.y;
      }
#pragma unroll
      for (int cl = 0; cl < kTailClips; ++cl) {
        if (cl >= clips) break;
        const uint4 xq = sx[(cl * ntaps + t) * nch + c8];
        const uint32_t xx[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 xf;
          if (p.operand == MS_BF16) xf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&xx[j]));
          else xf = __half22float2(*reinterpret_cast<const __half2*>(&xx[j]));
          acc[cl] = fmaf(xf.x, wf[2 * j], acc[cl]);
          acc[cl] = fmaf(xf.y, wf[2 * j + 1], acc[cl]);
        }
      }
    }
  }
#pragma unroll
  for (int cl = 0; cl < kTailClips; ++cl) {
    const int b = b0 + cl;
    // fixed-order combination of the 8 slices (adjacent lanes)
    float a = acc[cl];
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    a += __shfl_xor_sync(0xffffffffu, a, 4);
    if (!live || slice != 0 || cl >= clips || b >= p.B) continue;
    float v = a * p.alpha + (p.bias != nullptr ? p.bias[co] : 0.f);
    if (p.leaky == 1) v = leaky02(v);
    const size_t idx = ((static_cast<size_t>(b) * (p.cout >> 3) + (co >> 3)) * p.Lout + orow) * 8 + (co & 7);
    if (p.res32 != nullptr) v += p.res32[idx];
    if (p.leaky == 2) v = leaky02(v);
    if (p.y32 != nullptr) p.y32[idx] = v;
    if (p.y16 != nullptr) {
      if (p.operand == MS_BF16) {
        __nv_bfloat16 h = __float2bfloat16_rn(v);
        p.y16[idx] = *reinterpret_cast<uint16_t*>(&h);
      } else {
        __half h = __float2half_rn(v);
        p.y16[idx] = *reinterpret_cast<uint16_t*>(&h);
      }
    }
  }
}

// ------------------------------------------------------------ weight packing
// packed[nt][kb][tap][chunk][nn][e]  <-  reference-layout fp32 weights
__global__ void pack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out,
                                   int kind, int cin, int cout, int ksize, int stride,
                                   int taps, int NT, int KB, int nnt, int nkb, int operand,
                                   size_t total) {
  size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  size_t r = i;
  const int e = r % 8; r /= 8;
  const int nn = r % NT; r /= NT;
  const int c = r % (KB / 8); r /= (KB / 8);
  const int t = r % taps; r /= taps;
  const int kb = r % nkb; r /= nkb;
  const int nt = static_cast<int>(r);
  const int n = nt * NT + nn;
  const int ci = kb * KB + c * 8 + e;
  float v;
  if (kind == MS_CONV) {
    v = w[(static_cast<size_t>(n) * cin + ci) * ksize + t];
  } else {
    const int ph = convt_col_phase(n, stride);
    const int co = convt_col_channel(n, stride);
    const int kk = ph + stride * (taps - 1 - t);
    v = w[(static_cast<size_t>(ci) * cout + co) * ksize + kk];
  }
  if (operand == MS_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  } else {
    __half h = __float2half_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

ms_status launch_conv_pair(const ConvGemmParams& p, size_t smem_bytes, cudaStream_t stream);

ms_status launch_conv(const ms_conv_desc& d, const ConvCfg& c, const void* x16,
                      const void* w_packed, const float* bias, const float* res32, void* y16,
                      float* y32, cudaStream_t stream) {
  ConvGemmParams p;
  p.x = static_cast<const uint16_t*>(x16);
  p.w = static_cast<const uint16_t*>(w_packed);
  p.bias = bias;
  p.res32 = res32;
  p.y16 = static_cast<uint16_t*>(y16);
  p.y32 = y32;
  p.B = d.batch; p.cin = d.cin; p.lin = d.lin; p.cout = d.cout;
  p.xcin = d.cin / (d.x_repeat > 1 ? d.x_repeat : 1);
  p.xnkb = p.xcin / c.KB;
  p.Lout = c.Lout; p.Lm = c.Lm;
  p.taps = c.taps;
  for (int t = 0; t < kMaxTaps; ++t) p.off[t] = t < c.taps ? c.off[t] : 0;
  p.min_off = c.min_off; p.RA = c.RA;
  p.Ntot = c.Ntot; p.NT = c.NT; p.KB = c.KB; p.nnt = c.nnt; p.nkb = c.nkb;
  p.mtiles = c.mtiles; p.MBLK = c.MBLK; p.acc_stages = c.acc_stages; p.pair = c.pair; p.wres = c.wres; p.out_stage = c.out_stage;
  p.debug = 0;
  p.dbg = nullptr;
  {
    // MSB_CONV_TABCACHE=0: rebuild the epilogue tables for every tile (A/B timing)
    static int tabcache = -1;
    if (tabcache < 0) {
      const char* e = getenv("MSB_CONV_TABCACHE");
      tabcache = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (tabcache == 0) p.debug |= 256;
    // MSB_CONV_FASTEPI=0: plain convs take the generic epilogue (A/B timing)
    static int fastepi = -1;
    if (fastepi < 0) {
      const char* e = getenv("MSB_CONV_FASTEPI");
      fastepi = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    if (fastepi == 0) p.debug |= 512;
  }
#ifdef MSB_CONV_ABLATE
  if (const char* e = getenv("MSB_CONV_ABLATE")) p.debug |= atoi(e);
  p.dbg = g_conv_dbg;
#endif
  p.stages = c.stages; p.a_stage_bytes = c.a_stage_bytes; p.w_stage_bytes = c.w_stage_bytes;
  p.stage_bytes = c.stage_bytes; p.tmem_cols = c.tmem_cols;
  p.kind = d.kind; p.stride = d.stride; p.pad = d.pad; p.leaky = d.leaky;
  p.operand = d.operand; p.alpha = d.alpha;
  p.fold_slots = c.fold_slots; p.fold_stride = c.fold_stride;
  p.btiles = c.fold_slots > 0 ? (d.batch + c.fold_slots - 1) / c.fold_slots : d.batch;
  const long long tiles = static_cast<long long>(p.btiles) * c.mtiles * c.nnt;
  if (tiles > 0x7fffffffLL) return MS_ERR_INVALID;
  p.total_tiles = static_cast<int>(tiles);

  if (d.kind == MS_CONVT) {
    // the rows q >= lin first (independent outputs: the last rows of each clip)
    int clips = kTailClips;
    const size_t per_clip = static_cast<size_t>(c.taps - 1) * d.cin * 2;
    while (clips > 1 && clips * per_clip > static_cast<size_t>(kTailMaxSmem)) clips /= 2;
    if (clips * per_clip > static_cast<size_t>(kTailMaxSmem)) return MS_ERR_INVALID;
    dim3 tgrid((c.Ntot + 15) / 16, (d.batch + clips - 1) / clips, c.taps - 1);
    convt_tail_kernel<<<tgrid, 128, clips * per_clip, stream>>>(p, c.pair ? c.NT / 2 : c.NT, clips);
    ms_status ts = after_launch("convt_tail_kernel");
    if (ts != MS_OK) return ts;
  }
  if (c.pair) return launch_conv_pair(p, c.smem_bytes, stream);
  static thread_local size_t attr_set = 0;
  if (c.smem_bytes > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBudget);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(conv_gemm_kernel)");
    attr_set = kSmemBudget;
  }
  int sms = sm_count();
  if (sms <= 0) return check_cuda(cudaGetLastError(), "sm_count");
  int grid = p.total_tiles < sms ? p.total_tiles : sms;
  conv_gemm_kernel<<<grid, kConvThreads, c.smem_bytes, stream>>>(p);
  return after_launch("conv_gemm_kernel");
}

}  // namespace msb

using namespace msb;

extern "C" {

int ms_conv_out_len(const ms_conv_desc* d) {
  ConvCfg c;
  if (d == nullptr || !make_conv_cfg(*d, &c)) return MS_ERR_INVALID;
  return c.Lout;
}

size_t ms_conv_packed_weight_bytes(const ms_conv_desc* d) {
  ConvCfg c;
  if (d == nullptr || !make_conv_cfg(*d, &c)) return 0;
  return c.packed_weight_bytes;
}

ms_status ms_conv_pack_weight(const ms_conv_desc* d, const float* w_f32, void* w_packed,
                              void* stream) {
  ConvCfg c;
  if (d == nullptr || w_f32 == nullptr || w_packed == nullptr || !make_conv_cfg(*d, &c))
    return MS_ERR_INVALID;
  const size_t total = c.packed_weight_bytes / 2;
  const int threads = 256;
  const unsigned blocks = static_cast<unsigned>((total + threads - 1) / threads);
  pack_weight_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(
      w_f32, static_cast<uint16_t*>(w_packed), d->kind, d->cin, d->cout, d->ksize, d->stride,
      c.taps, c.pair ? c.NT / 2 : c.NT, c.KB, c.pair ? 2 * c.nnt : c.nnt, c.nkb, d->operand,
      total);
  return after_launch("pack_weight_kernel");
}

ms_status ms_conv_fwd(const ms_conv_desc* d, const void* x16, const void* w_packed,
                      const float* bias, const float* res32, void* y16, float* y32,
                      void* stream) {
  ConvCfg c;
  if (d == nullptr || x16 == nullptr || w_packed == nullptr || !make_conv_cfg(*d, &c))
    return MS_ERR_INVALID;
  if (y16 == nullptr && y32 == nullptr) return MS_ERR_INVALID;
  auto misaligned = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) != 0; };
  if (misaligned(x16) || misaligned(w_packed) || misaligned(bias) || misaligned(res32) ||
      misaligned(y16) || misaligned(y32))
    return MS_ERR_INVALID;
  return launch_conv(*d, c, x16, w_packed, bias, res32, y16, y32,
                     static_cast<cudaStream_t>(stream));
}

/* debugging aid (not in the public header): clock64 trace buffer (512 x int64, device) filled
 * by cluster 0's leader in subsequent pair-kernel launches (MSB_CONV_ABLATE builds only) */
void ms_debug_set_conv_trace(long long* dev_buf) { msb::g_conv_dbg = dev_buf; }

}  // extern "C"
