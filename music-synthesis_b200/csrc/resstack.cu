// Fused ResidualStack: three ResidualAtoms (six k=3 convolutions, dilations d0,d1,d2 /
// 1,1,1) per time tile in ONE kernel, activations never leaving the SM.
//   replaces ResidualStack.forward / ResidualAtom.forward,
//   featuresynth/util/modules.py:350-405, as used at generator/full.py:29,33,37,41.
//
// Per CTA tile of R = MB*128 rows (MB = 256/C M-blocks; R*C = 32768 always):
//   SMEM  X16 [C/8][R][8]   16-bit operand copy of the residual stream        (64 KB)
//         Y16 [C/8][R][8]   16-bit inner activation of the current atom       (64 KB)
//         W ring            one k-tap weight image (C x C, 16-bit) per slot   (<= 96 KB)
//   TMEM  per M-block: C columns fp32 residual stream + C columns accumulator (512 cols)
// The tile covers rows [t0-16, t0+R-16): 16 halo rows each side are recomputed (the
// stack's receptive field is +-16), only the middle R-32 rows are stored.  Rows outside
// the clip are forced to zero after every layer -- that IS each conv's zero padding.
// A tap with dilation shift s is the same SMEM buffer addressed s rows further
// (descriptor start address + 16*s bytes); reads that run off a tile edge only ever
// contaminate halo rows.
//
// Warp roles: warp 0 streams weight taps (bulk async copies) through the ring, warp 1
// issues tcgen05.mma, warps 2-9 are the prologue/epilogue: TMEM -> bias, LeakyReLU,
// residual add (fp32, TMEM resident) -> 16-bit operand for the next conv in SMEM.
// MMA and epilogue overlap at M-block granularity: conv l+1 starts on M-block 0 while
// the epilogue of conv l is still draining M-block 1.
#include <cstdlib>
#include <type_traits>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

constexpr int kStackHalo = 16;
constexpr int kStackHeader = 1024;

struct StackParams {
  const float* x32;      // BLK f32 (B, C/8, L, 8): stack input (upsampler output)
  const uint16_t* x16in; // alternative input: BLK 16-bit (the stream then starts from the rounded
                         // operand: 2 instead of 4 bytes per element in and out of HBM)
  const uint16_t* w;     // [6 convs][3 taps][C/8][C][8] 16-bit
  const float* bias;     // [6][C]
  uint16_t* y16;         // BLK 16-bit output or null
  float* y32;            // BLK f32 output or null
  int B, L;
  int dil[3];
  int tiles_per_clip, total_tiles;
  int halo, V;           // rows recomputed per tile side / rows stored per tile (R - 2*halo)
  int operand;
  // optional fused tail (C == 32 only): y = tanh(conv_k7_pad3(x, mono_w) + mono_b),
  // generator/full.py:43-44.  When mono_out != null, y16 / y32 are not written.
  const float* mono_w;   // (1, 32, 7) fp32, reference layout
  const float* mono_b;   // (1)
  float* mono_out;       // (B, 1, L) fp32
  long long* dbg;        // optional clock trace of CTA 0 (null in production)
  int ablate;            // MSB_STACK_ABLATE builds only (tools/stack_bench.py): bit 0 no MMAs,
                         // 1 empty epilogue, 2 no operand stores, 3 no tensor-memory traffic in
                         // the epilogue, 4 no bias pre-load
};

#ifdef MSB_STACK_ABLATE
#define MSB_SABL(bit) ((p.ablate & (bit)) != 0)
#else
#define MSB_SABL(bit) false
#endif

#ifdef MSB_STACK_TRACE
#define MSB_TRACE(slot)                                                     \
  do {                                                                      \
    if (p.dbg != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 &&   \
        nconv < 12)                                                         \
      p.dbg[(slot)] = clock64();                                            \
  } while (0)
#else
#define MSB_TRACE(slot) do { } while (0)
#endif

template <int C>
struct StackGeom {
  static constexpr int MB = 256 / C;             // M-blocks per tile
  static constexpr int NP = 2;                         // pipeline parts per tile (4 measured slower)
  static constexpr int HB = MB / NP;                   // M-blocks per part
  static constexpr int MSPLIT = HB >= 2 ? 2 : 1;       // M-block split of the epilogue
  static constexpr int PARTS = 4 / MSPLIT;             // column split of the epilogue
  static constexpr int EW = 16;                        // epilogue warps = 4 * PARTS * MSPLIT
  static constexpr int R = MB * 128;             // rows per tile
  static constexpr int NCH = C / 8;
  static constexpr int ACT_BYTES = R * C * 2;    // 65536
  static constexpr int TAP_BYTES = C * C * 2;
  static constexpr int NSLOT = (96 * 1024 / TAP_BYTES) < 18 ? (96 * 1024 / TAP_BYTES) : 18;
  // CTA-pair mode: each CTA stages half of every tap (the other N-half lives in the peer)
  static constexpr int TAP_BYTES_P = TAP_BYTES / 2;
  static constexpr int NSLOT_P = (96 * 1024 / TAP_BYTES_P) < 18 ? (96 * 1024 / TAP_BYTES_P) : 18;
  static constexpr int MONO_BYTES = 1024;        // [7][32] fp32 tail-conv weights
  // fused-tail partial sums P[2 parts][7 taps][R rows] (C == 32 only): a region of their own,
  // so that the X buffer is free during the last conv for the NEXT tile's operand
  static constexpr int P_BYTES = (C == 32) ? 14 * R * 4 : 0;
  // fold the next tile's prologue into the last conv's epilogue (needs COLS + 2 x 16 live
  // registers per row: fits the 96-register budget of 18-warp CTAs only at 16 columns per thread)
  static constexpr bool MERGE_PROLOGUE = (C == 32);   // C = 64 / 128 (two 16-column groups per row): measured slower (spills + a second exposed load)
  static constexpr int SMEM = kStackHeader + 2 * ACT_BYTES + NSLOT * TAP_BYTES + MONO_BYTES + P_BYTES;
  // producer, MMA issuer A, epilogue warps, MMA issuer B (single-CTA kernels only)
  static constexpr int THREADS = 64 + 32 * EW + 32;
  static constexpr int THREADS_PAIR = THREADS;
};

constexpr int kStackIssuerB = 18;   // warp index of the second MMA issuer

__device__ __forceinline__ uint32_t pack2s(float a, float b, int operand) {
  if (operand == MS_BF16) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

// BF: bf16 operands (compile-time: a runtime format makes ptxas emit BOTH predicated F2FP
// variants in the epilogue, which is instruction-issue-bound)
template <int C, bool PAIR, bool BF>
__device__ __forceinline__ void resstack_body(const StackParams& p) {
  constexpr int kOp = BF ? MS_BF16 : MS_F16;
  using G = StackGeom<C>;
  constexpr int MB = G::MB, HB = G::HB, R = G::R, EW = G::EW;
  constexpr int NSLOT = PAIR ? G::NSLOT_P : G::NSLOT;
  constexpr int TAPB = PAIR ? G::TAP_BYTES_P : G::TAP_BYTES;   // bytes of a tap staged per CTA
  constexpr int NB = PAIR ? C / 2 : C;                         // B rows staged per CTA
  constexpr int NP = G::NP;
  static_assert(MB >= 2 && MB == NP * HB && NP <= 8, "tile must split into NP parts");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // [0,18) wfull  [18,36) wempty  [36,44) acc_full  [44,52) act_ready  [52,70) peer wfull
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 640);
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t sX = smem_u32(smem + kStackHeader);
  const uint32_t sY = sX + G::ACT_BYTES;
  const uint32_t sW = sY + G::ACT_BYTES;
  float* sMono = reinterpret_cast<float*>(smem + kStackHeader + 2 * G::ACT_BYTES +
                                          G::NSLOT * G::TAP_BYTES);
  float* sP = sMono + G::MONO_BYTES / 4;
  const bool mono = (C == 32) && !PAIR && (p.mono_out != nullptr);
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const bool leader = (rank == 0);
  // tiles: CTA (or CTA pair) j takes tiles j, j + stride, ...; in pair mode the two CTAs of
  // a cluster take adjacent tiles 2j + rank (rank 1's may be a masked dummy past the end)
  const int tile_first = PAIR ? 2 * static_cast<int>(blockIdx.x >> 1) + static_cast<int>(rank)
                              : static_cast<int>(blockIdx.x);
  const int tile_stride = gridDim.x;
  auto pair_has_work = [&](int tile) { return tile - static_cast<int>(rank) < p.total_tiles; };
  auto wfull = [&](int s) { return bar_base + 8u * s; };
  auto wempty = [&](int s) { return bar_base + 8u * (18 + s); };
  auto acc_full = [&](int m) { return bar_base + 8u * (36 + m); };
  auto act_ready = [&](int m) { return bar_base + 8u * (44 + m); };
  auto pwfull = [&](int s) { return bar_base + 8u * (52 + s); };
  const uint32_t a_issued = bar_base + 8u * 70;   // issuer A -> issuer B, once per conv

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 18; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 2);              // released by both MMA issuers
      mbar_init(pwfull(s), 1);
    }
    mbar_init(a_issued, 1);
    for (int m = 0; m < 8; ++m) {
      mbar_init(acc_full(m), 1);
      mbar_init(act_ready(m), (PAIR ? 2 : 1) * EW);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc2(smem_u32(tmem_slot), 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(smem_u32(tmem_slot), 512);
      tmem_relinquish();
    }
  }
  if (mono) {
    // tail-conv weights, transposed to [tap][channel] for vector broadcast loads
    for (int i = threadIdx.x; i < 7 * 32; i += blockDim.x)
      sMono[i] = p.mono_w[(i & 31) * 7 + (i >> 5)];
  }
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // peer barriers initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ======================= weight-tap producer =========================
    uint32_t pos = 0;  // running tap counter (ring position)
    for (int tile = tile_first; pair_has_work(tile); tile += tile_stride) {
      for (int tap = 0; tap < 18; ++tap, ++pos) {
        const int slot = pos % NSLOT;
        const uint32_t par = (pos / NSLOT) & 1u;
        mbar_wait(wempty(slot), par ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(wfull(slot), TAPB);
          // pair layout: [conv][tap][N-half][chunk][C/2][8]
          bulk_g2s(sW + slot * TAPB,
                   reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(tap) * G::TAP_BYTES +
                       (PAIR ? rank * TAPB : 0u),
                   TAPB, wfull(slot));
        }
        __syncwarp();
      }
    }
  } else if (PAIR && warp == 1 && !leader) {
    // ============ relay (rank 1): "my half of tap landed" -> leader's pwfull ============
    uint32_t pos = 0;
    for (int tile = tile_first; pair_has_work(tile); tile += tile_stride) {
      for (int tap = 0; tap < 18; ++tap, ++pos) {
        mbar_wait(wfull(pos % NSLOT), (pos / NSLOT) & 1u);
        if (elect_one()) mbar_arrive_remote_relaxed(pwfull(pos % NSLOT), 0);
        __syncwarp();
      }
    }
  } else if ((warp == 1 || warp == kStackIssuerB) && (!PAIR || leader)) {
    // ======================== MMA issuers (two, one per part) ========================
    // The issue path is one thread's instruction stream; a lone issuer sharing its scheduler
    // with four epilogue warps was busy 100 % of the time at ~85 cycles per MMA (measured on
    // upstack.cu, tools/upstack_trace.py) while an N <= 128 MMA retires in 42-64 cycles.  Each
    // accumulator block is written by exactly one issuer (fixed summation order); weight taps
    // are released by both (wempty count 2).
    // CTA-pair kernel: both issuers live in the leader CTA and issue M = 256 MMAs that cover one
    // M-block of each CTA's tile; commits are multicast to both CTAs, act_ready collects both
    // CTAs' epilogue warps (cluster-scope waits), rank 1 relays "my half of the tap landed".
    static_assert(NP == 2, "one MMA issuer warp per part");
    const int part = (warp == 1) ? 0 : 1;
    const uint32_t idesc = PAIR ? umma_idesc_f16_m256(C, kOp) : umma_idesc_f16(C, kOp);
    auto commit_to = [&](uint32_t bar) {
      if (PAIR) umma2_commit_mc(bar); else umma_commit(bar);
    };
    auto wait_act = [&](int m, uint32_t par) {
      if (PAIR) mbar_wait_cluster(act_ready(m), par); else mbar_wait(act_ready(m), par);
    };
    const uint64_t adesc0 = umma_desc_base_nosw(R * 16, 128);
    const uint64_t bdesc0 = umma_desc_base_nosw(NB * 16, 128);
    uint32_t pos = 0;
    uint32_t nconv = 0;
    for (int tile = tile_first; pair_has_work(tile); tile += tile_stride) {
      for (int l = 0; l < 6; ++l, pos += 3, ++nconv) {
        const int d = (l & 1) ? 1 : p.dil[l >> 1];
        const uint32_t src16 = ((l & 1) ? sY : sX) >> 4;
        const uint32_t ready_par = nconv & 1u;
        // tap t on M-blocks [mb0, mb1) (call inside the elected region)
        auto issue = [&](int t, int mb0, int mb1) {
          const uint32_t slot = (pos + t) % NSLOT;
          const uint64_t bd = bdesc0 + ((sW + slot * TAPB) >> 4);
          if (MSB_SABL(1)) return;
          for (int mb = mb0; mb < mb1; ++mb) {
            const uint64_t ad = adesc0 + (src16 + static_cast<uint32_t>(mb * 128 + (t - 1) * d));
            const uint32_t dst = tmem_base + static_cast<uint32_t>(mb * 2 * C + C);
#pragma unroll
            for (int k16 = 0; k16 < C / 16; ++k16) {
              const uint64_t a_k = ad + static_cast<uint64_t>(k16 * 2 * R);
              const uint64_t b_k = bd + static_cast<uint64_t>(k16 * 2 * NB);
              if (PAIR) umma2_f16_ss(dst, a_k, b_k, idesc, 1u);
              else umma_f16_ss(dst, a_k, b_k, idesc, 1u);
            }
          }
        };
        auto wait_tap = [&](int t) {
          const uint32_t q = pos + t;
          mbar_wait(wfull(q % NSLOT), (q / NSLOT) & 1u);
          if (PAIR) mbar_wait_cluster(pwfull(q % NSLOT), (q / NSLOT) & 1u);
        };
        const int m0 = part * HB, m1 = m0 + HB;
        MSB_TRACE(nconv * 16 + part * 4 + 0);
        wait_act(part, ready_par);
        wait_tap(0);
        tc_fence_after();
        MSB_TRACE(nconv * 16 + part * 4 + 1);
        if (part == NP - 1) {
          // issuer B: the tile's last part has no successor (its tap +d runs off the tile edge).
          // It queues behind issuer A's MMAs of this conv: part 0 must COMPLETE first so that
          // its epilogue overlaps this part's MMAs (interleaved, both parts finish together
          // and the tensor pipe idles through the whole first epilogue)
          mbar_wait(a_issued, ready_par);
          if (elect_one()) {
            issue(0, m0, m1);
            commit_to(wempty((pos + 0) % NSLOT));
          }
          __syncwarp();
          wait_tap(1);
          if (elect_one()) {
            issue(1, m0, m1);
            commit_to(wempty((pos + 1) % NSLOT));
          }
          __syncwarp();
          wait_tap(2);
          if (elect_one()) {
            issue(2, m0, m1);
            commit_to(wempty((pos + 2) % NSLOT));
            commit_to(acc_full(part));
          }
          __syncwarp();
        } else {
          // issuer A: tap +d of the part's last M-block reads into the next part; it waits
          // for those rows (and stays with this issuer: fixed summation order per accumulator)
          if (elect_one()) {
            issue(0, m0, m1);
            commit_to(wempty((pos + 0) % NSLOT));
          }
          __syncwarp();
          wait_tap(1);
          if (elect_one()) {
            issue(1, m0, m1);
            commit_to(wempty((pos + 1) % NSLOT));
          }
          __syncwarp();
          wait_tap(2);
          if (elect_one()) issue(2, m0, m1 - 1);
          __syncwarp();
          MSB_TRACE(nconv * 16 + part * 4 + 2);
          wait_act(part + 1, ready_par);
          tc_fence_after();
          if (elect_one()) {
            issue(2, m1 - 1, m1);
            commit_to(wempty((pos + 2) % NSLOT));
            commit_to(acc_full(part));
            mbar_arrive(a_issued);
          }
          __syncwarp();
        }
        MSB_TRACE(nconv * 16 + part * 4 + 3);
      }
    }
  } else if (warp == 1 || warp == kStackIssuerB) {
    // (rank 1 of a CTA pair: warp 18 has no role; warp 1 was the relay above)
  } else {
    // ====================== prologue / epilogue warps ======================
    // Thread = one row (TMEM lane) x COLS channels of IT M-blocks per pipeline part.  The
    // epilogue is instruction-issue bound, so everything that is not arithmetic on the
    // elements is hoisted: the conv index is a compile-time parameter of the loop body
    // (no per-element branches), the bias is not added here at all -- the accumulator
    // columns are PRE-LOADED with the next conv's bias (one tcgen05.st per 16 columns; every
    // MMA accumulates), and addresses are a per-thread base + compile-time offsets.
    const int e = warp - 2;            // 0 .. EW-1
    const int q = warp & 3;            // TMEM lane quarter accessible to this warp
    const int part = (e >> 2) % G::PARTS;    // which slice of the channels
    const int ms = (e >> 2) / G::PARTS;      // which M-blocks of a half
    constexpr int COLS = C / G::PARTS; // channels per thread
    constexpr int NG = COLS / 16;      // 16-column groups per thread
    constexpr int IT = HB / G::MSPLIT; // M-blocks per thread per pipeline part
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int chunk0 = part * (COLS / 8);
    const int row0 = q * 32 + lane;                                     // row within an M-block
    const uint32_t tm0 = tmem_base + lane_off + static_cast<uint32_t>(part * COLS);
    const uint32_t so0 = static_cast<uint32_t>((chunk0 * R + row0) * 16);   // offset in sX / sY
    uint32_t nconv = 0;
    // one arrival per warp on act_ready (local, or on the leader's barrier in pair mode)
    auto arrive_act = [&](int h) {
      __syncwarp();
      if (lane == 0) {
        if (!PAIR || leader) mbar_arrive(act_ready(h));
        // the rows this arrival publishes live in THIS CTA's shared / tensor memory and were
        // made visible to the async proxy by fence.proxy.async + tcgen05 fence above; only the
        // signal crosses CTAs.  (The release.cluster form costs a MEMBAR.ALL.GPU per warp.)
        else mbar_arrive_remote_relaxed(act_ready(h), 0);
      }
    };
    // accumulator columns [ta, ta + COLS) <- bias of conv l (the MMAs of conv l accumulate)
    // accumulator columns [ta + c0, ta + c0 + n) <- bias of conv l (n a multiple of 16)
    auto store_bias_cols = [&](int l, uint32_t ta, int c0, int n) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + l * C + part * COLS + c0);
#pragma unroll
      for (int g = 0; g < n / 16; ++g) {
        uint32_t bv[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 t4 = __ldg(b4 + g * 4 + j4);
          bv[j4 * 4 + 0] = __float_as_uint(t4.x); bv[j4 * 4 + 1] = __float_as_uint(t4.y);
          bv[j4 * 4 + 2] = __float_as_uint(t4.z); bv[j4 * 4 + 3] = __float_as_uint(t4.w);
        }
        tmem_st16p(ta + c0 + g * 16, bv);
      }
    };
    auto store_bias = [&](int l, uint32_t ta) { store_bias_cols(l, ta, 0, COLS); };
    // row (M-block mb, this thread's lane) of the tile with origin tt0 in clip tb: this thread's
    // COLS channels as fp32 bit patterns (zeros outside the clip)
    // columns [c0, c0 + n) of this thread's COLS (n a multiple of 8) into v[0 .. n)
    auto load_cols = [&](int tb, int tt0, int mb, int c0, int n, uint32_t* v) {
      const int t = tt0 + mb * 128 + row0;
      const bool inside = (t >= 0) && (t < p.L);
      const size_t cstride = static_cast<size_t>(p.L) * 8;   // elements per channel chunk
      const size_t off =
          ((static_cast<size_t>(tb) * G::NCH + chunk0 + c0 / 8) * p.L + (inside ? t : 0)) * 8;
      if (p.x16in != nullptr) {
#pragma unroll
        for (int c = 0; c < n / 8; ++c) {
          uint4 q16 = make_uint4(0u, 0u, 0u, 0u);
          if (inside) q16 = __ldg(reinterpret_cast<const uint4*>(p.x16in + off + c * cstride));
          const uint32_t w16[4] = {q16.x, q16.y, q16.z, q16.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float2 f2;
            if (BF) f2 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w16[j]));
            else f2 = __half22float2(*reinterpret_cast<const __half2*>(&w16[j]));
            v[c * 8 + 2 * j] = __float_as_uint(f2.x);
            v[c * 8 + 2 * j + 1] = __float_as_uint(f2.y);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < n / 8; ++c) {
          float a8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (inside) ld_global_nc_v8(p.x32 + off + c * cstride, a8);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[c * 8 + j] = __float_as_uint(a8[j]);
        }
      }
    };
    // ... -> 16-bit operand in sX, fp32 residual stream in TMEM, conv 0's bias in the accumulator
    auto store_cols = [&](int mb, int c0, int n, const uint32_t* v) {
      const uint32_t tx = tm0 + static_cast<uint32_t>(mb * 2 * C);
#pragma unroll
      for (int c = 0; c < n / 8; ++c) {
        const uint32_t dst = sX + so0 + static_cast<uint32_t>(((c0 / 8 + c) * R + mb * 128) * 16);
        st_shared_v4(dst,
                     pack2s(__uint_as_float(v[c * 8 + 0]), __uint_as_float(v[c * 8 + 1]), kOp),
                     pack2s(__uint_as_float(v[c * 8 + 2]), __uint_as_float(v[c * 8 + 3]), kOp),
                     pack2s(__uint_as_float(v[c * 8 + 4]), __uint_as_float(v[c * 8 + 5]), kOp),
                     pack2s(__uint_as_float(v[c * 8 + 6]), __uint_as_float(v[c * 8 + 7]), kOp));
      }
#pragma unroll
      for (int g = 0; g < n / 16; ++g) tmem_st16p(tx + c0 + g * 16, &v[g * 16]);
      store_bias_cols(0, tx + C, c0, n);
    };
    auto load_row = [&](int tb, int tt0, int mb, uint32_t* v) { load_cols(tb, tt0, mb, 0, COLS, v); };
    auto store_row = [&](int mb, const uint32_t* v) { store_cols(mb, 0, COLS, v); };
    auto tile_origin = [&](int tl, int& tb, int& tt0) {
      const bool lv = tl < p.total_tiles;          // false: rank 1's dummy tile of an odd pair
      tb = lv ? tl / p.tiles_per_clip : 0;
      // a dummy tile sits entirely before the clip: every row is masked to zero
      tt0 = lv ? (tl % p.tiles_per_clip) * p.V - p.halo : -2 * R;
    };
    for (int tile = tile_first; pair_has_work(tile); tile += tile_stride) {
      const bool live = tile < p.total_tiles;
      int b, t0;
      tile_origin(tile, b, t0);
      const bool edge = (t0 < 0) || (t0 + R > p.L);                 // warp-uniform
      // the NEXT tile's prologue is folded into this tile's last conv epilogue (below): its
      // global loads are issued before the wait for the last accumulator, its operand / stream /
      // bias stores follow each row's output, so the MMA warp never idles between tiles
      const bool has_next = G::MERGE_PROLOGUE && pair_has_work(tile + tile_stride);
      int nb = 0, nt0 = 0;
      if (has_next) tile_origin(tile + tile_stride, nb, nt0);
      // ---- stand-alone prologue (first tile of this CTA only):
      //      x32 (global) -> TMEM residual stream + 16-bit operand in sX
      if (warp == 2) MSB_TRACE(480 + (nconv / 6) * 2);
      if (tile == tile_first || !G::MERGE_PROLOGUE) {
        for (int h = 0; h < NP; ++h) {
#pragma unroll
          for (int u = 0; u < IT; ++u) {
            const int mb = h * HB + ms + u * G::MSPLIT;
            uint32_t v[COLS];
            load_row(b, t0, mb, v);
            store_row(mb, v);
          }
          tmem_st_wait();
          fence_proxy_async_smem();
          tc_fence_before();
          arrive_act(h);
        }
      }
      if (warp == 2) MSB_TRACE(480 + (nconv / 6) * 2 + 1);
      // ---- six convolutions: conv_epi<second conv of an atom, last conv of the stack>(l)
      auto conv_epi = [&](auto second_t, auto last_t, const int l) {
        constexpr bool second = decltype(second_t)::value;   // residual add
        constexpr bool last = decltype(last_t)::value;
        const uint32_t dstbuf = second ? sX : sY;
        if (l == 4) {
          // warm L2 with the next tile's input while this tile finishes
          const int ntile = tile + tile_stride;
          if (ntile < p.total_tiles && (lane & 3) == 0) {
            const int nb = ntile / p.tiles_per_clip;
            const int nt0 = (ntile % p.tiles_per_clip) * p.V - p.halo;
            for (int mb = ms; mb < MB; mb += G::MSPLIT) {
              const int t = nt0 + mb * 128 + row0;
              if (t >= 0 && t < p.L) {
#pragma unroll
                for (int c = 0; c < COLS / 8; ++c) {
                  const size_t e = ((static_cast<size_t>(nb) * G::NCH + chunk0 + c) * p.L + t) * 8;
                  if (p.x16in != nullptr) prefetch_l2(p.x16in + e); else prefetch_l2(p.x32 + e);
                }
              }
            }
          }
        }
        for (int h = 0; h < NP; ++h) {
          if (warp == 2) MSB_TRACE(256 + nconv * 16 + h * 4 + 0);
          // the next tile's rows travel through registers 16 columns at a time: group 0 is
          // requested before the wait for this conv's accumulator, group k+1 while group k of the
          // current tile is being finished (64 live registers per row would not fit)
          uint32_t nx[last ? IT : 1][last ? 16 : 1];
          if (last && has_next) {
#pragma unroll
            for (int u = 0; u < IT; ++u)
              load_cols(nb, nt0, h * HB + ms + u * G::MSPLIT, 0, 16, nx[last ? u : 0]);
          }
          mbar_wait(acc_full(h), nconv & 1u);
          tc_fence_after();
          if (warp == 2) MSB_TRACE(256 + nconv * 16 + h * 4 + 1);
#pragma unroll
          for (int u = 0; u < IT; ++u) {
            if (MSB_SABL(2)) continue;
            const int mb = h * HB + ms + u * G::MSPLIT;
            const int row = mb * 128 + row0;
            const int t = t0 + row;
            const uint32_t tx = tm0 + static_cast<uint32_t>(mb * 2 * C);
            const uint32_t ta = tx + C;
            const bool zero_row = edge && (t < 0 || t >= p.L);
            // the last conv also carries the next tile's row in registers: it works through its
            // columns 16 at a time to stay inside the register budget
            constexpr int GS = last ? 16 : COLS;
#pragma unroll
            for (int c0 = 0; c0 < COLS; c0 += GS) {
              uint32_t v[GS];
              if (MSB_SABL(8)) {
#pragma unroll
                for (int j = 0; j < GS; ++j) v[j] = static_cast<uint32_t>(t + j);
              } else {
#pragma unroll
                for (int g = 0; g < GS / 16; ++g) tmem_ld16p(ta + c0 + g * 16, &v[g * 16]);
              }
              float f[GS];
              if (second) {
                uint32_t xr[GS];
                if (MSB_SABL(8)) {
#pragma unroll
                  for (int j = 0; j < GS; ++j) xr[j] = static_cast<uint32_t>(t - j);
                } else {
#pragma unroll
                  for (int g = 0; g < GS / 16; ++g) tmem_ld16p(tx + c0 + g * 16, &xr[g * 16]);
                }
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < GS; j += 2) {
                  float l0, l1;
                  leaky02x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), l0, l1);
                  add_x2(__uint_as_float(xr[j]), __uint_as_float(xr[j + 1]), l0, l1, f[j], f[j + 1]);
                }
              } else {
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < GS; j += 2)
                  leaky02x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), f[j], f[j + 1]);
              }
              if (zero_row) {
#pragma unroll
                for (int j = 0; j < GS; ++j) f[j] = 0.f;
              }
              if (!last) {
                if (second && !MSB_SABL(8)) {
#pragma unroll
                  for (int j = 0; j < GS; ++j) v[j] = __float_as_uint(f[j]);
#pragma unroll
                  for (int g = 0; g < GS / 16; ++g) tmem_st16p(tx + c0 + g * 16, &v[g * 16]);
                }
#pragma unroll
                for (int c = 0; c < GS / 8; ++c) {
                  if (MSB_SABL(4)) continue;
                  const uint32_t dst =
                      dstbuf + so0 + static_cast<uint32_t>(((c0 / 8 + c) * R + mb * 128) * 16);
                  st_shared_v4(dst, pack2s(f[c * 8 + 0], f[c * 8 + 1], kOp),
                               pack2s(f[c * 8 + 2], f[c * 8 + 3], kOp),
                               pack2s(f[c * 8 + 4], f[c * 8 + 5], kOp),
                               pack2s(f[c * 8 + 6], f[c * 8 + 7], kOp));
                }
              } else if (mono) {
                // fused tail, step 1 (C == 32: GS == COLS == 16): this thread's 16 channels of
                // its row contribute 7 per-tap partial dot products, P[part][tap][row]
                float pk[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                  float a0 = 0.f, a1 = 0.f;
#pragma unroll
                  for (int j4 = 0; j4 < GS / 4; ++j4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(sMono + k * 32 + part * COLS + c0 + j4 * 4);
                    fma_x2(f[j4 * 4 + 0], f[j4 * 4 + 1], w4.x, w4.y, a0, a1);
                    fma_x2(f[j4 * 4 + 2], f[j4 * 4 + 3], w4.z, w4.w, a0, a1);
                  }
                  pk[k] = a0 + a1;
                }
                float* P = sP + (part * 7) * R + row;
#pragma unroll
                for (int k = 0; k < 7; ++k) P[k * R] = pk[k];
              } else if (live && t >= 0 && t < p.L && row >= p.halo && row < R - p.halo) {
#pragma unroll
                for (int c = 0; c < GS / 8; ++c) {
                  const size_t idx = (static_cast<size_t>(b) * G::NCH + chunk0 + c0 / 8 + c) * p.L + t;
                  if (p.y16 != nullptr)
                    *reinterpret_cast<uint4*>(p.y16 + idx * 8) =
                        make_uint4(pack2s(f[c * 8 + 0], f[c * 8 + 1], kOp),
                                   pack2s(f[c * 8 + 2], f[c * 8 + 3], kOp),
                                   pack2s(f[c * 8 + 4], f[c * 8 + 5], kOp),
                                   pack2s(f[c * 8 + 6], f[c * 8 + 7], kOp));
                  if (p.y32 != nullptr) {
                    const float o8[8] = {f[c * 8 + 0], f[c * 8 + 1], f[c * 8 + 2], f[c * 8 + 3],
                                         f[c * 8 + 4], f[c * 8 + 5], f[c * 8 + 6], f[c * 8 + 7]};
                    st_global_v8(p.y32 + idx * 8, o8);
                  }
                }
              }
              // this row's columns [c0, c0 + 16) are done with the old tile: bring in the next tile's
              if (last && has_next) {
                store_cols(mb, c0, 16, nx[last ? u : 0]);
                if (c0 + 16 < COLS) load_cols(nb, nt0, mb, c0 + 16, 16, nx[last ? u : 0]);
              }
            }
            // the accumulator has been read: seed it with the next conv's bias
            if (!last && !MSB_SABL(8) && !MSB_SABL(16)) store_bias(l + 1, ta);
          }
          if (!last || has_next) {
            tmem_st_wait();
            fence_proxy_async_smem();
            tc_fence_before();
            arrive_act(h);
          }
          if (warp == 2) MSB_TRACE(256 + nconv * 16 + h * 4 + 2);
        }
        ++nconv;
      };
      using T_ = std::true_type;
      using F_ = std::false_type;
      conv_epi(F_{}, F_{}, 0);
      conv_epi(T_{}, F_{}, 1);
      conv_epi(F_{}, F_{}, 2);
      conv_epi(T_{}, F_{}, 3);
      conv_epi(F_{}, F_{}, 4);
      conv_epi(T_{}, T_{}, 5);
      if (mono) {
        // ---- fused tail, step 2: y[row] = tanh(b + sum_k sum_part P[part][k][row+k-3])
        named_bar_sync(1, 32 * EW);            // partial sums complete (epilogue warps only)
        const float bias0 = __ldg(p.mono_b);
        const float* P = sP;
        for (int row = threadIdx.x - 64; row < R - p.halo; row += 32 * EW) {
          const int t = t0 + row;
          if (row < p.halo || t < 0 || t >= p.L) continue;
          float a0 = bias0, a1 = 0.f;
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            a0 += P[k * R + row + k - 3];
            a1 += P[(7 + k) * R + row + k - 3];
          }
          p.mono_out[static_cast<size_t>(b) * p.L + t] = tanhf(a0 + a1);
        }
        // no second barrier: no warp can reach the next tile's last conv (the next writer of P)
        // before every warp has passed this point -- the convs in between need all 16 arrivals
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // nobody leaves while the peer may still signal / read us
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int C>
__global__ void __launch_bounds__(StackGeom<C>::THREADS, 1)
resstack_kernel(const __grid_constant__ StackParams p) {
  resstack_body<C, false, false>(p);
}

// CTA-pair variant (cta_group::2): the two CTAs of a cluster work on adjacent tiles in
// lock-step; the leader issues M = 256 MMAs covering one M-block of each tile, each CTA
// stages half of every weight tap.  Used for C = 128, where a single CTA's N = 128 MMAs
// saturate the shared-memory port (A 4 KB + B 4 KB per 64 cycles) and starve the epilogue.
template <int C>
__global__ void __launch_bounds__(StackGeom<C>::THREADS, 1)
resstack_bf16_kernel(const __grid_constant__ StackParams p) {
  resstack_body<C, false, true>(p);
}

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(StackGeom<C>::THREADS_PAIR, 1)
resstack_pair_kernel(const __grid_constant__ StackParams p) {
  resstack_body<C, true, false>(p);
}

// CTA-pair variant for C = 128 (default; MSB_STACK_PAIR=0 selects the single-CTA kernel).  Each CTA
// stages half of every tap, so the 96 KB ring holds two convs of weights instead of one and the
// tensor core reads 6 KB instead of 8 KB of operands per MMA -- the single-CTA kernel is bound by
// the shared-memory port.  Round 1 measured the pair form 8 % slower; the cause was the
// release.cluster form of the remote act_ready arrivals (a MEMBAR.ALL.GPU per epilogue warp and
// hand-off, ~1.2 k cycles on the critical path of every conv).  With relaxed remote arrivals and the
// two-issuer scheme: 579 vs 614 us at 64 clips (tools/stack_bench.py).  fp16 operands only.
bool stack_pair_enabled(int channels) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_STACK_PAIR");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1 && channels == 128;
}

template <int C>
ms_status launch_stack_pair(const StackParams& p, cudaStream_t stream) {
  using G = StackGeom<C>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(resstack_pair_kernel<C>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(resstack_pair_kernel)");
    attr_set = true;
  }
  const int sms = sm_count();
  if (sms <= 1) return check_cuda(cudaGetLastError(), "sm_count");
  const int pairs = (p.total_tiles + 1) / 2;
  const int clusters = pairs < sms / 2 ? pairs : sms / 2;
  resstack_pair_kernel<C><<<2 * clusters, G::THREADS_PAIR, G::SMEM, stream>>>(p);
  return after_launch("resstack_pair_kernel");
}

template <int C>
ms_status launch_stack(const StackParams& p, cudaStream_t stream) {
  using G = StackGeom<C>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(resstack_kernel<C>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(resstack_kernel)");
    attr_set = true;
  }
  const int sms = sm_count();
  if (sms <= 0) return check_cuda(cudaGetLastError(), "sm_count");
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (p.operand == MS_BF16) {
    static thread_local bool attr_set_bf = false;
    if (!attr_set_bf) {
      cudaError_t e = cudaFuncSetAttribute(resstack_bf16_kernel<C>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(resstack_bf16_kernel)");
      attr_set_bf = true;
    }
    resstack_bf16_kernel<C><<<grid, G::THREADS, G::SMEM, stream>>>(p);
    return after_launch("resstack_bf16_kernel");
  }
  resstack_kernel<C><<<grid, G::THREADS, G::SMEM, stream>>>(p);
  return after_launch("resstack_kernel");
}

static thread_local long long* g_stack_dbg = nullptr;

ms_status resstack_fwd(int channels, int batch, int len, const int* dil, int operand,
                       const float* x32, const void* packed, void* y16, float* y32,
                       cudaStream_t stream, const float* mono_w, const float* mono_b,
                       float* mono_out, const void* x16in) {
  if (batch <= 0 || len <= 0 || (x32 == nullptr && x16in == nullptr) || packed == nullptr ||
      (y16 == nullptr && y32 == nullptr && mono_out == nullptr))
    return MS_ERR_INVALID;
  if (mono_out != nullptr && (channels != 32 || mono_w == nullptr || mono_b == nullptr))
    return MS_ERR_INVALID;
  for (int i = 0; i < 3; ++i)
    if (dil[i] < 1 || dil[i] > 9) return MS_ERR_INVALID;
  // receptive field of the stack (+3 for the fused k7 tail) must fit the halo
  const int halo = mono_out != nullptr ? kStackHalo + 4 : kStackHalo;
  if (dil[0] + dil[1] + dil[2] + 3 > kStackHalo) return MS_ERR_INVALID;
  StackParams p;
  p.x32 = x32;
  p.x16in = static_cast<const uint16_t*>(x16in);
  p.w = static_cast<const uint16_t*>(packed);
  p.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) +
                                          static_cast<size_t>(18) * channels * channels * 2);
  p.y16 = static_cast<uint16_t*>(y16);
  p.y32 = y32;
  p.B = batch; p.L = len;
  p.dil[0] = dil[0]; p.dil[1] = dil[1]; p.dil[2] = dil[2];
  p.operand = operand;
  p.mono_w = mono_w; p.mono_b = mono_b; p.mono_out = mono_out;
  p.dbg = g_stack_dbg;
  p.ablate = 0;
#ifdef MSB_STACK_ABLATE
  if (const char* e = getenv("MSB_STACK_ABLATE")) p.ablate = atoi(e);
#endif
  int V;
  switch (channels) {
    case 128: V = StackGeom<128>::R - 2 * halo; break;
    case 64: V = StackGeom<64>::R - 2 * halo; break;
    case 32: V = StackGeom<32>::R - 2 * halo; break;
    default: return MS_ERR_INVALID;
  }
  p.halo = halo; p.V = V;
  p.tiles_per_clip = (len + V - 1) / V;
  const long long tiles = static_cast<long long>(batch) * p.tiles_per_clip;
  if (tiles > 0x7fffffffLL) return MS_ERR_INVALID;
  p.total_tiles = static_cast<int>(tiles);
  switch (channels) {
    case 128:
      return (stack_pair_enabled(128) && p.operand != MS_BF16) ? launch_stack_pair<128>(p, stream)
                                     : launch_stack<128>(p, stream);
    case 64: return launch_stack<64>(p, stream);
    default: return launch_stack<32>(p, stream);
  }
}

// packed[conv][tap][chunk][n][e] <- w_l (C, C, 3) fp32, then 6 x C fp32 biases
__global__ void pack_stack_weight_kernel(const float* __restrict__ w, uint16_t* __restrict__ out,
                                         int C, int operand, int pair) {
  const int total = 3 * C * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int e = i % 8;
  int n, c;
  const int t = i / (C * C);
  if (pair) {   // [tap][N-half][chunk][C/2][8]
    const int nn = (i / 8) % (C / 2);
    c = (i / (8 * (C / 2))) % (C / 8);
    const int half = (i / (8 * (C / 2) * (C / 8))) % 2;
    n = half * (C / 2) + nn;
  } else {      // [tap][chunk][C][8]
    n = (i / 8) % C;
    c = (i / (8 * C)) % (C / 8);
  }
  const float v = w[(static_cast<size_t>(n) * C + c * 8 + e) * 3 + t];
  if (operand == MS_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  } else {
    __half h = __float2half_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

}  // namespace msb

using namespace msb;

extern "C" {

/* debugging aid (not in the public header): clock64 trace buffer (512 x int64, device)
 * filled by CTA 0 of subsequent ms_resstack_fwd launches from this thread */
void ms_debug_set_stack_trace(long long* dev_buf) { msb::g_stack_dbg = dev_buf; }

int ms_resstack_supported(int channels) {
  return channels == 128 || channels == 64 || channels == 32;
}

size_t ms_resstack_packed_weight_bytes(int channels) {
  if (!ms_resstack_supported(channels)) return 0;
  return static_cast<size_t>(18) * channels * channels * 2 + sizeof(float) * 6 * channels;
}

ms_status ms_resstack_pack_weights(const float* const* params, int channels, int operand,
                                   void* packed, void* stream) {
  if (params == nullptr || packed == nullptr || !ms_resstack_supported(channels))
    return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(packed);
  const size_t conv_bytes = static_cast<size_t>(3) * channels * channels * 2;
  float* bias = reinterpret_cast<float*>(base + 6 * conv_bytes);
  const int total = 3 * channels * channels;
  for (int l = 0; l < 6; ++l) {
    pack_stack_weight_kernel<<<(total + 255) / 256, 256, 0, st>>>(
        params[2 * l], reinterpret_cast<uint16_t*>(base + l * conv_bytes), channels, operand,
        (stack_pair_enabled(channels) && operand != MS_BF16) ? 1 : 0);
    ms_status s = after_launch("pack_stack_weight_kernel");
    if (s != MS_OK) return s;
    s = check_cuda(cudaMemcpyAsync(bias + l * channels, params[2 * l + 1],
                                   sizeof(float) * channels, cudaMemcpyDeviceToDevice, st),
                   "cudaMemcpyAsync(stack bias)");
    if (s != MS_OK) return s;
  }
  return MS_OK;
}

ms_status ms_resstack_fwd(int channels, int batch, int len, const int* dilations, int operand,
                          const float* x32, const void* packed, void* y16, float* y32,
                          void* stream) {
  if (dilations == nullptr || !ms_resstack_supported(channels)) return MS_ERR_INVALID;
  return resstack_fwd(channels, batch, len, dilations, operand, x32, packed, y16, y32,
                      static_cast<cudaStream_t>(stream), nullptr, nullptr, nullptr, nullptr);
}

ms_status ms_resstack_tail_fwd(int batch, int len, const int* dilations, int operand,
                               const float* x32, const void* packed, const float* tail_w,
                               const float* tail_b, float* y, void* stream) {
  if (dilations == nullptr || tail_w == nullptr || tail_b == nullptr || y == nullptr)
    return MS_ERR_INVALID;
  return resstack_fwd(32, batch, len, dilations, operand, x32, packed, nullptr, nullptr,
                      static_cast<cudaStream_t>(stream), tail_w, tail_b, y, nullptr);
}

}  // extern "C"
