// Fused upsampling stage: ConvTranspose1d(2C -> C, k = 4, stride 2, pad 1) + LeakyReLU(0.2) +
// ResidualStack(C) [+ the generator's 32 -> 1 k7 conv + tanh] in ONE kernel.
//   replaces generator/full.py:35-37 (C = 64), 39-44 (C = 32) and
//   ResidualStack.forward / ResidualAtom.forward, util/modules.py:350-405.
//
// The upsampler's fp32 output (the residual stream the stack starts from) never exists in
// HBM: it is produced by tcgen05 MMAs straight into the stack's accumulator columns in tensor
// memory, and the stage reads only the previous stage's 16-bit operand image (2 bytes per
// input element instead of a 4-byte write + 4-byte read per OUTPUT element).
//
// Phase-split tile.  A tile covers R = MB*128 output rows t = t0 + 2m + r (t0 even).  Even rows
// (r = 0, "E") and odd rows (r = 1, "O") live in separate 128-row M-blocks: SMEM operand row
// r*R/2 + m, TMEM lane m % 128.  Why: the polyphase form of the transposed convolution,
//     out[2u]     = x[u] W[..,1] + x[u-1] W[..,3]
//     out[2u + 1] = x[u+1] W[..,0] + x[u] W[..,2],
// makes each phase a plain GEMM over the INPUT-rate rows u, so with phase-split M-blocks its
// result lands in exactly the TMEM lanes the stack uses -- no cross-lane shuffle.  The stack's
// dilations (1, 3, 9 and the inner 1) are all odd, so a tap +-d of an E block reads a shifted
// window of the O rows and vice versa: still "the same buffer, descriptor start advanced".
//     E block, m0:  -d -> O[m0 - (d+1)/2]   0 -> E[m0]   +d -> O[m0 + (d-1)/2]
//     O block, m0:  -d -> E[m0 - (d-1)/2]   0 -> O[m0]   +d -> E[m0 + (d+1)/2]
//
// Per tile seven pipelined stages (MMA warp <-> 16 epilogue warps, two parts per tile exactly as
// in resstack.cu): stage 0 = the transposed conv (input: R/2 + 2 low-rate rows x 2C channels,
// brought in by bulk async copies into the activation buffer the stack is not using), stages
// 1..6 = the six k3 convs.  Buffers alternate per tile: tile i's input lands in buf[i & 1]
// while tile i-1's stage 6 still reads the other one.
// Rows: the low-rate window holds R/2 rows (q0 = t0/2 - 1 onwards), two short of what the last
// three output rows need; together with the stack's reach of 16 that costs 4 stored rows per
// tile (V = R - 2*halo - 4).
#include <cstdlib>
#include <type_traits>

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "ptx.cuh"
#include "runtime.cuh"

namespace msb {

namespace {

constexpr int kUpHalo = 16;
constexpr int kUpHeader = 1024;
constexpr int kUpStages = 7;            // transposed conv + six convs
constexpr int kUpEntries = 8 + 18;      // weight images per tile (C x C 16-bit each)
constexpr int kUpFills = 3 + 6;         // weight-slot fills per tile (3 images each; one of 2)
constexpr int kIssuerB = 18;            // warp index of the second MMA issuer

struct UpStackParams {
  const uint16_t* x16;   // BLK 16-bit (B, 2C/8, lin, 8): previous stage's operand image
  const uint16_t* w;     // [26][C/8][C][8] 16-bit: 8 transposed-conv images, 18 conv taps
  const float* bias;     // [7][C]: transposed conv, six convs
  uint16_t* y16;         // BLK 16-bit (B, C/8, L, 8) or null
  float* y32;            // BLK f32 or null
  int B, lin, L;         // L = 2 * lin
  int dil[3];
  int tiles_per_clip, total_tiles;
  int halo, V;
  const float* mono_w;   // fused tail (C == 32): (1, 32, 7) fp32
  const float* mono_b;
  float* mono_out;       // (B, 1, L) fp32
  long long* dbg;        // MSB_UP_ABLATE builds only: clock trace buffer (256 x int64) or null
  int ablate;            // MSB_UP_ABLATE builds only (tools/upstack_bench.py): bit 0 no MMAs,
                         // 1 empty epilogue, 2 no operand stores, 3 no tensor-memory traffic
};

#ifdef MSB_UP_ABLATE
#define MSB_ABL(bit) ((p.ablate & (bit)) != 0)
// clock64 trace of CTA 0's third tile (tools/upstack_trace.py)
#define MSB_UTRACE(slot)                                                              \
  do {                                                                                \
    if (p.dbg != nullptr && blockIdx.x == 0 && it == 2 && (threadIdx.x & 31) == 0)    \
      p.dbg[(slot)] = clock64();                                                      \
  } while (0)
#else
#define MSB_ABL(bit) false
#define MSB_UTRACE(slot) do { } while (0)
#endif

template <int C>
struct UpGeom {
  static constexpr int MB = 256 / C;          // M-blocks per tile
  static constexpr int NP = 2;                // pipeline parts per tile
  static constexpr int HB = MB / NP;          // M-blocks per part
  static constexpr int HE = HB / 2;           // E (and O) blocks per part
  static constexpr int MSPLIT = 2;            // M-block split of the epilogue
  static constexpr int PARTS = 2;             // column split of the epilogue
  static constexpr int EW = 16;               // epilogue warps = 4 * PARTS * MSPLIT
  static constexpr int R = MB * 128;          // output rows per tile
  static constexpr int RH = R / 2;            // rows per phase = low-rate rows staged
  static constexpr int NCH = C / 8;
  static constexpr int ACT_BYTES = R * C * 2; // 65536
  static constexpr int TAP_BYTES = C * C * 2;
  static constexpr int SLOT_BYTES = 3 * TAP_BYTES;   // a weight slot = the three taps of a conv
  static constexpr int NSLOT = 4;                    // 96 KB at C = 64, 24 KB at C = 32
  static constexpr int MONO_BYTES = 1024;
  static constexpr int P_BYTES = (C == 32) ? 14 * R * 4 : 0;
  static constexpr int SMEM = kUpHeader + 2 * ACT_BYTES + NSLOT * SLOT_BYTES + MONO_BYTES + P_BYTES;
  static constexpr int THREADS = 64 + 32 * EW + 32;  // producer, issuer A, epilogue, issuer B
  static_assert(HB >= 2 && HB % 2 == 0, "a part needs E and O blocks");
};

template <bool BF>
__device__ __forceinline__ uint32_t pack2u(float a, float b) {
  if (BF) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_h2(a, b);
}

template <int C, bool BF>
__device__ __forceinline__ void upstack_body(const UpStackParams& p) {
  constexpr int kOp = BF ? MS_BF16 : MS_F16;
  using G = UpGeom<C>;
  constexpr int HB = G::HB, HE = G::HE, R = G::R, RH = G::RH, EW = G::EW, NP = G::NP;
  constexpr int NSLOT = G::NSLOT, TAPB = G::TAP_BYTES, SLOTB = G::SLOT_BYTES;
  static_assert(NP == 2, "one MMA issuer warp per part");
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  // [0,4) wfull  [18,22) wempty  [36,38) acc_full  [38,40) act_ready  40 in_full  41 in_free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 640);
  const uint32_t bar_base = smem_u32(bars);
  const uint32_t sBuf0 = smem_u32(smem + kUpHeader);
  const uint32_t sW = sBuf0 + 2 * G::ACT_BYTES;
  float* sMono = reinterpret_cast<float*>(smem + kUpHeader + 2 * G::ACT_BYTES + NSLOT * SLOTB);
  float* sP = sMono + G::MONO_BYTES / 4;
  const bool mono = (C == 32) && (p.mono_out != nullptr);
  auto wfull = [&](int s) { return bar_base + 8u * s; };
  auto wempty = [&](int s) { return bar_base + 8u * (18 + s); };
  auto acc_full = [&](int m) { return bar_base + 8u * (36 + m); };
  auto act_ready = [&](int m) { return bar_base + 8u * (38 + m); };
  const uint32_t in_full = bar_base + 8u * 40;
  const uint32_t in_free = bar_base + 8u * 41;
  const uint32_t a_issued = bar_base + 8u * 42;   // issuer A -> issuer B, once per stage
  // HE == 2 (C = 32): the first E / O blocks of part 1 are done -- all that the deferred taps of
  // part 0 read; issuer A no longer waits for the whole part
  const uint32_t act_half = bar_base + 8u * 43;
  auto buf = [&](int i) { return sBuf0 + static_cast<uint32_t>(i & 1) * G::ACT_BYTES; };

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSLOT; ++s) {
      mbar_init(wfull(s), 1);
      mbar_init(wempty(s), 1);     // released by epilogue warp 2 (it has seen both acc_full)
    }
    for (int m = 0; m < NP; ++m) {
      mbar_init(acc_full(m), 1);
      mbar_init(act_ready(m), EW);
    }
    mbar_init(in_full, 1);
    mbar_init(a_issued, 1);
    mbar_init(act_half, EW);
    mbar_init(in_free, 1);         // epilogue warp 2, once stage 5's MMAs have completed
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  if (mono) {
    for (int i = threadIdx.x; i < 7 * 32; i += blockDim.x)
      sMono[i] = p.mono_w[(i & 31) * 7 + (i >> 5)];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int tile_first = static_cast<int>(blockIdx.x);
  const int tile_stride = static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ================= producer: low-rate input window + weight slots =================
    constexpr int CH2 = 2 * C / 8;    // 16-byte channel chunks of the input
    uint32_t pos = 0;                 // weight slots filled so far
    int it = 0;
    for (int tile = tile_first; tile < p.total_tiles; tile += tile_stride, ++it) {
      const int b = tile / p.tiles_per_clip;
      const int t0 = (tile % p.tiles_per_clip) * p.V - p.halo;     // even
      const int q0 = (t0 >> 1) - 1;                                // low-rate row of window row 0
      const int lo = q0 < 0 ? 0 : q0;
      const int hi = (q0 + RH) > p.lin ? p.lin : (q0 + RH);
      const int nrows = hi > lo ? hi - lo : 0;
      const uint32_t dstb = buf(it);
      // the buffer was last read by stage 5 of this CTA's previous tile
      MSB_UTRACE(240);
      mbar_wait(in_free, static_cast<uint32_t>(it & 1) ^ 1u);
      MSB_UTRACE(241);
      if (nrows != RH) {
        // rows outside [0, lin): the transposed conv sees zeros there
        const int head = lo - q0;
        const int tail0 = head + nrows;
        const int nz = head + (RH - tail0);
        for (int i = lane; i < nz * CH2; i += 32) {
          const int c = i / nz;
          int r = i - c * nz;
          r = r < head ? r : tail0 + (r - head);
          st_shared_v4(dstb + static_cast<uint32_t>(c * RH + r) * 16u, 0u, 0u, 0u, 0u);
        }
        fence_proxy_async_smem();
        __syncwarp();
      }
      if (lane == 0)
        mbar_arrive_expect_tx(in_full, static_cast<uint32_t>(nrows) * 16u * CH2);
      __syncwarp();
      if (lane < CH2 && nrows > 0) {
        const uint16_t* src =
            p.x16 + ((static_cast<size_t>(b) * CH2 + lane) * p.lin + lo) * 8;
        bulk_g2s(dstb + static_cast<uint32_t>(lane * RH + (lo - q0)) * 16u, src,
                 static_cast<uint32_t>(nrows) * 16u, in_full);
      }
      __syncwarp();
      // nine slot fills per tile: images 0-2, 3-5, 6-7 (transposed conv), then one conv each
      for (int f = 0; f < kUpFills; ++f, ++pos) {
        const int slot = pos & (NSLOT - 1);
        const uint32_t par = (pos / NSLOT) & 1u;
        const uint32_t bytes = (f == 2 ? 2u : 3u) * TAPB;
        const int first = f < 3 ? 3 * f : 8 + 3 * (f - 3);         // first image of the fill
        mbar_wait(wempty(slot), par ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(wfull(slot), bytes);
          bulk_g2s(sW + slot * SLOTB,
                   reinterpret_cast<const uint8_t*>(p.w) + static_cast<size_t>(first) * TAPB,
                   bytes, wfull(slot));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1 || warp == kIssuerB) {
    // ============================ MMA issuers (two) ============================
    // One warp per pipeline part: the issue path is a single thread's instruction stream
    // (descriptor arithmetic + tcgen05.mma), and with N = C <= 64 an MMA retires in 42-48
    // cycles -- one issuer cannot keep the tensor pipe fed while sharing its scheduler with
    // four epilogue warps (measured: the lone issuer was busy 100 % of the time, 85 cycles per
    // MMA).  Each accumulator block is written by exactly one issuer; tcgen05.commit tracks
    // the issuing thread's own MMAs, so weight slots and the input window are released by
    // both (barrier count 2).
    const int part = (warp == 1) ? 0 : 1;
    const uint32_t idesc = umma_idesc_f16(C, kOp);
    const uint64_t adesc_up = umma_desc_base_nosw(RH * 16, 128);   // input window [2C/8][RH][8]
    const uint64_t adesc0 = umma_desc_base_nosw(R * 16, 128);      // activations [C/8][R][8]
    const uint64_t bdesc0 = umma_desc_base_nosw(C * 16, 128);
    uint32_t pos = 0;     // weight slot of the current stage
    uint32_t g = 0;       // stages issued so far (parity of the act_ready waits)
    int it = 0;
    auto wait_w = [&](uint32_t q) { mbar_wait(wfull(q & (NSLOT - 1)), (q / NSLOT) & 1u); };
    auto slot_addr = [&](uint32_t q) { return (sW + (q & (NSLOT - 1)) * SLOTB) >> 4; };
    // tap image `img` of weight slot `wq` on block (r, kk) of part `pt`, A rows from `arow`
    auto mma_block = [&](uint64_t ad, uint32_t wq16, int img, int mb, int kstepA) {
      const uint64_t bd = bdesc0 + (wq16 + static_cast<uint32_t>(img * (TAPB >> 4)));
      const uint32_t dst = tmem_base + static_cast<uint32_t>(mb * 2 * C + C);
#pragma unroll
      for (int k16 = 0; k16 < C / 16; ++k16)
        umma_f16_ss(dst, ad + static_cast<uint64_t>(k16 * kstepA),
                    bd + static_cast<uint64_t>(k16 * 2 * C), idesc, 1u);
    };
    for (int tile = tile_first; tile < p.total_tiles; tile += tile_stride, ++it) {
      const uint32_t inb = buf(it), oth = buf(it + 1);
      // ---------------- stage 0: transposed conv (image e: phase e >> 2, tap, K half) --------
      MSB_UTRACE(243 + 2 * part);
      mbar_wait(in_full, static_cast<uint32_t>(it & 1));
      MSB_UTRACE(244 + 2 * part);
      MSB_UTRACE(part * 4 + 0);
      mbar_wait(act_ready(part), g & 1u);
      wait_w(pos);
      wait_w(pos + 1);
      wait_w(pos + 2);
      tc_fence_after();
      // issuer B queues behind issuer A's MMAs of the stage: part 0 has to COMPLETE first so
      // that its epilogue overlaps part 1's MMAs (interleaved, both parts would finish together
      // and the tensor pipe would idle through the whole first epilogue)
      if (part == 1) mbar_wait(a_issued, g & 1u);
      MSB_UTRACE(part * 4 + 1);
      if (elect_one()) {
        if (!MSB_ABL(1)) {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int r = e >> 2, tapsel = (e >> 1) & 1, kh = e & 1;
            const int shift = (r ? 2 : 1) - tapsel;
            const uint32_t wq16 = slot_addr(pos + e / 3);
#pragma unroll
            for (int kk = 0; kk < HE; ++kk) {
              const int mb = part * HB + r * HE + kk;
              const int k = part * HE + kk;
              const uint64_t ad =
                  adesc_up + ((inb >> 4) + static_cast<uint32_t>(kh * (C / 8) * RH + k * 128 + shift));
              mma_block(ad, wq16, e % 3, mb, 2 * RH);
            }
          }
        }
        umma_commit(acc_full(part));
        if (part == 0) mbar_arrive(a_issued);
      }
      __syncwarp();
      MSB_UTRACE(part * 4 + 2);
      pos += 3;
      ++g;
      // ---------------- stages 1..6: the six k3 convs ----------------
      for (int l = 0; l < 6; ++l, ++pos, ++g) {
        const int d = (l & 1) ? 1 : p.dil[l >> 1];
        const uint32_t src16 = ((l & 1) ? inb : oth) >> 4;
        // source row (relative to the block's first row) of tap t for E / O blocks
        const int offE[3] = {RH - (d + 1) / 2, 0, RH + (d - 1) / 2};
        const int offO[3] = {-(d - 1) / 2, RH, (d + 1) / 2};
        MSB_UTRACE((l + 1) * 16 + part * 4 + 0);
        mbar_wait(act_ready(part), g & 1u);
        wait_w(pos);
        tc_fence_after();
        MSB_UTRACE((l + 1) * 16 + part * 4 + 1);
        const uint32_t wq16 = slot_addr(pos);
        // taps [t0_, t1_) on blocks [kk0, kk1) of both phases of part pt
        auto issue_part = [&](int pt, int t0_, int t1_, int kk0, int kk1) {
#pragma unroll
          for (int t = t0_; t < t1_; ++t) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
#pragma unroll
              for (int kk = kk0; kk < kk1; ++kk) {
                const int mb = pt * HB + r * HE + kk;
                const int row = (pt * HE + kk) * 128 + (r ? offO[t] : offE[t]);
                mma_block(adesc0 + (src16 + static_cast<uint32_t>(row)), wq16, t, mb, 2 * R);
              }
            }
          }
        };
        if (part == NP - 1) {
          mbar_wait(a_issued, g & 1u);
          if (elect_one()) {
            if (!MSB_ABL(1)) issue_part(1, 0, 3, 0, HE);
            umma_commit(acc_full(1));
          }
          __syncwarp();
        } else {
          // everything except tap +d of the part's last E and O blocks, which reads into the
          // next part: those two wait for the next part's operand rows.  (They stay with this
          // issuer: one thread per accumulator keeps the fp32 summation order fixed.)
          if (elect_one() && !MSB_ABL(1)) {
            issue_part(0, 0, 2, 0, HE);
            if (HE > 1) issue_part(0, 2, 3, 0, HE - 1);
          }
          __syncwarp();
          MSB_UTRACE((l + 1) * 16 + part * 4 + 2);
          mbar_wait(HE == 2 ? act_half : act_ready(1), g & 1u);
          tc_fence_after();
          if (elect_one()) {
            if (!MSB_ABL(1)) issue_part(0, 2, 3, HE - 1, HE);
            umma_commit(acc_full(0));
            mbar_arrive(a_issued);
          }
          __syncwarp();
        }
        MSB_UTRACE((l + 1) * 16 + part * 4 + 3);
      }
    }
  } else {
    // ========================== epilogue warps (16) ==========================
    const int e = warp - 2;
    const int q = warp & 3;                   // TMEM lane quarter of this warp
    const int cpart = (e >> 2) % G::PARTS;    // column slice
    const int ms = (e >> 2) / G::PARTS;       // M-block slice
    constexpr int COLS = C / G::PARTS;
    constexpr int NG = COLS / 16;
    constexpr int IT = HB / G::MSPLIT;
    const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
    const int chunk0 = cpart * (COLS / 8);
    const int row0 = q * 32 + lane;
    const uint32_t tm0 = tmem_base + lane_off + static_cast<uint32_t>(cpart * COLS);
    uint32_t g = 0;
    uint32_t wpos = 0;      // weight slots released so far (warp 2 only)
    auto arrive_act = [&](int h) {
      __syncwarp();
      if (lane == 0) mbar_arrive(act_ready(h));
    };
    auto store_bias = [&](int s, uint32_t ta) {
      const float4* b4 = reinterpret_cast<const float4*>(p.bias + s * C + cpart * COLS);
#pragma unroll
      for (int gg = 0; gg < NG; ++gg) {
        uint32_t bv[16];
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 t4 = __ldg(b4 + gg * 4 + j4);
          bv[j4 * 4 + 0] = __float_as_uint(t4.x); bv[j4 * 4 + 1] = __float_as_uint(t4.y);
          bv[j4 * 4 + 2] = __float_as_uint(t4.z); bv[j4 * 4 + 3] = __float_as_uint(t4.w);
        }
        tmem_st16p(ta + gg * 16, bv);
      }
    };
    // block j of a part: phase r, SMEM row of this thread, time-order row in the tile
    auto geom = [&](int h, int u, int& mb, int& srow, int& trow) {
      // HE == 2: ms selects the phase and u the block of that phase, so that after u = 0 the
      // first E and O blocks of the part are complete (see act_half)
      const int j = (HE == 2) ? ms * HE + u : ms + u * G::MSPLIT;
      const int r = j / HE, kk = j % HE;
      mb = h * HB + j;
      const int m = (h * HE + kk) * 128 + row0;
      srow = r * RH + m;
      trow = 2 * m + r;
    };
    // before the first tile: every accumulator holds the transposed conv's bias
    for (int h = 0; h < NP; ++h) {
#pragma unroll
      for (int u = 0; u < IT; ++u) {
        int mb, srow, trow;
        geom(h, u, mb, srow, trow);
        store_bias(0, tm0 + static_cast<uint32_t>(mb * 2 * C + C));
      }
      tmem_st_wait();
      tc_fence_before();
      if (HE == 2 && h == 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(act_half);
      }
      arrive_act(h);
    }
    int it = 0;
    for (int tile = tile_first; tile < p.total_tiles; tile += tile_stride, ++it) {
      const int b = tile / p.tiles_per_clip;
      const int t0 = (tile % p.tiles_per_clip) * p.V - p.halo;
      const bool edge = (t0 < 0) || (t0 + R > p.L);
      const uint32_t inb = buf(it), oth = buf(it + 1);
      // KIND 0: transposed conv (stream := leaky(acc)), 1: first conv of an atom,
      //      2: second conv (stream += leaky(acc)), 3: last conv of the stack
      auto stage_epi = [&](auto kind_t, const int s) {
        constexpr int KIND = decltype(kind_t)::value;
        constexpr bool resid = KIND >= 2;
        constexpr bool last = KIND == 3;
        const uint32_t dstbuf = (KIND == 1) ? inb : oth;
        for (int h = 0; h < NP; ++h) {
          if (warp == 2) MSB_UTRACE(128 + s * 16 + h * 4 + 0);
          mbar_wait(acc_full(h), g & 1u);
          tc_fence_after();
          if (warp == 2) MSB_UTRACE(128 + s * 16 + h * 4 + 1);
          if (warp == 2 && h == NP - 1) {
            // every MMA of this stage has completed: its weight slot(s) may be refilled, and
            // after stage 5 the buffer `oth` may take the next tile's input window
            if (lane == 0) {
              mbar_arrive(wempty(wpos & (NSLOT - 1)));
              if (KIND == 0) {
                mbar_arrive(wempty((wpos + 1) & (NSLOT - 1)));
                mbar_arrive(wempty((wpos + 2) & (NSLOT - 1)));
              }
              if (s == 5) mbar_arrive(in_free);
            }
            wpos += (KIND == 0) ? 3 : 1;
          }
          if (last && C == 32 && mono) {
            // Last stage with the fused k7 tail (C = 32): each thread takes its lane's row of the
            // E AND the O block (rows 2m, 2m + 1) of one block pair for its 16 channels: the tap
            // weights are read once for both rows and the per-tap partial sums leave as 8-byte
            // pairs (one row per thread meant stride-2 shared-memory writes and twice the reads).
            if (!MSB_ABL(2)) {
              const int m = (h * HE + ms) * 128 + row0;          // block pair kk = ms
              const int trow = 2 * m;
              float f[2][16];
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const int t = t0 + trow + r;
                const uint32_t txr = tmem_base + lane_off +
                                     static_cast<uint32_t>((h * HB + r * HE + ms) * 2 * C + cpart * COLS);
                uint32_t v[16], xr[16];
                tmem_ld16p(txr + C, v);
                tmem_ld16p(txr, xr);
                tmem_ld_wait();
                const bool zero_row = edge && (t < 0 || t >= p.L);
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float l0, l1;
                  leaky02x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), l0, l1);
                  add_x2(__uint_as_float(xr[j]), __uint_as_float(xr[j + 1]), l0, l1, f[r][j], f[r][j + 1]);
                  if (zero_row) { f[r][j] = 0.f; f[r][j + 1] = 0.f; }
                }
                if (!MSB_ABL(8)) {
                  const float4* b4 = reinterpret_cast<const float4*>(p.bias + cpart * COLS);
                  uint32_t bv[16];
#pragma unroll
                  for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 t4 = __ldg(b4 + j4);
                    bv[j4 * 4 + 0] = __float_as_uint(t4.x); bv[j4 * 4 + 1] = __float_as_uint(t4.y);
                    bv[j4 * 4 + 2] = __float_as_uint(t4.z); bv[j4 * 4 + 3] = __float_as_uint(t4.w);
                  }
                  tmem_st16p(txr + C, bv);
                }
              }
              float* P = sP + (cpart * 7) * R + trow;
#pragma unroll
              for (int k = 0; k < 7; ++k) {
                float e0 = 0.f, e1 = 0.f, o0 = 0.f, o1 = 0.f;
#pragma unroll
                for (int j4 = 0; j4 < 4; ++j4) {
                  const float4 w4 =
                      *reinterpret_cast<const float4*>(sMono + k * 32 + cpart * COLS + j4 * 4);
                  fma_x2(f[0][j4 * 4 + 0], f[0][j4 * 4 + 1], w4.x, w4.y, e0, e1);
                  fma_x2(f[0][j4 * 4 + 2], f[0][j4 * 4 + 3], w4.z, w4.w, e0, e1);
                  fma_x2(f[1][j4 * 4 + 0], f[1][j4 * 4 + 1], w4.x, w4.y, o0, o1);
                  fma_x2(f[1][j4 * 4 + 2], f[1][j4 * 4 + 3], w4.z, w4.w, o0, o1);
                }
                *reinterpret_cast<float2*>(P + k * R) = make_float2(e0 + e1, o0 + o1);
              }
            }
            if (h == 1) {      // keep act_half's phase in step (nobody waits on this arrival)
              __syncwarp();
              if (lane == 0) mbar_arrive(act_half);
            }
          } else if (last && C == 64 && !mono) {
            // Last stage at C = 64 (output = the next stage's 16-bit image in natural row order):
            // each thread takes its lane's row of BOTH the E and the O block of this part for a
            // quarter of the channels, so rows 2m and 2m + 1 leave as 32 contiguous bytes per
            // channel chunk and a warp writes 1 KB runs.  (One block per thread meant 16-byte
            // pieces at a 32-byte stride: this epilogue took twice as long as the others.)
            if (!MSB_ABL(2)) {
              const int cq = cpart * COLS + ms * 16;             // this thread's 16 channels
              const int m = h * 128 + row0;                      // HE == 1
              const int t = t0 + 2 * m;                          // even row; t + 1 is the odd one
              float f[2][16];
#pragma unroll
              for (int r = 0; r < 2; ++r) {
                const uint32_t txr = tmem_base + lane_off + static_cast<uint32_t>((h * HB + r) * 2 * C + cq);
                uint32_t v[16], xr[16];
                tmem_ld16p(txr + C, v);
                tmem_ld16p(txr, xr);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                  float l0, l1;
                  leaky02x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), l0, l1);
                  add_x2(__uint_as_float(xr[j]), __uint_as_float(xr[j + 1]), l0, l1, f[r][j], f[r][j + 1]);
                }
                // seed the accumulator with the next tile's stage-0 bias
                if (!MSB_ABL(8)) {
                  const float4* b4 = reinterpret_cast<const float4*>(p.bias + cq);
                  uint32_t bv[16];
#pragma unroll
                  for (int j4 = 0; j4 < 4; ++j4) {
                    const float4 t4 = __ldg(b4 + j4);
                    bv[j4 * 4 + 0] = __float_as_uint(t4.x); bv[j4 * 4 + 1] = __float_as_uint(t4.y);
                    bv[j4 * 4 + 2] = __float_as_uint(t4.z); bv[j4 * 4 + 3] = __float_as_uint(t4.w);
                  }
                  tmem_st16p(txr + C, bv);
                }
              }
              // rows outside the clip are never stored (t, t + 1 are in or out together: L even)
              const int trow = 2 * m;
              if (t >= 0 && t < p.L && trow >= p.halo && trow < p.halo + p.V) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                  const size_t idx = (static_cast<size_t>(b) * G::NCH + (cq >> 3) + c) * p.L + t;
                  if (p.y16 != nullptr) {
                    uint4* dst = reinterpret_cast<uint4*>(p.y16 + idx * 8);
                    dst[0] = make_uint4(pack2u<BF>(f[0][c * 8 + 0], f[0][c * 8 + 1]),
                                        pack2u<BF>(f[0][c * 8 + 2], f[0][c * 8 + 3]),
                                        pack2u<BF>(f[0][c * 8 + 4], f[0][c * 8 + 5]),
                                        pack2u<BF>(f[0][c * 8 + 6], f[0][c * 8 + 7]));
                    dst[1] = make_uint4(pack2u<BF>(f[1][c * 8 + 0], f[1][c * 8 + 1]),
                                        pack2u<BF>(f[1][c * 8 + 2], f[1][c * 8 + 3]),
                                        pack2u<BF>(f[1][c * 8 + 4], f[1][c * 8 + 5]),
                                        pack2u<BF>(f[1][c * 8 + 6], f[1][c * 8 + 7]));
                  }
                  if (p.y32 != nullptr) {
#pragma unroll
                    for (int r = 0; r < 2; ++r) {
                      const float o8[8] = {f[r][c * 8 + 0], f[r][c * 8 + 1], f[r][c * 8 + 2],
                                           f[r][c * 8 + 3], f[r][c * 8 + 4], f[r][c * 8 + 5],
                                           f[r][c * 8 + 6], f[r][c * 8 + 7]};
                      st_global_v8(p.y32 + (idx + r) * 8, o8);
                    }
                  }
                }
              }
            }
          } else
#pragma unroll
          for (int u = 0; u < IT; ++u) {
            if (MSB_ABL(2)) continue;
            int mb, srow, trow;
            geom(h, u, mb, srow, trow);
            const int t = t0 + trow;
            const uint32_t tx = tm0 + static_cast<uint32_t>(mb * 2 * C);
            const uint32_t ta = tx + C;
            const bool zero_row = edge && (t < 0 || t >= p.L);
            constexpr int GS = last ? 16 : COLS;
#pragma unroll
            for (int c0 = 0; c0 < COLS; c0 += GS) {
              uint32_t v[GS];
              if (MSB_ABL(8)) {
#pragma unroll
                for (int j = 0; j < GS; ++j) v[j] = static_cast<uint32_t>(t + j);
              } else {
#pragma unroll
                for (int gg = 0; gg < GS / 16; ++gg) tmem_ld16p(ta + c0 + gg * 16, &v[gg * 16]);
              }
              float f[GS];
              if (resid) {
                uint32_t xr[GS];
                if (MSB_ABL(8)) {
#pragma unroll
                  for (int j = 0; j < GS; ++j) xr[j] = static_cast<uint32_t>(t - j);
                } else {
#pragma unroll
                  for (int gg = 0; gg < GS / 16; ++gg) tmem_ld16p(tx + c0 + gg * 16, &xr[gg * 16]);
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
                if (KIND != 1 && !MSB_ABL(8)) {
#pragma unroll
                  for (int j = 0; j < GS; ++j) v[j] = __float_as_uint(f[j]);
#pragma unroll
                  for (int gg = 0; gg < GS / 16; ++gg) tmem_st16p(tx + c0 + gg * 16, &v[gg * 16]);
                }
#pragma unroll
                for (int c = 0; c < GS / 8; ++c) {
                  if (MSB_ABL(4)) continue;
                  const uint32_t dst =
                      dstbuf + static_cast<uint32_t>(((chunk0 + c0 / 8 + c) * R + srow) * 16);
                  st_shared_v4(dst, pack2u<BF>(f[c * 8 + 0], f[c * 8 + 1]),
                               pack2u<BF>(f[c * 8 + 2], f[c * 8 + 3]),
                               pack2u<BF>(f[c * 8 + 4], f[c * 8 + 5]),
                               pack2u<BF>(f[c * 8 + 6], f[c * 8 + 7]));
                }
              } else if (mono) {
                // fused tail, step 1: 7 per-tap partial dot products of this thread's channels
                float pk[7];
#pragma unroll
                for (int k = 0; k < 7; ++k) {
                  float a0 = 0.f, a1 = 0.f;
#pragma unroll
                  for (int j4 = 0; j4 < GS / 4; ++j4) {
                    const float4 w4 = *reinterpret_cast<const float4*>(
                        sMono + k * 32 + cpart * COLS + c0 + j4 * 4);
                    fma_x2(f[j4 * 4 + 0], f[j4 * 4 + 1], w4.x, w4.y, a0, a1);
                    fma_x2(f[j4 * 4 + 2], f[j4 * 4 + 3], w4.z, w4.w, a0, a1);
                  }
                  pk[k] = a0 + a1;
                }
                float* P = sP + (cpart * 7) * R + trow;
#pragma unroll
                for (int k = 0; k < 7; ++k) P[k * R] = pk[k];
              } else if (t >= 0 && t < p.L && trow >= p.halo && trow < p.halo + p.V) {
#pragma unroll
                for (int c = 0; c < GS / 8; ++c) {
                  const size_t idx =
                      (static_cast<size_t>(b) * G::NCH + chunk0 + c0 / 8 + c) * p.L + t;
                  if (p.y16 != nullptr)
                    *reinterpret_cast<uint4*>(p.y16 + idx * 8) =
                        make_uint4(pack2u<BF>(f[c * 8 + 0], f[c * 8 + 1]),
                                   pack2u<BF>(f[c * 8 + 2], f[c * 8 + 3]),
                                   pack2u<BF>(f[c * 8 + 4], f[c * 8 + 5]),
                                   pack2u<BF>(f[c * 8 + 6], f[c * 8 + 7]));
                  if (p.y32 != nullptr) {
                    const float o8[8] = {f[c * 8 + 0], f[c * 8 + 1], f[c * 8 + 2], f[c * 8 + 3],
                                         f[c * 8 + 4], f[c * 8 + 5], f[c * 8 + 6], f[c * 8 + 7]};
                    st_global_v8(p.y32 + idx * 8, o8);
                  }
                }
              }
            }
            // the accumulator has been read: seed it with the next stage's bias
            if (!MSB_ABL(8)) store_bias(last ? 0 : s + 1, ta);
            if (HE == 2 && h == 1 && u == 0) {
              // (the last stage writes no operand rows, but the barrier's phase must advance)
              tmem_st_wait();
              fence_proxy_async_smem();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(act_half);
            }
          }
          tmem_st_wait();
          fence_proxy_async_smem();
          tc_fence_before();
          arrive_act(h);
          if (warp == 2) MSB_UTRACE(128 + s * 16 + h * 4 + 2);
        }
        ++g;
      };
      using K0 = std::integral_constant<int, 0>;
      using K1 = std::integral_constant<int, 1>;
      using K2 = std::integral_constant<int, 2>;
      using K3 = std::integral_constant<int, 3>;
      stage_epi(K0{}, 0);
      stage_epi(K1{}, 1);
      stage_epi(K2{}, 2);
      stage_epi(K1{}, 3);
      stage_epi(K2{}, 4);
      stage_epi(K1{}, 5);
      stage_epi(K3{}, 6);
      if (mono) {
        // fused tail, step 2: y[t] = tanh(b + sum_k sum_part P[part][k][row + k - 3])
        named_bar_sync(1, 32 * EW);
        const float bias0 = __ldg(p.mono_b);
        const float* P = sP;
        for (int row = p.halo + static_cast<int>(threadIdx.x) - 64; row < p.halo + p.V;
             row += 32 * EW) {
          const int t = t0 + row;
          if (t < 0 || t >= p.L) continue;
          float a0 = bias0, a1 = 0.f;
#pragma unroll
          for (int k = 0; k < 7; ++k) {
            a0 += P[k * R + row + k - 3];
            a1 += P[(7 + k) * R + row + k - 3];
          }
          p.mono_out[static_cast<size_t>(b) * p.L + t] = tanhf(a0 + a1);
        }
        // the next writer of P is the next tile's stage 6, six full hand-offs away
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int C>
__global__ void __launch_bounds__(UpGeom<C>::THREADS, 1)
upstack_kernel(const __grid_constant__ UpStackParams p) {
  upstack_body<C, false>(p);
}

template <int C>
__global__ void __launch_bounds__(UpGeom<C>::THREADS, 1)
upstack_bf16_kernel(const __grid_constant__ UpStackParams p) {
  upstack_body<C, true>(p);
}

template <int C>
ms_status launch_upstack(const UpStackParams& p, int operand, cudaStream_t stream) {
  using G = UpGeom<C>;
  static thread_local bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(upstack_kernel<C>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(upstack_bf16_kernel<C>,
                               cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM);
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(upstack_kernel)");
    attr_set = true;
  }
  const int sms = sm_count();
  if (sms <= 0) return check_cuda(cudaGetLastError(), "sm_count");
  const int grid = p.total_tiles < sms ? p.total_tiles : sms;
  if (operand == MS_BF16)
    upstack_bf16_kernel<C><<<grid, G::THREADS, G::SMEM, stream>>>(p);
  else
    upstack_kernel<C><<<grid, G::THREADS, G::SMEM, stream>>>(p);
  return after_launch("upstack_kernel");
}

// out[e][chunk][n][j] for the 8 transposed-conv images e = (phase r, tap, K half):
//   B[n = co][k = ci'] = Wt[kh*C + ci'][co][ktap],  ktap = (r ? 0 : 1) + 2*tap
// (Wt: ConvTranspose1d weight (2C, C, 4), reference layout)
__global__ void pack_up_weight_kernel(const float* __restrict__ wt, uint16_t* __restrict__ out,
                                      int C, int operand) {
  const int total = 8 * C * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = i % 8;
  const int n = (i / 8) % C;
  const int c = (i / (8 * C)) % (C / 8);
  const int e = i / (C * C);
  const int r = e >> 2, tap = (e >> 1) & 1, kh = e & 1;
  const int ktap = (r ? 0 : 1) + 2 * tap;
  const int ci = kh * C + c * 8 + j;
  const float v = wt[(static_cast<size_t>(ci) * C + n) * 4 + ktap];
  if (operand == MS_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  } else {
    __half h = __float2half_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

// [tap][chunk][C][8] <- w (C, C, 3) fp32 (same image as the fused stack's)
__global__ void pack_up_tap_kernel(const float* __restrict__ w, uint16_t* __restrict__ out, int C,
                                   int operand) {
  const int total = 3 * C * C;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int j = i % 8;
  const int n = (i / 8) % C;
  const int c = (i / (8 * C)) % (C / 8);
  const int t = i / (C * C);
  const float v = w[(static_cast<size_t>(n) * C + c * 8 + j) * 3 + t];
  if (operand == MS_BF16) {
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  } else {
    __half h = __float2half_rn(v);
    out[i] = *reinterpret_cast<uint16_t*>(&h);
  }
}

}  // namespace

static thread_local long long* g_up_dbg = nullptr;

ms_status upstack_fwd(int channels, int batch, int lin, const int* dil, int operand,
                      const void* x16, const void* packed, void* y16, float* y32,
                      const float* mono_w, const float* mono_b, float* mono_out,
                      cudaStream_t stream) {
  if (batch <= 0 || lin <= 0 || x16 == nullptr || packed == nullptr ||
      (y16 == nullptr && y32 == nullptr && mono_out == nullptr))
    return MS_ERR_INVALID;
  if (channels != 64 && channels != 32) return MS_ERR_INVALID;
  if (mono_out != nullptr && (channels != 32 || mono_w == nullptr || mono_b == nullptr))
    return MS_ERR_INVALID;
  for (int i = 0; i < 3; ++i)
    if (dil[i] < 1 || dil[i] > 9 || (dil[i] & 1) == 0) return MS_ERR_INVALID;   // odd only
  if (dil[0] + dil[1] + dil[2] + 3 > kUpHalo) return MS_ERR_INVALID;
  UpStackParams p;
  p.x16 = static_cast<const uint16_t*>(x16);
  p.w = static_cast<const uint16_t*>(packed);
  p.bias = reinterpret_cast<const float*>(static_cast<const uint8_t*>(packed) +
                                          static_cast<size_t>(kUpEntries) * channels * channels * 2);
  p.y16 = static_cast<uint16_t*>(y16);
  p.y32 = y32;
  p.B = batch; p.lin = lin; p.L = 2 * lin;
  p.dil[0] = dil[0]; p.dil[1] = dil[1]; p.dil[2] = dil[2];
  p.mono_w = mono_w; p.mono_b = mono_b; p.mono_out = mono_out;
  p.ablate = 0;
  p.dbg = nullptr;
#ifdef MSB_UP_ABLATE
  if (const char* e = getenv("MSB_UP_ABLATE")) p.ablate = atoi(e);
  p.dbg = g_up_dbg;
#endif
  // rows lost per tile: the halo on both sides, the 3 rows the low-rate window cannot feed
  // (rounded to keep V even) and, with the fused k7 tail, its reach of 3 on both sides
  const int halo = mono_out != nullptr ? kUpHalo + 4 : kUpHalo;
  const int R = channels == 64 ? UpGeom<64>::R : UpGeom<32>::R;
  p.halo = halo;
  p.V = R - 2 * halo - (mono_out != nullptr ? 2 : 4);
  p.tiles_per_clip = (p.L + p.V - 1) / p.V;
  const long long tiles = static_cast<long long>(batch) * p.tiles_per_clip;
  if (tiles > 0x7fffffffLL) return MS_ERR_INVALID;
  p.total_tiles = static_cast<int>(tiles);
  return channels == 64 ? launch_upstack<64>(p, operand, stream)
                        : launch_upstack<32>(p, operand, stream);
}

}  // namespace msb

using namespace msb;

extern "C" {

/* debugging aid (not in the public header; effective in MSB_UP_ABLATE builds only) */
void ms_debug_set_upstack_trace(long long* dev_buf) { msb::g_up_dbg = dev_buf; }

int ms_upstack_supported(int channels) { return channels == 64 || channels == 32; }

size_t ms_upstack_packed_weight_bytes(int channels) {
  if (!ms_upstack_supported(channels)) return 0;
  return static_cast<size_t>(kUpEntries) * channels * channels * 2 +
         sizeof(float) * kUpStages * channels;
}

ms_status ms_upstack_pack_weights(const float* const* params, int channels, int operand,
                                  void* packed, void* stream) {
  if (params == nullptr || packed == nullptr || !ms_upstack_supported(channels))
    return MS_ERR_INVALID;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(packed);
  const size_t image = static_cast<size_t>(channels) * channels * 2;
  float* bias = reinterpret_cast<float*>(base + kUpEntries * image);
  pack_up_weight_kernel<<<(8 * channels * channels + 255) / 256, 256, 0, st>>>(
      params[0], reinterpret_cast<uint16_t*>(base), channels, operand);
  ms_status s = after_launch("pack_up_weight_kernel");
  if (s != MS_OK) return s;
  s = check_cuda(cudaMemcpyAsync(bias, params[1], sizeof(float) * channels,
                                 cudaMemcpyDeviceToDevice, st),
                 "cudaMemcpyAsync(upstack bias)");
  if (s != MS_OK) return s;
  for (int l = 0; l < 6; ++l) {
    pack_up_tap_kernel<<<(3 * channels * channels + 255) / 256, 256, 0, st>>>(
        params[2 + 2 * l], reinterpret_cast<uint16_t*>(base + (8 + 3 * l) * image), channels,
        operand);
    s = after_launch("pack_up_tap_kernel");
    if (s != MS_OK) return s;
    s = check_cuda(cudaMemcpyAsync(bias + (l + 1) * channels, params[3 + 2 * l],
                                   sizeof(float) * channels, cudaMemcpyDeviceToDevice, st),
                   "cudaMemcpyAsync(upstack bias)");
    if (s != MS_OK) return s;
  }
  return MS_OK;
}

ms_status ms_upstack_fwd(int channels, int batch, int lin, const int* dilations, int operand,
                         const void* x16, const void* packed, void* y16, float* y32,
                         const float* tail_w, const float* tail_b, float* tail_y, void* stream) {
  if (dilations == nullptr || !ms_upstack_supported(channels)) return MS_ERR_INVALID;
  return upstack_fwd(channels, batch, lin, dilations, operand, x16, packed, y16, y32, tail_w,
                     tail_b, tail_y, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
