// Host-visible description of one tcgen05 implicit-GEMM convolution launch.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/msb200.h"

namespace msb {

constexpr int kMaxTaps = 32;   // k41 stride-2 convs: 21 taps over space-to-depth input; 511-tap banks: 32 taps of dilation 16
constexpr int kMaxStages = 8;
constexpr int kSmemHeader = 3072;          // barriers [0,512) | bias tables [512,2560) | index tables [2560,3072)
constexpr int kSmemBudget = 227 * 1024;    // max dynamic smem per CTA on sm_100
constexpr int kPairEpiWarps = 8;
constexpr int kPairStageRow = 144;         // staged fp32 row: 4 phases x 32 B + 16 B (bank spread)
constexpr int kPairStageWarp = 32 * kPairStageRow;   // per epilogue warp

// Tile configuration + derived geometry, a pure function of the descriptor so that
// weight packing and the forward launch always agree.
struct ConvCfg {
  int taps;            // GEMM taps (MS_CONV: ksize; MS_CONVT: ksize / stride)
  int off[kMaxTaps];   // input row of tap t for GEMM row m is m + off[t]
  int pair;            // 1: CTA-pair kernel (cta_group::2, 256-row cluster tile, W split)
  int out_stage;       // bytes of epilogue staging (pair kernel, ConvTranspose with stride % 4 == 0:
                       // outputs go through shared memory so that every store covers whole lines)
  int wres;            // 1 (pair kernel): the n-tile's weight slice stays resident in shared memory
                       // and only activations stream through the stage ring
  int MBLK;            // 128-row M-blocks per CTA tile (1 or 2): they share each W stage
  int min_off, RA;     // RA = 128*MBLK + max_off - min_off rows of A staged per tile
  int Ntot;            // GEMM N (MS_CONV: cout; MS_CONVT: stride*cout)
  int NT, KB;          // n-tile (columns per CTA tile), k-block (input channels per stage)
  int nnt, nkb;        // Ntot/NT, cin/KB
  int Lm;              // GEMM rows per clip (MS_CONV: Lout; MS_CONVT: lin+1)
  int Lout;
  int mtiles;          // ceil(Lm/(128*MBLK))
  int acc_stages;      // TMEM accumulator sets (2 = epilogue overlaps the next tile)
  int fold_slots;      // >0: short sequences, this many clips share one 128-row tile
  int fold_stride;     // rows per clip slot (Lm + tap span)
  int a_stage_bytes, w_stage_bytes, stage_bytes;
  int stages;
  int tmem_cols;       // power of two >= 2*NT
  size_t smem_bytes;
  size_t packed_weight_bytes;
};

// GEMM column order of the polyphase ConvTranspose: n = ((co / 8) * stride + phase) * 8 + co % 8.
// A thread of the epilogue (one input-rate row q) then meets the `stride` phases of an
// 8-channel block in consecutive 8-column chunks = consecutive output rows s*q + r - pad = one
// contiguous stride*32-byte (fp32) run, so every 128-byte line is written whole by one thread
// within one tile.  (With the phase-major order n = r * cout + co the four rows of a line came
// from different n-tiles, i.e. different CTAs at different times: partial-line writes.)
__host__ __device__ inline int convt_col(int phase, int co, int stride) {
  return ((co >> 3) * stride + phase) * 8 + (co & 7);
}
__host__ __device__ inline int convt_col_phase(int n, int stride) { return (n >> 3) % stride; }
__host__ __device__ inline int convt_col_channel(int n, int stride) {
  return ((n >> 3) / stride) * 8 + (n & 7);
}

// returns false when the descriptor is unsupported
bool make_conv_cfg(const ms_conv_desc& d, ConvCfg* cfg);

struct ConvGemmParams {
  const uint16_t* x;    // BLK 16-bit (B, cin/8, lin, 8)
  const uint16_t* w;    // packed
  const float* bias;    // [cout] or null
  const float* res32;   // BLK f32 (B, cout/8, Lout, 8) or null
  uint16_t* y16;        // or null
  float* y32;           // or null
  int B, cin, lin, cout, Lout, Lm;
  int xcin, xnkb;       // channels of the x tensor (cin / x_repeat) and its k-blocks (xcin / KB)
  int taps;
  int off[kMaxTaps];
  int min_off, RA;
  int Ntot, NT, KB, nnt, nkb, mtiles, MBLK, acc_stages, pair, wres, out_stage;
  int fold_slots, fold_stride, btiles;
  int stages, a_stage_bytes, w_stage_bytes, stage_bytes, tmem_cols;
  int kind, stride, pad, leaky, operand;
  float alpha;
  int total_tiles;
  long long* dbg;       // MSB_CONV_ABLATE builds only: clock64 trace buffer (tools/pair_trace.py)
  int debug;            // MSB_CONV_ABLATE builds only (timing experiments): bit 0 no global stores,
                        // 1 no MMAs, 2 no residual loads
};

}  // namespace msb
