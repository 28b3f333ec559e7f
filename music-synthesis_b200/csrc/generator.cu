// MelGanGenerator forward (inference): layer schedule over the tcgen05 conv kernels.
//   replaces featuresynth/generator/full.py:16-50 (+ util/modules.py:350-405).
//
// Numerics: 16-bit operands (fp16 default), fp32 accumulation in tensor memory, and an
// fp32 residual stream: each ResidualAtom output is kept in fp32 (BLK f32) for the next
// skip connection and, rounded once, in 16-bit as the next conv's operand.  The final
// 32->1 conv + tanh runs in fp32 on the fp32 stream.
#include <cstdlib>
#include <vector>

#include "conv_gemm.cuh"
#include "runtime.cuh"

namespace msb {

ms_status launch_conv(const ms_conv_desc& d, const ConvCfg& c, const void* x16,
                      const void* w_packed, const float* bias, const float* res32, void* y16,
                      float* y32, cudaStream_t stream);
ms_status conv_to_mono(const float* x32, const float* w, const float* bias, float* y, int batch,
                       int cin, int len, int ksize, int pad, int tanh_out, cudaStream_t stream);
ms_status pack_ncl_to_blk16(const float* x, void* y16, int batch, int channels, int len,
                            int pad, int pad_mode, int operand, cudaStream_t stream);
ms_status resstack_fwd(int channels, int batch, int len, const int* dil, int operand,
                       const float* x32, const void* packed, void* y16, float* y32,
                       cudaStream_t stream, const float* mono_w, const float* mono_b,
                       float* mono_out, const void* x16in);
ms_status upstack_fwd(int channels, int batch, int lin, const int* dil, int operand,
                      const void* x16, const void* packed, void* y16, float* y32,
                      const float* mono_w, const float* mono_b, float* mono_out,
                      cudaStream_t stream);

namespace {

// MSB_FUSE_UP=0: run the stride-2 upsamplers as separate launches in front of the fused stacks
// (the round-1 schedule; kept for A/B timing).  Default: one launch per stage (upstack.cu).
bool fuse_upsamplers() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_FUSE_UP");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// MSB_STAGE1_CHUNK: clips per sub-chunk of the leading low-rate layers (0 = whole pass)
int stage1_chunk() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_STAGE1_CHUNK");
    v = e != nullptr ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return v;
}

// MSB_STAGE_ENTRY16: bit mask of fused stages whose input stream starts from the 16-bit operand
// instead of an fp32 tensor (1: C=32, 2: C=64, 4: C=128): halves the HBM bytes between the
// upsampler and the stack at the price of one fp16 rounding of the stream per stage
// (measured waveform rel-L2 7.2e-4 -> 7.9e-4 for mask 3, 8.3e-4 for mask 7)
int stage_entry16() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("MSB_STAGE_ENTRY16");
    v = e != nullptr ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return v;
}

bool entry16_for(int channels) {
  const int m = stage_entry16();
  return (channels == 32 && (m & 1)) || (channels == 64 && (m & 2)) || (channels == 128 && (m & 4));
}

struct GenLayer {
  ms_conv_desc d;   // batch / lin filled per call
  int w_param, b_param;
  size_t w_off, b_off;
  int role;         // 0 first conv, 1 upsampler, 2 atom conv1, 3 atom conv2,
                    // 4 fused ResidualStack (w_param = first of its 12 params), 5 = upsampler
                    // feeding a fused stack (fp32 output only), 6 = upsampler + ResidualStack
                    // (+ tail) in one kernel (w_param = first of its 14 params)
  int len_mult;     // lin = len_mult * T (+6 for the first conv)
};

struct GenPlan {
  std::vector<GenLayer> layers;
  size_t final_w_off, final_b_off;
  size_t total_bytes;
};

// (cin, cout, ksize, stride, pad) of the four upsamplers, generator/full.py:27-39
const int kUp[4][5] = {{512, 256, 16, 8, 4}, {256, 128, 16, 8, 4}, {128, 64, 4, 2, 1},
                       {64, 32, 4, 2, 1}};
const int kDil[3] = {1, 3, 9};

bool build_plan(int in_channels, int operand, GenPlan* plan) {
  plan->layers.clear();
  size_t off = 0;
  int param = 0;
  auto add = [&](ms_conv_desc d, int role, int len_mult) -> bool {
    d.batch = 1;
    d.lin = 1024;  // nominal: the tile config does not depend on batch / length
    d.operand = operand;
    d.alpha = 1.0f;
    ConvCfg c;
    if (!make_conv_cfg(d, &c)) return false;
    GenLayer L;
    L.d = d; L.role = role; L.len_mult = len_mult;
    L.w_param = param++; L.b_param = param++;
    L.w_off = off; off = align_up(off + c.packed_weight_bytes, 256);
    L.b_off = off; off = align_up(off + sizeof(float) * d.cout, 256);
    plan->layers.push_back(L);
    return true;
  };
  ms_conv_desc d{};
  d.kind = MS_CONV; d.cin = in_channels; d.cout = 512; d.ksize = 7; d.dilation = 1;
  d.pad = 0; d.stride = 1; d.leaky = 1;
  if (!add(d, 0, 1)) return false;
  int mult = 1;
  for (int s = 0; s < 4; ++s) {
    ms_conv_desc u{};
    u.kind = MS_CONVT; u.cin = kUp[s][0]; u.cout = kUp[s][1]; u.ksize = kUp[s][2];
    u.stride = kUp[s][3]; u.pad = kUp[s][4]; u.dilation = 1; u.leaky = 1;
    if (fuse_upsamplers() && u.ksize == 4 && u.stride == 2 && ms_upstack_supported(u.cout)) {
      GenLayer L{};
      L.d = u; L.d.operand = operand;
      L.role = 6; L.len_mult = mult;    // lin = mult * T
      L.w_param = param; param += 14; L.b_param = -1;
      L.w_off = off; off = align_up(off + ms_upstack_packed_weight_bytes(u.cout), 256);
      L.b_off = 0;
      plan->layers.push_back(L);
      mult *= 2;
      continue;
    }
    if (!add(u, 1, mult)) return false;
    mult *= kUp[s][3];
    if (ms_resstack_supported(kUp[s][1])) {
      plan->layers.back().role = 5;
      GenLayer L{};
      L.d = u; L.d.cin = L.d.cout = kUp[s][1]; L.d.operand = operand;
      L.role = 4; L.len_mult = mult;
      L.w_param = param; param += 12; L.b_param = -1;
      L.w_off = off; off = align_up(off + ms_resstack_packed_weight_bytes(kUp[s][1]), 256);
      L.b_off = 0;
      plan->layers.push_back(L);
      continue;
    }
    for (int a = 0; a < 3; ++a) {
      ms_conv_desc c1{};
      c1.kind = MS_CONV; c1.cin = c1.cout = kUp[s][1]; c1.ksize = 3; c1.dilation = kDil[a];
      c1.pad = kDil[a]; c1.stride = 1; c1.leaky = 1;
      if (!add(c1, 2, mult)) return false;
      ms_conv_desc c2 = c1;
      c2.dilation = 1; c2.pad = 1;
      if (!add(c2, 3, mult)) return false;
    }
  }
  plan->final_w_off = off; off = align_up(off + sizeof(float) * 32 * 7, 256);
  plan->final_b_off = off; off = align_up(off + sizeof(float), 256);
  plan->total_bytes = off;
  return param == MS_MELGAN_NUM_PARAMS - 2;
}

size_t per_clip_workspace(int frames, int in_channels) {
  const size_t T = frames;
  const size_t E = 8192 * T;  // largest stage: C*L = 128*64T = 64*128T = 32*256T
  size_t b = 0;
  b += align_up(static_cast<size_t>(in_channels) * (T + 6) * 2, 256);
  b += align_up(512 * T * 2, 256);
  b += 3 * align_up(E * 2, 256);
  b += 2 * align_up(E * 4, 256);
  return b;
}

}  // namespace
}  // namespace msb

using namespace msb;

extern "C" {

size_t ms_melgan_packed_weight_bytes(int in_channels, int operand) {
  GenPlan plan;
  if (!build_plan(in_channels, operand, &plan)) return 0;
  return plan.total_bytes;
}

ms_status ms_melgan_pack_weights(const float* const* params, int in_channels, int operand,
                                 void* packed, void* stream) {
  GenPlan plan;
  if (params == nullptr || packed == nullptr || !build_plan(in_channels, operand, &plan))
    return MS_ERR_INVALID;
  uint8_t* base = static_cast<uint8_t*>(packed);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (const GenLayer& L : plan.layers) {
    if (L.role == 4) {
      ms_status s4 = ms_resstack_pack_weights(params + L.w_param, L.d.cout, operand,
                                              base + L.w_off, stream);
      if (s4 != MS_OK) return s4;
      continue;
    }
    if (L.role == 6) {
      ms_status s6 = ms_upstack_pack_weights(params + L.w_param, L.d.cout, operand,
                                             base + L.w_off, stream);
      if (s6 != MS_OK) return s6;
      continue;
    }
    ms_status s = ms_conv_pack_weight(&L.d, params[L.w_param], base + L.w_off, stream);
    if (s != MS_OK) return s;
    s = check_cuda(cudaMemcpyAsync(base + L.b_off, params[L.b_param],
                                   sizeof(float) * L.d.cout, cudaMemcpyDeviceToDevice, st),
                   "cudaMemcpyAsync(bias)");
    if (s != MS_OK) return s;
  }
  ms_status s = check_cuda(
      cudaMemcpyAsync(base + plan.final_w_off, params[MS_MELGAN_NUM_PARAMS - 2],
                      sizeof(float) * 32 * 7, cudaMemcpyDeviceToDevice, st),
      "cudaMemcpyAsync(final w)");
  if (s != MS_OK) return s;
  return check_cuda(cudaMemcpyAsync(base + plan.final_b_off, params[MS_MELGAN_NUM_PARAMS - 1],
                                    sizeof(float), cudaMemcpyDeviceToDevice, st),
                    "cudaMemcpyAsync(final b)");
}

size_t ms_melgan_workspace_bytes(int batch, int frames, int in_channels) {
  if (batch <= 0 || frames <= 0 || in_channels <= 0) return 0;
  return static_cast<size_t>(batch) * per_clip_workspace(frames, in_channels);
}

ms_status ms_melgan_generator_fwd(const void* packed_weights, int in_channels, int operand,
                                  const float* x, float* y, int batch, int frames,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  GenPlan plan;
  if (packed_weights == nullptr || x == nullptr || y == nullptr || workspace == nullptr ||
      batch <= 0 || frames < 4 || !build_plan(in_channels, operand, &plan))
    return MS_ERR_INVALID;
  const size_t per_clip = per_clip_workspace(frames, in_channels);
  int bc = static_cast<int>(workspace_bytes / per_clip);
  if (bc < 1) return MS_ERR_WORKSPACE;
  if (bc > batch) bc = batch;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const uint8_t* wb = static_cast<const uint8_t*>(packed_weights);
  const size_t T = frames;
  const size_t E = 8192 * T;

  // carve the workspace for `bc` clips
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  size_t o = 0;
  auto carve = [&](size_t bytes_per_clip) {
    uint8_t* p = ws + o;
    o += static_cast<size_t>(bc) * align_up(bytes_per_clip, 256);
    return p;
  };
  void* xin16 = carve(static_cast<size_t>(in_channels) * (T + 6) * 2);
  void* h16 = carve(512 * T * 2);
  void* x16[2] = {carve(E * 2), carve(E * 2)};
  void* y16 = carve(E * 2);
  float* x32[2] = {reinterpret_cast<float*>(carve(E * 4)), reinterpret_cast<float*>(carve(E * 4))};

  for (int b0 = 0; b0 < batch; b0 += bc) {
    const int nb = (batch - b0) < bc ? (batch - b0) : bc;
    ms_status s = pack_ncl_to_blk16(x + static_cast<size_t>(b0) * in_channels * T, xin16, nb,
                                    in_channels, frames, 3, 1, operand, st);
    if (s != MS_OK) return s;
    bool fused_tail = false;
    const void* cur16 = nullptr;  // 16-bit operand of the next layer
    int cur = 0;                  // which x16/x32 buffer holds the residual stream
    // The leading low-rate layers (first conv, first upsampler, C=256 residual stack) are
    // run in sub-chunks of `sc` clips so that their fp32 residual stream stays L2-resident
    // between consecutive launches; everything after runs on the whole pass.
    size_t nlead = 0;
    while (nlead < plan.layers.size() && plan.layers[nlead].role <= 3) ++nlead;
    int sc = stage1_chunk();
    if (sc <= 0 || sc > nb) sc = nb;
    for (int c0 = 0; c0 < nb; c0 += sc) {
      const int ncb = (nb - c0) < sc ? (nb - c0) : sc;
      const size_t o1 = static_cast<size_t>(c0) * 2048 * T;          // stage-1 elements / clip
      uint8_t* lx16[2] = {static_cast<uint8_t*>(x16[0]) + o1 * 2, static_cast<uint8_t*>(x16[1]) + o1 * 2};
      uint8_t* ly16 = static_cast<uint8_t*>(y16) + o1 * 2;
      float* lx32[2] = {x32[0] + o1, x32[1] + o1};
      const uint8_t* lxin = static_cast<const uint8_t*>(xin16) + static_cast<size_t>(c0) * in_channels * (T + 6) * 2;
      uint8_t* lh16 = static_cast<uint8_t*>(h16) + static_cast<size_t>(c0) * 512 * T * 2;
      const void* l16 = nullptr;
      int lc = 0;
      for (size_t li = 0; li < nlead; ++li) {
        const GenLayer& L = plan.layers[li];
        ms_conv_desc d = L.d;
        d.batch = ncb;
        d.lin = L.len_mult * frames + (L.role == 0 ? 6 : 0);
        ConvCfg c;
        if (!make_conv_cfg(d, &c)) return MS_ERR_INVALID;
        const void* w = wb + L.w_off;
        const float* bias = reinterpret_cast<const float*>(wb + L.b_off);
        switch (L.role) {
          case 0:
            s = launch_conv(d, c, lxin, w, bias, nullptr, lh16, nullptr, st);
            l16 = lh16;
            break;
          case 1:
            lc = 0;
            s = launch_conv(d, c, l16, w, bias, nullptr, lx16[0], lx32[0], st);
            l16 = lx16[0];
            break;
          case 2:
            s = launch_conv(d, c, l16, w, bias, nullptr, ly16, nullptr, st);
            break;
          default: {
            // the fp32 stream is only written when a later layer reads it (the next atom's skip
            // connection, or the fp32 final conv): an upsampler next reads the 16-bit image alone
            const bool feeds_up = li + 1 < plan.layers.size() &&
                                  (plan.layers[li + 1].role == 1 || plan.layers[li + 1].role >= 5);
            s = launch_conv(d, c, ly16, w, bias, lx32[lc], lx16[lc ^ 1],
                            feeds_up ? nullptr : lx32[lc ^ 1], st);
            lc ^= 1;
            l16 = lx16[lc];
            break;
          }
        }
        if (s != MS_OK) return s;
      }
      cur = lc;
    }
    cur16 = x16[cur];
    for (size_t li = nlead; li < plan.layers.size(); ++li) {
      const GenLayer& L = plan.layers[li];
      if (L.role == 6) {
        // upsampler + ResidualStack (+ tail) in one kernel: 16-bit operand in, 16-bit out
        const bool tail = L.d.cout == 32;
        void* out16 = (cur16 == x16[1]) ? x16[0] : x16[1];
        s = upstack_fwd(L.d.cout, nb, L.len_mult * frames, kDil, operand, cur16, wb + L.w_off,
                        tail ? nullptr : out16, nullptr,
                        tail ? reinterpret_cast<const float*>(wb + plan.final_w_off) : nullptr,
                        tail ? reinterpret_cast<const float*>(wb + plan.final_b_off) : nullptr,
                        tail ? y + static_cast<size_t>(b0) * 256 * T : nullptr, st);
        if (s != MS_OK) return s;
        if (tail) fused_tail = true;
        cur16 = out16;
        continue;
      }
      if (L.role == 4) {
        // fused ResidualStack: fp32 stream in, 16-bit operand (+ fp32 for the last stage) out
        if (L.d.cout == 32) {
          // last stage: the 32->1 k7 conv + tanh is fused into the stack kernel's tail
          s = resstack_fwd(32, nb, L.len_mult * frames, kDil, operand, x32[0], wb + L.w_off,
                           nullptr, nullptr, st,
                           reinterpret_cast<const float*>(wb + plan.final_w_off),
                           reinterpret_cast<const float*>(wb + plan.final_b_off),
                           y + static_cast<size_t>(b0) * 256 * T,
                           entry16_for(32) ? x16[0] : nullptr);
          if (s != MS_OK) return s;
          fused_tail = true;
          continue;
        }
        s = resstack_fwd(L.d.cout, nb, L.len_mult * frames, kDil, operand, x32[0],
                         wb + L.w_off, x16[1], nullptr, st, nullptr, nullptr, nullptr,
                         entry16_for(L.d.cout) ? x16[0] : nullptr);
        if (s != MS_OK) return s;
        cur = 1;
        cur16 = x16[1];
        continue;
      }
      ms_conv_desc d = L.d;
      d.batch = nb;
      d.lin = L.len_mult * frames + (L.role == 0 ? 6 : 0);
      ConvCfg c;
      if (!make_conv_cfg(d, &c)) return MS_ERR_INVALID;
      const void* w = wb + L.w_off;
      const float* bias = reinterpret_cast<const float*>(wb + L.b_off);
      switch (L.role) {
        case 0:
          s = launch_conv(d, c, xin16, w, bias, nullptr, h16, nullptr, st);
          cur16 = h16;
          break;
        case 1:  // upsampler: starts a new residual stream
          cur = 0;
          s = launch_conv(d, c, cur16, w, bias, nullptr, x16[0], x32[0], st);
          cur16 = x16[0];
          break;
        case 5:  // upsampler in front of a fused stack: only the stream is needed (fp32, or
                 // the 16-bit operand when the stage is selected by MSB_STAGE_ENTRY16)
          cur = 0;
          if (entry16_for(d.cout))
            s = launch_conv(d, c, cur16, w, bias, nullptr, x16[0], nullptr, st);
          else
            s = launch_conv(d, c, cur16, w, bias, nullptr, nullptr, x32[0], st);
          cur16 = nullptr;
          break;
        case 2:  // y = leaky(conv_dil(x))
          s = launch_conv(d, c, cur16, w, bias, nullptr, y16, nullptr, st);
          break;
        default: {  // x' = x + leaky(conv(y))
          const bool feeds_up = li + 1 < plan.layers.size() &&
                                (plan.layers[li + 1].role == 1 || plan.layers[li + 1].role >= 5);
          s = launch_conv(d, c, y16, w, bias, x32[cur], x16[cur ^ 1],
                          feeds_up ? nullptr : x32[cur ^ 1], st);
          cur ^= 1;
          cur16 = x16[cur];
          break;
        }
      }
      if (s != MS_OK) return s;
    }
    if (fused_tail) continue;
    s = conv_to_mono(x32[cur], reinterpret_cast<const float*>(wb + plan.final_w_off),
                     reinterpret_cast<const float*>(wb + plan.final_b_off),
                     y + static_cast<size_t>(b0) * 256 * T, nb, 32, 256 * frames, 7, 3, 1, st);
    if (s != MS_OK) return s;
  }
  return MS_OK;
}

}  // extern "C"
