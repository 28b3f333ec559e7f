// CPU check of the ConvTranspose GEMM column order (csrc/conv_gemm.cuh: convt_col and its two
// inverses): a bijection between columns n < stride * cout and (phase, channel), with the
// `stride` phases of an 8-channel block in consecutive 8-column chunks.
#include <cstdio>
#include <vector>

#include "conv_gemm.cuh"

int main() {
  using namespace msb;
  for (int stride : {1, 2, 4, 8, 16})
    for (int cout = 8; cout <= 512; cout += 8) {
      std::vector<char> seen(static_cast<size_t>(stride) * cout, 0);
      for (int n = 0; n < stride * cout; ++n) {
        const int r = convt_col_phase(n, stride), co = convt_col_channel(n, stride);
        if (r < 0 || r >= stride || co < 0 || co >= cout) return 1;
        if (convt_col(r, co, stride) != n) return 2;
        if (seen[static_cast<size_t>(r) * cout + co]++) return 3;
        // chunk q = n / 8 holds one phase of one 8-channel block; chunks of a block are adjacent
        if ((n >> 3) / stride != (co >> 3) || (n >> 3) % stride != r || (n & 7) != (co & 7)) return 4;
      }
    }
  std::puts("ok");
  return 0;
}
