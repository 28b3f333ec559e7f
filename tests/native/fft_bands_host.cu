// CPU harness for csrc/fft_passes.cuh: runs the band split and merge of fft_bands.cu -- the same
// pass bodies, pass plan and transform sequences, a loop over butterflies standing in for each
// kernel launch -- on float32 rows read from a file.
//   usage: fft_bands_host <in.bin> <out.bin> <batch> <n> <min_size> [mode = 3]
//   mode 0: full-length complex transforms, 1: real-input packing with accumulation passes,
//   2: real-input packing, merge gathered by the inverse loader, 3 (what the library runs):
//   as 2 with two passes per launch
//   out.bin = the bands in ascending size, (batch, size) float32 each, then the (batch, n)
//   recomposition of those bands.  tests/test_abi.py compares both with the oracle.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "fft_passes.cuh"

using namespace msb::fftb;

int main(int argc, char** argv) {
  if (argc != 6 && argc != 7) return 2;
  const int mode = argc == 7 ? std::atoi(argv[6]) : 3;
  const bool packed = mode != 0;
  const int batch = std::atoi(argv[3]), n = std::atoi(argv[4]), min_size = std::atoi(argv[5]);
  std::vector<float> x(static_cast<size_t>(batch) * n);
  FILE* fi = std::fopen(argv[1], "rb");
  if (fi == nullptr || std::fread(x.data(), sizeof(float), x.size(), fi) != x.size()) return 3;
  std::fclose(fi);
  const size_t bn = static_cast<size_t>(batch) * n;
  std::vector<float2> coef(bn), w0(bn), w1(bn), acc(static_cast<size_t>(batch) * (n / 2 + 1));
  std::vector<std::vector<float>> bands;
  std::vector<float*> band_ptr;
  std::vector<int> sizes;
  for (int s = min_size; s <= n; s <<= 1) {
    bands.emplace_back(static_cast<size_t>(batch) * s);
    sizes.push_back(s);
  }
  for (auto& b : bands) band_ptr.push_back(b.data());
  struct HostLauncher {
    bool fuse;
    bool fuses() const { return fuse; }
    int operator()(int radix, int load, int store, const PassArgs& a) const {
      return dispatch(radix, load, store, [&](auto r, auto ld, auto st) -> int {
        for (size_t gid = 0; gid < a.total; ++gid)
          pass_thread<decltype(r)::value, decltype(ld)::value, decltype(st)::value>(a, gid);
        return 0;
      });
    }
    int fused(int r1, int load, int store, const PassArgs& a) const {
      return dispatch_fused(r1, load, [&](auto r, auto ld) -> int {
        for (size_t g = 0; g < a.total; ++g)
          fused_group_host<decltype(r)::value, decltype(ld)::value>(a, store, g);
        return 0;
      });
    }
  };
  HostLauncher launch{mode == 3};
  auto accum = [&](const float2* zs, float2* ac, int S, int D, int lo, float scale, int first) {
    const size_t total = static_cast<size_t>(batch) * (D / 2 + 1);
    for (size_t gid = 0; gid < total; ++gid) accumulate_one(zs, ac, S, D, lo, scale, first, gid);
    return 0;
  };
  auto accum_pk = [&](const float2* zs, float2* ac, int S, int D, int lo, float scale, int first) {
    const size_t total = static_cast<size_t>(batch) * (D / 2 + 1);
    for (size_t gid = 0; gid < total; ++gid)
      accumulate_one_packed(zs, ac, S, D, lo, scale, first, gid);
    return 0;
  };
  // mode 3 also reads the loaders' twiddles from the table the library builds per call
  std::vector<float2> table(n / 2 + 1);
  for (int j = 0; j <= n / 2; ++j) {
    float s, c;
    sincospif(-static_cast<float>(j) / static_cast<float>(n / 2), &s, &c);
    table[j] = make_float2(c, s);
  }
  const float2* tw = mode == 3 ? table.data() : nullptr;
  if ((packed ? decompose_packed(x.data(), batch, n, min_size, band_ptr.data(), coef.data(),
                                 w0.data(), w1.data(), launch, tw, n / 2)
              : decompose(x.data(), batch, n, min_size, band_ptr.data(), coef.data(), w0.data(),
                          w1.data(), launch)) != 0)
    return 4;
  std::vector<float> y(bn);
  int rc_merge = -2;
  if (mode >= 2)
    rc_merge = recompose_merged(band_ptr.data(), sizes.data(), static_cast<int>(sizes.size()),
                                batch, n, y.data(), coef.data(), w0.data(), w1.data(), launch, tw,
                                n / 2);
  if (rc_merge != 0 && rc_merge != -2) return 7;
  std::fprintf(stderr, "merge path: %s\n", rc_merge == 0 ? "gathered" : "accumulated");
  if (rc_merge == -2 &&
      (packed ? recompose_packed(band_ptr.data(), sizes.data(), static_cast<int>(sizes.size()),
                                 batch, n, y.data(), acc.data(), w0.data(), w1.data(), launch,
                                 accum_pk)
              : recompose(band_ptr.data(), sizes.data(), static_cast<int>(sizes.size()), batch, n,
                          y.data(), acc.data(), w0.data(), w1.data(), launch, accum)) != 0)
    return 5;
  FILE* fo = std::fopen(argv[2], "wb");
  if (fo == nullptr) return 6;
  for (auto& b : bands) std::fwrite(b.data(), sizeof(float), b.size(), fo);
  std::fwrite(y.data(), sizeof(float), y.size(), fo);
  std::fclose(fo);
  return 0;
}
