// CPU harness for csrc/a2m_fft.cuh: runs the per-thread pass bodies of the Audio2Mel FFT
// (the same __host__ __device__ functions the kernel calls) thread by thread, a loop boundary
// standing in for each __syncthreads(), and compares with a double-precision DFT.
// Prints the relative L2 error; exit status 0 iff it is below 2e-6.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "a2m_fft.cuh"

using namespace msb::a2m;

int main() {
  std::vector<float> xr(kN), xi(kN);
  unsigned s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return static_cast<float>((s >> 8) & 0xffff) / 32768.f - 1.f;
  };
  for (int i = 0; i < kN; ++i) { xr[i] = rnd(); xi[i] = rnd(); }
  std::vector<float> sr(kPlane, 0.f), si(kPlane, 0.f);
  for (int t = 0; t < 64; ++t) {
    float re[16], im[16];
    for (int a = 0; a < 16; ++a) { re[a] = xr[64 * a + t]; im[a] = xi[64 * a + t]; }
    pass_a(t, re, im, sr.data(), si.data());
  }
  for (int t = 0; t < 64; ++t) pass_b(t, sr.data(), si.data());
  std::vector<float> keep_r(64 * 16), keep_i(64 * 16);
  for (int t = 0; t < 64; ++t) {
    float re[16], im[16];
    pass_c_read(t, sr.data(), si.data(), re, im);
    for (int j = 0; j < 16; ++j) { keep_r[t * 16 + j] = re[j]; keep_i[t * 16 + j] = im[j]; }
  }
  for (int t = 0; t < 64; ++t) {
    float re[16], im[16];
    for (int j = 0; j < 16; ++j) { re[j] = keep_r[t * 16 + j]; im[j] = keep_i[t * 16 + j]; }
    pass_c_write(t, re, im, sr.data(), si.data());
  }
  double num = 0, den = 0;
  for (int k = 0; k < kN; ++k) {
    double ar = 0, ai = 0;
    for (int n = 0; n < kN; ++n) {
      const double ph = -2.0 * M_PI * static_cast<double>((static_cast<long long>(n) * k) % kN) / kN;
      const double c = std::cos(ph), sn = std::sin(ph);
      ar += xr[n] * c - xi[n] * sn;
      ai += xr[n] * sn + xi[n] * c;
    }
    num += (sr[k] - ar) * (sr[k] - ar) + (si[k] - ai) * (si[k] - ai);
    den += ar * ar + ai * ai;
  }
  const double rel = std::sqrt(num / den);
  std::printf("rel_l2 %.3e\n", rel);
  return rel < 2e-6 ? 0 : 1;
}
