// Register-resident 1024-point complex FFT for Audio2Mel (n_fft = 1024 = 16 * 16 * 4).
//
// 64 threads own one transform.  Decimation in frequency with n = 64 a + b, k = c + 16 d:
//   pass A  thread b          : DFT16 over a of x[64 a + b] (registers, loaded straight from
//                               global memory), twiddle W_1024^(b c), store S[c][b]
//   pass B  thread (c, b')    : DFT16 over a' of S[c][4 a' + b'], twiddle W_64^(b' c'), store
//                               S[c][4 c' + b'] IN PLACE (the thread owns the same 16 slots)
//   pass C  thread (c, c'+4i) : DFT4 over b' -> X[c + 16 c' + 256 d'] (natural order)
// S rows are padded to 68 floats: every shared-memory access of the three passes is
// bank-conflict free.  Two exchanges through shared memory instead of the five
// read-modify-write sweeps (plus twiddle-table reads) of the radix-4 in-place form.
//
// The per-thread pass bodies are __host__ __device__ so tests/native/a2m_fft_host.cu can run
// the very same arithmetic on the CPU, thread by thread, against a double-precision DFT.
#pragma once

#include <cuda_runtime.h>

namespace msb {
namespace a2m {

constexpr int kN = 1024;
constexpr int kLd = 68;             // padded row of the 16 x 64 exchange array
constexpr int kPlane = 16 * kLd;    // floats per real (or imaginary) plane of one transform

#define A2M_HD __host__ __device__ __forceinline__

// forward 4-point DFT (W_4 = -i), in place
A2M_HD void dft4(float& r0, float& i0, float& r1, float& i1, float& r2, float& i2, float& r3,
                 float& i3) {
  const float t0r = r0 + r2, t0i = i0 + i2, t1r = r0 - r2, t1i = i0 - i2;
  const float t2r = r1 + r3, t2i = i1 + i3, t3r = r1 - r3, t3i = i1 - i3;
  r0 = t0r + t2r; i0 = t0i + t2i;
  r2 = t0r - t2r; i2 = t0i - t2i;
  r1 = t1r + t3i; i1 = t1i - t3r;   // t1 - i t3
  r3 = t1r - t3i; i3 = t1i + t3r;   // t1 + i t3
}

A2M_HD void cmul(float& r, float& i, float c, float s) {
  const float t = r * c - i * s;
  i = r * s + i * c;
  r = t;
}

// forward 16-point DFT, natural order in and out, fully unrolled (registers only).
// n = 4 a + b, k = c + 4 d:  X[c + 4 d] = sum_b W_4^(b d) W_16^(b c) sum_a x[4 a + b] W_4^(a c)
A2M_HD void dft16(float (&re)[16], float (&im)[16]) {
  constexpr float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f;   // cos, sin (pi / 8)
  constexpr float H = 0.70710678118654752f;
#pragma unroll
  for (int b = 0; b < 4; ++b)
    dft4(re[b], im[b], re[4 + b], im[4 + b], re[8 + b], im[8 + b], re[12 + b], im[12 + b]);
  // now re[4 c + b] = y[c][b]; twiddles W_16^(b c) = exp(-2 pi i b c / 16)
  cmul(re[4 * 1 + 1], im[4 * 1 + 1], C1, -S1);     // bc = 1
  cmul(re[4 * 1 + 2], im[4 * 1 + 2], H, -H);       // 2
  cmul(re[4 * 1 + 3], im[4 * 1 + 3], S1, -C1);     // 3
  cmul(re[4 * 2 + 1], im[4 * 2 + 1], H, -H);       // 2
  {                                                // 4: multiply by -i
    const float t = re[4 * 2 + 2];
    re[4 * 2 + 2] = im[4 * 2 + 2];
    im[4 * 2 + 2] = -t;
  }
  cmul(re[4 * 2 + 3], im[4 * 2 + 3], -H, -H);      // 6
  cmul(re[4 * 3 + 1], im[4 * 3 + 1], S1, -C1);     // 3
  cmul(re[4 * 3 + 2], im[4 * 3 + 2], -H, -H);      // 6
  cmul(re[4 * 3 + 3], im[4 * 3 + 3], -C1, S1);     // 9
#pragma unroll
  for (int c = 0; c < 4; ++c)
    dft4(re[4 * c], im[4 * c], re[4 * c + 1], im[4 * c + 1], re[4 * c + 2], im[4 * c + 2],
         re[4 * c + 3], im[4 * c + 3]);
  // re[4 c + d] = X[c + 4 d]: transpose to natural order
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int d = c + 1; d < 4; ++d) {
      float t = re[4 * c + d]; re[4 * c + d] = re[4 * d + c]; re[4 * d + c] = t;
      t = im[4 * c + d]; im[4 * c + d] = im[4 * d + c]; im[4 * d + c] = t;
    }
}

// v[k] *= w^k, k = 1..15, w = (c1, s1) on the unit circle.  Powers by squaring / products of
// at most four factors (w^2, w^4, w^8 then sums of those), so the rounding error stays at a
// few ulp instead of growing linearly as in a running product.
A2M_HD void twiddle16(float (&re)[16], float (&im)[16], float c1, float s1) {
  const float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
  const float c4 = c2 * c2 - s2 * s2, s4 = 2.f * c2 * s2;
  const float c8 = c4 * c4 - s4 * s4, s8 = 2.f * c4 * s4;
  const float c3 = c2 * c1 - s2 * s1, s3 = c2 * s1 + s2 * c1;
  cmul(re[1], im[1], c1, s1);
  cmul(re[2], im[2], c2, s2);
  cmul(re[3], im[3], c3, s3);
  cmul(re[4], im[4], c4, s4);
  cmul(re[8], im[8], c8, s8);
  float c, s;
  c = c4 * c1 - s4 * s1; s = c4 * s1 + s4 * c1; cmul(re[5], im[5], c, s);
  c = c4 * c2 - s4 * s2; s = c4 * s2 + s4 * c2; cmul(re[6], im[6], c, s);
  c = c4 * c3 - s4 * s3; s = c4 * s3 + s4 * c3; cmul(re[7], im[7], c, s);
  c = c8 * c1 - s8 * s1; s = c8 * s1 + s8 * c1; cmul(re[9], im[9], c, s);
  c = c8 * c2 - s8 * s2; s = c8 * s2 + s8 * c2; cmul(re[10], im[10], c, s);
  c = c8 * c3 - s8 * s3; s = c8 * s3 + s8 * c3; cmul(re[11], im[11], c, s);
  const float c12 = c8 * c4 - s8 * s4, s12 = c8 * s4 + s8 * c4;
  cmul(re[12], im[12], c12, s12);
  c = c12 * c1 - s12 * s1; s = c12 * s1 + s12 * c1; cmul(re[13], im[13], c, s);
  c = c12 * c2 - s12 * s2; s = c12 * s2 + s12 * c2; cmul(re[14], im[14], c, s);
  c = c12 * c3 - s12 * s3; s = c12 * s3 + s12 * c3; cmul(re[15], im[15], c, s);
}

// pass A.  t = thread within the transform (0..63) = b; re/im hold x[64 a + b], a = 0..15.
// sr / si: the two planes of this transform (kPlane floats each).
A2M_HD void pass_a(int t, float (&re)[16], float (&im)[16], float* sr, float* si) {
  dft16(re, im);
  float s1, c1;
  sincospif(-static_cast<float>(t) * (2.f / kN), &s1, &c1);
  twiddle16(re, im, c1, s1);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    sr[c * kLd + t] = re[c];
    si[c * kLd + t] = im[c];
  }
}

// pass B, in place.  t -> (c = t / 4, b' = t % 4)
A2M_HD void pass_b(int t, float* sr, float* si) {
  const int base = (t >> 2) * kLd + (t & 3);
  float re[16], im[16];
#pragma unroll
  for (int a = 0; a < 16; ++a) {
    re[a] = sr[base + 4 * a];
    im[a] = si[base + 4 * a];
  }
  dft16(re, im);
  float s1, c1;
  sincospif(-static_cast<float>(t & 3) * (2.f / 64.f), &s1, &c1);
  twiddle16(re, im, c1, s1);
#pragma unroll
  for (int c = 0; c < 16; ++c) {
    sr[base + 4 * c] = re[c];
    si[base + 4 * c] = im[c];
  }
}

// pass C, read half.  t -> c = t % 16, c' = t / 16 + 4 i.  Results stay in registers until
// every thread of the CTA has finished reading the padded layout.
A2M_HD void pass_c_read(int t, const float* sr, const float* si, float (&re)[16],
                        float (&im)[16]) {
  const int c = t & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cp = (t >> 4) + 4 * i;
    const float4 r = *reinterpret_cast<const float4*>(sr + c * kLd + 4 * cp);
    const float4 m = *reinterpret_cast<const float4*>(si + c * kLd + 4 * cp);
    re[4 * i] = r.x; re[4 * i + 1] = r.y; re[4 * i + 2] = r.z; re[4 * i + 3] = r.w;
    im[4 * i] = m.x; im[4 * i + 1] = m.y; im[4 * i + 2] = m.z; im[4 * i + 3] = m.w;
    dft4(re[4 * i], im[4 * i], re[4 * i + 1], im[4 * i + 1], re[4 * i + 2], im[4 * i + 2],
         re[4 * i + 3], im[4 * i + 3]);
  }
}

// pass C, write half: natural-order spectrum X[k], k < 1024, over the same planes
A2M_HD void pass_c_write(int t, const float (&re)[16], const float (&im)[16], float* xr,
                         float* xi) {
  const int c = t & 15;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int cp = (t >> 4) + 4 * i;
#pragma unroll
    for (int d = 0; d < 4; ++d) {
      xr[c + 16 * cp + 256 * d] = re[4 * i + d];
      xi[c + 16 * cp + 256 * d] = im[4 * i + d];
    }
  }
}

}  // namespace a2m
}  // namespace msb
