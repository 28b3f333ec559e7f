// Stockham autosort FFT passes of radix 2 / 4 / 8 / 16 for the octave-band split / merge
// (fft_bands.cu), with the boundary work folded into the first and last pass of a transform:
//   first pass loads   complex data | real data (imag = 0) | a band's Hermitian spectrum cut
//                      out of the full-length coefficients | the full Hermitian spectrum
//                      expanded from its stored half
//   last pass stores   complex data | the real part only
// Every pass is forward (exp(-i...)) arithmetic: an inverse transform is conj . forward . conj,
// the input conjugation sits in the two spectrum loaders and the output conjugation is moot
// because only the real part is kept.
//
// One pass, radix R, p = product of the radices already applied, t = n / R, thread i < t of a
// row:  k = i mod p,  v_q = x[i + q t] W_(R p)^(q k),  y[(i - k) R + k + m p] = sum_q v_q W_R^(q m).
//
// Two consecutive passes (radix R1, then radix 16) can go out as one launch that exchanges
// through shared memory ("two passes in one launch" below): the launcher opts in with fuses().
//
// The per-thread body and the pass plan are __host__ __device__ / host templates so that
// tests/native/fft_bands_host.cu runs the same code on the CPU.
#pragma once

#include <cuda_runtime.h>

#include <cmath>
#include <type_traits>

#include "a2m_fft.cuh"

namespace msb {
namespace fftb {

enum Load {
  kLoadComplex = 0, kLoadReal = 1, kLoadBand = 2, kLoadHalf = 3,
  // real-input packing (default): a real sequence of length 2h is transformed as the h-point
  // complex sequence x[2m] + i x[2m+1]; these loaders build the packed inverse input V[k]
  // straight from the packed forward output Z (band cut-out) or from the stored half spectrum
  kLoadBandPk = 4, kLoadHalfPk = 5,
  // merge: the Hermitian spectrum is the scaled sum of the kept bins of every band's packed
  // forward transform, gathered while loading (no accumulation passes over a spectrum array)
  kLoadMergePk = 6
};

constexpr int kMaxBands = 12;
struct BandTable {        // kLoadMergePk: band b = rows of size[b]/2 float2 at x + offset[b]
  int count;
  int size[kMaxBands];
  int lo[kMaxBands];
  float scale[kMaxBands];
  long long offset[kMaxBands];
  // which bands can keep bin k without walking the list: the band whose half-length is the power
  // of two at or above k (keeps [hb/2, hb]), the next one up when k is a power of two (k is its
  // lowest kept bin), and the smallest band (keeps [0, hb]).  by_log[L] = band with hb = 2^L or -1.
  int by_log[32];
  int smallest;
};
enum Store { kStoreComplex = 0, kStoreReal = 1, kStoreConj = 2 };

struct PassArgs {
  const void* x;   // float2 rows of n | float rows of n | float2 rows of src_n (band / half)
  void* y;         // float2 rows of n | float rows of n
  int n;           // transform length
  int p;           // product of the radices of the earlier passes
  size_t total;    // batch * n / R threads
  int src_n;       // kLoadBand: row length of the coefficient array; kLoadHalf: n / 2 + 1
  int lo;          // kLoadBand: first kept bin
  float scale;     // kLoadBand: applied to the kept bins
  BandTable bands; // kLoadMergePk
  // optional table tw[j] = exp(-i pi j / tw_h), 0 <= j <= tw_h (tw_h a power of two that every
  // half-length of the call divides): the three loaders that un-tangle packed spectra read their
  // twiddles from it instead of evaluating sincospif three to five times per element
  const float2* tw = nullptr;
  int tw_h = 0;
};

// log2 of a power of two (every length in this file is one: divisions become shifts)
A2M_HD int ilog2(int v) {
#ifdef __CUDA_ARCH__
  return 31 - __clz(v);
#else
  int l = 0;
  while ((1 << l) < v) ++l;
  return l;
#endif
}

// smallest L with 2^L >= k, k >= 1
A2M_HD int ceil_log2(int k) {
#ifdef __CUDA_ARCH__
  return 32 - __clz(k - 1);
#else
  int l = 0;
  while ((1 << l) < k) ++l;
  return l;
#endif
}

// exp(-i pi num / den), 0 <= num <= den, den a power of two: same bits from the table (which
// holds sincospif of the same argument) or computed in place
A2M_HD void unit_minus(const PassArgs& a, int num, int den, float& c, float& s) {
  if (a.tw != nullptr) {
    const float2 w = a.tw[num << (ilog2(a.tw_h) - ilog2(den))];
    c = w.x;
    s = w.y;
  } else {
    sincospif(-static_cast<float>(num) / static_cast<float>(den), &s, &c);
  }
}

// forward 8-point DFT, natural order in and out (n = 4a + b, k = c + 2d)
A2M_HD void dft8(float (&re)[8], float (&im)[8]) {
  constexpr float H = 0.70710678118654752f;
#pragma unroll
  for (int b = 0; b < 4; ++b) {
    const float ur = re[b], ui = im[b];
    re[b] = ur + re[4 + b]; im[b] = ui + im[4 + b];
    re[4 + b] = ur - re[4 + b]; im[4 + b] = ui - im[4 + b];
  }
  a2m::cmul(re[5], im[5], H, -H);                 // W_8^1
  {                                               // W_8^2 = -i
    const float t = re[6];
    re[6] = im[6];
    im[6] = -t;
  }
  a2m::cmul(re[7], im[7], -H, -H);                // W_8^3
  a2m::dft4(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3]);
  a2m::dft4(re[4], im[4], re[5], im[5], re[6], im[6], re[7], im[7]);
  // position 4c + d holds X[c + 2d]
  const float r[8] = {re[0], re[4], re[1], re[5], re[2], re[6], re[3], re[7]};
  const float i[8] = {im[0], im[4], im[1], im[5], im[2], im[6], im[3], im[7]};
#pragma unroll
  for (int k = 0; k < 8; ++k) { re[k] = r[k]; im[k] = i[k]; }
}

// v[q] *= w^q, q = 1..7
A2M_HD void twiddle8(float (&re)[8], float (&im)[8], float c1, float s1) {
  const float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
  const float c4 = c2 * c2 - s2 * s2, s4 = 2.f * c2 * s2;
  const float c3 = c2 * c1 - s2 * s1, s3 = c2 * s1 + s2 * c1;
  a2m::cmul(re[1], im[1], c1, s1);
  a2m::cmul(re[2], im[2], c2, s2);
  a2m::cmul(re[3], im[3], c3, s3);
  a2m::cmul(re[4], im[4], c4, s4);
  float c, s;
  c = c4 * c1 - s4 * s1; s = c4 * s1 + s4 * c1; a2m::cmul(re[5], im[5], c, s);
  c = c4 * c2 - s4 * s2; s = c4 * s2 + s4 * c2; a2m::cmul(re[6], im[6], c, s);
  c = c4 * c3 - s4 * s3; s = c4 * s3 + s4 * c3; a2m::cmul(re[7], im[7], c, s);
}

// X[k], 0 <= k <= h, of a real sequence of length 2h from Z = FFT_h(x[2m] + i x[2m+1]):
//   X[k] = (Z[k] + conj(Z[h-k])) / 2 - i W_2h^k (Z[k] - conj(Z[h-k])) / 2,   Z[h] := Z[0]
A2M_HD float2 real_bin_cs(const float2* Z, int h, int k, float c, float s) {
  const float2 a = Z[k == h ? 0 : k];
  const float2 b = Z[(k == 0 || k == h) ? 0 : h - k];
  const float er = 0.5f * (a.x + b.x), ei = 0.5f * (a.y - b.y);
  const float dr = 0.5f * (a.x - b.x), di = 0.5f * (a.y + b.y);
  return make_float2(er + (c * di + s * dr), ei - (c * dr - s * di));
}
A2M_HD float2 real_bin(const PassArgs& pa, const float2* Z, int h, int k) {
  float s, c;
  unit_minus(pa, k, h, c, s);
  return real_bin_cs(Z, h, k, c, s);
}
A2M_HD float2 real_bin(const float2* Z, int h, int k) {     // accumulation passes: no table
  float s, c;
  sincospif(-static_cast<float>(k) / static_cast<float>(h), &s, &c);
  return real_bin_cs(Z, h, k, c, s);
}

// conj(V[k]), 0 <= k < H, the packed input of the H-point transform whose output is
// y[2m] - i y[2m+1] for the real sequence y of length 2H with Hermitian spectrum Y[0..H]:
//   V[k] = (Y[k] + conj(Y[H-k])) + i W_2H^-k (Y[k] - conj(Y[H-k]))
A2M_HD void packed_inverse_input(const PassArgs& pa, float2 ya, float2 yb, int k, int H,
                                 float& re, float& im) {
  const float px = ya.x + yb.x, py = ya.y - yb.y;
  const float qx = ya.x - yb.x, qy = ya.y + yb.y;
  float s, c;
  unit_minus(pa, k, H, c, s);      // exp(+i pi k / H) is its conjugate
  s = -s;
  const float ex = c * qx - s * qy, ey = c * qy + s * qx;
  re = px - ey;
  im = -(py + ex);
}

template <int LD>
A2M_HD void load_one(const PassArgs& a, size_t row, int idx, float& re, float& im) {
  if (LD == kLoadComplex) {
    const float2 v = static_cast<const float2*>(a.x)[row * a.n + idx];
    re = v.x;
    im = v.y;
  } else if (LD == kLoadReal) {
    re = static_cast<const float*>(a.x)[row * a.n + idx];
    im = 0.f;
  } else if (LD == kLoadBand) {
    // Z[k] = scale * C[k] for lo <= k <= n/2 (imaginary part of k = 0 and k = n/2 dropped, as a
    // c2r transform ignores it), Z[n-k] = conj(Z[k]), 0 elsewhere; conjugated for the inverse
    const int half = a.n >> 1;
    const int kk = idx <= half ? idx : a.n - idx;
    re = 0.f;
    im = 0.f;
    if (kk >= a.lo) {
      const float2 v = static_cast<const float2*>(a.x)[row * a.src_n + kk];
      re = v.x * a.scale;
      im = (kk == 0 || kk == half) ? 0.f : (idx <= half ? -a.scale : a.scale) * v.y;
    }
  } else if (LD == kLoadHalf) {
    // full Hermitian spectrum from its n/2+1 half; conjugated for the inverse
    const int half = a.n >> 1;
    const int kk = idx <= half ? idx : a.n - idx;
    const float2 v = static_cast<const float2*>(a.x)[row * a.src_n + kk];
    re = v.x;
    im = (kk == 0 || kk == half) ? 0.f : (idx <= half ? -v.y : v.y);
  } else if (LD == kLoadBandPk) {
    // band of size S = 2 a.n cut out of the packed forward transform Z (rows of src_n = h):
    // Y[k] = scale * X[k] for lo <= k <= S/2 (imaginary part of k = 0 and k = S/2 dropped)
    const int H = a.n;
    const float2* Z = static_cast<const float2*>(a.x) + row * a.src_n;
    float2 y[2];
    if (2 * a.lo == H) {
      // upper-half band (every band but the lowest): bins idx and H - idx lie on either side of
      // lo, so ONE kept bin feeds the element (both are the same bin when idx = lo) -- one
      // un-tangling and no data-dependent branch
      const int k = idx >= a.lo ? idx : H - idx;
      const float2 x = real_bin(a, Z, a.src_n, k);
      const float2 v = make_float2(x.x * a.scale, k == H ? 0.f : x.y * a.scale);
      y[0] = idx >= a.lo ? v : make_float2(0.f, 0.f);
      y[1] = idx <= a.lo ? v : make_float2(0.f, 0.f);
    } else {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int k = e == 0 ? idx : H - idx;
        y[e] = make_float2(0.f, 0.f);
        if (k >= a.lo) {
          const float2 x = real_bin(a, Z, a.src_n, k);
          y[e] = make_float2(x.x * a.scale, (k == 0 || k == H) ? 0.f : x.y * a.scale);
        }
      }
    }
    packed_inverse_input(a, y[0], y[1], idx, H, re, im);
  } else if (LD == kLoadMergePk) {
    // Y[k] = sum over bands of scale_b * X_b[k] for lo_b <= k <= S_b/2 (fft_resample +
    // the sum of fft_frequency_recompose, audio/transform.py:85-115), X_b un-tangled from
    // band b's packed transform; imaginary part of k = 0 and k = H dropped
    const int H = a.n;
    const float2* base = static_cast<const float2*>(a.x);
    float2 y[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int k = e == 0 ? idx : H - idx;
      // primary band: half-length = the power of two at or above k (k lies in its kept upper
      // half), else the smallest band when k is inside it; un-tangled without a branch (scale 0
      // and a safe address when no band keeps the bin)
      const int L = k > 0 ? ceil_log2(k) : 0;
      const int sm = a.bands.smallest;
      int p = k > 0 ? a.bands.by_log[L] : -1;
      if (p < 0 && k <= (a.bands.size[sm] >> 1)) p = sm;
      const int pb = p < 0 ? sm : p;
      const int hb = a.bands.size[pb] >> 1;
      const float sc = p < 0 ? 0.f : a.bands.scale[pb];
      const float2 x = real_bin(a, base + a.bands.offset[pb] + row * hb, hb, p < 0 ? 0 : k);
      float yr = x.x * sc, yi = x.y * sc;
      if (k > 0 && (k & (k - 1)) == 0) {
        // a power of two is also the lowest kept bin of the next band up (the reference's
        // double-counted boundary bins): a handful of bins per row
        const int b2 = a.bands.by_log[L + 1];
        if (b2 >= 0 && b2 != p && k >= a.bands.lo[b2]) {
          const int h2 = a.bands.size[b2] >> 1;
          const float2 x2 = real_bin(a, base + a.bands.offset[b2] + row * h2, h2, k);
          yr += x2.x * a.bands.scale[b2];
          yi += x2.y * a.bands.scale[b2];
        }
      }
      y[e] = make_float2(yr, (k == 0 || k == H) ? 0.f : yi);
    }
    packed_inverse_input(a, y[0], y[1], idx, H, re, im);
  } else {
    // kLoadHalfPk: Y[0 .. H] stored (rows of src_n = H + 1), transform length H = a.n
    const int H = a.n;
    const float2* Y = static_cast<const float2*>(a.x) + row * a.src_n;
    float2 ya = Y[idx], yb = Y[H - idx];
    if (idx == 0) { ya.y = 0.f; yb.y = 0.f; }
    packed_inverse_input(a, ya, yb, idx, H, re, im);
  }
}

// v[q] *= exp(-2 pi i q k / (R p)), then the R-point DFT (natural order in and out)
template <int R>
A2M_HD void twiddle_dft(float (&re)[R], float (&im)[R], int k, int p) {
  float s1 = 0.f, c1 = 1.f;
  if (p > 1)
    sincospif(-2.f * static_cast<float>(k) / (static_cast<float>(R) * static_cast<float>(p)),
              &s1, &c1);
  if constexpr (R == 16) {
    if (p > 1) a2m::twiddle16(re, im, c1, s1);
    a2m::dft16(re, im);
  } else if constexpr (R == 8) {
    if (p > 1) twiddle8(re, im, c1, s1);
    dft8(re, im);
  } else if constexpr (R == 4) {
    if (p > 1) {
      const float c2 = c1 * c1 - s1 * s1, s2 = 2.f * c1 * s1;
      const float c3 = c2 * c1 - s2 * s1, s3 = c2 * s1 + s2 * c1;
      a2m::cmul(re[1], im[1], c1, s1);
      a2m::cmul(re[2], im[2], c2, s2);
      a2m::cmul(re[3], im[3], c3, s3);
    }
    a2m::dft4(re[0], im[0], re[1], im[1], re[2], im[2], re[3], im[3]);
  } else {
    if (p > 1) a2m::cmul(re[1], im[1], c1, s1);
    const float ur = re[0], ui = im[0];
    re[0] = ur + re[1]; im[0] = ui + im[1];
    re[1] = ur - re[1]; im[1] = ui - im[1];
  }
}

// loads, twiddles and transforms the R inputs of butterfly `gid`; returns the output base
// index j (outputs go to row * n + j + m * p)
template <int R, int LD>
A2M_HD void pass_compute(const PassArgs& a, size_t gid, float (&re)[R], float (&im)[R],
                         size_t& row, int& j) {
  const int t = a.n / R;
  const int i = static_cast<int>(gid & static_cast<size_t>(t - 1));
  row = gid >> ilog2(t);
  const int k = i & (a.p - 1);
  j = (i - k) * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) load_one<LD>(a, row, i + q * t, re[q], im[q]);
  twiddle_dft<R>(re, im, k, a.p);
}

template <int R, int LD, int ST>
A2M_HD void pass_thread(const PassArgs& a, size_t gid) {
  float re[R], im[R];
  size_t row;
  int j;
  pass_compute<R, LD>(a, gid, re, im, row, j);
  const size_t o = row * a.n + j;
#pragma unroll
  for (int m = 0; m < R; ++m) {
    if (ST == kStoreComplex)
      static_cast<float2*>(a.y)[o + static_cast<size_t>(m) * a.p] = make_float2(re[m], im[m]);
    else if (ST == kStoreConj)     // packed inverse: (y[2j], y[2j+1]) = conj of the forward result
      static_cast<float2*>(a.y)[o + static_cast<size_t>(m) * a.p] = make_float2(re[m], -im[m]);
    else
      static_cast<float*>(a.y)[o + static_cast<size_t>(m) * a.p] = re[m];
  }
}

// ---- two passes in one launch ---------------------------------------------------------------
// A radix-R1 pass at p followed by a radix-16 pass at R1 p is one radix-G pass, G = 16 R1: group
// g = a2 p + k1 (k1 < p, g < n / G) reads x[g + e n / G], e < G, and writes
// y[a2 G p + k1 + e' p], e' < G.  Inside the group, first-half thread q1 < 16 takes the inputs
// e = q1 + 16 q (q < R1), and its output m is input q1 of second-half thread m2 = m < R1, which
// twiddles with k2 = m2 p + k1 and produces e' = m2 + R1 m, m < 16.  The 16 x R1 exchange goes
// through `ex` (row m2 at ex + 17 m2: pitch 17 keeps both sides free of bank conflicts).
// PassArgs.total counts groups (batch * n / G).
constexpr int kExPitch = 17;

A2M_HD void store_any(void* y, int store, size_t idx, float re, float im) {
  if (store == kStoreComplex) static_cast<float2*>(y)[idx] = make_float2(re, im);
  else if (store == kStoreConj) static_cast<float2*>(y)[idx] = make_float2(re, -im);
  else static_cast<float*>(y)[idx] = re;
}

template <int R1, int LD>
A2M_HD void fused_first_half(const PassArgs& a, size_t group, int q1, float2* ex) {
  const int per_row = a.n / (16 * R1);
  const size_t row = group >> ilog2(per_row);
  const int g = static_cast<int>(group & static_cast<size_t>(per_row - 1));
  float re[R1], im[R1];
#pragma unroll
  for (int q = 0; q < R1; ++q) load_one<LD>(a, row, g + (q1 + 16 * q) * per_row, re[q], im[q]);
  twiddle_dft<R1>(re, im, g & (a.p - 1), a.p);
#pragma unroll
  for (int m = 0; m < R1; ++m) ex[m * kExPitch + q1] = make_float2(re[m], im[m]);
}

// output m of the thread goes to index o + (m2 + R1 m) p
template <int R1>
A2M_HD void fused_second_half(const PassArgs& a, size_t group, int m2, const float2* ex,
                              float (&re)[16], float (&im)[16], size_t& o) {
  const int per_row = a.n / (16 * R1);
  const size_t row = group >> ilog2(per_row);
  const int g = static_cast<int>(group & static_cast<size_t>(per_row - 1));
  const int k1 = g & (a.p - 1);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    const float2 v = ex[m2 * kExPitch + q];
    re[q] = v.x;
    im[q] = v.y;
  }
  twiddle_dft<16>(re, im, m2 * a.p + k1, R1 * a.p);
  o = row * a.n + static_cast<size_t>(g - k1) * (16 * R1) + k1;
}

// the whole group on one thread (host harness)
template <int R1, int LD>
void fused_group_host(const PassArgs& a, int store, size_t group) {
  float2 ex[R1 * kExPitch];
  for (int q1 = 0; q1 < 16; ++q1) fused_first_half<R1, LD>(a, group, q1, ex);
  for (int m2 = 0; m2 < R1; ++m2) {
    float re[16], im[16];
    size_t o;
    fused_second_half<R1>(a, group, m2, ex, re, im, o);
    for (int m = 0; m < 16; ++m)
      store_any(a.y, store, o + static_cast<size_t>(m2 + R1 * m) * a.p, re[m], im[m]);
  }
}

// a launcher may offer fused launches: bool fuses() const; int fused(r1, load, store, args)
template <class L>
auto launcher_fuses(const L& l, int) -> decltype(l.fuses()) { return l.fuses(); }
template <class L>
bool launcher_fuses(const L&, long) { return false; }
template <class L>
auto launch_fused(L& l, int r1, int load, int store, const PassArgs& a, int)
    -> decltype(l.fused(r1, load, store, a)) { return l.fused(r1, load, store, a); }
template <class L>
int launch_fused(L&, int, int, int, const PassArgs&, long) { return -1; }

// run-time (radix, load, store) -> compile-time constants: f(IC<R>, IC<LD>, IC<ST>) -> int
template <int V>
using IC = std::integral_constant<int, V>;

template <class F>
int dispatch(int radix, int load, int store, F&& f) {
  auto with_store = [&](auto r, auto ld) -> int {
    if (store == kStoreComplex) return f(r, ld, IC<kStoreComplex>{});
    if (store == kStoreConj) return f(r, ld, IC<kStoreConj>{});
    return f(r, ld, IC<kStoreReal>{});
  };
  auto with_load = [&](auto r) -> int {
    switch (load) {
      case kLoadComplex: return with_store(r, IC<kLoadComplex>{});
      case kLoadReal: return with_store(r, IC<kLoadReal>{});
      case kLoadBand: return with_store(r, IC<kLoadBand>{});
      case kLoadHalf: return with_store(r, IC<kLoadHalf>{});
      case kLoadBandPk: return with_store(r, IC<kLoadBandPk>{});
      case kLoadHalfPk: return with_store(r, IC<kLoadHalfPk>{});
      default: return with_store(r, IC<kLoadMergePk>{});
    }
  };
  switch (radix) {
    case 2: return with_load(IC<2>{});
    case 4: return with_load(IC<4>{});
    case 8: return with_load(IC<8>{});
    case 16: return with_load(IC<16>{});
    default: return -1;
  }
}

// run-time (first radix, load) of a fused launch -> compile-time constants: f(IC<R1>, IC<LD>)
template <class F>
int dispatch_fused(int r1, int load, F&& f) {
  auto with_load = [&](auto r) -> int {
    switch (load) {
      case kLoadComplex: return f(r, IC<kLoadComplex>{});
      case kLoadReal: return f(r, IC<kLoadReal>{});
      case kLoadBand: return f(r, IC<kLoadBand>{});
      case kLoadHalf: return f(r, IC<kLoadHalf>{});
      case kLoadBandPk: return f(r, IC<kLoadBandPk>{});
      case kLoadHalfPk: return f(r, IC<kLoadHalfPk>{});
      default: return f(r, IC<kLoadMergePk>{});
    }
  };
  switch (r1) {
    case 2: return with_load(IC<2>{});
    case 4: return with_load(IC<4>{});
    case 8: return with_load(IC<8>{});
    case 16: return with_load(IC<16>{});
    default: return -1;
  }
}

// ---- the plan of one batched transform --------------------------------------------------
struct Xform {
  int n, batch;
  int load;            // Load of the first pass
  const void* src;
  int src_n, lo;
  float scale;
  int store;           // Store of the last pass
  void* dst;
  float2* w0;          // ping-pong scratch for the passes in between (batch * n each)
  float2* w1;
  const BandTable* bands = nullptr;   // kLoadMergePk
  const float2* tw = nullptr;         // PassArgs::tw of the first pass
  int tw_h = 0;
};

inline int plan_radices(int n, int* radix) {   // n = power of two >= 2; returns the pass count
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  int c = 0;
  if ((log2n & 3) == 3) radix[c++] = 8;
  else if (log2n & 1) radix[c++] = 2;
  else if (log2n & 2) radix[c++] = 4;
  for (int s = log2n >> 2; s > 0; --s) radix[c++] = 16;
  return c;
}

// number of launches of an n-point transform: with `fuse`, the first two passes and every
// following pair of radix-16 passes go out as one launch each
inline int plan_launch_count(int n, bool fuse) {
  int radix[16];
  const int count = plan_radices(n, radix);
  return fuse ? (count + 1) / 2 : count;
}

// launch(radix, load, store, args) -> 0 on success
template <class Launch>
int run_xform(const Xform& x, Launch&& launch) {
  int radix[16];
  const int count = plan_radices(x.n, radix);
  const bool fuse = launcher_fuses(launch, 0);
  const void* src = x.src;
  int p = 1;
  for (int i = 0, li = 0; i < count; ++li) {
    const bool pair = fuse && i + 1 < count;       // radix[i + 1] is always 16
    const int step = pair ? 2 : 1;
    const bool last = i + step >= count;
    PassArgs a;
    a.x = src;
    a.y = last ? x.dst : static_cast<void*>((li & 1) ? x.w1 : x.w0);
    a.n = x.n;
    a.p = p;
    a.total = static_cast<size_t>(x.batch) * (x.n / (pair ? 16 * radix[i] : radix[i]));
    a.src_n = x.src_n;
    a.lo = x.lo;
    a.scale = x.scale;
    if (i == 0 && x.bands != nullptr) a.bands = *x.bands; else a.bands.count = 0;
    if (i == 0) { a.tw = x.tw; a.tw_h = x.tw_h; }
    const int load = i == 0 ? x.load : static_cast<int>(kLoadComplex);
    const int store = last ? x.store : static_cast<int>(kStoreComplex);
    const int rc = pair ? launch_fused(launch, radix[i], load, store, a, 0)
                        : launch(radix[i], load, store, a);
    if (rc != 0) return rc;
    src = a.y;
    p *= pair ? 16 * radix[i] : radix[i];
    i += step;
  }
  return 0;
}

// ---- the two public operations as sequences of transforms ---------------------------------
// decompose (audio/transform.py:50-82): forward transform of the whole clip into `coef`, then
// per band of size S an inverse transform of bins [S/4, S/2] (lowest band: [0, S/2]).
// ortho scaling of both directions (1/sqrt(n), 1/sqrt(S)) is applied where the band is cut out.
template <class Launch>
int decompose(const float* x, int batch, int n, int min_size, float* const* bands_out,
              float2* coef, float2* w0, float2* w1, Launch&& launch) {
  Xform f;
  f.n = n; f.batch = batch; f.load = kLoadReal; f.src = x; f.src_n = n; f.lo = 0; f.scale = 1.f;
  f.store = kStoreComplex; f.dst = coef; f.w0 = w0; f.w1 = w1;
  int rc = run_xform(f, launch);
  if (rc != 0) return rc;
  int bi = 0;
  for (int S = min_size; S <= n; S <<= 1, ++bi) {
    Xform b;
    b.n = S; b.batch = batch; b.load = kLoadBand; b.src = coef; b.src_n = n;
    b.lo = (S > min_size) ? S / 4 : 0;
    b.scale = 1.0f / (sqrtf(static_cast<float>(n)) * sqrtf(static_cast<float>(S)));
    b.store = kStoreReal; b.dst = bands_out[bi]; b.w0 = w0; b.w1 = w1;
    rc = run_xform(b, launch);
    if (rc != 0) return rc;
  }
  return 0;
}

// recompose (audio/transform.py:85-115): forward transform of each band, its kept bins scaled
// and summed into the half spectrum `acc` (batch rows of D/2+1), one inverse transform.
// accum(spectrum, acc, S, D, lo, scale, first) -> 0 on success
template <class Launch, class Accum>
int recompose(const float* const* bands, const int* sizes, int nbands, int batch, int D,
              float* out, float2* acc, float2* w0, float2* w1, Launch&& launch, Accum&& accum) {
  int smin = sizes[0];
  for (int i = 1; i < nbands; ++i) smin = sizes[i] < smin ? sizes[i] : smin;
  for (int i = 0; i < nbands; ++i) {
    const int S = sizes[i];
    const int count = plan_launch_count(S, launcher_fuses(launch, 0));
    Xform f;
    f.n = S; f.batch = batch; f.load = kLoadReal; f.src = bands[i]; f.src_n = S; f.lo = 0;
    f.scale = 1.f; f.store = kStoreComplex;
    f.dst = ((count - 1) & 1) ? w1 : w0;      // continues the ping-pong: never the last source
    f.w0 = w0; f.w1 = w1;
    int rc = run_xform(f, launch);
    if (rc != 0) return rc;
    // fft_resample: the lowest band keeps bins [0, S/2], the others [S/4, S/2]
    // (n_coeffs // 2 with n_coeffs = S/2 + 1), audio/transform.py:93-96
    const int lo = (S == smin) ? 0 : (S / 2 + 1) / 2;
    const float scale = 1.0f / (sqrtf(static_cast<float>(S)) * sqrtf(static_cast<float>(D)));
    rc = accum(static_cast<const float2*>(f.dst), acc, S, D, lo, scale, i == 0 ? 1 : 0);
    if (rc != 0) return rc;
  }
  Xform inv;
  inv.n = D; inv.batch = batch; inv.load = kLoadHalf; inv.src = acc; inv.src_n = D / 2 + 1;
  inv.lo = 0; inv.scale = 1.f; inv.store = kStoreReal; inv.dst = out; inv.w0 = w0; inv.w1 = w1;
  return run_xform(inv, launch);
}

// ---- the same two operations with real-input packing: every transform runs at half length --
// Z (batch rows of n/2) = packed forward transform of the clip; a band's inverse transform of
// length S/2 builds its input from Z while loading and stores (y[2j], y[2j+1]) pairs.
template <class Launch>
int decompose_packed(const float* x, int batch, int n, int min_size, float* const* bands_out,
                     float2* Z, float2* w0, float2* w1, Launch&& launch,
                     const float2* tw = nullptr, int tw_h = 0) {
  Xform f;
  f.n = n / 2; f.batch = batch; f.load = kLoadComplex; f.src = x; f.src_n = n / 2; f.lo = 0;
  f.scale = 1.f; f.store = kStoreComplex; f.dst = Z; f.w0 = w0; f.w1 = w1;
  int rc = run_xform(f, launch);
  if (rc != 0) return rc;
  int bi = 0;
  for (int S = min_size; S <= n; S <<= 1, ++bi) {
    Xform b;
    b.n = S / 2; b.batch = batch; b.load = kLoadBandPk; b.src = Z; b.src_n = n / 2;
    b.lo = (S > min_size) ? S / 4 : 0;
    b.scale = 1.0f / (sqrtf(static_cast<float>(n)) * sqrtf(static_cast<float>(S)));
    b.store = kStoreConj; b.dst = bands_out[bi]; b.w0 = w0; b.w1 = w1;
    b.tw = tw; b.tw_h = tw_h;
    rc = run_xform(b, launch);
    if (rc != 0) return rc;
  }
  return 0;
}

// accum(packed spectrum (batch rows of S/2), acc, S, D, lo, scale, first) -> 0 on success
template <class Launch, class Accum>
int recompose_packed(const float* const* bands, const int* sizes, int nbands, int batch, int D,
                     float* out, float2* acc, float2* w0, float2* w1, Launch&& launch,
                     Accum&& accum, const float2* tw = nullptr, int tw_h = 0) {
  int smin = sizes[0];
  for (int i = 1; i < nbands; ++i) smin = sizes[i] < smin ? sizes[i] : smin;
  for (int i = 0; i < nbands; ++i) {
    const int S = sizes[i];
    const int count = plan_launch_count(S / 2, launcher_fuses(launch, 0));
    Xform f;
    f.n = S / 2; f.batch = batch; f.load = kLoadComplex; f.src = bands[i]; f.src_n = S / 2;
    f.lo = 0; f.scale = 1.f; f.store = kStoreComplex;
    f.dst = ((count - 1) & 1) ? w1 : w0;
    f.w0 = w0; f.w1 = w1;
    int rc = run_xform(f, launch);
    if (rc != 0) return rc;
    const int lo = (S == smin) ? 0 : (S / 2 + 1) / 2;
    const float scale = 1.0f / (sqrtf(static_cast<float>(S)) * sqrtf(static_cast<float>(D)));
    rc = accum(static_cast<const float2*>(f.dst), acc, S, D, lo, scale, i == 0 ? 1 : 0);
    if (rc != 0) return rc;
  }
  Xform inv;
  inv.n = D / 2; inv.batch = batch; inv.load = kLoadHalfPk; inv.src = acc; inv.src_n = D / 2 + 1;
  inv.lo = 0; inv.scale = 1.f; inv.store = kStoreConj; inv.dst = out; inv.w0 = w0; inv.w1 = w1;
  inv.tw = tw; inv.tw_h = tw_h;
  return run_xform(inv, launch);
}

// merge without accumulation passes: every band's packed forward transform is kept (in
// `spectra`, batch * sum(S_b / 2) float2 <= batch * D when the sizes are distinct powers of
// two) and the inverse transform gathers its input from all of them while loading.
// Returns -2 when the table or the buffer cannot hold the bands (caller falls back).
template <class Launch>
int recompose_merged(const float* const* bands, const int* sizes, int nbands, int batch, int D,
                     float* out, float2* spectra, float2* w0, float2* w1, Launch&& launch,
                     const float2* tw = nullptr, int tw_h = 0) {
  if (nbands > kMaxBands) return -2;
  long long total = 0;
  int smin = sizes[0];
  for (int i = 0; i < nbands; ++i) {
    total += sizes[i] / 2;
    smin = sizes[i] < smin ? sizes[i] : smin;
  }
  if (total > D) return -2;
  BandTable t;
  t.count = nbands;
  for (int l = 0; l < 32; ++l) t.by_log[l] = -1;
  t.smallest = 0;
  for (int i = 0; i < nbands; ++i) {
    const int l = ilog2(sizes[i] / 2);
    if (sizes[i] < 4 || t.by_log[l] >= 0) return -2;     // the lookup needs distinct sizes
    t.by_log[l] = i;
    if (sizes[i] < sizes[t.smallest]) t.smallest = i;
  }
  long long off = 0;
  for (int i = 0; i < nbands; ++i) {
    const int S = sizes[i];
    Xform f;
    f.n = S / 2; f.batch = batch; f.load = kLoadComplex; f.src = bands[i]; f.src_n = S / 2;
    f.lo = 0; f.scale = 1.f; f.store = kStoreComplex; f.dst = spectra + off;
    f.w0 = w0; f.w1 = w1;
    const int rc = run_xform(f, launch);
    if (rc != 0) return rc;
    t.size[i] = S;
    t.lo[i] = (S == smin) ? 0 : (S / 2 + 1) / 2;
    t.scale[i] = 1.0f / (sqrtf(static_cast<float>(S)) * sqrtf(static_cast<float>(D)));
    t.offset[i] = off;
    off += static_cast<long long>(batch) * (S / 2);
  }
  Xform inv;
  inv.n = D / 2; inv.batch = batch; inv.load = kLoadMergePk; inv.src = spectra; inv.src_n = 0;
  inv.lo = 0; inv.scale = 1.f; inv.store = kStoreConj; inv.dst = out; inv.w0 = w0; inv.w1 = w1;
  inv.bands = &t;
  inv.tw = tw; inv.tw_h = tw_h;
  return run_xform(inv, launch);
}

// one element of the packed band accumulation: bin k of the band = real_bin of its packed
// transform (gid over batch * (D/2+1))
A2M_HD void accumulate_one_packed(const float2* zs, float2* acc, int S, int D, int lo,
                                  float scale, int first, size_t gid) {
  const int k = static_cast<int>(gid % (D / 2 + 1));
  const size_t b = gid / (D / 2 + 1);
  float2 v = first ? make_float2(0.f, 0.f) : acc[gid];
  if (k >= lo && k <= S / 2) {
    const float2 c = real_bin(zs + b * (S / 2), S / 2, k);
    v.x += c.x * scale;
    v.y += c.y * scale;
  }
  acc[gid] = v;
}

// one element of the band accumulation (gid over batch * (D/2+1))
A2M_HD void accumulate_one(const float2* zs, float2* acc, int S, int D, int lo, float scale,
                           int first, size_t gid) {
  const int k = static_cast<int>(gid % (D / 2 + 1));
  const size_t b = gid / (D / 2 + 1);
  float2 v = first ? make_float2(0.f, 0.f) : acc[gid];
  if (k >= lo && k <= S / 2) {
    const float2 c = zs[b * S + k];
    v.x += c.x * scale;
    v.y += c.y * scale;
  }
  acc[gid] = v;
}

}  // namespace fftb
}  // namespace msb
