// Audio2Mel in one fused kernel: right zero-pad -> framing -> window -> real FFT ->
// magnitude -> mel projection -> log10(clamp(., 1e-5)).
//   replaces Audio2Mel.forward, featuresynth/feature/feature.py:39-59.
//
// fp32 throughout (the 1e-3 log-mel bar is a max-abs bar: bins with little energy make
// 16-bit DFT operands unsafe).  One CTA handles 8 consecutive frames of
// one clip: the 8 real frames are packed pairwise into 4 complex radix-2 FFTs in shared
// memory (two real transforms for the price of one), magnitudes stay in shared memory,
// and the mel basis is streamed once per CTA in coalesced 128x32 tiles and reused for
// all 8 frames.  HBM traffic = audio in + log-mel out (+ the L2-resident basis).
#include <math_constants.h>

#include <cstdlib>

#include "a2m_fft.cuh"
#include "runtime.cuh"

namespace msb {

constexpr int kFramesPerCta = 8;
constexpr int kA2MThreads = 256;

struct A2MParams {
  const float* audio;   // (B, 1, N)
  const float* window;  // (n_fft)
  const float* basis;   // (n_mels, bins)
  float* out;           // (B, n_mels, F)
  const int2* ranges;   // per mel row [first, last+1) non-zero bin, or null (dense basis)
  int N, n_fft, log2n, hop, n_mels, bins, F, groups;
};

// Shared tail of both kernels: untangle the two real spectra packed in each complex transform
// (transform j: re plane zr + j * zstride, im plane zi + j * zstride, natural order), take
// magnitudes, project onto the mel basis, log10(clamp).
// One-instruction square root / logarithm (MUFU): relative error ~1e-7 on the magnitude, absolute
// ~1e-7 on the log -- four orders of magnitude inside the 1e-3 log-mel bar; the IEEE forms cost
// ~8 and ~15 instructions per call in a kernel that is instruction-issue bound.
__device__ __forceinline__ float sqrt_fast(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float log10_fast(float x) {     // x >= 1e-5 here
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y * 0.30102999566398120f;
}

// BINS > 0: compile-time bin count (the index split of the un-tangling loop becomes a multiply)
template <int BINS = 0>
__device__ __forceinline__ void a2m_tail(const A2MParams& p, const float* zr, const float* zi,
                                         const int zstride, float* mag, float* tile, const int n,
                                         const int b, const int f0) {
  const int tid = threadIdx.x;
  const int bins = BINS > 0 ? BINS : p.bins;
  // magnitudes as mag4[gh][k] = frames 4 gh .. 4 gh + 3 of bin k: the projection reads one
  // 16-byte vector per bin (quarter-warp wavefronts: the rows' different start bins no longer
  // collide on banks) instead of four scalars
  float4* mag4 = reinterpret_cast<float4*>(mag);
  // untangle the two real spectra of each complex transform, take magnitudes
  for (int i = tid; i < 4 * bins; i += kA2MThreads) {
    const int j = i / bins;
    const int k = i - j * bins;
    const int kn = (n - k) & (n - 1);
    const float ar = zr[j * zstride + k], ai = zi[j * zstride + k];
    const float br = zr[j * zstride + kn], bi = zi[j * zstride + kn];
    const float xar = 0.5f * (ar + br), xai = 0.5f * (ai - bi);
    const float xbr = 0.5f * (ai + bi), xbi = -0.5f * (ar - br);
    // frames 2j, 2j+1 = slots 2 (j & 1), 2 (j & 1) + 1 of group j >> 1
    reinterpret_cast<float2*>(mag4 + (j >> 1) * bins + k)[j & 1] =
        make_float2(sqrt_fast(xar * xar + xai * xai), sqrt_fast(xbr * xbr + xbi * xbi));
  }
  __syncthreads();
  // mel projection: thread (m, gh) accumulates frames gh*4 .. gh*4+3 of mel row m
  const int ml = tid & 127;
  const int gh = tid >> 7;
  if (p.ranges != nullptr) {
    // banded basis (triangular mel filters): walk only the row's non-zero bins, in the same
    // ascending order as the dense product -> bit-identical sums
    for (int mb = 0; mb < p.n_mels; mb += 128) {
      const int m = mb + ml;
      if (m >= p.n_mels) continue;
      const int2 r = __ldg(p.ranges + m);
      const float* brow = p.basis + static_cast<size_t>(m) * bins;
      const float4* mg = mag4 + gh * bins;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k = r.x; k < r.y; ++k) {
        const float w = __ldg(brow + k);
        const float4 v = mg[k];
        acc[0] = fmaf(w, v.x, acc[0]);
        acc[1] = fmaf(w, v.y, acc[1]);
        acc[2] = fmaf(w, v.z, acc[2]);
        acc[3] = fmaf(w, v.w, acc[3]);
      }
      float* o = p.out + (static_cast<size_t>(b) * p.n_mels + m) * p.F;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int f = f0 + gh * 4 + g;
        if (f < p.F) o[f] = log10_fast(fmaxf(acc[g], 1e-5f));
      }
    }
    return;
  }
  for (int mb = 0; mb < p.n_mels; mb += 128) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < bins; k0 += 32) {
      // stage basis[mb .. mb+128)[k0 .. k0+32) : coalesced 128-byte row segments
      for (int i = tid; i < 128 * 32; i += kA2MThreads) {
        const int r = i >> 5, kk = i & 31;
        const int m = mb + r, k = k0 + kk;
        tile[r * 33 + kk] =
            (m < p.n_mels && k < bins) ? __ldg(p.basis + static_cast<size_t>(m) * bins + k) : 0.f;
      }
      __syncthreads();
      const int kmax = (bins - k0) < 32 ? (bins - k0) : 32;
      const float4* mg = mag4 + gh * bins + k0;
      for (int kk = 0; kk < kmax; ++kk) {
        const float w = tile[ml * 33 + kk];
        const float4 v = mg[kk];
        acc[0] = fmaf(w, v.x, acc[0]);
        acc[1] = fmaf(w, v.y, acc[1]);
        acc[2] = fmaf(w, v.z, acc[2]);
        acc[3] = fmaf(w, v.w, acc[3]);
      }
      __syncthreads();
    }
    const int m = mb + ml;
    if (m < p.n_mels) {
      float* o = p.out + (static_cast<size_t>(b) * p.n_mels + m) * p.F;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int f = f0 + gh * 4 + g;
        if (f < p.F) o[f] = log10_fast(fmaxf(acc[g], 1e-5f));
      }
    }
  }
}

// LOG2N > 0: compile-time FFT size (index arithmetic becomes shifts / masks); 0: runtime size
template <int LOG2N>
__global__ void __launch_bounds__(kA2MThreads, 4)
audio2mel_kernel(const A2MParams p) {
  extern __shared__ float sm[];
  const int n = LOG2N > 0 ? (1 << LOG2N) : p.n_fft;
  const int log2n = LOG2N > 0 ? LOG2N : p.log2n;
  const int bins = p.bins;
  const int magld = bins + 3;           // row stride of the magnitude array
  // per-stage CONTIGUOUS twiddle tables: stage with butterfly span `half` uses entries
  // [half-1, 2*half-1) = exp(-2*pi*i*pos/(2*half)) -- conflict-free (a single strided table
  // would be read with a power-of-two stride: up to 32-way bank conflicts)
  // twiddle tables: [n] each for the radix-2 form, [3n/4] for the radix-4 form of n = 1024
  // (indices k1, 2k1, 3k1 < 3n/4) -- 55.4 KB per CTA, four CTAs per SM
  const int twn = LOG2N == 10 ? 3 * n / 4 : n;
  float* tw_c = sm;                     // [twn]
  float* tw_s = tw_c + twn;             // [twn]
  float* mag = tw_s + twn;              // [8][magld]
  float* zr = mag + kFramesPerCta * magld;   // [4][n]   (re)  -- reused as basis tile
  float* zi = zr + 4 * n;                    // [4][n]   (im)
  float* tile = zr;                          // [128][33]

  const int tid = threadIdx.x;
  const int b = blockIdx.x / p.groups;
  const int f0 = (blockIdx.x % p.groups) * kFramesPerCta;
  const float* a = p.audio + static_cast<size_t>(b) * p.N;

  if (LOG2N == 10) {
    // ---- n = 1024 = 4^5: radix-4 decimation in time, 5 passes (half the shared-memory traffic
    // and barriers of the radix-2 form).  One twiddle table tw[k] = exp(-2*pi*i*k/n), k < n.
    for (int i = tid; i < twn; i += kA2MThreads) {
      float sn, cs;
      sincospif(-2.f * static_cast<float>(i) / static_cast<float>(n), &sn, &cs);
      tw_c[i] = cs;
      tw_s[i] = sn;
    }
    // load 8 windowed frames in base-4 digit-reversed order; frame 2j -> re, 2j+1 -> im
    for (int i = tid; i < 4 * n; i += kA2MThreads) {
      const int j = i >> 10;
      const int t = i & 1023;
      const int rev = ((t & 3) << 8) | (((t >> 2) & 3) << 6) | (((t >> 4) & 3) << 4) |
                      (((t >> 6) & 3) << 2) | ((t >> 8) & 3);
      const float w = __ldg(p.window + t);
      const int fa = f0 + 2 * j, fb = fa + 1;
      const long long sa = static_cast<long long>(fa) * p.hop + t;
      const long long sb = static_cast<long long>(fb) * p.hop + t;
      const float va = (fa < p.F && sa < p.N) ? __ldg(a + sa) : 0.f;
      const float vb = (fb < p.F && sb < p.N) ? __ldg(a + sb) : 0.f;
      zr[j * n + rev] = va * w;
      zi[j * n + rev] = vb * w;
    }
    __syncthreads();
#pragma unroll 1
    for (int st = 0; st < 5; ++st) {
      const int q = 1 << (2 * st);                 // butterfly span
      const int tstep = 256 >> (2 * st);           // n / (4q): twiddle index step
      for (int i = tid; i < n; i += kA2MThreads) {  // 4 transforms x n/4 butterflies
        const int j = i >> 8;
        const int bf = i & 255;
        const int pos = bf & (q - 1);
        const int i0 = ((bf >> (2 * st)) << (2 * st + 2)) + pos + j * n;
        const int i1 = i0 + q, i2 = i1 + q, i3 = i2 + q;
        const int k1 = pos * tstep;
        const float c1 = tw_c[k1], s1 = tw_s[k1];
        const float c2 = tw_c[2 * k1], s2 = tw_s[2 * k1];
        const float c3 = tw_c[3 * k1], s3 = tw_s[3 * k1];
        const float ar = zr[i0], ai = zi[i0];
        float xr = zr[i1], xi = zi[i1];
        const float br = xr * c1 - xi * s1, bi = xr * s1 + xi * c1;
        xr = zr[i2]; xi = zi[i2];
        const float cr = xr * c2 - xi * s2, ci = xr * s2 + xi * c2;
        xr = zr[i3]; xi = zi[i3];
        const float dr = xr * c3 - xi * s3, di = xr * s3 + xi * c3;
        const float t0r = ar + cr, t0i = ai + ci, t1r = ar - cr, t1i = ai - ci;
        const float t2r = br + dr, t2i = bi + di, t3r = br - dr, t3i = bi - di;
        zr[i0] = t0r + t2r; zi[i0] = t0i + t2i;
        zr[i2] = t0r - t2r; zi[i2] = t0i - t2i;
        zr[i1] = t1r + t3i; zi[i1] = t1i - t3r;     // (a - c) - j (b - d)
        zr[i3] = t1r - t3i; zi[i3] = t1i + t3r;     // (a - c) + j (b - d)
      }
      __syncthreads();
    }
  } else {
    for (int i = tid; i < n - 1; i += kA2MThreads) {
      const int half = 1 << (31 - __clz(i + 1));      // largest power of two <= i+1
      const int pos = i + 1 - half;
      float s, c;
      sincospif(-static_cast<float>(pos) / static_cast<float>(half), &s, &c);
      tw_c[i] = c;
      tw_s[i] = s;
    }
    // load 8 windowed frames, bit-reversed, frame 2j -> real part, 2j+1 -> imaginary part
    for (int i = tid; i < 4 * n; i += kA2MThreads) {
      const int j = i / n;
      const int t = i - j * n;
      const int rev = static_cast<int>(__brev(static_cast<unsigned>(t)) >> (32 - log2n));
      const float w = __ldg(p.window + t);
      const int fa = f0 + 2 * j, fb = fa + 1;
      const long long sa = static_cast<long long>(fa) * p.hop + t;
      const long long sb = static_cast<long long>(fb) * p.hop + t;
      const float va = (fa < p.F && sa < p.N) ? __ldg(a + sa) : 0.f;
      const float vb = (fb < p.F && sb < p.N) ? __ldg(a + sb) : 0.f;
      zr[j * n + rev] = va * w;
      zi[j * n + rev] = vb * w;
    }
    __syncthreads();
    // 4 in-place radix-2 DIT FFTs side by side
    for (int s = 0; s < log2n; ++s) {
      const int half = 1 << s;
      const float* twc = tw_c + half - 1;
      const float* tws_ = tw_s + half - 1;
      for (int i = tid; i < 2 * n; i += kA2MThreads) {   // 4 * n/2 butterflies
        const int j = i / (n / 2);
        const int bf = i - j * (n / 2);
        const int pos = bf & (half - 1);
        const int i0 = ((bf >> s) << (s + 1)) + pos + j * n;
        const int i1 = i0 + half;
        const float c = twc[pos], sn = tws_[pos];
        const float xr = zr[i1], xi = zi[i1];
        const float tr = xr * c - xi * sn;
        const float ti = xr * sn + xi * c;
        const float ur = zr[i0], ui = zi[i0];
        zr[i0] = ur + tr; zi[i0] = ui + ti;
        zr[i1] = ur - tr; zi[i1] = ui - ti;
      }
      __syncthreads();
    }
  }
  a2m_tail(p, zr, zi, n, mag, tile, n, b, f0);
}

// n_fft = 1024, register-resident form (a2m_fft.cuh): 64 threads per complex transform, the
// 16 samples of a thread come straight from global memory (coalesced 256-byte runs), two
// exchanges through a conflict-free padded array, 5 CTA barriers in all.  Shared memory:
// 8 planes of 1088 floats + the magnitudes = 51.3 KB -> four CTAs per SM.
__global__ void __launch_bounds__(kA2MThreads, 4)
audio2mel_r16_kernel(const A2MParams p) {
  extern __shared__ float sm[];
  float* zr = sm;                            // [4][kPlane] (re)  -- reused as basis tile
  float* zi = sm + 4 * a2m::kPlane;          // [4][kPlane] (im)
  float* mag = sm + 8 * a2m::kPlane;         // [8][bins + 3]
  const int tid = threadIdx.x;
  const int j = tid >> 6, t = tid & 63;
  const int b = blockIdx.x / p.groups;
  const int f0 = (blockIdx.x % p.groups) * kFramesPerCta;
  const float* a = p.audio + static_cast<size_t>(b) * p.N;
  float* sr = zr + j * a2m::kPlane;
  float* si = zi + j * a2m::kPlane;
  {
    // frame 2j -> real part, frame 2j+1 -> imaginary part; right zero padding past the clip
    float re[16], im[16];
    const int fa = f0 + 2 * j;
    const long long sa = static_cast<long long>(fa) * p.hop + t;
    const long long sb = sa + p.hop;
    const bool oka = fa < p.F, okb = fa + 1 < p.F;
    // samples left in the clip from each frame's first sample (<= 0: the whole frame is padding)
    const long long la = oka ? p.N - sa : 0, lb = okb ? p.N - sb : 0;
    const int lefta = la > 1024 ? 1024 : static_cast<int>(la > 0 ? la : 0);
    const int leftb = lb > 1024 ? 1024 : static_cast<int>(lb > 0 ? lb : 0);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const float w = __ldg(p.window + 64 * q + t);
      const float va = (64 * q < lefta) ? __ldg(a + sa + 64 * q) : 0.f;
      const float vb = (64 * q < leftb) ? __ldg(a + sb + 64 * q) : 0.f;
      re[q] = va * w;
      im[q] = vb * w;
    }
    a2m::pass_a(t, re, im, sr, si);
  }
  __syncthreads();
  a2m::pass_b(t, sr, si);
  __syncthreads();
  {
    float re[16], im[16];
    a2m::pass_c_read(t, sr, si, re, im);
    __syncthreads();
    a2m::pass_c_write(t, re, im, sr, si);
  }
  __syncthreads();
  a2m_tail<a2m::kN / 2 + 1>(p, zr, zi, a2m::kPlane, mag, zr, a2m::kN, b, f0);
}

}  // namespace msb

using namespace msb;

extern "C" {

int ms_audio2mel_frames(int samples, int n_fft, int hop) {
  if (samples <= 0 || n_fft <= 0 || hop <= 0 || hop > n_fft) return MS_ERR_INVALID;
  const int padded = samples + (n_fft - hop) / 2;
  if (padded < n_fft) return 0;
  return (padded - n_fft) / hop + 1;
}

ms_status ms_audio2mel_fwd(const float* audio, const float* window, const float* mel_basis,
                           const int* row_ranges, float* out, int batch, int samples, int n_fft,
                           int hop, int n_mels, void* stream) {
  if (audio == nullptr || window == nullptr || mel_basis == nullptr || out == nullptr ||
      batch <= 0 || n_mels <= 0)
    return MS_ERR_INVALID;
  int log2n = 0;
  while ((1 << log2n) < n_fft) ++log2n;
  if ((1 << log2n) != n_fft || n_fft < 64 || n_fft > 2048) return MS_ERR_INVALID;
  const int F = ms_audio2mel_frames(samples, n_fft, hop);
  if (F < 0) return MS_ERR_INVALID;
  if (F == 0) return MS_OK;
  A2MParams p;
  p.audio = audio; p.window = window; p.basis = mel_basis; p.out = out;
  p.ranges = reinterpret_cast<const int2*>(row_ranges);
  p.N = samples; p.n_fft = n_fft; p.log2n = log2n; p.hop = hop; p.n_mels = n_mels;
  p.bins = n_fft / 2 + 1; p.F = F; p.groups = (F + kFramesPerCta - 1) / kFramesPerCta;
  const long long blocks = static_cast<long long>(batch) * p.groups;
  if (blocks > 0x7fffffffLL) return MS_ERR_INVALID;
  const size_t tile_floats = 128 * 33;
  // MSB_A2M_RADIX4=1: the shared-memory radix-4 form of n_fft = 1024 (kept for A/B timing;
  // read per call so one process can time both)
  const char* env_r4 = getenv("MSB_A2M_RADIX4");
  const bool radix4 = env_r4 != nullptr && env_r4[0] == '1';
  if (n_fft == 1024 && !radix4) {
    const size_t smem = sizeof(float) * (8 * a2m::kPlane + kFramesPerCta * (p.bins + 3));
    static_assert(8 * a2m::kPlane >= 128 * 33, "basis tile must fit in the FFT planes");
    static thread_local bool attr_r16 = false;
    if (!attr_r16) {
      cudaError_t e = cudaFuncSetAttribute(audio2mel_r16_kernel,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem));
      if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(audio2mel_r16_kernel)");
      attr_r16 = true;
    }
    audio2mel_r16_kernel<<<static_cast<unsigned>(blocks), kA2MThreads, smem,
                           static_cast<cudaStream_t>(stream)>>>(p);
    return after_launch("audio2mel_r16_kernel");
  }
  const size_t fft_floats = 8 * static_cast<size_t>(n_fft);
  const size_t tw_floats = n_fft == 1024 ? 2 * (3 * n_fft / 4) : 2 * n_fft;
  const size_t smem = sizeof(float) * (tw_floats + kFramesPerCta * (p.bins + 3) +
                                       (fft_floats > tile_floats ? fft_floats : tile_floats));
  static thread_local size_t attr_set = 0;
  if (smem > attr_set) {
    cudaError_t e = cudaFuncSetAttribute(audio2mel_kernel<10>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(audio2mel_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem));
    if (e != cudaSuccess) return check_cuda(e, "cudaFuncSetAttribute(audio2mel_kernel)");
    attr_set = smem;
  }
  if (n_fft == 1024)
    audio2mel_kernel<10><<<static_cast<unsigned>(blocks), kA2MThreads, smem,
                           static_cast<cudaStream_t>(stream)>>>(p);
  else
    audio2mel_kernel<0><<<static_cast<unsigned>(blocks), kA2MThreads, smem,
                          static_cast<cudaStream_t>(stream)>>>(p);
  return after_launch("audio2mel_kernel");
}

}  // extern "C"
