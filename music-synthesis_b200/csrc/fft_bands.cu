// FFT octave-band split / merge ("fixed multiscale filterbank").
//   replaces fft_frequency_decompose / fft_resample / fft_frequency_recompose,
//   featuresynth/audio/transform.py:50-115 (ortho-normalised rFFT of the whole clip, band of
//   size S keeps bins [S/4, S/2] -- the lowest keeps [0, S/2] -- and is inverse-transformed
//   at length S; recompose zero-stuffs each band's spectrum to the full length and sums,
//   including the reference's double-counted boundary bins).
// HBM/L2-bound, fp32.  Transforms are power-of-two Stockham autosort FFTs, out of place between
// two workspace buffers.  Default form (fft_passes.cuh): register-resident radix-16 passes
// (+ one radix-2 / radix-4 pass for the odd bits of log2 n), real-input load, band cut-out,
// Hermitian expansion and real-part store folded into the first / last pass of a transform:
// n = 65536 is 4 passes and no boundary kernels instead of 8 radix-4 passes + 3 copies.
// Two passes per launch (default; MSB_FFT_FUSE=0: one): the first two passes of a transform and
// every following pair of radix-16 passes exchange through shared memory inside one kernel, so
// a half-length transform of 32768 points reads and writes HBM twice instead of four times.
// Real-input packing: a real sequence of length 2h runs as the h-point complex transform of
// x[2m] + i x[2m+1]; the band cut-out builds the packed inverse input straight from the packed
// forward output (MSB_FFT_PACKED=0: full-length complex transforms).
// MSB_FFT_LEGACY=1 selects the first version (radix-4 passes, separate boundary kernels).
// Because irFFT is linear, recompose sums the bands' spectra first and runs ONE inverse
// transform.
#include <cstdint>
#include <cstdlib>

#include "fft_passes.cuh"
#include "runtime.cuh"

namespace msb {

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 twiddle(float frac_pi, float sign) {  // exp(sign*i*pi*frac)
  float s, c;
  sincospif(frac_pi, &s, &c);
  return make_float2(c, sign * s);
}

// One Stockham radix-2 pass: p = 1, 2, 4, ... n/2.  sign = -1 forward, +1 inverse.
__global__ void fft_pass2_kernel(const float2* __restrict__ x, float2* __restrict__ y, int n, int p,
                                 float sign, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = n >> 1;
  const int i = static_cast<int>(gid % t);
  const size_t base = (gid / t) * n;
  const int k = i & (p - 1);
  const int j = ((i - k) << 1) + k;
  const float2 u0 = x[base + i];
  const float2 u1 = cmul(x[base + i + t], twiddle(static_cast<float>(k) / static_cast<float>(p), sign));
  y[base + j] = make_float2(u0.x + u1.x, u0.y + u1.y);
  y[base + j + p] = make_float2(u0.x - u1.x, u0.y - u1.y);
}

// One Stockham radix-4 pass: p = 1, 4, 16, ...
__global__ void fft_pass4_kernel(const float2* __restrict__ x, float2* __restrict__ y, int n, int p,
                                 float sign, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int t = n >> 2;
  const int i = static_cast<int>(gid % t);
  const size_t base = (gid / t) * n;
  const int k = i & (p - 1);
  const int j = ((i - k) << 2) + k;
  const float a = static_cast<float>(k) / static_cast<float>(2 * p);   // alpha / pi
  float2 v0 = x[base + i];
  float2 v1 = cmul(x[base + i + t], twiddle(a, sign));
  float2 v2 = cmul(x[base + i + 2 * t], twiddle(2.f * a, sign));
  float2 v3 = cmul(x[base + i + 3 * t], twiddle(3.f * a, sign));
  // radix-4 butterfly
  const float2 s02 = make_float2(v0.x + v2.x, v0.y + v2.y), d02 = make_float2(v0.x - v2.x, v0.y - v2.y);
  const float2 s13 = make_float2(v1.x + v3.x, v1.y + v3.y), d13 = make_float2(v1.x - v3.x, v1.y - v3.y);
  // multiply d13 by (sign * i):  forward (sign=-1): -i*d13 = (d13.y, -d13.x)
  const float2 r13 = make_float2(-sign * d13.y, sign * d13.x);
  y[base + j] = make_float2(s02.x + s13.x, s02.y + s13.y);
  y[base + j + p] = make_float2(d02.x + r13.x, d02.y + r13.y);
  y[base + j + 2 * p] = make_float2(s02.x - s13.x, s02.y - s13.y);
  y[base + j + 3 * p] = make_float2(d02.x - r13.x, d02.y - r13.y);
}

// complex FFT of `batch` rows of length n (power of two); data starts in `a`; returns the
// buffer holding the result (a or b)
static float2* fft_c2c(float2* a, float2* b, int batch, int n, float sign, cudaStream_t st,
                       ms_status* status) {
  int log2n = 0;
  while ((1 << log2n) < n) ++log2n;
  float2* src = a;
  float2* dst = b;
  int p = 1;
  int stages = log2n;
  const int threads = 256;
  if (stages & 1) {
    const size_t total = static_cast<size_t>(batch) * (n >> 1);
    fft_pass2_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, st>>>(
        src, dst, n, p, sign, total);
    *status = after_launch("fft_pass2_kernel");
    if (*status != MS_OK) return nullptr;
    p <<= 1;
    float2* t = src; src = dst; dst = t;
    --stages;
  }
  for (; stages > 0; stages -= 2) {
    const size_t total = static_cast<size_t>(batch) * (n >> 2);
    fft_pass4_kernel<<<static_cast<unsigned>((total + threads - 1) / threads), threads, 0, st>>>(
        src, dst, n, p, sign, total);
    *status = after_launch("fft_pass4_kernel");
    if (*status != MS_OK) return nullptr;
    p <<= 2;
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

__global__ void real_to_complex_kernel(const float* __restrict__ x, float2* __restrict__ z,
                                       size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < total) z[i] = make_float2(__ldg(x + i), 0.f);
}

// Hermitian spectrum of one band: Z[k] = scale * C[k] for lo <= k <= S/2 (imag of k = 0 and
// k = S/2 dropped, as a c2r transform ignores them), Z[S-k] = conj(Z[k]), 0 elsewhere.
__global__ void band_spectrum_kernel(const float2* __restrict__ coeffs, float2* __restrict__ z,
                                     int n, int S, int lo, float scale, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int k = static_cast<int>(gid % S);
  const size_t b = gid / S;
  const int kk = k <= S / 2 ? k : S - k;
  float2 v = make_float2(0.f, 0.f);
  if (kk >= lo) {
    v = coeffs[b * n + kk];
    v.x *= scale;
    v.y *= (kk == 0 || kk == S / 2) ? 0.f : (k <= S / 2 ? scale : -scale);
  }
  z[gid] = v;
}

__global__ void complex_real_part_kernel(const float2* __restrict__ z, float* __restrict__ y,
                                         size_t total) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i < total) y[i] = z[i].x;
}

// accumulate a band's spectrum (size S transform in `zs`) into the full-size Hermitian
// spectrum `acc` (length D): bins [lo, S/2] of the band land on the same bin indices.
__global__ void accumulate_band_kernel(const float2* __restrict__ zs, float2* __restrict__ acc,
                                       int S, int D, int lo, float scale, int first, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int k = static_cast<int>(gid % (D / 2 + 1));
  const size_t b = gid / (D / 2 + 1);
  float2 v = first ? make_float2(0.f, 0.f) : acc[b * (D / 2 + 1) + k];
  if (k >= lo && k <= S / 2) {
    const float2 c = zs[b * S + k];
    v.x += c.x * scale;
    v.y += c.y * scale;
  }
  acc[b * (D / 2 + 1) + k] = v;
}

// full Hermitian spectrum of length D from its D/2+1 half (imag of DC / Nyquist dropped)
__global__ void hermitian_expand_kernel(const float2* __restrict__ half, float2* __restrict__ z,
                                        int D, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid >= total) return;
  const int k = static_cast<int>(gid % D);
  const size_t b = gid / D;
  const int kk = k <= D / 2 ? k : D - k;
  float2 v = half[b * (D / 2 + 1) + kk];
  if (kk == 0 || kk == D / 2) v.y = 0.f;
  else if (k > D / 2) v.y = -v.y;
  z[gid] = v;
}

// ---- register-resident passes (fft_passes.cuh) ------------------------------------------
template <int R, int LD, int ST>
__global__ void __launch_bounds__(256) fft_pass_kernel(const fftb::PassArgs a) {
  const size_t gid = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (gid < a.total) fftb::pass_thread<R, LD, ST>(a, gid);
}

// First radix-16 pass of a transform (p = 1): a thread's 16 outputs are consecutive and a
// warp's 512 outputs are ONE contiguous 4 KB run, so they are exchanged through shared memory
// and stored as full 256-byte rows per instruction (direct stores would hit 32 different
// 128-byte lines per instruction, a quarter of a sector each).  Requires (n / 16) % 32 == 0
// and complex output.  Row pitch 17 float2: both the per-thread writes and the transposed
// reads are bank-conflict free.
template <int LD>
__global__ void __launch_bounds__(256) fft_pass16_first_kernel(const fftb::PassArgs a) {
  __shared__ float2 stage[8][32][17];
  const size_t gid = blockIdx.x * static_cast<size_t>(256) + threadIdx.x;
  if (gid >= a.total) return;      // total % 32 == 0: whole warps leave together
  float re[16], im[16];
  size_t row;
  int j;
  fftb::pass_compute<16, LD>(a, gid, re, im, row, j);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int m = 0; m < 16; ++m) stage[warp][lane][m] = make_float2(re[m], im[m]);
  __syncwarp();
  float2* y = static_cast<float2*>(a.y) + row * a.n + (j - 16 * lane);
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    const int idx = r * 32 + lane;
    y[idx] = stage[warp][idx >> 4][idx & 15];
  }
}

// Two passes per launch (fft_passes.cuh, "two passes in one launch"): 16 groups of G = 16 R1
// points per CTA, thread (hi, gl) = (tid / 16, tid % 16) works on group gl, so that for every
// load / store instruction 16 neighbouring lanes touch one contiguous 128-byte run (groups are
// adjacent in g, and in k1 for the stores when p >= 16).  STAGED (p = 1): a group's outputs are
// G consecutive elements, the CTA's 16 G outputs one contiguous run -- they go back through the
// exchange buffer and are stored in order.
template <int R1, bool STAGED, int LD>
__global__ void __launch_bounds__(256) fft_pass2x_kernel(const fftb::PassArgs a, const int store) {
  constexpr int G = 16 * R1;
  constexpr int GP = fftb::kExPitch * R1 + 1;     // float2 per group: word pitch = 2 mod 4
  __shared__ float2 ex[16 * GP];
  const int tid = threadIdx.x, gl = tid & 15, hi = tid >> 4;
  const size_t group = blockIdx.x * static_cast<size_t>(16) + gl;
  const bool valid = group < a.total;
  if (valid) fftb::fused_first_half<R1, LD>(a, group, hi, ex + gl * GP);
  __syncthreads();
  const bool act = valid && hi < R1;
  float re[16], im[16];
  size_t o = 0;
  if (act) fftb::fused_second_half<R1>(a, group, hi, ex + gl * GP, re, im, o);
  if (!STAGED) {
    if (act) {
#pragma unroll
      for (int m = 0; m < 16; ++m)
        fftb::store_any(a.y, store, o + static_cast<size_t>(hi + R1 * m) * a.p, re[m], im[m]);
    }
    return;
  }
  __syncthreads();
  if (act) {
#pragma unroll
    for (int m = 0; m < 16; ++m) ex[gl * GP + hi + R1 * m] = make_float2(re[m], im[m]);
  }
  __syncthreads();
  const size_t first = blockIdx.x * static_cast<size_t>(16);
  const size_t left = a.total - first;
  const int count = (left < 16 ? static_cast<int>(left) : 16) * G;
  for (int i = tid; i < count; i += 256) {
    const float2 v = ex[(i / G) * GP + (i % G)];
    fftb::store_any(a.y, store, first * G + i, v.x, v.y);
  }
}

// special loaders only occur in the first pass of a transform (p = 1: the staged form)
static int launch_pass2x(int r1, int load, int store, const fftb::PassArgs& a, cudaStream_t st) {
  const unsigned grid = static_cast<unsigned>((a.total + 15) / 16);
  if (a.p > 1 && load != fftb::kLoadComplex) return MS_ERR_INVALID;
  return fftb::dispatch_fused(r1, load, [&](auto r, auto ld) -> int {
    constexpr int R1 = decltype(r)::value;
    if (a.p == 1)
      fft_pass2x_kernel<R1, true, decltype(ld)::value><<<grid, 256, 0, st>>>(a, store);
    else
      fft_pass2x_kernel<R1, false, fftb::kLoadComplex><<<grid, 256, 0, st>>>(a, store);
    return after_launch("fft_pass2x_kernel");
  });
}

// tw[j] = exp(-i pi j / h), 0 <= j <= h (PassArgs::tw)
__global__ void fft_twiddle_table_kernel(float2* __restrict__ tw, int h) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > h) return;
  float s, c;
  sincospif(-static_cast<float>(j) / static_cast<float>(h), &s, &c);
  tw[j] = make_float2(c, s);
}

static bool env_flag(const char* name, bool dflt) {
  const char* e = getenv(name);
  if (e == nullptr || e[0] == 0) return dflt;
  return e[0] != '0';
}

// the knobs are read per call (not cached) so one process can time both settings
struct PassLauncher {
  cudaStream_t st;
  bool staged;
  bool fuse;         // MSB_FFT_FUSE=0: one pass per launch
  explicit PassLauncher(cudaStream_t s)
      : st(s), staged(env_flag("MSB_FFT_STAGED", true)), fuse(env_flag("MSB_FFT_FUSE", true)) {}
  bool fuses() const { return fuse; }
  int fused(int r1, int load, int store, const fftb::PassArgs& a) const {
    return launch_pass2x(r1, load, store, a, st);
  }
  int operator()(int radix, int load, int store, const fftb::PassArgs& a) const {
    const unsigned grid = static_cast<unsigned>((a.total + 255) / 256);
    if (staged && radix == 16 && a.p == 1 && store == fftb::kStoreComplex &&
        (a.n / 16) % 32 == 0) {
      switch (load) {
        case fftb::kLoadComplex:
          fft_pass16_first_kernel<fftb::kLoadComplex><<<grid, 256, 0, st>>>(a); break;
        case fftb::kLoadReal:
          fft_pass16_first_kernel<fftb::kLoadReal><<<grid, 256, 0, st>>>(a); break;
        case fftb::kLoadBand:
          fft_pass16_first_kernel<fftb::kLoadBand><<<grid, 256, 0, st>>>(a); break;
        case fftb::kLoadHalf:
          fft_pass16_first_kernel<fftb::kLoadHalf><<<grid, 256, 0, st>>>(a); break;
        case fftb::kLoadBandPk:
          fft_pass16_first_kernel<fftb::kLoadBandPk><<<grid, 256, 0, st>>>(a); break;
        case fftb::kLoadHalfPk:
          fft_pass16_first_kernel<fftb::kLoadHalfPk><<<grid, 256, 0, st>>>(a); break;
        default:
          fft_pass16_first_kernel<fftb::kLoadMergePk><<<grid, 256, 0, st>>>(a); break;
      }
      return after_launch("fft_pass16_first_kernel");
    }
    cudaStream_t s = st;
    return fftb::dispatch(radix, load, store, [&](auto r, auto ld, auto sto) -> int {
      fft_pass_kernel<decltype(r)::value, decltype(ld)::value, decltype(sto)::value>
          <<<grid, 256, 0, s>>>(a);
      return after_launch("fft_pass_kernel");
    });
  }
};

__global__ void accumulate_band2_kernel(const float2* __restrict__ zs, float2* __restrict__ acc,
                                        int S, int D, int lo, float scale, int first,
                                        size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid < total) fftb::accumulate_one(zs, acc, S, D, lo, scale, first, gid);
}

__global__ void accumulate_band_packed_kernel(const float2* __restrict__ zs,
                                              float2* __restrict__ acc, int S, int D, int lo,
                                              float scale, int first, size_t total) {
  const size_t gid = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (gid < total) fftb::accumulate_one_packed(zs, acc, S, D, lo, scale, first, gid);
}

inline ms_status twiddle_table(float2* tw, int h, cudaStream_t st) {
  fft_twiddle_table_kernel<<<(h + 256) / 256, 256, 0, st>>>(tw, h);
  return after_launch("fft_twiddle_table_kernel");
}

inline bool aligned8(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

inline unsigned nblk(size_t total) { return static_cast<unsigned>((total + 255) / 256); }
inline bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

}  // namespace msb

using namespace msb;

extern "C" {

size_t ms_fft_bands_workspace_bytes(int batch, int n) {
  if (batch <= 0 || n <= 0) return 0;
  // coefficient buffer + two ping-pong transform buffers (complex, full length)
  // + the twiddle table of the packed loaders (n / 2 + 1 entries)
  return 3 * static_cast<size_t>(batch) * n * sizeof(float2) +
         (static_cast<size_t>(n) / 2 + 1) * sizeof(float2) + 1024;
}

ms_status ms_fft_frequency_decompose(const float* x, int batch, int n, int min_size,
                                     float* const* bands_out, int nbands, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (x == nullptr || bands_out == nullptr || workspace == nullptr || batch <= 0 ||
      !is_pow2(n) || !is_pow2(min_size) || min_size < 4 || min_size > n)
    return MS_ERR_INVALID;
  int expect = 0;
  for (int s = min_size; s <= n; s <<= 1) ++expect;
  if (nbands != expect) return MS_ERR_INVALID;
  if (workspace_bytes < ms_fft_bands_workspace_bytes(batch, n)) return MS_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t bn = static_cast<size_t>(batch) * n;
  float2* coef = static_cast<float2*>(workspace);
  float2* w0 = coef + bn;
  float2* w1 = w0 + bn;
  const bool legacy = env_flag("MSB_FFT_LEGACY", false);
  if (!legacy) {
    // real-input packing (half-length transforms) needs 8-byte aligned rows: float2 accesses
    bool packed = env_flag("MSB_FFT_PACKED", true) && aligned8(x);
    for (int i = 0; i < nbands; ++i) packed = packed && aligned8(bands_out[i]);
    if (packed) {
      const float2* tw = nullptr;
      if (env_flag("MSB_FFT_TABLE", true)) {
        ms_status ts = twiddle_table(w1 + bn, n / 2, st);
        if (ts != MS_OK) return ts;
        tw = w1 + bn;
      }
      return static_cast<ms_status>(fftb::decompose_packed(x, batch, n, min_size, bands_out, coef,
                                                           w0, w1, PassLauncher(st), tw, n / 2));
    }
    return static_cast<ms_status>(
        fftb::decompose(x, batch, n, min_size, bands_out, coef, w0, w1, PassLauncher(st)));
  }
  real_to_complex_kernel<<<nblk(bn), 256, 0, st>>>(x, w0, bn);
  ms_status s = after_launch("real_to_complex_kernel");
  if (s != MS_OK) return s;
  float2* f = fft_c2c(w0, w1, batch, n, -1.f, st, &s);
  if (f == nullptr) return s;
  s = check_cuda(cudaMemcpyAsync(coef, f, bn * sizeof(float2), cudaMemcpyDeviceToDevice, st),
                 "cudaMemcpyAsync(coeffs)");
  if (s != MS_OK) return s;
  int bi = 0;
  for (int S = min_size; S <= n; S <<= 1, ++bi) {
    const size_t bs = static_cast<size_t>(batch) * S;
    const int lo = (S > min_size) ? S / 4 : 0;
    // ortho forward (1/sqrt(n)) and ortho inverse at length S (1/sqrt(S))
    const float scale = 1.0f / (sqrtf(static_cast<float>(n)) * sqrtf(static_cast<float>(S)));
    band_spectrum_kernel<<<nblk(bs), 256, 0, st>>>(coef, w0, n, S, lo, scale, bs);
    s = after_launch("band_spectrum_kernel");
    if (s != MS_OK) return s;
    float2* r = fft_c2c(w0, w1, batch, S, +1.f, st, &s);
    if (r == nullptr) return s;
    complex_real_part_kernel<<<nblk(bs), 256, 0, st>>>(r, bands_out[bi], bs);
    s = after_launch("complex_real_part_kernel");
    if (s != MS_OK) return s;
  }
  return MS_OK;
}

ms_status ms_fft_frequency_recompose(const float* const* bands, const int* sizes, int nbands,
                                     int batch, int desired_size, float* out, void* workspace,
                                     size_t workspace_bytes, void* stream) {
  if (bands == nullptr || sizes == nullptr || out == nullptr || workspace == nullptr ||
      nbands <= 0 || batch <= 0 || !is_pow2(desired_size))
    return MS_ERR_INVALID;
  int smin = sizes[0];
  for (int i = 0; i < nbands; ++i) {
    if (!is_pow2(sizes[i]) || sizes[i] > desired_size || sizes[i] < 4) return MS_ERR_INVALID;
    if (sizes[i] < smin) smin = sizes[i];
  }
  if (workspace_bytes < ms_fft_bands_workspace_bytes(batch, desired_size)) return MS_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = desired_size;
  const size_t bd = static_cast<size_t>(batch) * D;
  float2* acc = static_cast<float2*>(workspace);      // (batch, D/2+1)
  float2* w0 = acc + bd;
  float2* w1 = w0 + bd;
  const size_t bh = static_cast<size_t>(batch) * (D / 2 + 1);
  ms_status s = MS_OK;
  const bool legacy = env_flag("MSB_FFT_LEGACY", false);
  if (!legacy) {
    auto accum = [&](const float2* zs, float2* ac, int S, int Dd, int lo, float scale,
                     int first) -> int {
      accumulate_band2_kernel<<<nblk(bh), 256, 0, st>>>(zs, ac, S, Dd, lo, scale, first, bh);
      return after_launch("accumulate_band2_kernel");
    };
    bool packed = env_flag("MSB_FFT_PACKED", true) && aligned8(out);
    for (int i = 0; i < nbands; ++i) packed = packed && aligned8(bands[i]);
    const float2* tw = nullptr;
    if (packed && env_flag("MSB_FFT_TABLE", true)) {
      s = twiddle_table(w1 + bd, D / 2, st);
      if (s != MS_OK) return s;
      tw = w1 + bd;
    }
    if (packed && env_flag("MSB_FFT_MERGE_GATHER", true)) {
      // the bands' packed spectra are kept side by side in the first workspace region and the
      // inverse transform gathers from them while loading: no accumulation passes
      const int rc = fftb::recompose_merged(bands, sizes, nbands, batch, D, out, acc, w0, w1,
                                            PassLauncher(st), tw, D / 2);
      if (rc != -2) return static_cast<ms_status>(rc);
    }
    if (packed) {
      auto accum_pk = [&](const float2* zs, float2* ac, int S, int Dd, int lo, float scale,
                          int first) -> int {
        accumulate_band_packed_kernel<<<nblk(bh), 256, 0, st>>>(zs, ac, S, Dd, lo, scale, first,
                                                                bh);
        return after_launch("accumulate_band_packed_kernel");
      };
      return static_cast<ms_status>(fftb::recompose_packed(bands, sizes, nbands, batch, D, out,
                                                           acc, w0, w1, PassLauncher(st),
                                                           accum_pk, tw, D / 2));
    }
    return static_cast<ms_status>(fftb::recompose(bands, sizes, nbands, batch, D, out, acc, w0,
                                                  w1, PassLauncher(st), accum));
  }
  for (int i = 0; i < nbands; ++i) {
    const int S = sizes[i];
    const size_t bs = static_cast<size_t>(batch) * S;
    real_to_complex_kernel<<<nblk(bs), 256, 0, st>>>(bands[i], w0, bs);
    s = after_launch("real_to_complex_kernel");
    if (s != MS_OK) return s;
    float2* f = fft_c2c(w0, w1, batch, S, -1.f, st, &s);
    if (f == nullptr) return s;
    // fft_resample: the lowest band keeps bins [0, S/2], the others [S/4, S/2]
    // (n_coeffs // 2 with n_coeffs = S/2 + 1), audio/transform.py:93-96
    const int lo = (S == smin) ? 0 : (S / 2 + 1) / 2;
    const float scale = 1.0f / (sqrtf(static_cast<float>(S)) * sqrtf(static_cast<float>(D)));
    accumulate_band_kernel<<<nblk(bh), 256, 0, st>>>(f, acc, S, D, lo, scale, i == 0 ? 1 : 0, bh);
    s = after_launch("accumulate_band_kernel");
    if (s != MS_OK) return s;
  }
  hermitian_expand_kernel<<<nblk(bd), 256, 0, st>>>(acc, w0, D, bd);
  s = after_launch("hermitian_expand_kernel");
  if (s != MS_OK) return s;
  float2* r = fft_c2c(w0, w1, batch, D, +1.f, st, &s);
  if (r == nullptr) return s;
  complex_real_part_kernel<<<nblk(bd), 256, 0, st>>>(r, out, bd);
  return after_launch("complex_real_part_kernel");
}

/* ---- exact adjoints (training with decompose=True / recompose=True) ----
 * With orthonormal transforms the band merge is the adjoint of the band split and vice versa,
 * except at ONE bin per band: bin S/2 is the Nyquist bin of the S-point real transform (weight 1
 * in its adjoint) but an interior bin of the n-point one (weight 1/2 resp. 2).  The difference is a
 * rank-1 term per band with S < n (derivation and numerical check: DESIGN.md section 5.4):
 *   split^T(g)[u]   = merge(g)[u]   - sum_S cos(pi S u / n) * (sum_t (-1)^t g_S[t]) / sqrt(n S)
 *   merge^T(g)_S[t] = split(g)_S[t] + (-1)^t * (sum_u g[u] cos(pi S u / n)) / sqrt(n S)
 * One block per row; reductions in a fixed order (deterministic). */
}  // extern "C"

namespace msb {
namespace {

constexpr int kMaxAdjBands = 12;
struct AdjBands {
  float* ptr[kMaxAdjBands];
  int size[kMaxAdjBands];
  int nbands;
};

// sum over the block of v, fixed order; result valid in every thread
__device__ __forceinline__ float block_sum_fixed(float v, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) t += sh[w];
  return t;
}

__device__ __forceinline__ float cos_band(int u, int S, int n) {
  // cos(pi * S * u / n), S and n powers of two: reduce u modulo the period 2n/S exactly
  const int period = 2 * (n / S);
  return cospif(static_cast<float>(u & (period - 1)) * (static_cast<float>(S) / static_cast<float>(n)));
}

__global__ void __launch_bounds__(256)
split_adjoint_fix_kernel(AdjBands bands, float* __restrict__ dx, int n) {
  __shared__ float sh[8];
  __shared__ float coef[kMaxAdjBands];
  const size_t row = blockIdx.x;
  for (int i = 0; i < bands.nbands; ++i) {
    const int S = bands.size[i];
    float a = 0.f;
    if (S < n) {
      const float* g = bands.ptr[i] + row * S;
      for (int t = threadIdx.x; t < S; t += blockDim.x) a += (t & 1) ? -g[t] : g[t];
    }
    a = block_sum_fixed(a, sh);
    if (threadIdx.x == 0)
      coef[i] = S < n ? a / (sqrtf(static_cast<float>(n)) * sqrtf(static_cast<float>(S))) : 0.f;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < n; u += blockDim.x) {
    float c = 0.f;
    for (int i = 0; i < bands.nbands; ++i)
      if (bands.size[i] < n) c += coef[i] * cos_band(u, bands.size[i], n);
    dx[row * n + u] -= c;
  }
}

__global__ void __launch_bounds__(256)
merge_adjoint_fix_kernel(AdjBands bands, const float* __restrict__ dy, int n) {
  __shared__ float sh[8];
  __shared__ float coef[kMaxAdjBands];
  const size_t row = blockIdx.x;
  const float* g = dy + row * n;
  for (int i = 0; i < bands.nbands; ++i) {
    const int S = bands.size[i];
    float a = 0.f;
    if (S < n)
      for (int u = threadIdx.x; u < n; u += blockDim.x) a += g[u] * cos_band(u, S, n);
    a = block_sum_fixed(a, sh);
    if (threadIdx.x == 0)
      coef[i] = S < n ? a / (sqrtf(static_cast<float>(n)) * sqrtf(static_cast<float>(S))) : 0.f;
  }
  __syncthreads();
  for (int i = 0; i < bands.nbands; ++i) {
    const int S = bands.size[i];
    if (S >= n) continue;
    float* d = bands.ptr[i] + row * S;
    const float c = coef[i];
    for (int t = threadIdx.x; t < S; t += blockDim.x) d[t] += (t & 1) ? -c : c;
  }
}

bool fill_adj(AdjBands* a, float* const* ptrs, const int* sizes, int nbands, int n) {
  if (ptrs == nullptr || sizes == nullptr || nbands <= 0 || nbands > kMaxAdjBands) return false;
  if (n <= 0 || (n & (n - 1)) != 0) return false;
  a->nbands = nbands;
  for (int i = 0; i < nbands; ++i) {
    if (ptrs[i] == nullptr || sizes[i] <= 1 || (sizes[i] & (sizes[i] - 1)) != 0 || sizes[i] > n)
      return false;
    a->ptr[i] = ptrs[i];
    a->size[i] = sizes[i];
  }
  return true;
}

}  // namespace
}  // namespace msb

extern "C" {

ms_status ms_fft_decompose_adjoint_fix(const float* const* dbands, const int* sizes, int nbands,
                                       int batch, int n, float* dx, void* stream) {
  msb::AdjBands a;
  if (batch <= 0 || dx == nullptr ||
      !msb::fill_adj(&a, const_cast<float* const*>(dbands), sizes, nbands, n))
    return MS_ERR_INVALID;
  msb::split_adjoint_fix_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dx, n);
  return msb::after_launch("split_adjoint_fix_kernel");
}

ms_status ms_fft_recompose_adjoint_fix(const float* dy, int batch, int n, float* const* dbands,
                                       const int* sizes, int nbands, void* stream) {
  msb::AdjBands a;
  if (batch <= 0 || dy == nullptr || !msb::fill_adj(&a, dbands, sizes, nbands, n))
    return MS_ERR_INVALID;
  msb::merge_adjoint_fix_kernel<<<batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, dy, n);
  return msb::after_launch("merge_adjoint_fix_kernel");
}

}  // extern "C"
