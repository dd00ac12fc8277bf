// STFT / iSTFT (+ fused cIRM decompress & complex mask) and the elementwise mask ops.
//
// Reference semantics: audio_zen/acoustics/feature.py:9-91 (-> torch.stft / torch.istft),
// audio_zen/acoustics/mask.py:7-64, recipes/dns_interspeech_2020/inferencer.py:136-143.
//
// FFT: shared-memory radix-2 DIT, two real frames packed into one complex transform
// (frame A -> real lane, frame B -> imaginary lane), FR frames per CTA so that the [B,F,T]
// (T-contiguous) stores / loads of the reference layout are FR*4-byte segments.
#include "fsn_common.cuh"

namespace fsn {

constexpr int kFR = 16;        // frames per CTA
constexpr int kDspThreads = 256;

__device__ __forceinline__ int ilog2(int n) { return 31 - __clz(n); }

// in-place radix-2 DIT over `npairs` independent transforms whose inputs are already in
// bit-reversed order; tw[k] = exp(-2*pi*i*k/n)
template <bool INVERSE>
__device__ __forceinline__ void fft_radix2_smem(float2* z, int zstride, int npairs, int n, int log2n,
                                                const float2* tw) {
  const int nb_log = log2n - 1;
  const int total = npairs << nb_log;
  for (int s = 1; s <= log2n; ++s) {
    const int half = 1 << (s - 1);
    const int tw_shift = log2n - s;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
      const int p = idx >> nb_log;
      const int j = idx & ((1 << nb_log) - 1);
      const int pos = j & (half - 1);
      const int i0 = ((j >> (s - 1)) << s) + pos;
      const int i1 = i0 + half;
      float2 w = tw[pos << tw_shift];
      if (INVERSE) w.y = -w.y;
      float2* zp = z + p * zstride;
      const float2 a = zp[i0], b = zp[i1];
      const float tx = b.x * w.x - b.y * w.y;
      const float ty = b.x * w.y + b.y * w.x;
      zp[i0] = make_float2(a.x + tx, a.y + ty);
      zp[i1] = make_float2(a.x - tx, a.y - ty);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void init_tables(float2* tw, float* win, int n, int win_length) {
  for (int k = threadIdx.x; k < n / 2; k += blockDim.x) {
    float s, c;
    sincospif(-2.0f * (float)k / (float)n, &s, &c);
    tw[k] = make_float2(c, s);
  }
  const int left = (n - win_length) / 2;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int m = i - left;
    win[i] = (m >= 0 && m < win_length) ? 0.5f - 0.5f * cospif(2.0f * (float)m / (float)win_length) : 0.0f;
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDspThreads)
stft_kernel(const float* __restrict__ wav, int L, int n, int hop, int win_length, int T,
            float* __restrict__ mag, float* __restrict__ phase, float* __restrict__ real,
            float* __restrict__ imag, float* __restrict__ magT, int T_pad) {
  extern __shared__ float2 smem2[];
  const int log2n = ilog2(n);
  const int zstride = n + 1;
  constexpr int NP = kFR / 2;
  float2* z = smem2;
  float2* tw = z + NP * zstride;
  float* win = reinterpret_cast<float*>(tw + n / 2);
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kFR;
  const int F = n / 2 + 1;
  init_tables(tw, win, n, win_length);
  __syncthreads();
  const float* x = wav + (size_t)b * L;
  for (int idx = threadIdx.x; idx < NP * n; idx += blockDim.x) {
    const int p = idx >> log2n;
    const int i = idx & (n - 1);
    const int ta = t0 + 2 * p, tb = ta + 1;
    const float w = win[i];
    float va = 0.f, vb = 0.f;
    if (ta < T) va = x[reflect_idx(ta * hop + i - n / 2, L)] * w;
    if (tb < T) vb = x[reflect_idx(tb * hop + i - n / 2, L)] * w;
    z[p * zstride + (int)(__brev((unsigned)i) >> (32 - log2n))] = make_float2(va, vb);
  }
  __syncthreads();
  fft_radix2_smem<false>(z, zstride, NP, n, log2n, tw);

  // un-pack the two real transforms and store in the reference layout [B,F,T]
  const size_t plane = (size_t)F * T;
  for (int idx = threadIdx.x; idx < F * kFR; idx += blockDim.x) {
    const int k = idx / kFR;
    const int j = idx - k * kFR;
    const int t = t0 + j;
    if (t >= T) continue;
    const float2 zk = z[(j >> 1) * zstride + k];
    const float2 zn = z[(j >> 1) * zstride + ((n - k) & (n - 1))];
    float re, im;
    if ((j & 1) == 0) { re = 0.5f * (zk.x + zn.x); im = 0.5f * (zk.y - zn.y); }
    else              { re = 0.5f * (zk.y + zn.y); im = -0.5f * (zk.x - zn.x); }
    const size_t o = (size_t)b * plane + (size_t)k * T + t;
    if (real) real[o] = re;
    if (imag) imag[o] = im;
    if (mag) mag[o] = hypotf(re, im);
    if (phase) phase[o] = atan2f(im, re);
  }
  if (magT) {  // time-major copy with the look-ahead rows zeroed
    for (int idx = threadIdx.x; idx < F * kFR; idx += blockDim.x) {
      const int j = idx / F;
      const int k = idx - j * F;
      const int t = t0 + j;
      if (t >= T_pad) continue;
      float m = 0.f;
      if (t < T) {
        const float2 zk = z[(j >> 1) * zstride + k];
        const float2 zn = z[(j >> 1) * zstride + ((n - k) & (n - 1))];
        float re, im;
        if ((j & 1) == 0) { re = 0.5f * (zk.x + zn.x); im = 0.5f * (zk.y - zn.y); }
        else              { re = 0.5f * (zk.y + zn.y); im = -0.5f * (zk.x - zn.x); }
        m = hypotf(re, im);
      }
      magT[((size_t)b * T_pad + t) * F + k] = m;
    }
  }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kDspThreads)
istft_kernel(const float* __restrict__ real, const float* __restrict__ imag, int cstride,
             const float* __restrict__ crm, int mask_mode, int T, int n, int hop, int win_length, int out_len,
             int seg, int np_max, float* __restrict__ wav, unsigned int* __restrict__ peak_bits) {
  extern __shared__ float2 smem2[];
  const int log2n = ilog2(n);
  const int zstride = n + 1;
  float2* z = smem2;
  float2* tw = z + np_max * zstride;
  float* win = reinterpret_cast<float*>(tw + n / 2);
  const int b = blockIdx.y;
  const int F = n / 2 + 1;
  const int s_begin = n / 2 + blockIdx.x * seg;
  const int s_end = min(s_begin + seg, n / 2 + out_len);
  const int t_min = (s_begin >= n) ? (s_begin - n) / hop + 1 : 0;
  const int t_max = min(T - 1, (s_end - 1) / hop);
  const int nframes = t_max - t_min + 1;
  const int np = nframes > 0 ? (nframes + 1) / 2 : 0;
  init_tables(tw, win, n, win_length);
  const size_t plane = (size_t)F * T;
  const float* xr = real + (size_t)b * plane * cstride;
  const float* xi = imag + (size_t)b * plane * cstride;
  const float* cr = crm ? crm + (size_t)b * 2 * plane : nullptr;
  const float* ci = crm ? cr + plane : nullptr;
  for (int idx = threadIdx.x; idx < F * np; idx += blockDim.x) {
    const int k = idx / np;
    const int p = idx - k * np;
    float e[2][2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int t = t_min + 2 * p + q;
      float r = 0.f, i = 0.f;
      if (t <= t_max) {
        const size_t o = (size_t)k * T + t;
        r = xr[o * cstride];
        i = xi[o * cstride];
        if (crm && mask_mode == 2) {  // improved_fullsubnet/model.py:575-576: element-wise, no decompression
          r *= cr[o];
          i *= ci[o];
        } else if (crm) {  // mask.py:58-63 then inferencer.py:139-140
          const float mr = decompress_cirm_f(cr[o], 10.0f, 9.9f);
          const float mi = decompress_cirm_f(ci[o], 10.0f, 9.9f);
          const float er = mr * r - mi * i;
          const float ei = mi * r + mr * i;
          r = er; i = ei;
        }
      }
      e[q][0] = r;
      e[q][1] = (k == 0 || k == n / 2) ? 0.f : i;  // irfft ignores Im of DC / Nyquist
    }
    // Z = Ea + i*Eb on the full circle (Hermitian extension of both)
    z[p * zstride + (int)(__brev((unsigned)k) >> (32 - log2n))] =
        make_float2(e[0][0] - e[1][1], e[0][1] + e[1][0]);
    if (k > 0 && k < n / 2)
      z[p * zstride + (int)(__brev((unsigned)(n - k)) >> (32 - log2n))] =
          make_float2(e[0][0] + e[1][1], -e[0][1] + e[1][0]);
  }
  __syncthreads();
  fft_radix2_smem<true>(z, zstride, np, n, log2n, tw);

  const int full = n + hop * (T - 1);
  const float inv_n = 1.0f / (float)n;
  float* out = wav + (size_t)b * out_len;
  float peak = 0.f;
  for (int s = s_begin + threadIdx.x; s < s_end; s += blockDim.x) {
    float acc = 0.f, env = 0.f;
    if (s < full) {
      const int tl = max(t_min, (s >= n) ? (s - n) / hop + 1 : 0);
      const int th = min(t_max, s / hop);
      for (int t = tl; t <= th; ++t) {
        const int i = s - t * hop;
        const int q = t - t_min;
        const float2 v = z[(q >> 1) * zstride + i];
        const float w = win[i];
        acc += ((q & 1) ? v.y : v.x) * inv_n * w;
        env += w * w;
      }
    }
    const float y = (env > 1e-11f) ? acc / env : 0.f;
    out[s - n / 2] = y;
    peak = fmaxf(peak, fabsf(y));
  }
  if (peak_bits) {
    // max|y| of the clip for the int16 scaling of the host loop (base_inferencer.py:181-182): non-negative floats
    // order like their bit patterns, and max is order-independent, so the atomic is deterministic
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) peak = fmaxf(peak, __shfl_xor_sync(0xffffffffu, peak, o));
    if ((threadIdx.x & 31) == 0 && peak > 0.f) atomicMax(peak_bits + b, __float_as_uint(peak));
  }
}

// out = int16(gain * wav / peak) per clip, peak from the iSTFT epilogue (float32 mul, div, truncation like numpy)
__global__ void scale_int16_kernel(const float* __restrict__ wav, const unsigned int* __restrict__ peak_bits, int L, float gain,
                                   int16_t* __restrict__ out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float m = __uint_as_float(peak_bits[i / L]);
    out[i] = (m > 0.f) ? (int16_t)__fdiv_rn(__fmul_rn(gain, wav[i]), m) : (int16_t)0;
  }
}

// ------------------------------------------------------------------------------------------
__global__ void decompress_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float K,
                                  float limit) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = decompress_cirm_f(in[i], K, limit);
}

// mask.py:38-40
__device__ __forceinline__ float compress_cirm_f(float m, float K, float C) {
  m = (m <= -100.f) ? -100.f : m;
  const float e = expf(-C * m);
  return K * (1.f - e) / (1.f + e);
}

__global__ void compress_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n, float K, float C) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = compress_cirm_f(in[i], K, C);
}

// mask.py:22-29
__global__ void build_cirm_kernel(const float* __restrict__ nr, const float* __restrict__ ni,
                                  const float* __restrict__ cr, const float* __restrict__ ci,
                                  float2* __restrict__ out, int64_t n) {
  const float eps = 1.1920928955078125e-07f;  // audio_zen/constant.py:9
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float a = nr[i], b = ni[i], c = cr[i], d = ci[i];
    const float den = a * a + b * b + eps;
    out[i] = make_float2(compress_cirm_f((a * c + b * d) / den, 10.f, 0.1f),
                         compress_cirm_f((a * d - b * c) / den, 10.f, 0.1f));
  }
}

// feature.py:332-345: output clip b' of group g <- clip g + G*i, frequency f' <- g + G*f'
__global__ void drop_band_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int C, int F,
                                 int T, int G) {
  const int Fo = F / G;
  const int64_t total = (int64_t)B * C * Fo * T;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i % T);
    int64_t r = i / T;
    const int fo = (int)(r % Fo); r /= Fo;
    const int c = (int)(r % C);
    int bo = (int)(r / C);
    int g = 0;
    for (; g < G; ++g) {  // group g holds ceil((B-g)/G) clips
      const int cnt = (B - g + G - 1) / G;
      if (bo < cnt) break;
      bo -= cnt;
    }
    const int bi = g + G * bo;
    const int fi = g + G * fo;
    out[i] = in[(((int64_t)bi * C + c) * F + fi) * T + t];
  }
}

// base_inferencer.py:181-182: int16(0.8 * 32767 * y / max|y|) per clip; one CTA per clip (max reduce, then scale)
__global__ void peak_normalize_int16_kernel(const float* __restrict__ wav, int L, float gain, int16_t* __restrict__ out) {
  __shared__ float sh[256];
  const float* x = wav + (size_t)blockIdx.x * L;
  float m = 0.f;
  for (int i = threadIdx.x; i < L; i += blockDim.x) m = fmaxf(m, fabsf(x[i]));
  sh[threadIdx.x] = m;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] = fmaxf(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  m = sh[0];
  int16_t* o = out + (size_t)blockIdx.x * L;
  for (int i = threadIdx.x; i < L; i += blockDim.x)
    o[i] = (m > 0.f) ? (int16_t)__fdiv_rn(__fmul_rn(gain, x[i]), m) : (int16_t)0;  // float32 mul, div, truncation like numpy
}

// audio_zen/metrics.py:6-31 SI_SDR(reference, estimation) per clip: alpha = <ref,est>/<ref,ref>;
// 10 log10(|alpha ref|^2 / |est - alpha ref|^2).  One CTA per clip, two passes, fixed-order tree reductions in
// double (the reference sums in float32 pairwise; both agree to ~1e-5 dB on 4 s clips).
__global__ void si_sdr_kernel(const float* __restrict__ ref, const float* __restrict__ est, int L, float* __restrict__ out) {
  __shared__ double sa[256], sb[256];
  const float* r = ref + (size_t)blockIdx.x * L;
  const float* e = est + (size_t)blockIdx.x * L;
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) { a += (double)r[i] * r[i]; b += (double)r[i] * e[i]; }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  const float alpha = (float)sb[0] / (float)sa[0];  // float32 like the reference's optimal_scaling
  __syncthreads();
  a = 0.0; b = 0.0;
  for (int i = threadIdx.x; i < L; i += blockDim.x) {
    const float p = alpha * r[i];
    const float n = e[i] - p;
    a += (double)p * p; b += (double)n * n;
  }
  sa[threadIdx.x] = a; sb[threadIdx.x] = b;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sa[threadIdx.x] += sa[threadIdx.x + s]; sb[threadIdx.x] += sb[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(10.0 * log10(sa[0] / sb[0]));
}

static int ew_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

// fsn_dsp_dft.cu: direct-DFT variants for even transform sizes that are not a power of two (e.g. 960)
int stft_dft_launch(const float* wav, int B, int L, int n_fft, int hop, int win_length, int T, int Tg, float* mag,
                    float* phase, float* real, float* imag, float* magT, int T_pad, cudaStream_t st);
int istft_dft_launch(const float* real, const float* imag, int cstride, const float* crm, int mask_mode, int B, int T,
                     int n_fft, int hop, int win_length, int out_len, float* wav, cudaStream_t st);
static bool dft_size_ok(int n) { return !is_pow2(n) && (n & 1) == 0 && n >= 16 && n <= 1200; }

int stft_launch(const float* wav, int B, int L, int n_fft, int hop, int win_length, float* mag, float* phase,
                float* real, float* imag, float* magT, int T_pad, cudaStream_t st) {
  FSN_REQUIRE(B > 0 && L > 0, FSN_ERR_SHAPE, "stft: empty input (B=%d, L=%d)", B, L);
  if (dft_size_ok(n_fft)) {
    FSN_REQUIRE(hop > 0 && win_length > 0 && win_length <= n_fft, FSN_ERR_SHAPE, "stft: bad hop/win_length");
    FSN_REQUIRE(n_fft / 2 < L, FSN_ERR_SHAPE, "stft: reflect padding %d needs L > pad (L=%d)", n_fft / 2, L);
    FSN_REQUIRE(!magT || T_pad >= 1 + L / hop, FSN_ERR_SHAPE, "stft: T_pad < T");
    const int Td = 1 + L / hop;
    return stft_dft_launch(wav, B, L, n_fft, hop, win_length, Td, magT ? (T_pad > Td ? T_pad : Td) : Td, mag, phase, real,
                           imag, magT, T_pad, st);
  }
  FSN_REQUIRE(is_pow2(n_fft) && n_fft >= 16 && n_fft <= 2048, FSN_ERR_UNSUPPORTED,
              "stft: n_fft=%d unsupported (power of two in [16,2048], or even and <= 1200)", n_fft);
  FSN_REQUIRE(hop > 0 && win_length > 0 && win_length <= n_fft, FSN_ERR_SHAPE, "stft: bad hop/win_length");
  FSN_REQUIRE(n_fft / 2 < L, FSN_ERR_SHAPE, "stft: reflect padding %d needs L > pad (L=%d)", n_fft / 2, L);
  const int T = 1 + L / hop;
  const int Tg = magT ? (T_pad > T ? T_pad : T) : T;
  FSN_REQUIRE(!magT || T_pad >= T, FSN_ERR_SHAPE, "stft: T_pad < T");
  const size_t smem = (size_t)(kFR / 2) * (n_fft + 1) * 8 + (size_t)n_fft / 2 * 8 + (size_t)n_fft * 4;
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "stft smem attr");
    if (rc) return rc;
  }
  dim3 grid(cdiv(Tg, kFR), B);
  stft_kernel<<<grid, kDspThreads, smem, st>>>(wav, L, n_fft, hop, win_length, T, mag, phase, real, imag, magT,
                                               T_pad);
  FSN_CHECK_LAUNCH("stft_kernel");
  return FSN_OK;
}

int istft_launch(const float* real, const float* imag, int cstride, const float* crm, int B, int T, int n_fft,
                 int hop, int win_length, int length, float* wav, cudaStream_t st, int mask_mode, unsigned int* peak_bits) {
  FSN_REQUIRE(B > 0 && T > 0, FSN_ERR_SHAPE, "istft: empty input");
  if (peak_bits) {
    FSN_REQUIRE(!dft_size_ok(n_fft), FSN_ERR_UNSUPPORTED, "istft: the fused peak is built for the power-of-two transform");
    int rc = check_cuda(cudaMemsetAsync(peak_bits, 0, (size_t)B * sizeof(unsigned int), st), "istft peak memset");
    if (rc) return rc;
  }
  if (dft_size_ok(n_fft)) {
    FSN_REQUIRE(hop > 0 && hop <= n_fft && win_length > 0 && win_length <= n_fft, FSN_ERR_SHAPE,
                "istft: bad hop/win_length");
    FSN_REQUIRE(cstride == 1 || cstride == 2, FSN_ERR_SHAPE, "istft: cstride must be 1 or 2");
    const int olen = length > 0 ? length : hop * (T - 1);
    FSN_REQUIRE(olen > 0, FSN_ERR_SHAPE, "istft: output length %d", olen);
    return istft_dft_launch(real, imag, cstride, crm, mask_mode, B, T, n_fft, hop, win_length, olen, wav, st);
  }
  FSN_REQUIRE(is_pow2(n_fft) && n_fft >= 16 && n_fft <= 2048, FSN_ERR_UNSUPPORTED,
              "istft: n_fft=%d unsupported (power of two in [16,2048], or even and <= 1200)", n_fft);
  FSN_REQUIRE(hop > 0 && hop <= n_fft && win_length > 0 && win_length <= n_fft, FSN_ERR_SHAPE,
              "istft: bad hop/win_length");
  FSN_REQUIRE(cstride == 1 || cstride == 2, FSN_ERR_SHAPE, "istft: cstride must be 1 or 2");
  const int out_len = length > 0 ? length : hop * (T - 1);
  FSN_REQUIRE(out_len > 0, FSN_ERR_SHAPE, "istft: output length %d", out_len);
  const int seg = kFR * hop;
  const int np_max = (kFR + cdiv(n_fft, hop) + 2) / 2;
  const size_t smem = (size_t)np_max * (n_fft + 1) * 8 + (size_t)n_fft / 2 * 8 + (size_t)n_fft * 4;
  if (smem > 48 * 1024) {
    int rc = check_cuda(cudaFuncSetAttribute(istft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                        "istft smem attr");
    if (rc) return rc;
  }
  dim3 grid(cdiv(out_len, seg), B);
  istft_kernel<<<grid, kDspThreads, smem, st>>>(real, imag, cstride, crm, mask_mode, T, n_fft, hop, win_length, out_len,
                                                seg, np_max, wav, peak_bits);
  FSN_CHECK_LAUNCH("istft_kernel");
  return FSN_OK;
}

}  // namespace fsn

using namespace fsn;

extern "C" int fsn_stft(const float* wav, int B, int L, int n_fft, int hop, int win_length, float* mag,
                        float* phase, float* real, float* imag, float* magT, int T_pad, fsn_stream_t stream) {
  return stft_launch(wav, B, L, n_fft, hop, win_length, mag, phase, real, imag, magT, T_pad, (cudaStream_t)stream);
}

extern "C" int fsn_istft(const float* real, const float* imag, int cstride, const float* crm, int B, int T,
                         int n_fft, int hop, int win_length, int length, float* wav, fsn_stream_t stream) {
  return istft_launch(real, imag, cstride, crm, B, T, n_fft, hop, win_length, length, wav, (cudaStream_t)stream, 1, nullptr);
}

extern "C" int fsn_peak_normalize_int16(const float* wav, int B, int L, float gain, int16_t* out, fsn_stream_t stream) {
  FSN_REQUIRE(B > 0 && L > 0, FSN_ERR_SHAPE, "peak_normalize: empty input");
  peak_normalize_int16_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(wav, L, gain, out);
  FSN_CHECK_LAUNCH("peak_normalize_int16_kernel");
  return FSN_OK;
}

extern "C" int fsn_si_sdr(const float* reference, const float* estimation, int B, int L, float* out, fsn_stream_t stream) {
  FSN_REQUIRE(B > 0 && L > 0, FSN_ERR_SHAPE, "si_sdr: empty input");
  si_sdr_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(reference, estimation, L, out);
  FSN_CHECK_LAUNCH("si_sdr_kernel");
  return FSN_OK;
}

// int16 scaling with the per-clip peak the iSTFT epilogue produced (fsn_enhance_pcm)
namespace fsn {
int scale_int16_launch(const float* wav, const unsigned int* peak_bits, int B, int L, float gain, int16_t* out, cudaStream_t st) {
  const size_t n = (size_t)B * L;
  scale_int16_kernel<<<ew_grid((int64_t)n), 256, 0, st>>>(wav, peak_bits, L, gain, out, n);
  FSN_CHECK_LAUNCH("scale_int16_kernel");
  return FSN_OK;
}
}  // namespace fsn

extern "C" int fsn_decompress_cirm(const float* in, float* out, int64_t n, float K, float limit,
                                   fsn_stream_t stream) {
  if (n <= 0) return FSN_OK;
  decompress_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, K, limit);
  FSN_CHECK_LAUNCH("decompress_kernel");
  return FSN_OK;
}

extern "C" int fsn_compress_cirm(const float* in, float* out, int64_t n, float K, float C, fsn_stream_t stream) {
  if (n <= 0) return FSN_OK;
  compress_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(in, out, n, K, C);
  FSN_CHECK_LAUNCH("compress_kernel");
  return FSN_OK;
}

extern "C" int fsn_build_cirm(const float* nr, const float* ni, const float* cr, const float* ci, float* out,
                              int64_t n, fsn_stream_t stream) {
  if (n <= 0) return FSN_OK;
  build_cirm_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(nr, ni, cr, ci, reinterpret_cast<float2*>(out), n);
  FSN_CHECK_LAUNCH("build_cirm_kernel");
  return FSN_OK;
}

extern "C" int fsn_drop_band(const float* in, float* out, int B, int C, int F, int T, int G, fsn_stream_t stream) {
  FSN_REQUIRE(B > G, FSN_ERR_SHAPE,
              "Batch size = %d, num_groups = %d. The batch size should larger than the num_groups.", B, G);
  FSN_REQUIRE(G >= 2, FSN_ERR_SHAPE, "drop_band: G < 2 is the identity, handle on the host");
  const int64_t n = (int64_t)B * C * (F / G) * T;
  if (n <= 0) return FSN_OK;
  drop_band_kernel<<<ew_grid(n), 256, 0, (cudaStream_t)stream>>>(in, out, B, C, F, T, G);
  FSN_CHECK_LAUNCH("drop_band_kernel");
  return FSN_OK;
}
