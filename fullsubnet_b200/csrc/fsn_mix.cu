// Training-data mixing on the device: the arithmetic of Dataset.snr_mix
// (recipes/dns_interspeech_2020/dataset_train.py:136-199) for a whole batch of (clean, noise) pairs -
// optional reverberation (scipy.signal.fftconvolve(clean, rir)[:L], :161), norm_amplitude + tailor_dB_FS of both
// (audio_zen/acoustics/feature.py:99-111), SNR scaling, the common dBFS of the mixture and the anti-clipping rescale.
// The random draws of the reference (SNR, noisy target dBFS, RIR choice) stay on the host and come in as arrays,
// so that the result is a pure function of its inputs.  One CTA per clip, fixed-order reductions.
#include "fsn_internal.cuh"

namespace fsn {
namespace mix {

constexpr int THREADS = 1024;

__device__ __forceinline__ double block_sum(double v, double* sh) {
  __syncthreads();
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}
__device__ __forceinline__ float block_max(float v, double* sh) {
  __syncthreads();
  sh[threadIdx.x] = (double)v;
  __syncthreads();
  for (int s = THREADS / 2; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] = fmax(sh[threadIdx.x], sh[threadIdx.x + s]);
    __syncthreads();
  }
  return (float)sh[0];
}

// out[b, n] = sum_k x[b, n-k] rir[b, k], n < L (the first L samples of the full convolution); clips with
// rir_len[b] == 0 are copied.  Direct form: 256 outputs per CTA, taps staged through shared memory in chunks.
constexpr int CO = 256, CK = 1024;
__global__ void __launch_bounds__(CO) rir_conv_kernel(const float* __restrict__ x, const float* __restrict__ rir,
                                                      const int* __restrict__ rir_len, int L, int Lr_max,
                                                      float* __restrict__ out) {
  __shared__ float taps[CK];
  __shared__ float seg[CK + CO];
  const int b = blockIdx.y, n0 = blockIdx.x * CO, n = n0 + threadIdx.x;
  const float* xb = x + (size_t)b * L;
  const int lr = rir_len ? min(rir_len[b], Lr_max) : Lr_max;
  if (lr <= 0) {
    if (n < L) out[(size_t)b * L + n] = xb[n];
    return;
  }
  const float* rb = rir + (size_t)b * Lr_max;
  double acc = 0.0;  // double accumulation: the reference's FFT convolution carries ~1e-7 relative error per output
  for (int k0 = 0; k0 < lr && k0 <= n0 + CO - 1; k0 += CK) {
    __syncthreads();
    for (int i = threadIdx.x; i < CK; i += CO) taps[i] = (k0 + i < lr) ? rb[k0 + i] : 0.f;
    // x[n0 - k0 - (CK-1) .. n0 - k0 + CO - 1]
    const int base = n0 - k0 - (CK - 1);
    for (int i = threadIdx.x; i < CK + CO; i += CO) {
      const int j = base + i;
      seg[i] = (j >= 0 && j < L) ? xb[j] : 0.f;
    }
    __syncthreads();
    // x[n - (k0 + kk)] = seg[(n - n0) + (CK - 1) - kk]
    const float* sp = seg + threadIdx.x + (CK - 1);
#pragma unroll 8
    for (int kk = 0; kk < CK; ++kk) acc += (double)taps[kk] * (double)sp[-kk];
  }
  if (n < L) out[(size_t)b * L + n] = (float)acc;
}

__global__ void __launch_bounds__(THREADS) snr_mix_kernel(const float* __restrict__ clean, const float* __restrict__ noise,
                                                          const float* __restrict__ snr, const float* __restrict__ noisy_target,
                                                          float target_dB_FS, float eps, int L, float* __restrict__ noisy_out,
                                                          float* __restrict__ clean_out) {
  __shared__ double sh[THREADS];
  const int b = blockIdx.x;
  const float* c = clean + (size_t)b * L;
  const float* n = noise + (size_t)b * L;
  float mc = 0.f, mn = 0.f;
  for (int i = threadIdx.x; i < L; i += THREADS) { mc = fmaxf(mc, fabsf(c[i])); mn = fmaxf(mn, fabsf(n[i])); }
  const float sc1 = block_max(mc, sh) + eps, sn1 = block_max(mn, sh) + eps;  // norm_amplitude: y / (max|y| + eps)
  double a = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < L; i += THREADS) {
    const float u = c[i] / sc1, v = n[i] / sn1;
    a += (double)u * u; d += (double)v * v;
  }
  const float tgt = powf(10.0f, target_dB_FS / 20.0f);
  const float kc = tgt / ((float)sqrt(block_sum(a, sh) / L) + eps);   // tailor_dB_FS scalar of the clean speech
  const float kn = tgt / ((float)sqrt(block_sum(d, sh) / L) + eps);   // ... of the noise
  a = 0.0; d = 0.0;
  for (int i = threadIdx.x; i < L; i += THREADS) {
    const float u = (c[i] / sc1) * kc, v = (n[i] / sn1) * kn;
    a += (double)u * u; d += (double)v * v;
  }
  const float clean_rms = (float)sqrt(block_sum(a, sh) / L), noise_rms = (float)sqrt(block_sum(d, sh) / L);
  const float snr_scalar = clean_rms / powf(10.0f, snr[b] / 20.0f) / (noise_rms + eps);
  a = 0.0;
  for (int i = threadIdx.x; i < L; i += THREADS) {
    const float y = (c[i] / sc1) * kc + ((n[i] / sn1) * kn) * snr_scalar;
    a += (double)y * y;
  }
  const float ky = powf(10.0f, noisy_target[b] / 20.0f) / ((float)sqrt(block_sum(a, sh) / L) + eps);
  float my = 0.f;
  for (int i = threadIdx.x; i < L; i += THREADS) {
    const float y = ((c[i] / sc1) * kc + ((n[i] / sn1) * kn) * snr_scalar) * ky;
    my = fmaxf(my, fabsf(y));
  }
  my = block_max(my, sh);
  const bool clipped = my > 0.999f;                       // is_clipped (feature.py:113-114)
  const float kclip = clipped ? my / (0.99f - eps) : 1.0f;
  for (int i = threadIdx.x; i < L; i += THREADS) {
    const float u = (c[i] / sc1) * kc;
    float y = (u + ((n[i] / sn1) * kn) * snr_scalar) * ky;
    float cl = u * ky;
    if (clipped) { y = y / kclip; cl = cl / kclip; }
    noisy_out[(size_t)b * L + i] = y;
    clean_out[(size_t)b * L + i] = cl;
  }
}

}  // namespace mix
}  // namespace fsn

using namespace fsn;

extern "C" int fsn_rir_convolve(const float* x, const float* rir, const int* rir_len, int B, int L, int Lr_max, float* out,
                                fsn_stream_t stream) {
  FSN_REQUIRE(B > 0 && L > 0 && Lr_max > 0, FSN_ERR_SHAPE, "rir_convolve: empty input");
  mix::rir_conv_kernel<<<dim3(cdiv(L, mix::CO), B), mix::CO, 0, (cudaStream_t)stream>>>(x, rir, rir_len, L, Lr_max, out);
  FSN_CHECK_LAUNCH("rir_conv_kernel");
  return FSN_OK;
}

extern "C" int fsn_snr_mix(const float* clean, const float* noise, const float* snr, const float* noisy_target_dB_FS,
                           float target_dB_FS, float eps, int B, int L, float* noisy_out, float* clean_out,
                           fsn_stream_t stream) {
  FSN_REQUIRE(B > 0 && L > 0, FSN_ERR_SHAPE, "snr_mix: empty input");
  mix::snr_mix_kernel<<<B, mix::THREADS, 0, (cudaStream_t)stream>>>(clean, noise, snr, noisy_target_dB_FS, target_dB_FS, eps, L,
                                                                    noisy_out, clean_out);
  FSN_CHECK_LAUNCH("snr_mix_kernel");
  return FSN_OK;
}
