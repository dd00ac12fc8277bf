// Full-band 2-layer LSTM as ONE persistent cooperative kernel (fp32 FMA, exact-class arithmetic).
//
// Reference semantics: audio_zen/model/module/sequence_model.py:106-125 (nn.LSTM part) as used by
// recipes/dns_interspeech_2020/fullsubnet/model.py:92-95 (rows = clips, input = normalised magnitude).
//
// Decomposition.  The batch is small (B clips) and the recurrence is serial in t, so the hidden
// dimension is spread over the whole chip: CTA j owns `upc` hidden units of BOTH layers, i.e.
// 4*upc gate rows of W_ih/W_hh per layer, which it keeps resident in shared memory as fp32 for the
// whole sequence (F=257, H=512, upc=4: 16 x (769 + 1024) x 4 B = 115 KB) - weights are read from
// HBM exactly once.  The two layers run as a wavefront: in phase p every CTA computes its slice of
// layer 0 at step p and of layer 1 at step p-1; both only need data of phase p-1
// (x_p, h0_{p-1}, h1_{p-2}), so there is ONE grid-wide barrier per time step.  h is exchanged through
// global memory (L2); c stays in registers.
//
// Thread mapping (512 threads): row = tid/2 (clip), half = tid%2 -> gate columns [8*half, 8*half+8)
// of the CTA's 16 (= 2 complete hidden units), for both layers: 32 accumulators per thread.
#include <cooperative_groups.h>

#include "fsn_internal.cuh"

namespace fsn {
namespace fb {

constexpr int ROWS = 256;      // clips per launch (host loops over chunks)
constexpr int THREADS = 512;
constexpr int KC = 32;         // k-chunk staged in shared memory
constexpr int MAX_UPC = 4;     // hidden units per CTA (=> 16 gate columns)

struct Args {
  const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  const float* x;        // magT [B, Tp, F]
  const float* inv1;     // [B]
  float* h0buf;          // [2][B][H] ping-pong
  float* h1all;          // [B][Tp][H]
  unsigned int* barrier; // grid barrier counter (zeroed by the host before launch)
  int B, F, H, Tp, upc, G;
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    unsigned int spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      if (++spins > (1u << 28)) { printf("fsn fb: grid barrier timeout\n"); __trap(); }
    } while (v < target);
  }
  __syncthreads();
}

// shared-memory weight slice layout: Ws[layer][k][16] (gate column c = unit_local*4 + gate), k over
// [x | h_prev] of that layer; zero for units beyond H.
__global__ void __launch_bounds__(THREADS, 1) fb_lstm_kernel(const Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const int F = a.F, H = a.H, Tp = a.Tp, B = a.B;
  const int K0 = F + H, K1 = 2 * H;
  float* W0 = smem_f;                 // [K0][16]
  float* W1 = W0 + (size_t)K0 * 16;   // [K1][16]
  float* At = W1 + (size_t)K1 * 16;   // [ROWS][KC+1]
  const int tid = threadIdx.x;
  const int row = tid >> 1, half = tid & 1;
  const int u0 = blockIdx.x * a.upc;  // first hidden unit of this CTA

  // ---- one-time: weight slice -> shared memory (gate column c: unit u0 + c/4, gate c%4)
  for (int idx = tid; idx < K0 * 16; idx += THREADS) {
    const int k = idx >> 4, c = idx & 15;
    const int ul = c >> 2, g = c & 3, u = u0 + ul;
    float w = 0.f;
    if (ul < a.upc && u < H) {
      const size_t wr = (size_t)g * H + u;
      w = (k < F) ? a.w_ih[0][wr * F + k] : a.w_hh[0][wr * H + (k - F)];
    }
    W0[idx] = w;
  }
  for (int idx = tid; idx < K1 * 16; idx += THREADS) {
    const int k = idx >> 4, c = idx & 15;
    const int ul = c >> 2, g = c & 3, u = u0 + ul;
    float w = 0.f;
    if (ul < a.upc && u < H) {
      const size_t wr = (size_t)g * H + u;
      w = (k < H) ? a.w_ih[1][wr * H + k] : a.w_hh[1][wr * H + (k - H)];
    }
    W1[idx] = w;
  }
  float bias0[8], bias1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = half * 8 + j, ul = c >> 2, g = c & 3, u = u0 + ul;
    const bool ok = ul < a.upc && u < H;
    bias0[j] = ok ? a.b_ih[0][g * H + u] + a.b_hh[0][g * H + u] : 0.f;
    bias1[j] = ok ? a.b_ih[1][g * H + u] + a.b_hh[1][g * H + u] : 0.f;
  }
  float c0[2] = {0.f, 0.f}, c1[2] = {0.f, 0.f};  // cell state of the thread's 2 units, both layers
  const float scale = (row < B) ? a.inv1[row] : 0.f;
  __syncthreads();

  // A-tile loader: thread -> (row = tid/2, 16 consecutive k of the 32-chunk)
  const int l_row = tid >> 1, l_k = (tid & 1) * 16;

  for (int p = 0; p <= Tp; ++p) {
    const bool do0 = p < Tp;    // layer 0 at step p
    const bool do1 = p >= 1;    // layer 1 at step p-1
    float acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc0[j] = bias0[j]; acc1[j] = bias1[j]; }
    const float* h0_prev = a.h0buf + (size_t)((p + 1) & 1) * B * H;       // h0_{p-1}
    const float* h1_prev = a.h1all + (size_t)(p >= 2 ? p - 2 : 0) * H;    // h1_{p-2}, row stride Tp*H
    // three k segments: x_p (F, layer 0), h0_{p-1} (H, both layers), h1_{p-2} (H, layer 1)
    for (int seg = 0; seg < 3; ++seg) {
      if (seg == 0 && !do0) continue;
      if (seg == 1 && p == 0) continue;           // h0_{-1} = 0
      if (seg == 2 && p < 2) continue;            // h1_{-1} = 0
      const int klen = (seg == 0) ? F : H;
      const float* w0 = (seg == 0) ? W0 : W0 + (size_t)F * 16;          // layer-0 rows of this segment
      const float* w1 = (seg == 1) ? W1 : W1 + (size_t)H * 16;          // layer-1 rows of this segment
      const bool use0 = (seg <= 1) && do0, use1 = (seg >= 1) && do1;
      for (int k0 = 0; k0 < klen; k0 += KC) {
        __syncthreads();
        if (l_row < B) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int k = k0 + l_k + j;
            float v = 0.f;
            if (k < klen) {
              if (seg == 0) v = a.x[((size_t)l_row * Tp + p) * F + k] * scale;
              else if (seg == 1) v = __ldcg(h0_prev + (size_t)l_row * H + k);
              else v = __ldcg(h1_prev + (size_t)l_row * Tp * H + k);
            }
            At[l_row * (KC + 1) + l_k + j] = v;
          }
        }
        __syncthreads();
        const int kmax = min(KC, klen - k0);
        const float* ar = At + row * (KC + 1);
        if (use0 && use1) {
          for (int kk = 0; kk < kmax; ++kk) {
            const float av = ar[kk];
            const float4* p0 = reinterpret_cast<const float4*>(w0 + (size_t)(k0 + kk) * 16 + half * 8);
            const float4* p1 = reinterpret_cast<const float4*>(w1 + (size_t)(k0 + kk) * 16 + half * 8);
            const float4 wa = p0[0], wb = p0[1], wc = p1[0], wd = p1[1];
            acc0[0] = fmaf(av, wa.x, acc0[0]); acc0[1] = fmaf(av, wa.y, acc0[1]);
            acc0[2] = fmaf(av, wa.z, acc0[2]); acc0[3] = fmaf(av, wa.w, acc0[3]);
            acc0[4] = fmaf(av, wb.x, acc0[4]); acc0[5] = fmaf(av, wb.y, acc0[5]);
            acc0[6] = fmaf(av, wb.z, acc0[6]); acc0[7] = fmaf(av, wb.w, acc0[7]);
            acc1[0] = fmaf(av, wc.x, acc1[0]); acc1[1] = fmaf(av, wc.y, acc1[1]);
            acc1[2] = fmaf(av, wc.z, acc1[2]); acc1[3] = fmaf(av, wc.w, acc1[3]);
            acc1[4] = fmaf(av, wd.x, acc1[4]); acc1[5] = fmaf(av, wd.y, acc1[5]);
            acc1[6] = fmaf(av, wd.z, acc1[6]); acc1[7] = fmaf(av, wd.w, acc1[7]);
          }
        } else if (use0 || use1) {
          const float* w = use0 ? w0 : w1;
          float* acc = use0 ? acc0 : acc1;
          for (int kk = 0; kk < kmax; ++kk) {
            const float av = ar[kk];
            const float4* pw = reinterpret_cast<const float4*>(w + (size_t)(k0 + kk) * 16 + half * 8);
            const float4 wa = pw[0], wb = pw[1];
            acc[0] = fmaf(av, wa.x, acc[0]); acc[1] = fmaf(av, wa.y, acc[1]);
            acc[2] = fmaf(av, wa.z, acc[2]); acc[3] = fmaf(av, wa.w, acc[3]);
            acc[4] = fmaf(av, wb.x, acc[4]); acc[5] = fmaf(av, wb.y, acc[5]);
            acc[6] = fmaf(av, wb.z, acc[6]); acc[7] = fmaf(av, wb.w, acc[7]);
          }
        }
      }
    }
    // ---- cell updates of the thread's 2 units (gate order i,f,g,o), write h
    if (row < B) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int u = u0 + half * 2 + q;
        if (half * 2 + q < a.upc && u < H) {
          if (do0) {
            const float c = sigmoidf_(acc0[q * 4 + 1]) * c0[q] + sigmoidf_(acc0[q * 4 + 0]) * tanhf(acc0[q * 4 + 2]);
            c0[q] = c;
            a.h0buf[(size_t)(p & 1) * B * H + (size_t)row * H + u] = sigmoidf_(acc0[q * 4 + 3]) * tanhf(c);
          }
          if (do1) {
            const float c = sigmoidf_(acc1[q * 4 + 1]) * c1[q] + sigmoidf_(acc1[q * 4 + 0]) * tanhf(acc1[q * 4 + 2]);
            c1[q] = c;
            a.h1all[((size_t)row * Tp + (p - 1)) * H + u] = sigmoidf_(acc1[q * 4 + 3]) * tanhf(c);
          }
        }
      }
    }
    grid_barrier(a.barrier, (unsigned int)(p + 1) * gridDim.x);
  }
}

}  // namespace fb

size_t fb_persistent_smem(int F, int H) {
  return ((size_t)(F + H) * 16 + (size_t)2 * H * 16 + (size_t)fb::ROWS * (fb::KC + 1)) * sizeof(float);
}

bool fb_persistent_supported(int F, int H) {
  static int coop = -1, max_smem = 0;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return coop == 1 && fb_persistent_smem(F, H) <= (size_t)max_smem;
}

// rows [b0, b0+nb) of the batch (nb <= 256); h0buf [2][nb][H], h1all [nb][Tp][H] are for this chunk
int fb_persistent_launch(const fsn_seq_weights* w, const float* magT_chunk, const float* inv1_chunk, float* h0buf,
                         float* h1all_chunk, unsigned int* barrier, int nb, int F, int H, int Tp, cudaStream_t st) {
  fb::Args a;
  for (int l = 0; l < 2; ++l) { a.w_ih[l] = w->w_ih[l]; a.w_hh[l] = w->w_hh[l]; a.b_ih[l] = w->b_ih[l]; a.b_hh[l] = w->b_hh[l]; }
  a.x = magT_chunk; a.inv1 = inv1_chunk; a.h0buf = h0buf; a.h1all = h1all_chunk; a.barrier = barrier;
  a.B = nb; a.F = F; a.H = H; a.Tp = Tp;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  int upc = 1;
  while (upc < fb::MAX_UPC && cdiv(H, upc) > sms) ++upc;
  FSN_REQUIRE(cdiv(H, upc) <= sms, FSN_ERR_UNSUPPORTED, "fb persistent: hidden size %d too large for %d SMs", H, sms);
  a.upc = upc; a.G = cdiv(H, upc);
  const size_t smem = fb_persistent_smem(F, H);
  int rc = check_cuda(cudaFuncSetAttribute(fb::fb_lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "fb_lstm smem attr");
  if (rc) return rc;
  rc = check_cuda(cudaMemsetAsync(barrier, 0, sizeof(unsigned int), st), "fb barrier memset");
  if (rc) return rc;
  void* params[] = {(void*)&a};
  rc = check_cuda(cudaLaunchCooperativeKernel((const void*)fb::fb_lstm_kernel, dim3(a.G), dim3(fb::THREADS), params,
                                              smem, st), "fb_lstm cooperative launch");
  if (rc) return rc;
  FSN_CHECK_LAUNCH("fb_lstm_kernel");
  return FSN_OK;
}

}  // namespace fsn
