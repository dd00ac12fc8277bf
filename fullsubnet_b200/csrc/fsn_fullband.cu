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
#include <type_traits>

#include "fsn_internal.cuh"

namespace fsn {
namespace fb {

constexpr int ROWS = 256;      // clips per launch (host loops over chunks)
constexpr int THREADS = 256;
constexpr int KC = 16;         // k-chunk staged in shared memory
constexpr int RS = KC + 4;      // row stride (floats) of the A tile [row][k]: 80 B keeps 16-byte cp.async aligned and,
                               // with rows rq+8i per thread, makes the 128-bit reads bank-conflict free
constexpr int MAX_UPC = 4;     // hidden units per CTA (=> 16 gate columns)
constexpr int NSTAGE = 5;      // cp.async ring depth of the A tiles

struct Args {
  const float* w_ih[2]; const float* w_hh[2]; const float* b_ih[2]; const float* b_hh[2];
  const float* x;        // magT [B, Tp, F]
  const float* inv1;     // [B]
  float* h0buf;          // [2][B][H] ping-pong
  float* h1all;          // [B][Tp][H]
  unsigned int* barrier; // grid barrier counter (zeroed by the host before launch)
  int B, F, H0, H1, Tp, upc, G;  // layer 0: F -> H0, layer 1: H0 -> H1
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

// 4 rows x 4 gate columns (one hidden unit) per thread and layer: acc[r][g] += a[r] * w[g]
__device__ __forceinline__ void fma_4x4(float (&acc)[16], const float (&av)[4], const float4 w) {
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    acc[r * 4 + 0] = fmaf(av[r], w.x, acc[r * 4 + 0]);
    acc[r * 4 + 1] = fmaf(av[r], w.y, acc[r * 4 + 1]);
    acc[r * 4 + 2] = fmaf(av[r], w.z, acc[r * 4 + 2]);
    acc[r * 4 + 3] = fmaf(av[r], w.w, acc[r * 4 + 3]);
  }
}

// shared-memory weight slice layout: W[layer][k][16] (gate column c = unit_local*4 + gate), k over
// [x | h_prev] of that layer; zero for units beyond H.
// Thread mapping (256 threads): warp w, rq = lane/4 -> rows 32w+rq+8i (i<4), cq = lane%4 -> unit u0+cq (4 gates),
// both layers: a 256x16(x2) register-tiled GEMM per phase, FMA-pipe bound.
__global__ void __launch_bounds__(THREADS, 1) fb_lstm_kernel(const Args a) {
  extern __shared__ __align__(16) float smem_f[];
  const int F = a.F, H0 = a.H0, H1 = a.H1, Tp = a.Tp, B = a.B;
  const int K0 = F + H0, K1 = H0 + H1;
  float* W0 = smem_f;                 // [K0][16]
  float* W1 = W0 + (size_t)K0 * 16;   // [K1][16]
  float* At = W1 + (size_t)K1 * 16;   // [NSTAGE][ROWS][RS]
  const int tid = threadIdx.x;
  const int cq = tid & 3;
  const int row_base = (tid >> 5) * 32 + ((tid & 31) >> 2);  // thread rows: row_base + 8*i
  const int u0 = blockIdx.x * a.upc;  // first hidden unit of this CTA
  const int u = u0 + cq;
  const bool unit_ok0 = cq < a.upc && u < H0, unit_ok1 = cq < a.upc && u < H1;

  // ---- one-time: weight slice -> shared memory
  for (int idx = tid; idx < K0 * 16; idx += THREADS) {
    const int k = idx >> 4, c = idx & 15;
    const int ul = c >> 2, g = c & 3, uu = u0 + ul;
    float w = 0.f;
    if (ul < a.upc && uu < H0) {
      const size_t wr = (size_t)g * H0 + uu;
      w = (k < F) ? a.w_ih[0][wr * F + k] : a.w_hh[0][wr * H0 + (k - F)];
    }
    W0[idx] = w;
  }
  for (int idx = tid; idx < K1 * 16; idx += THREADS) {
    const int k = idx >> 4, c = idx & 15;
    const int ul = c >> 2, g = c & 3, uu = u0 + ul;
    float w = 0.f;
    if (ul < a.upc && uu < H1) {
      const size_t wr = (size_t)g * H1 + uu;
      w = (k < H0) ? a.w_ih[1][wr * H0 + k] : a.w_hh[1][wr * H1 + (k - H0)];
    }
    W1[idx] = w;
  }
  float bias0[4], bias1[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    bias0[g] = unit_ok0 ? a.b_ih[0][g * H0 + u] + a.b_hh[0][g * H0 + u] : 0.f;
    bias1[g] = unit_ok1 ? a.b_ih[1][g * H1 + u] + a.b_hh[1][g * H1 + u] : 0.f;
  }
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};  // cell state: 4 rows, both layers
  float rs[4];  // 1/(mu+1e-5) of the thread's 4 clips (model.py:92), applied to the x segment
#pragma unroll
  for (int r = 0; r < 4; ++r) rs[r] = (row_base + 8 * r < B) ? (a.inv1 ? a.inv1[row_base + 8 * r] : 1.f) : 0.f;

  // A-tile loader: 16-byte cp.async for the h segments (thread -> rows tid/4 + 64 j, k = 4*(tid%4));
  // 4-byte cp.async for the x segment whose rows (F floats) are not 16-byte aligned
  const int l_row0 = (tid >> 5) * 32 + ((tid & 31) >> 4), l_k = tid & 15;
  const int v_row0 = tid >> 2, v_k = (tid & 3) * 4;
  const bool vec_ok1 = (H0 % KC) == 0, vec_ok2 = (H1 % KC) == 0;  // 16-byte cp.async needs aligned, full chunks
  const uint32_t At_s = (uint32_t)__cvta_generic_to_shared(At);
  __syncthreads();

  for (int p = 0; p <= Tp; ++p) {
    const bool do0 = p < Tp;    // layer 0 at step p
    const bool do1 = p >= 1;    // layer 1 at step p-1
    float acc0[16], acc1[16];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int g = 0; g < 4; ++g) { acc0[r * 4 + g] = bias0[g]; acc1[r * 4 + g] = bias1[g]; }
    const float* h0_prev = a.h0buf + (size_t)((p + 1) & 1) * B * H0;      // h0_{p-1}
    const float* h1_prev = a.h1all + (size_t)(p >= 2 ? p - 2 : 0) * H1;   // h1_{p-2}, row stride Tp*H1
    // three k segments: x_p (F, layer 0), h0_{p-1} (H, both layers), h1_{p-2} (H, layer 1), walked as one
    // flat list of KC-wide chunks through a cp.async ring of NSTAGE tiles (NSTAGE-1 chunks in flight hide
    // the L2 latency; out-of-range elements are zero-filled with src-size 0).  Plain (non-.cg) loads are
    // correct here: every phase starts after the acquire of the grid barrier.
    const int nch_f = (F + KC - 1) / KC, nch_h = (H0 + KC - 1) / KC, nch_h1 = (H1 + KC - 1) / KC;
    const int c_begin = do0 ? 0 : nch_f;                       // skip x when layer 0 is finished
    const int c_end = nch_f + (p >= 1 ? nch_h : 0) + (p >= 2 ? nch_h1 : 0);
    auto issue = [&](int ci) {
      if (ci < c_end) {
        int seg = 0, k0 = ci * KC;
        if (ci >= nch_f) { seg = 1; k0 = (ci - nch_f) * KC; }
        if (ci >= nch_f + nch_h) { seg = 2; k0 = (ci - nch_f - nch_h) * KC; }
        const uint32_t Ab = At_s + (uint32_t)((ci - c_begin) % NSTAGE) * (ROWS * RS * 4);
        if ((seg == 1 && vec_ok1) || (seg == 2 && vec_ok2)) {
          const float* base = (seg == 1) ? h0_prev + k0 : h1_prev + k0;
          const unsigned rstride = (seg == 1) ? (unsigned)H0 : (unsigned)(Tp * H1);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int r = v_row0 + 64 * j;
            const bool ok = r < B;  // H % KC == 0: these chunks are always full
            const float* src = ok ? base + (size_t)r * rstride + v_k : a.x;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(Ab + (uint32_t)((r * RS + v_k) * 4)), "l"(src),
                         "r"(ok ? 16 : 0)
                         : "memory");
          }
        } else {
          const int klen = (seg == 0) ? F : ((seg == 1) ? H0 : H1);
          const int k = k0 + l_k;
          const bool kok = k < klen;
          const float* base = (seg == 0) ? a.x + (size_t)p * F + k : ((seg == 1) ? h0_prev + k : h1_prev + k);
          const size_t rstride = (seg == 0) ? (size_t)Tp * F : ((seg == 1) ? (size_t)H0 : (size_t)Tp * H1);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int r = l_row0 + 2 * j;
            const bool ok = kok && r < B;
            const float* src = ok ? base + (size_t)r * rstride : a.x;  // any valid address when zero-filled
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(Ab + (uint32_t)((r * RS + l_k) * 4)), "l"(src),
                         "r"(ok ? 4 : 0)
                         : "memory");
          }
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");  // (empty groups keep the wait count uniform)
    };
#pragma unroll 1
    for (int i = 0; i < NSTAGE - 1; ++i) issue(c_begin + i);
    for (int ci = c_begin; ci < c_end; ++ci) {
      asm volatile("cp.async.wait_group %0;" ::"n"(NSTAGE - 2) : "memory");
      __syncthreads();                 // chunk ci landed for everyone; buffer of chunk ci-1 is free
      issue(ci + NSTAGE - 1);
      const float* Ab = At + (size_t)((ci - c_begin) % NSTAGE) * ROWS * RS;
      if (row_base >= B) continue;
      int seg = 0, k0 = ci * KC;
      if (ci >= nch_f) { seg = 1; k0 = (ci - nch_f) * KC; }
      if (ci >= nch_f + nch_h) { seg = 2; k0 = (ci - nch_f - nch_h) * KC; }
      const float* w0 = ((seg == 0) ? W0 : W0 + (size_t)F * 16) + (size_t)k0 * 16 + cq * 4;   // layer-0 rows
      const float* w1 = ((seg == 1) ? W1 : W1 + (size_t)H0 * 16) + (size_t)k0 * 16 + cq * 4;  // layer-1 rows
      const bool use0 = (seg <= 1) && do0, use1 = (seg >= 1) && do1;
      const float* ar = Ab + row_base * RS;
      // rows beyond klen inside the chunk are zero-filled, so the full KC is always safe to consume; the three
      // segment kinds get their own straight-line code (branch once per chunk, not per FMA group)
      auto consume = [&](auto use0_c, auto use1_c, auto scale_c) {
        constexpr bool U0 = decltype(use0_c)::value, U1 = decltype(use1_c)::value, SC = decltype(scale_c)::value;
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
          float4 a4[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) a4[r] = *reinterpret_cast<const float4*>(ar + (8 * r) * RS + k4);
          if (SC) {
#pragma unroll
            for (int r = 0; r < 4; ++r) { a4[r].x *= rs[r]; a4[r].y *= rs[r]; a4[r].z *= rs[r]; a4[r].w *= rs[r]; }
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float av[4] = {q == 0 ? a4[0].x : q == 1 ? a4[0].y : q == 2 ? a4[0].z : a4[0].w,
                                 q == 0 ? a4[1].x : q == 1 ? a4[1].y : q == 2 ? a4[1].z : a4[1].w,
                                 q == 0 ? a4[2].x : q == 1 ? a4[2].y : q == 2 ? a4[2].z : a4[2].w,
                                 q == 0 ? a4[3].x : q == 1 ? a4[3].y : q == 2 ? a4[3].z : a4[3].w};
            if (U0) fma_4x4(acc0, av, *reinterpret_cast<const float4*>(w0 + (k4 + q) * 16));
            if (U1) fma_4x4(acc1, av, *reinterpret_cast<const float4*>(w1 + (k4 + q) * 16));
          }
        }
      };
      using T_ = std::true_type; using F_ = std::false_type;
      if (seg == 0) { if (use0) consume(T_{}, F_{}, T_{}); }
      else if (use0 && use1) consume(T_{}, T_{}, F_{});
      else if (use0) consume(T_{}, F_{}, F_{});
      else if (use1) consume(F_{}, T_{}, F_{});
    }
    // ---- cell updates of the thread's unit for its 4 rows (gate order i,f,g,o), write h
    if (unit_ok0 || unit_ok1) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int row = row_base + 8 * r;
        if (row >= B) continue;
        if (do0 && unit_ok0) {
          const float c = sigmoidf_(acc0[r * 4 + 1]) * c0[r] + sigmoidf_(acc0[r * 4 + 0]) * tanhf(acc0[r * 4 + 2]);
          c0[r] = c;
          a.h0buf[(size_t)(p & 1) * B * H0 + (size_t)row * H0 + u] = sigmoidf_(acc0[r * 4 + 3]) * tanhf(c);
        }
        if (do1 && unit_ok1) {
          const float c = sigmoidf_(acc1[r * 4 + 1]) * c1[r] + sigmoidf_(acc1[r * 4 + 0]) * tanhf(acc1[r * 4 + 2]);
          c1[r] = c;
          a.h1all[((size_t)row * Tp + (p - 1)) * H1 + u] = sigmoidf_(acc1[r * 4 + 3]) * tanhf(c);
        }
      }
    }
    grid_barrier(a.barrier, (unsigned int)(p + 1) * gridDim.x);
  }
}

}  // namespace fb

size_t fb_persistent_smem(int F, int H0, int H1) {
  return ((size_t)(F + H0) * 16 + (size_t)(H0 + H1) * 16 + (size_t)fb::NSTAGE * fb::ROWS * fb::RS) * sizeof(float);
}

bool fb_persistent_supported(int F, int H0, int H1) {
  static int coop = -1, max_smem = 0, sms = 148;
  if (coop < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const int Hm = H0 > H1 ? H0 : H1;
  return coop == 1 && fb_persistent_smem(F, H0, H1) <= (size_t)max_smem && cdiv(Hm, fb::MAX_UPC) <= sms;
}

// One persistent launch of a 2-layer LSTM wavefront (layer 0: F -> H0 with optional per-row input scale `inv1`,
// layer 1: H0 -> H1) for rows [0, nb), nb <= 256; h0buf [2][nb][H0] scratch, h1all [nb][Tp][H1] output.
int fb_persistent_launch(const fsn_seq_weights* w, const float* x_chunk, const float* inv1_chunk, float* h0buf,
                         float* h1all_chunk, unsigned int* barrier, int nb, int F, int H0, int H1, int Tp,
                         cudaStream_t st) {
  fb::Args a;
  for (int l = 0; l < 2; ++l) { a.w_ih[l] = w->w_ih[l]; a.w_hh[l] = w->w_hh[l]; a.b_ih[l] = w->b_ih[l]; a.b_hh[l] = w->b_hh[l]; }
  a.x = x_chunk; a.inv1 = inv1_chunk; a.h0buf = h0buf; a.h1all = h1all_chunk; a.barrier = barrier;
  a.B = nb; a.F = F; a.H0 = H0; a.H1 = H1; a.Tp = Tp;
  int sms = 148;
  { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
  const int Hm = H0 > H1 ? H0 : H1;
  int upc = 1;
  while (upc < fb::MAX_UPC && cdiv(Hm, upc) > sms) ++upc;
  FSN_REQUIRE(cdiv(Hm, upc) <= sms, FSN_ERR_UNSUPPORTED, "persistent LSTM: hidden size %d too large for %d SMs", Hm, sms);
  a.upc = upc; a.G = cdiv(Hm, upc);
  const size_t smem = fb_persistent_smem(F, H0, H1);
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
