// One LSTM layer for a SMALL batch of sequences (rows = clips) on the tensor cores (tcgen05, sm_100a):
//   1. the input projection of ALL steps hoisted into one GEMM  P[r,t,:] = x[r,t,:] W_ih^T   (tgemm_tma_kernel, tf32;
//      three passes on tf32 hi/lo splits for the fp32 error class), and
//   2. the recurrence  gates_t = P_t + b + h_{t-1} W_hh^T  as ONE persistent cooperative kernel (this file).
//
// Reference semantics: audio_zen/model/module/sequence_model.py:52-58,117 (nn.LSTM, gate order i,f,g,o, zero initial
// state) as used for the full-band stacks (recipes/dns_interspeech_2020/fullsubnet/model.py:43-51,92-95; rows = clips).
//
// Mapping of the recurrence.  The batch is small (<= 128 rows per group) and the steps are serial, so the HIDDEN
// dimension is spread over the chip: CTA j of a group owns 8 hidden units = 32 gate columns and keeps that slice of
// W_hh resident in shared memory for the whole sequence (fp16, UMMA K-major 128B-swizzled; compensated mode: hi and lo
// parts).  Per step every CTA needs ALL of h_{t-1}: the epilogues publish h_t as fp16 (hi [+ lo]) in a ping-pong array
// in global memory (it lives in L2), a per-group counter is the step barrier, and each CTA's producer warp streams the
// [128 rows x K] state through a TMA ring.  Rows sit on the M side of the MMA:
//     D[128 rows, 32 | 64 gate columns] += h_hi[128, 16] . [W_hi ; W_lo]^T   (one MMA, N = 64: the lo product lands in
//     D[128 rows, 32]                    += h_lo[128, 16] . W_hi^T             columns 32..63 and is added in the epilogue)
// so the thread that owns TMEM lane r sees all four gates of the CTA's 8 units for row r: c stays in 8 registers, no
// exchange.  Two groups of 64 CTAs cover 256 clips with H = 512.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>
#include <stdlib.h>

#include "fsn_internal.cuh"
#include "fsn_tc_ptx.cuh"

namespace fsn {
namespace rec {
using namespace ptx;

constexpr int MR = 128;               // rows per group (MMA M)
constexpr int U = 8;                  // hidden units per CTA
constexpr int NG = 4 * U;             // gate columns per CTA
constexpr int KB = 64;                // k-block: one 128-byte swizzled row of fp16
constexpr int A_TILE = MR * KB * 2;   // 16 KB
constexpr int NTHREADS = 192;         // warp 0: TMA producer, warp 1: MMA issue + TMEM, warps 2-5: epilogue
constexpr int MAX_STAGES = 6;

struct Bars {
  uint64_t full[MAX_STAGES], empty[MAX_STAGES];
  uint64_t acc_full, acc_empty;
  uint32_t tmem_base;
};

struct Args {
  const float* w_hh; const float* b_ih; const float* b_hh;   // [4H,H], [4H], [4H] (PyTorch layout)
  const float* P; size_t p_row, p_t;                          // P[r*p_row + t*p_t + gate*H + u]
  float* hall; size_t h_row, h_t;                             // hall[r*h_row + t*h_t + u]
  __half* state;                                              // [2 parity][PARTS][Rpad][Kp]
  unsigned int* barrier;                                      // one counter per group, 32 words apart
  int R, T, H, Kp, Rpad, C, stages, fence_all;
};

__device__ __forceinline__ void tc_mma1_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <bool X3> __device__ __forceinline__ float act_sigmoid(float x) {
  return X3 ? 1.0f / (1.0f + expf(-x)) : fast_sigmoid(x);
}
template <bool X3> __device__ __forceinline__ float act_tanh(float x) {
  return X3 ? 1.0f - 2.0f / (1.0f + expf(2.0f * x)) : fast_tanh(x);
}

template <bool X3>
__global__ void __launch_bounds__(NTHREADS, 1) lstm_rec_tc_kernel(const __grid_constant__ CUtensorMap tmap, const Args a) {
  constexpr int PARTS = X3 ? 2 : 1;
  constexpr int STAGE_BYTES = PARTS * A_TILE;
  constexpr int WB_ROWS = PARTS * NG;          // rows of the resident weight operand per k-block: [W_hi ; W_lo]
  constexpr int WB_BYTES = WB_ROWS * 128;
  constexpr uint32_t kIdescMain = (1u << 4) | ((uint32_t)(WB_ROWS >> 3) << 17) | ((128u >> 4) << 24);
  constexpr uint32_t kIdescLo = (1u << 4) | ((uint32_t)(NG >> 3) << 17) | ((128u >> 4) << 24);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int nkb = a.Kp / KB;
  uint8_t* ring = smem;
  uint8_t* wsm = smem + (size_t)a.stages * STAGE_BYTES;
  Bars& bars = *reinterpret_cast<Bars*>(wsm + (size_t)nkb * WB_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = blockIdx.x / a.C, j = blockIdx.x - g * a.C;   // group, unit slice
  const int u0 = j * U;
  const int H = a.H, T = a.T;
  unsigned int* counter = a.barrier + g * 32;

  // ---------------- one-time setup: barriers, TMEM, resident weight slice
  if (threadIdx.x == 0) {
    for (int s = 0; s < a.stages; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    mbar_init(&bars.acc_empty, 128);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&bars.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // weight operand: k-block kb = rows [W_hi (gate col n = gate*8 + ul) ; W_lo] x 64 k, 128B swizzle
  for (int idx = threadIdx.x; idx < nkb * NG * (KB / 8); idx += NTHREADS) {
    const int kb = idx / (NG * (KB / 8));
    const int rem = idx - kb * (NG * (KB / 8));
    const int n = rem / (KB / 8), ch = rem - n * (KB / 8);
    const int gate = n / U, u = u0 + (n % U);
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = kb * KB + ch * 8 + e;
      const float w = (u < H && k < H) ? a.w_hh[((size_t)gate * H + u) * H + k] : 0.f;
      hi[e] = __float2half_rn(w);
      lo[e] = __float2half_rn(w - __half2float(hi[e]));
    }
    uint8_t* blk = wsm + (size_t)kb * WB_BYTES;
    *reinterpret_cast<uint4*>(blk + swz128_off(n, ch * 8)) = *reinterpret_cast<const uint4*>(hi);
    if (X3) *reinterpret_cast<uint4*>(blk + swz128_off(NG + n, ch * 8)) = *reinterpret_cast<const uint4*>(lo);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp == 0) {
    // ================= producer: after the group has published h_{p-1}, stream it through the ring
    uint32_t it = 0;
    for (int p = 1; p < T; ++p) {
      if (lane == 0) {
        const unsigned int target = (unsigned int)a.C * (unsigned int)p;
        unsigned int v, spins = 0;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
          if (++spins > (1u << 27)) { printf("fsn rec: step barrier timeout (block %d step %d)\n", blockIdx.x, p); __trap(); }
        } while (v < target);
      }
      __syncwarp();
      fence_proxy_async();  // the TMA (async proxy) reads below come after the acquire above
      const int par = (p - 1) & 1;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)a.stages, use = it / (uint32_t)a.stages;
        mbar_wait<false>(&bars.empty[s], (use & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&bars.full[s], STAGE_BYTES);
#pragma unroll
          for (int part = 0; part < PARTS; ++part)
            tma_load_2d(ring + (size_t)s * STAGE_BYTES + part * A_TILE, &tmap, kb * KB,
                        (par * PARTS + part) * a.Rpad + g * MR, &bars.full[s]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issue (converged warp, one elected lane)
    uint32_t it = 0;
    const uint64_t wdesc0 = desc_sw128(smem_u32(wsm));
    const uint64_t adesc0 = desc_sw128(smem_u32(ring));
    for (int p = 1; p < T; ++p) {
      mbar_wait<false>(&bars.acc_empty, (p - 1) & 1);  // the epilogue of step p-1 has drained the accumulator
      tc_fence_after();
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const uint32_t s = it % (uint32_t)a.stages, use = it / (uint32_t)a.stages;
        mbar_wait<false>(&bars.full[s], use & 1);
        tc_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)((s * STAGE_BYTES) >> 4);
          const uint64_t wd = wdesc0 + (uint64_t)(((uint32_t)kb * WB_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < KB / 16; ++k) {
            tc_mma1_f16(tmem_base, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), kIdescMain, (kb | k) ? 1u : 0u);
            if (X3) tc_mma1_f16(tmem_base, ad + (uint64_t)((A_TILE >> 4) + 2 * k), wd + (uint64_t)(2 * k), kIdescLo, 1u);
          }
          tc_commit1(&bars.empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit1(&bars.acc_full);
      __syncwarp();
    }
  } else {
    // ================= epilogue: thread = row (TMEM lane), 8 units x 4 gates in its columns
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int r = g * MR + row;
    const bool valid = r < a.R;
    const bool vec_ok = (H & 3) == 0 && (a.p_row & 3) == 0 && (a.p_t & 3) == 0 && (a.h_row & 3) == 0 && (a.h_t & 3) == 0;
    float bias[NG];
#pragma unroll
    for (int n = 0; n < NG; ++n) {
      const int gate = n / U, u = u0 + (n % U);
      bias[n] = (u < H) ? a.b_ih[gate * H + u] + a.b_hh[gate * H + u] : 0.f;
    }
    float c[U];
#pragma unroll
    for (int i = 0; i < U; ++i) c[i] = 0.f;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    const int nu = (H - u0 < U) ? (H - u0) : U;  // real units of this CTA
    for (int p = 0; p < T; ++p) {
      float pre[NG];
      if (valid) {  // input projection of this step (issued before the wait below: the latency hides behind the MMAs)
        const float* pp = a.P + (size_t)r * a.p_row + (size_t)p * a.p_t + u0;
        if (vec_ok && nu == U) {
#pragma unroll
          for (int gate = 0; gate < 4; ++gate) {
            const float4 v0 = __ldg(reinterpret_cast<const float4*>(pp + (size_t)gate * H));
            const float4 v1 = __ldg(reinterpret_cast<const float4*>(pp + (size_t)gate * H + 4));
            pre[gate * U + 0] = v0.x; pre[gate * U + 1] = v0.y; pre[gate * U + 2] = v0.z; pre[gate * U + 3] = v0.w;
            pre[gate * U + 4] = v1.x; pre[gate * U + 5] = v1.y; pre[gate * U + 6] = v1.z; pre[gate * U + 7] = v1.w;
          }
        } else {
#pragma unroll
          for (int n = 0; n < NG; ++n) pre[n] = ((n % U) < nu) ? __ldg(pp + (size_t)(n / U) * H + (n % U)) : 0.f;
        }
      } else {
#pragma unroll
        for (int n = 0; n < NG; ++n) pre[n] = 0.f;
      }
      if (p > 0) {
        mbar_wait<true>(&bars.acc_full, (p - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < NG; cb += 8) {
          float d[8];
          tc_ld8(taddr + cb, d);
          tc_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) pre[cb + e] += d[e];
          if (X3) {
            tc_ld8(taddr + NG + cb, d);
            tc_wait_ld();
#pragma unroll
            for (int e = 0; e < 8; ++e) pre[cb + e] += d[e];
          }
        }
        tc_fence_before();
      }
      mbar_arrive(&bars.acc_empty);
      float h[U];
#pragma unroll
      for (int i = 0; i < U; ++i) {
        const float gi = pre[0 * U + i] + bias[0 * U + i], gf = pre[1 * U + i] + bias[1 * U + i];
        const float gg = pre[2 * U + i] + bias[2 * U + i], go = pre[3 * U + i] + bias[3 * U + i];
        const float cn = act_sigmoid<X3>(gf) * c[i] + act_sigmoid<X3>(gi) * act_tanh<X3>(gg);
        c[i] = cn;
        h[i] = act_sigmoid<X3>(go) * act_tanh<X3>(cn);
      }
      if (valid) {
        float* hp = a.hall + (size_t)r * a.h_row + (size_t)p * a.h_t + u0;
        if (vec_ok && nu == U) {
          *reinterpret_cast<float4*>(hp) = make_float4(h[0], h[1], h[2], h[3]);
          *reinterpret_cast<float4*>(hp + 4) = make_float4(h[4], h[5], h[6], h[7]);
        } else {
#pragma unroll
          for (int i = 0; i < U; ++i) if (i < nu) hp[i] = h[i];
        }
        if (p + 1 < T) {  // publish h_p for the next step: fp16 hi (and lo), 16 bytes each
          __half hi[U], lo[U];
#pragma unroll
          for (int i = 0; i < U; ++i) {
            const float v = (i < nu) ? h[i] : 0.f;
            hi[i] = __float2half_rn(v);
            lo[i] = __float2half_rn(v - __half2float(hi[i]));
          }
          __half* sp = a.state + ((size_t)((p & 1) * PARTS) * a.Rpad + r) * a.Kp + u0;
          *reinterpret_cast<uint4*>(sp) = *reinterpret_cast<const uint4*>(hi);
          if (X3) *reinterpret_cast<uint4*>(sp + (size_t)a.Rpad * a.Kp) = *reinterpret_cast<const uint4*>(lo);
        }
      }
      if (p + 1 < T) {
        // publish h_p: CTA-scope barrier over the 128 writers, then ONE gpu-scope fence (cumulative over the writes
        // observed through the barrier - the grid.sync pattern) + async-proxy fence + release by the arriving thread.
        // FSN_REC_FENCE_ALL=1 (debug) makes every writer fence for itself.
        if (a.fence_all) { __threadfence(); fence_proxy_async(); }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          __threadfence();
          fence_proxy_async();
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
        }
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem_base));
  }
}

// Operand preparation for the tf32 GEMM.  v = in[r, k] * scale.
//   cat == 0: out[r, :Kp] = v (zero padded to Kp): plain scaled, 16-byte aligned copy, single-pass GEMM.
//   cat != 0: compensated GEMM as ONE pass over a 3x longer K: hi = tf32-truncated v (what the tensor core reads),
//             lo = v - hi (exact); the A operand (cat == 1) is laid out [hi | lo | hi], the B operand (cat == 2)
//             [hi | hi | lo], so that  A' B'^T = hi.hi + lo.hi + hi.lo  - the product to ~2^-21, one output write.
// row_scale index of row r: r / rows_per_scale (per clip), or with scale_B > 0 the time-major entry
// (r % rows_per_scale) * scale_B + r / rows_per_scale of a [T', B] table (cumulative norm).
__global__ void split_tf32_kernel(const float* __restrict__ in, size_t rows, int K, size_t ldi, const float* __restrict__ row_scale,
                                  int rows_per_scale, int scale_B, float* __restrict__ out, int Kp, int cat) {
  const size_t n = rows * (size_t)Kp;
  const size_t ldo = cat ? (size_t)3 * Kp : (size_t)Kp;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / Kp;
    const int k = (int)(i - r * Kp);
    float v = 0.f;
    if (k < K) {
      v = in[r * ldi + k];
      if (row_scale)
        v *= scale_B > 0 ? row_scale[(r % rows_per_scale) * scale_B + r / rows_per_scale] : row_scale[r / rows_per_scale];
    }
    float* o = out + r * ldo + k;
    if (cat) {
      const float hi = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
      const float lo = v - hi;
      o[0] = hi;
      o[Kp] = (cat == 1) ? lo : hi;
      o[2 * Kp] = (cat == 1) ? hi : lo;
    } else {
      o[0] = v;
    }
  }
}

// x[r, :N] = act(x[r, :N] + bias[:N]) in place, row stride ld
__global__ void bias_act_kernel(float* __restrict__ x, size_t rows, int N, size_t ld, const float* __restrict__ bias, int act) {
  const size_t n = rows * (size_t)N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / N;
    const int c = (int)(i - r * N);
    float v = x[r * ld + c] + (bias ? bias[c] : 0.f);
    switch (act) {
      case FSN_ACT_RELU: v = fmaxf(v, 0.f); break;
      case FSN_ACT_TANH: v = tanhf(v); break;
      case FSN_ACT_RELU6: v = fminf(fmaxf(v, 0.f), 6.f); break;
      default: break;
    }
    x[r * ld + c] = v;
  }
}

static PFN_cuTensorMapEncodeTiled_v12000 encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    cudaGetLastError();
  }
  return fn;
}

static size_t smem_bytes(int Kp, bool x3, int stages) {
  const int parts = x3 ? 2 : 1;
  return (size_t)stages * parts * A_TILE + (size_t)(Kp / KB) * parts * NG * 128 + sizeof(Bars) + 1024;
}

}  // namespace rec

static int rec_sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// tensor-core recurrence available for hidden size H on this device?
bool lstm_rec_tc_supported(int H, bool x3) {
  static const bool off = getenv("FSN_NO_REC_TC") != nullptr;
  if (off || H < 64 || !rec::encoder()) return false;  // H >= 64: see lstm_rec_tc_scratch_bytes
  int dev = 0, coop = 0, max_smem = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  const int Kp = (H + rec::KB - 1) / rec::KB * rec::KB;
  return coop == 1 && major == 10 && cdiv(H, rec::U) <= rec_sm_count() && rec::smem_bytes(Kp, x3, 2) <= (size_t)max_smem;
}

// rows one launch covers (groups of 128 that fit on the chip next to each other)
int lstm_rec_tc_rows_per_launch(int H) {
  const int C = cdiv(H, rec::U);
  int G = rec_sm_count() / C;
  if (G < 1) G = 1;
  return G * rec::MR;
}

// scratch of one launch: fp16 state ping-pong [2][parts][rows][Kp] + the group counters.  The same scratch serves
// layers of DIFFERENT hidden sizes (fast_fullsubnet: 384 / 257 / 512), so its size must not depend on H:
// rows_per_launch * Kp <= (SMs * 8 / H) * 128 * (H + 63) < 2 * SMs * 8 * 128 elements for H >= 64.
static size_t rec_state_capacity_bytes() { return align_up((size_t)2 * 2 * 2 * rec_sm_count() * 8 * 128 * sizeof(__half), 256); }
size_t lstm_rec_tc_scratch_bytes(int H, bool x3) {
  (void)H; (void)x3;
  return rec_state_capacity_bytes() + 64 * 32 * sizeof(unsigned int);
}

// h_t for every step of one layer, given the hoisted input projection P (see Args for the strides); R rows (any
// count: chunks of lstm_rec_tc_rows_per_launch are launched back to back)
int lstm_rec_tc_launch(const float* w_hh, const float* b_ih, const float* b_hh, const float* P, size_t p_row, size_t p_t,
                       float* hall, size_t h_row, size_t h_t, int R, int T, int H, bool x3, void* scratch,
                       cudaStream_t st) {
  FSN_REQUIRE(lstm_rec_tc_supported(H, x3), FSN_ERR_UNSUPPORTED, "lstm_rec_tc: hidden size %d not supported", H);
  const int Kp = (H + rec::KB - 1) / rec::KB * rec::KB;
  const int C = cdiv(H, rec::U);
  const int rows_max = lstm_rec_tc_rows_per_launch(H);
  const int parts = x3 ? 2 : 1;
  int dev = 0, max_smem = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  int stages = x3 ? 4 : rec::MAX_STAGES;
  while (stages > 2 && rec::smem_bytes(Kp, x3, stages) > (size_t)max_smem) --stages;
  const size_t smem = rec::smem_bytes(Kp, x3, stages);
  __half* state = (__half*)scratch;
  const size_t state_bytes = align_up((size_t)2 * parts * rows_max * Kp * sizeof(__half), 256);
  FSN_REQUIRE(state_bytes <= rec_state_capacity_bytes(), FSN_ERR_UNSUPPORTED, "lstm_rec_tc: state of H=%d exceeds the scratch", H);
  unsigned int* barrier = (unsigned int*)((uint8_t*)scratch + rec_state_capacity_bytes());  // fixed place, whatever H
  int rc;
  const void* kern = x3 ? (const void*)rec::lstm_rec_tc_kernel<true> : (const void*)rec::lstm_rec_tc_kernel<false>;
  if ((rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "lstm_rec_tc smem attr")))
    return rc;
  for (int r0 = 0; r0 < R; r0 += rows_max) {
    const int nr = (R - r0 < rows_max) ? R - r0 : rows_max;
    const int G = cdiv(nr, rec::MR);
    const int Rpad = G * rec::MR;
    // state of padded rows / padded k stays zero for the whole launch; counters start at zero
    if ((rc = check_cuda(cudaMemsetAsync(scratch, 0, state_bytes, st), "lstm_rec_tc memset"))) return rc;
    if ((rc = check_cuda(cudaMemsetAsync(barrier, 0, 64 * 32 * sizeof(unsigned int), st), "lstm_rec_tc memset"))) return rc;
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)(2 * parts * Rpad)};
    cuuint64_t gstr[1] = {(cuuint64_t)Kp * sizeof(__half)};
    cuuint32_t box[2] = {(cuuint32_t)rec::KB, (cuuint32_t)rec::MR};
    cuuint32_t estr[2] = {1, 1};
    FSN_REQUIRE(rec::encoder()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)state, gdim, gstr, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS,
                FSN_ERR_CUDA, "lstm_rec_tc: cuTensorMapEncodeTiled failed");
    rec::Args a;
    a.w_hh = w_hh; a.b_ih = b_ih; a.b_hh = b_hh;
    a.P = P + (size_t)r0 * p_row; a.p_row = p_row; a.p_t = p_t;
    a.hall = hall + (size_t)r0 * h_row; a.h_row = h_row; a.h_t = h_t;
    a.state = state; a.barrier = barrier;
    a.R = nr; a.T = T; a.H = H; a.Kp = Kp; a.Rpad = Rpad; a.C = C; a.stages = stages;
    a.fence_all = getenv("FSN_REC_FENCE_ALL") != nullptr;
    void* params[] = {(void*)&tm, (void*)&a};
    if ((rc = check_cuda(cudaLaunchCooperativeKernel(kern, dim3(G * C), dim3(rec::NTHREADS), params, smem, st),
                         "lstm_rec_tc cooperative launch")))
      return rc;
    FSN_CHECK_LAUNCH("lstm_rec_tc_kernel");
  }
  return FSN_OK;
}

int split_tf32_launch(const float* in, size_t rows, int K, size_t ldi, const float* row_scale, int rows_per_scale,
                      float* out, int Kp, int cat, cudaStream_t st, int scale_B) {
  if (rows == 0) return FSN_OK;
  const size_t n = rows * (size_t)Kp;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  rec::split_tf32_kernel<<<(int)blocks, 256, 0, st>>>(in, rows, K, ldi, row_scale, rows_per_scale < 1 ? 1 : rows_per_scale,
                                                      scale_B, out, Kp, cat);
  FSN_CHECK_LAUNCH("split_tf32_kernel");
  return FSN_OK;
}

int bias_act_launch(float* x, size_t rows, int N, size_t ld, const float* bias, int act, cudaStream_t st) {
  if (rows == 0) return FSN_OK;
  const size_t n = rows * (size_t)N;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  rec::bias_act_kernel<<<(int)blocks, 256, 0, st>>>(x, rows, N, ld, bias, act);
  FSN_CHECK_LAUNCH("bias_act_kernel");
  return FSN_OK;
}

// C[M,N] = A[M,K] W[N,K]^T on the tf32 tensor cores.  `a` is the prepared A operand (prep_operand: [M, Keff], lda);
// W row-major [N, K] is prepared here into `w` ([N, Keff]); x3: Keff = 3 * Kp (see split_tf32_kernel).
int gemm_tc_split_launch(const float* a, size_t lda, const float* W, int N, int K, float* w, float* C, size_t ldc, size_t M,
                         bool x3, cudaStream_t st) {
  const int Kp = (K + 3) & ~3;
  const int Keff = x3 ? 3 * Kp : K;
  int rc;
  if ((rc = split_tf32_launch(W, (size_t)N, K, (size_t)K, nullptr, 1, w, Kp, x3 ? 2 : 0, st, 0))) return rc;
  FSN_REQUIRE(M < ((size_t)1 << 31), FSN_ERR_SHAPE, "gemm_tc: too many rows");
  return tgemm_launch(a, lda, w, x3 ? (size_t)3 * Kp : (size_t)Kp, C, ldc, (int)M, N, Keff, false, nullptr, 0, st);
}

// ---------------------------------------------------------------------------------------------------------------
// One full LSTM layer and one Linear layer on top of the pieces above (what the model files call).
void lstm_tc_carve(char* base, size_t& off, size_t rows_T, int Kmax, int Hmax, bool x3, LstmTcWs& ws) {
  auto take = [&](size_t bytes) { char* r = base ? base + off : nullptr; off = align_up(off + bytes, 256); return r; };
  const size_t wa = (size_t)((Kmax + 3) & ~3) * (x3 ? 3 : 1);
  ws.a = (float*)take(rows_T * wa * sizeof(float));
  ws.w = (float*)take((size_t)4 * Hmax * wa * sizeof(float));
  ws.P = (float*)take(rows_T * 4 * Hmax * sizeof(float));
  ws.rec = take(lstm_rec_tc_scratch_bytes(Hmax, x3));
}

// operand A of a GEMM: x [rows, K] (row stride ldx) -> 16-byte aligned rows, scaled; x3: the [hi | lo | hi] layout
static int prep_operand(const float* x, size_t ldx, int K, size_t rows, const float* row_scale, int rows_per_scale, int scale_B,
                        bool x3, const LstmTcWs& ws, const float*& a, size_t& lda, cudaStream_t st) {
  if (!x3 && !row_scale && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    a = x; lda = ldx;
    return FSN_OK;
  }
  const int Kp = (K + 3) & ~3;
  a = ws.a; lda = (size_t)Kp * (x3 ? 3 : 1);
  return split_tf32_launch(x, rows, K, ldx, row_scale, rows_per_scale, ws.a, Kp, x3 ? 1 : 0, st, scale_B);
}

// hall[r, t, :] (row stride T*H) of one LSTM layer over x[(r*T + t), :K] * scale
int lstm_layer_tc(const fsn_lstm_layer& L, const float* x, size_t ldx, int K, const float* row_scale, int rows_per_scale,
                  int scale_B, int R, int T, int H, bool x3, const LstmTcWs& ws, float* hall, cudaStream_t st) {
  const size_t rows = (size_t)R * T;
  const float* a;
  size_t lda;
  int rc;
  if ((rc = prep_operand(x, ldx, K, rows, row_scale, rows_per_scale, scale_B, x3, ws, a, lda, st))) return rc;
  if ((rc = gemm_tc_split_launch(a, lda, L.w_ih, 4 * H, K, ws.w, ws.P, (size_t)4 * H, rows, x3, st))) return rc;
  return lstm_rec_tc_launch(L.w_hh, L.b_ih, L.b_hh, ws.P, (size_t)T * 4 * H, (size_t)4 * H, hall, (size_t)T * H, (size_t)H, R, T,
                            H, x3, ws.rec, st);
}

// out[rows, :N] (row stride ldo) = act(x[rows, :K] W[N,K]^T + bias)
int linear_tc(const float* x, size_t ldx, int K, const float* W, const float* bias, int N, int act, float* out, size_t ldo,
              size_t rows, bool x3, const LstmTcWs& ws, cudaStream_t st) {
  const float* a;
  size_t lda;
  int rc;
  if ((rc = prep_operand(x, ldx, K, rows, nullptr, 1, 0, x3, ws, a, lda, st))) return rc;
  if ((rc = gemm_tc_split_launch(a, lda, W, N, K, ws.w, out, ldo, rows, x3, st))) return rc;
  if (!bias && act == FSN_ACT_NONE) return FSN_OK;
  return bias_act_launch(out, rows, N, ldo, bias, act, st);
}

}  // namespace fsn

// unit-test hooks (tests/test_gpu_rec_tc.py): one LSTM layer / one Linear layer on the tensor-core path
extern "C" size_t fsn_debug_lstm_tc_workspace_bytes(int R, int T, int K, int H, int x3) {
  size_t off = 0;
  fsn::LstmTcWs ws;
  fsn::lstm_tc_carve(nullptr, off, (size_t)R * T, K > H ? K : H, H, x3 != 0, ws);
  return off;
}
extern "C" int fsn_debug_lstm_layer_tc(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                       const float* x, int R, int T, int K, int H, int x3, float* hall, void* workspace,
                                       size_t workspace_bytes, fsn_stream_t stream) {
  FSN_REQUIRE(fsn::lstm_rec_tc_supported(H, x3 != 0), FSN_ERR_UNSUPPORTED, "lstm_layer_tc: hidden size %d not supported", H);
  size_t off = 0;
  fsn::LstmTcWs ws;
  fsn::lstm_tc_carve((char*)workspace, off, (size_t)R * T, K > H ? K : H, H, x3 != 0, ws);
  FSN_REQUIRE(workspace && workspace_bytes >= off, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, off);
  fsn_lstm_layer L{w_ih, w_hh, b_ih, b_hh};
  return fsn::lstm_layer_tc(L, x, (size_t)K, K, nullptr, 1, 0, R, T, H, x3 != 0, ws, hall, (cudaStream_t)stream);
}
extern "C" int fsn_debug_linear_tc(const float* x, int rows, int K, const float* W, const float* bias, int N, int act, int x3,
                                   float* out, void* workspace, size_t workspace_bytes, fsn_stream_t stream) {
  size_t off = 0;
  fsn::LstmTcWs ws;
  const int Hm = (N + 3) / 4 > 8 ? (N + 3) / 4 : 8;
  fsn::lstm_tc_carve((char*)workspace, off, (size_t)rows, K, Hm, x3 != 0, ws);
  FSN_REQUIRE(workspace && workspace_bytes >= off, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu", workspace_bytes, off);
  return fsn::linear_tc(x, (size_t)K, K, W, bias, N, act, out, (size_t)N, (size_t)rows, x3 != 0, ws, (cudaStream_t)stream);
}
