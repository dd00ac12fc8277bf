// tf32 tensor-core GEMM for the training path:  C[M,N] (+)= A[M,K] * B[N,K]^T, all fp32 in global memory, both
// operands K-contiguous ("K-major"), tcgen05.mma kind::tf32 with the accumulator in TMEM.
//
// One CTA = one 128 x BN output tile (BN = 128 or 256), optional split-K slice (blockIdx.z).
//   warps 0-3  producers: every thread owns one row of the A tile and one (BN=128) or two (BN=256) rows of the B
//              tile; per k-step (32 floats = one 128-byte swizzle row) it issues 16-byte cp.async copies straight
//              into the 128B-swizzled K-major layout the UMMA descriptors expect (chunk ^= row & 7); rows / k beyond
//              the matrix are zero-filled.  cp.async.wait_group + fence.proxy.async + mbarrier arrive hand the stage
//              to the tensor core.  After the last k-step the same warps drain TMEM (warp w <-> lanes 32w..32w+31).
//   warp 4     MMA issue (converged warp, one elected lane): 4 MMAs (K = 8) per stage, tcgen05.commit frees it.
// No operand conversion pass: the tensor core reads fp32 bits as tf32 (10-bit mantissa, truncation).
//
// Also in this file, built from the same pieces:
//   tgemm_tma_kernel        the same tile fed by TMA (default); BlockedOps mode streams block-tiled operand copies
//                           (transpose_blocked_kernel) for the long-K weight-gradient GEMMs dW = dG^T X
//   lstm_fwd_step_kernel    one LSTM step of the training forward: [x_t | h_{t-1}] [W_ih | W_hh]^T (fp16 / tf32 operands)
//                           and the cell in one launch; tile = 128 rows x (4 gates x 32 hidden units)
#include <cuda.h>
#include <cudaTypedefs.h>
#include <string.h>

#include "fsn_internal.cuh"
#include "fsn_tc_ptx.cuh"

namespace fsn {
namespace tg {

using namespace ptx;

constexpr int BM = 128, BK = 32;
constexpr int A_BYTES = BM * BK * 4;  // 16 KB

template <int BN> struct Cfg {
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  // BN = 128: 3 stages (96 KB) so two CTAs share an SM - one drains its accumulator while the other runs its main loop
  static constexpr int STAGES = (BN == 128) ? 3 : (BN == 256 ? 4 : 3);
  static constexpr int TMEM_COLS = (BN <= 128) ? 128 : (BN <= 256 ? 256 : 512);
  static constexpr int LAG = (BN == 128) ? 2 : 3;      // cp.async groups in flight per thread
  static constexpr int MIN_CTAS = (BN == 128) ? 2 : 1;
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

// block-tiled operand addressing of tgemm_tma_kernel (weight-gradient GEMMs, see tgemm_blocked_launch)
struct BlockedOps { int on, a_nkb, b_nkb, a_kb0, b_kb0; };

struct Bars {
  uint64_t full[8];
  uint64_t empty[8];
  uint64_t acc_full;
  uint32_t tmem_base;
};

// drain the accumulator: TMEM -> registers -> per-warp padded smem slab -> full 128-byte rows in global memory
// (all MMAs have completed, so the stage buffers are free)
template <int BN>
__device__ __forceinline__ void epilogue(uint8_t* smem, Bars& bars, uint32_t tmem_base, float* __restrict__ C, size_t ldc,
                                         int M, int N, int m0, int n0, int accumulate, size_t split_stride) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  mbar_wait<true>(&bars.acc_full, 0);
  tc_fence_after();
  float* slab = reinterpret_cast<float*>(smem) + warp * (32 * 33);
  float* cbase = C + (size_t)blockIdx.z * split_stride;
  const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int cb = 0; cb < BN / 32; ++cb) {
    if (n0 + cb * 32 >= N) break;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float v[8];
      tc_ld8(taddr + cb * 32 + q * 8, v);
      tc_wait_ld();
#pragma unroll
      for (int j = 0; j < 8; ++j) slab[lane * 33 + q * 8 + j] = v[j];
    }
    __syncwarp();
    const int col = n0 + cb * 32 + lane;
    if (col < N) {
      const int row0 = m0 + warp * 32;
#pragma unroll
      for (int r8 = 0; r8 < 32; r8 += 8) {
        float old[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)  // batch the read-modify-write loads: one memory round trip per 8 rows
          old[j] = (accumulate && row0 + r8 + j < M) ? cbase[(size_t)(row0 + r8 + j) * ldc + col] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (row0 + r8 + j < M) cbase[(size_t)(row0 + r8 + j) * ldc + col] = old[j] + slab[(r8 + j) * 33 + lane];
      }
    }
    __syncwarp();
  }
}

template <int BN>
__device__ __forceinline__ void mma_loop(uint8_t* smem, Bars& bars, uint32_t tmem_base, int nk) {
  using CF = Cfg<BN>;
  constexpr int STAGES = CF::STAGES;
  constexpr int N0 = BN > 256 ? 256 : BN, N1 = BN - N0;  // one MMA covers at most 256 columns
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N0 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  const uint32_t idesc1 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)((N1 > 0 ? N1 : 8) >> 3) << 17) |
                          ((uint32_t)(BM >> 4) << 24);
  const uint32_t smem_base = smem_u32(smem);
  for (int i = 0; i < nk; ++i) {
    const int s = i % STAGES;
    mbar_wait<false>(&bars.full[s], (uint32_t)((i / STAGES) & 1));
    tc_fence_after();
    if (elect_one()) {
      const uint32_t sa = smem_base + s * CF::STAGE_BYTES;
      const uint32_t sb = sa + A_BYTES;
#pragma unroll
      for (int kk = 0; kk < BK / 8; ++kk) {
        tc_mma1_tf32(tmem_base, desc_sw128(sa + kk * 32), desc_sw128(sb + kk * 32), idesc, (i > 0 || kk > 0) ? 1u : 0u);
        if (N1 > 0)  // columns 256.. of the tile: B rows 256.. start 256 * 128 bytes further
          tc_mma1_tf32(tmem_base + 256, desc_sw128(sa + kk * 32), desc_sw128(sb + 256 * 128 + kk * 32), idesc1,
                       (i > 0 || kk > 0) ? 1u : 0u);
      }
      tc_commit1(&bars.empty[s]);
    }
    __syncwarp();
  }
  if (elect_one()) tc_commit1(&bars.acc_full);
  __syncwarp();
}

template <int BN>
__global__ void __launch_bounds__(160, Cfg<BN>::MIN_CTAS)
tgemm_kernel(const float* __restrict__ A, size_t lda, const float* __restrict__ Bm, size_t ldb, float* __restrict__ C,
             size_t ldc, int M, int N, int K, int k_per_split, int accumulate, size_t split_stride) {
  using CF = Cfg<BN>;
  constexpr int STAGES = CF::STAGES, LAG = CF::LAG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Bars& bars = *reinterpret_cast<Bars*>(smem + STAGES * CF::STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb = blockIdx.z * k_per_split;
  const int ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  const int nk = (ke - kb + BK - 1) / BK;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 128); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                 "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    // lane -> (16-byte chunk c = tid & 7 of a 128-byte row, rows (tid >> 3) + 16 j): 8 lanes read one full line
    const int c = tid & 7, rbase = tid >> 3;
    const uint32_t sw = (uint32_t)(rbase & 7);  // (rbase + 16 j) & 7
    const uint32_t dst_off = (uint32_t)((rbase >> 3) * 1024 + (rbase & 7) * 128) + (((uint32_t)c ^ sw) << 4);
    const uint32_t smem_base = smem_u32(smem);
    const float* a_row = A + (size_t)(m0 + rbase) * lda + c * 4;
    const float* b_row = Bm + (size_t)(n0 + rbase) * ldb + c * 4;
    for (int i = 0; i < nk + LAG; ++i) {
      if (i < nk) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait<false>(&bars.empty[s], (uint32_t)(((i / STAGES) - 1) & 1));
        const int k0 = kb + i * BK;
        int rem = (ke - (k0 + c * 4)) * 4;
        rem = rem < 0 ? 0 : (rem > 16 ? 16 : rem);
        const uint32_t sa = smem_base + s * CF::STAGE_BYTES + dst_off;
#pragma unroll
        for (int j = 0; j < BM / 16; ++j) {
          const uint32_t nb = (m0 + rbase + 16 * j < M) ? (uint32_t)rem : 0u;
          cp_async16_zfill(sa + j * 2048, nb ? (const void*)(a_row + (size_t)(16 * j) * lda + k0) : (const void*)A, nb);
        }
#pragma unroll
        for (int j = 0; j < BN / 16; ++j) {
          const uint32_t nb = (n0 + rbase + 16 * j < N) ? (uint32_t)rem : 0u;
          cp_async16_zfill(sa + A_BYTES + j * 2048, nb ? (const void*)(b_row + (size_t)(16 * j) * ldb + k0) : (const void*)Bm,
                           nb);
        }
      }
      cp_async_commit();
      if (i >= LAG) {
        cp_async_wait<LAG>();   // group i-LAG (k-step i-LAG) has landed
        fence_proxy_async();    // generic-proxy writes -> visible to the tensor core (async proxy)
        mbar_arrive(&bars.full[(i - LAG) % STAGES]);
      }
    }
    epilogue<BN>(smem, bars, tmem_base, C, ldc, M, N, m0, n0, accumulate, split_stride);
  } else {
    mma_loop<BN>(smem, bars, tmem_base, nk);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN));
  }
}

// Same tile, operands fed by TMA: one elected lane issues two tiled loads per stage (128B hardware swizzle, rows / k
// beyond the matrix zero-filled by the tensor map), the full barrier counts the transaction bytes.
template <int BN>
__global__ void __launch_bounds__(160, Cfg<BN>::MIN_CTAS)
tgemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmB2, float* __restrict__ C, size_t ldc, int M, int N, int K,
                 int k_per_split, int accumulate, size_t split_stride, BlockedOps bo) {
  using CF = Cfg<BN>;
  constexpr int STAGES = CF::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Bars& bars = *reinterpret_cast<Bars*>(smem + STAGES * CF::STAGE_BYTES);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kb = blockIdx.z * k_per_split;
  const int ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  const int nk = (ke - kb + BK - 1) / BK;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                 "n"(CF::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  if (warp < 4) {
    if (warp == 0) {
      for (int i = 0; i < nk; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait<false>(&bars.empty[s], (uint32_t)(((i / STAGES) - 1) & 1));
        if (elect_one()) {
          uint8_t* sa = smem + s * CF::STAGE_BYTES;
          mbar_expect_tx(&bars.full[s], CF::STAGE_BYTES);
          if (bo.on) {
            // block-tiled operands: tile (128 rows, k block of 32) = 16 contiguous KB, see tgemm_blocked_launch
            const int kblk = (kb >> 5) + i;
            tma_load_2d(sa, &tmA, 0, (blockIdx.x * bo.a_nkb + bo.a_kb0 + kblk) * 128, &bars.full[s]);
#pragma unroll
            for (int j = 0; j < BN / 128; ++j)
              tma_load_2d(sa + A_BYTES + j * 128 * 128, &tmB, 0, ((blockIdx.y * (BN / 128) + j) * bo.b_nkb + bo.b_kb0 + kblk) * 128,
                          &bars.full[s]);
          } else {
            tma_load_2d(sa, &tmA, kb + i * BK, m0, &bars.full[s]);
            tma_load_2d(sa + A_BYTES, &tmB, kb + i * BK, n0, &bars.full[s]);
            if (BN > 256) tma_load_2d(sa + A_BYTES + 256 * 128, &tmB2, kb + i * BK, n0 + 256, &bars.full[s]);
          }
        }
        __syncwarp();
      }
    }
    epilogue<BN>(smem, bars, tmem_base, C, ldc, M, N, m0, n0, accumulate, split_stride);
  } else {
    mma_loop<BN>(smem, bars, tmem_base, nk);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CF::TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// One LSTM step of the training forward, GEMM and cell in one kernel (torch.nn.LSTM forward, saved for autograd):
//   z = P_t + b_ih + b_hh + h_{t-1} W_hh^T;  (i,f,g,o) = act(z);  c_t = f c_{t-1} + i g;  h_t = o tanh(c_t)
// The CTA tile is 128 rows x 128 gate columns = the FOUR gates of 32 hidden units: the B operand is four 32-row TMA
// boxes of W_hh (rows g*H + 32 j ..), so the accumulator holds everything the cell of those units needs and the
// recurrent product never goes to memory.  The epilogue turns the accumulator through shared memory so that every
// global access is a full 128-byte row piece: reads P_t (the hoisted input projection, overwritten in place by the
// post-activation gates the backward pass needs) and c_{t-1}, writes gates, c_t, h_t.  Two CTAs per SM: one drains
// while the other multiplies.
// 1 / (1 + 2^(-x log2 e)) and 2 sigmoid(2x) - 1 on the MUFU unit (ex2.approx + rcp.approx: ~2 ulp; saturate correctly:
// ex2 -> inf gives rcp -> 0, ex2 -> 0 gives 1)
__device__ __forceinline__ float sigmoid_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return r;
}
__device__ __forceinline__ float tanh_mufu(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
  return fmaf(2.0f, r, -1.0f);
}

// HT: hidden size known at compile time (0: runtime) - every stride of the epilogue becomes an immediate offset
// ST: TMA stages (3: 2 CTAs per SM; 2: 3 CTAs per SM - the kernel is bound by the latency chain of one CTA - load, multiply,
// drain, cell - not by any unit, so more CTAs in flight is what raises the throughput)
template <int ST> struct StepCfg {
  static constexpr int MAIN = (ST * Cfg<128>::STAGE_BYTES > 4 * 4 * 32 * 33 * 4 ? ST * Cfg<128>::STAGE_BYTES : 4 * 4 * 32 * 33 * 4 + 1023) & ~1023;
  static constexpr int SMEM = MAIN + 1024 /*align*/ + 256 /*barriers*/;
};
template <bool FOLD, int HT, int ST>
__global__ void __launch_bounds__(192, (ST == 2 && HT != 0 && FOLD) ? 3 : 2)
lstm_fwd_step_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmWx, float* __restrict__ G,
                     const float* __restrict__ b_ih, const float* __restrict__ b_hh, const float* __restrict__ C_prev,
                     float* __restrict__ C_out, float* __restrict__ H_out, __half* __restrict__ H16_out, int R, int H_rt, int nkx,
                     int nkh, int x16, int h16) {
  const int H = HT ? HT : H_rt;
  using CF = Cfg<128>;
  constexpr int STAGES = ST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  Bars& bars = *reinterpret_cast<Bars*>(smem + StepCfg<ST>::MAIN);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, u0 = blockIdx.y * 32;
  // k blocks: first nkx of x_t W_ih^T (a narrow layer input is folded in here instead of a hoisted projection: G then
  // carries no P and is only written), then nkh of h_{t-1} W_hh^T (0 at the first step)
  const int nk = nkx + nkh;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); }
    mbar_init(&bars.acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars.tmem_base)),
                 "n"(CF::TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;
  if (warp == 5) {
    // TMA producer (its own warp: the four epilogue warps start their global loads right away)
    for (int i = 0; i < nk; ++i) {
      const int s = i % STAGES;
      if (i >= STAGES) mbar_wait<false>(&bars.empty[s], (uint32_t)(((i / STAGES) - 1) & 1));
      if (elect_one()) {
        uint8_t* sa = smem + s * CF::STAGE_BYTES;
        mbar_expect_tx(&bars.full[s], CF::STAGE_BYTES);
        const CUtensorMap* ma = i < nkx ? &tmX : &tmA;
        const CUtensorMap* mb = i < nkx ? &tmWx : &tmB;
        // a stage row is 128 bytes: 32 tf32 or 64 fp16 k values
        const int k0 = (i < nkx ? i * (x16 ? 2 * BK : BK) : (i - nkx) * (h16 ? 2 * BK : BK));
        tma_load_2d(sa, ma, k0, m0, &bars.full[s]);
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_load_2d(sa + A_BYTES + g * 32 * 128, mb, k0, g * H + u0, &bars.full[s]);
      }
      __syncwarp();
    }
  } else if (warp < 4) {
    // ---- epilogue: lane = hidden unit u0 + lane (H % 32 == 0), the warp walks its 32 rows in batches of 8.  Kept
    // small on purpose: the straight-line version of this loop overflowed the instruction cache (ncu: 29 % of the
    // samples on stall_no_inst) and the cell costs more issue slots than the MMAs
    const int u = u0 + lane;
    const int row0 = m0 + warp * 32;
    const int nrows = R - row0 < 32 ? R - row0 : 32;  // may be <= 0: nothing to do but the barriers
    constexpr int RB = 8;
    const unsigned H4 = 4u * (unsigned)H;
    float b[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) b[g] = b_ih[g * H + u] + b_hh[g * H + u];
    const float* cp_ptr = C_prev ? C_prev + (size_t)row0 * H + u : nullptr;
    float* g_ptr = G + (size_t)row0 * H4 + u;
    float cp[RB], pz[FOLD ? 1 : RB][4];
    auto load_rows = [&](int r0) {
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        const bool ok = r0 + j < nrows;
        cp[j] = (ok && cp_ptr) ? (cp_ptr + (size_t)r0 * H)[j * H] : 0.f;
        if (!FOLD) {
#pragma unroll
          for (int g = 0; g < 4; ++g) pz[j][g] = ok ? (g_ptr + (size_t)r0 * H4)[j * 4 * H + g * H] : 0.f;
        }
      }
    };
    load_rows(0);  // in flight while the main loop runs
    mbar_wait<true>(&bars.acc_full, 0);
    tc_fence_after();
    // accumulator -> per-warp slab [gate][row][33] in the (now free) stage buffers, shared-space addresses
    const uint32_t slab = smem_u32(smem) + (uint32_t)warp * (4 * 32 * 33 * 4);
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float v[4][8];
#pragma unroll
      for (int q = 0; q < 4; ++q) tc_ld8(taddr + g * 32 + q * 8, v[q]);
      tc_wait_ld();
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 8; ++j)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(slab + (uint32_t)(((g * 32 + lane) * 33 + q * 8 + j) * 4)), "f"(v[q][j]));
    }
    __syncwarp();
    float* c_ptr = C_out + (size_t)row0 * H + u;
    float* h_ptr = H_out + (size_t)row0 * H + u;
    __half* h16_ptr = H16_out ? H16_out + (size_t)row0 * H + u : nullptr;
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += RB) {
      float gi[RB], gf[RB], gg[RB], go[RB], cn[RB], hn[RB];
#pragma unroll
      for (int j = 0; j < RB; ++j) {
        float z[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float a;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(slab + (uint32_t)(((g * 32 + r0 + j) * 33 + lane) * 4)));
          z[g] = FOLD ? a + b[g] : (pz[j][g] + b[g]) + a;
        }
        // ex2 + rcp forms (2 ulp class; the tf32 / fp16 products around them are 1e-3 class).  FSN_TRAIN_FAST_ACT=0
        // selects the unfused GEMM + lstm_cell_fwd_kernel path with expf / IEEE division instead
        gi[j] = sigmoid_mufu(z[0]); gf[j] = sigmoid_mufu(z[1]); gg[j] = tanh_mufu(z[2]); go[j] = sigmoid_mufu(z[3]);
        cn[j] = fmaf(gf[j], cp[j], gi[j] * gg[j]);
        hn[j] = go[j] * tanh_mufu(cn[j]);
      }
      // one base per array and batch, everything else an offset (immediate when HT != 0)
      float* gr = g_ptr + (size_t)r0 * H4;
      float* cr = c_ptr + (size_t)r0 * H;
      float* hr = h_ptr + (size_t)r0 * H;
      __half* h16r = h16_ptr + (size_t)r0 * H;
      if (r0 + RB < 32) load_rows(r0 + RB);  // next batch in flight under this batch's stores
      if (r0 + RB <= nrows) {                // whole batch inside the matrix: no per-row predicates
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          gr[j * 4 * H] = gi[j]; gr[j * 4 * H + H] = gf[j]; gr[j * 4 * H + 2 * H] = gg[j]; gr[j * 4 * H + 3 * H] = go[j];
          cr[j * H] = cn[j];
          hr[j * H] = hn[j];
        }
        if (h16_ptr) {
#pragma unroll
          for (int j = 0; j < RB; ++j) h16r[j * H] = __float2half_rn(hn[j]);  // next step's / next layer's MMA operand
        }
      } else {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          if (r0 + j < nrows) {
            gr[j * 4 * H] = gi[j]; gr[j * 4 * H + H] = gf[j]; gr[j * 4 * H + 2 * H] = gg[j]; gr[j * 4 * H + 3 * H] = go[j];
            cr[j * H] = cn[j];
            hr[j * H] = hn[j];
            if (h16_ptr) h16r[j * H] = __float2half_rn(hn[j]);
          }
        }
      }
    }
  } else if (warp == 4) {
    // MMA issue: per stage four instructions over 32 bytes of k each (8 tf32 or 16 fp16 values); both kinds accumulate
    // into the same fp32 tile
    const uint32_t idesc32 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t idesc16 = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    const uint32_t smem_base = smem_u32(smem);
    for (int i = 0; i < nk; ++i) {
      const int s = i % STAGES;
      mbar_wait<false>(&bars.full[s], (uint32_t)((i / STAGES) & 1));
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sa = smem_base + s * CF::STAGE_BYTES;
        const uint32_t sb = sa + A_BYTES;
        const bool half_blk = i < nkx ? (x16 != 0) : (h16 != 0);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (half_blk) tc_mma1_f16(tmem_base, desc_sw128(sa + kk * 32), desc_sw128(sb + kk * 32), idesc16, (i > 0 || kk > 0) ? 1u : 0u);
          else          tc_mma1_tf32(tmem_base, desc_sw128(sa + kk * 32), desc_sw128(sb + kk * 32), idesc32, (i > 0 || kk > 0) ? 1u : 0u);
        }
        tc_commit1(&bars.empty[s]);
      }
      __syncwarp();
    }
    if (elect_one()) tc_commit1(&bars.acc_full);
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(CF::TMEM_COLS));
  }
}

}  // namespace tg

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
static PFN_cuTensorMapEncodeTiled_v12000 tmap_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (getenv("FSN_TGEMM_FEED") == nullptr || strcmp(getenv("FSN_TGEMM_FEED"), "cpasync") != 0)
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
          q == cudaDriverEntryPointSuccess)
        fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    cudaGetLastError();
  }
  return fn;
}

// fp32 [rows, K] row-major (ld floats) -> boxes of box_rows x 32 floats, 128B swizzle, zero fill outside
static bool make_tmap(CUtensorMap* m, const float* base, int K, int rows, size_t ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = tmap_encoder();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)tg::BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// fp16 [rows, K] row-major (ld halfs) -> boxes of box_rows x 64 halfs (128 bytes), 128B swizzle, zero fill outside
static bool make_tmap16(CUtensorMap* m, const __half* base, int K, int rows, size_t ld, int box_rows) {
  PFN_cuTensorMapEncodeTiled_v12000 fn = tmap_encoder();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)(2 * tg::BK), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ==
         CUDA_SUCCESS;
}

// fixed-order sum of split-K slabs (fsn_train.cu)
int splitk_reduce_launch(const float* part, int S, int M, int N, float* C, size_t ldc, bool accumulate, cudaStream_t st);

bool tgemm_supported(const float* A, size_t lda, const float* Bm, size_t ldb, int K) {
  static int ok = -1;
  if (ok < 0) {
    int dev = 0, major = 0, smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    ok = (major == 10 && smem >= tg::Cfg<128>::SMEM && getenv("FSN_NO_TGEMM") == nullptr) ? 1 : 0;
  }
  return ok == 1 && K >= 4 && (lda & 3) == 0 && (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(Bm) & 15) == 0;
}

// C[M,N] (+)= A[M,K] B[N,K]^T; `scratch` (>= scratch_floats) enables split-K for long-K / few-tile problems
int tgemm_launch(const float* A, size_t lda, const float* Bm, size_t ldb, float* C, size_t ldc, int M, int N, int K,
                 bool accumulate, float* scratch, size_t scratch_floats, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return FSN_OK;
  FSN_REQUIRE(tgemm_supported(A, lda, Bm, ldb, K), FSN_ERR_UNSUPPORTED, "tgemm: operands must be 16-byte aligned rows");
  static const int force_bn = getenv("FSN_TGEMM_BN") ? atoi(getenv("FSN_TGEMM_BN")) : 0;
  int BN = force_bn ? force_bn : ((N >= 256 && N % 256 == 0) ? 256 : 128);
  // a handful of tiles (per-step GEMMs of the full-band stack, 64 rows): narrow tiles + split-K so that the weight
  // matrix is streamed by ~64 CTAs instead of 2-8
  const bool few = scratch && K < 8192 && cdiv(M, tg::BM) * cdiv(N, 128) <= 32;
  if (few && !force_bn) BN = 128;
  // short K, many tiles (hoisted input projections: output-write bound): 128-wide tiles, two CTAs per SM, so one CTA's
  // epilogue overlaps the other's main loop (measured 1081 -> 945 us at K = 384, 778 -> 522 us at K = 32 per 195 k rows)
  static const int smallk_bn = getenv("FSN_TGEMM_SMALLK_BN") ? atoi(getenv("FSN_TGEMM_SMALLK_BN")) : 128;
  if (!force_bn && K <= 512 && BN == 256 && cdiv(M, tg::BM) >= 1024) BN = smallk_bn == 256 ? 256 : 128;
  // long-K, narrow output (weight gradients): the whole N extent in one CTA so the big A operand is read exactly once
  static const bool wide = getenv("FSN_TGEMM_NO384") == nullptr;
  if (!force_bn && wide && tmap_encoder() && scratch && K >= 65536 && N > 256 && N <= 384) BN = 384;
  const int tiles = cdiv(M, tg::BM) * cdiv(N, BN);
  int S = 1;
  if (scratch && K >= 8192 && tiles < 296) {
    // minimise waves(tiles * S) / S over the SM slots (1 or 2 resident CTAs per SM), slices of >= 2048 k
    const int slots = 148 * (BN == 128 ? 2 : 1);  // resident CTAs
    double best = 1e30;
    for (int s = 1; s <= 64 && s <= cdiv(K, 2048); ++s) {
      if ((size_t)s * M * N > scratch_floats) break;
      const double cost = (double)cdiv(tiles * s, slots) / s + 1e-4 * s;
      if (cost < best) { best = cost; S = s; }
    }
  }
  if (few && K >= 256) {
    S = 64 / tiles;
    if (S > K / 128) S = K / 128;
    while (S > 1 && (size_t)S * M * N > scratch_floats) --S;
    if (S < 1) S = 1;
  }
  const int kps = cdiv(cdiv(K, S), tg::BK) * tg::BK;
  S = cdiv(K, kps);
  dim3 grid(cdiv(M, tg::BM), cdiv(N, BN), S);
  float* dst = S > 1 ? scratch : C;
  const size_t ldd = S > 1 ? (size_t)N : ldc;
  const int acc = (S > 1) ? 0 : (accumulate ? 1 : 0);
  int rc;
  CUtensorMap tmA, tmB;
  CUtensorMap tmB2;
  if (make_tmap(&tmA, A, K, M, lda, tg::BM) && make_tmap(&tmB, Bm, K, N, ldb, BN > 256 ? 256 : BN) &&
      make_tmap(&tmB2, Bm, K, N, ldb, BN > 256 ? BN - 256 : 128)) {
    static bool attr_by_dev[64] = {};  // the opt-in is per device
    int cur_dev_ = 0; cudaGetDevice(&cur_dev_); bool& attr = attr_by_dev[cur_dev_ & 63];
    if (!attr) {
      if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                tg::Cfg<256>::SMEM), "tgemm smem attr")))
        return rc;
      if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                tg::Cfg<128>::SMEM), "tgemm smem attr")))
        return rc;
      if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                tg::Cfg<384>::SMEM), "tgemm smem attr")))
        return rc;
      attr = true;
    }
    if (BN == 384)
      tg::tgemm_tma_kernel<384><<<grid, 160, tg::Cfg<384>::SMEM, st>>>(tmA, tmB, tmB2, dst, ldd, M, N, K, kps, acc,
                                                                       (size_t)M * N, tg::BlockedOps{0, 0, 0, 0, 0});
    else if (BN == 256)
      tg::tgemm_tma_kernel<256><<<grid, 160, tg::Cfg<256>::SMEM, st>>>(tmA, tmB, tmB2, dst, ldd, M, N, K, kps, acc,
                                                                       (size_t)M * N, tg::BlockedOps{0, 0, 0, 0, 0});
    else
      tg::tgemm_tma_kernel<128><<<grid, 160, tg::Cfg<128>::SMEM, st>>>(tmA, tmB, tmB2, dst, ldd, M, N, K, kps, acc,
                                                                       (size_t)M * N, tg::BlockedOps{0, 0, 0, 0, 0});
    FSN_CHECK_LAUNCH("tgemm_tma_kernel");
    if (S > 1) return splitk_reduce_launch(scratch, S, M, N, C, ldc, accumulate, st);
    return FSN_OK;
  }
  FSN_REQUIRE(BN != 384, FSN_ERR_CUDA, "tgemm: tensor-map encoding failed for the 128x384 tile");
  if (BN == 256) {
    static bool attr_by_dev[64] = {};  // the opt-in is per device
    int cur_dev_ = 0; cudaGetDevice(&cur_dev_); bool& attr = attr_by_dev[cur_dev_ & 63];
    if (!attr) {
      if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                tg::Cfg<256>::SMEM), "tgemm smem attr")))
        return rc;
      attr = true;
    }
    tg::tgemm_kernel<256><<<grid, 160, tg::Cfg<256>::SMEM, st>>>(A, lda, Bm, ldb, dst, ldd, M, N, K, kps, acc,
                                                                 (size_t)M * N);
  } else {
    static bool attr_by_dev[64] = {};  // the opt-in is per device
    int cur_dev_ = 0; cudaGetDevice(&cur_dev_); bool& attr = attr_by_dev[cur_dev_ & 63];
    if (!attr) {
      if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                tg::Cfg<128>::SMEM), "tgemm smem attr")))
        return rc;
      attr = true;
    }
    tg::tgemm_kernel<128><<<grid, 160, tg::Cfg<128>::SMEM, st>>>(A, lda, Bm, ldb, dst, ldd, M, N, K, kps, acc,
                                                                 (size_t)M * N);
  }
  FSN_CHECK_LAUNCH("tgemm_kernel");
  if (S > 1) return splitk_reduce_launch(scratch, S, M, N, C, ldc, accumulate, st);
  return FSN_OK;
}


// fused recurrent GEMM + LSTM cell of one training-forward step (tg::lstm_fwd_step_kernel); G_t [R,4H] holds the hoisted
// input projection and receives the post-activation gates
bool lstm_fwd_step_supported(const float* Hbuf, const float* w_hh, int H) {
  static const int mode = getenv("FSN_TRAIN_FUSED_FWD") ? atoi(getenv("FSN_TRAIN_FUSED_FWD")) : 1;
  static const int fast_act = getenv("FSN_TRAIN_FAST_ACT") ? atoi(getenv("FSN_TRAIN_FAST_ACT")) : 1;
  return mode != 0 && fast_act != 0 && (H % 32) == 0 && tmap_encoder() != nullptr && tgemm_supported(Hbuf, H, w_hh, H, H);
}
// the layer input is multiplied inside the step kernel (no hoisted projection, no P round trip through HBM: 2 x 9.6 GB per
// sub-band layer at config 3).  Measured per training step: no fold 110.5 ms, fold K0 <= 64 105.6 ms, all layers 99.1 ms
bool lstm_fwd_step_folds_input(const float* X, const float* w_ih, int K0) {
  static const int maxk = getenv("FSN_TRAIN_FOLD_K") ? atoi(getenv("FSN_TRAIN_FOLD_K")) : 512;
  return K0 <= maxk && tgemm_supported(X, K0, w_ih, K0, K0);
}
// fp16 copies of the MMA operands (h and the weights rounded to nearest: the same 11-bit significand as tf32 reads, half the
// bytes through L2 and twice the tensor rate); the fp32 state, the saved activations and the backward pass are unchanged
bool lstm_fwd_step_half_enabled(int H) {
  static const int on = getenv("FSN_TRAIN_F16_FWD") ? atoi(getenv("FSN_TRAIN_F16_FWD")) : 1;
  return on != 0 && (H % 8) == 0;
}
// Hprev == nullptr: first step (no recurrent term).  Xt / w_ih (nullable together): fold x_t W_ih^T in, G_t is then
// write-only; otherwise G_t holds the hoisted projection P_t.  h: optional fp16 operands (see LstmStepHalf)
int lstm_fwd_step_launch(const float* Hprev, const float* w_hh, const float* Xt, const float* w_ih, int K0, float* Gt,
                         const float* b_ih, const float* b_hh, const float* C_prev, float* C_out, float* H_out, int R, int H,
                         cudaStream_t st, const LstmStepHalf* h) {
  CUtensorMap tmA, tmB, tmX, tmWx;
  const bool h16 = h && h->w_hh16 && h->H16_out, x16 = h16 && Xt && h->Xt16 && h->w_ih16;
  bool ok;
  if (h16) {
    const __half* a = Hprev ? h->Hprev16 : h->H16_out;  // any valid [R,H] block: not read when nkh == 0
    ok = make_tmap16(&tmA, a, H, R, H, tg::BM) && make_tmap16(&tmB, h->w_hh16, H, 4 * H, H, 32);
  } else {
    const float* a = Hprev ? Hprev : H_out;
    ok = make_tmap(&tmA, a, H, R, H, tg::BM) && make_tmap(&tmB, w_hh, H, 4 * H, H, 32);
  }
  if (Xt && x16) ok = ok && make_tmap16(&tmX, h->Xt16, K0, R, K0, tg::BM) && make_tmap16(&tmWx, h->w_ih16, K0, 4 * H, K0, 32);
  else if (Xt)   ok = ok && make_tmap(&tmX, Xt, K0, R, K0, tg::BM) && make_tmap(&tmWx, w_ih, K0, 4 * H, K0, 32);
  else { tmX = tmA; tmWx = tmB; }
  FSN_REQUIRE(ok, FSN_ERR_CUDA, "lstm_fwd_step: tensor-map encoding failed");
  static bool attr_by_dev[64] = {};
  int dev = 0; cudaGetDevice(&dev); bool& attr = attr_by_dev[dev & 63];
  if (!attr) {
    int rc;
#define FSN_STEP_ATTR(FOLD, HT)                                                                                               \
  if ((rc = check_cuda(cudaFuncSetAttribute(tg::lstm_fwd_step_kernel<FOLD, HT, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            tg::StepCfg<2>::SMEM), "lstm_fwd_step smem attr")))                             \
    return rc;                                                                                                                \
  if ((rc = check_cuda(cudaFuncSetAttribute(tg::lstm_fwd_step_kernel<FOLD, HT, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                            tg::StepCfg<3>::SMEM), "lstm_fwd_step smem attr")))                             \
    return rc
    FSN_STEP_ATTR(true, 384); FSN_STEP_ATTR(true, 512); FSN_STEP_ATTR(true, 0);
    FSN_STEP_ATTR(false, 384); FSN_STEP_ATTR(false, 512); FSN_STEP_ATTR(false, 0);
#undef FSN_STEP_ATTR
    attr = true;
  }
  const int nkx = Xt ? cdiv(K0, x16 ? 2 * tg::BK : tg::BK) : 0, nkh = Hprev ? cdiv(H, h16 ? 2 * tg::BK : tg::BK) : 0;
  FSN_REQUIRE(nkx + nkh > 0, FSN_ERR_SHAPE, "lstm_fwd_step: nothing to multiply");
  const dim3 grid(cdiv(R, tg::BM), H / 32);
  static const int stages = getenv("FSN_TRAIN_STEP_STAGES") ? atoi(getenv("FSN_TRAIN_STEP_STAGES")) : 2;
#define FSN_STEP_LAUNCH(FOLD, HT)                                                                                              \
  do {                                                                                                                         \
    if (stages == 2)                                                                                                           \
      tg::lstm_fwd_step_kernel<FOLD, HT, 2><<<grid, 192, tg::StepCfg<2>::SMEM, st>>>(                                          \
          tmA, tmB, tmX, tmWx, Gt, b_ih, b_hh, C_prev, C_out, H_out, h16 ? h->H16_out : nullptr, R, H, nkx, nkh, x16 ? 1 : 0,  \
          h16 ? 1 : 0);                                                                                                        \
    else                                                                                                                       \
      tg::lstm_fwd_step_kernel<FOLD, HT, 3><<<grid, 192, tg::StepCfg<3>::SMEM, st>>>(                                          \
          tmA, tmB, tmX, tmWx, Gt, b_ih, b_hh, C_prev, C_out, H_out, h16 ? h->H16_out : nullptr, R, H, nkx, nkh, x16 ? 1 : 0,  \
          h16 ? 1 : 0);                                                                                                        \
  } while (0)
  if (Xt) {
    if (H == 384) FSN_STEP_LAUNCH(true, 384); else if (H == 512) FSN_STEP_LAUNCH(true, 512); else FSN_STEP_LAUNCH(true, 0);
  } else {
    if (H == 384) FSN_STEP_LAUNCH(false, 384); else if (H == 512) FSN_STEP_LAUNCH(false, 512); else FSN_STEP_LAUNCH(false, 0);
  }
  (void)0;
#undef FSN_STEP_LAUNCH
  FSN_CHECK_LAUNCH("lstm_fwd_step_kernel");
  return FSN_OK;
}

namespace tg {
__global__ void to_half_kernel(const float* __restrict__ in, size_t n, __half* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(in[i]);
}
}  // namespace tg
int to_half_launch(const float* in, size_t n, __half* out, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  tg::to_half_kernel<<<blocks, 256, 0, st>>>(in, n, out);
  FSN_CHECK_LAUNCH("to_half_kernel");
  return FSN_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Weight-gradient GEMMs  C[M,N] = A^T B  with A [K,M] and B [K,N] row-major (dW = dG^T X, K = T'R up to 1.5 M).
// The tensor-core kernel wants K-major operands; a plain transposed copy [M,K] has a row pitch of K floats (6 MB), so
// every 128-row TMA box touches 128 DRAM pages and the GEMM ran at 15 % of the HBM bandwidth.  Here the transposed copy
// is BLOCK-TILED: tile (128 rows of M, k block of 32) = 16 contiguous KB at ((mt * nkb + kb) * 128 + row) * 32 + kk
// (zero padded in M and K), viewed by TMA as a 2-D array [rows, 32] - one box = one contiguous burst.
namespace tg {
// one CTA: m tile of 128 columns x `kb_per` k blocks of 32 rows; 512-byte row pieces in, one contiguous 16 KB tile out per
// k block.  CTAs are numbered m tile fastest, so the CTAs resident together read whole rows.  colsum_part (nullable):
// [gridDim.y][M] column sums of this CTA's rows (bias gradients = column sums of dG, free while the tile is in SMEM)
__global__ void __launch_bounds__(256) transpose_blocked_kernel(const float* __restrict__ in, size_t K, int M, size_t ld, int nkb,
                                                                 int kb_per, float* __restrict__ out, float* __restrict__ colsum_part) {
  __shared__ float tile[32][129];
  const int tid = threadIdx.x;
  const int mt = blockIdx.x, m0 = mt * 128;
  const int kb0 = blockIdx.y * kb_per;
  const int kb1 = (kb0 + kb_per < nkb) ? kb0 + kb_per : nkb;
  const int col = tid & 127, rsub = tid >> 7;
  const bool col_ok = m0 + col < M;
  float csum = 0.f;
  for (int kb = kb0; kb < kb1; ++kb) {
    const size_t k0 = (size_t)kb * 32;
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const size_t k = k0 + i * 2 + rsub;
      v[i] = (col_ok && k < K) ? __ldcs(in + k * ld + m0 + col) : 0.f;
    }
    __syncthreads();  // the previous tile has been read out
#pragma unroll
    for (int i = 0; i < 16; ++i) tile[i * 2 + rsub][col] = v[i];
    __syncthreads();
    float* o = out + ((size_t)mt * nkb + kb) * (128 * 32);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int e = i * 256 + tid;
      o[e] = tile[e & 31][e >> 5];
    }
    if (colsum_part && tid < 128) {
#pragma unroll
      for (int r = 0; r < 32; ++r) csum += tile[r][tid];  // fixed order
    }
  }
  if (colsum_part && tid < 128 && m0 + tid < M) colsum_part[(size_t)blockIdx.y * M + m0 + tid] = csum;
}
}  // namespace tg

size_t tgemm_blocked_floats(size_t K, int M) { return (size_t)cdiv(M, 128) * 128 * ((K + 31) / 32) * 32; }

// in [K, M] (row stride ld) -> block-tiled transposed copy (tgemm_blocked_floats(K, M) floats).  colsum_part (nullable,
// >= max_slabs * M floats): per-slab column sums, *slabs receives the number of slabs written
int transpose_blocked_launch(const float* in, size_t K, int M, size_t ld, float* out, cudaStream_t st, float* colsum_part,
                             int max_slabs, int* slabs) {
  const int nkb = (int)((K + 31) / 32);
  int kb_per = 8;  // 256 rows per CTA
  if (colsum_part && cdiv(nkb, kb_per) > max_slabs) kb_per = cdiv(nkb, max_slabs);
  const int S = cdiv(nkb, kb_per);
  if (slabs) *slabs = S;
  dim3 grid((unsigned)cdiv(M, 128), (unsigned)S);
  FSN_REQUIRE(S <= 65535, FSN_ERR_SHAPE, "transpose_blocked: K too long");
  tg::transpose_blocked_kernel<<<grid, 256, 0, st>>>(in, K, M, ld, nkb, kb_per, out, colsum_part);
  FSN_CHECK_LAUNCH("transpose_blocked_kernel");
  return FSN_OK;
}

bool tgemm_blocked_enabled() {
  static const bool on = getenv("FSN_TGEMM_BLOCKED") == nullptr || atoi(getenv("FSN_TGEMM_BLOCKED")) != 0;
  return on && tmap_encoder() != nullptr;
}

// C[M,N] (+)= A^T B over K, operands block-tiled (transpose_blocked_launch) with nkb_a / nkb_b k blocks per tile row;
// a_kb0 / b_kb0: first k block of each operand (lets dW_hh pair dG[1:] with H[:-1])
int tgemm_blocked_launch(const float* Ablk, int nkb_a, int a_kb0, const float* Bblk, int nkb_b, int b_kb0, float* C, size_t ldc,
                         int M, int N, int K, bool accumulate, float* scratch, size_t scratch_floats, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return FSN_OK;
  PFN_cuTensorMapEncodeTiled_v12000 fn = tmap_encoder();
  FSN_REQUIRE(fn, FSN_ERR_UNSUPPORTED, "tgemm_blocked: cuTensorMapEncodeTiled unavailable");
  int BN = (N > 256 && N <= 384) ? 384 : ((N > 128 && N <= 256) || N % 256 == 0 ? 256 : 128);
  const int tiles = cdiv(M, tg::BM) * cdiv(N, BN);
  int S = 1;
  if (scratch && K >= 8192 && tiles < 296) {
    const int slots = 148 * (BN == 128 ? 2 : 1);
    double best = 1e30;
    for (int s = 1; s <= 64 && s <= cdiv(K, 2048); ++s) {
      if ((size_t)s * M * N > scratch_floats) break;
      const double cost = (double)cdiv(tiles * s, slots) / s + 1e-4 * s;
      if (cost < best) { best = cost; S = s; }
    }
  }
  const int kps = cdiv(cdiv(K, S), tg::BK) * tg::BK;
  S = cdiv(K, kps);
  dim3 grid(cdiv(M, tg::BM), cdiv(N, BN), S);
  float* dst = S > 1 ? scratch : C;
  const size_t ldd = S > 1 ? (size_t)N : ldc;
  const int acc = (S > 1) ? 0 : (accumulate ? 1 : 0);
  CUtensorMap tmA, tmB;
  cuuint64_t gstr[1] = {128};
  cuuint32_t box[2] = {32, 128}, estr[2] = {1, 1};
  cuuint64_t gA[2] = {32, (cuuint64_t)cdiv(M, 128) * 128 * (cuuint64_t)nkb_a};
  cuuint64_t gB[2] = {32, (cuuint64_t)cdiv(N, 128) * 128 * (cuuint64_t)nkb_b};  // tiles past it: zero fill
  FSN_REQUIRE(fn(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)Ablk, gA, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                  fn(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)Bblk, gB, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS,
              FSN_ERR_CUDA, "tgemm_blocked: tensor-map encoding failed");
  int rc;
  if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::Cfg<384>::SMEM), "tgemm smem attr"))) return rc;
  if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::Cfg<256>::SMEM), "tgemm smem attr"))) return rc;
  if ((rc = check_cuda(cudaFuncSetAttribute(tg::tgemm_tma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::Cfg<128>::SMEM), "tgemm smem attr"))) return rc;
  const tg::BlockedOps bo{1, nkb_a, nkb_b, a_kb0, b_kb0};
  if (BN == 384)
    tg::tgemm_tma_kernel<384><<<grid, 160, tg::Cfg<384>::SMEM, st>>>(tmA, tmB, tmB, dst, ldd, M, N, K, kps, acc, (size_t)M * N, bo);
  else if (BN == 256)
    tg::tgemm_tma_kernel<256><<<grid, 160, tg::Cfg<256>::SMEM, st>>>(tmA, tmB, tmB, dst, ldd, M, N, K, kps, acc, (size_t)M * N, bo);
  else
    tg::tgemm_tma_kernel<128><<<grid, 160, tg::Cfg<128>::SMEM, st>>>(tmA, tmB, tmB, dst, ldd, M, N, K, kps, acc, (size_t)M * N, bo);
  FSN_CHECK_LAUNCH("tgemm_tma_kernel");
  if (S > 1) return splitk_reduce_launch(scratch, S, M, N, C, ldc, accumulate, st);
  return FSN_OK;
}
}  // namespace fsn

// debug / unit-test entry point (tests/test_gpu_train.py): C[M,N] (+)= A[M,K] B[N,K]^T on the tcgen05 path
extern "C" int fsn_debug_tgemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M,
                               int N, int K, int accumulate, float* scratch, int64_t scratch_floats,
                               fsn_stream_t stream) {
  return fsn::tgemm_launch(A, (size_t)lda, B, (size_t)ldb, C, (size_t)ldc, M, N, K, accumulate != 0, scratch,
                           (size_t)scratch_floats, (cudaStream_t)stream);
}

// unit-test hook: C[M,N] = A^T B for row-major A [K,M], B [K,N] through the block-tiled transposes + tgemm_blocked_launch
// (a_k0 / b_k0: first k row of each operand, multiples of 32); scratch: tgemm_blocked_floats(K, M) + (K, N padded to the
// tile) floats for the copies, then split-K space
extern "C" int fsn_debug_tgemm_blocked(const float* A, const float* B, float* C, int M, int N, int K, int a_k0, int b_k0,
                                       float* scratch, int64_t scratch_floats, fsn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Ka = (size_t)K + a_k0, Kb = (size_t)K + b_k0;
  const size_t fa = fsn::tgemm_blocked_floats(Ka, M), fb = fsn::tgemm_blocked_floats(Kb, N);
  FSN_REQUIRE((a_k0 & 31) == 0 && (b_k0 & 31) == 0 && scratch && (size_t)scratch_floats >= fa + fb, FSN_ERR_SHAPE,
              "tgemm_blocked hook: bad offsets or scratch");
  float *Ab = scratch, *Bb = scratch + fa;
  int rc;
  if ((rc = fsn::transpose_blocked_launch(A, Ka, M, (size_t)M, Ab, st, nullptr, 0, nullptr))) return rc;
  if ((rc = fsn::transpose_blocked_launch(B, Kb, N, (size_t)N, Bb, st, nullptr, 0, nullptr))) return rc;
  return fsn::tgemm_blocked_launch(Ab, (int)((Ka + 31) / 32), a_k0 / 32, Bb, (int)((Kb + 31) / 32), b_k0 / 32, C, (size_t)N, M, N, K,
                                   false, scratch + fa + fb, (size_t)scratch_floats - fa - fb, st);
}

// unit-test hook for the fused training-forward step (tg::lstm_fwd_step_kernel; torch.nn.LSTM cell math): one step
//   z = [X W_ih^T  or  the P already in G] + b_ih + b_hh + Hprev W_hh^T;  G <- act(z) (i,f,g,o), C_out, H_out
// Hprev nullable (first step), X nullable (then G [R,4H] holds the hoisted projection on entry).  half != 0: fp16 MMA
// operands, converted here into `scratch` (>= 2 * (2 R H + 4 H (H + K0) + R K0) bytes)
extern "C" int fsn_debug_lstm_fwd_step(const float* Hprev, const float* w_hh, const float* X, const float* w_ih, int K0, float* G,
                                       const float* b_ih, const float* b_hh, const float* C_prev, float* C_out, float* H_out,
                                       int R, int H, int half, void* scratch, int64_t scratch_bytes, fsn_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  FSN_REQUIRE(R > 0 && H > 0 && w_hh && G && b_ih && b_hh && C_out && H_out && (!X || (w_ih && K0 > 0)), FSN_ERR_SHAPE,
              "lstm_fwd_step hook: missing arguments");
  FSN_REQUIRE(fsn::lstm_fwd_step_supported(H_out, w_hh, H), FSN_ERR_UNSUPPORTED,
              "lstm_fwd_step hook: needs H %% 32 == 0, 16-byte aligned operands and the TMA driver entry point");
  FSN_REQUIRE(!X || fsn::tgemm_supported(X, K0, w_ih, K0, K0), FSN_ERR_UNSUPPORTED, "lstm_fwd_step hook: X rows must be 16-byte aligned");
  if (!half) return fsn::lstm_fwd_step_launch(Hprev, w_hh, X, w_ih, K0, G, b_ih, b_hh, C_prev, C_out, H_out, R, H, st, nullptr);
  const size_t nh = (size_t)R * H, nw = (size_t)4 * H * H, nx = X ? (size_t)R * K0 : 0, nwx = X ? (size_t)4 * H * K0 : 0;
  FSN_REQUIRE((H % 8) == 0 && (!X || (K0 % 8) == 0), FSN_ERR_UNSUPPORTED, "lstm_fwd_step hook: fp16 rows must be 16-byte aligned");
  FSN_REQUIRE(scratch && (size_t)scratch_bytes >= 2 * (2 * nh + nw + nx + nwx) + 1024, FSN_ERR_WORKSPACE,
              "lstm_fwd_step hook: scratch too small");
  auto up = [](size_t n) { return (n + 127) & ~(size_t)127; };  // keep every block 256-byte aligned
  __half* hp = (__half*)scratch;
  __half* ho = hp + up(nh);
  __half* wh = ho + up(nh);
  __half* xx = wh + up(nw);
  __half* wx = xx + up(nx);
  FSN_REQUIRE((size_t)((char*)(wx + up(nwx)) - (char*)scratch) <= (size_t)scratch_bytes, FSN_ERR_WORKSPACE,
              "lstm_fwd_step hook: scratch too small");
  int rc;
  if (Hprev && (rc = fsn::to_half_launch(Hprev, nh, hp, st))) return rc;
  if ((rc = fsn::to_half_launch(w_hh, nw, wh, st))) return rc;
  if (X && ((rc = fsn::to_half_launch(X, nx, xx, st)) || (rc = fsn::to_half_launch(w_ih, nwx, wx, st)))) return rc;
  fsn::LstmStepHalf hs{Hprev ? hp : nullptr, wh, X ? xx : nullptr, X ? wx : nullptr, ho};
  return fsn::lstm_fwd_step_launch(Hprev, w_hh, X, w_ih, K0, G, b_ih, b_hh, C_prev, C_out, H_out, R, H, st, &hs);
}
