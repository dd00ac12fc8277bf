// Sub-band LSTM stack, CTA-PAIR variant: tcgen05.mma.cta_group::2 (sm_100a).
//
// Same mathematics and reference rows as fsn_subband_tc.cu (model.py:98-135), different mapping onto the
// machine.  A cluster of two CTAs (two SMs of one TPC) owns 64 sub-band units (rows): CTA r holds the
// recurrent state (x_t, h0, h1; fp16, UMMA K-major layout) of rows [32r, 32r+32) and HALF of the hidden
// units of both layers.  The leader CTA issues every MMA for the pair:
//     J1: M = 256 (CTA0: units 0..127, CTA1: units 128..255), N = 64, 4 gate accumulators x 64 TMEM columns
//     J2: M = 128 (CTA0: units 256..319, CTA1: units 320..383), N = 64, 4 x 32 columns
//         (2-CTA M=128 accumulator layout: lane = unit + 64*(n/32), column = n%32)
// so that every 4 KB weight tile an SM reads from shared memory is multiplied against 64 rows instead of 32,
// and each SM streams only its own half of the weights (1.78 MB per step instead of 3.55 MB): the tensor pipe,
// bound by the shared-memory read of the weight operand, does half the work per row.
// The epilogue thread that owns (hidden unit, 32-row half) keeps c in registers and writes h_t (fp16) into
// the state buffer of the CTA that owns those rows - its own or the peer's, through distributed shared memory
// (st.shared::cluster + fence.proxy.async), and publishes it with cluster-scope mbarrier arrives on the leader.
//
// Warp roles per CTA (512 threads): 0 = weight producer (own half of the stream), 1 = MMA issuer (leader) /
// stage relay (peer), 2 = x gather + TMEM alloc, 3 = Linear(H->2) reduce + output, 4-7 = epilogue J1 rows 0..31
// (-> CTA0), 8-11 = epilogue J1 rows 32..63 (-> CTA1), 12-15 = epilogue J2.
//
// X3 = true is the ERROR-COMPENSATED variant (FSN_PREC_F16X3_TC): weights and state are each split into two fp16
// terms (hi = rn(v), lo = rn(v - hi), 22 significand bits together) and every product is issued as three MMAs into the
// SAME fp32 accumulator, W_hi.S_hi + W_hi.S_lo + W_lo.S_hi (the dropped W_lo.S_lo term is 2^-22 relative).  The weight
// stream carries a hi stage and a lo stage per k range (the hi stage is used twice from shared memory); the state
// buffers hold hi and lo copies, h0 becomes single-buffered (h0_t is held in registers until the `h0_free` commit
// says the last MMA reading h0_{t-1} has completed, like h1) so that a 7-granule weight ring still fits, and the gate
// non-linearities use expf / IEEE division instead of the MUFU approximations.  Measured against the fp32 reference
// this is the fp32 class (cRM 1e-6 relative); the single-pass variant is 5e-4.
#include <stdlib.h>
#include <string.h>

#include "fsn_internal.cuh"
#include "fsn_tc_ptx.cuh"

#ifndef FSN_TC2_ACCOUNT
#define FSN_TC2_ACCOUNT 0   // 1: cycle accounting of the leader's MMA warp (diagnostic builds, FSN_TC_TRACE=1)
#endif
#if FSN_TC2_ACCOUNT
#define ACCT(...) __VA_ARGS__
#else
#define ACCT(...)
#endif

namespace fsn {
namespace tc2 {
using namespace ptx;

constexpr int H = 384;                 // hidden size this variant is built for
constexpr int NB = 32;                 // rows per CTA (64 per pair)
constexpr int KB = 64, KS = 32;
constexpr int S_KBLK = NB * KB * 2;    // 4096 B: one 64-k block of the state operand
constexpr int NKH = H / KB;            // 6 blocks per h
constexpr int W_SUB1 = 128 * KS * 2;   // J1 gate sub-tile (128 units x 32 k)
constexpr int W_ST1 = 4 * W_SUB1;      // 32 KB
constexpr int W_SUB2 = 64 * KS * 2;    // J2 gate sub-tile (64 units x 32 k)
constexpr int W_ST2 = 4 * W_SUB2;      // 16 KB
constexpr int NK0 = 1 + H / KS;        // 13 k-ranges: x, h0_{t-1}
constexpr int NK1 = 2 * H / KS;        // 24 k-ranges: h0_t, h1_{t-1}
constexpr int GRAN = 16384;            // ring granule; a J1 stage takes 2 granules ((i,f),(g,o)), a J2 stage 1
constexpr int OUT_T = 8;
constexpr int NTHREADS = 512;
constexpr int FC_SLOTS = 12;

// Everything that differs between the single-pass (X3 = false) and the error-compensated (X3 = true) variant.
template <bool X3>
struct Plan {
  static constexpr int PARTS = X3 ? 2 : 1;   // weight stages per k range: hi (, lo)
  static constexpr int NG = X3 ? 7 : 8;      // ring granules: 112 / 128 KB of weights in flight per CTA
  // packed stream, per rank and step: [L0 J1 stages][L0 J2 stages][L1 J1 stages][L1 J2 stages]; stage s = k*PARTS + part
  static constexpr size_t STREAM_BYTES = (size_t)PARTS * (NK0 + NK1) * (W_ST1 + W_ST2);
  static constexpr size_t OFF_L0J1 = 0;
  static constexpr size_t OFF_L0J2 = OFF_L0J1 + (size_t)PARTS * NK0 * W_ST1;
  static constexpr size_t OFF_L1J1 = OFF_L0J2 + (size_t)PARTS * NK0 * W_ST2;
  static constexpr size_t OFF_L1J2 = OFF_L1J1 + (size_t)PARTS * NK1 * W_ST1;
  static constexpr size_t OFF_BIAS = 2 * STREAM_BYTES;
  static constexpr size_t OFF_FCW = OFF_BIAS + (size_t)2 * 4 * H * sizeof(float);
  static constexpr size_t OFF_FCB = OFF_FCW + (size_t)2 * H * sizeof(float);
  static constexpr size_t PACKED_BYTES = OFF_FCB + 256;
  // shared-memory plan (identical in both CTAs).  x block: k 0..31 = hi, k 32..63 = lo (X3) / unused
  static constexpr int H0_BUFS = X3 ? 1 : 2;
  static constexpr uint32_t SM_W = 0;
  static constexpr uint32_t SM_X = SM_W + NG * GRAN;
  static constexpr uint32_t SM_H0 = SM_X + 2 * S_KBLK;
  static constexpr uint32_t SM_H1 = SM_H0 + H0_BUFS * NKH * S_KBLK;      // h1 single-buffered
  static constexpr uint32_t SM_LO = SM_H1 + NKH * S_KBLK;                // X3: [h0 lo][h1 lo]
  static constexpr uint32_t LO_OFF = SM_LO - SM_H0;                      // lo copy of a state block = hi + LO_OFF
  static constexpr uint32_t SM_FC = SM_LO + (X3 ? 2 * NKH * S_KBLK : 0); // [FC_SLOTS][2][NB] float
  static constexpr uint32_t SM_OUT = SM_FC + FC_SLOTS * 2 * NB * 4;      // [NB][2][OUT_T] float
  static constexpr uint32_t SM_ROWS = SM_OUT + NB * 2 * OUT_T * 4;
  static constexpr uint32_t SM_BARS = SM_ROWS + NB * 16;
  static constexpr uint32_t SM_TOTAL = SM_BARS + 384;
  static_assert(SM_TOTAL + 1024 <= 232448, "shared-memory plan exceeds 227 KB");
  static_assert(!X3 || LO_OFF == 2 * NKH * S_KBLK, "lo copies must sit at one offset from both h0 and h1");
};

constexpr uint32_t kIdesc256 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);  // f16 x f16 -> f32, N=64
constexpr uint32_t kIdesc128 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// ---------------------------------------------------------------- weight packer
// per rank r: [L0 J1 stages][L0 J2 stages][L1 J1 stages][L1 J2 stages]; a stage = 4 gate sub-tiles of one k range
// (X3: stage 2k = fp16 hi part, stage 2k+1 = fp16 lo part = rn(w - hi))
template <bool X3>
__global__ void pack2_kernel(const float* __restrict__ wih0, const float* __restrict__ whh0,
                             const float* __restrict__ wih1, const float* __restrict__ whh1,
                             const float* __restrict__ bih0, const float* __restrict__ bhh0,
                             const float* __restrict__ bih1, const float* __restrict__ bhh1,
                             const float* __restrict__ fcw, const float* __restrict__ fcb, int Ksb, int fc_out,
                             uint8_t* __restrict__ out) {
  using P = Plan<X3>;
  // one thread per 16-byte chunk (8 halves): chunks per rank = STREAM_BYTES / 16
  const size_t chunks_per_rank = P::STREAM_BYTES / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < 2 * chunks_per_rank;
       i += (size_t)gridDim.x * blockDim.x) {
    const int rank = (int)(i / chunks_per_rank);
    size_t off = (i % chunks_per_rank) * 16;  // byte offset inside the rank's stream
    int layer, j2;
    if (off < P::OFF_L0J2) { layer = 0; j2 = 0; }
    else if (off < P::OFF_L1J1) { layer = 0; j2 = 1; off -= P::OFF_L0J2; }
    else if (off < P::OFF_L1J2) { layer = 1; j2 = 0; off -= P::OFF_L1J1; }
    else { layer = 1; j2 = 1; off -= P::OFF_L1J2; }
    const int st_bytes = j2 ? W_ST2 : W_ST1, sub_bytes = j2 ? W_SUB2 : W_SUB1;
    const int stage = (int)(off / st_bytes);
    const int kb = stage / P::PARTS, part = stage % P::PARTS;
    const int in_st = (int)(off % st_bytes);
    const int g = in_st / sub_bytes;
    const int in_sub = in_st % sub_bytes;
    // invert swz64_off: 512-byte groups of 8 rows, 64-byte rows, 16-byte chunks XOR-ed with (row>>1)&3
    const int grp = in_sub / 512, rr = (in_sub % 512) / 64, cx = (in_sub % 64) / 16;
    const int row = grp * 8 + rr;
    const int c = cx ^ ((row >> 1) & 3);
    const int unit = j2 ? 256 + rank * 64 + row : rank * 128 + row;
    const int wrow = g * H + unit;
    __half v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = c * 8 + e;
      float w = 0.f;
      if (layer == 0) {
        if (kb == 0) { if (kk < Ksb) w = wih0[(size_t)wrow * Ksb + kk]; }
        else w = whh0[(size_t)wrow * H + (kb - 1) * KS + kk];
      } else {
        const int k = kb * KS + kk;
        w = (k < H) ? wih1[(size_t)wrow * H + k] : whh1[(size_t)wrow * H + (k - H)];
      }
      const __half hi = __float2half_rn(w);
      v[e] = part ? __float2half_rn(w - __half2float(hi)) : hi;
    }
    *reinterpret_cast<uint4*>(out + (size_t)rank * P::STREAM_BYTES + (i % chunks_per_rank) * 16) =
        *reinterpret_cast<const uint4*>(v);
  }
  float* bias = reinterpret_cast<float*>(out + P::OFF_BIAS);
  float* pfcw = reinterpret_cast<float*>(out + P::OFF_FCW);
  float* pfcb = reinterpret_cast<float*>(out + P::OFF_FCB);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * H; i += gridDim.x * blockDim.x) {
    bias[i] = bih0[i] + bhh0[i];
    bias[4 * H + i] = bih1[i] + bhh1[i];
  }
  // the kernel always evaluates 2 Linear outputs; a model with fewer (fast_fullsubnet bottleneck: 1) gets zero rows
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * H; i += gridDim.x * blockDim.x)
    pfcw[i] = (i < fc_out * H) ? fcw[i] : 0.f;
  if (blockIdx.x == 0 && threadIdx.x < 2) pfcb[threadIdx.x] = ((int)threadIdx.x < fc_out) ? fcb[threadIdx.x] : 0.f;
}

struct Bars {
  uint64_t w_full[8], w_empty[8];   // indexed by granule; a stage uses the full barrier of its first granule
  uint64_t x_full[2], x_empty[2];
  uint64_t accf_j1, accf_j2[2];     // MMA -> epilogue (multicast commit)
  uint64_t acce_j1, acce_j2[2];     // epilogue (both CTAs) -> leader MMA
  uint64_t h0_ready, h1_ready;      // epilogue (both CTAs) -> leader MMA
  uint64_t fc_ready, fc_done;       // Linear partials ready (12 writer warps) / consumed (both FC warps)
  uint64_t l1_done;                 // every layer-1 MMA of the step has completed: h1 may be overwritten
  uint64_t h0_free;                 // X3: the last MMA reading h0_{t-1} has completed: h0 may be overwritten with h0_t
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 384, "barrier block too large");

struct RowInfo {
  int src_b, src_f;
  float scale;
  int out_idx;
};

struct KArgs {
  const uint8_t* packed;
  const float* magT; const float* fbT; const float* inv2;
  const float* unit_scale;  // nullable: cumulative norm, scale of (step t, row r) at [t*R + r] instead of inv2[clip]
  float* crm;
  int R, F, Tp, la, T, Ns, Nf, Ksb, act, Fsub;
  int src_T, shrink;  // frames in magT/fbT; x_t = mean of `shrink` source frames (fast_fullsubnet down-sampling), 1 = none
  long long* dbg;  // FSN_TC_TRACE: cycle accounting of the leader's MMA warp (block 0)
  RowMap map;
};

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case FSN_ACT_RELU: return fmaxf(v, 0.f);
    case FSN_ACT_TANH: return tanhf(v);
    case FSN_ACT_RELU6: return fminf(fmaxf(v, 0.f), 6.f);
    default: return v;
  }
}

__device__ __forceinline__ void st_cluster_b16(uint32_t addr, __half v) {
  asm volatile("st.shared::cluster.b16 [%0], %1;" ::"r"(addr), "h"(__half_as_ushort(v)) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// gate non-linearities: MUFU approximations (~5e-7) for the single-pass variant, whose operand rounding is 2.4e-4
// anyway; libm-class expf + IEEE division for the compensated one, whose whole point is the fp32 error class
template <bool PRECISE> __device__ __forceinline__ float gate_sigmoid(float x) {
  if (PRECISE) return 1.0f / (1.0f + expf(-x));
  return fast_sigmoid(x);
}
template <bool PRECISE> __device__ __forceinline__ float gate_tanh(float x) {
  if (PRECISE) return 1.0f - 2.0f / (1.0f + expf(2.0f * x));
  return fast_tanh(x);
}

template <bool X3>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) sb_lstm_tc2_kernel(const KArgs a) {
  using P = Plan<X3>;
  constexpr int NG = P::NG;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem + P::SM_BARS);
  RowInfo* rows = reinterpret_cast<RowInfo*>(smem + P::SM_ROWS);
  float* fc_part = reinterpret_cast<float*>(smem + P::SM_FC);
  float* outst = reinterpret_cast<float*>(smem + P::SM_OUT);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int row0 = (blockIdx.x >> 1) * (2 * NB) + (int)rank * NB;  // first row of this CTA
  const int Tp = a.Tp;
  const uint8_t* my_stream = a.packed + (size_t)rank * P::STREAM_BYTES;

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    // leader: a stage is full when its own half has landed (arrive.expect_tx + bytes) AND the peer's relay arrived
    for (int s = 0; s < NG; ++s) { mbar_init(&bars.w_full[s], leader ? 2 : 1); mbar_init(&bars.w_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.x_full[i], 2); mbar_init(&bars.x_empty[i], 1); }
    mbar_init(&bars.accf_j1, 1);
    mbar_init(&bars.acce_j1, 16);
    for (int l = 0; l < 2; ++l) { mbar_init(&bars.accf_j2[l], 1); mbar_init(&bars.acce_j2[l], 8); }
    mbar_init(&bars.h0_ready, 24);
    mbar_init(&bars.h1_ready, 24);
    mbar_init(&bars.fc_ready, FC_SLOTS);
    mbar_init(&bars.fc_done, 2);
    mbar_init(&bars.l1_done, 1);
    mbar_init(&bars.h0_free, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (threadIdx.x < NB) {
    RowInfo ri;
    const int r = row0 + threadIdx.x;
    ri.src_b = -1; ri.src_f = 0; ri.scale = 0.f; ri.out_idx = 0;
    if (r < a.R) {
      row_to_unit(a.map, r, ri.src_b, ri.src_f);
      ri.scale = a.inv2[ri.src_b];
      const int bq = r / a.Fsub, fq = r - bq * a.Fsub;
      ri.out_idx = bq * 2 * a.Fsub + fq;
    }
    rows[threadIdx.x] = ri;
  }
  {  // zero the state (h_{-1} = 0, x padding)
    uint4* z = reinterpret_cast<uint4*>(smem + P::SM_X);
    const int n16 = (P::SM_FC - P::SM_X) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers, state and TMEM are ready before any cross-CTA traffic
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ================= weight producer: this CTA's half of the stream, same sequence every step.
      // Ring of NG granules; phases are tracked per barrier (bit g of the masks) because a granule's full
      // barrier is only used when the granule is the first one of a stage.
      uint32_t g = 0, empty_ph = 0;
      long long pw = 0, pi = 0; const long long p0 = clock64();
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          const int nst = (layer ? NK1 : NK0) * P::PARTS;
          for (int j2 = 0; j2 < 2; ++j2) {
            const uint8_t* src = my_stream + (layer ? (j2 ? P::OFF_L1J2 : P::OFF_L1J1) : (j2 ? P::OFF_L0J2 : P::OFF_L0J1));
            const int ng = j2 ? 1 : 2;
            for (int k = 0; k < nst; ++k, src += ng * GRAN) {
              const uint32_t g0 = g;
              const long long q0 = clock64();
              for (int i = 0; i < ng; ++i) {
                const uint32_t gi = (g0 + i) % NG;
                mbar_wait<false>(&bars.w_empty[gi], ((empty_ph >> gi) & 1) ^ 1);
                empty_ph ^= 1u << gi;
              }
              const long long q1 = clock64();
              if (elect_one()) {
                mbar_expect_tx(&bars.w_full[g0], ng * GRAN);
                for (int i = 0; i < ng; ++i)
                  bulk_g2s(smem + P::SM_W + ((g0 + i) % NG) * GRAN, src + i * GRAN, GRAN, &bars.w_full[g0]);
              }
              __syncwarp();
              pw += q1 - q0; pi += clock64() - q1;
              g = (g0 + ng) % NG;
            }
          }
        }
      }
      if (a.dbg && (blockIdx.x >> 1) == 0 && lane == 0) { a.dbg[8 + rank * 4] = clock64() - p0; a.dbg[9 + rank * 4] = pw; a.dbg[10 + rank * 4] = pi; }
    } else if (warp == 1 && !leader) {
      // ================= peer: relay "my half of this stage has landed" to the leader's stage barrier
      uint32_t g = 0, full_ph = 0;
      long long rw = 0; const long long r0 = clock64();
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          const int nst = (layer ? NK1 : NK0) * P::PARTS;
          for (int j2 = 0; j2 < 2; ++j2) {
            const int ng = j2 ? 1 : 2;
            for (int k = 0; k < nst; ++k) {
              const long long q0 = clock64();
              mbar_wait<false>(&bars.w_full[g], (full_ph >> g) & 1);
              rw += clock64() - q0;
              full_ph ^= 1u << g;
              // relaxed: the relay publishes no data of its own (the weights were written by the TMA engine into
              // this CTA's shared memory, where this CTA's tensor core reads them)
              if (elect_one()) mbar_arrive_cluster_relaxed(&bars.w_full[g], 0);
              __syncwarp();
              g = (g + ng) % NG;
            }
          }
        }
      }
      if (a.dbg && (blockIdx.x >> 1) == 0 && lane == 0) { a.dbg[16] = clock64() - r0; a.dbg[17] = rw; }
    } else if (warp == 1) {
      // ================= leader: MMA issuer for the pair (converged warp, one elected lane issues)
      uint32_t g = 0, full_ph = 0;
      const uint64_t adesc0 = desc_sw64(smem_u32(smem + P::SM_W));
      int h0_seen = 0, h1_seen = 0;
      uint32_t j1_uses = 0, j2_uses[2] = {0, 0};
      bool w_ready = false;
      ACCT(long long c_full = 0, c_acce = 0, c_h = 0, c_issue = 0; const long long c_start = clock64();)
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          ACCT(long long c0 = clock64();)
          if (layer == 0) {
            mbar_wait<false>(&bars.x_full[t & 1], (t >> 1) & 1);
            for (; h0_seen < t; ++h0_seen) mbar_wait<false>(&bars.h0_ready, h0_seen & 1);
          } else {
            for (; h0_seen < t + 1; ++h0_seen) mbar_wait<false>(&bars.h0_ready, h0_seen & 1);
            for (; h1_seen < t; ++h1_seen) mbar_wait<false>(&bars.h1_ready, h1_seen & 1);
          }
          ACCT(c_h += clock64() - c0;)
          tc_fence_after();
          const uint32_t x_addr = smem_u32(smem + P::SM_X + (t & 1) * S_KBLK);
          const uint32_t h0_cur = smem_u32(smem + P::SM_H0 + (X3 ? 0 : (t & 1) * NKH * S_KBLK));
          const uint32_t h0_prev = smem_u32(smem + P::SM_H0 + (X3 ? 0 : ((t + 1) & 1) * NKH * S_KBLK));
          const uint32_t h1_prev = smem_u32(smem + P::SM_H1);
          const uint64_t bd_a = desc_sw128(layer ? h0_cur : x_addr);
          const uint64_t bd_b = desc_sw128(layer ? h1_prev : h0_prev);
          // X3: descriptor delta from the hi copy of a k range to its lo copy (x: k + 32 inside the block)
          const uint64_t lo_a = layer ? (uint64_t)(P::LO_OFF >> 4) : 4ull;
          const uint64_t lo_b = (uint64_t)(P::LO_OFF >> 4);
          const int n_a = layer ? H / KS : 1;
          const int n_b = H / KS;
          // one weight stage (one k range of 32, 4 gates) against `nb` state operands: 2 k16 slices x 4 gates per
          // operand (8 MMAs on 4 different accumulators); ONE barrier wait per stage (own half + the peer's relay
          // arrive on the same barrier), probed early for the next stage
          auto issue_stage = [&](uint32_t d0, uint64_t bd0, uint64_t bd1, int nb, bool first, bool j2) {
            const uint32_t g0 = g, g1 = (g + 1) % NG;
            ACCT(const long long c1 = clock64();)
            if (!w_ready) mbar_wait<false>(&bars.w_full[g0], (full_ph >> g0) & 1);
            ACCT(const long long c2 = clock64(); c_full += c2 - c1;)
            full_ph ^= 1u << g0;
            tc_fence_after();
            const uint64_t ad0 = adesc0 + (uint64_t)(g0 * (GRAN >> 4));
            const uint64_t ad1 = adesc0 + (uint64_t)(g1 * (GRAN >> 4));
            if (elect_one()) {
              if (!j2) {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                  if (b < nb) {
                    const uint64_t bd = b ? bd1 : bd0;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                      const uint32_t acc = (first && b == 0 && k == 0) ? 0u : 1u;
                      tc_mma2_f16(d0 + 0 * 64, ad0 + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc256, acc);
                      tc_mma2_f16(d0 + 1 * 64, ad0 + (uint64_t)((W_SUB1 >> 4) + 2 * k), bd + (uint64_t)(2 * k), kIdesc256, acc);
                      tc_mma2_f16(d0 + 2 * 64, ad1 + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), kIdesc256, acc);
                      tc_mma2_f16(d0 + 3 * 64, ad1 + (uint64_t)((W_SUB1 >> 4) + 2 * k), bd + (uint64_t)(2 * k), kIdesc256, acc);
                    }
                  }
                }
                tc_commit2_mc(&bars.w_empty[g0], 3);  // frees the granules in both CTAs
                tc_commit2_mc(&bars.w_empty[g1], 3);
              } else {
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                  if (b < nb) {
                    const uint64_t bd = b ? bd1 : bd0;
#pragma unroll
                    for (int k = 0; k < 2; ++k)
#pragma unroll
                      for (int gt = 0; gt < 4; ++gt)
                        tc_mma2_f16(d0 + (uint32_t)gt * 32, ad0 + (uint64_t)(gt * (W_SUB2 >> 4) + 2 * k), bd + (uint64_t)(2 * k),
                                    kIdesc128, (first && b == 0 && k == 0) ? 0u : 1u);
                  }
                }
                tc_commit2_mc(&bars.w_empty[g0], 3);
              }
            }
            __syncwarp();
            g = (g0 + (j2 ? 1 : 2)) % NG;
            w_ready = mbar_test_wait(&bars.w_full[g], (full_ph >> g) & 1);  // probe the next stage early
            ACCT(c_issue += clock64() - c2;)
          };
          // one k range: single-pass = W.S; X3 = W_hi.S_hi + W_hi.S_lo (hi stage), then W_lo.S_hi (lo stage)
          auto issue_krange = [&](uint32_t d0, uint64_t bd, uint64_t lo_delta, bool first, bool j2) {
            if (X3) {
              issue_stage(d0, bd, bd + lo_delta, 2, first, j2);
              issue_stage(d0, bd, bd, 1, false, j2);
            } else {
              issue_stage(d0, bd, bd, 1, first, j2);
            }
          };
          auto run_job = [&](uint32_t d0, bool j2) {
            uint64_t bd = bd_a;
#pragma unroll 1
            for (int j = 0; j < n_a; ++j) {
              issue_krange(d0, bd, lo_a, j == 0, j2);
              bd += (j & 1) ? (uint64_t)((S_KBLK >> 4) - 4) : 4ull;
            }
            if (X3 && j2 && layer == 1) {
              // every MMA that reads h0_{t} (= h0_{it-1}) has been issued: h0 may be overwritten with h0_{it}
              if (elect_one()) tc_commit2_mc(&bars.h0_free, 3);
              __syncwarp();
            }
            bd = bd_b;
#pragma unroll 1
            for (int j = 0; j < n_b; ++j) {
              issue_krange(d0, bd, lo_b, false, j2);
              bd += (j & 1) ? (uint64_t)((S_KBLK >> 4) - 4) : 4ull;
            }
          };
          // ---- J1: M=256, accumulators at columns [0,256): gate g at g*64
          ACCT(c0 = clock64();)
          mbar_wait<false>(&bars.acce_j1, (j1_uses & 1) ^ 1);
          ACCT(c_acce += clock64() - c0;)
          tc_fence_after();
          run_job(tmem_base, false);
          if (elect_one()) tc_commit2_mc(&bars.accf_j1, 3);
          __syncwarp();
          ++j1_uses;
          // ---- J2: M=128 over the pair, accumulators at columns 256 + layer*128: gate g at g*32
          ACCT(c0 = clock64();)
          mbar_wait<false>(&bars.acce_j2[layer], (j2_uses[layer] & 1) ^ 1);
          ACCT(c_acce += clock64() - c0;)
          tc_fence_after();
          run_job(tmem_base + 256 + layer * 128, true);
          if (elect_one()) {
            tc_commit2_mc(&bars.accf_j2[layer], 3);
            if (layer == 0) {
              tc_commit2_mc(&bars.x_empty[t & 1], 3);
              if (X3 && it == 0) tc_commit2_mc(&bars.h0_free, 3);  // step 0: h0_{-1} (zeros) is read by layer 0 only
            } else {
              tc_commit2_mc(&bars.l1_done, 3);  // every layer-1 MMA of step t has read h1_{t-1}
            }
          }
          __syncwarp();
          ++j2_uses[layer];
        }
      }
      ACCT(if (a.dbg && blockIdx.x == 0 && lane == 0) {
        a.dbg[0] = clock64() - c_start; a.dbg[1] = c_full; a.dbg[2] = 0; a.dbg[3] = c_acce; a.dbg[4] = c_h; a.dbg[5] = c_issue;
      })
    } else if (warp == 2) {
      // ================= x_t gather for this CTA's 32 rows (base_model.py:35-44, model.py:98-111)
      const int nmag = 2 * a.Ns + 1;
      for (int t = 0; t < Tp; ++t) {
        mbar_wait<true>(&bars.x_empty[t & 1], ((t >> 1) & 1) ^ 1);
        uint8_t* xb = smem + P::SM_X + (t & 1) * S_KBLK;
        // source frames of step t: itself, or (fast_fullsubnet/model.py:108-129) frame 0 alone, then blocks of
        // `shrink` frames, the last one over its own length
        int f0 = t, f1 = t + 1;
        if (a.shrink > 1 && t > 0) { f0 = 1 + (t - 1) * a.shrink; f1 = min(f0 + a.shrink, a.src_T); }
        const float wmean = 1.0f / (float)(f1 - f0);
        auto put = [&](int n, float v) {
          const __half hi = __float2half_rn(v);
          *reinterpret_cast<__half*>(xb + swz128_off(n, lane)) = hi;
          if (X3) *reinterpret_cast<__half*>(xb + swz128_off(n, KS + lane)) = __float2half_rn(v - __half2float(hi));
        };
        if (a.shrink <= 1) {  // fullsubnet: one source frame per step, 4 rows of loads in flight
#pragma unroll 4
          for (int n = 0; n < NB; ++n) {
            const RowInfo ri = rows[n];
            float v = 0.f;
            if (ri.src_b >= 0 && lane < a.Ksb) {
              const size_t base = ((size_t)ri.src_b * a.src_T + t) * a.F;
              if (lane < nmag) v = a.magT[base + reflect_idx(ri.src_f + lane - a.Ns, a.F)];
              else             v = a.fbT[base + reflect_idx(ri.src_f + (lane - nmag) - a.Nf, a.F)];
              v *= a.unit_scale ? a.unit_scale[(size_t)t * a.R + row0 + n] : ri.scale;
            }
            put(n, v);
          }
        } else {
#pragma unroll 2
          for (int n = 0; n < NB; ++n) {
            const RowInfo ri = rows[n];
            float v = 0.f;
            if (ri.src_b >= 0 && lane < a.Ksb) {
              const int col = (lane < nmag) ? reflect_idx(ri.src_f + lane - a.Ns, a.F)
                                            : reflect_idx(ri.src_f + (lane - nmag) - a.Nf, a.F);
              const float* src = (lane < nmag) ? a.magT : a.fbT;
              for (int fr = f0; fr < f1; ++fr) v += src[((size_t)ri.src_b * a.src_T + fr) * a.F + col];
              v *= wmean * ri.scale;
            }
            put(n, v);
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars.x_full[t & 1], 0);
      }
    } else {
      // ================= Linear(H -> 2): sum the 12 fp32 partials of this CTA's rows, bias, stage, store
      const float fcb0 = reinterpret_cast<const float*>(a.packed + P::OFF_FCB)[0];
      const float fcb1 = reinterpret_cast<const float*>(a.packed + P::OFF_FCB)[1];
      const RowInfo ri = rows[lane];
      int staged = 0, t_stage0 = 0;
      for (int t = 0; t < Tp; ++t) {
        mbar_wait<true>(&bars.fc_ready, t & 1);
        if (t >= a.la) {
          float s0 = fcb0, s1 = fcb1;
#pragma unroll
          for (int w = 0; w < FC_SLOTS; ++w) {
            s0 += fc_part[(w * 2 + 0) * NB + lane];
            s1 += fc_part[(w * 2 + 1) * NB + lane];
          }
          if (staged == 0) t_stage0 = t - a.la;
          outst[(lane * 2 + 0) * OUT_T + staged] = act_apply(s0, a.act);
          outst[(lane * 2 + 1) * OUT_T + staged] = act_apply(s1, a.act);
          ++staged;
        }
        __syncwarp();
        if (lane == 0) { mbar_arrive_cluster(&bars.fc_done, 0); mbar_arrive_cluster(&bars.fc_done, 1); }
        if (staged == OUT_T || (t == Tp - 1 && staged > 0)) {
          if (ri.src_b >= 0) {
#pragma unroll
            for (int o = 0; o < 2; ++o) {
              float* dst = a.crm + ((size_t)ri.out_idx + (size_t)o * a.Fsub) * a.T + t_stage0;
              for (int i = 0; i < staged; ++i) dst[i] = outst[(lane * 2 + o) * OUT_T + i];
            }
          }
          staged = 0;
        }
      }
    }
  } else {
    // ================= epilogue warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 152;");
    const int wg = (warp - 4) >> 2;   // 0: J1 rows 0..31, 1: J1 rows 32..63, 2: J2
    const int q = warp & 3;           // TMEM lane quadrant
    const bool is_j2 = wg == 2;
    const int L = q * 32 + lane;      // TMEM lane
    const int u = is_j2 ? 256 + (int)rank * 64 + (L & 63) : (int)rank * 128 + L;   // hidden unit
    const uint32_t dest = is_j2 ? (uint32_t)(L >> 6) : (uint32_t)wg;               // CTA that owns the rows
    const float* bias_g = reinterpret_cast<const float*>(a.packed + P::OFF_BIAS);
    float b0[4], b1[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { b0[g] = bias_g[g * H + u]; b1[g] = bias_g[4 * H + g * H + u]; }
    const float wfc0 = reinterpret_cast<const float*>(a.packed + P::OFF_FCW)[u];
    const float wfc1 = reinterpret_cast<const float*>(a.packed + P::OFF_FCW)[H + u];
    float c0[NB], c1[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i) c0[i] = c1[i] = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int kbu = u >> 6, chunk = (u & 63) >> 3, el = u & 7;
    // state / partial-sum addresses in the destination CTA (distributed shared memory)
    const uint32_t dst_base = mapa(smem_u32(smem), dest);
    const int slot = (int)rank * 6 + (is_j2 ? 4 + (q & 1) : q);
    const uint32_t part_addr = dst_base + P::SM_FC + (uint32_t)slot * 2 * NB * 4;
    uint64_t* accf = is_j2 ? nullptr : &bars.accf_j1;
    uint32_t job = 0, job2[2] = {0, 0};
    for (int it = 0; it <= Tp; ++it) {
      for (int layer = 0; layer < 2; ++layer) {
        const int t = it - layer;
        if (t < 0 || t >= Tp) continue;
        uint32_t tacc;
        if (is_j2) {
          mbar_wait<true>(&bars.accf_j2[layer], job2[layer] & 1);
          ++job2[layer];
          tacc = tmem_base + lane_addr + 256 + layer * 128;
        } else {
          mbar_wait<true>(accf, job & 1);
          ++job;
          tacc = tmem_base + lane_addr + (uint32_t)wg * 32;
        }
        tc_fence_after();
        if (layer == 1 && t >= 1) mbar_wait<true>(&bars.fc_done, (t - 1) & 1);  // both Linear warps read step t-1
        const uint32_t gstep = is_j2 ? 32u : 64u;
        const uint32_t hb = dst_base + (layer ? P::SM_H1 : P::SM_H0 + (X3 ? 0u : (uint32_t)((t & 1) * NKH * S_KBLK))) +
                            (uint32_t)(kbu * S_KBLK + el * 2);
        // held-back state: layer 1 keeps h1_t until every layer-1 MMA of this step is done (h1 is single-buffered);
        // X3 does the same for h0_t (h0_free) and also carries the lo halves
        const bool hold = X3 || layer == 1;
        __half2 hst[NB / 2];
        __half2 lst[X3 ? NB / 2 : 1];
#pragma unroll
        for (int j0 = 0; j0 < NB; j0 += 8) {
          float gi[8], gf[8], gg[8], go[8];
          tc_ld8(tacc + 0 * gstep + j0, gi);
          tc_ld8(tacc + 1 * gstep + j0, gf);
          tc_ld8(tacc + 2 * gstep + j0, gg);
          tc_ld8(tacc + 3 * gstep + j0, go);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float cp, bi, bf, bg, bo;
            if (layer == 0) { cp = c0[j0 + j]; bi = b0[0]; bf = b0[1]; bg = b0[2]; bo = b0[3]; }
            else            { cp = c1[j0 + j]; bi = b1[0]; bf = b1[1]; bg = b1[2]; bo = b1[3]; }
            const float cn = gate_sigmoid<X3>(gf[j] + bf) * cp + gate_sigmoid<X3>(gi[j] + bi) * gate_tanh<X3>(gg[j] + bg);
            if (layer == 0) c0[j0 + j] = cn; else c1[j0 + j] = cn;
            const float h = gate_sigmoid<X3>(go[j] + bo) * gate_tanh<X3>(cn);
            // row n = j0 + j of the destination CTA: (n>>3)*1024 + (n&7)*128 + ((chunk ^ (n&7)) << 4)
            if (!hold)
              st_cluster_b16(hb + (uint32_t)((j0 >> 3) * 1024 + j * 128 + ((chunk ^ j) << 4)), __float2half_rn(h));
            go[j] = h;
          }
          if (hold) {
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              const __half2 hi = __floats2half2_rn(go[j], go[j + 1]);
              hst[(j0 + j) >> 1] = hi;
              if (X3) lst[(j0 + j) >> 1] = __floats2half2_rn(go[j] - __low2float(hi), go[j + 1] - __high2float(hi));
            }
          }
          if (layer == 1) {
            // Linear(H->2) in fp32: 2 outputs x 8 rows, summed over the warp's 32 hidden units
            float v[16];
#pragma unroll
            for (int j = 0; j < 8; ++j) { v[j] = go[j] * wfc0; v[8 + j] = go[j] * wfc1; }
#pragma unroll
            for (int half = 8, off = 16; half >= 1; half >>= 1, off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < half; ++i) {
                const float send = up ? v[i] : v[i + half];
                const float keep = up ? v[i + half] : v[i];
                v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
            if ((lane & 1) == 0)
              st_cluster_f32(part_addr + (uint32_t)((((lane >> 4) & 1) * NB + j0 + ((lane >> 1) & 7)) * 4), v[0]);
          }
        }
        tc_fence_before();
        if (hold) {
          // accumulators drained (and Linear partials written): release them, then wait until the last MMA that
          // reads the previous state has completed before overwriting it
          __syncwarp();
          if (lane == 0) {
            if (layer == 1) fence_release_cluster();  // Linear partials (st.shared::cluster) before the arrives below
            mbar_arrive_cluster_relaxed(is_j2 ? &bars.acce_j2[layer] : &bars.acce_j1, 0);
            if (layer == 1) mbar_arrive_cluster_relaxed(&bars.fc_ready, dest);
          }
          mbar_wait<true>(layer ? &bars.l1_done : &bars.h0_free, t & 1);
#pragma unroll
          for (int n = 0; n < NB; n += 2) {
            const uint32_t o0 = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((chunk ^ (n & 7)) << 4));
            const uint32_t o1 = (uint32_t)(((n + 1) >> 3) * 1024 + ((n + 1) & 7) * 128 + ((chunk ^ ((n + 1) & 7)) << 4));
            st_cluster_b16(hb + o0, __low2half(hst[n >> 1]));
            st_cluster_b16(hb + o1, __high2half(hst[n >> 1]));
            if (X3) {
              st_cluster_b16(hb + P::LO_OFF + o0, __low2half(lst[n >> 1]));
              st_cluster_b16(hb + P::LO_OFF + o1, __high2half(lst[n >> 1]));
            }
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          fence_release_cluster();  // h (st.shared::cluster) of the whole warp before the arrives below
          if (!hold) mbar_arrive_cluster_relaxed(is_j2 ? &bars.acce_j2[layer] : &bars.acce_j1, 0);
          mbar_arrive_cluster_relaxed(layer ? &bars.h1_ready : &bars.h0_ready, 0);
        }
      }
    }
  }

  // ---------------- teardown
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

}  // namespace tc2

bool sb_tc2_supported(const fsn_model_desc* d) {
  static int pair_env = -1;
  if (pair_env < 0) {
    const char* e = getenv("FSN_TC_PAIR");
    pair_env = e ? atoi(e) : 1;  // default: CTA-pair kernel (FSN_TC_PAIR=0 selects the single-CTA kernel)
  }
  const int Ksb = (2 * d->sb_num_neighbors + 1) + (2 * d->fb_num_neighbors + 1);
  if (d->precision == FSN_PREC_F16X3_TC) return d->sb_hidden == tc2::H && Ksb <= tc2::KS;  // pair kernel only
  return pair_env != 0 && d->sb_hidden == tc2::H && Ksb <= tc2::KS;
}

size_t sb_tc2_packed_bytes(bool x3) { return x3 ? tc2::Plan<true>::PACKED_BYTES : tc2::Plan<false>::PACKED_BYTES; }

int sb_tc2_pack_raw(const fsn_seq_weights* sb, int Ksb, int fc_out, void* packed, cudaStream_t st, bool x3) {
  if (x3)
    tc2::pack2_kernel<true><<<148 * 4, 256, 0, st>>>(sb->w_ih[0], sb->w_hh[0], sb->w_ih[1], sb->w_hh[1], sb->b_ih[0],
                                                     sb->b_hh[0], sb->b_ih[1], sb->b_hh[1], sb->fc_w, sb->fc_b, Ksb,
                                                     fc_out, (uint8_t*)packed);
  else
    tc2::pack2_kernel<false><<<148 * 4, 256, 0, st>>>(sb->w_ih[0], sb->w_hh[0], sb->w_ih[1], sb->w_hh[1], sb->b_ih[0],
                                                      sb->b_hh[0], sb->b_ih[1], sb->b_hh[1], sb->fc_w, sb->fc_b, Ksb,
                                                      fc_out, (uint8_t*)packed);
  FSN_CHECK_LAUNCH("sb pack2_kernel");
  return FSN_OK;
}

bool sb_tc2_enabled() {
  fsn_model_desc d;
  memset(&d, 0, sizeof(d));
  d.sb_hidden = tc2::H; d.sb_num_neighbors = 0; d.fb_num_neighbors = 0;
  return sb_tc2_supported(&d);
}

int sb_tc2_pack(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, cudaStream_t st) {
  const int Ksb = (2 * d->sb_num_neighbors + 1) + (2 * d->fb_num_neighbors + 1);
  return sb_tc2_pack_raw(sb, Ksb, 2, packed, st, d->precision == FSN_PREC_F16X3_TC);
}

template <bool X3>
static int tc2_launch(const tc2::KArgs& a, cudaStream_t st) {
  const size_t smem = tc2::Plan<X3>::SM_TOTAL + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(tc2::sb_lstm_tc2_kernel<X3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem), "sb_lstm_tc2 smem attr");
  if (rc) return rc;
  const int pairs = cdiv(a.R, 2 * tc2::NB);
  tc2::sb_lstm_tc2_kernel<X3><<<2 * pairs, tc2::NTHREADS, smem, st>>>(a);
  FSN_CHECK_LAUNCH("sb_lstm_tc2_kernel");
  return FSN_OK;
}

int sb_tc2_forward(const SbTcArgs& s, cudaStream_t st) {
  tc2::KArgs a;
  a.packed = (const uint8_t*)s.packed;
  a.magT = s.magT; a.fbT = s.fbT; a.inv2 = s.inv2; a.unit_scale = s.shrink > 1 ? nullptr : s.unit_scale; a.crm = s.crm;
  a.R = s.map.B * s.map.Fsub; a.F = s.F; a.Tp = s.steps > 0 ? s.steps : s.Tp; a.la = s.la; a.T = a.Tp - s.la;
  a.src_T = s.Tp; a.shrink = s.shrink > 1 ? s.shrink : 1;
  a.Ns = s.Ns; a.Nf = s.Nf; a.Ksb = (2 * s.Ns + 1) + (2 * s.Nf + 1); a.act = s.act;
  a.Fsub = s.map.Fsub; a.map = s.map;
  a.dbg = nullptr;
  static long long* dbg_buf = nullptr;
  if (getenv("FSN_TC_TRACE")) {
    if (!dbg_buf) { cudaMalloc(&dbg_buf, 32 * sizeof(long long)); cudaMemset(dbg_buf, 0, 32 * sizeof(long long)); }
    a.dbg = dbg_buf;
  }
  const int rc = s.x3 ? tc2_launch<true>(a, st) : tc2_launch<false>(a, st);
  if (rc) return rc;
  if (a.dbg) {
    long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc2 producers] leader: total %lld wait-empty %lld issue %lld | peer: total %lld wait-empty %lld issue %lld | "
                    "relay: total %lld wait-full %lld\n", h[8], h[9], h[10], h[12], h[13], h[14], h[16], h[17]);
    fprintf(stderr, "[tc2 leader cycles] total %lld | wait own stage %lld | wait peer stage %lld | wait acc drained %lld | "
                    "wait state ready %lld | issue+commit %lld  (Tp=%d)\n", h[0], h[1], h[2], h[3], h[4], h[5], a.Tp);
  }
  return FSN_OK;
}

}  // namespace fsn
