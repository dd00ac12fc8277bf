// Sub-band LSTM stack, CLUSTER variant (single fp16 pass): three CTA pairs (6 SMs = one CPC) share 128 sub-band units.
// EXPERIMENTAL, opt-in (FSN_TC_CLUSTER4=1, precision f16_tc): validated on hardware (tests/test_gpu_parity.py), faster
// than the pair kernel for <= 21 clusters (B <= 10 clips: 3.7 vs 5.3 ms per launch) but slower at full batch (105 vs
// 83 ms) - see "What was measured" below.
//
// Same mathematics and reference rows as fsn_subband_tc2.cu (recipes/dns_interspeech_2020/fullsubnet/model.py:98-135).
// Why another mapping: the pair kernel is bound by the shared-memory port - every weight byte is written into SMEM by
// the TMA engine and read once by the tensor core, and it meets only N = 64 rows (9 KB of port traffic per 2-CTA MMA =
// 72 cycles for 32 cycles of math).  N is pinned by TMEM (gates of one layer-step for N rows) and by the register file
// (fp32 cell state) as long as ONE pair holds all 384 hidden units.  Here a cluster of NPAIR = 3 cta_group::2 pairs owns
// 128 rows; every CTA holds 64 hidden units of both layers (256 gate rows = two 128-lane tiles), every MMA is M = 256
// (pair) x N = 128 x K = 16: per MMA an SM moves 4 KB (TMA write) + 4 KB (A read) + 2 KB (B read) for 64 cycles of
// math, i.e. half the weight bytes per FLOP.
//   rank = 2*pair + r.  CTA (pair, r): units [64*rank, +64); state (x_t, h0, h1 as UMMA B operand) of rows
//   [64 r, 64 r + 64) of the cluster's 128 rows - every pair holds a full copy of the state (each computes other gate
//   rows over ALL rows), so every epilogue thread sends h to one CTA of each pair.
// TMEM lane order of a tile (32 units x 4 gates): inside each 32-lane quadrant q, lanes 0-7 = gate i of units 8q..8q+7,
// 8-15 = f, 16-23 = g, 24-31 = o.  tcgen05.ld.16x256b (thread t: lanes t/4 and t/4 + 8, columns 2(t%4) + {0,1} of
// every 8-column block; measured with tools/probe_ld16.cu) at lane offsets 0 and 16 then hands thread t all four
// gates of unit 8q + t/4 for its columns: no gate exchange, c stays in 32 registers per layer.  An 8 x 8 shuffle
// transpose collects the 8 units of a quadrant for one row, so h leaves the thread as ONE 16-byte st.async per
// destination CTA, whose completion is counted as tx bytes on an mbarrier of the destination (no fence, no arrive; the
// peer's state warp relays "my half has landed" to the pair leader).
// Four 128-column TMEM slots: layer 0 tiles in slots 0-1, layer 1 in 2-3.  h0 is double-buffered, h1 single-buffered
// (h1_t is held in registers until a tcgen05.commit of ALL pairs says the last MMA reading h1_{t-1} has completed); x_t
// sits in 4 KB 64B-swizzled blocks, which leaves an 8-granule weight ring (stages of 1-2 granules: one barrier wait per
// 2-4 MMAs - with one wait per 2 MMAs the issuing thread lost 135 cycles per stage).
// Linear(384 -> 2): every CTA ends a step with the complete h1_t (fp16) of its 64 rows in shared memory, so one warp
// of ranks 0 and 1 evaluates the two dot products from there (no partial-sum traffic).  The LSTM cell uses 7 MUFU
// operations instead of 10 (lstm_cell7: the gate factors share reciprocals).
//
// What was measured (B200, 256 x 4 s): as a 4-CTA cluster (2 pairs, 96 units per CTA) only ~26 clusters were resident
// (a cluster must sit inside one CPC of 3 TPCs) - 92 ms; as a 6-CTA cluster 22 are resident (132 SMs), one cluster
// lifetime is 3.7 ms alone / 4.6 ms under load = 14.6-18.3 us per LSTM step for 148 MMAs (tensor time ~5.5 us): the step
// is a chain of hand-offs (commit -> epilogue -> st.async -> relay -> leader) that the 2.6 us of the other layer's MMAs
// cannot cover.  Findings on the way (cycle accounting, FSN_TC4_ACCOUNT=1): a gather with the multiply inside the
// (lane-dependent) branch serialises its loads (43 k cycles per step -> 14 k branch-free); gating the START of layer 1
// on the Linear warp instead of the l1_done commit cost 14 k cycles per step; remote 2-byte stores + cluster fences cost
// 5.3 k cycles per tile against 0.6-1.5 k for 16-byte st.async.
//
// Warp roles per CTA (384 threads): 0 = weight producer, 1 = MMA issuer (pair leader) / stage relay (peer),
// 2 = x gather + TMEM alloc, 3 = state warp (expect_tx, relay, Linear + output on ranks 0, 1), 4-11 = epilogue:
// warpgroup j = tile j, warp%4 = quadrant.
#include <stdlib.h>
#include <string.h>

#include "fsn_internal.cuh"
#include "fsn_tc_ptx.cuh"

#ifndef FSN_TC4_ACCOUNT
#define FSN_TC4_ACCOUNT 0   // 1: cycle accounting of the leader's MMA warp and the epilogue warps (diagnostic builds, FSN_TC_TRACE=1)
#endif
#if FSN_TC4_ACCOUNT
#define ACCT4(...) __VA_ARGS__
#else
#define ACCT4(...)
#endif

namespace fsn {
namespace tc4 {
using namespace ptx;

constexpr int H = 384;
constexpr int NBR = 64;                // rows per CTA (B-operand half of N = 128)
constexpr int NCL = 2 * NBR;           // rows per cluster
constexpr int NPAIR = 3;               // CTA pairs per cluster: 6 SMs = the 3 TPCs of one CPC (a 4-CTA cluster leaves a
                                       // third of every CPC idle: measured 26 instead of 37 resident clusters)
constexpr int NCTA = 2 * NPAIR;
constexpr int UC = H / NCTA;           // 64 hidden units per CTA
constexpr int TU = 32;                 // units per 128-lane tile (x 4 gates)
constexpr int TILES = UC / TU;         // 2 tiles per layer
constexpr int KB = 64, KS = 32;
constexpr int S_KBLK = NBR * KB * 2;   // 8192 B: one 64-k block of the state operand
constexpr int NKH = H / KB;            // 6
constexpr int W_ST = 128 * KS * 2;     // 8192 B: one stage = one tile x one k range of 32
constexpr int NK0 = 1 + H / KS;        // 13
constexpr int NK1 = 2 * H / KS;        // 24
constexpr int NSTG = 8;
constexpr int X_BLK = NBR * KS * 2;     // 4096 B: x_t as a [64 rows x 32 k] 64B-swizzled block
constexpr int OUT_T = 8;
constexpr int NTHREADS = 128 + TILES * 128;   // 4 service warps + one epilogue warpgroup per tile

constexpr size_t STREAM_BYTES = (size_t)TILES * (NK0 + NK1) * W_ST;   // per rank: [L0: tile 0 stages, 1, 2][L1: ...]
constexpr size_t OFF_L1 = (size_t)TILES * NK0 * W_ST;
constexpr size_t OFF_BIAS = NCTA * STREAM_BYTES;
constexpr size_t OFF_FCW = OFF_BIAS + (size_t)2 * 4 * H * sizeof(float);
constexpr size_t OFF_FCB = OFF_FCW + (size_t)2 * H * sizeof(float);
constexpr size_t PACKED_BYTES = OFF_FCB + 256;

constexpr uint32_t SM_W = 0;
constexpr uint32_t SM_X = SM_W + NSTG * W_ST;
constexpr uint32_t SM_H0 = SM_X + 2 * X_BLK;
constexpr uint32_t SM_H1 = SM_H0 + 2 * NKH * S_KBLK;   // h0 double-buffered, h1 single (held in registers, see l1_done)
constexpr uint32_t SM_FCW = SM_H1 + NKH * S_KBLK;                     // [2][H] float
constexpr uint32_t SM_OUT = SM_FCW + 2 * H * 4;                       // [NBR][2][OUT_T] float
constexpr uint32_t SM_ROWS = SM_OUT + NBR * 2 * OUT_T * 4;
constexpr uint32_t SM_BARS = SM_ROWS + NBR * 16;
constexpr uint32_t SM_TOTAL = SM_BARS + 512;
static_assert(SM_TOTAL + 1024 <= 232448, "shared-memory plan exceeds 227 KB");

constexpr uint32_t kIdesc = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);  // f16 x f16 -> f32, M = 256, N = 128

// ---------------------------------------------------------------- weight packer
// rank -> [L0 tile 0: NK0 stages][L0 tile 1][L0 tile 2][L1 tile 0: NK1 stages]...; stage = [128 lanes x 32 k] K-major, 64B swizzle
__global__ void pack4_kernel(const float* __restrict__ wih0, const float* __restrict__ whh0, const float* __restrict__ wih1,
                             const float* __restrict__ whh1, const float* __restrict__ bih0, const float* __restrict__ bhh0,
                             const float* __restrict__ bih1, const float* __restrict__ bhh1, const float* __restrict__ fcw,
                             const float* __restrict__ fcb, int Ksb, int fc_out, uint8_t* __restrict__ out) {
  const size_t chunks_per_rank = STREAM_BYTES / 16;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < NCTA * chunks_per_rank; i += (size_t)gridDim.x * blockDim.x) {
    const int rank = (int)(i / chunks_per_rank);
    size_t off = (i % chunks_per_rank) * 16;
    int layer = 0;
    if (off >= OFF_L1) { layer = 1; off -= OFF_L1; }
    const int nk = layer ? NK1 : NK0;
    const int stage = (int)(off / W_ST);
    const int tile = stage / nk, kr = stage % nk;
    const int in_st = (int)(off % W_ST);
    // invert swz64_off: 512-byte groups of 8 rows, 64-byte rows, 16-byte chunks XOR-ed with (row>>1)&3
    const int grp = in_st / 512, rr = (in_st % 512) / 64, cx = (in_st % 64) / 16;
    const int lane = grp * 8 + rr;                       // TMEM lane of the tile
    const int c = cx ^ ((lane >> 1) & 3);
    const int q = lane >> 5, rho = lane & 31;
    const int gate = rho >> 3, unit = rank * UC + tile * TU + q * 8 + (rho & 7);
    const int wrow = gate * H + unit;
    __half v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int kk = c * 8 + e;
      float w = 0.f;
      if (layer == 0) {
        if (kr == 0) { if (kk < Ksb) w = wih0[(size_t)wrow * Ksb + kk]; }
        else w = whh0[(size_t)wrow * H + (kr - 1) * KS + kk];
      } else {
        const int k = kr * KS + kk;
        w = (k < H) ? wih1[(size_t)wrow * H + k] : whh1[(size_t)wrow * H + (k - H)];
      }
      v[e] = __float2half_rn(w);
    }
    *reinterpret_cast<uint4*>(out + (size_t)rank * STREAM_BYTES + (i % chunks_per_rank) * 16) = *reinterpret_cast<const uint4*>(v);
  }
  float* bias = reinterpret_cast<float*>(out + OFF_BIAS);
  float* pfcw = reinterpret_cast<float*>(out + OFF_FCW);
  float* pfcb = reinterpret_cast<float*>(out + OFF_FCB);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 4 * H; i += gridDim.x * blockDim.x) {
    bias[i] = bih0[i] + bhh0[i];
    bias[4 * H + i] = bih1[i] + bhh1[i];
  }
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * H; i += gridDim.x * blockDim.x)
    pfcw[i] = (i < fc_out * H) ? fcw[i] : 0.f;
  if (blockIdx.x == 0 && threadIdx.x < 2) pfcb[threadIdx.x] = ((int)threadIdx.x < fc_out) ? fcb[threadIdx.x] : 0.f;
}

struct Bars {
  uint64_t w_full[NSTG], w_empty[NSTG];
  uint64_t x_full[2], x_empty[2];
  uint64_t accf[4], acce[4];     // MMA -> epilogue (commit, multicast to the pair) / epilogue (both CTAs) -> pair leader
  uint64_t h0_tx, h1_tx;         // every CTA: the 48 KB of h_t of this CTA's 64 rows have landed (st.async tx bytes)
  uint64_t h0_peer, h1_peer;     // pair leader: the peer CTA's h_t has landed (relayed by the peer's state warp)
  uint64_t l1_done;              // both leaders' commits (multicast to the cluster): layer-1 MMAs of the step are complete
  uint64_t fc_done;              // Linear warps of ranks 0, 1 -> both leaders: h1_t has been consumed
  uint32_t tmem_base;
};
static_assert(sizeof(Bars) <= 512, "barrier block too large");

struct RowInfo { int src_b, src_f; float scale; int out_idx; };

struct KArgs {
  const uint8_t* packed;
  const float* magT; const float* fbT; const float* inv2; const float* unit_scale;
  float* crm;
  int R, F, Tp, la, T, Ns, Nf, Ksb, act, Fsub;
  long long* dbg;  // FSN_TC_TRACE: cycle accounting of cluster 0 (leader MMA warp, producer)
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
// 16 bytes into (remote) shared memory, completion signalled as 16 tx-bytes on an mbarrier of the SAME destination CTA:
// the writer does not wait, fence or arrive - the destination's waiters see the data when the phase completes
__device__ __forceinline__ void st_async_v4(uint32_t addr, uint4 v, uint32_t mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(addr),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar)
               : "memory");
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// LSTM cell with 7 MUFU operations instead of 10 (the epilogue is MUFU-bound): the three sigmoid/tanh factors of the
// cell update share ONE reciprocal, sigma(o) tanh(c) another:
//   c' = c/(1+ef) + (1-eg)/((1+ei)(1+eg)) = [c (1+ei)(1+eg) + (1-eg)(1+ef)] / [(1+ef)(1+ei)(1+eg)],  e_x = exp(-x), eg = exp(-2g)
//   h  = (1-ec) / ((1+eo)(1+ec)),  ec = exp(-2c')
// Pre-activations are clamped to +-25 (sigma saturates to 1e-11 there) so that the products stay inside fp32.
__device__ __forceinline__ float lstm_cell7(float xi, float xf, float xg, float xo, float& c) {
  const float L2E = 1.4426950408889634f;
  xi = fminf(fmaxf(xi, -25.f), 25.f); xf = fminf(fmaxf(xf, -25.f), 25.f);
  xg = fminf(fmaxf(xg, -12.5f), 12.5f); xo = fminf(fmaxf(xo, -25.f), 25.f);
  const float ei = ex2_approx(-L2E * xi), ef = ex2_approx(-L2E * xf), eg = ex2_approx(-2.f * L2E * xg), eo = ex2_approx(-L2E * xo);
  const float A = 1.f + ei, F = 1.f + ef, G = 1.f + eg;
  const float AG = A * G;
  const float cn = fmaf(c, AG, (1.f - eg) * F) * rcp_approx(F * AG);
  c = cn;
  const float cc = fminf(fmaxf(cn, -12.5f), 12.5f);
  const float ec = ex2_approx(-2.f * L2E * cc);
  return (1.f - ec) * rcp_approx((1.f + eo) * (1.f + ec));
}
// 8 x 8 transpose across the 8 lanes {t : t % 4 == cp} of a warp (lane = 4 * u8 + cp): in: a[j] = value of MY unit u8
// for cell j; out: a[i] = value of unit i for cell u8.  Three butterfly stages on the unit bits (lane bits 4, 3, 2).
__device__ __forceinline__ void transpose8_units(float (&a)[8], int u8) {
#pragma unroll
  for (int k = 2; k >= 0; --k) {
    const int bit = 1 << k;
    const bool beta = (u8 >> k) & 1;
#pragma unroll
    for (int idx = 0; idx < 8; ++idx) {
      if (idx & bit) continue;
      const float send = beta ? a[idx] : a[idx | bit];
      const float recv = __shfl_xor_sync(0xffffffffu, send, 4 << k);
      if (beta) a[idx] = recv; else a[idx | bit] = recv;
    }
  }
}
__device__ __forceinline__ uint4 pack8_half(const float (&a)[8]) {
  uint4 r;
  __half2 h0 = __floats2half2_rn(a[0], a[1]), h1 = __floats2half2_rn(a[2], a[3]);
  __half2 h2 = __floats2half2_rn(a[4], a[5]), h3 = __floats2half2_rn(a[6], a[7]);
  r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
  r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
  return r;
}
// 32 columns of two TMEM lanes (t/4, t/4 + 8 of the 16 lanes at taddr): v[4b + 0,1] = lane A cols 8b + 2(t%4) + {0,1},
// v[4b + 2,3] = lane B same columns
__device__ __forceinline__ void tc_ld16x256_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__global__ void __cluster_dims__(NCTA, 1, 1) __launch_bounds__(NTHREADS, 1) sb_lstm_tc4_kernel(const KArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  Bars& bars = *reinterpret_cast<Bars*>(smem + SM_BARS);
  RowInfo* rows = reinterpret_cast<RowInfo*>(smem + SM_ROWS);
  float* outst = reinterpret_cast<float*>(smem + SM_OUT);
  float* fcw_s = reinterpret_cast<float*>(smem + SM_FCW);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const uint32_t pair = rank >> 1, r_in = rank & 1;
  const bool leader = r_in == 0;
  const uint32_t lead_rank = pair * 2;
  const uint16_t pair_mask = (uint16_t)(3u << (2 * pair));
  const int row0 = (blockIdx.x / NCTA) * NCL + (int)r_in * NBR;  // first row of this CTA's state half
  const int Tp = a.Tp;
  const uint8_t* my_stream = a.packed + (size_t)rank * STREAM_BYTES;

  // ---------------- one-time setup
  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTG; ++s) { mbar_init(&bars.w_full[s], leader ? 2 : 1); mbar_init(&bars.w_empty[s], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bars.x_full[i], 2); mbar_init(&bars.x_empty[i], 1); }
    for (int i = 0; i < 4; ++i) { mbar_init(&bars.accf[i], 1); mbar_init(&bars.acce[i], 8); }
    mbar_init(&bars.h0_tx, 1);
    mbar_init(&bars.h1_tx, 1);
    mbar_init(&bars.h0_peer, 1);
    mbar_init(&bars.h1_peer, 1);
    mbar_init(&bars.l1_done, NPAIR);
    mbar_init(&bars.fc_done, 2);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&bars.tmem_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  if (threadIdx.x < NBR) {
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
  for (int i = threadIdx.x; i < 2 * H; i += blockDim.x) fcw_s[i] = reinterpret_cast<const float*>(a.packed + OFF_FCW)[i];
  {  // zero the state (h_{-1} = 0, x padding)
    uint4* z = reinterpret_cast<uint4*>(smem + SM_X);
    const int n16 = (SM_FCW - SM_X) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars.tmem_base;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
      // ================= weight producer: this CTA's slice, same sequence every step.  Ring of NSTG granules of
      // 8 KB (one tile x one k range); a STAGE = 1 granule (x) or 2 granules (64 k of h) signalled on the full barrier
      // of its first granule, so the consumer pays one barrier wait per 2-4 MMAs.  Phases are tracked per barrier.
      uint32_t g = 0, empty_ph = 0;
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          const int nk = layer ? NK1 : NK0;
          const uint8_t* src = my_stream + (layer ? OFF_L1 : 0);
          for (int tile = 0; tile < TILES; ++tile) {
            for (int k = 0; k < nk;) {
              const int ng = (layer == 0 && k == 0) ? 1 : 2;
              const uint32_t g0 = g;
              for (int i = 0; i < ng; ++i) {
                const uint32_t gi = (g0 + i) % NSTG;
                mbar_wait<false>(&bars.w_empty[gi], ((empty_ph >> gi) & 1) ^ 1);
                empty_ph ^= 1u << gi;
              }
              if (elect_one()) {
                mbar_expect_tx(&bars.w_full[g0], ng * W_ST);
                for (int i = 0; i < ng; ++i)
                  bulk_g2s(smem + SM_W + ((g0 + i) % NSTG) * W_ST, src + i * W_ST, W_ST, &bars.w_full[g0]);
              }
              __syncwarp();
              src += ng * W_ST;
              k += ng;
              g = (g0 + ng) % NSTG;
            }
          }
        }
      }
    } else if (warp == 1 && !leader) {
      // ================= peer: relay "my half of this stage has landed" to the pair leader's stage barrier
      uint32_t g = 0, full_ph = 0;
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          const int nk = layer ? NK1 : NK0;
          for (int tile = 0; tile < TILES; ++tile) {
            for (int k = 0; k < nk;) {
              const int ng = (layer == 0 && k == 0) ? 1 : 2;
              mbar_wait<false>(&bars.w_full[g], (full_ph >> g) & 1);
              full_ph ^= 1u << g;
              if (elect_one()) mbar_arrive_cluster_relaxed(&bars.w_full[g], lead_rank);
              __syncwarp();
              k += ng;
              g = (g + ng) % NSTG;
            }
          }
        }
      }
    } else if (warp == 1) {
      // ================= pair leader: MMA issuer (converged warp, one elected lane issues)
      uint32_t g = 0, full_ph = 0, seq = 0;
      bool w_ready = false;
      const uint64_t adesc0 = desc_sw64(smem_u32(smem + SM_W));
      int h0_seen = 0, h1_seen = 0;
      ACCT4(long long c_state = 0, c_acce = 0, c_full = 0, c_issue = 0, c_x = 0, c_h0 = 0; const long long c_start = clock64();)
      for (int it = 0; it <= Tp; ++it) {
        for (int layer = 0; layer < 2; ++layer) {
          const int t = it - layer;
          if (t < 0 || t >= Tp) continue;
          ACCT4(long long q0 = clock64();)
          auto wait_h0 = [&](int upto) {
            for (; h0_seen < upto; ++h0_seen) {
              mbar_wait<false>(&bars.h0_tx, h0_seen & 1);
              mbar_wait<false>(&bars.h0_peer, h0_seen & 1);
            }
          };
          if (layer == 0) {
            mbar_wait<false>(&bars.x_full[t & 1], (t >> 1) & 1);
            ACCT4(const long long q1 = clock64(); c_x += q1 - q0;)
            wait_h0(t);
            ACCT4(c_h0 += clock64() - q1;)
          } else {
            wait_h0(t + 1);
            ACCT4(c_h0 += clock64() - q0;)
            for (; h1_seen < t; ++h1_seen) {
              mbar_wait<false>(&bars.h1_tx, h1_seen & 1);
              mbar_wait<false>(&bars.h1_peer, h1_seen & 1);
            }
          }
          ACCT4(c_state += clock64() - q0;)
          tc_fence_after();  // (st.async writes, like TMA writes, are observed through the mbarrier: no proxy fence)
          const uint32_t x_addr = smem_u32(smem + SM_X + (t & 1) * X_BLK);
          const uint32_t h0_cur = smem_u32(smem + SM_H0 + (t & 1) * NKH * S_KBLK);
          const uint32_t h0_prev = smem_u32(smem + SM_H0 + ((t + 1) & 1) * NKH * S_KBLK);
          const uint32_t h1_prev = smem_u32(smem + SM_H1);
          const uint64_t bd_a = layer ? desc_sw128(h0_cur) : desc_sw64(x_addr);
          const uint64_t bd_b = desc_sw128(layer ? h1_prev : h0_prev);
          const int n_a = layer ? H / KS : 1, n_b = H / KS;
          for (int tile = 0; tile < TILES; ++tile, ++seq) {
            const uint32_t slot = seq & 3;
            ACCT4(q0 = clock64();)
            mbar_wait<false>(&bars.acce[slot], ((seq >> 2) & 1) ^ 1);
            ACCT4(c_acce += clock64() - q0;)
            tc_fence_after();
            const uint32_t d = tmem_base + slot * 128;
            bool first = true;
            // one stage = ng granules (k ranges of 32) of this tile against consecutive k ranges of the state operand
            auto issue = [&](uint64_t bd, int ng, int j0) {
              const uint32_t g0 = g;
              ACCT4(const long long w0 = clock64();)
              if (!w_ready) mbar_wait<false>(&bars.w_full[g0], (full_ph >> g0) & 1);
              ACCT4(const long long w1 = clock64(); c_full += w1 - w0;)
              full_ph ^= 1u << g0;
              tc_fence_after();
              if (elect_one()) {
                uint64_t b = bd;
                for (int i = 0; i < ng; ++i) {
                  const uint32_t gi = (g0 + i) % NSTG;
                  const uint64_t ad = adesc0 + (uint64_t)(gi * (W_ST >> 4));
                  tc_mma2_f16(d, ad, b, kIdesc, first ? 0u : 1u);
                  tc_mma2_f16(d, ad + 2ull, b + 2ull, kIdesc, 1u);
                  tc_commit2_mc(&bars.w_empty[gi], pair_mask);
                  first = false;
                  b += ((j0 + i) & 1) ? (uint64_t)((S_KBLK >> 4) - 4) : 4ull;
                }
              }
              __syncwarp();
              first = false;
              g = (g0 + ng) % NSTG;
              w_ready = mbar_test_wait(&bars.w_full[g], (full_ph >> g) & 1);  // probe the next stage early
              ACCT4(c_issue += clock64() - w1;)
            };
            auto advance = [&](uint64_t bd, int j0, int n) {  // descriptor of k range j0 + n given the one of j0
              for (int i = 0; i < n; ++i) bd += ((j0 + i) & 1) ? (uint64_t)((S_KBLK >> 4) - 4) : 4ull;
              return bd;
            };
            uint64_t bd = bd_a;
            if (layer == 0) {
              issue(bd, 1, 0);  // x_t: one k range
            } else {
#pragma unroll 1
              for (int j = 0; j < n_a; j += 2) { issue(bd, 2, j); bd = advance(bd, j, 2); }
            }
            bd = bd_b;
#pragma unroll 1
            for (int j = 0; j < n_b; j += 2) { issue(bd, 2, j); bd = advance(bd, j, 2); }
            if (elect_one()) tc_commit2_mc(&bars.accf[slot], pair_mask);
            __syncwarp();
          }
          // l1_done lets the epilogues overwrite h1_{t-1} with h1_t: the Linear warps (ranks 0, 1) must have consumed it
          if (layer == 1 && t >= 1) mbar_wait<false>(&bars.fc_done, (t - 1) & 1);
          if (elect_one()) {
            if (layer == 0) {
              tc_commit2_mc(&bars.x_empty[t & 1], pair_mask);
            } else {
              tc_commit2_mc(&bars.l1_done, (uint16_t)((1u << NCTA) - 1));  // every layer-1 MMA of step t (this pair) has completed
            }
          }
          __syncwarp();
        }
      }
      ACCT4(if (a.dbg && blockIdx.x == 0 && lane == 0) {
        a.dbg[0] = clock64() - c_start; a.dbg[1] = c_state; a.dbg[2] = c_acce; a.dbg[3] = c_full; a.dbg[4] = c_issue; a.dbg[5] = c_x; a.dbg[6] = c_h0;
      })
    } else if (warp == 2) {
      // ================= x_t gather for this CTA's 64 rows (base_model.py:35-44, model.py:98-111)
      const int nmag = 2 * a.Ns + 1;
      ACCT4(long long g_wait = 0, g_body = 0, g_sig = 0;)
      for (int t = 0; t < Tp; ++t) {
        ACCT4(const long long g0 = clock64();)
        mbar_wait<true>(&bars.x_empty[t & 1], ((t >> 1) & 1) ^ 1);
        ACCT4(const long long g1 = clock64(); g_wait += g1 - g0;)
        uint8_t* xb = smem + SM_X + (t & 1) * X_BLK;
        // two explicit phases per batch of 16 rows so that 16 global loads are in flight (the step is ~10 us: a
        // dependent load per row would cost 64 x ~0.4 us)
#pragma unroll 1
        for (int n0 = 0; n0 < NBR; n0 += 16) {
          float raw[16], sc[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {  // branch-free: every lane loads (a safe address when the row / column is padding)
            const RowInfo ri = rows[n0 + i];
            const bool ok = ri.src_b >= 0 && lane < a.Ksb;
            const size_t base = ((size_t)(ok ? ri.src_b : 0) * Tp + t) * a.F;
            const int col = (lane < nmag) ? reflect_idx(ri.src_f + lane - a.Ns, a.F) : reflect_idx(ri.src_f + (lane - nmag) - a.Nf, a.F);
            const float* src = ((lane < nmag) ? a.magT : a.fbT) + base + (ok ? col : 0);
            raw[i] = __ldg(src);
            sc[i] = ok ? ri.scale : 0.f;
          }
          if (a.unit_scale) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (sc[i] != 0.f) sc[i] = a.unit_scale[(size_t)t * a.R + row0 + n0 + i];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) *reinterpret_cast<__half*>(xb + swz64_off(n0 + i, lane)) = __float2half_rn(raw[i] * sc[i]);
        }
        ACCT4(const long long g2 = clock64(); g_body += g2 - g1;)
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(&bars.x_full[t & 1], lead_rank);
        ACCT4(g_sig += clock64() - g2;)
      }
      ACCT4(if (a.dbg && blockIdx.x == 0 && lane == 0) { a.dbg[20] = g_wait; a.dbg[21] = g_body; a.dbg[22] = g_sig; })
    } else {
      // ================= state warp (every CTA): posts the expect_tx of the h barriers, relays "my half has landed"
      // to the pair leader (peers), and on ranks 0, 1 evaluates Linear(H -> 2) + output for rows [64 rank, +64) from
      // the complete h1_t (fp16) in this CTA's shared memory
      constexpr uint32_t H_BYTES = NKH * S_KBLK;  // 48 KB per layer-step and CTA
      const float fcb0 = reinterpret_cast<const float*>(a.packed + OFF_FCB)[0];
      const float fcb1 = reinterpret_cast<const float*>(a.packed + OFF_FCB)[1];
      const uint8_t* h1 = smem + SM_H1;
      int staged = 0, t_stage0 = 0;
      if (lane == 0) { mbar_expect_tx(&bars.h0_tx, H_BYTES); mbar_expect_tx(&bars.h1_tx, H_BYTES); }
      __syncwarp();
      for (int it = 0; it <= Tp; ++it) {
        if (it < Tp) {  // h0_it
          mbar_wait<false>(&bars.h0_tx, it & 1);
          if (lane == 0) {
            if (it + 1 < Tp) mbar_expect_tx(&bars.h0_tx, H_BYTES);
            if (!leader) mbar_arrive_cluster(&bars.h0_peer, lead_rank);
          }
          __syncwarp();
        }
        if (it >= 1) {  // h1_{it-1}
          const int t = it - 1;
          mbar_wait<false>(&bars.h1_tx, t & 1);
          if (lane == 0) {
            if (t + 1 < Tp) mbar_expect_tx(&bars.h1_tx, H_BYTES);
            if (!leader) mbar_arrive_cluster(&bars.h1_peer, lead_rank);
          }
          __syncwarp();
          if (rank < 2) {
            if (t >= a.la) {
              float s[2][2] = {{fcb0, fcb1}, {fcb0, fcb1}};
#pragma unroll 2
              for (int ch = 0; ch < H / 8; ++ch) {
                const int kb = ch >> 3, kk = (ch & 7) * 8;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                  const int row = lane + 32 * rr;
                  const uint4 raw = *reinterpret_cast<const uint4*>(h1 + kb * S_KBLK + swz128_off(row, kk));
                  const __half2* hp = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
                  for (int e = 0; e < 4; ++e) {
                    const float2 f = __half22float2(hp[e]);
                    const int u = ch * 8 + 2 * e;
                    s[rr][0] = fmaf(f.x, fcw_s[u], s[rr][0]);     s[rr][0] = fmaf(f.y, fcw_s[u + 1], s[rr][0]);
                    s[rr][1] = fmaf(f.x, fcw_s[H + u], s[rr][1]); s[rr][1] = fmaf(f.y, fcw_s[H + u + 1], s[rr][1]);
                  }
                }
              }
              if (staged == 0) t_stage0 = t - a.la;
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                outst[((lane + 32 * rr) * 2 + 0) * OUT_T + staged] = act_apply(s[rr][0], a.act);
                outst[((lane + 32 * rr) * 2 + 1) * OUT_T + staged] = act_apply(s[rr][1], a.act);
              }
              ++staged;
            }
            __syncwarp();
            if (lane == 0)
              for (int pr = 0; pr < NPAIR; ++pr) mbar_arrive_cluster(&bars.fc_done, (uint32_t)(2 * pr));
            if (staged == OUT_T || (t == Tp - 1 && staged > 0)) {
#pragma unroll
              for (int rr = 0; rr < 2; ++rr) {
                const RowInfo ri = rows[lane + 32 * rr];
                if (ri.src_b >= 0) {
#pragma unroll
                  for (int o = 0; o < 2; ++o) {
                    float* dst = a.crm + ((size_t)ri.out_idx + (size_t)o * a.Fsub) * a.T + t_stage0;
                    for (int i = 0; i < staged; ++i) dst[i] = outst[((lane + 32 * rr) * 2 + o) * OUT_T + i];
                  }
                }
              }
              staged = 0;
            }
          }
        }
      }
    }
  } else {
    // ================= epilogue: warpgroup = tile, warp & 3 = TMEM lane quadrant, thread = (unit, column pair)
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int tile = (warp - 4) >> 2;
    const int q = warp & 3;
    const int u8 = lane >> 2, cp = lane & 3;
    const int u = (int)rank * UC + tile * TU + q * 8 + u8;  // hidden unit of this thread
    const float* bias_g = reinterpret_cast<const float*>(a.packed + OFF_BIAS);
    float b0[4], b1[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) { b0[g] = bias_g[g * H + u]; b1[g] = bias_g[4 * H + g * H + u]; }
    float c0[32], c1[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) c0[i] = c1[i] = 0.f;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const int kbu = u >> 6, ku = u & 63;
    uint32_t dst_base[NCTA];
#pragma unroll
    for (int i = 0; i < NCTA; ++i) dst_base[i] = mapa(smem_u32(smem), (uint32_t)i);
    const int chunk = ku >> 3;  // 16-byte chunk of the k block that holds units 8q..8q+7 of this tile
    uint32_t seq = (uint32_t)tile;  // tile sequence number of this warpgroup's next tile (advances by 3 per layer-step)
    ACCT4(long long e_accf = 0, e_comp = 0, e_free = 0, e_store = 0;)
    for (int it = 0; it <= Tp; ++it) {
      for (int layer = 0; layer < 2; ++layer) {
        const int t = it - layer;
        if (t < 0 || t >= Tp) continue;
        const uint32_t slot = seq & 3;
        ACCT4(long long e0 = clock64();)
        mbar_wait<false>(&bars.accf[slot], (seq >> 2) & 1);
        ACCT4(long long e1 = clock64(); e_accf += e1 - e0;)
        seq += TILES;
        tc_fence_after();
        const uint32_t tacc = tmem_base + lane_addr + slot * 128;
        // 16-byte destination of (cell row, the 8 units of this quadrant): units 8q..8q+7 of the tile share one chunk
        const uint32_t hoff = (layer ? SM_H1 : SM_H0 + (uint32_t)((t & 1) * NKH * S_KBLK)) + (uint32_t)(kbu * S_KBLK);
        uint32_t bar_tx[NCTA];  // the destination CTAs' h barrier of this layer
#pragma unroll
        for (int i = 0; i < NCTA; ++i) bar_tx[i] = dst_base[i] + SM_BARS + (uint32_t)(layer ? offsetof(Bars, h1_tx) : offsetof(Bars, h0_tx));
        uint4 hst[4];
#pragma unroll
        for (int ck = 0; ck < 4; ++ck) {
          float gif[16], ggo[16];
          tc_ld16x256_x4(tacc + ck * 32, gif);                      // lanes 0-15 of the quadrant: gates i, f
          tc_ld16x256_x4(tacc + (16u << 16) + ck * 32, ggo);        // lanes 16-31: gates g, o
          tc_wait_ld();
          float hh[8];
#pragma unroll
          for (int b = 0; b < 4; ++b) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int ci = ck * 8 + b * 2 + e;
              const float gi = gif[4 * b + e], gf = gif[4 * b + 2 + e], gg = ggo[4 * b + e], go = ggo[4 * b + 2 + e];
              if (layer == 0) hh[b * 2 + e] = lstm_cell7(gi + b0[0], gf + b0[1], gg + b0[2], go + b0[3], c0[ci]);
              else            hh[b * 2 + e] = lstm_cell7(gi + b1[0], gf + b1[1], gg + b1[2], go + b1[3], c1[ci]);
            }
          }
          // gather the 8 units of this quadrant for ONE row per thread: hh[i] = h(unit 8q + i, cell u8)
          transpose8_units(hh, u8);
          const uint4 pk = pack8_half(hh);
          hst[ck] = pk;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_relaxed(&bars.acce[slot], lead_rank);
        ACCT4(e0 = clock64(); e_comp += e0 - e1;)
        {
          // h1 is single-buffered: every layer-1 MMA of step t (both pairs) must have read h1_{t-1} before h1_t lands.
          // (h0_t goes into buffer t&1, last read by the layer-1 MMAs of step t-2, whose l1_done this warp has waited for
          //  in its own layer-1 epilogue of that step.)  Rows = columns of chunk ck live in CTA r = ck >> 1 of both pairs;
          // my cell: b = u8 >> 1, e = u8 & 1.
          if (layer == 1) mbar_wait<false>(&bars.l1_done, t & 1);
          ACCT4(e1 = clock64(); e_free += e1 - e0;)
#pragma unroll
          for (int ck = 0; ck < 4; ++ck) {
            const int nrow = ((ck & 1) * 32) + 8 * (u8 >> 1) + 2 * cp + (u8 & 1);
            const uint32_t o = hoff + (uint32_t)((nrow >> 3) * 1024 + (nrow & 7) * 128 + ((chunk ^ (nrow & 7)) << 4));
#pragma unroll
            for (int pr = 0; pr < NPAIR; ++pr)  // every pair holds a copy of the state: CTA r = ck >> 1 of each pair
              st_async_v4(dst_base[2 * pr + (ck >> 1)] + o, hst[ck], bar_tx[2 * pr + (ck >> 1)]);
          }
        }
        ACCT4(e_store += clock64() - e1;)
      }
    }
    ACCT4(if (a.dbg && blockIdx.x == 0 && q == 0 && lane == 0) {
      a.dbg[8 + tile * 4 + 0] = e_accf; a.dbg[8 + tile * 4 + 1] = e_comp; a.dbg[8 + tile * 4 + 2] = e_free; a.dbg[8 + tile * 4 + 3] = e_store;
    })
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

}  // namespace tc4

bool sb_tc4_supported(const fsn_model_desc* d) {
  static const int on = getenv("FSN_TC_CLUSTER4") ? atoi(getenv("FSN_TC_CLUSTER4")) : 0;  // opt-in until validated
  const int Ksb = (2 * d->sb_num_neighbors + 1) + (2 * d->fb_num_neighbors + 1);
  return on != 0 && d->cell_type == FSN_CELL_LSTM && d->precision == FSN_PREC_F16_TC && d->sb_hidden == tc4::H && Ksb <= tc4::KS;
}

size_t sb_tc4_packed_bytes() { return tc4::PACKED_BYTES; }

int sb_tc4_pack(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, cudaStream_t st) {
  const int Ksb = (2 * d->sb_num_neighbors + 1) + (2 * d->fb_num_neighbors + 1);
  tc4::pack4_kernel<<<148 * 4, 256, 0, st>>>(sb->w_ih[0], sb->w_hh[0], sb->w_ih[1], sb->w_hh[1], sb->b_ih[0], sb->b_hh[0],
                                             sb->b_ih[1], sb->b_hh[1], sb->fc_w, sb->fc_b, Ksb, 2, (uint8_t*)packed);
  FSN_CHECK_LAUNCH("sb pack4_kernel");
  return FSN_OK;
}

int sb_tc4_forward(const SbTcArgs& s, cudaStream_t st) {
  tc4::KArgs a;
  a.packed = (const uint8_t*)s.packed;
  a.magT = s.magT; a.fbT = s.fbT; a.inv2 = s.inv2; a.unit_scale = s.unit_scale; a.crm = s.crm;
  a.R = s.map.B * s.map.Fsub; a.F = s.F; a.Tp = s.Tp; a.la = s.la; a.T = s.Tp - s.la;
  a.Ns = s.Ns; a.Nf = s.Nf; a.Ksb = (2 * s.Ns + 1) + (2 * s.Nf + 1); a.act = s.act;
  a.Fsub = s.map.Fsub; a.map = s.map;
  if (getenv("FSN_TC4_NOGATHER")) a.Ksb = 0;  // timing experiment only: x_t = 0 without touching global memory
  a.dbg = nullptr;
  static long long* dbg_buf = nullptr;
  if (getenv("FSN_TC_TRACE")) {
    if (!dbg_buf) { cudaMalloc(&dbg_buf, 32 * sizeof(long long)); cudaMemset(dbg_buf, 0, 32 * sizeof(long long)); }
    a.dbg = dbg_buf;
  }
  const size_t smem = tc4::SM_TOTAL + 1024;
  int rc = check_cuda(cudaFuncSetAttribute(tc4::sb_lstm_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                      "sb_lstm_tc4 smem attr");
  if (rc) return rc;
  const int clusters = cdiv(a.R, tc4::NCL);
  if (getenv("FSN_TC_TRACE")) {  // how many clusters fit on the chip at once
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tc4::NCTA * clusters); cfg.blockDim = dim3(tc4::NTHREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = tc4::NCTA; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, tc4::sb_lstm_tc4_kernel, &cfg);
    fprintf(stderr, "[tc4] cluster of %d CTAs: max active clusters %d (%s), %d clusters to run\n", tc4::NCTA, nc, cudaGetErrorString(e),
            clusters);
  }
  tc4::sb_lstm_tc4_kernel<<<tc4::NCTA * clusters, tc4::NTHREADS, smem, st>>>(a);
  FSN_CHECK_LAUNCH("sb_lstm_tc4_kernel");
  if (a.dbg) {
    long long h[32];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, a.dbg, sizeof(h), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[tc4 leader cycles] total %lld | wait state (x, h0, h1, fc) %lld | wait accumulator slot %lld | wait weight stage "
                    "%lld | issue + commit %lld | of the state wait: x %lld, h0 %lld  (Tp=%d, %d granules/step)\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], a.Tp,
            tc4::TILES * (tc4::NK0 + tc4::NK1));
    fprintf(stderr, "[tc4 gather warp, per step] wait x_empty %lld | loads + stores %lld | fence + arrive %lld\n", h[20] / a.Tp,
            h[21] / a.Tp, h[22] / a.Tp);
    for (int j = 0; j < 3; ++j)
      fprintf(stderr, "[tc4 epilogue tile %d, per layer-step] wait accumulator %lld | ld + cell math + transpose %lld | wait h free %lld | "
                      "store + fence + arrive %lld\n", j, h[8 + j * 4] / (2 * a.Tp), h[9 + j * 4] / (2 * a.Tp), h[10 + j * 4] / (2 * a.Tp),
              h[11 + j * 4] / (2 * a.Tp));
  }
  return FSN_OK;
}

}  // namespace fsn
