// Training step of recipes/dns_interspeech_2020/fullsubnet/trainer.py:56-68, fp32:
//   fsn_train_forward   Model.forward in train mode, keeping what back-propagation through time needs
//   fsn_mse_loss        audio_zen/loss.py:4 (MSELoss) + d loss / d cRM
//   fsn_train_backward  BPTT through sub-band stack -> second norm (closed form) -> drop_band row map ->
//                       full-band Linear/ReLU -> full-band stack; weight gradients as split-K GEMMs over all steps
//   fsn_clip_adam       clip_grad_norm_ + Adam in three launches, no host synchronisation
// All saved activations are time-major ([Tp, rows, ...]) so every per-step operand is one contiguous block and
// every weight gradient is one GEMM over K = Tp*rows.  oracle/train_oracle.py:manual_backward is the same
// algorithm on the CPU.
#include <string.h>

#include "fsn_internal.cuh"

namespace fsn {

// ------------------------------------------------------------------------------------------ GEMM
// C[M,N] (+)= op(A) B,  B [K,N] row-major (ldb); op(A) = A [M,K] (lda) or, TA, A stored [K,M] (lda).
// 64x64 tile, 4x4 per thread.  blockIdx.z = split-K slice writing its own [M,N] slab (ldc = N) at C + z*M*N.
template <bool TA>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, size_t lda, const float* __restrict__ Bm, size_t ldb, float* __restrict__ C,
             size_t ldc, int M, int N, int K, int k_per_split, int accumulate, size_t split_stride) {
  __shared__ __align__(16) float As[16][64];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int kb = blockIdx.z * k_per_split;
  const int ke = (kb + k_per_split < K) ? kb + k_per_split : K;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = kb; k0 < ke; k0 += 16) {
    if (!TA) {
      const int row = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        As[kq + j][row] = (m0 + row < M && k < ke) ? A[(size_t)(m0 + row) * lda + k] : 0.f;
      }
    } else {
      const int kk = tid >> 4, mq = (tid & 15) * 4, k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + mq + j;
        As[kk][mq + j] = (k < ke && m < M) ? A[(size_t)k * lda + m] : 0.f;
      }
    }
    {
      const int kk = tid >> 4, nq = (tid & 15) * 4, k = k0 + kk;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + nq + j;
        Bs[kk][nq + j] = (k < ke && n < N) ? Bm[(size_t)k * ldb + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float b = Bs[kk][tx + 16 * j];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][j] = fmaf(a[i], b, acc[i][j]);
      }
    }
    __syncthreads();
  }
  float* Cz = C + (size_t)blockIdx.z * split_stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx + 16 * j;
      if (col >= N) continue;
      float* dst = Cz + (size_t)row * ldc + col;
      *dst = accumulate ? *dst + acc[i][j] : acc[i][j];
    }
  }
}

// C[m,n] (+)= sum_s part[s][m][n]  (fixed order: deterministic)
__global__ void splitk_reduce_kernel(const float* __restrict__ part, int S, int M, int N, float* __restrict__ C,
                                     size_t ldc, int accumulate) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)M * N) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += part[(size_t)z * M * N + i];
  float* dst = C + (i / N) * ldc + (i % N);
  *dst = accumulate ? *dst + s : s;
}

int splitk_reduce_launch(const float* part, int S, int M, int N, float* C, size_t ldc, bool accumulate, cudaStream_t st) {
  const size_t n = (size_t)M * N;
  splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(part, S, M, N, C, ldc, accumulate ? 1 : 0);
  FSN_CHECK_LAUNCH("splitk_reduce_kernel");
  return FSN_OK;
}

static const size_t SPLITK_SCRATCH_FLOATS = (size_t)16 << 20;  // 64 MB

static int sgemm_launch(bool ta, const float* A, size_t lda, const float* Bm, size_t ldb, float* C, size_t ldc, int M,
                        int N, int K, bool accumulate, float* scratch, cudaStream_t st) {
  if (M <= 0 || N <= 0 || K <= 0) return FSN_OK;
  const int tiles = cdiv(M, 64) * cdiv(N, 64);
  int S = 1;
  if (scratch && K >= 4096 && tiles < 592) {  // fill 148 SMs x 4 CTAs; slices of >= 1024 k
    S = cdiv(592, tiles);
    if (S > cdiv(K, 1024)) S = cdiv(K, 1024);
    while (S > 1 && (size_t)S * M * N > SPLITK_SCRATCH_FLOATS) --S;
  }
  const int kps = cdiv(cdiv(K, S), 16) * 16;
  S = cdiv(K, kps);
  dim3 grid(cdiv(M, 64), cdiv(N, 64), S);
  float* dst = S > 1 ? scratch : C;
  const size_t ldd = S > 1 ? (size_t)N : ldc;
  const int acc = (S > 1) ? 0 : (accumulate ? 1 : 0);
  if (ta) sgemm_kernel<true><<<grid, 256, 0, st>>>(A, lda, Bm, ldb, dst, ldd, M, N, K, kps, acc, (size_t)M * N);
  else    sgemm_kernel<false><<<grid, 256, 0, st>>>(A, lda, Bm, ldb, dst, ldd, M, N, K, kps, acc, (size_t)M * N);
  FSN_CHECK_LAUNCH("sgemm_kernel");
  if (S > 1) {
    const size_t n = (size_t)M * N;
    splitk_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(scratch, S, M, N, C, ldc, accumulate ? 1 : 0);
    FSN_CHECK_LAUNCH("splitk_reduce_kernel");
  }
  return FSN_OK;
}

// out[c] = sum_r X[r*ldx + c]: slabs of rows -> part[S][cols] -> fixed-order sum
__global__ void colsum_part_kernel(const float* __restrict__ X, size_t rows, int cols, size_t ldx, size_t rows_per,
                                   float* __restrict__ part) {
  __shared__ float sh[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const size_t r0 = (size_t)blockIdx.y * rows_per;
  const size_t r1 = (r0 + rows_per < rows) ? r0 + rows_per : rows;
  float s = 0.f;
  if (c < cols)
    for (size_t r = r0 + threadIdx.y; r < r1; r += 8) s += X[r * ldx + c];
  sh[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][threadIdx.x];
    part[(size_t)blockIdx.y * cols + c] = t;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int S, int cols, float* __restrict__ out,
                                    float* __restrict__ out2) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int z = 0; z < S; ++z) s += part[(size_t)z * cols + c];
  out[c] = s;
  if (out2) out2[c] = s;
}
static const int COLSUM_MAX_S = 512;

// dW [O,H] = dout^T Hm for a Linear with a few outputs (the sub-band Linear, O = 2; model.py:129-135 backwards):
// dout [rows,O], Hm [rows,H].  One pass over Hm: a CTA owns a slab of rows, a thread one column; part [S][O][H], then
// the fixed-order sum of colsum_final_kernel over S slabs of O*H "columns".
template <int O>
__global__ void __launch_bounds__(128) small_out_wgrad_kernel(const float* __restrict__ dout, const float* __restrict__ Hm,
                                                              size_t rows, int H, size_t rows_per, float* __restrict__ part) {
  const int c = blockIdx.x * 128 + threadIdx.x;
  const size_t r0 = (size_t)blockIdx.y * rows_per;
  const size_t r1 = (r0 + rows_per < rows) ? r0 + rows_per : rows;
  float acc[O];
#pragma unroll
  for (int o = 0; o < O; ++o) acc[o] = 0.f;
  if (c < H) {
    size_t r = r0;
    for (; r + 4 <= r1; r += 4) {
      float h[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __ldcs(Hm + (r + j) * H + c);
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int o = 0; o < O; ++o) acc[o] = fmaf(dout[(r + j) * O + o], h[j], acc[o]);
    }
    for (; r < r1; ++r) {
      const float h = Hm[r * H + c];
#pragma unroll
      for (int o = 0; o < O; ++o) acc[o] = fmaf(dout[r * O + o], h, acc[o]);
    }
#pragma unroll
    for (int o = 0; o < O; ++o) part[((size_t)blockIdx.y * O + o) * H + c] = acc[o];
  }
}
static int colsum_launch(const float* X, size_t rows, int cols, size_t ldx, float* out, float* out2, float* scratch,
                         cudaStream_t st) {
  int S = (int)((rows + 2047) / 2048);
  if (S > COLSUM_MAX_S) S = COLSUM_MAX_S;
  if (S < 1) S = 1;
  const size_t rows_per = (rows + S - 1) / S;
  colsum_part_kernel<<<dim3(cdiv(cols, 32), S), dim3(32, 8), 0, st>>>(X, rows, cols, ldx, rows_per, scratch);
  FSN_CHECK_LAUNCH("colsum_part_kernel");
  colsum_final_kernel<<<cdiv(cols, 128), 128, 0, st>>>(scratch, S, cols, out, out2);
  FSN_CHECK_LAUNCH("colsum_final_kernel");
  return FSN_OK;
}

// ------------------------------------------------------------------------------------------ forward helpers
// per-clip sums of noisy_mag [B,F,T]: (sum, sum_f c_Ns[f] * row sum), one CTA per clip, fixed-order tree
__global__ void train_mag_stats_kernel(const float* __restrict__ mag, int F, int T, int Ns, float2* __restrict__ sums) {
  __shared__ float2 sh[256];
  const int b = blockIdx.x;
  const float* p = mag + (size_t)b * F * T;
  float2 a = make_float2(0.f, 0.f);
  for (int i = threadIdx.x; i < F * T; i += 256) {
    const float v = p[i];
    a.x += v;
    a.y += v * (float)reflect_count(i / T, F, Ns);
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sh[threadIdx.x].x += sh[threadIdx.x + s].x; sh[threadIdx.x].y += sh[threadIdx.x + s].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[b] = sh[0];
}

// same for a time-major tensor x [Tp,B,F]
__global__ void train_tm_stats_kernel(const float* __restrict__ x, int B, int F, int Tp, int N, float2* __restrict__ sums) {
  __shared__ float2 sh[256];
  const int b = blockIdx.x;
  float2 a = make_float2(0.f, 0.f);
  for (int i = threadIdx.x; i < F * Tp; i += 256) {
    const int t = i / F, f = i - t * F;
    const float v = x[((size_t)t * B + b) * F + f];
    a.x += v;
    a.y += v * (float)reflect_count(f, F, N);
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) { sh[threadIdx.x].x += sh[threadIdx.x + s].x; sh[threadIdx.x].y += sh[threadIdx.x + s].y; }
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[b] = sh[0];
}

// mag [B,F,T] -> raw [Tp,B,F] (zero look-ahead frames, model.py:85) and scaled = raw * inv1[b] (model.py:92)
__global__ void train_transpose_kernel(const float* __restrict__ mag, const float* __restrict__ inv1,
                                       float* __restrict__ raw, float* __restrict__ scaled, int B, int F, int T, int Tp) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    tile[i][tx] = (f < F && t < T) ? mag[((size_t)b * F + f) * T + t] : 0.f;
  }
  __syncthreads();
  const float s = inv1[b];
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (t < Tp && f < F) {
      const float v = tile[tx][i];
      const size_t o = ((size_t)t * B + b) * F + f;
      raw[o] = v;
      scaled[o] = v * s;
    }
  }
}

// sub-band input X[t,r,k] (base_model.py:13-46 + model.py:98-119): unit (b,f) of row r, scaled by inv2[b]
__global__ void train_gather_kernel(const float* __restrict__ raw, const float* __restrict__ fbz,
                                    const float* __restrict__ inv2, const float* __restrict__ unit_scale,
                                    float* __restrict__ X, RowMap map, int Tp, int R, int Ns, int Nf) {
  const int K = 2 * Ns + 1 + 2 * Nf + 1;
  const size_t n = (size_t)Tp * R * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const size_t tr = i / K;
    const int r = (int)(tr % R), t = (int)(tr / R);
    int b, f;
    row_to_unit(map, r, b, f);
    const size_t base = ((size_t)t * map.B + b) * map.F;
    float v;
    if (k < 2 * Ns + 1) v = raw[base + reflect_idx(f + k - Ns, map.F)];
    else                v = fbz[base + reflect_idx(f + (k - 2 * Ns - 1) - Nf, map.F)];
    X[i] = v * (unit_scale ? unit_scale[tr] : inv2[b]);  // cumulative norm: scale of (step t, unit r) at [t*R + r]
  }
}

// ---- cumulative_laplace_norm in the training step (audio_zen/model/base_model.py:220-251)
// frame sums of a time-major tensor raw [Tp,B,F] in the layout cum_clip_scale_launch reads: fs[b*Tp + t].x
__global__ void train_frame_sum_kernel(const float* __restrict__ raw, int B, int F, int Tp, float2* __restrict__ fs) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // row = t*B + b
  const int lane = threadIdx.x & 31;
  if (row >= Tp * B) return;
  const float* p = raw + (size_t)row * F;
  float a = 0.f;
  for (int f = lane; f < F; f += 32) a += p[f];
  a = warp_sum(a);
  if (lane == 0) { const int t = row / B, b = row - t * B; fs[(size_t)b * Tp + t] = make_float2(a, a); }
}
// xfb[t,b,f] = raw[t,b,f] * scale1T[t*B + b]
__global__ void train_scale_tm_kernel(const float* __restrict__ raw, const float* __restrict__ scale1T, int F, size_t n,
                                      float* __restrict__ out) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = raw[i] * scale1T[i / F];
}
// Backward of X[t,r,k] = u[t,r,k] * s[t,r], s = 1/(m + eps), m[t,r] = sum_{t'<=t} sum_k u[t',r,k] / (K (t+1)):
//   d u[t,r,k] = dX[t,r,k] s[t,r] + sum_{t''>=t} q[t'',r],   q[t,r] = -s[t,r] <dX[t,r,:], X[t,r,:]> / (K (t+1)).
// Only the full-band row k = K-1 has a parameter behind it (Nf = 0): dunit[t,r] = its gradient.  One thread per
// unit, sequential in t (suffix sum in a fixed order).
__global__ void train_cum_unit_bwd_kernel(const float* __restrict__ dX, const float* __restrict__ X,
                                          const float* __restrict__ scaleT, int Tp, int R, int K, float* __restrict__ dunit) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  float suffix = 0.f;
  for (int t = Tp - 1; t >= 0; --t) {
    const size_t o = ((size_t)t * R + r) * K;
    float dot = 0.f;
    for (int k = 0; k < K; ++k) dot = fmaf(dX[o + k], X[o + k], dot);
    const float s = scaleT[(size_t)t * R + r];
    suffix += -s * dot / ((float)K * (float)(t + 1));
    dunit[(size_t)t * R + r] = fmaf(dX[o + K - 1], s, suffix);
  }
}
// dz[t,b,f] = act'(fbz) * dunit[t, row(b,f)]  (units removed by drop_band carry no gradient: their norm is their own)
__global__ void train_dfbz_cum_kernel(const float* __restrict__ dunit, const float* __restrict__ fbz, RowMap map, int Tp,
                                      int R, int act, float* __restrict__ dz) {
  const size_t n = (size_t)Tp * map.B * map.F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % map.F);
    const size_t tb = i / map.F;
    const int b = (int)(tb % map.B), t = (int)(tb / map.B);
    const int r = unit_to_row(map, b, f);
    float v = r >= 0 ? dunit[(size_t)t * R + r] : 0.f;
    const float y = fbz[i];
    if (act == FSN_ACT_RELU) v = y > 0.f ? v : 0.f;
    else if (act == FSN_ACT_TANH) v *= 1.f - y * y;
    else if (act == FSN_ACT_RELU6) v = (y > 0.f && y < 6.f) ? v : 0.f;
    dz[i] = v;
  }
}

// LSTM cell of one step (tensor-core path): G_t [R,4H] holds x_t W_ih^T (all steps from one hoisted GEMM), rec the
// recurrent product of this step; G_t is overwritten with the post-activation gates (i,f,g,o); writes c_t and h_t
__global__ void lstm_cell_fwd_kernel(float* __restrict__ G, const float* __restrict__ rec, const float* __restrict__ b_ih,
                                     const float* __restrict__ b_hh, const float* __restrict__ C_prev,
                                     float* __restrict__ C_out, float* __restrict__ H_out, int R, int H) {
  const size_t n = (size_t)R * H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % H);
    const size_t r = idx / H;
    float* g = G + r * 4 * H + u;
    float z[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      z[q] = g[q * H] + b_ih[q * H + u] + b_hh[q * H + u];
      if (rec) z[q] += rec[r * 4 * H + q * H + u];  // h_{t-1} W_hh^T of this step
    }
    const float si = sigmoidf_(z[0]), sf = sigmoidf_(z[1]), tg = tanhf(z[2]), so = sigmoidf_(z[3]);
    const float c = sf * (C_prev ? C_prev[idx] : 0.f) + si * tg;
    g[0] = si; g[H] = sf; g[2 * H] = tg; g[3 * H] = so;
    C_out[idx] = c;
    H_out[idx] = so * tanhf(c);
  }
}

// out[c, r] = in[r, c]   (in [rows, cols] row-major -> out [cols, rows]); operands of the K-major tensor-core GEMM
__global__ void transpose_kernel(const float* __restrict__ in, size_t rows, int cols, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const size_t r0 = (size_t)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const size_t r = r0 + i;
    const int c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i;
    const size_t r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}
static int transpose_launch(const float* in, size_t rows, int cols, float* out, cudaStream_t st) {
  dim3 grid((unsigned)((rows + 31) / 32), cdiv(cols, 32));
  transpose_kernel<<<grid, dim3(32, 8), 0, st>>>(in, rows, cols, out);
  FSN_CHECK_LAUNCH("transpose_kernel");
  return FSN_OK;
}

// ------------------------------------------------------------------------------------------ backward kernels
// dout[t,r,o] = dcrm[b',o,f',t-la] (0 for the look-ahead steps)  (model.py:129-135 backwards)
__global__ void train_dout_kernel(const float* __restrict__ dcrm, float* __restrict__ dout, int R, int Fsub, int T,
                                  int Tp, int la) {
  const size_t n = (size_t)Tp * R * 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(i & 1);
    const size_t tr = i >> 1;
    const int r = (int)(tr % R), t = (int)(tr / R);
    float v = 0.f;
    if (t >= la) {
      const int bq = r / Fsub, fq = r - bq * Fsub;
      v = dcrm[(((size_t)bq * 2 + o) * Fsub + fq) * T + (t - la)];
    }
    dout[i] = v;
  }
}

struct BwdPoint {
  int R, H;
  float* G;             // [R,4H] in: gates (post-activation), out: d(pre-activation)
  const float* C;       // [R,H] cell state of this step
  const float* C_prev;  // nullable (t == 0)
  const float* dh_above;  // nullable [R,H]
  const float* dh_rec;    // nullable [R,H]
  float* dc;              // [R,H] in (ignored when first_dc): d c_t from step t+1, out: d c_{t-1}
  int first_dc;
  const float* dout;  // nullable [R,O]: dh_above += dout W_fc   (Linear backward, small O)
  const float* fc_w;  // [O,H]
  int O;
};

__global__ void lstm_bwd_point_kernel(const BwdPoint p) {
  const size_t n = (size_t)p.R * p.H;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int u = (int)(idx % p.H);
    const size_t r = idx / p.H;
    float dh = 0.f;
    if (p.dh_above) dh += p.dh_above[idx];
    if (p.dh_rec) dh += p.dh_rec[idx];
    if (p.dout)
      for (int o = 0; o < p.O; ++o) dh = fmaf(p.dout[r * p.O + o], p.fc_w[(size_t)o * p.H + u], dh);
    float* g = p.G + r * 4 * p.H + u;
    const float gi = g[0], gf = g[p.H], gg = g[2 * p.H], go = g[3 * p.H];
    const float tc = tanhf(p.C[idx]);
    const float dc_tot = (p.first_dc ? 0.f : p.dc[idx]) + dh * go * (1.f - tc * tc);
    const float c_prev = p.C_prev ? p.C_prev[idx] : 0.f;
    g[0] = dc_tot * gg * gi * (1.f - gi);
    g[p.H] = dc_tot * c_prev * gf * (1.f - gf);
    g[2 * p.H] = dc_tot * gi * (1.f - gg * gg);
    g[3 * p.H] = dh * tc * go * (1.f - go);
    p.dc[idx] = dc_tot * gf;
  }
}

// dot[b'] = sum over the rows of output clip b' and all t,k of dX * X   (second-norm backward)
__global__ void train_dot_kernel(const float* __restrict__ dX, const float* __restrict__ X, int Tp, int R, int Fsub,
                                 int K, float* __restrict__ dot) {
  __shared__ float sh[256];
  const int bq = blockIdx.x;
  const size_t per_t = (size_t)Fsub * K;
  float a = 0.f;
  for (int t = 0; t < Tp; ++t) {
    const size_t base = ((size_t)t * R + (size_t)bq * Fsub) * K;
    for (size_t i = threadIdx.x; i < per_t; i += 256) a = fmaf(dX[base + i], X[base + i], a);
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) dot[bq] = sh[0];
}

// dz[t,b,f] = act'(fbz) * ( dX[t, row(b,f), K-1] * inv2[b]  -  inv2[b] * dot[b'(b)] * c_Nf[f] / cnt2 )   (Nf = 0)
__global__ void train_dfbz_kernel(const float* __restrict__ dX, const float* __restrict__ fbz,
                                  const float* __restrict__ inv2, const float* __restrict__ dot, RowMap map, int Tp,
                                  int R, int K, float cnt2, int act, float* __restrict__ dz) {
  const size_t n = (size_t)Tp * map.B * map.F;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i % map.F);
    const size_t tb = i / map.F;
    const int b = (int)(tb % map.B), t = (int)(tb / map.B);
    // output clip of b: row of any kept frequency of that clip / Fsub
    int bq;
    if (map.G <= 1) bq = b;
    else bq = unit_to_row(map, b, b % map.G) / map.Fsub;
    const float s = inv2[b];
    float v = -s * dot[bq] / cnt2;
    const int r = unit_to_row(map, b, f);
    if (r >= 0) v = fmaf(dX[((size_t)t * R + r) * K + (K - 1)], s, v);
    const float y = fbz[i];
    if (act == FSN_ACT_RELU) v = y > 0.f ? v : 0.f;
    else if (act == FSN_ACT_TANH) v *= 1.f - y * y;
    else if (act == FSN_ACT_RELU6) v = (y > 0.f && y < 6.f) ? v : 0.f;
    dz[i] = v;
  }
}

// ------------------------------------------------------------------------------------------ workspace

struct TrainWs {
  float *raw, *xfb, *fbz, *inv1, *inv2;
  float2 *sums_mag, *sums_fb;
  LayerSave fb[2], sb[2];
  float *xsb, *dxsb, *dout, *dz, *dfh1;
  float *dh_rec[2], *dc[2], *dh_mid, *dot;
  float *splitk, *colsum;
  float *splitk2, *colsum2;  // scratch of the side stream (full-band backward overlapped with the sub-band weight gradients)
  // FSN_PREC_TF32_TC: transposed weights ([H,4H], [K0,4H]) and transposed dG / layer inputs for the weight gradients
  float *sb_whhT[2], *sb_wihT[2], *fb_whhT[2], *fb_wihT1;
  float *gT, *xT, *rec;
  __half *fb_h16[2], *sb_h16[2], *w16;  // fp16 MMA operands of the forward step kernel (hidden states, weights)
  float *cum1, *cum2, *dunit;  // cumulative norm: scale of (step, clip), of (step, unit); gradient of the fb row per unit
  float2* fs;
  size_t bytes;
};

struct TCarver {
  char* base; size_t off;
  explicit TCarver(void* p) : base((char*)p), off(0) {}
  float* take(size_t n) {
    float* r = base ? (float*)(base + off) : nullptr;
    off = align_up(off + n * sizeof(float), 256);
    return r;
  }
};

static void carve_train(const fsn_model_desc* d, const Dims& m, void* base, TrainWs& w) {
  TCarver c(base);
  const size_t Tp = m.Tp, B = m.B, F = m.F, R = m.R, Hf = d->fb_hidden, Hs = d->sb_hidden;
  w.raw = c.take(Tp * B * F); w.xfb = c.take(Tp * B * F); w.fbz = c.take(Tp * B * F);
  w.inv1 = c.take(B); w.inv2 = c.take(B);
  w.sums_mag = (float2*)c.take(2 * B); w.sums_fb = (float2*)c.take(2 * B);
  for (int l = 0; l < 2; ++l) {
    w.fb[l].G = c.take(Tp * B * 4 * Hf); w.fb[l].C = c.take(Tp * B * Hf); w.fb[l].H = c.take(Tp * B * Hf);
    w.sb[l].G = c.take(Tp * R * 4 * Hs); w.sb[l].C = c.take(Tp * R * Hs); w.sb[l].H = c.take(Tp * R * Hs);
  }
  w.xsb = c.take(Tp * R * m.Ksb); w.dxsb = c.take(Tp * R * m.Ksb);
  w.dout = c.take(Tp * R * 2);
  w.dz = c.take(Tp * B * F); w.dfh1 = c.take(Tp * B * Hf);
  const size_t RH = (R * Hs > B * Hf) ? R * Hs : B * Hf;
  for (int i = 0; i < 2; ++i) { w.dh_rec[i] = c.take(RH); w.dc[i] = c.take(RH); }
  w.dh_mid = c.take(RH);
  w.dot = c.take(B);
  w.cum1 = w.cum2 = w.dunit = nullptr; w.fs = nullptr;
  if (d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE) {
    w.cum1 = c.take(Tp * B); w.cum2 = c.take(Tp * R); w.dunit = c.take(Tp * R);
    w.fs = (float2*)c.take(2 * Tp * B);
  }
  w.splitk = c.take(SPLITK_SCRATCH_FLOATS);
  const size_t maxcols = 4 * (Hf > Hs ? Hf : Hs) > F ? 4 * (Hf > Hs ? Hf : Hs) : F;
  w.colsum = c.take((size_t)COLSUM_MAX_S * maxcols);
  w.splitk2 = c.take(SPLITK_SCRATCH_FLOATS);
  w.colsum2 = c.take((size_t)COLSUM_MAX_S * maxcols);
  if (d->precision == FSN_PREC_TF32_TC) {
    for (int l = 0; l < 2; ++l) {
      w.sb_whhT[l] = c.take(Hs * 4 * Hs);
      w.sb_wihT[l] = c.take((l == 0 ? (size_t)m.Ksb : Hs) * 4 * Hs);
      w.fb_whhT[l] = c.take(Hf * 4 * Hf);
    }
    w.fb_wihT1 = c.take(Hf * 4 * Hf);
    // K-major (block-tiled, zero padded: tgemm_blocked_floats) copies of dG and of the layer input / hidden states
    const size_t g_sb = tgemm_blocked_floats(Tp * R, 4 * (int)Hs), g_fb = tgemm_blocked_floats(Tp * B, 4 * (int)Hf);
    w.gT = c.take(g_sb > g_fb ? g_sb : g_fb);
    const size_t x_sb = tgemm_blocked_floats(Tp * R, Hs > (size_t)m.Ksb ? (int)Hs : m.Ksb),
                 x_fb = tgemm_blocked_floats(Tp * B, Hf > F ? (int)Hf : (int)F);
    w.xT = c.take(x_sb > x_fb ? x_sb : x_fb);
    w.rec = c.take(4 * RH);
    for (int l = 0; l < 2; ++l) {
      w.fb_h16[l] = (__half*)c.take((Tp * B * Hf + 1) / 2);
      w.sb_h16[l] = (__half*)c.take((Tp * R * Hs + 1) / 2);
    }
    const size_t wmax = Hf > Hs ? Hf : Hs;
    w.w16 = (__half*)c.take((4 * wmax * 2 * wmax + 1) / 2);
  } else {
    w.fb_h16[0] = w.fb_h16[1] = w.sb_h16[0] = w.sb_h16[1] = w.w16 = nullptr;
  }
  w.bytes = c.off;
}

static int train_check(const fsn_model_desc* d) {
  FSN_REQUIRE(d->norm_type == FSN_NORM_OFFLINE_LAPLACE || d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE, FSN_ERR_UNSUPPORTED,
              "training: offline_laplace_norm and cumulative_laplace_norm are built");
  FSN_REQUIRE(d->fb_num_neighbors == 0, FSN_ERR_UNSUPPORTED,
              "training: fb_num_neighbors > 0 is not built (every shipped recipe uses 0)");
  FSN_REQUIRE(d->cell_type == FSN_CELL_LSTM, FSN_ERR_UNSUPPORTED, "training: the GRU cell is built for inference only");
  return FSN_OK;
}

// one layer forward over all steps, saving gates / cell / hidden:  X [Tp,R,K0] (row_scale == nullptr)
static int layer_forward_save(const fsn_seq_weights* w, int l, const float* X, int R, int K0, int H, int Tp,
                              const LayerSave& s, cudaStream_t st) {
  for (int t = 0; t < Tp; ++t) {
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.R = R; p.K0 = K0; p.H = H; p.first = (t == 0);
    p.w_ih = w->w_ih[l]; p.w_hh = w->w_hh[l]; p.b_ih = w->b_ih[l]; p.b_hh = w->b_hh[l];
    p.x0 = X + (size_t)t * R * K0; p.x0_row_stride = K0;
    p.h_prev = s.H + (size_t)(t > 0 ? t - 1 : 0) * R * H; p.h_prev_stride = H;
    p.h_out = s.H + (size_t)t * R * H; p.h_out_stride = H;
    p.c = s.C + (size_t)t * R * H;
    p.c_in = s.C + (size_t)(t > 0 ? t - 1 : 0) * R * H;
    p.save_gates = s.G + (size_t)t * R * 4 * H;
    int rc = lstm_step_launch(p, SEG0_DENSE, st);
    if (rc) return rc;
  }
  return FSN_OK;
}

// tensor-core variant.  Default: ONE kernel per step (lstm_fwd_step_kernel, fsn_tgemm.cu): [x_t | h_{t-1}] [W_ih | W_hh]^T on
// tcgen05 and the cell in its epilogue, gates / cell / hidden saved.  Fallbacks, step by step: a layer input whose rows
// are not 16-byte aligned keeps a hoisted projection of all steps (one GEMM into the gate buffer) that the step kernel
// adds; with the fused kernel switched off (or H % 32 != 0) every step is a recurrent GEMM into `rec` + lstm_cell_fwd_kernel
int layer_forward_save_tc(const fsn_seq_weights* w, int l, const float* X, int R, int K0, int H, int Tp,
                                 const LayerSave& s, float* rec, cudaStream_t st, float* splitk, size_t splitk_floats,
                                 const LayerHalf* half) {
  int rc;
  const int rows = Tp * R;
  static const int fused_min_rows = getenv("FSN_TRAIN_FUSED_MIN_ROWS") ? atoi(getenv("FSN_TRAIN_FUSED_MIN_ROWS")) : 1;
  const bool fused = R >= fused_min_rows && lstm_fwd_step_supported(s.H, w->w_hh[l], H);
  // x_t W_ih^T as leading k blocks of the step kernel - no hoisted projection, G is written once and never read in the
  // forward pass (FSN_TRAIN_FOLD_K bounds the input width this is done for)
  const bool fold = fused && lstm_fwd_step_folds_input(X, w->w_ih[l], K0);
  // fp16 MMA operands: h_t (written by the step kernel next to the fp32 copy) and the weights; the folded layer input too
  // when the layer below left an fp16 copy (K0 % 8: 16-byte rows)
  const bool h16 = fused && half && half->H16 && half->w16 && lstm_fwd_step_half_enabled(H);
  const bool x16 = h16 && fold && half->X16 && (K0 % 8) == 0;
  __half* w_hh16 = h16 ? half->w16 : nullptr;
  __half* w_ih16 = x16 ? half->w16 + (size_t)4 * H * H : nullptr;
  if (h16 && (rc = to_half_launch(w->w_hh[l], (size_t)4 * H * H, w_hh16, st))) return rc;
  if (x16 && (rc = to_half_launch(w->w_ih[l], (size_t)4 * H * K0, w_ih16, st))) return rc;
  if (fold) {
  } else if (tgemm_supported(X, K0, w->w_ih[l], K0, K0)) {
    if ((rc = tgemm_launch(X, K0, w->w_ih[l], K0, s.G, 4 * H, rows, 4 * H, K0, false, nullptr, 0, st))) return rc;
  } else if ((rc = fc_gemm_launch(X, w->w_ih[l], nullptr, s.G, rows, K0, 4 * H, FSN_ACT_NONE, st))) {
    return rc;  // rows of X not 16-byte aligned (K0 % 4 != 0): fp32 SIMT GEMM
  }
  const size_t n = (size_t)R * H;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  for (int t = 0; t < Tp; ++t) {
    float* Gt = s.G + (size_t)t * R * 4 * H;
    if (fused && (t > 0 || fold)) {  // GEMM + cell in one kernel, the recurrent product stays in TMEM
      LstmStepHalf hs{nullptr, nullptr, nullptr, nullptr, nullptr};
      if (h16) {
        hs.Hprev16 = t > 0 ? half->H16 + (size_t)(t - 1) * R * H : nullptr;
        hs.w_hh16 = w_hh16;
        hs.H16_out = half->H16 + (size_t)t * R * H;
        if (x16) { hs.Xt16 = half->X16 + (size_t)t * R * K0; hs.w_ih16 = w_ih16; }
      }
      if ((rc = lstm_fwd_step_launch(t > 0 ? s.H + (size_t)(t - 1) * R * H : nullptr, w->w_hh[l],
                                     fold ? X + (size_t)t * R * K0 : nullptr, w->w_ih[l], K0, Gt, w->b_ih[l], w->b_hh[l],
                                     t > 0 ? s.C + (size_t)(t - 1) * R * H : nullptr, s.C + (size_t)t * R * H,
                                     s.H + (size_t)t * R * H, R, H, st, h16 ? &hs : nullptr)))
        return rc;
      continue;
    }
    // (first step of a layer with a hoisted projection: no product at all, the plain cell kernel; its h_0 also goes out
    // in fp16 below)
    if (t > 0)
      if ((rc = tgemm_launch(s.H + (size_t)(t - 1) * R * H, H, w->w_hh[l], H, rec, 4 * H, R, 4 * H, H, false, splitk, splitk_floats, st)))
        return rc;
    lstm_cell_fwd_kernel<<<blocks, 256, 0, st>>>(Gt, t > 0 ? rec : nullptr, w->b_ih[l], w->b_hh[l],
                                                 t > 0 ? s.C + (size_t)(t - 1) * R * H : nullptr,
                                                 s.C + (size_t)t * R * H, s.H + (size_t)t * R * H, R, H);
    FSN_CHECK_LAUNCH("lstm_cell_fwd_kernel");
    if (h16 && (rc = to_half_launch(s.H + (size_t)t * R * H, (size_t)R * H, half->H16 + (size_t)t * R * H, st))) return rc;
  }
  return FSN_OK;
}

static bool tc_layer_ok(const fsn_model_desc* d, int H) { return d->precision == FSN_PREC_TF32_TC && (H & 3) == 0; }

struct LayerBwd {
  const float *w_ih, *w_hh;
  LayerSave s;
  int R, K0, H;
  float *dh_rec, *dc;
  const float *w_hhT, *w_ihT;  // tensor-core path: [H,4H] / [K0,4H] transposed copies (else nullptr)
  float* splitk;               // split-K space of the per-step GEMMs (used when the layer has only a few tiles)
};

// step t of one layer: pointwise gate gradients, then dh_rec = dG W_hh and (optionally) dx = dG W_ih
static int layer_bwd_step(const LayerBwd& L, int t, int Tp, const float* dh_above, const float* dout, const float* fc_w,
                          int O, float* dx, cudaStream_t st) {
  BwdPoint p;
  memset(&p, 0, sizeof(p));
  p.R = L.R; p.H = L.H;
  p.G = L.s.G + (size_t)t * L.R * 4 * L.H;
  p.C = L.s.C + (size_t)t * L.R * L.H;
  p.C_prev = t > 0 ? L.s.C + (size_t)(t - 1) * L.R * L.H : nullptr;
  p.dh_above = dh_above;
  p.dh_rec = (t == Tp - 1) ? nullptr : L.dh_rec;
  p.dc = L.dc; p.first_dc = (t == Tp - 1);
  p.dout = dout; p.fc_w = fc_w; p.O = O;
  const size_t n = (size_t)L.R * L.H;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  lstm_bwd_point_kernel<<<blocks, 256, 0, st>>>(p);
  FSN_CHECK_LAUNCH("lstm_bwd_point_kernel");
  int rc;
  if (t > 0) {
    if (L.w_hhT) rc = tgemm_launch(p.G, 4 * L.H, L.w_hhT, 4 * L.H, L.dh_rec, L.H, L.R, L.H, 4 * L.H, false, L.splitk, SPLITK_SCRATCH_FLOATS, st);
    else         rc = sgemm_launch(false, p.G, 4 * L.H, L.w_hh, L.H, L.dh_rec, L.H, L.R, L.H, 4 * L.H, false, nullptr, st);
    if (rc) return rc;
  }
  if (dx) {
    if (L.w_ihT) rc = tgemm_launch(p.G, 4 * L.H, L.w_ihT, 4 * L.H, dx, L.K0, L.R, L.K0, 4 * L.H, false, L.splitk, SPLITK_SCRATCH_FLOATS, st);
    else         rc = sgemm_launch(false, p.G, 4 * L.H, L.w_ih, L.K0, dx, L.K0, L.R, L.K0, 4 * L.H, false, nullptr, st);
    if (rc) return rc;
  }
  return FSN_OK;
}

// weight / bias gradients of one layer from dG [Tp*R,4H] (in L.s.G), its input X [Tp*R,K0] and hidden states
static int layer_weight_grads(const LayerBwd& L, int Tp, const float* X, float* g_w_ih, float* g_w_hh, float* g_b_ih,
                              float* g_b_hh, const TrainWs& w, cudaStream_t st) {
  const int H4 = 4 * L.H;
  const int rows = Tp * L.R;
  int rc;
  if (L.w_hhT && tgemm_blocked_enabled()) {
    // tensor-core path, block-tiled K-major copies (one contiguous 16 KB burst per TMA box instead of 128 rows with a
    // pitch of `rows` floats): dW_ih = dG^T X, dW_hh = dG[1:]^T H[:-1]
    const int nkb = (rows + 31) / 32;
    int slabs = 0;  // bias gradients = column sums of dG, taken while its tiles pass through shared memory
    if ((rc = transpose_blocked_launch(L.s.G, (size_t)rows, H4, (size_t)H4, w.gT, st, w.colsum, COLSUM_MAX_S, &slabs))) return rc;
    colsum_final_kernel<<<cdiv(H4, 128), 128, 0, st>>>(w.colsum, slabs, H4, g_b_ih, g_b_hh);
    FSN_CHECK_LAUNCH("colsum_final_kernel");
    if ((rc = transpose_blocked_launch(X, (size_t)rows, L.K0, (size_t)L.K0, w.xT, st, nullptr, 0, nullptr))) return rc;
    if ((rc = tgemm_blocked_launch(w.gT, nkb, 0, w.xT, nkb, 0, g_w_ih, L.K0, H4, L.K0, rows, false, w.splitk, SPLITK_SCRATCH_FLOATS, st)))
      return rc;
    if (Tp > 1) {
      if ((rc = transpose_blocked_launch(L.s.H, (size_t)rows, L.H, (size_t)L.H, w.xT, st, nullptr, 0, nullptr))) return rc;
      int a_kb0 = L.R / 32, a_nkb = nkb;
      if (L.R & 31) {  // step offset not on a k block: a second copy that starts at step 1
        a_kb0 = 0; a_nkb = (rows - L.R + 31) / 32;
        if ((rc = transpose_blocked_launch(L.s.G + (size_t)L.R * H4, (size_t)(rows - L.R), H4, (size_t)H4, w.gT, st, nullptr, 0, nullptr)))
          return rc;
      }
      if ((rc = tgemm_blocked_launch(w.gT, a_nkb, a_kb0, w.xT, nkb, 0, g_w_hh, L.H, H4, L.H, rows - L.R, false, w.splitk,
                                     SPLITK_SCRATCH_FLOATS, st)))
        return rc;
    } else if ((rc = check_cuda(cudaMemsetAsync(g_w_hh, 0, (size_t)H4 * L.H * sizeof(float), st), "memset"))) {
      return rc;
    }
    return FSN_OK;
  }
  if (L.w_hhT && (L.R & 3) == 0) {
    // tensor-core path: K-major operands = transposed copies dG^T [4H, rows], X^T [K0, rows], H^T [H, rows]
    if ((rc = transpose_launch(L.s.G, (size_t)rows, H4, w.gT, st))) return rc;
    if ((rc = transpose_launch(X, (size_t)rows, L.K0, w.xT, st))) return rc;
    if ((rc = tgemm_launch(w.gT, rows, w.xT, rows, g_w_ih, L.K0, H4, L.K0, rows, false, w.splitk, SPLITK_SCRATCH_FLOATS, st)))
      return rc;
    if (Tp > 1) {
      if ((rc = transpose_launch(L.s.H, (size_t)rows, L.H, w.xT, st))) return rc;
      if ((rc = tgemm_launch(w.gT + L.R, rows, w.xT, rows, g_w_hh, L.H, H4, L.H, rows - L.R, false, w.splitk,
                             SPLITK_SCRATCH_FLOATS, st)))
        return rc;
    } else if ((rc = check_cuda(cudaMemsetAsync(g_w_hh, 0, (size_t)H4 * L.H * sizeof(float), st), "memset"))) {
      return rc;
    }
    return colsum_launch(L.s.G, (size_t)rows, H4, H4, g_b_ih, g_b_hh, w.colsum, st);
  }
  if ((rc = sgemm_launch(true, L.s.G, H4, X, L.K0, g_w_ih, L.K0, H4, L.K0, rows, false, w.splitk, st))) return rc;
  if (Tp > 1) {
    if ((rc = sgemm_launch(true, L.s.G + (size_t)L.R * H4, H4, L.s.H, L.H, g_w_hh, L.H, H4, L.H, rows - L.R, false,
                           w.splitk, st)))
      return rc;
  } else if ((rc = check_cuda(cudaMemsetAsync(g_w_hh, 0, (size_t)H4 * L.H * sizeof(float), st), "memset"))) {
    return rc;
  }
  return colsum_launch(L.s.G, (size_t)rows, H4, H4, g_b_ih, g_b_hh, w.colsum, st);
}

}  // namespace fsn

using namespace fsn;

extern "C" size_t fsn_train_workspace_bytes(const fsn_model_desc* d, int B, int T) {
  Dims m;
  if (make_dims(d, B, T, m)) return 0;
  TrainWs w;
  carve_train(d, m, nullptr, w);
  return w.bytes;
}

extern "C" int fsn_train_forward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                                 const float* noisy_mag, int B, int T, float* crm, void* workspace,
                                 size_t workspace_bytes, fsn_stream_t stream) {
  launch_counter() = 0;
  Dims m;
  int rc = make_dims(d, B, T, m);
  if (rc) return rc;
  if ((rc = train_check(d))) return rc;
  TrainWs w;
  carve_train(d, m, workspace, w);
  FSN_REQUIRE(workspace && workspace_bytes >= w.bytes, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = m.Tp, F = m.F, Hf = d->fb_hidden, Hs = d->sb_hidden;
  // first norm (model.py:92) and the time-major copies
  train_mag_stats_kernel<<<B, 256, 0, st>>>(noisy_mag, F, T, d->sb_num_neighbors, w.sums_mag);
  FSN_CHECK_LAUNCH("train_mag_stats_kernel");
  if ((rc = norm_scales_launch(w.sums_mag, w.sums_mag, B, (float)F * Tp, 1.f, w.inv1, nullptr, st))) return rc;
  train_transpose_kernel<<<dim3(cdiv(Tp, 32), cdiv(F, 32), B), dim3(32, 8), 0, st>>>(noisy_mag, w.inv1, w.raw, w.xfb, B,
                                                                                     F, T, Tp);
  FSN_CHECK_LAUNCH("train_transpose_kernel");
  const bool cum = d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE;
  const float cum_eps = 1.1920928955078125e-07f;  // audio_zen/constant.py:9
  if (cum) {  // causal running mean per clip instead of the clip mean (base_model.py:220-251)
    train_frame_sum_kernel<<<cdiv(Tp * B, 8), 256, 0, st>>>(w.raw, B, F, Tp, w.fs);
    FSN_CHECK_LAUNCH("train_frame_sum_kernel");
    if ((rc = cum_clip_scale_launch(w.fs, B, Tp, F, cum_eps, w.cum1, st))) return rc;
    train_scale_tm_kernel<<<148 * 8, 256, 0, st>>>(w.raw, w.cum1, F, (size_t)Tp * B * F, w.xfb);
    FSN_CHECK_LAUNCH("train_scale_tm_kernel");
  }
  // full-band stack + Linear/activation (model.py:92-95)
  const bool tc_fb = tc_layer_ok(d, Hf), tc_sb = tc_layer_ok(d, Hs);
  // fp16 operand copies: layer 0's hidden states double as layer 1's input
  const LayerHalf hf0{w.fb_h16[0], nullptr, w.w16}, hf1{w.fb_h16[1], w.fb_h16[0], w.w16};
  const LayerHalf hs0{w.sb_h16[0], nullptr, w.w16}, hs1{w.sb_h16[1], w.sb_h16[0], w.w16};
  if (tc_fb) {
    if ((rc = layer_forward_save_tc(fb, 0, w.xfb, B, F, Hf, Tp, w.fb[0], w.rec, st, w.splitk, SPLITK_SCRATCH_FLOATS, &hf0))) return rc;
    if ((rc = layer_forward_save_tc(fb, 1, w.fb[0].H, B, Hf, Hf, Tp, w.fb[1], w.rec, st, w.splitk, SPLITK_SCRATCH_FLOATS, &hf1))) return rc;
  } else {
    if ((rc = layer_forward_save(fb, 0, w.xfb, B, F, Hf, Tp, w.fb[0], st))) return rc;
    if ((rc = layer_forward_save(fb, 1, w.fb[0].H, B, Hf, Hf, Tp, w.fb[1], st))) return rc;
  }
  if ((rc = fc_gemm_launch(w.fb[1].H, fb->fc_w, fb->fc_b, w.fbz, Tp * B, Hf, F, d->fb_activation, st))) return rc;
  // second norm in closed form (model.py:110-111)
  train_tm_stats_kernel<<<B, 256, 0, st>>>(w.fbz, B, F, Tp, d->fb_num_neighbors, w.sums_fb);
  FSN_CHECK_LAUNCH("train_tm_stats_kernel");
  if ((rc = norm_scales_launch(w.sums_mag, w.sums_fb, B, 1.f, (float)F * m.Ksb * Tp, nullptr, w.inv2, st))) return rc;
  // sub-band units (unfold + concat + norm + drop_band as one gather), then the sub-band stack (model.py:98-128)
  RowMap map{B, F, m.Fsub, m.G};
  if (cum && (rc = cum_unit_scale_launch(w.raw, w.fbz, map, m.R, Tp, d->sb_num_neighbors, d->fb_num_neighbors, cum_eps, w.cum2,
                                         st, /*time_major=*/true)))
    return rc;
  train_gather_kernel<<<148 * 8, 256, 0, st>>>(w.raw, w.fbz, w.inv2, cum ? w.cum2 : nullptr, w.xsb, map, Tp, m.R,
                                               d->sb_num_neighbors, d->fb_num_neighbors);
  FSN_CHECK_LAUNCH("train_gather_kernel");
  if (tc_sb) {
    if ((rc = layer_forward_save_tc(sb, 0, w.xsb, m.R, m.Ksb, Hs, Tp, w.sb[0], w.rec, st, w.splitk, SPLITK_SCRATCH_FLOATS, &hs0))) return rc;
    if ((rc = layer_forward_save_tc(sb, 1, w.sb[0].H, m.R, Hs, Hs, Tp, w.sb[1], w.rec, st, w.splitk, SPLITK_SCRATCH_FLOATS, &hs1))) return rc;
  } else {
    if ((rc = layer_forward_save(sb, 0, w.xsb, m.R, m.Ksb, Hs, Tp, w.sb[0], st))) return rc;
    if ((rc = layer_forward_save(sb, 1, w.sb[0].H, m.R, Hs, Hs, Tp, w.sb[1], st))) return rc;
  }
  // sub-band Linear of every output frame in one launch (model.py:129-135; the first look_ahead steps have no frame)
  return sb_fc_steps_launch(w.sb[1].H + (size_t)d->look_ahead * m.R * Hs, m.R, Hs, Tp - d->look_ahead, sb->fc_w, sb->fc_b, 2,
                            d->sb_activation, crm, m.Fsub, m.T, 0, st);
}

namespace fsn {
// Side stream of the backward pass: the full-band BPTT is a chain of ~1 900 tiny launches (64 rows), the sub-band weight
// gradients are a dozen HBM-bound kernels with no dependency on it - they run side by side.  One high-priority
// non-blocking stream and two events per device, created on first use; fork / join through events only, so the pattern
// is also legal inside a stream capture.
struct SideStream { cudaStream_t s; cudaEvent_t fork, join; bool ok; };
static SideStream* side_stream() {
  static SideStream per_dev[64] = {};
  static bool tried[64] = {};
  static const bool enabled = getenv("FSN_TRAIN_OVERLAP") == nullptr || atoi(getenv("FSN_TRAIN_OVERLAP")) != 0;
  if (!enabled) return nullptr;
  int dev = 0;
  cudaGetDevice(&dev);
  SideStream& x = per_dev[dev & 63];
  if (!tried[dev & 63]) {
    tried[dev & 63] = true;
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    x.ok = cudaStreamCreateWithPriority(&x.s, cudaStreamNonBlocking, hi) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
    if (!x.ok) cudaGetLastError();
  }
  return x.ok ? &x : nullptr;
}
}  // namespace fsn

extern "C" int fsn_train_backward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                                  const float* dcrm, int B, int T, const fsn_seq_grads* gfb, const fsn_seq_grads* gsb,
                                  void* workspace, size_t workspace_bytes, fsn_stream_t stream) {
  launch_counter() = 0;
  Dims m;
  int rc = make_dims(d, B, T, m);
  if (rc) return rc;
  if ((rc = train_check(d))) return rc;
  FSN_REQUIRE(d->sb_activation == FSN_ACT_NONE, FSN_ERR_UNSUPPORTED,
              "training: sb_output_activate_function must be off (as in every shipped recipe)");
  TrainWs w;
  carve_train(d, m, workspace, w);
  FSN_REQUIRE(workspace && workspace_bytes >= w.bytes, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int Tp = m.Tp, F = m.F, R = m.R, Hf = d->fb_hidden, Hs = d->sb_hidden, K = m.Ksb;
  RowMap map{B, F, m.Fsub, m.G};
  // ---- sub-band Linear (model.py:129-135 backwards)
  train_dout_kernel<<<148 * 8, 256, 0, st>>>(dcrm, w.dout, R, m.Fsub, T, Tp, d->look_ahead);
  FSN_CHECK_LAUNCH("train_dout_kernel");
  {  // dW of the 2-output Linear: one streaming pass over h1 (2.4 GB at config 3)
    const size_t rows = (size_t)Tp * R;
    int S = (int)((rows + 2047) / 2048);
    if (S > COLSUM_MAX_S) S = COLSUM_MAX_S;
    while (S > 1 && (size_t)S * 2 * Hs > SPLITK_SCRATCH_FLOATS) --S;
    const size_t rows_per = (rows + S - 1) / S;
    small_out_wgrad_kernel<2><<<dim3(cdiv(Hs, 128), S), 128, 0, st>>>(w.dout, w.sb[1].H, rows, Hs, rows_per, w.splitk);
    FSN_CHECK_LAUNCH("small_out_wgrad_kernel");
    colsum_final_kernel<<<cdiv(2 * Hs, 128), 128, 0, st>>>(w.splitk, S, 2 * Hs, gsb->fc_w, nullptr);
    FSN_CHECK_LAUNCH("colsum_final_kernel");
  }
  if ((rc = colsum_launch(w.dout, (size_t)Tp * R, 2, 2, gsb->fc_b, nullptr, w.colsum, st))) return rc;
  // ---- sub-band stack, both layers one step apart
  const bool tc_fb = tc_layer_ok(d, Hf), tc_sb = tc_layer_ok(d, Hs);
  if (tc_sb) {
    for (int l = 0; l < 2; ++l) {
      if ((rc = transpose_launch(sb->w_hh[l], (size_t)4 * Hs, Hs, w.sb_whhT[l], st))) return rc;
      if ((rc = transpose_launch(sb->w_ih[l], (size_t)4 * Hs, l == 0 ? K : Hs, w.sb_wihT[l], st))) return rc;
    }
  }
  if (tc_fb) {
    for (int l = 0; l < 2; ++l)
      if ((rc = transpose_launch(fb->w_hh[l], (size_t)4 * Hf, Hf, w.fb_whhT[l], st))) return rc;
    if ((rc = transpose_launch(fb->w_ih[1], (size_t)4 * Hf, Hf, w.fb_wihT1, st))) return rc;
  }
  LayerBwd s1{sb->w_ih[1], sb->w_hh[1], w.sb[1], R, Hs, Hs, w.dh_rec[1], w.dc[1], tc_sb ? w.sb_whhT[1] : nullptr,
              tc_sb ? w.sb_wihT[1] : nullptr, w.splitk};
  LayerBwd s0{sb->w_ih[0], sb->w_hh[0], w.sb[0], R, K, Hs, w.dh_rec[0], w.dc[0], tc_sb ? w.sb_whhT[0] : nullptr,
              tc_sb ? w.sb_wihT[0] : nullptr, w.splitk};
  for (int t = Tp - 1; t >= 0; --t) {
    if ((rc = layer_bwd_step(s1, t, Tp, nullptr, w.dout + (size_t)t * R * 2, sb->fc_w, 2, w.dh_mid, st))) return rc;
    if ((rc = layer_bwd_step(s0, t, Tp, w.dh_mid, nullptr, nullptr, 0, w.dxsb + (size_t)t * R * K, st))) return rc;
  }
  // ---- fork: sub-band weight gradients on the caller's stream, the rest of the chain (second norm, full-band Linear, full-band
  // BPTT) on the side stream with its own split-K / column-sum scratch
  SideStream* side = side_stream();
  cudaStream_t st2 = st;
  float *splitk2 = w.splitk, *colsum2 = w.colsum;
  if (side) {
    if ((rc = check_cuda(cudaEventRecord(side->fork, st), "event record"))) return rc;
    if ((rc = check_cuda(cudaStreamWaitEvent(side->s, side->fork, 0), "stream wait"))) return rc;
    st2 = side->s; splitk2 = w.splitk2; colsum2 = w.colsum2;
  }
  if ((rc = layer_weight_grads(s1, Tp, w.sb[0].H, gsb->w_ih[1], gsb->w_hh[1], gsb->b_ih[1], gsb->b_hh[1], w, st)))
    return rc;
  if ((rc = layer_weight_grads(s0, Tp, w.xsb, gsb->w_ih[0], gsb->w_hh[0], gsb->b_ih[0], gsb->b_hh[0], w, st))) return rc;
  // ---- second norm + drop_band + full-band Linear/activation
  if (d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE) {
    train_cum_unit_bwd_kernel<<<cdiv(R, 128), 128, 0, st2>>>(w.dxsb, w.xsb, w.cum2, Tp, R, K, w.dunit);
    FSN_CHECK_LAUNCH("train_cum_unit_bwd_kernel");
    train_dfbz_cum_kernel<<<148 * 8, 256, 0, st2>>>(w.dunit, w.fbz, map, Tp, R, d->fb_activation, w.dz);
    FSN_CHECK_LAUNCH("train_dfbz_cum_kernel");
  } else {
    train_dot_kernel<<<B, 256, 0, st2>>>(w.dxsb, w.xsb, Tp, R, m.Fsub, K, w.dot);
    FSN_CHECK_LAUNCH("train_dot_kernel");
    train_dfbz_kernel<<<148 * 8, 256, 0, st2>>>(w.dxsb, w.fbz, w.inv2, w.dot, map, Tp, R, K, (float)F * K * Tp,
                                                d->fb_activation, w.dz);
    FSN_CHECK_LAUNCH("train_dfbz_kernel");
  }
  if ((rc = sgemm_launch(true, w.dz, F, w.fb[1].H, Hf, gfb->fc_w, Hf, F, Hf, Tp * B, false, splitk2, st2))) return rc;
  if ((rc = colsum_launch(w.dz, (size_t)Tp * B, F, F, gfb->fc_b, nullptr, colsum2, st2))) return rc;
  if ((rc = sgemm_launch(false, w.dz, F, fb->fc_w, Hf, w.dfh1, Hf, Tp * B, Hf, F, false, nullptr, st2))) return rc;
  // ---- full-band stack
  LayerBwd f1{fb->w_ih[1], fb->w_hh[1], w.fb[1], B, Hf, Hf, w.dh_rec[1], w.dc[1], tc_fb ? w.fb_whhT[1] : nullptr,
              tc_fb ? w.fb_wihT1 : nullptr, splitk2};
  LayerBwd f0{fb->w_ih[0], fb->w_hh[0], w.fb[0], B, F, Hf, w.dh_rec[0], w.dc[0], tc_fb ? w.fb_whhT[0] : nullptr, nullptr, splitk2};
  for (int t = Tp - 1; t >= 0; --t) {
    if ((rc = layer_bwd_step(f1, t, Tp, w.dfh1 + (size_t)t * B * Hf, nullptr, nullptr, 0, w.dh_mid, st2))) return rc;
    if ((rc = layer_bwd_step(f0, t, Tp, w.dh_mid, nullptr, nullptr, 0, nullptr, st2))) return rc;
  }
  if (side) {  // join: the full-band weight gradients share gT / xT / splitk / colsum with the sub-band ones
    if ((rc = check_cuda(cudaEventRecord(side->join, side->s), "event record"))) return rc;
    if ((rc = check_cuda(cudaStreamWaitEvent(st, side->join, 0), "stream wait"))) return rc;
  }
  if ((rc = layer_weight_grads(f1, Tp, w.fb[0].H, gfb->w_ih[1], gfb->w_hh[1], gfb->b_ih[1], gfb->b_hh[1], w, st)))
    return rc;
  return layer_weight_grads(f0, Tp, w.xfb, gfb->w_ih[0], gfb->w_hh[0], gfb->b_ih[0], gfb->b_hh[0], w, st);
}

// ------------------------------------------------------------------------------------------ loss
namespace fsn {
// loss = mean((cirm - crm)^2) with cirm [B,Fs,T,2] (trainer.py:49-54) and crm [B,2,Fs,T] (Model.forward);
// dcrm = 2 (crm - cirm) / n.  Stage 1: per-CTA partial sums; stage 2: fixed-order sum.
__global__ void mse_part_kernel(const float* __restrict__ cirm, const float* __restrict__ crm, int Fs, int T, size_t n,
                                float* __restrict__ dcrm, float* __restrict__ part) {
  __shared__ float sh[256];
  float a = 0.f;
  const float k = 2.0f / (float)n;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    // i indexes crm [b,o,f,t]
    const int t = (int)(i % T);
    size_t q = i / T;
    const int f = (int)(q % Fs); q /= Fs;
    const int o = (int)(q & 1);
    const size_t b = q >> 1;
    const float dlt = crm[i] - cirm[((b * Fs + f) * T + t) * 2 + o];
    a = fmaf(dlt, dlt, a);
    if (dcrm) dcrm[i] = k * dlt;
  }
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[blockIdx.x] = sh[0];
}
__global__ void mse_final_kernel(const float* __restrict__ part, int nb, size_t n, float* __restrict__ loss) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) a += (double)part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(sh[0] / (double)n);
}
constexpr int MSE_BLOCKS = 1024;
}  // namespace fsn

extern "C" size_t fsn_mse_loss_scratch_bytes(void) { return MSE_BLOCKS * sizeof(float); }

extern "C" int fsn_mse_loss(const float* cirm, const float* crm, int B, int Fsub, int T, float* loss, float* dcrm,
                            void* scratch, size_t scratch_bytes, fsn_stream_t stream) {
  FSN_REQUIRE(B > 0 && Fsub > 0 && T > 0, FSN_ERR_SHAPE, "mse_loss: empty input");
  FSN_REQUIRE(scratch && scratch_bytes >= MSE_BLOCKS * sizeof(float), FSN_ERR_WORKSPACE, "mse_loss: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * 2 * Fsub * T;
  int nb = (int)((n + 255) / 256);
  if (nb > MSE_BLOCKS) nb = MSE_BLOCKS;
  mse_part_kernel<<<nb, 256, 0, st>>>(cirm, crm, Fsub, T, n, dcrm, (float*)scratch);
  FSN_CHECK_LAUNCH("mse_part_kernel");
  mse_final_kernel<<<1, 256, 0, st>>>((const float*)scratch, nb, n, loss);
  FSN_CHECK_LAUNCH("mse_final_kernel");
  return FSN_OK;
}

// ------------------------------------------------------------------------------------------ clip + Adam
namespace fsn {
constexpr int ADAM_CHUNKS = 32;

__global__ void gradsq_part_kernel(const fsn_param_list L, float* __restrict__ part) {
  __shared__ float sh[256];
  const int ti = blockIdx.y;
  const float* g = L.grad[ti];
  const int64_t n = L.numel[ti];
  float a = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) a = fmaf(g[i], g[i], a);
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) part[ti * ADAM_CHUNKS + blockIdx.x] = sh[0];
}

// out[0] = total L2 norm of (grad * grad_scale); out[1] = grad_scale * min(1, max_norm / (norm + 1e-6))
__global__ void gradnorm_final_kernel(const float* __restrict__ part, int n, float grad_scale, float max_norm,
                                      float* __restrict__ out) {
  __shared__ double sh[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += (double)part[i];
  sh[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const float norm = (float)sqrt(sh[0]) * grad_scale;
    float coef = max_norm > 0.f ? max_norm / (norm + 1e-6f) : 1.f;
    if (coef > 1.f) coef = 1.f;
    out[0] = norm;
    out[1] = coef * grad_scale;
  }
}

// torch.optim.Adam single-tensor update (no weight decay / amsgrad); the clipped, scaled gradient is written back
__global__ void adam_kernel(const fsn_param_list L, const float* __restrict__ coef_ptr, float lr, float b1, float b2,
                            float eps, float bc1, float bc2_sqrt) {
  const int ti = blockIdx.y;
  float* p = L.param[ti];
  float* g = L.grad[ti];
  float* m = L.exp_avg[ti];
  float* v = L.exp_avg_sq[ti];
  const int64_t n = L.numel[ti];
  const float coef = coef_ptr[1];
  const float step_size = lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    const float gi = g[i] * coef;
    g[i] = gi;
    const float mi = m[i] * b1 + (1.f - b1) * gi;   // lerp(m, g, 1 - b1)
    const float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] -= step_size * (mi / denom);
  }
}
}  // namespace fsn

extern "C" size_t fsn_clip_adam_scratch_bytes(void) { return (FSN_MAX_PARAM_TENSORS * ADAM_CHUNKS + 2) * sizeof(float); }

extern "C" int fsn_clip_adam(const fsn_param_list* L, float max_norm, float grad_scale, float lr, float beta1,
                             float beta2, float eps, int step, float* norm_out, void* scratch, size_t scratch_bytes,
                             fsn_stream_t stream) {
  FSN_REQUIRE(L && L->n > 0 && L->n <= FSN_MAX_PARAM_TENSORS, FSN_ERR_SHAPE, "clip_adam: 1..%d tensors",
              FSN_MAX_PARAM_TENSORS);
  FSN_REQUIRE(step >= 1, FSN_ERR_SHAPE, "clip_adam: step starts at 1");
  FSN_REQUIRE(scratch && scratch_bytes >= fsn_clip_adam_scratch_bytes(), FSN_ERR_WORKSPACE, "clip_adam: scratch too small");
  cudaStream_t st = (cudaStream_t)stream;
  float* part = (float*)scratch;
  float* res = norm_out ? norm_out : part + FSN_MAX_PARAM_TENSORS * ADAM_CHUNKS;
  gradsq_part_kernel<<<dim3(ADAM_CHUNKS, L->n), 256, 0, st>>>(*L, part);
  FSN_CHECK_LAUNCH("gradsq_part_kernel");
  gradnorm_final_kernel<<<1, 256, 0, st>>>(part, L->n * ADAM_CHUNKS, grad_scale, max_norm, res);
  FSN_CHECK_LAUNCH("gradnorm_final_kernel");
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2 = 1.f - powf(beta2, (float)step);
  adam_kernel<<<dim3(ADAM_CHUNKS * 4, L->n), 256, 0, st>>>(*L, res, lr, beta1, beta2, eps, bc1, sqrtf(bc2));
  FSN_CHECK_LAUNCH("adam_kernel");
  return FSN_OK;
}
