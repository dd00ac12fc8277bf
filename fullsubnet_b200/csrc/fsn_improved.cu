// improved_fullsubnet (recipes/dns_interspeech_2020/improved_fullsubnet/model.py:452-591, BASELINE config 5):
// host orchestration + the section unfold / output kernels, on top of the shared fp32 building blocks
// (STFT/iSTFT, persistent full-band LSTM, LSTM step kernel).
#include <string.h>

#include "fsn_internal.cuh"

namespace fsn {

// |X|^fdrc with the Nyquist bin dropped, time-major: mag [B,F,T] -> out [B,T,F-1]  (model.py:564-565)
__global__ void imp_compress_kernel(const float* __restrict__ mag, float* __restrict__ out, int F, int T, float fdrc) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, Fu = F - 1;
  const int f0 = blockIdx.y * 32, t0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int f = f0 + i, t = t0 + tx;
    float v = 0.f;
    if (f < Fu && t < T) {
      const float m = mag[((size_t)b * F + f) * T + t];
      v = (fdrc == 0.5f) ? sqrtf(m) : powf(m, fdrc);
    }
    tile[i][tx] = v;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int t = t0 + i, f = f0 + tx;
    if (t < T && f < Fu) out[((size_t)b * T + t) * Fu + f] = tile[tx][i];
  }
}

struct SecGeom { int lo, N, cs, ns, cf, nf, W; };

// section input (model.py:321-405, 425-442): unit n of clip b at frame t = noisy rows lo+n*cs-ns .. (+cs+2ns) and
// full-band rows lo+n*cf-nf .. (+cf+2nf), reflected (no edge repeat) at row 0 / row Fu-1.  One CTA per (b,t):
// writes X[t][b*N+n][w] and the per-(b,t) sum (for the section norm).
__global__ void imp_section_input_kernel(const float* __restrict__ magc, const float* __restrict__ fbT, int B, int T,
                                         int Fu, SecGeom g, float* __restrict__ X, float2* __restrict__ fs) {
  __shared__ float red[256];
  const int b = blockIdx.x / T, t = blockIdx.x % T;
  const int Wn = g.cs + 2 * g.ns;
  const size_t base = ((size_t)b * T + t) * Fu;
  float local = 0.f;
  for (int i = threadIdx.x; i < g.N * g.W; i += blockDim.x) {
    const int n = i / g.W, w = i - n * g.W;
    int row;
    const float* src;
    if (w < Wn) { row = g.lo + n * g.cs - g.ns + w; src = magc; }
    else        { row = g.lo + n * g.cf - g.nf + (w - Wn); src = fbT; }
    row = reflect_idx(row, Fu);
    const float v = src[base + row];
    X[((size_t)t * B * g.N + (size_t)b * g.N + n) * g.W + w] = v;
    local += v;
  }
  red[threadIdx.x] = local;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) fs[(size_t)b * T + t] = make_float2(red[0], red[0]);
}

// Linear(H -> 2c) of one section for one frame, written into crm[b, ch, lo + n*c + j, t] with o = ch*c + j
// (SubBandSequenceWrapper.forward, model.py:239-247); one warp per (row, output)
__global__ void imp_fc_step_kernel(const float* __restrict__ h, int R, int H, const float* __restrict__ W,
                                   const float* __restrict__ bias, int c, int N, int lo, int act, float* __restrict__ crm,
                                   int F, int T, int t) {
  const int O = 2 * c;
  const size_t wid = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (wid >= (size_t)R * O) return;
  const int row = (int)(wid / O), o = (int)(wid % O);
  const float* hp = h + (size_t)row * H;
  float s = 0.f;
  for (int k = lane; k < H; k += 32) s = fmaf(hp[k], W[(size_t)o * H + k], s);
  s = warp_sum(s);
  if (lane == 0) {
    s += bias[o];
    if (act == FSN_ACT_RELU) s = fmaxf(s, 0.f);
    const int b = row / N, n = row - b * N, ch = o / c, j = o - ch * c;
    crm[(((size_t)b * 2 + ch) * F + (lo + n * c + j)) * T + t] = s;
  }
}

// X[t][r][w] *= inv[b(r)] with r = b*N + n: element i of the [T, R*W] tensor belongs to clip (i % (R*W)) / (N*W)
__global__ void imp_scale_rows_kernel(float* __restrict__ X, const float* __restrict__ inv, size_t n, size_t per_t,
                                      size_t per_clip, int B) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = (i % per_t) / per_clip;
    X[i] *= inv[b < (size_t)B ? b : 0];
  }
}

struct ImpDims { int B, L, T, F, Fu, S; SecGeom sec[FSN_IMP_MAX_SECTIONS]; int maxRW, maxR; };

struct ImpWs {
  float *mag, *real, *imag, *crm, *magc, *fbT, *X, *inv1, *invs;
  float2 *fs, *sums;
  float *fb_pp, *fb_h1all;
  unsigned int* barrier;
  float *h0[2], *h1[2], *c0, *c1;
  float *fbs_h0[2], *fbs_c0, *fbs_c1;  // per-step full-band fallback
  LayerSave tc;                        // FSN_PREC_TF32_TC: gates / cell / hidden of every step of one layer
  float *tc_h1, *tc_rec;
  LstmTcWs fbtc;                       // FSN_PREC_TF32_TC: full band on the hoisted-GEMM + persistent-recurrence kernels
  float* fbtc_mid;
  size_t bytes;
};

struct ICarver {
  char* base; size_t off;
  explicit ICarver(void* p) : base((char*)p), off(0) {}
  template <class T> T* take(size_t n) {
    T* r = base ? (T*)(base + off) : nullptr;
    off = align_up(off + n * sizeof(T), 256);
    return r;
  }
};

static bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }

static int imp_dims(const fsn_improved_desc* d, int B, int L, ImpDims& m) {
  FSN_REQUIRE(d && B > 0 && L > 0, FSN_ERR_SHAPE, "improved model: empty input");
  FSN_REQUIRE((is_pow2(d->n_fft) && d->n_fft <= 2048) || (d->n_fft % 2 == 0 && d->n_fft >= 16 && d->n_fft <= 1200),
              FSN_ERR_UNSUPPORTED, "improved model: n_fft=%d unsupported (power of two <= 2048, or even and <= 1200)",
              d->n_fft);
  FSN_REQUIRE(d->num_freqs == d->n_fft / 2 + 1, FSN_ERR_SHAPE, "improved model: num_freqs != n_fft/2+1");
  FSN_REQUIRE(d->num_sections >= 1 && d->num_sections <= FSN_IMP_MAX_SECTIONS, FSN_ERR_SHAPE, "improved model: sections");
  FSN_REQUIRE(d->precision == FSN_PREC_FP32 || (d->precision == FSN_PREC_TF32_TC && (d->sb_hidden & 3) == 0),
              FSN_ERR_UNSUPPORTED, "improved model: precision must be FSN_PREC_FP32 or FSN_PREC_TF32_TC (sb_hidden %% 4 == 0)");
  m.B = B; m.L = L; m.T = 1 + L / d->hop_length; m.F = d->num_freqs; m.Fu = m.F - 1; m.S = d->num_sections;
  m.maxRW = 0; m.maxR = 0;
  for (int s = 0; s < m.S; ++s) {
    SecGeom& g = m.sec[s];
    g.lo = s == 0 ? 0 : d->freq_cutoffs[s - 1];
    const int hi = s == m.S - 1 ? m.Fu : d->freq_cutoffs[s];
    g.cs = d->sb_num_center[s]; g.ns = d->sb_num_neighbor[s]; g.cf = d->fb_num_center[s]; g.nf = d->fb_num_neighbor[s];
    FSN_REQUIRE(g.cs > 0 && hi > g.lo && (hi - g.lo) % g.cs == 0 && (hi - g.lo) % g.cf == 0, FSN_ERR_SHAPE,
                "The number of center frequencies should be divisible by the subband freqency interval.");
    FSN_REQUIRE(g.cs == g.cf, FSN_ERR_UNSUPPORTED, "improved model: sb/fb centre widths of a section must match");
    FSN_REQUIRE(g.ns < m.Fu && g.nf < m.Fu, FSN_ERR_SHAPE, "improved model: neighbours >= num_freqs");
    g.N = (hi - g.lo) / g.cs;
    g.W = (g.cs + 2 * g.ns) + (g.cf + 2 * g.nf);
    if (g.N * g.W > m.maxRW) m.maxRW = g.N * g.W;
    if (g.N > m.maxR) m.maxR = g.N;
  }
  return FSN_OK;
}

static void imp_carve(const fsn_improved_desc* d, const ImpDims& m, void* base, ImpWs& w) {
  ICarver c(base);
  const size_t BFT = (size_t)m.B * m.F * m.T, BT = (size_t)m.B * m.T;
  w.mag = c.take<float>(BFT); w.real = c.take<float>(BFT); w.imag = c.take<float>(BFT);
  w.crm = c.take<float>(2 * BFT);
  w.magc = c.take<float>(BT * m.Fu);
  w.fbT = c.take<float>(BT * m.Fu);
  w.X = c.take<float>(BT * m.maxRW);
  w.inv1 = c.take<float>(m.B); w.invs = c.take<float>(m.B);
  w.fs = c.take<float2>(BT); w.sums = c.take<float2>(m.B);
  w.fb_pp = c.take<float>((size_t)2 * 256 * d->fb_hidden);
  w.fb_h1all = c.take<float>(BT * d->fb_hidden);
  w.barrier = c.take<unsigned int>(64);
  const size_t RH = (size_t)m.B * m.maxR * d->sb_hidden;
  for (int i = 0; i < 2; ++i) { w.h0[i] = c.take<float>(RH); w.h1[i] = c.take<float>(RH); }
  w.c0 = c.take<float>(RH); w.c1 = c.take<float>(RH);
  const size_t BH = (size_t)m.B * d->fb_hidden;
  w.fbs_h0[0] = c.take<float>(BH); w.fbs_h0[1] = c.take<float>(BH);
  w.fbs_c0 = c.take<float>(BH); w.fbs_c1 = c.take<float>(BH);
  w.tc.G = w.tc.C = w.tc.H = w.tc_h1 = w.tc_rec = nullptr;
  if (d->precision == FSN_PREC_TF32_TC) {
    const size_t TR = (size_t)m.T * m.B * m.maxR;
    w.tc.G = c.take<float>(TR * 4 * d->sb_hidden);
    w.tc.C = c.take<float>(TR * d->sb_hidden);
    w.tc.H = c.take<float>(TR * d->sb_hidden);
    w.tc_h1 = c.take<float>(TR * d->sb_hidden);
    w.tc_rec = c.take<float>(4 * RH);
  }
  memset(&w.fbtc, 0, sizeof(w.fbtc));
  w.fbtc_mid = nullptr;
  if (d->precision == FSN_PREC_TF32_TC && lstm_rec_tc_supported(d->fb_hidden, false)) {
    lstm_tc_carve(c.base, c.off, BT, m.Fu > d->fb_hidden ? m.Fu : d->fb_hidden, d->fb_hidden, false, w.fbtc);
    w.fbtc_mid = c.take<float>(BT * d->fb_hidden);
  }
  w.bytes = c.off;
}

}  // namespace fsn

using namespace fsn;

extern "C" size_t fsn_improved_workspace_bytes(const fsn_improved_desc* d, int B, int L) {
  ImpDims m;
  if (imp_dims(d, B, L, m)) return 0;
  ImpWs w;
  imp_carve(d, m, nullptr, w);
  return w.bytes;
}

extern "C" int fsn_improved_forward(const fsn_improved_desc* d, const fsn_improved_weights* wt, const float* wav, int B,
                                    int L, float* enhanced, float* crm_out, void* workspace, size_t workspace_bytes,
                                    fsn_stream_t stream) {
  launch_counter() = 0;
  ImpDims m;
  int rc = imp_dims(d, B, L, m);
  if (rc) return rc;
  ImpWs w;
  imp_carve(d, m, workspace, w);
  FSN_REQUIRE(workspace && workspace_bytes >= w.bytes, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  const int T = m.T, F = m.F, Fu = m.Fu, Hf = d->fb_hidden, Hs = d->sb_hidden;
  const float eps = 1.1920928955078125e-07f;  // np.finfo(np.float32).eps (model.py:23,148)
  float* crm = crm_out ? crm_out : w.crm;
  // STFT (model.py:550-557), |X|^fdrc without the Nyquist bin (564-565)
  if ((rc = stft_launch(wav, B, L, d->n_fft, d->hop_length, d->win_length, w.mag, nullptr, w.real, w.imag, nullptr, 0, st)))
    return rc;
  {
    dim3 grid(cdiv(T, 32), cdiv(Fu, 32), B);
    imp_compress_kernel<<<grid, dim3(32, 8), 0, st>>>(w.mag, w.magc, F, T, d->fdrc);
    FSN_CHECK_LAUNCH("imp_compress_kernel");
  }
  // full band: norm (566) -> 2xLSTM + Linear (567)
  if ((rc = clip_stats_launch(w.magc, B, T, Fu, 0, w.fs, w.sums, st))) return rc;
  if ((rc = norm_scales_launch(w.sums, w.sums, B, (float)Fu * T, 1.f, w.inv1, nullptr, st, eps))) return rc;
  const bool fb_tc = w.fbtc_mid != nullptr;
  if (fb_tc) {
    // tensor cores: per layer one hoisted input-projection GEMM (tf32) + the persistent tcgen05 recurrence, Linear likewise
    fsn_lstm_layer L0{wt->fb.w_ih[0], wt->fb.w_hh[0], wt->fb.b_ih[0], wt->fb.b_hh[0]};
    fsn_lstm_layer L1{wt->fb.w_ih[1], wt->fb.w_hh[1], wt->fb.b_ih[1], wt->fb.b_hh[1]};
    if ((rc = lstm_layer_tc(L0, w.magc, (size_t)Fu, Fu, w.inv1, T, 0, B, T, Hf, false, w.fbtc, w.fbtc_mid, st))) return rc;
    if ((rc = lstm_layer_tc(L1, w.fbtc_mid, (size_t)Hf, Hf, nullptr, 1, 0, B, T, Hf, false, w.fbtc, w.fb_h1all, st))) return rc;
    if ((rc = linear_tc(w.fb_h1all, (size_t)Hf, Hf, wt->fb.fc_w, wt->fb.fc_b, Fu, d->fb_activation, w.fbT, (size_t)Fu,
                        (size_t)B * T, false, w.fbtc, st)))
      return rc;
  } else if (fb_persistent_supported(Fu, Hf, Hf)) {
    for (int b0 = 0; b0 < B; b0 += 256) {
      const int nb = (B - b0 < 256) ? B - b0 : 256;
      if ((rc = fb_persistent_launch(&wt->fb, w.magc + (size_t)b0 * T * Fu, w.inv1 + b0, w.fb_pp,
                                     w.fb_h1all + (size_t)b0 * T * Hf, w.barrier, nb, Fu, Hf, Hf, T, st)))
        return rc;
    }
  } else {
    // weights of a (Fu + 3 Hf) x 16 slice exceed one SM's shared memory (n_fft = 1024): per-step kernels
    for (int t = 0; t < T; ++t) {
      StepParams p;
      memset(&p, 0, sizeof(p));
      p.R = B; p.H = Hf; p.first = (t == 0);
      p.K0 = Fu;
      p.w_ih = wt->fb.w_ih[0]; p.w_hh = wt->fb.w_hh[0]; p.b_ih = wt->fb.b_ih[0]; p.b_hh = wt->fb.b_hh[0];
      p.h_prev = w.fbs_h0[(t + 1) & 1]; p.h_prev_stride = Hf;
      p.h_out = w.fbs_h0[t & 1]; p.h_out_stride = Hf;
      p.c = w.fbs_c0;
      p.x0 = w.magc + (size_t)t * Fu; p.x0_row_stride = (size_t)T * Fu; p.row_scale = w.inv1;
      if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
      p.K0 = Hf;
      p.w_ih = wt->fb.w_ih[1]; p.w_hh = wt->fb.w_hh[1]; p.b_ih = wt->fb.b_ih[1]; p.b_hh = wt->fb.b_hh[1];
      p.x0 = w.fbs_h0[t & 1]; p.x0_row_stride = Hf; p.row_scale = nullptr;
      p.h_prev = w.fb_h1all + (size_t)(t > 0 ? t - 1 : 0) * Hf; p.h_prev_stride = (size_t)T * Hf;
      p.h_out = w.fb_h1all + (size_t)t * Hf; p.h_out_stride = (size_t)T * Hf;
      p.c = w.fbs_c1;
      if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
    }
  }
  if (!fb_tc && (rc = fc_gemm_launch(w.fb_h1all, wt->fb.fc_w, wt->fb.fc_b, w.fbT, B * T, Hf, Fu, d->fb_activation, st)))
    return rc;
  // cRM, Nyquist row = 0 (572)
  if ((rc = check_cuda(cudaMemsetAsync(crm, 0, (size_t)2 * B * F * T * sizeof(float), st), "crm memset"))) return rc;
  // sub-band sections (408-447)
  for (int s = 0; s < m.S; ++s) {
    const SecGeom& g = m.sec[s];
    const int R = B * g.N;
    imp_section_input_kernel<<<B * T, 256, 0, st>>>(w.magc, w.fbT, B, T, Fu, g, w.X, w.fs);
    FSN_CHECK_LAUNCH("imp_section_input_kernel");
    if ((rc = clip_reduce_only_launch(w.fs, B, T, w.sums, st))) return rc;
    if ((rc = norm_scales_launch(w.sums, w.sums, B, (float)g.N * g.W * T, 1.f, w.invs, nullptr, st, eps))) return rc;
    const fsn_seq_weights& sw = wt->sb[s];
    if (d->precision == FSN_PREC_TF32_TC) {
      // layer by layer over all steps: hoisted input projection + per-step recurrent GEMM on tcgen05 (tf32), fused cell
      LayerSave l1{w.tc.G, w.tc.C, w.tc_h1};
      {  // scale X by the section norm in place (the tensor-core GEMM reads plain fp32 rows)
        const size_t n = (size_t)T * R * g.W;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 148 * 16) blocks = 148 * 16;
        imp_scale_rows_kernel<<<blocks, 256, 0, st>>>(w.X, w.invs, n, (size_t)R * g.W, (size_t)g.N * g.W, B);
        FSN_CHECK_LAUNCH("imp_scale_rows_kernel");
      }
      if ((rc = layer_forward_save_tc(&sw, 0, w.X, R, g.W, Hs, T, w.tc, w.tc_rec, st))) return rc;
      if ((rc = layer_forward_save_tc(&sw, 1, w.tc.H, R, Hs, Hs, T, l1, w.tc_rec, st))) return rc;
      for (int t = 0; t < T; ++t) {
        const size_t warps = (size_t)R * 2 * g.cs;
        imp_fc_step_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(w.tc_h1 + (size_t)t * R * Hs, R, Hs, sw.fc_w, sw.fc_b,
                                                                    g.cs, g.N, g.lo, d->sb_activation, crm, F, T, t);
        FSN_CHECK_LAUNCH("imp_fc_step_kernel");
      }
      continue;
    }
    for (int t = 0; t < T; ++t) {
      StepParams p;
      memset(&p, 0, sizeof(p));
      p.R = R; p.first = (t == 0);
      p.K0 = g.W; p.H = Hs;
      p.w_ih = sw.w_ih[0]; p.w_hh = sw.w_hh[0]; p.b_ih = sw.b_ih[0]; p.b_hh = sw.b_hh[0];
      p.h_prev = w.h0[(t + 1) & 1]; p.h_prev_stride = Hs;
      p.h_out = w.h0[t & 1]; p.h_out_stride = Hs;
      p.c = w.c0;
      p.x0 = w.X + (size_t)t * R * g.W; p.x0_row_stride = g.W; p.row_scale = w.invs; p.row_scale_div = g.N;
      if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
      p.K0 = Hs;
      p.w_ih = sw.w_ih[1]; p.w_hh = sw.w_hh[1]; p.b_ih = sw.b_ih[1]; p.b_hh = sw.b_hh[1];
      p.x0 = w.h0[t & 1]; p.x0_row_stride = Hs; p.row_scale = nullptr; p.row_scale_div = 0;
      p.h_prev = w.h1[(t + 1) & 1]; p.h_prev_stride = Hs;
      p.h_out = w.h1[t & 1]; p.h_out_stride = Hs;
      p.c = w.c1;
      if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
      const size_t warps = (size_t)R * 2 * g.cs;
      imp_fc_step_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, st>>>(w.h1[t & 1], R, Hs, sw.fc_w, sw.fc_b, g.cs, g.N, g.lo,
                                                                  d->sb_activation, crm, F, T, t);
      FSN_CHECK_LAUNCH("imp_fc_step_kernel");
    }
  }
  // element-wise mask on (re, im) + iSTFT (575-589)
  return istft_launch(w.real, w.imag, 1, crm, B, T, d->n_fft, d->hop_length, d->win_length, L, enhanced, st, 2);
}
