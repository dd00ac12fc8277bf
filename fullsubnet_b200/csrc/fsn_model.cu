// Host-side orchestration of Model.forward and the fused wav->wav enhancement call, plus the
// error / version entry points of the C ABI (include/fsn_b200.h).
//
// Reference: recipes/dns_interspeech_2020/fullsubnet/model.py:72-136 (Model.forward),
//            recipes/dns_interspeech_2020/inferencer.py:130-145 (full_band_crm_mask).
#include <stdlib.h>
#include <string.h>

#include "fsn_internal.cuh"

namespace fsn {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
static thread_local int g_err_code = 0;
int& last_error_code() { return g_err_code; }
int64_t& launch_counter() { return g_launches; }
static int64_t g_total_launches = 0;
int64_t& total_launch_counter() { return g_total_launches; }

// opt-in stage timing (bench.py): 5 events bracket the 4 stages
static thread_local bool g_prof = false;
static thread_local cudaEvent_t g_ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
static thread_local bool g_ev_valid[5] = {false, false, false, false, false};
static void prof_mark(int i, cudaStream_t st) {
  if (!g_prof) return;
  if (!g_ev[i] && cudaEventCreate(&g_ev[i]) != cudaSuccess) return;
  g_ev_valid[i] = (cudaEventRecord(g_ev[i], st) == cudaSuccess);
}
static void prof_reset() { for (int i = 0; i < 5; ++i) g_ev_valid[i] = false; }

struct Carver {
  char* base; size_t off;
  explicit Carver(void* p) : base((char*)p), off(0) {}
  template <class T> T* take(size_t n) {
    T* r = base ? (T*)(base + off) : nullptr;
    off = align_up(off + n * sizeof(T), 256);
    return r;
  }
};

struct ModelWs {
  float *magT, *fbT, *inv1, *inv2;
  float2 *fs, *sums_mag, *sums_fb;
  float *fb_h0[2], *fb_c0, *fb_c1, *fb_h1all, *fb_pp;
  float *cum1, *cum2;  // cumulative norm: per-(step, clip) and per-(step, unit) scales
  unsigned int* fb_barrier;
  float *sb_h0[2], *sb_h1[2], *sb_c0, *sb_c1;
  // tensor-core full-band path (fb_tc_forward): split operands, hoisted projection, layer-0 output, scratch
  LstmTcWs tc;
  float* tc_h0all;
  size_t bytes;
};

// full-band stack on the tensor cores?  (tensor-core precisions, offline norm, enough clips to fill an MMA tile)
static bool fb_tc_enabled(const fsn_model_desc* d, int B) {
  // every batch size takes the same path, so a clip's result does not depend on the batch it is enhanced in
  static const int min_b = getenv("FSN_FB_TC_MIN_B") ? atoi(getenv("FSN_FB_TC_MIN_B")) : 1;
  if (d->precision != FSN_PREC_F16_TC && d->precision != FSN_PREC_F16X3_TC) return false;
  if (d->cell_type != FSN_CELL_LSTM) return false;
  if (B < min_b) return false;
  return lstm_rec_tc_supported(d->fb_hidden, d->precision == FSN_PREC_F16X3_TC);
}

int make_dims(const fsn_model_desc* d, int B, int T, Dims& m) {
  FSN_REQUIRE(d && d->num_freqs > 1 && d->fb_hidden > 0 && d->sb_hidden > 0 && d->look_ahead >= 0, FSN_ERR_SHAPE,
              "model: bad descriptor");
  FSN_REQUIRE(B > 0 && T > 0, FSN_ERR_SHAPE, "model: empty input (B=%d, T=%d)", B, T);
  FSN_REQUIRE(d->norm_type == FSN_NORM_OFFLINE_LAPLACE || d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE,
              FSN_ERR_UNSUPPORTED, "You must set up a type of Norm. (offline_laplace_norm / cumulative_laplace_norm are built)");
  FSN_REQUIRE(d->cell_type == FSN_CELL_LSTM || (d->cell_type == FSN_CELL_GRU && d->precision == FSN_PREC_FP32),
              FSN_ERR_UNSUPPORTED, "model: sequence_model must be LSTM, or GRU on the fp32 kernels (precision fp32)");
  FSN_REQUIRE(d->sb_num_neighbors >= 0 && d->fb_num_neighbors >= 0 && d->sb_num_neighbors < d->num_freqs &&
                  d->fb_num_neighbors < d->num_freqs,
              FSN_ERR_SHAPE, "model: reflect padding needs num_neighbors < num_freqs");
  m.B = B; m.T = T; m.Tp = T + d->look_ahead; m.F = d->num_freqs;
  m.G = (B > 1 && d->num_groups_in_drop_band > 1) ? d->num_groups_in_drop_band : 1;
  if (B > 1)  // model.py:114-117 -> feature.py:317-319 (asserted even when G == 1)
    FSN_REQUIRE(B > d->num_groups_in_drop_band, FSN_ERR_SHAPE,
                "Batch size = %d, num_groups = %d. The batch size should larger than the num_groups.", B,
                d->num_groups_in_drop_band);
  m.Fsub = (m.G > 1) ? m.F / m.G : m.F;
  FSN_REQUIRE(m.Fsub > 0, FSN_ERR_SHAPE, "model: num_freqs < num_groups");
  m.R = B * m.Fsub;
  m.Ksb = (2 * d->sb_num_neighbors + 1) + (2 * d->fb_num_neighbors + 1);
  return FSN_OK;
}

static void carve_model(const fsn_model_desc* d, const Dims& m, void* base, ModelWs& w) {
  Carver c(base);
  const size_t BTF = (size_t)m.B * m.Tp * m.F;
  w.magT = c.take<float>(BTF);
  w.fbT = c.take<float>(BTF);
  w.fs = c.take<float2>((size_t)m.B * m.Tp);
  w.sums_mag = c.take<float2>(m.B);
  w.sums_fb = c.take<float2>(m.B);
  w.inv1 = c.take<float>(m.B);
  w.inv2 = c.take<float>(m.B);
  const size_t BH = (size_t)m.B * d->fb_hidden;
  w.fb_h0[0] = c.take<float>(BH);
  w.fb_h0[1] = c.take<float>(BH);
  w.fb_c0 = c.take<float>(BH);
  w.fb_c1 = c.take<float>(BH);
  w.fb_h1all = c.take<float>(BH * m.Tp);
  w.fb_pp = c.take<float>((size_t)2 * 256 * d->fb_hidden);  // h0 ping-pong of the persistent kernel
  w.fb_barrier = c.take<unsigned int>(64);
  if (d->precision == FSN_PREC_FP32) {
    const size_t RH = (size_t)m.R * d->sb_hidden;
    for (int i = 0; i < 2; ++i) { w.sb_h0[i] = c.take<float>(RH); w.sb_h1[i] = c.take<float>(RH); }
    w.sb_c0 = c.take<float>(RH);
    w.sb_c1 = c.take<float>(RH);
  }
  memset(&w.tc, 0, sizeof(w.tc));
  w.tc_h0all = nullptr;
  if (fb_tc_enabled(d, m.B)) {
    const int Hf = d->fb_hidden;
    const size_t rows = (size_t)m.B * m.Tp;
    lstm_tc_carve(c.base, c.off, rows, m.F > Hf ? m.F : Hf, Hf, d->precision == FSN_PREC_F16X3_TC, w.tc);
    w.tc_h0all = c.take<float>(rows * Hf);
  }
  w.cum1 = w.cum2 = nullptr;
  if (d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE) {
    w.cum1 = c.take<float>((size_t)m.Tp * m.B);
    w.cum2 = c.take<float>((size_t)m.Tp * m.R);
  }
  w.bytes = c.off;
}

// Full-band stack on the tensor cores (model.py:92-95; sequence_model.py:106-125): per layer the input projection of
// all steps as one tf32 GEMM (three passes on hi/lo splits when x3) and the recurrence in the persistent tcgen05
// kernel; Linear(Hf -> F) + activation as the same GEMM + a bias/activation pass.  x3 keeps the fp32 error class.
static int fb_tc_forward(const fsn_model_desc* d, const fsn_seq_weights* fb, const Dims& m, const ModelWs& w, cudaStream_t st) {
  const int F = m.F, Tp = m.Tp, B = m.B, Hf = d->fb_hidden;
  const bool x3 = d->precision == FSN_PREC_F16X3_TC;
  const bool cum = d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE;
  int rc;
  fsn_lstm_layer L0{fb->w_ih[0], fb->w_hh[0], fb->b_ih[0], fb->b_hh[0]}, L1{fb->w_ih[1], fb->w_hh[1], fb->b_ih[1], fb->b_hh[1]};
  // layer 0: x = magT * 1/(mu + 1e-5) of the clip (model.py:92); cumulative norm: the scale of (clip, step) from the
  // time-major table cum1[t*B + b] (base_model.py:220-251)
  if ((rc = lstm_layer_tc(L0, w.magT, (size_t)F, F, cum ? w.cum1 : w.inv1, Tp, cum ? B : 0, B, Tp, Hf, x3, w.tc, w.tc_h0all, st)))
    return rc;
  if ((rc = lstm_layer_tc(L1, w.tc_h0all, (size_t)Hf, Hf, nullptr, 1, 0, B, Tp, Hf, x3, w.tc, w.fb_h1all, st))) return rc;
  // Linear(Hf -> F) + activation (sequence_model.py:119-123) -> fbT [B, Tp, F]
  return linear_tc(w.fb_h1all, (size_t)Hf, Hf, fb->fc_w, fb->fc_b, F, d->fb_activation, w.fbT, (size_t)F, (size_t)B * Tp, x3,
                   w.tc, st);
}

// everything after the time-major magnitude exists: norms, full-band stack, sub-band stack
static int model_core(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                      const void* sb_packed, const Dims& m, const ModelWs& w, float* crm, cudaStream_t st) {
  int rc;
  const int F = m.F, Tp = m.Tp, B = m.B, Hf = d->fb_hidden, Hs = d->sb_hidden;
  // per-clip statistics of the look-ahead-padded magnitude (model.py:92, :111)
  if ((rc = clip_stats_launch(w.magT, B, Tp, F, d->sb_num_neighbors, w.fs, w.sums_mag, st))) return rc;
  if ((rc = norm_scales_launch(w.sums_mag, w.sums_mag, B, (float)F * Tp, 1.f, w.inv1, nullptr, st))) return rc;
  const bool cum = d->norm_type == FSN_NORM_CUMULATIVE_LAPLACE;
  const float cum_eps = 1.1920928955078125e-07f;  // audio_zen/constant.py:9 (np.finfo(np.float32).eps)
  if (cum && (rc = cum_clip_scale_launch(w.fs, B, Tp, F, cum_eps, w.cum1, st))) return rc;

  // ---- full-band stack (model.py:92-95): 2-layer LSTM(F -> Hf -> Hf), rows = clips
  static const bool fb_stepwise = getenv("FSN_FB_STEPWISE") != nullptr;  // debug: force the per-step kernels
  const bool fb_tc = !fb_stepwise && fb_tc_enabled(d, B);
  if (fb_tc) {
    if ((rc = fb_tc_forward(d, fb, m, w, st))) return rc;
  } else if (!fb_stepwise && !cum && d->cell_type == FSN_CELL_LSTM && fb_persistent_supported(F, Hf, Hf)) {
    // one persistent cooperative kernel per chunk of <= 256 clips: weights resident in shared memory,
    // layer wavefront, one grid barrier per time step
    for (int b0 = 0; b0 < B; b0 += 256) {
      const int nb = (B - b0 < 256) ? B - b0 : 256;
      if ((rc = fb_persistent_launch(fb, w.magT + (size_t)b0 * Tp * F, w.inv1 + b0, w.fb_pp,
                                     w.fb_h1all + (size_t)b0 * Tp * Hf, w.fb_barrier, nb, F, Hf, Hf, Tp, st)))
        return rc;
    }
  } else {
  for (int t = 0; t < Tp; ++t) {
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.R = B; p.H = Hf; p.first = (t == 0); p.gru = d->cell_type == FSN_CELL_GRU;
    // layer 0: x_t = magT[b,t,:] * inv1[b]
    p.K0 = F;
    p.w_ih = fb->w_ih[0]; p.w_hh = fb->w_hh[0]; p.b_ih = fb->b_ih[0]; p.b_hh = fb->b_hh[0];
    p.h_prev = w.fb_h0[(t + 1) & 1]; p.h_prev_stride = Hf;
    p.h_out = w.fb_h0[t & 1]; p.h_out_stride = Hf;
    p.c = w.fb_c0;
    p.x0 = w.magT + (size_t)t * F; p.x0_row_stride = (size_t)Tp * F;
    p.row_scale = cum ? w.cum1 + (size_t)t * B : w.inv1;  // cumulative: scale of (step t, clip b)
    if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
    // layer 1: x_t = h0_t, output kept for every t (input of the Linear layer)
    p.K0 = Hf;
    p.w_ih = fb->w_ih[1]; p.w_hh = fb->w_hh[1]; p.b_ih = fb->b_ih[1]; p.b_hh = fb->b_hh[1];
    p.x0 = w.fb_h0[t & 1]; p.x0_row_stride = Hf; p.row_scale = nullptr;
    p.h_prev = w.fb_h1all + (size_t)(t > 0 ? t - 1 : 0) * Hf; p.h_prev_stride = (size_t)Tp * Hf;
    p.h_out = w.fb_h1all + (size_t)t * Hf; p.h_out_stride = (size_t)Tp * Hf;
    p.c = w.fb_c1;
    if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
  }
  }
  // Linear(Hf -> F) + activation over all (b,t): fbT[b,t,f]  (sequence_model.py:119-123)
  if (!fb_tc && (rc = fc_gemm_launch(w.fb_h1all, fb->fc_w, fb->fc_b, w.fbT, B * Tp, Hf, F, d->fb_activation, st))) return rc;

  // ---- second norm (model.py:110-111) in closed form: never materialise [B,F,Ksb,T']
  if ((rc = clip_stats_launch(w.fbT, B, Tp, F, d->fb_num_neighbors, w.fs, w.sums_fb, st))) return rc;
  if ((rc = norm_scales_launch(w.sums_mag, w.sums_fb, B, 1.f, (float)F * m.Ksb * Tp, nullptr, w.inv2, st)))
    return rc;

  prof_mark(2, st);
  RowMap map{B, F, m.Fsub, m.G};
  if (cum && (rc = cum_unit_scale_launch(w.magT, w.fbT, map, m.R, Tp, d->sb_num_neighbors, d->fb_num_neighbors, cum_eps,
                                         w.cum2, st)))
    return rc;
  if (d->precision == FSN_PREC_F16_TC || d->precision == FSN_PREC_F16X3_TC) {
    FSN_REQUIRE(sb_packed, FSN_ERR_SHAPE, "model: the tensor-core precisions need packed sub-band weights");
    SbTcArgs a;
    memset(&a, 0, sizeof(a));
    a.packed = sb_packed; a.magT = w.magT; a.fbT = w.fbT; a.inv2 = w.inv2; a.crm = crm;
    a.B = B; a.F = F; a.Tp = Tp; a.la = d->look_ahead; a.Ns = d->sb_num_neighbors; a.Nf = d->fb_num_neighbors;
    a.H = Hs; a.act = d->sb_activation; a.map = map; a.pair = sb_tc2_supported(d); a.x3 = d->precision == FSN_PREC_F16X3_TC;
    a.unit_scale = cum ? w.cum2 : nullptr;
    a.quad = sb_tc4_supported(d);
    rc = sb_tc_forward(a, st);
    prof_mark(3, st);
    return rc;
  }

  // ---- sub-band stack, fp32 path (model.py:121-135): rows = (clip, frequency) units
  for (int t = 0; t < Tp; ++t) {
    StepParams p;
    memset(&p, 0, sizeof(p));
    p.R = m.R; p.H = Hs; p.first = (t == 0); p.gru = d->cell_type == FSN_CELL_GRU;
    p.K0 = m.Ksb;
    p.w_ih = sb->w_ih[0]; p.w_hh = sb->w_hh[0]; p.b_ih = sb->b_ih[0]; p.b_hh = sb->b_hh[0];
    p.h_prev = w.sb_h0[(t + 1) & 1]; p.h_prev_stride = Hs;
    p.h_out = w.sb_h0[t & 1]; p.h_out_stride = Hs;
    p.c = w.sb_c0;
    p.magT = w.magT; p.fbT = w.fbT; p.inv2 = w.inv2;
    p.unit_scale = cum ? w.cum2 + (size_t)t * m.R : nullptr;
    p.F = F; p.Tp = Tp; p.t = t; p.Ns = d->sb_num_neighbors; p.Nf = d->fb_num_neighbors; p.map = map;
    if ((rc = lstm_step_launch(p, SEG0_GATHER, st))) return rc;
    p.K0 = Hs;
    p.w_ih = sb->w_ih[1]; p.w_hh = sb->w_hh[1]; p.b_ih = sb->b_ih[1]; p.b_hh = sb->b_hh[1];
    p.x0 = w.sb_h0[t & 1]; p.x0_row_stride = Hs; p.row_scale = nullptr;
    p.h_prev = w.sb_h1[(t + 1) & 1]; p.h_prev_stride = Hs;
    p.h_out = w.sb_h1[t & 1]; p.h_out_stride = Hs;
    p.c = w.sb_c1;
    if ((rc = lstm_step_launch(p, SEG0_DENSE, st))) return rc;
    if (t >= d->look_ahead)
      if ((rc = sb_fc_step_launch(w.sb_h1[t & 1], m.R, Hs, sb->fc_w, sb->fc_b, 2, d->sb_activation, crm, m.Fsub,
                                  m.T, t - d->look_ahead, st)))
        return rc;
  }
  prof_mark(3, st);
  return FSN_OK;
}

}  // namespace fsn

using namespace fsn;

extern "C" int fsn_version(void) { return 100; }
extern "C" const char* fsn_last_error(void) { return g_err; }
extern "C" int fsn_last_error_code(void) { return g_err_code; }
extern "C" int64_t fsn_last_launch_count(void) { return g_launches; }
extern "C" int64_t fsn_total_launch_count(void) { return g_total_launches; }
extern "C" int fsn_set_profiling(int enable) { g_prof = enable != 0; return FSN_OK; }
extern "C" float fsn_last_stage_ms(int stage) {
  if (stage < 0 || stage > 3 || !g_ev_valid[stage] || !g_ev_valid[stage + 1]) return -1.0f;
  float ms = -1.0f;
  if (cudaEventElapsedTime(&ms, g_ev[stage], g_ev[stage + 1]) != cudaSuccess) return -1.0f;
  return ms;
}
// host-side views of the index helpers the kernels share (tests/test_cpu_host.py checks them against the oracle)
extern "C" int fsn_debug_row_to_unit(int B, int F, int G, int r, int* b, int* f) {
  const int g = (B > 1 && G > 1) ? G : 1;
  RowMap m{B, F, g > 1 ? F / g : F, g};
  if (r < 0 || r >= B * m.Fsub) return FSN_ERR_SHAPE;
  row_to_unit(m, r, *b, *f);
  return FSN_OK;
}
extern "C" int fsn_debug_unit_to_row(int B, int F, int G, int b, int f) {
  const int g = (B > 1 && G > 1) ? G : 1;
  RowMap m{B, F, g > 1 ? F / g : F, g};
  return unit_to_row(m, b, f);
}
extern "C" int fsn_debug_reflect_count(int r, int F, int N) { return reflect_count(r, F, N); }

extern "C" int fsn_built_arch(void) {
#ifdef FSN_BUILT_ARCH
  return FSN_BUILT_ARCH;
#else
  return 0;
#endif
}

extern "C" size_t fsn_model_workspace_bytes(const fsn_model_desc* d, int B, int T) {
  Dims m;
  if (make_dims(d, B, T, m)) return 0;
  ModelWs w;
  carve_model(d, m, nullptr, w);
  return w.bytes;
}

extern "C" size_t fsn_sb_packed_bytes(const fsn_model_desc* d) { return sb_tc_packed_bytes(d); }

extern "C" int fsn_pack_sb_weights(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed,
                                   fsn_stream_t stream) {
  return sb_tc_pack(d, sb, packed, (cudaStream_t)stream);
}

extern "C" int fsn_model_forward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                                 const void* sb_packed, const float* noisy_mag, int B, int T, float* crm,
                                 void* workspace, size_t workspace_bytes, fsn_stream_t stream) {
  g_launches = 0;
  Dims m;
  int rc = make_dims(d, B, T, m);
  if (rc) return rc;
  FSN_REQUIRE(d->precision == FSN_PREC_FP32 || sb_tc_supported(d), FSN_ERR_UNSUPPORTED,
              "FSN_PREC_F16_TC needs sb_hidden %% 128 == 0 (FSN_PREC_F16X3_TC: sb_hidden = 384) and sub-band input width <= 32");
  ModelWs w;
  carve_model(d, m, workspace, w);
  FSN_REQUIRE(workspace && workspace_bytes >= w.bytes, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, w.bytes);
  cudaStream_t st = (cudaStream_t)stream;
  prof_reset();
  prof_mark(0, st);
  if ((rc = transpose_mag_launch(noisy_mag, w.magT, B, m.F, T, m.Tp, st))) return rc;
  prof_mark(1, st);
  return model_core(d, fb, sb, sb_packed, m, w, crm, st);
}

struct EnhanceWs {
  float *real, *imag, *crm;
  unsigned int* peak;
  void* model;
  size_t bytes;
};

static int carve_enhance(const fsn_model_desc* d, int B, int L, int n_fft, int hop, void* base, EnhanceWs& e,
                         Dims& m) {
  FSN_REQUIRE(hop > 0 && n_fft > 0, FSN_ERR_SHAPE, "enhance: bad n_fft/hop");
  const int T = 1 + L / hop;
  int rc = make_dims(d, B, T, m);
  if (rc) return rc;
  FSN_REQUIRE(n_fft / 2 + 1 == d->num_freqs, FSN_ERR_SHAPE, "enhance: n_fft/2+1 = %d != num_freqs = %d",
              n_fft / 2 + 1, d->num_freqs);
  Carver c(base);
  const size_t BFT = (size_t)B * m.F * T;
  e.real = c.take<float>(BFT);
  e.imag = c.take<float>(BFT);
  e.crm = c.take<float>(2 * BFT);
  e.peak = c.take<unsigned int>(B);
  ModelWs w;
  carve_model(d, m, nullptr, w);
  e.model = base ? (char*)base + c.off : nullptr;
  e.bytes = c.off + w.bytes;
  return FSN_OK;
}

extern "C" size_t fsn_enhance_workspace_bytes(const fsn_model_desc* d, int B, int L, int n_fft, int hop) {
  EnhanceWs e;
  Dims m;
  fsn_model_desc dd = *d;
  dd.num_groups_in_drop_band = 1;
  if (carve_enhance(&dd, B, L, n_fft, hop, nullptr, e, m)) return 0;
  return e.bytes;
}

static int enhance_impl(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                        const void* sb_packed, const float* wav, int B, int L, int n_fft, int hop, int win_length,
                        float* enhanced, float* crm_out, int16_t* pcm, float pcm_gain, void* workspace,
                        size_t workspace_bytes, fsn_stream_t stream) {
  g_launches = 0;
  // batched inference == loop of B=1 calls of the reference inferencer: drop_band off
  // (audio_zen/inferencer/base_inferencer.py:78,173)
  fsn_model_desc dd = *d;
  dd.num_groups_in_drop_band = 1;
  EnhanceWs e;
  Dims m;
  int rc = carve_enhance(&dd, B, L, n_fft, hop, workspace, e, m);
  if (rc) return rc;
  FSN_REQUIRE(dd.precision == FSN_PREC_FP32 || sb_tc_supported(&dd), FSN_ERR_UNSUPPORTED,
              "FSN_PREC_F16_TC needs sb_hidden %% 128 == 0 (FSN_PREC_F16X3_TC: sb_hidden = 384) and sub-band input width <= 32");
  FSN_REQUIRE(workspace && workspace_bytes >= e.bytes, FSN_ERR_WORKSPACE, "workspace too small: %zu < %zu",
              workspace_bytes, e.bytes);
  ModelWs w;
  carve_model(&dd, m, e.model, w);
  cudaStream_t st = (cudaStream_t)stream;
  float* crm = crm_out ? crm_out : e.crm;
  prof_reset();
  prof_mark(0, st);
  if ((rc = stft_launch(wav, B, L, n_fft, hop, win_length, nullptr, nullptr, e.real, e.imag, w.magT, m.Tp, st)))
    return rc;
  prof_mark(1, st);
  if ((rc = model_core(&dd, fb, sb, sb_packed, m, w, crm, st))) return rc;
  rc = istft_launch(e.real, e.imag, 1, crm, B, m.T, n_fft, hop, win_length, L, enhanced, st, 1, pcm ? e.peak : nullptr);
  if (!rc && pcm) rc = scale_int16_launch(enhanced, e.peak, B, L, pcm_gain, pcm, st);
  prof_mark(4, st);
  return rc;
}

extern "C" int fsn_enhance(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                           const void* sb_packed, const float* wav, int B, int L, int n_fft, int hop,
                           int win_length, float* enhanced, float* crm_out, void* workspace,
                           size_t workspace_bytes, fsn_stream_t stream) {
  return enhance_impl(d, fb, sb, sb_packed, wav, B, L, n_fft, hop, win_length, enhanced, crm_out, nullptr, 0.f, workspace,
                      workspace_bytes, stream);
}

extern "C" int fsn_enhance_pcm(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                               const void* sb_packed, const float* wav, int B, int L, int n_fft, int hop,
                               int win_length, float* enhanced, int16_t* pcm, float gain, void* workspace,
                               size_t workspace_bytes, fsn_stream_t stream) {
  FSN_REQUIRE(pcm && enhanced, FSN_ERR_SHAPE, "enhance_pcm: output buffers missing");
  return enhance_impl(d, fb, sb, sb_packed, wav, B, L, n_fft, hop, win_length, enhanced, nullptr, pcm, gain, workspace,
                      workspace_bytes, stream);
}
