// Internal (non-ABI) declarations shared by the translation units of libfsn_b200.
#pragma once
#include <cuda_fp16.h>

#include "fsn_common.cuh"

namespace fsn {

// Sub-band row -> (clip, frequency) map.  Row r = b' * Fsub + f' of the sub-band batch.
// G <= 1: identity (Fsub = F).  G > 1: drop_band (audio_zen/acoustics/feature.py:332-345):
// output clip b' of group g is input clip g + G*i, frequency f' is input bin g + G*f'.
struct RowMap {
  int B, F, Fsub, G;
};

__host__ __device__ inline void row_to_unit(const RowMap& m, int r, int& b, int& f) {
  int bq = r / m.Fsub;
  const int fq = r - bq * m.Fsub;
  if (m.G <= 1) { b = bq; f = fq; return; }
  int g = 0;
  for (; g < m.G; ++g) {
    const int cnt = (m.B - g + m.G - 1) / m.G;
    if (bq < cnt) break;
    bq -= cnt;
  }
  b = g + m.G * bq;
  f = g + m.G * fq;
}

enum { SEG0_DENSE = 0, SEG0_GATHER = 1 };

struct StepParams {
  int R, K0, H, first;
  int gru;  // 0: LSTM (4 gates i,f,g,o); 1: GRU (weights [3H,.] r,z,n; h_prev is read in the update; no cell state)
  const float* w_ih; const float* w_hh; const float* b_ih; const float* b_hh;
  const float* h_prev; size_t h_prev_stride;
  float* h_out; size_t h_out_stride;
  float* c;
  // training only (both nullable): previous cell state read from c_in instead of c; post-activation gates
  // (i,f,g,o) of this step stored at save_gates[row*4H + g*H + u]
  const float* c_in; float* save_gates;
  // SEG0_DENSE: x = x0[row * x0_row_stride + k] * (row_scale ? row_scale[row] : 1)
  const float* x0; size_t x0_row_stride; const float* row_scale;
  int row_scale_div;  // scale index = row / row_scale_div (0 or 1: per row)
  // SEG0_GATHER: sub-band unit of row r at frame t (base_model.py:13-46 + model.py:98-111)
  const float* magT; const float* fbT; const float* inv2;
  const float* unit_scale;  // nullable: per-row scale of this step (cumulative norm) instead of inv2[clip]
  int F, Tp, t, Ns, Nf;
  RowMap map;
};

int lstm_step_launch(const StepParams& p, int mode, cudaStream_t st);
// cumulative_laplace_norm (base_model.py:220-251): scale1T[t*B+b] from the frame sums fs[b*Tp+t].x, and
// scaleT[t*R+r] of every sub-band unit (running mean over its K rows and the frames so far)
int cum_clip_scale_launch(const float2* fs, int B, int Tp, int F, float eps, float* scale1T, cudaStream_t st);
int cum_unit_scale_launch(const float* magT, const float* fbT, RowMap map, int R, int Tp, int Ns, int Nf, float eps,
                          float* scaleT, cudaStream_t st, bool time_major = false);

// tf32 tcgen05 GEMM (fsn_tgemm.cu): C[M,N] (+)= A[M,K] B[N,K]^T, fp32 row-major operands with 16-byte aligned rows
bool tgemm_supported(const float* A, size_t lda, const float* Bm, size_t ldb, int K);
int tgemm_launch(const float* A, size_t lda, const float* Bm, size_t ldb, float* C, size_t ldc, int M, int N, int K,
                 bool accumulate, float* scratch, size_t scratch_floats, cudaStream_t st);

// weight-gradient GEMMs C[M,N] (+)= A^T B over a long K with block-tiled K-major operand copies (fsn_tgemm.cu)
size_t tgemm_blocked_floats(size_t K, int M);
int transpose_blocked_launch(const float* in, size_t K, int M, size_t ld, float* out, cudaStream_t st, float* colsum_part,
                             int max_slabs, int* slabs);
bool tgemm_blocked_enabled();
int tgemm_blocked_launch(const float* Ablk, int nkb_a, int a_kb0, const float* Bblk, int nkb_b, int b_kb0, float* C, size_t ldc,
                         int M, int N, int K, bool accumulate, float* scratch, size_t scratch_floats, cudaStream_t st);

bool lstm_fwd_step_supported(const float* Hbuf, const float* w_hh, int H);
bool lstm_fwd_step_folds_input(const float* X, const float* w_ih, int K0);
// optional fp16 MMA operands of the step kernel: previous hidden state, weights, folded layer input; H16_out receives h_t
struct LstmStepHalf {
  const __half *Hprev16, *w_hh16, *Xt16, *w_ih16;
  __half* H16_out;
};
bool lstm_fwd_step_half_enabled(int H);
int to_half_launch(const float* in, size_t n, __half* out, cudaStream_t st);
int lstm_fwd_step_launch(const float* Hprev, const float* w_hh, const float* Xt, const float* w_ih, int K0, float* Gt,
                         const float* b_ih, const float* b_hh, const float* C_prev, float* C_out, float* H_out, int R, int H,
                         cudaStream_t st, const LstmStepHalf* h = nullptr);

// one LSTM layer over all steps on the tf32 tensor-core path (fsn_train.cu): input projection of all steps hoisted
// into one GEMM, then per step the recurrent GEMM into `rec` [R,4H] and the fused cell kernel.  G [Tp,R,4H]
// (post-activation gates), C, H [Tp,R,H] receive every step.  X [Tp,R,K0] contiguous.
struct LayerSave { float *G, *C, *H; };
// fp16 side buffers of one layer (all nullable): H16 [Tp,R,H] copy of the hidden states (written by the step kernel, the
// next layer's X16), X16 [Tp,R,K0] copy of the layer input, w16: 4H*(H+K0) halfs for the weight copies
struct LayerHalf { __half* H16; const __half* X16; __half* w16; };
int layer_forward_save_tc(const fsn_seq_weights* w, int l, const float* X, int R, int K0, int H, int Tp,
                          const LayerSave& s, float* rec, cudaStream_t st, float* splitk = nullptr, size_t splitk_floats = 0,
                          const LayerHalf* half = nullptr);

// shapes of one Model.forward call (fsn_model.cu)
struct Dims {
  int B, T, Tp, F, Fsub, G, R, Ksb;
};
int make_dims(const fsn_model_desc* d, int B, int T, Dims& m);

// (clip, frequency) -> sub-band row, or -1 when drop_band removed the unit (inverse of row_to_unit)
__host__ __device__ inline int unit_to_row(const RowMap& m, int b, int f) {
  if (m.G <= 1) return b * m.Fsub + f;
  const int g = b % m.G;
  if (f % m.G != g || f / m.G >= m.Fsub) return -1;
  int off = 0;
  for (int gg = 0; gg < g; ++gg) off += (m.B - gg + m.G - 1) / m.G;
  return (off + b / m.G) * m.Fsub + f / m.G;
}
int fc_gemm_launch(const float* A, const float* W, const float* bias, float* out, int M, int K, int O, int act,
                   cudaStream_t st, bool w_kmajor = false);
// out[row*row_stride + o*o_stride] = act(h[row,:] . W[o,:] + b[o]), one warp per row (small O)
int rows_fc_launch(const float* h, int R, int H, const float* W, const float* bias, int O, int act, float* out,
                   size_t row_stride, size_t o_stride, cudaStream_t st);
int sb_fc_step_launch(const float* h, int R, int H, const float* W, const float* bias, int O, int act, float* crm,
                      int Fsub, int T_out, int t_out, cudaStream_t st);
int sb_fc_steps_launch(const float* h, int R, int H, int steps, const float* W, const float* bias, int O, int act, float* crm,
                       int Fsub, int T_out, int t_out0, cudaStream_t st);
int transpose_mag_launch(const float* in, float* out, int B, int F, int T, int T_pad, cudaStream_t st);
int clip_stats_launch(const float* x, int B, int T_pad, int F, int N, float2* fs, float2* sums, cudaStream_t st);
int clip_reduce_only_launch(const float2* fs, int B, int T_pad, float2* sums, cudaStream_t st);
int norm_scales_launch(const float2* mag_sums, const float2* fb_sums, int B, float cnt1, float cnt2, float* inv1,
                       float* inv2, cudaStream_t st, float eps = 1e-5f);

int stft_launch(const float* wav, int B, int L, int n_fft, int hop, int win_length, float* mag, float* phase,
                float* real, float* imag, float* magT, int T_pad, cudaStream_t st);
// mask_mode: 1 = decompress_cIRM + complex product (fullsubnet), 2 = element-wise re*crm0, im*crm1 (improved_fullsubnet)
int istft_launch(const float* real, const float* imag, int cstride, const float* crm, int B, int T, int n_fft,
                 int hop, int win_length, int length, float* wav, cudaStream_t st, int mask_mode = 1,
                 unsigned int* peak_bits = nullptr);
// peak_bits (optional, [B]): max|wav| per clip as float bits, reduced in the iSTFT epilogue; scale_int16_launch turns
// it into the int16 scaling of the reference host loop (audio_zen/inferencer/base_inferencer.py:181-182)
int scale_int16_launch(const float* wav, const unsigned int* peak_bits, int B, int L, float gain, int16_t* out, cudaStream_t st);

// persistent cooperative full-band LSTM (fsn_fullband.cu)
bool fb_persistent_supported(int F, int H0, int H1);
int fb_persistent_launch(const fsn_seq_weights* w, const float* x_chunk, const float* inv1_chunk, float* h0buf,
                         float* h1all_chunk, unsigned int* barrier, int nb, int F, int H0, int H1, int Tp,
                         cudaStream_t st);

// tensor-core LSTM layer for a small batch of sequences (fsn_lstm_rec_tc.cu): hoisted input projection on the tf32
// GEMM (x3: three passes on tf32 hi/lo splits) + persistent cooperative tcgen05 recurrence (x3: fp16 hi/lo splits)
bool lstm_rec_tc_supported(int H, bool x3);
int lstm_rec_tc_rows_per_launch(int H);
size_t lstm_rec_tc_scratch_bytes(int H, bool x3);
int lstm_rec_tc_launch(const float* w_hh, const float* b_ih, const float* b_hh, const float* P, size_t p_row, size_t p_t,
                       float* hall, size_t h_row, size_t h_t, int R, int T, int H, bool x3, void* scratch,
                       cudaStream_t st);
int split_tf32_launch(const float* in, size_t rows, int K, size_t ldi, const float* row_scale, int rows_per_scale,
                      float* out, int Kp, int cat, cudaStream_t st, int scale_B = 0);
int bias_act_launch(float* x, size_t rows, int N, size_t ld, const float* bias, int act, cudaStream_t st);
// workspace of lstm_layer_tc / linear_tc: prepared A operand [rows_T, Kmax (x3: 3 Kmax)], prepared weights
// [4 Hmax, same], hoisted projection P [rows_T, 4 Hmax], recurrence scratch
struct LstmTcWs { float *a, *w, *P; void* rec; };
void lstm_tc_carve(char* base, size_t& off, size_t rows_T, int Kmax, int Hmax, bool x3, LstmTcWs& ws);
int lstm_layer_tc(const fsn_lstm_layer& L, const float* x, size_t ldx, int K, const float* row_scale, int rows_per_scale,
                  int scale_B, int R, int T, int H, bool x3, const LstmTcWs& ws, float* hall, cudaStream_t st);
int linear_tc(const float* x, size_t ldx, int K, const float* W, const float* bias, int N, int act, float* out, size_t ldo,
              size_t rows, bool x3, const LstmTcWs& ws, cudaStream_t st);
int gemm_tc_split_launch(const float* a, size_t lda, const float* W, int N, int K, float* w, float* C, size_t ldc, size_t M,
                         bool x3, cudaStream_t st);

// tcgen05 sub-band stack (fsn_subband_tc.cu)
struct SbTcArgs {
  const void* packed;       // tile-ordered fp16 weights (fsn_pack_sb_weights)
  const float* magT; const float* fbT; const float* inv2;
  const float* unit_scale;  // nullable: per-row scale of this step (cumulative norm) instead of inv2[clip]
  float* crm;
  int B, F, Tp, la, Ns, Nf, H, act;
  int steps, shrink;      // pair kernel only: LSTM steps (0 = Tp) and time down-sampling of the gathered input (0/1 = none)
  bool pair;              // packed for / run by the CTA-pair kernel
  bool x3;                // pair kernel only: error-compensated variant (FSN_PREC_F16X3_TC image)
  bool quad;              // packed for / run by the 6-CTA-cluster kernel (fsn_subband_tc4.cu)
  RowMap map;
};
size_t sb_tc_packed_bytes(const fsn_model_desc* d);
int sb_tc_pack(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, cudaStream_t st);
int sb_tc_forward(const SbTcArgs& a, cudaStream_t st);
bool sb_tc_supported(const fsn_model_desc* d);

// CTA-pair (cta_group::2) variant (fsn_subband_tc2.cu); preferred when the shape allows (H = 384)
bool sb_tc2_supported(const fsn_model_desc* d);
size_t sb_tc2_packed_bytes(bool x3 = false);
int sb_tc2_pack(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, cudaStream_t st);
int sb_tc2_forward(const SbTcArgs& a, cudaStream_t st);
int sb_tc2_pack_raw(const fsn_seq_weights* sb, int Ksb, int fc_out, void* packed, cudaStream_t st, bool x3 = false);
// cluster (three pairs = 6 CTAs, N = 128) variant (fsn_subband_tc4.cu); experimental, single fp16 pass, H = 384; FSN_TC_CLUSTER4=1
bool sb_tc4_supported(const fsn_model_desc* d);
size_t sb_tc4_packed_bytes();
int sb_tc4_pack(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, cudaStream_t st);
int sb_tc4_forward(const SbTcArgs& a, cudaStream_t st);
bool sb_tc2_enabled();  // H = 384 stacks may use the pair kernel (FSN_TC_PAIR != 0)

}  // namespace fsn
