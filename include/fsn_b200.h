/*
 * fsn_b200.h - C ABI of libfsn_b200.so: the B200 (sm_100a) implementation of FullSubNet's
 * enhancement hot path.
 *
 * The reference (Audio-WestlakeU/FullSubNet) is pure Python and has no FFI; its "operator"
 * boundary is the set of Python callables listed below.  Each entry point here replaces the
 * device work of one of them and is what `fullsubnet_b200/` (the Python host mirroring the
 * reference API) binds through ctypes.  Paths are relative to the upstream repository.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless noted; the caller allocates
 *     all buffers including the workspace; the library never allocates, frees or retains
 *     DATA pointers across calls, so compute calls are re-entrant across streams and threads.  The
 *     only mutable state it keeps is diagnostic and host-side: the thread-local last-error string /
 *     code, the thread-local profiling switch with its per-stage CUDA events (fsn_set_profiling,
 *     fsn_last_stage_ms), a thread-local launch counter (fsn_last_launch_count) and one process-wide
 *     launch total (fsn_total_launch_count, a relaxed counter); none of it influences results;
 *   - `stream` is a cudaStream_t; nothing synchronises the host;
 *   - return value 0 = ok, non-zero = error; fsn_last_error() gives a thread-local message.
 *     Shape-contract violations that are AssertionError / NotImplementedError in the reference
 *     come back as FSN_ERR_SHAPE / FSN_ERR_UNSUPPORTED and the Python host re-raises them as
 *     the reference's exception types.
 */
#ifndef FSN_B200_H
#define FSN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fsn_stream_t; /* cudaStream_t */

enum {
  FSN_OK = 0,
  FSN_ERR_SHAPE = 1,       /* reference: AssertionError */
  FSN_ERR_UNSUPPORTED = 2, /* reference: NotImplementedError */
  FSN_ERR_CUDA = 3,
  FSN_ERR_WORKSPACE = 4
};

enum { FSN_ACT_NONE = 0, FSN_ACT_RELU = 1, FSN_ACT_TANH = 2, FSN_ACT_RELU6 = 3 };
/* FSN_NORM_CUMULATIVE_LAPLACE (audio_zen/model/base_model.py:220-251): causal running mean per clip (first norm) and
 * per sub-band unit (second norm); built for the fp32 inference path of fsn_model_forward / fsn_enhance */
enum { FSN_NORM_OFFLINE_LAPLACE = 0, FSN_NORM_CUMULATIVE_LAPLACE = 1 };
enum { FSN_CELL_LSTM = 0, FSN_CELL_GRU = 1 };
/* arithmetic of the sub-band LSTM stack (99 % of the FLOPs):
 *   FSN_PREC_FP32     - fp32 FMA everywhere (bit-for-bit class of the reference CPU path, ~1e-6)
 *   FSN_PREC_TF32_TC  - training step only (fsn_train_*): every GEMM of the forward, of back-propagation through
 *                       time and of the weight gradients on tcgen05 kind::tf32 (fp32 data read as tf32, fp32
 *                       accumulate in TMEM); gate / cell arithmetic and all reductions stay fp32
 *   FSN_PREC_F16_TC   - fp16 operands (11-bit significand, like TF32) x fp32 accumulate on the
 *                       tcgen05 tensor cores, fp32 cell state; cRM within 1e-3 rel (tests)
 *   FSN_PREC_F16X3_TC - error-compensated tensor-core path: weights and state split into fp16 hi + lo terms, every
 *                       product issued as W_hi.S_hi + W_hi.S_lo + W_lo.S_hi into one fp32 TMEM accumulator (22
 *                       significand bits per operand), libm-class gate functions; the fp32 error class (cRM ~1e-6
 *                       rel), needed where decompress_cIRM amplifies mask errors x100 (|cRM| near the 9.9 clip) */
enum { FSN_PREC_FP32 = 0, FSN_PREC_F16_TC = 1, FSN_PREC_TF32_TC = 2, FSN_PREC_F16X3_TC = 3 };

int fsn_version(void);
const char* fsn_last_error(void);
/* status code (FSN_ERR_*) of the last failed call on this thread: lets the *_workspace_bytes() functions, which
 * return 0 on failure, report WHY (shape error vs unsupported configuration) */
int fsn_last_error_code(void);
/* compile-time facts for the host (sm arch the kernels were built for, e.g. 100) */
int fsn_built_arch(void);

/* ------------------------------------------------------------------------------------------
 * audio_zen/acoustics/feature.py:9-50  stft(y, n_fft, hop_length, win_length)
 *   wav [B,L] -> mag, phase, real, imag, each [B,F,T]; F = n_fft/2+1, T = 1 + L/hop.
 *   torch.stft semantics: center=True reflect pad n_fft/2, periodic hann (zero-padded to n_fft
 *   when win_length < n_fft), one-sided, un-normalised.  `phase` may be NULL (not computed).
 *   `magT` (optional, may be NULL): a second, time-major copy [B, T_pad, F] with rows
 *   T..T_pad-1 zeroed - the layout the model kernels consume (look-ahead pad fused,
 *   recipes/dns_interspeech_2020/fullsubnet/model.py:85).
 * ---------------------------------------------------------------------------------------- */
int fsn_stft(const float* wav, int B, int L, int n_fft, int hop, int win_length,
             float* mag, float* phase, float* real, float* imag,
             float* magT, int T_pad, fsn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * audio_zen/acoustics/feature.py:53-91  istft(features, n_fft, hop, win, length, input_type)
 *   real/imag [B,F,T] with element stride `cstride` (1 = planar "real_imag", 2 = interleaved
 *   complex64) -> wav [B,out_len].  torch.istft semantics: irfft (1/N), window, overlap-add,
 *   divide by the window-square envelope, trim n_fft/2, cut/zero-pad to `length`
 *   (length <= 0: hop*(T-1)).
 *   If `crm` != NULL ([B,2,F,T], recipes/dns_interspeech_2020/inferencer.py:130-145) the
 *   spectrum is first multiplied by decompress_cIRM(crm) (audio_zen/acoustics/mask.py:47-64,
 *   K=10, limit=9.9) as a complex mask - rows A9 of SURVEY 8a in one kernel.
 * ---------------------------------------------------------------------------------------- */
int fsn_istft(const float* real, const float* imag, int cstride, const float* crm,
              int B, int T, int n_fft, int hop, int win_length, int length,
              float* wav, fsn_stream_t stream);

/* audio_zen/acoustics/mask.py:47-64 / :32-44 / :7-29 (elementwise, n = number of elements) */
int fsn_decompress_cirm(const float* in, float* out, int64_t n, float K, float limit, fsn_stream_t stream);
int fsn_compress_cirm(const float* in, float* out, int64_t n, float K, float C, fsn_stream_t stream);
/* noisy/clean real/imag [n] -> cIRM [n,2] (compressed, K=10, C=0.1) */
int fsn_build_cirm(const float* nr, const float* ni, const float* cr, const float* ci,
                   float* out, int64_t n, fsn_stream_t stream);
/* audio_zen/acoustics/feature.py:309-345  drop_band: in [B,C,F,T] -> out [B,C,F/G,T] */
int fsn_drop_band(const float* in, float* out, int B, int C, int F, int T, int G, fsn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * recipes/dns_interspeech_2020/fullsubnet/model.py:9-136  Model
 * ---------------------------------------------------------------------------------------- */
typedef struct fsn_model_desc {
  int32_t num_freqs;        /* F */
  int32_t look_ahead;
  int32_t fb_num_neighbors; /* Nf */
  int32_t sb_num_neighbors; /* Ns */
  int32_t fb_hidden;        /* 512 */
  int32_t sb_hidden;        /* 384 */
  int32_t fb_activation;    /* FSN_ACT_* (fb_output_activate_function) */
  int32_t sb_activation;    /* FSN_ACT_* (sb_output_activate_function) */
  int32_t norm_type;        /* FSN_NORM_* */
  int32_t num_groups_in_drop_band; /* applied when B > 1 (model.py:114), 1 = off */
  int32_t precision;        /* FSN_PREC_* for the sub-band stack */
  int32_t cell_type;        /* FSN_CELL_*: `sequence_model` = "LSTM" | "GRU" (sequence_model.py:52-66); GRU: weights
                             * [3H,K] with gate order r,z,n, inference on the fp32 kernels (FSN_PREC_FP32) only */
} fsn_model_desc;

/* One SequenceModel (audio_zen/model/module/sequence_model.py:26-125): 2-layer nn.LSTM +
 * Linear, PyTorch parameter layout (gate order i,f,g,o; weight [4H,K] row-major).  Pointers
 * go straight into the nn.Parameter storage so Adam / checkpoints / DDP keep working. */
typedef struct fsn_seq_weights {
  const float* w_ih[2];
  const float* w_hh[2];
  const float* b_ih[2];
  const float* b_hh[2];
  const float* fc_w; /* [out, H] */
  const float* fc_b; /* [out] */
} fsn_seq_weights;

/* bytes of caller-provided scratch for fsn_model_forward / fsn_enhance at batch B, T frames */
size_t fsn_model_workspace_bytes(const fsn_model_desc* d, int B, int T);

/* FSN_PREC_F16_TC / FSN_PREC_F16X3_TC only: bytes of, and packer for, the tile-ordered fp16 image of the sub-band
 * weights that the tcgen05 kernel streams (cache it keyed on the parameters' version AND the precision: the
 * compensated image carries a hi and a lo stage per k range). */
size_t fsn_sb_packed_bytes(const fsn_model_desc* d);
int fsn_pack_sb_weights(const fsn_model_desc* d, const fsn_seq_weights* sb, void* packed, fsn_stream_t stream);

/* Model.forward (model.py:72-136): noisy_mag [B,1,F,T] -> crm [B,2,F',T]
 *   F' = F, or F/G with the drop_band batch permutation when B > 1 and G > 1.
 *   sb_packed: NULL for FSN_PREC_FP32. */
int fsn_model_forward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                      const void* sb_packed, const float* noisy_mag, int B, int T, float* crm,
                      void* workspace, size_t workspace_bytes, fsn_stream_t stream);

/* recipes/dns_interspeech_2020/inferencer.py:130-145  Inferencer.full_band_crm_mask, batched
 * over independent clips (drop_band off): wav [B,L] -> enhanced [B,L]; crm_out (optional,
 * [B,2,F,T]) receives the mask.  One call = stft -> model -> decompress/mask -> istft. */
size_t fsn_enhance_workspace_bytes(const fsn_model_desc* d, int B, int L, int n_fft, int hop);
int fsn_enhance(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                const void* sb_packed, const float* wav, int B, int L, int n_fft, int hop, int win_length,
                float* enhanced, float* crm_out, void* workspace, size_t workspace_bytes,
                fsn_stream_t stream);
/* The same call followed by the int16 scaling of the reference host loop (audio_zen/inferencer/base_inferencer.py:
 * 181-182: int16(gain * y / max|y|), gain = 0.8 * 32767): max|y| per clip is reduced in the iSTFT epilogue, so the
 * float waveform is read once more and only 2 bytes per sample have to go back to the host.  `enhanced` (float,
 * [B,L]) is still written; pcm [B,L] int16. */
int fsn_enhance_pcm(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                    const void* sb_packed, const float* wav, int B, int L, int n_fft, int hop, int win_length,
                    float* enhanced, int16_t* pcm, float gain, void* workspace, size_t workspace_bytes,
                    fsn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * recipes/dns_interspeech_2020/fast_fullsubnet/model.py:11-202  Model (BASELINE config 4)
 *   MelScale(F->M) -> norm -> encoder LSTM(M->He1), LSTM(He1->He2)+Linear(M)+ReLU -> unfold(noisy mel, Nn) ||
 *   unfold(encoder out, Ne) -> real-time down-sampling x`shrink` -> norm -> bottleneck 2xLSTM(Hb)+Linear(1)+ReLU on
 *   B*M rows -> up-sampling -> decoder LSTM(2M->Hd), LSTM(Hd->Hd)+Linear(2F) -> [B,2,F,T].  fp32 kernels.
 * ---------------------------------------------------------------------------------------- */
typedef struct fsn_lstm_layer {
  const float* w_ih; /* [4H, K] */
  const float* w_hh; /* [4H, H] */
  const float* b_ih;
  const float* b_hh;
} fsn_lstm_layer;

typedef struct fsn_fast_desc {
  int32_t num_freqs;   /* encoder_input_size (257) */
  int32_t look_ahead;
  int32_t shrink_size;
  int32_t num_mels;    /* 64 */
  int32_t enc1_hidden; /* 384 */
  int32_t enc2_hidden; /* 257 */
  int32_t bn_hidden;   /* bottleneck_hidden_size */
  int32_t bn_layers;   /* bottleneck_num_layers (2) */
  int32_t dec_hidden;  /* 512 */
  int32_t noisy_num_neighbors; /* noisy_input_num_neighbors */
  int32_t enc_num_neighbors;   /* encoder_output_num_neighbors */
  int32_t precision;           /* FSN_PREC_* for the bottleneck stack (the tensor-core path needs bn_hidden = 384) */
} fsn_fast_desc;

typedef struct fsn_fast_weights {
  const float* mel_fb;  /* mel_scale.fb [F, M] */
  fsn_lstm_layer enc1, enc2;
  const float* enc_fc_w; const float* enc_fc_b;   /* [M, He2], [M] */
  fsn_lstm_layer bn[2];
  const float* bn_fc_w; const float* bn_fc_b;     /* [1, Hb], [1] */
  fsn_lstm_layer dec1, dec2;
  const float* dec_fc_w; const float* dec_fc_b;   /* [2F, Hd], [2F] */
  const void* bn_packed;  /* FSN_PREC_F16_TC: fsn_fast_pack_bn_weights() image of the bottleneck stack, else NULL */
} fsn_fast_weights;

size_t fsn_fast_workspace_bytes(const fsn_fast_desc* d, int B, int T);
/* FSN_PREC_F16_TC: 0 when the tensor-core path cannot run this descriptor */
size_t fsn_fast_packed_bytes(const fsn_fast_desc* d);
int fsn_fast_pack_bn_weights(const fsn_fast_desc* d, const fsn_fast_weights* w, void* packed, fsn_stream_t stream);
/* Model.forward (fast_fullsubnet/model.py:143-202): mix_mag [B,1,F,T] -> [B,2,F,T] */
int fsn_fast_model_forward(const fsn_fast_desc* d, const fsn_fast_weights* w, const float* mix_mag, int B, int T,
                           float* out, void* workspace, size_t workspace_bytes, fsn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * recipes/dns_interspeech_2020/improved_fullsubnet/model.py:452-591  Model (BASELINE config 5)
 *   wav -> STFT -> |X|^fdrc, Nyquist bin dropped -> norm -> full-band 2xLSTM + Linear -> per sub-band section:
 *   strided unfold (centre/neighbour widths) of noisy and full-band output, concat, per-section norm, 2xLSTM +
 *   Linear(2*centre) -> cRM (Nyquist row 0) -> element-wise mask on (re, im) -> iSTFT -> wav.  fp32 kernels;
 *   n_fft must be a power of two (the reference's n_fft=960 example needs a mixed-radix FFT: not built).
 * ---------------------------------------------------------------------------------------- */
#define FSN_IMP_MAX_SECTIONS 8
typedef struct fsn_improved_desc {
  int32_t n_fft, hop_length, win_length, num_freqs;
  float fdrc;
  int32_t num_sections;                       /* len(sb_num_center_freqs) = len(freq_cutoffs) + 1 */
  int32_t freq_cutoffs[FSN_IMP_MAX_SECTIONS];
  int32_t sb_num_center[FSN_IMP_MAX_SECTIONS], sb_num_neighbor[FSN_IMP_MAX_SECTIONS];
  int32_t fb_num_center[FSN_IMP_MAX_SECTIONS], fb_num_neighbor[FSN_IMP_MAX_SECTIONS];
  int32_t fb_hidden, sb_hidden, fb_activation, sb_activation;
  int32_t precision; /* FSN_PREC_FP32, or FSN_PREC_TF32_TC: the sub-band sections' GEMMs on tcgen05 kind::tf32 */
} fsn_improved_desc;

typedef struct fsn_improved_weights {
  fsn_seq_weights fb;                          /* fb_model */
  fsn_seq_weights sb[FSN_IMP_MAX_SECTIONS];    /* sb_model.sb_models[s] */
} fsn_improved_weights;

size_t fsn_improved_workspace_bytes(const fsn_improved_desc* d, int B, int L);
/* Model.forward (improved_fullsubnet/model.py:541-591): wav [B,L] -> enhanced [B,L] (the reference returns
 * [B,1,L]); crm_out optional [B,2,F,T] */
int fsn_improved_forward(const fsn_improved_desc* d, const fsn_improved_weights* w, const float* wav, int B, int L,
                         float* enhanced, float* crm_out, void* workspace, size_t workspace_bytes,
                         fsn_stream_t stream);

/* Opt-in stage timing for bench.py: when enabled, fsn_model_forward / fsn_enhance bracket their
 * stages with CUDA events on `stream` (thread-local, created lazily).  After the caller has
 * synchronised the stream, fsn_last_stage_ms(stage) returns the device time of the last call:
 * stage 0 = stft (or mag transpose), 1 = norms + full-band stack, 2 = sub-band stack, 3 = mask+istft. */
int fsn_set_profiling(int enable);
float fsn_last_stage_ms(int stage);

/* number of kernel launches issued by the last fsn_model_forward / fsn_enhance on this thread
 * (bench.py reports it as gpu_launches) */
int64_t fsn_last_launch_count(void);
/* kernels launched by this library since it was loaded (never reset): difference two readings */
int64_t fsn_total_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Training step: recipes/dns_interspeech_2020/fullsubnet/trainer.py:56-68, fp32.
 *   fsn_train_forward   = Model.forward in train mode (model.py:72-136, drop_band on) that keeps the activations
 *                         back-propagation through time needs in `workspace` (same buffer must be passed to
 *                         fsn_train_backward, untouched in between)
 *   fsn_mse_loss        = audio_zen/loss.py:4 (torch.nn.MSELoss) between cIRM [B',F',T,2] (trainer.py:49-54) and
 *                         cRM [B',2,F',T]; also writes d loss / d cRM when dcrm != NULL.  loss: device scalar.
 *   fsn_train_backward  = loss.backward() (trainer.py:63): gradients of the 20 parameters, OVERWRITTEN into the
 *                         buffers of gfb / gsb (same shapes as the parameters)
 *   fsn_clip_adam       = clip_grad_norm_(max_norm) + Adam step (trainer.py:65-68, train.py:55-59) over a list of
 *                         tensors, no host synchronisation.  grad_scale multiplies every gradient first (1/world
 *                         after a sum all-reduce).  norm_out (optional, 2 floats on the device) receives the total
 *                         norm and the applied coefficient; gradients are left clipped like the reference. */
typedef struct fsn_seq_grads {
  float* w_ih[2];
  float* w_hh[2];
  float* b_ih[2];
  float* b_hh[2];
  float* fc_w;
  float* fc_b;
} fsn_seq_grads;

size_t fsn_train_workspace_bytes(const fsn_model_desc* d, int B, int T);
int fsn_train_forward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                      const float* noisy_mag, int B, int T, float* crm, void* workspace, size_t workspace_bytes,
                      fsn_stream_t stream);
int fsn_train_backward(const fsn_model_desc* d, const fsn_seq_weights* fb, const fsn_seq_weights* sb,
                       const float* dcrm, int B, int T, const fsn_seq_grads* gfb, const fsn_seq_grads* gsb,
                       void* workspace, size_t workspace_bytes, fsn_stream_t stream);

size_t fsn_mse_loss_scratch_bytes(void);
int fsn_mse_loss(const float* cirm, const float* crm, int B, int Fsub, int T, float* loss, float* dcrm,
                 void* scratch, size_t scratch_bytes, fsn_stream_t stream);

#define FSN_MAX_PARAM_TENSORS 64
typedef struct fsn_param_list {
  int n;
  float* param[FSN_MAX_PARAM_TENSORS];
  float* grad[FSN_MAX_PARAM_TENSORS];
  float* exp_avg[FSN_MAX_PARAM_TENSORS];
  float* exp_avg_sq[FSN_MAX_PARAM_TENSORS];
  int64_t numel[FSN_MAX_PARAM_TENSORS];
} fsn_param_list;

size_t fsn_clip_adam_scratch_bytes(void);
int fsn_clip_adam(const fsn_param_list* L, float max_norm, float grad_scale, float lr, float beta1, float beta2,
                  float eps, int step, float* norm_out, void* scratch, size_t scratch_bytes, fsn_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * recipes/dns_interspeech_2020/fullband_baseline/model.py:8-68  Model
 *   look-ahead pad -> norm -> num_layers x LSTM(F -> H) -> Linear(H -> 2F) [+ activation] -> [B,2,F,T]; fp32 kernels.
 *   layers: num_layers entries (PyTorch parameter layout); fc_w [2F,H], fc_b [2F]. */
typedef struct fsn_fullband_desc {
  int32_t num_freqs;
  int32_t hidden;
  int32_t num_layers; /* the reference builds 3 */
  int32_t look_ahead;
  int32_t activation; /* FSN_ACT_* */
  int32_t norm_type;  /* FSN_NORM_* */
} fsn_fullband_desc;

size_t fsn_fullband_workspace_bytes(const fsn_fullband_desc* d, int B, int T);
int fsn_fullband_forward(const fsn_fullband_desc* d, const fsn_lstm_layer* layers, const float* fc_w, const float* fc_b,
                         const float* noisy_mag, int B, int T, float* out, void* workspace, size_t workspace_bytes,
                         fsn_stream_t stream);

/* audio_zen/inferencer/base_inferencer.py:181-182: out = int16(gain * wav / max|wav|) per clip,
 * gain = 0.8 * 32767 in the reference; float32 multiply, divide, truncation toward zero like numpy; all-zero clip -> 0 */
int fsn_peak_normalize_int16(const float* wav, int B, int L, float gain, int16_t* out, fsn_stream_t stream);

/* audio_zen/metrics.py:6-31  SI_SDR(reference, estimation) (the validation metric of
 * fullsubnet/trainer.py:78-181 that is pure arithmetic; STOI / PESQ are third-party CPU packages and stay out).
 * reference, estimation [B,L] -> out[B] in dB; fixed-order reductions. */
int fsn_si_sdr(const float* reference, const float* estimation, int B, int L, float* out, fsn_stream_t stream);

/* recipes/dns_interspeech_2020/dataset_train.py:136-199  Dataset.snr_mix for a batch: the random
 * draws (snr, noisy target dBFS, which RIR) are made by the caller and passed in.
 *   fsn_rir_convolve: out[b,:L] = fftconvolve(x[b], rir[b,:rir_len[b]])[:L]  (dataset_train.py:161; direct form,
 *                     rir [B,Lr_max], rir_len[b] == 0 copies the clip; rir_len may be NULL = Lr_max everywhere)
 *   fsn_snr_mix:      norm_amplitude + tailor_dB_FS(target_dB_FS) of clean and noise, noise scaled to snr[b] dB,
 *                     mixture tailored to noisy_target_dB_FS[b] (clean by the same factor), both divided by
 *                     max|noisy| / (0.99 - eps) when the mixture exceeds 0.999 (audio_zen/acoustics/feature.py:99-114). */
int fsn_rir_convolve(const float* x, const float* rir, const int* rir_len, int B, int L, int Lr_max, float* out,
                     fsn_stream_t stream);
int fsn_snr_mix(const float* clean, const float* noise, const float* snr, const float* noisy_target_dB_FS,
                float target_dB_FS, float eps, int B, int L, float* noisy_out, float* clean_out, fsn_stream_t stream);

/* unit-test hooks (host code only): the drop_band row map of Model.forward and its inverse (-1 = unit dropped), and
 * the reflect-padding multiplicity c[r] of the closed-form second norm */
int fsn_debug_row_to_unit(int B, int F, int G, int r, int* b, int* f);
int fsn_debug_unit_to_row(int B, int F, int G, int b, int f);
int fsn_debug_reflect_count(int r, int F, int N);

/* unit-test hook for the tf32 tcgen05 GEMM of the training path: C[M,N] (+)= A[M,K] B[N,K]^T, fp32 row-major
 * operands with 16-byte aligned rows; scratch (optional) enables split-K */
int fsn_debug_tgemm(const float* A, int64_t lda, const float* B, int64_t ldb, float* C, int64_t ldc, int M, int N,
                    int K, int accumulate, float* scratch, int64_t scratch_floats, fsn_stream_t stream);

/* unit-test hook for the weight-gradient GEMMs of the training step (dW = dG^T X, torch autograd of nn.LSTM):
 * C[M,N] = A[a_k0:a_k0+K, :M]^T B[b_k0:b_k0+K, :N] for row-major A [a_k0+K, M], B [b_k0+K, N]; both operands are first
 * copied into the block-tiled K-major layout the TMA loads stream (a_k0, b_k0 multiples of 32); scratch holds the two
 * copies (rounded up to 128 x 32 tiles) followed by split-K space */
int fsn_debug_tgemm_blocked(const float* A, const float* B, float* C, int M, int N, int K, int a_k0, int b_k0,
                            float* scratch, int64_t scratch_floats, fsn_stream_t stream);

/* unit-test hook for the fused forward step of the training path (one LSTM step of torch.nn.LSTM as used by
 * audio_zen/model/module/sequence_model.py:52-58): z = x W_ih^T (or the projection already in G) + b_ih + b_hh +
 * h_prev W_hh^T; G [R,4H] <- post-activation gates (i,f,g,o), c_out = f c_prev + i g, h_out = o tanh(c_out).
 * h_prev, c_prev nullable (first step); x nullable (G then holds x W_ih^T on entry).  half != 0: fp16 MMA operands
 * (converted into scratch, >= 2 * (2 R H + 4 H (H + K0) + R K0) + 1024 bytes); H % 32 == 0 */
int fsn_debug_lstm_fwd_step(const float* h_prev, const float* w_hh, const float* x, const float* w_ih, int K0, float* G,
                            const float* b_ih, const float* b_hh, const float* c_prev, float* c_out, float* h_out, int R,
                            int H, int half, void* scratch, int64_t scratch_bytes, fsn_stream_t stream);

/* unit-test hooks for the tensor-core LSTM layer of the full-band stacks (fsn_lstm_rec_tc.cu;
 * audio_zen/model/module/sequence_model.py:52-58,117): hall[r,t,:] of nn.LSTM(K -> H, 1 layer) over x [R,T,K]
 * (hoisted input-projection GEMM + persistent tcgen05 recurrence), and out = act(x W^T + b) for x [rows,K], W [N,K];
 * x3 != 0 selects the compensated (fp32-class) arithmetic.  Workspace of the Linear hook: the LSTM one with
 * R*T = rows, H = max(8, ceil(N/4)). */
size_t fsn_debug_lstm_tc_workspace_bytes(int R, int T, int K, int H, int x3);
int fsn_debug_lstm_layer_tc(const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh, const float* x,
                            int R, int T, int K, int H, int x3, float* hall, void* workspace, size_t workspace_bytes,
                            fsn_stream_t stream);
int fsn_debug_linear_tc(const float* x, int rows, int K, const float* W, const float* bias, int N, int act, int x3,
                        float* out, void* workspace, size_t workspace_bytes, fsn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FSN_B200_H */
