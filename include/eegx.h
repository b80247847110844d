/*
 * eegx.h -- C ABI of libeegx.so: the B200 (sm_100a) implementation of the
 * EEG preprocessing + encoder hot path of alexsteinerr/imagined-speech-translation.
 *
 * The reference has no FFI / operator layer of its own;
 * its boundary is the Python API of main_model/src.  Each entry point below
 * therefore cites the reference *Python* code it replaces.  The Python classes
 * in imagined_speech_translation_b200/ keep the reference signatures and call
 * these functions through ctypes on tensor.data_ptr().
 *
 * Conventions (all entry points):
 *   - return 0 on success, a negative eegx_status otherwise; the message is in
 *     eegx_last_error() (thread-local).  No exceptions, no exit(), NO CPU
 *     FALLBACK: on a device that is not sm_100 every compute call returns
 *     EEGX_ERR_ARCH.
 *   - every pointer except plans is DEVICE memory owned by the caller
 *     (row-major, contiguous, float32 unless stated); inputs are const; outputs
 *     must not alias inputs.
 *   - work is enqueued asynchronously on `stream` (a cudaStream_t passed as
 *     void*; NULL = legacy default stream); no host sync, no allocation inside
 *     compute calls, so they are CUDA-graph capturable.
 *   - plans own only O(KB) constant tables (FIR taps, window, twiddles).
 *   - grids are sized for B200: the SM count (148) is a compile-time constant of the library
 *     (csrc/eegx_common.h, kNumSMsB200) -- persistent kernels launch one CTA (or CTA pair, or a fixed
 *     number of CTAs) per SM of that part, and the sm_100 check above is what keeps other devices out.
 */
#ifndef EEGX_H
#define EEGX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EEGX_VERSION 100 /* 0.1.0 */

typedef enum eegx_status {
    EEGX_OK = 0,
    EEGX_ERR_ARCH = -1,      /* device is not compute capability 10.x */
    EEGX_ERR_SHAPE = -2,     /* unsupported / inconsistent sizes */
    EEGX_ERR_ALIGN = -3,     /* pointer not aligned as required */
    EEGX_ERR_WORKSPACE = -4, /* workspace too small */
    EEGX_ERR_CUDA = -5,      /* a CUDA runtime call failed */
    EEGX_ERR_ARG = -6        /* NULL / invalid argument */
} eegx_status;

int eegx_version(void);
const char* eegx_last_error(void);
/* 0 if the current CUDA device can run this library (sm_100), else EEGX_ERR_ARCH. */
int eegx_device_check(void);

/* ------------------------------------------------------------------------
 * Reference-actual normalisation.
 * Replaces EEGDataset._process_raw_eeg + _normalize_eeg_sample
 * (main_model/src/data/dataset.py:172-191, 193-225) for a whole batch:
 *   v   = x[b, ch_idx[j], t]; nan -> 0, +inf -> 10, -inf -> -10
 *   out = (v - center[j]) / scale[j]                (RobustScaler.transform)
 * x: (B, C_in, T).  ch_idx/center/scale: (C_out,) device arrays (the four
 * regions concatenated in the order frontal, temporal, central, parietal).
 * Output addressing: channel j of trial b is written at
 *   out + out_off[j] + b * out_bstride[j] + t
 * so the four regions can land in four dense (B, C_r, T) buffers carved from
 * one allocation.  out_off / out_bstride may both be NULL: dense (B, C_out, T).
 * center/scale may both be NULL: only the gather + nan_to_num is applied.
 * ------------------------------------------------------------------------ */
int eegx_normalize_f32(const float* x, const int32_t* ch_idx, const float* center,
                       const float* scale, float* out, const int64_t* out_off,
                       const int64_t* out_bstride, int64_t B, int64_t C_in,
                       int64_t C_out, int64_t T, void* stream);

/* The scaler-less fallback branch of _normalize_eeg_sample (dataset.py:213-216):
 * per (trial, channel) over time, population std: (v - mean_t) / (std_t + 1e-8),
 * after the same gather + nan_to_num.  Same addressing as eegx_normalize_f32. */
int eegx_zscore_time_f32(const float* x, const int32_t* ch_idx, float* out,
                         const int64_t* out_off, const int64_t* out_bstride,
                         int64_t B, int64_t C_in, int64_t C_out, int64_t T,
                         void* stream);

/* ------------------------------------------------------------------------
 * north_star DSP chain (spec: SURVEY.md section 8(c); absent from the reference):
 *   trial windowing -> band-pass FIR ("same", zero padded, delay compensated)
 *   -> STFT (center=True, reflect pad, periodic Hann, one-sided)
 *   -> log(|X|^2 + log_eps) -> per-(trial, channel) z-score over all F*N_f
 *      values, population std, (L - mu) / (sigma + z_eps).
 * One fused launch: x is read once, out is written once.
 * ------------------------------------------------------------------------ */
typedef struct eegx_dsp_plan eegx_dsp_plan;

/* fir: HOST pointer to numtaps float32 taps (numtaps odd, <= 129).
 * n_fft: power of two in [32, 2048]; T > n_fft/2; hop >= 1.
 * Output dims: F = n_fft/2 + 1, N_f = 1 + T / hop. */
int eegx_dsp_plan_create(eegx_dsp_plan** plan, int C, int T, int n_fft, int hop,
                         const float* fir, int numtaps, float log_eps, float z_eps);
int eegx_dsp_plan_destroy(eegx_dsp_plan* plan);
int eegx_dsp_plan_dims(const eegx_dsp_plan* plan, int* F, int* N_f);
/* Which kernel the plan dispatches to: 0 = generic, 1 = tuned T=2048/n_fft=256/hop=64/K=65 (BASELINE configs[1-2]),
 * 2 = tuned T=4096/n_fft=1024/hop=256/K=65 (configs[3]), 3 = float64 (set_precise). */
int eegx_dsp_plan_kernel(const eegx_dsp_plan* plan);
/* Force the generic kernel (testing / A-B comparison). */
int eegx_dsp_plan_force_generic(eegx_dsp_plan* plan, int on);
/* Run the chain in float64 (FIR accumulation, FFT butterflies, log, statistics; float32 in and out).  About a
 * tenth of the tuned kernels' throughput; reaches the spec's 1e-5 bound at n_fft = 1024, where float32 arithmetic
 * measures 1.1e-5 .. 1.9e-5 against float64. */
int eegx_dsp_plan_set_precise(eegx_dsp_plan* plan, int on);

/* onsets == NULL: x is (B, C, T), trials already cut.
 * onsets != NULL: x is one continuous recording (C, rec_len) and
 *                 trial b = x[:, onsets[b] : onsets[b] + T]  (int64 device array,
 *                 every window must lie inside [0, rec_len)).
 * out: (B, C, F, N_f) float32. */
int eegx_dsp_forward(const eegx_dsp_plan* plan, const float* x, const int64_t* onsets,
                     int64_t rec_len, float* out, int64_t B, void* stream);

/* ------------------------------------------------------------------------
 * bf16 tensor-core GEMM (tcgen05 / TMEM / TMA), the contraction core of the
 * encoder.  Replaces the cuDNN / cuBLAS library calls behind nn.Conv1d and
 * nn.Linear in Conv1DWithAttention / BrainRegionEncoder
 * (main_model/src/models/layers.py:30-127, brain_encoder.py:31-92; SURVEY.md
 * section 2 "library-call sites that become sm_100a kernels").
 *
 *   D[g][b] (M x N) = alpha * A[g][b] (M x K) * B[g][b] (N x K)^T  (+ bias[g][N]) (+ GELU) (+ D[g][b])
 *
 * A, B: bf16.  a_mn_major = 0: A is stored M x K (K contiguous, leading dim lda);
 *              a_mn_major = 1: A is stored K x M (M contiguous).  Same for B / N.
 * With x (M x K), W (N x K), dy (M x N) this gives, without any transpose in HBM:
 *   forward  y  = x W^T      : A = x  (K-major),  B = W  (K-major)
 *   dgrad    dx = dy W       : A = dy (K-major),  B = W  (MN-major, "K" = N)
 *   wgrad    dW = dy^T x     : A = dy (MN-major), B = x  (MN-major, "K" = M)
 * D: bf16 (out_f32 = 0) or fp32 (out_f32 = 1), row-major, leading dim ldd.
 * epilogue: 0 none, 1 + bias, 2 + bias then exact (erf) GELU.  accumulate: D += result.
 * lda / ldb / batch strides: multiples of 8 elements; A, B, D 16-byte aligned.
 * lda (ldb) may be smaller than the contiguous extent: rows then overlap, which is how a
 * channels-last Conv1d (kernel k, C_in channels) runs as an implicit-im2col GEMM with
 * K = k * C_in and lda = C_in (layers.py:30-47).
 * ------------------------------------------------------------------------ */
typedef struct eegx_gemm_desc {
    int64_t M, N, K, batch;
    int64_t lda, ldb, ldd;
    int64_t stride_a, stride_b, stride_d; /* batch strides in elements */
    int32_t a_mn_major, b_mn_major;
    int32_t out_f32;
    int32_t epilogue;
    int32_t accumulate;
    int32_t force_block_n; /* 0 = auto (wave-quantisation cost model); 64 / 128 / 192 / 256 pins the N tile */
    float alpha;
    int32_t reserved;
    /* Grouped launch (0 or 1 = off): `groups` independent problem sets in ONE launch -- the four region encoders,
     * whose layers have identical shapes and their own weights (brain_encoder.py:148-150 runs them one after the
     * other).  Problem (g, s), s < batch, reads A + g*stride_a_g + s*stride_a (same for B, D) and bias +
     * g*stride_bias_g.  Group strides: multiples of 8 elements for A and B. */
    int64_t groups;
    int64_t stride_a_g, stride_b_g, stride_d_g, stride_bias_g;
} eegx_gemm_desc;

int eegx_gemm_bf16(const eegx_gemm_desc* desc, const void* A, const void* B, const float* bias,
                   void* D, void* stream);

/* ------------------------------------------------------------------------
 * Optimizer step of EEGTrainer.train_epoch (main_model/src/training/trainer.py:101-113;
 * wiring main_model/scripts/train.py:199-241) over flat fp32 buffers.
 *   eegx_sumsq_f32      out[0] (+)= sum(g^2), two fixed-order stages (bit-stable); with
 *                       accumulate != 0 several parameter groups add into one global norm.
 *   eegx_adamw_clip_f32 clip_grad_norm_ coefficient min(1, max_norm / (sqrt(*grad_norm_sq) *
 *                       grad_scale + 1e-6)) computed on the device (grad_norm_sq may be NULL:
 *                       no clipping), then torch.optim.AdamW's update:
 *                         p *= 1 - lr*wd;  m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
 *                         p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
 *                       grad_scale folds a constant factor on g (e.g. 1/world_size).
 *                       w16 (nullable): n bf16 values, receives the updated parameters rounded to bf16 -- the
 *                       operand copy the GEMMs of the next step read (no separate cast kernels).
 * ------------------------------------------------------------------------ */
size_t eegx_sumsq_workspace_bytes(void);
int eegx_sumsq_f32(const float* g, int64_t n, float* out, int accumulate, void* workspace,
                   size_t workspace_bytes, void* stream);
int eegx_adamw_clip_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr,
                        float beta1, float beta2, float eps, float weight_decay, int64_t step,
                        const float* grad_norm_sq, float max_norm, float grad_scale, void* w16, void* stream);


/* ------------------------------------------------------------------------
 * Fused glue of the encoder (bf16 activations, fp32 parameters / statistics, fp32 math).
 * Every entry point is one (or two) HBM-bound passes replacing a chain of element-wise
 * library kernels of the reference modules; forward and backward are separate calls so
 * torch.autograd.Function wrappers can own the tape.
 *
 * Dropout: (rng_state, site, p).  rng_state -> two device uint64 {seed, step}; the keep-mask
 * of element group i at call site `site` is Philox-4x32-10(seed, step, site, i), regenerated
 * in backward (no mask tensor).  rng_state == NULL or p == 0 disables dropout.
 *
 * "Guarded rows" (CNN stack): a (B, T, C) activation is rows m = b*(T+2*pad) + pad + t of an
 * (M = B*(T+2*pad)) x C matrix whose other rows are zero, with `pad` more zero rows before
 * m = 0 and after m = M-1, so nn.Conv1d is a GEMM over overlapping rows.  Pointers address row
 * m = 0; "out" buffers of this kind have every row (guards included) written.
 *
 * Parameter groups (G / groups): the four region encoders of BrainRegionEncoder have layers of identical
 * shape with their own parameters (brain_encoder.py:148-150 runs them one after the other).  Entry points
 * that read parameters take G: the activation then holds G equal blocks of rows (G * B trials in ONE guarded
 * buffer), block g uses parameter set g (gamma / beta / weights / statistics stacked (G, ...), stride given or
 * C), and parameter gradients come out per group.  G = 1 is a single module.
 * ------------------------------------------------------------------------ */

/* nn.LayerNorm(C) (+ nn.GELU if act = 1) (+ nn.Dropout) on (rows, C) bf16
 * (main_model/src/models/layers.py:61-71, 84-127, 232, 240).  mean / rstd: (rows) fp32 saved for backward. */
/* groups > 1: rows split evenly; gamma / beta (and dgamma / dbeta) of group g at + g * param_stride elements. */
int eegx_layernorm_fwd_bf16(const void* x, const float* gamma, const float* beta, void* y, float* mean,
                            float* rstd, int64_t rows, int64_t C, int64_t groups, int64_t param_stride, float eps,
                            int act, const uint64_t* rng_state, uint32_t site, float p, void* stream);
size_t eegx_layernorm_bwd_workspace_bytes(int64_t C);
/* accumulate != 0: dgamma / dbeta += (they may point straight into the parameters' gradient buffers). */
int eegx_layernorm_bwd_bf16(const void* dy, const void* x, const float* gamma, const float* beta,
                            const float* mean, const float* rstd, void* dx, float* dgamma, float* dbeta,
                            int accumulate, void* workspace, size_t workspace_bytes, int64_t rows, int64_t C,
                            int64_t groups, int64_t param_stride, int act, const uint64_t* rng_state, uint32_t site,
                            float p, void* stream);
/* Token assembly in front of the attention stack (layers.py:214-225), G groups of B sequences:
 * out[gb, s, :] = (s == 0 ? cls[g] : s < 4 ? temporal[g][s-1] : h[gb, s-4]) + pos[g][s]; h (G*B*T, d), out
 * (G*B*(T+4), d) bf16; cls (d), temporal (3, d), pos (T+4, d) fp32 per group at the given strides.
 * Backward: dh[gb, t] = dout[gb, t+4]; d(pos) is eegx_colsum_bf16 of dout viewed (G, B, (T+4)*d). */
int eegx_assemble_tokens_fwd_bf16(const void* h, const float* cls, int64_t cls_gstride, const float* temporal,
                                  int64_t temporal_gstride, const float* pos, int64_t pos_gstride, void* out, int64_t G,
                                  int64_t B, int64_t T, int64_t d, void* stream);
int eegx_assemble_tokens_bwd_bf16(const void* dout, void* dh, int64_t GB, int64_t T, int64_t d, void* stream);

/* out = a + scale * dropout(b): the residual adds of layers.py:234, 242, 251 / brain_encoder.py:165. */
int eegx_add_dropout_fwd_bf16(const void* a, const void* b, void* out, int64_t n, float scale,
                              const uint64_t* rng_state, uint32_t site, float p, void* stream);
/* out = scale * dropout_mask * in (the backward of the `b` operand above; also a plain dropout). */
int eegx_dropout_scale_bf16(const void* in, void* out, int64_t n, float scale, const uint64_t* rng_state,
                            uint32_t site, float p, void* stream);
/* out = dropout(gelu(x)) (nn.GELU + nn.Dropout after a Linear, brain_encoder.py:36-75) and its backward. */
int eegx_gelu_dropout_fwd_bf16(const void* x, void* out, int64_t n, const uint64_t* rng_state, uint32_t site,
                               float p, void* stream);
int eegx_gelu_dropout_bwd_bf16(const void* dout, const void* x, void* dx, int64_t n, const uint64_t* rng_state,
                               uint32_t site, float p, void* stream);
/* FeedForwardNetwork gate (layers.py:311-317): ag = [W1 x | Wg x] (rows, 2H) -> dropout(gelu(a) * sigmoid(g)). */
int eegx_glu_fwd_bf16(const void* ag, void* out, int64_t rows, int64_t H, const uint64_t* rng_state,
                      uint32_t site, float p, void* stream);
int eegx_glu_bwd_bf16(const void* dout, const void* ag, void* dag, int64_t rows, int64_t H,
                      const uint64_t* rng_state, uint32_t site, float p, void* stream);

/* BatchNorm1d statistics over the valid rows of a conv output (layers.py:31-48; train mode):
 * mean / rstd (C) fp32; running_mean / running_var updated in place with `momentum` (unbiased
 * variance) unless NULL.  Two fixed-order stages (bit-stable). */
size_t eegx_colreduce_workspace_bytes(int64_t C);
/* out[g][c] (+)= sum_r y[g * y_gstride + r * ld + c] of G (rows, C) bf16 matrices with row pitch ld, fp32, fixed
 * order: the bias gradients of nn.Linear / nn.Conv1d (accumulate != 0: added into the gradient buffer); out of
 * group g at + g * out_gstride. */
int eegx_colsum_bf16(const void* y, int64_t ld, int64_t G, int64_t y_gstride, int64_t rows, int64_t C, float* out,
                     int64_t out_gstride, int accumulate, void* workspace, size_t workspace_bytes, void* stream);
/* dst[g * dst_gstride + i] (+)= sum_s part[(s * G + g) * n + i]: folds the split-K partials of G weight-gradient
 * GEMMs into their gradients (G = 1: dst contiguous). */
int eegx_accumulate_partials_f32(const float* part, int64_t s, int64_t G, int64_t n, float* dst, int64_t dst_gstride,
                                 int accumulate, void* stream);
/* nn.Conv1d weight gradient: part is s partials of the GEMM layout (Cout, k * Cin) [index tap * Cin + ci];
 * dst, laid out like the parameter (Cout, Cin, k), (+)= their sum. */
int eegx_accumulate_conv_wgrad_f32(const float* part, int64_t s, int64_t Cout, int64_t Cin, int64_t k, float* dst,
                                   int accumulate, void* stream);
/* G groups of B trials: mean / rstd (G, C); running statistics of group g at + g * running_gstride. */
int eegx_bn_stats_bf16(const void* y, int64_t G, int64_t B, int64_t T, int64_t pad, int64_t C, float eps, float* mean,
                       float* rstd, float* running_mean, float* running_var, int64_t running_gstride, float momentum,
                       void* workspace, size_t workspace_bytes, void* stream);
/* out = zero_pad_rows( dropout( gelu( bn_a(ya) + residual ) ) )  (layers.py:142-174)
 * res_mode 0: none; 1: identity residual yr; 2: bn_r(yr) (the 1x1-conv + BatchNorm residual). */
int eegx_bn_act_fwd_bf16(const void* ya, const float* mean_a, const float* rstd_a, const float* gamma_a,
                         const float* beta_a, const void* yr, const float* mean_r, const float* rstd_r,
                         const float* gamma_r, const float* beta_r, int res_mode, void* out, int64_t G, int64_t B,
                         int64_t T, int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p, void* stream);
/* Backward of the above.  sums (G, 3, C) fp32: [0] dbeta (both sides), [1] dgamma_a, [2] dgamma_r.  Statistics and
 * affine vectors are (G, C).
 * da / dr: gradients w.r.t. ya / yr as guarded rows.  train = 0: statistics are constants (eval). */
int eegx_bn_act_bwd_bf16(const void* dout, const void* ya, const float* mean_a, const float* rstd_a,
                         const float* gamma_a, const float* beta_a, const void* yr, const float* mean_r,
                         const float* rstd_r, const float* gamma_r, const float* beta_r, int res_mode, int train,
                         void* da, void* dr, float* sums, void* workspace, size_t workspace_bytes, int64_t G, int64_t B,
                         int64_t T, int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p,
                         void* stream);
/* Depthwise Conv1d k = 5, groups = C (layers.py:157) on guarded rows; w (G, C, 5), bias (G, C) fp32.
 * Backward: dx guarded rows; dwdb (G, 6, C) fp32 = [dw tap 0..4, dbias]. */
int eegx_dwconv5_fwd_bf16(const void* x, const float* w, const float* bias, void* out, int64_t G, int64_t B, int64_t T,
                          int64_t pad, int64_t C, void* stream);
int eegx_dwconv5_bwd_bf16(const void* dout, const void* x, const float* w, void* dx, float* dwdb,
                          void* workspace, size_t workspace_bytes, int64_t G, int64_t B, int64_t T, int64_t pad,
                          int64_t C, void* stream);
/* SqueezeExciteBlock (layers.py:288-298): s = mean_t x (B, C) fp32; out = dropout(x * e[b, c]) written as
 * compact (B*T, C) rows; backward: dx (valid guarded rows) and de (B, C). */
int eegx_group_mean_bf16(const void* x, float* s, int64_t B, int64_t T, int64_t pad, int64_t C, void* stream);
int eegx_group_mean_bwd_bf16(const float* ds, void* dx, int64_t B, int64_t T, int64_t pad, int64_t C,
                             int accumulate, void* stream);
int eegx_se_scale_fwd_bf16(const void* x, const float* e, void* out, int64_t B, int64_t T, int64_t pad, int64_t C,
                           const uint64_t* rng_state, uint32_t site, float p, void* stream);
int eegx_se_scale_bwd_bf16(const void* dout, const void* x, const float* e, void* dx, float* de, int64_t B,
                           int64_t T, int64_t pad, int64_t C, const uint64_t* rng_state, uint32_t site, float p,
                           void* stream);
/* (B, C, T) fp32 (batch stride x_bstride elements) -> guarded channels-last bf16 rows: the layout change in
 * front of conv1 (layers.py:142 consumes (B, C, T)). */
int eegx_nct_to_rows_bf16(const float* x, int64_t x_bstride, void* out, int64_t B, int64_t T, int64_t pad,
                          int64_t C, void* stream);

/* ------------------------------------------------------------------------
 * Attention core of nn.MultiheadAttention for short sequences (S_q, S_k <= 64; head_dim in
 * {64, 96, 128, 192}): softmax(q k^T * scale) -> dropout -> (.) v, one CTA per (batch, head),
 * probabilities never leave the SM (layers.py:232-234, 245-251; brain_encoder.py:165-168).
 * q, k, v, o: rows of (B*S, row_stride) bf16 matrices, head h at columns [h*hd, (h+1)*hd) --
 * the packed in-projection output is consumed in place.  lse: (B, H, S_q) fp32 (saved).
 * Backward recomputes the probabilities; d_o has o's row stride.
 * ------------------------------------------------------------------------ */
typedef struct eegx_attn_desc {
    int64_t B, H, Sq, Sk, hd;
    int64_t q_rs, k_rs, v_rs, o_rs; /* row strides in elements (multiples of 8) */
    int32_t causal;
    float scale;
} eegx_attn_desc;
int eegx_attn_fwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, void* o, float* lse,
                       const uint64_t* rng_state, uint32_t site, float p, void* stream);
int eegx_attn_bwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, const void* o,
                       const void* d_o, const float* lse, void* dq, void* dk, void* dv, int64_t dq_rs, int64_t dk_rs,
                       int64_t dv_rs, const uint64_t* rng_state, uint32_t site, float p, void* stream);

/* The same attention core for LONG sequences (any S_q, S_k, no causal mask): flash style -- keys / values stream
 * through shared memory in tiles of 64, online softmax, the S_q x S_k scores and probabilities never exist in HBM
 * (main_model/src/models/layers.py:230-251 at the reference's real shapes, S = T + 4 = 1655 / 2052 / 4100, where
 * nn.MultiheadAttention materialises B*H*S*S scores).  Same descriptor and addressing as above.  Backward
 * recomputes the probabilities from lse; `dsum` is a caller-owned (B, H, S_q) fp32 workspace (D_i = dO_i . O_i,
 * written by the dQ kernel, read by the dK/dV kernel).  Every gradient has one owner CTA: bit-reproducible. */
int eegx_attn_flash_fwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, void* o, float* lse,
                             const uint64_t* rng_state, uint32_t site, float p, void* stream);
int eegx_attn_flash_bwd_bf16(const eegx_attn_desc* d, const void* q, const void* k, const void* v, const void* o,
                             const void* d_o, const float* lse, float* dsum, void* dq, void* dk, void* dv, int64_t dq_rs,
                             int64_t dk_rs, int64_t dv_rs, const uint64_t* rng_state, uint32_t site, float p,
                             void* stream);

/* ------------------------------------------------------------------------
 * Cross-entropy at the loss end of the train step (main_model/src/models/bart_decoder.py:41-48 ->
 * BartForConditionalGeneration: lm_head + final_logits_bias + CrossEntropyLoss(ignore_index)).
 * logits: (rows, ld) bf16 with V <= ld valid columns (the LM-head GEMM output, bias fused).
 *   fwd: lse[r] = logsumexp(logits[r, :V]); loss_rows[r] = lse[r] - logits[r, labels[r]] (0 if ignored)
 *   bwd: dlogits[r, v] = (exp(logits[r, v] - lse[r]) - [v == labels[r]]) * coef[0]  (0 on ignored rows and
 *        on the padding columns [V, ld)); coef: device scalar = dloss / n_valid.
 * ------------------------------------------------------------------------ */
int eegx_ce_fwd_bf16(const void* logits, int64_t ld, const int64_t* labels, int64_t rows, int64_t V,
                     int64_t ignore_index, float* loss_rows, float* lse, void* stream);
int eegx_ce_bwd_bf16(const void* logits, int64_t ld, const int64_t* labels, const float* lse, const float* coef,
                     void* dlogits, int64_t rows, int64_t V, int64_t ignore_index, void* stream);


/* ------------------------------------------------------------------------
 * Data-side rows of the path that the reference runs in numpy / sklearn on the host.
 *
 * eegx_robust_fit_f32: RobustScaler(quantile_range=(q_lo, q_hi)).fit of
 *   EEGDataset._initialize_scalers_efficiently (main_model/src/data/dataset.py:102-151).
 *   x: (C, n) fp32, row c = every fit sample of channel c concatenated over time (nan_to_num applied).
 *   center[c] = median, scale[c] = P(q_hi) - P(q_lo) with numpy's linear-interpolated percentiles, zero
 *   scales replaced by 1.  Exact (radix select), deterministic.
 * eegx_region_std_f32: population standard deviation of each row of a (B, n) matrix (np.std of a whole
 *   region, dataset.py:240).
 * eegx_augment_f32: EEGDataset._augment_eeg_regions (dataset.py:227-261) for a batch of one region:
 *   out[b, c, t] = scale[b] * (x[b, c, (t - shift[b]) mod T] + sigma[b] * N(0, 1)); sigma[b] = 0 disables the
 *   noise, scale[b] = 1 the scaling, shift[b] = 0 the roll (the caller draws the per-trial decisions).
 *   Noise: Philox(seed, step, site, source index) + Box-Muller.  out must not alias x.
 * ------------------------------------------------------------------------ */
int eegx_robust_fit_f32(const float* x, int64_t C, int64_t n, float q_lo, float q_hi, float* center, float* scale,
                        void* stream);
int eegx_region_std_f32(const float* x, int64_t B, int64_t n, float* out, void* stream);
int eegx_augment_f32(const float* x, float* out, int64_t B, int64_t C, int64_t T, const float* sigma,
                     const float* scale, const int32_t* shift, const uint64_t* rng_state, uint32_t site, void* stream);

/* ------------------------------------------------------------------------
 * wake_model dense head (BASELINE config 5): Linear(in, hidden, act) -> Linear(hidden, n_cls, softmax) ->
 * categorical cross-entropy, per-sample SGD in fp64, n samples IN ORDER in one persistent cooperative launch.
 *
 * Replaces wake_model/layers/linear.cpp:5-44 (Linear::forward), :47-72 (Linear::backward with the in-place SGD
 * update), layers/activations.h:29-41,64-95, layers/losses.h:8-22 and the Linear part of the loop
 * wake_model/train.cpp:68-117.
 *   w1 (hidden, in), b1 (hidden), w2 (n_cls, hidden), b2 (n_cls): fp64 device buffers, updated in place (train = 1).
 *   x (n, in) fp64, label (n) int32 class index (train.cpp:102 one-hot position).
 *   activation: 0 none, 1 relu, 2 sigmoid, 3 tanh (hidden layer; derivative taken at the layer OUTPUT as the
 *   reference does).  train = 0: forward only (loss / probs), no updates.
 *   loss (n), probs (n, n_cls), dx (n, in; = what Linear::backward of the hidden layer returns): nullable outputs.
 *   Needs about 2*in + n_cls*(ceil(hidden/SMs) + 3) doubles of shared memory (<= 227 KB), else EEGX_ERR_SHAPE.
 * Tolerance vs the reference: the SGD update itself is rounded exactly as the C++ (mul, mul, sub); dot products
 * are summed in a different fixed order -> parameters agree to ~1e-10 relative after tens of samples.
 * ------------------------------------------------------------------------ */
size_t eegx_wake_dense_workspace_bytes(int64_t in, int64_t hidden, int64_t n_cls, int want_dx);
int eegx_wake_dense_f64(double* w1, double* b1, double* w2, double* b2, const double* x, const int32_t* label,
                        int64_t n, int64_t in, int64_t hidden, int64_t n_cls, double lr, int activation, int train,
                        double* loss, double* probs, double* dx, void* workspace, size_t workspace_bytes,
                        void* stream);

/* ------------------------------------------------------------------------
 * wake_model convolution / max-pool front (wake_model/train.cpp:26-33), fp64, one sample per call.
 *
 * eegx_wake_conv2d_f64 replaces Convolution::forward (wake_model/layers/convolution.cpp:36-57) and, when dout is
 * given, Convolution::backward (:60-112) including its in-place SGD update of kernel and bias:
 *   x (H, W), kernel (kh, kw), bias (1), y and dout (H-kh+1, W-kw+1), dx (H, W): row-major fp64 device buffers.
 *   y = valid cross-correlation + bias (the constructor's activation is never applied by the reference).
 *   dout != NULL: dx = the layer's "input gradient" as written (flipped kernel at input position (y+ky, x+kx)),
 *   computed with the kernel BEFORE the update; then kernel -= lr * kernel_gradient, bias -= lr * sum(dout).
 *   y may be NULL (backward only).
 * eegx_wake_maxpool_f64 replaces MaxPool::forward (wake_model/layers/maxpool.cpp:6-43) and ::backward (:46-69):
 *   y (OH, OW) with OH = (H-pool_h)/stride+1, OW = (W-pool_w)/stride+1; argmax (OH, OW, 2) int32 = the recorded
 *   (row, col) of the first strict maximum, (-1, -1) if the window saw no element (the reference bounds rows by a
 *   member that maxpool.h:15 sets to the input WIDTH); dx (H, W) = scatter-add of dout to the maxima.
 * Every result is bit-exact against the reference's compiled classes (tests/test_wake.py): one thread owns one
 * output element and adds its terms in the reference's loop order without FMA contraction.
 * ------------------------------------------------------------------------ */
int eegx_wake_conv2d_f64(double* kernel, double* bias, const double* x, int64_t H, int64_t W, int64_t kh, int64_t kw,
                         const double* dout, double lr, double* y, double* dx, void* stream);
int eegx_wake_maxpool_f64(const double* x, int64_t H, int64_t W, int64_t pool_w, int64_t pool_h, int64_t stride,
                          const double* dout, double* y, int32_t* argmax, double* dx, void* stream);

/* ------------------------------------------------------------------------
 * Beam-search step: log-softmax + top-k of every logits row in one read (generation.py).
 * Replaces the per-step log_softmax / add / torch.topk over (batch*beams, V) fp32 logits of transformers'
 * GenerationMixin._beam_search (third-party; called from main_model/src/models/bart_decoder.py:59-79).
 *   logits (rows, ld) fp32, V <= ld valid columns; k <= 16.
 *   out_val[r, j] = logits[r, idx_j] - logsumexp(logits[r, :V]), j-th largest (ties: lower index first);
 *   out_idx[r, j] = idx_j.  banned >= 0: token excluded from the selection, not from the log-sum-exp
 *   (MinLengthLogitsProcessor semantics); pass -1 for none.
 * ------------------------------------------------------------------------ */
int eegx_logsoftmax_topk_f32(const float* logits, int64_t ld, int64_t rows, int64_t V, int32_t k, int64_t banned,
                             float* out_val, int64_t* out_idx, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EEGX_H */
