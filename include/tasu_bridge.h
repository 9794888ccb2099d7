/*
 * tasu_bridge.h — C ABI of libtasu_bridge.so: the B200 (sm_100a) implementation of the
 * TASU speech→LLM bridge hot path of PigeonDan1/ps-slm.
 *
 * The reference is pure Python/PyTorch and has NO FFI; these entry points are what a
 * ctypes binding placed in the reference's own methods would call (INTEGRATION.md shows
 * the stubs).  Every function cites the reference lines (relative to /root/reference/)
 * whose work it replaces.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - strides are in ELEMENTS, sizes in elements/rows; dtype codes are TASU_F32/TASU_BF16;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); nothing
 *    synchronises the device, allocates or frees — the caller owns every buffer;
 *  - return value: TASU_OK (0) or a negative TASU_ERR_*; tasu_last_error() gives the
 *    thread-local message.  Data-dependent error conditions (the two ValueErrors of
 *    ps-slm.py:783-785 and :861-865) are reported through the splice header words so the
 *    host can raise after its single read-back.
 */
#ifndef TASU_BRIDGE_H_
#define TASU_BRIDGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TASU_ABI_VERSION 2

enum { TASU_OK = 0, TASU_ERR_INVALID_ARG = -1, TASU_ERR_CUDA = -2, TASU_ERR_UNSUPPORTED = -3 };
enum { TASU_F32 = 0, TASU_BF16 = 1 };

/* frame-stats / collapse input kinds */
enum { TASU_INPUT_PROBS = 0,  /* probabilities or log-probabilities; auto-detected like ps-slm.py:256 */
       TASU_INPUT_LOGITS = 1  /* raw ctc_lo logits; softmax is fused (replaces ps-slm.py:451,582) */ };

/* GEMM epilogues */
enum { TASU_EPI_NONE = 0, TASU_EPI_BIAS = 1, TASU_EPI_BIAS_SILU = 2, TASU_EPI_BIAS_RELU = 3,
       TASU_EPI_LNFOLD_SILU = 4, /* silu(rstd[m]*(acc - mean[m]*colsum[n]) + bias[n]) */
       TASU_EPI_LNFOLD = 5,      /* the same without the SiLU (pre-activation kept for training) */
       TASU_EPI_SOFTMAX = 6      /* exp(acc + bias[n] - row_mean[m]) * row_rstd[m]: softmax with known row max / 1/sum */ };

/* splice header words (int64) written by tasu_splice_header */
enum { TASU_SH_SPLICED_LEN = 0,   /* S' = max_b sum(placeholders)            ps-slm.py:809 */
       TASU_SH_LEFT_PADDING = 1,  /* 1 = left padding                         ps-slm.py:771-785 */
       TASU_SH_ERR_BOTH_SIDES = 2,/* 1 → ValueError of ps-slm.py:783-785 */
       TASU_SH_TOTAL_SLOTS = 3,   /* number of audio slots found              ps-slm.py:861 */
       TASU_SH_TOTAL_AUDIO = 4,   /* sum(num_audio_tokens)                    ps-slm.py:861 */
       TASU_SH_N_SPEECH = 5,      /* number of <speech> tokens in input_ids */
       TASU_SH_WORDS = 8 };

/* collapse header words (int64) written by tasu_collapse_scan */
enum { TASU_CH_N_OUT = 0,         /* total compressed rows  sum_b M_b */
       TASU_CH_MAX_LEN = 1,       /* max_b M_b              ps-slm.py:303 */
       TASU_CH_IS_LOGPROB = 2,    /* 1 if input was detected as log-probs  ps-slm.py:256 */
       TASU_CH_KEPT_FRAMES = 3,   /* input frames feeding the kept rows (sum of kept run lengths) */
       TASU_CH_WORDS = 4 };

int tasu_abi_version(void);
const char* tasu_last_error(void);
/* Run-time options; the initial value comes from the environment variable of the same purpose when the library is first
 * used, else the default.
 *   TASU_OPT_GEMM_PAIR (env TASU_GEMM_PAIR, default 1): tasu_gemm_bf16_tn runs deep-K problems (K > 1024, M > 128) as CTA
 *   pairs — clusters of two CTAs, one tcgen05.mma.cta_group::2 of M = 256 per 256x256 tile, each CTA staging half of
 *   the B tile (a third less shared-memory fill per flop).  Same contract and bit-identical results as the one-CTA
 *   kernel (the accumulation order inside a tile is unchanged); 0 selects the one-CTA kernel. */
enum { TASU_OPT_GEMM_PAIR = 0, TASU_OPT_COUNT = 1 };
int tasu_set_option(int option, int value);
int tasu_get_option(int option);   /* value, or TASU_ERR_INVALID_ARG for an unknown option */
/* sm_count, compute capability of the current device */
int tasu_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host);

/* ---------------------------------------------------------------------------------------
 * Step 2a — per-frame greedy statistics over the vocab axis, one warp per frame.
 * Replaces torch.softmax (ps-slm.py:451,582), ctc_posterior.max() (:256) and
 * ctc_probs[b,:L].argmax(-1) (:265).
 *   x           [B, T, V] view: element (b,t,v) at x[b*batch_stride + t*row_stride + v]
 *   argmax      [B*T] int32   first index of the row maximum (torch tie rule)
 *   x_blank     [B*T] float   x[b,t,blank_id] as given (prob, log-prob or logit)
 *   row_max     [B*T] float
 *   row_sumexp  [B*T] float   sum_v exp(x - row_max); only for TASU_INPUT_LOGITS (else may be NULL)
 *   global_max_enc [1] uint32 order-preserving encoding of max over the WHOLE tensor;
 *               zero-initialised by this call (stream-ordered) before the kernel runs.
 *   lens        [B] int64 or NULL: with TASU_INPUT_LOGITS rows t >= lens[b] are skipped;
 *               with TASU_INPUT_PROBS every row is scanned (the reference's max is global).
 */
int tasu_frame_stats(const void* x, int dtype, int input_kind, int B, int T, int V,
                     int64_t batch_stride, int64_t row_stride, int blank_id, const int64_t* lens,
                     int32_t* argmax, float* x_blank, float* row_max, float* row_sumexp,
                     uint32_t* global_max_enc, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 2b — collapse plan: runs of equal greedy ids, blank frames kept one by one,
 * candidate score = mean blank probability, keep iff score < threshold (strict, fp32),
 * compaction by block scan.  Replaces the Python loop ps-slm.py:259-301.
 *   seg_start/seg_len/seg_score [B*T]  kept candidates of utterance b at [b*T, b*T+M_b)
 *   new_lens   [B] int64  M_b  (ps-slm.py:315)
 *   kept_frames [B] int32 or NULL: number of input frames covered by the kept candidates
 *   seg_frame_off [B*T] int32 or NULL: per kept candidate, offset of its first frame among the
 *               utterance's kept frames (compact layout of tasu_gather_kept_rows)
 */
int tasu_collapse_plan(const int32_t* argmax, const float* x_blank, const float* row_max,
                       const float* row_sumexp, const uint32_t* global_max_enc, int input_kind,
                       const int64_t* lens, int B, int T, int blank_id, float threshold,
                       int32_t* seg_start, int32_t* seg_len, float* seg_score, int64_t* new_lens,
                       int32_t* kept_frames, int32_t* seg_frame_off, void* stream);

/* Exact-decision mode of the fused CTC head.  tasu_ctc_head_stats computes its logits from bf16 operands, so a logit
 * differs from the fp32 one (ctc_lo of ps-slm.py:450, :581) by at most delta_f = ||x_f|| * max_v ||w_v|| * err_scale
 * (err_scale = 2^-8: both operands rounded to 8 significant bits; 2^-9 when the encoder rows were GIVEN in bf16).  The
 * greedy decisions of a frame — argmax (ps-slm.py:265) and the strict fp32 `score < threshold` test (:295-297) — can only
 * differ from the fp32 reference when their margin is below 2*delta_f.  tasu_flag_ambiguous_frames lists those frames:
 * top-2 logit gap (bounded from p1 = 1/row_sumexp and p2 <= min(1 - p1, sqrt(sum p^2 - p1^2))) below the margin, a blank
 * frame with |logit(p_blank) - logit(threshold)| below it, a non-blank frame whose blank probability could reach the
 * threshold.  The list has one slot per frame (no cap, nothing can be dropped): frame_idx / raw_row [B*T] int32,
 * count [1] int32 (zeroed by the call).  dec_max / dec_sum [B*T] receive a copy of row_max / row_sumexp: the normalisers
 * the collapse plan uses (tasu_collapse_plan), kept apart from the bf16-consistent ones pass 2 needs.
 *   x [B*(T+n_prefix), ldx] the encoder rows (fp32 or bf16), w_norm_max_enc [1] from tasu_row_norm_max.
 * tasu_ctc_head_refine recomputes the listed frames with fp32 FMAs from the fp32 weights (every dot product in ascending
 * k), for all `*count` frames, in-kernel, and writes argmax / x_blank / dec_max / dec_sum of those frames. */
int tasu_row_norm_max(const void* w, int dtype, int rows, int cols, int64_t ld,
                      uint32_t* out_enc /*[1] order-preserving encoding of max_r ||w_r||_2; zeroed by the call*/, void* stream);
int tasu_flag_ambiguous_frames(const int32_t* argmax, const float* x_blank, const float* row_max,
                               const float* row_sumexp, const float* row_sumexp2 /*or NULL*/, const int64_t* lens,
                               const void* x, int x_dtype, int64_t ldx, int K,
                               const float* x_sumsq /*[B*(T+n_prefix)] squared row norms (tasu_cast_rows_sumsq) or NULL: taken from x*/,
                               const uint32_t* w_norm_max_enc,
                               float err_scale, int B, int T, int n_prefix, int blank_id, float threshold,
                               float* dec_max, float* dec_sum, int32_t* frame_idx, int32_t* raw_row, int32_t* count,
                               void* stream);
int64_t tasu_ctc_head_refine_workspace(int64_t max_frames);
int tasu_ctc_head_refine(const void* x, int x_dtype, int64_t ldx, const float* w_f32, int64_t ldw, const float* bias,
                         int V, int K, int blank_id, const int32_t* frame_idx, const int32_t* raw_row,
                         const int32_t* count, int64_t max_frames /*list capacity = B*T*/, int32_t* argmax, float* x_blank,
                         float* dec_max, float* dec_sum, void* workspace, int64_t workspace_bytes, void* stream);

/* exclusive scans: new_lens → row_off [B+1] int32, kept_frames → frame_off [B+1] int32 (optional);
 * header [TASU_CH_WORDS] int64 */
int tasu_collapse_scan(const int64_t* new_lens, const int32_t* kept_frames, const uint32_t* global_max_enc,
                       int B, int32_t* row_off, int32_t* frame_off, int64_t* header,
                       int32_t* counts_dev /*[4] {N_out, max_len, kept_frames, 0} in device memory, or NULL*/,
                       void* stream);

/* tasu_collapse_plan + tasu_collapse_scan in ONE launch: the last CTA of the plan kernel to finish runs the scans and
 * writes the header.  `ticket` [1] int32 must be zero before the first use; the kernel hands it back zeroed, so one word
 * serves every later (stream-ordered) launch. */
int tasu_collapse_plan_scan(const int32_t* argmax, const float* x_blank, const float* row_max,
                            const float* row_sumexp, const uint32_t* global_max_enc, int input_kind,
                            const int64_t* lens, int B, int T, int blank_id, float threshold,
                            int32_t* seg_start, int32_t* seg_len, float* seg_score, int64_t* new_lens,
                            int32_t* kept_frames, int32_t* seg_frame_off, int32_t* row_off, int32_t* frame_off,
                            int64_t* header, int32_t* counts_dev, int32_t* ticket, void* stream);

/* Gather the encoder rows of the kept frames into a compact [F_kept, K] bf16 matrix together with their
 * softmax scalars, so that a second, ~3x smaller CTC-head GEMM (TASU_EPI_SOFTMAX) recomputes probabilities
 * only where PSD keeps them.  Layout: row r < N_out = FIRST frame of packed candidate r (so the GEMM writes
 * single-frame candidates — the majority — straight into their pooled row), rows >= N_out = the extra frames
 * of multi-frame runs in candidate order.  pk_len [N_out] = run length, tail_src [N_out] = compact row of a
 * candidate's second frame, multi_rows[0, *multi_count) = the packed rows with more than one frame (work list of
 * tasu_pool_tail, unordered).  With row_sumexp2 given, LayerNorm statistics of the single-frame rows are
 * emitted (mean = 1/V, rstd from sum p^2); multi-frame rows get theirs from tasu_pool_tail. */
int tasu_gather_kept_rows(const void* x_bf16, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                          const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_frame_off,
                          const int32_t* row_off, const int32_t* frame_off, const float* row_max,
                          const float* row_sumexp, const float* row_sumexp2, int64_t max_rows /*compact rows*/,
                          int64_t max_out /*packed rows*/, void* xg_bf16, int64_t ldg, float* g_max, float* g_inv_sum, int32_t* pk_len,
                          int32_t* tail_src, int32_t* multi_rows /*[N_out] or NULL*/, int32_t* multi_count /*[1]*/,
                          float* ln_mean, float* ln_rstd, float ln_eps, void* stream);
/* Natural-order index of the kept frames (fp32-accurate path: the kept frames' fp32 encoder rows are gathered, their
 * logits recomputed with the fp32-accurate GEMM and pooled by tasu_segment_meanpool with seg_src): compact row
 * frame_off[b] + seg_frame_off + f holds frame f of the candidate; frame_row [max_rows] = raw encoder row of every
 * compact row, seg_src [max_out] = first compact row of every packed candidate. */
int tasu_kept_frame_index(const int32_t* seg_start, const int32_t* seg_len, const int32_t* seg_frame_off,
                          const int32_t* row_off, const int32_t* frame_off, int B, int T, int n_prefix, int64_t max_rows,
                          int64_t max_out, int32_t* frame_row, int32_t* seg_src, void* stream);
/* In-place mean over the frames of every multi-frame candidate of the compact probability matrix
 * (ps-slm.py:286): probs[r] = (probs[r] + sum of its tail rows) / n, plus LayerNorm statistics.  max_rows = rows the
 * compact matrix holds: a candidate whose frames do not all lie below it is skipped (no access beyond the buffer). */
int tasu_pool_tail(void* probs_bf16, int64_t ld, int D, int64_t n_out, int64_t max_rows, const int32_t* pk_len,
                   const int32_t* tail_src, const int32_t* multi_rows, const int32_t* multi_count,
                   float* ln_mean, float* ln_rstd, float ln_eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 2b', grouped layout — the mean over the frames of a run (ps-slm.py:286) taken INSIDE the epilogue of the
 * kept-frame softmax GEMM: the per-frame probabilities of runs of 2-4 frames never reach HBM.  The frames of a run lie
 * on adjacent rows of the GEMM's A operand (= adjacent TMEM lanes = adjacent threads of one epilogue warp), runs of
 * one size class fill whole 128-row tiles:
 *   S  (1 frame)     1 row per candidate, A rows [0, NS);  pooled row = A row
 *   G2 (2 frames)    2 rows per candidate from A row A2;   pooled rows from A2 (a tile stores 64 pooled rows)
 *   G4 (3-4 frames)  4 rows per candidate from A row A4 (3-frame runs carry a zero row of weight 0); pooled rows from O4
 *   X  (> 4 frames)  first frame among the S rows, extra frames from A row AX → per-frame probabilities from output row
 *                    OX, averaged by tasu_pool_tail as in the plain layout (pk_len / tail_src / multi_rows by pooled row)
 * A2, A4, AX are multiples of 128; filler rows are zero rows of weight 0.  Pooled rows [0, OX) are what the projector
 * consumes (holes included: at most 127 + 63 + 31 rows); perm[r] = pooled row of packed candidate r.
 * Layout words (device int32 [TASU_GL_WORDS], written by tasu_group_plan): */
enum { TASU_GL_A2 = 0, TASU_GL_A4 = 1, TASU_GL_AX = 2,
       TASU_GL_A_ROWS = 3,        /* A rows in total = live M of the GEMM */
       TASU_GL_O4 = 4, TASU_GL_OX = 5,
       TASU_GL_N2 = 6, TASU_GL_N4 = 7, TASU_GL_NS = 8 /* S and X candidates */, TASU_GL_NXE = 9 /* extra frames of X */,
       TASU_GL_WORDS = 16 };
/* slot / xoff [B*T] (indexed like seg_len): index of a candidate inside its class within its utterance / its first extra
 * frame; cnt / base [4*B]: per-utterance class counts and their exclusive scans; ticket: int32 zero-initialised once. */
int tasu_group_plan(const int32_t* seg_len, const int64_t* new_lens, int B, int T, int32_t* slot, int32_t* xoff,
                    int32_t* cnt, int32_t* base, int32_t* lay, int32_t* ticket, void* stream);
/* Copies the kept frames' encoder rows into the grouped A matrix xg [max_a, ldg] with their softmax scalars
 * (g_inv = 1 / (sum exp * frames averaged in the epilogue); 0 marks a row of weight 0), writes perm [max_out], the
 * LayerNorm statistics of single-frame rows and pool_tail's work list for class X.  max_o = rows of the pooled matrix
 * and of pk_len / tail_src / ln_mean / ln_rstd, max_proj = pooled rows the projector's buffers hold (perm = -1 for a
 * candidate beyond it and for entries [N_out, max_out): "zero row" of tasu_gather_rows).  Capacities bound every write. */
int tasu_gather_kept_rows_grouped(const void* x_bf16, int64_t ldx, int B, int T, int n_prefix, int K, int V,
                                  const int32_t* seg_start, const int32_t* seg_len, const int32_t* row_off,
                                  const int32_t* slot, const int32_t* xoff, const int32_t* base, const int32_t* lay,
                                  const float* row_max, const float* row_sumexp, const float* row_sumexp2,
                                  int64_t max_a, int64_t max_o, int64_t max_out, int64_t max_proj, void* xg_bf16, int64_t ldg,
                                  float* g_max, float* g_inv, int32_t* perm, int32_t* pk_len, int32_t* tail_src,
                                  int32_t* multi_rows, int32_t* multi_count, float* ln_mean, float* ln_rstd,
                                  float ln_eps, void* stream);
/* The kept-frame softmax GEMM on the grouped layout: C = pooled probabilities (bf16) [c_rows, ldc]; A [a_rows, lda] bf16,
 * B = W_ctc [N, K] bf16, K <= 1024.  row_inv / row_max [a_rows] as written by tasu_gather_kept_rows_grouped.
 * q_part [n_parts][ldq] fp32 with n_parts = tasu_gemm_softmax_grouped_parts(N): partial sums of p^2 of the pooled rows of
 * the G regions (row index relative to A2), one per (column tile, epilogue group), every element written exactly once
 * (no atomics) — tasu_group_ln_finish adds them in a fixed order and writes the LayerNorm statistics of those rows. */
int tasu_gemm_softmax_grouped_parts(int N);
int tasu_gemm_softmax_grouped(const void* A, int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, int a_rows,
                              int c_rows, int N, int K, const float* bias, const float* row_inv, const float* row_max,
                              const int32_t* lay, float* q_part, int64_t ldq, void* stream);
int tasu_group_ln_finish(const float* q_part, int64_t ldq, int n_parts, const int32_t* lay, int V, int64_t max_o,
                         float* ln_mean, float* ln_rstd, float ln_eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 2c — segmented mean-pool of the kept candidates (ps-slm.py:275-287, :290, :297,
 * :303-314).  feats is a [B, T, D] view (same tensor as the posterior on the default path).
 *   softmax_max/softmax_sumexp: NULL → pool feats as given; else feats are logits and
 *               exp(x - max[b,t]) / sumexp[b,t] is pooled (fused softmax).
 *   seg_src: NULL → candidate (b,j) starts at feats[b, seg_start, :]; else feats is a compact [F_kept, D] matrix
 *               in natural order (tasu_kept_frame_index), packed row r starts at row seg_src[r] and the softmax
 *               statistics are indexed by compact row as well (layout 0 only).
 *   layout 0 (packed):  out row r in [0, N_out) at out + r*out_row_stride      (N_out = row_off[B])
 *   layout 1 (padded):  out row (b, j<max_len) at out + (b*max_len + j)*out_row_stride, rows
 *               j >= M_b are zero-filled (ps-slm.py:308-314)
 *   ln_mean/ln_rstd [rows] optional LayerNorm statistics of every pooled row (fp32, biased
 *               variance, eps) consumed by TASU_EPI_LNFOLD_SILU (projector.py:139,150).
 */
int tasu_segment_meanpool(const void* feats, int in_dtype, int B, int T, int D,
                          int64_t batch_stride, int64_t row_stride,
                          const float* softmax_max, const float* softmax_sumexp,
                          const int32_t* seg_start, const int32_t* seg_len, const int32_t* row_off,
                          const int32_t* seg_src, int layout, int64_t max_len, int64_t max_rows,
                          void* out, int out_dtype, int64_t out_row_stride,
                          float* ln_mean, float* ln_rstd, float ln_eps, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 1a — simulated posterior rows (ps-slm.py:346-358 clean, :380-408 noisy).  The random
 * decisions are drawn on the host in the reference's order; each output row r is
 *   tok[r] <  0 : all zeros (padding, :403-408)
 *   else        : base[r] everywhere, hot[r] at column tok[r]
 * written to out + dst_row[r]*out_row_stride (dst_row NULL → r).  Optional LayerNorm stats.
 */
int tasu_sim_posterior_rows(const int32_t* tok, const float* hot, const float* base,
                            const int64_t* dst_row, int64_t n_rows, int V,
                            void* out, int out_dtype, int64_t out_row_stride,
                            float* ln_mean, float* ln_rstd, float ln_eps, void* stream);

/* fp32 -> bf16 row cast (pad columns of the pitch zero-filled) that also emits the squared 2-norm of every fp32 row:
 * the per-frame ||x_f|| of the exact-decision error bound, taken where the encoder output is read anyway. */
int tasu_cast_rows_sumsq(const float* src, int64_t rows, int cols, int64_t src_stride, void* dst_bf16, int64_t dst_stride,
                         float* row_sumsq, void* stream);

/* Content fingerprint of up to 8 device buffers (4096 evenly spaced 32-bit words of each, position-dependent hash):
 * the host mirror validates its cached bf16 / folded weight copies with it, because optimizers that update parameters
 * through flat buffers (DeepSpeed ZeRO, finetune_deepspeed.py:147-149) change neither a tensor's address nor torch's
 * version counter.  `out` [1] may be device or pinned host memory (plain store by one thread). */
int tasu_fingerprint(const void* const* ptrs_host, const int64_t* nbytes_host, int n, uint64_t* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 3 helpers — operand preparation for the tensor-core GEMMs.
 * tasu_cast_rows: [rows, cols] src → dst (dtype conversion, arbitrary row strides) with optional
 *   per-row LayerNorm statistics (the generic A-operand producer for projector.py:150).
 * tasu_fold_layernorm: W1g[n,k] = bf16(W1[n,k]*gamma[k]); colsum[n] = sum_k W1g[n,k];
 *   dbias[n] = sum_k W1[n,k]*beta[k] + b1[n]  — folds nn.LayerNorm (projector.py:139) into
 *   nn.Linear (projector.py:141) so GEMM-1 runs directly on the pooled posterior.
 */
int tasu_cast_rows(const void* src, int src_dtype, int64_t rows, int cols, int64_t src_stride,
                   void* dst, int dst_dtype, int64_t dst_stride,
                   float* ln_mean, float* ln_rstd, float ln_eps, void* stream);
int tasu_fold_layernorm(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                        const float* b1, int N, int K, void* w1g_bf16, int64_t w1g_stride,
                        float* colsum, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 3 — C[M,N] = epilogue(A[M,K] · B[N,K]^T), bf16 operands (both K-major), fp32
 * accumulation in TMEM, tcgen05.mma fed by TMA, persistent over 148 SMs.  Replaces
 * ctc_lo (ps-slm.py:450,581), nn.Linear(25055,2048)+SiLU and nn.Linear(2048,1536)
 * (projector.py:141-143), the Linear/ReLU of projector.py:35-37 and :16.
 *   lda/ldb/ldc in elements; A, B and C base pointers and row pitches must be 16-byte aligned.
 *   Stores are clipped at row M and at column N rounded up to the next 16-byte boundary of the
 *   row: pad columns inside the pitch may be written with zeros, nothing is written past it.
 *   bias [N] fp32 (EPI_BIAS*, LNFOLD), row_rstd/row_mean [M] and colsum [N] (LNFOLD only).
 *   m_dev (optional): device int32 holding the live row count; M is then the capacity the buffers
 *   were allocated for and the kernel works on min(*m_dev, M) rows — data-dependent sizes (compressed
 *   rows, kept frames) need no host synchronisation between the plan and the GEMMs.
 */
int tasu_gemm_bf16_tn(const void* A, int64_t lda, const void* B, int64_t ldb,
                      void* C, int c_dtype, int64_t ldc, int M, int N, int K, int epilogue,
                      const float* bias, const float* row_rstd, const float* row_mean,
                      const float* colsum, const int32_t* m_dev, void* stream);
/* ---------------------------------------------------------------------------------------
 * Steps 1b+2a fused — CTC head with the softmax statistics computed in the GEMM epilogue:
 * logits = X·W^T + b live only in TMEM; per frame the running max / sum-exp / argmax / blank logit
 * are kept in registers while the CTA sweeps the vocabulary.  Replaces ctc_lo + softmax
 * (ps-slm.py:450-451, :581-582), .max() (:256) and .argmax (:265) without materialising the
 * [B, T+P, V] tensor.  X is [B*(T+P), K] bf16 (P = n_prefix query frames, dropped like :452-454),
 * W is [V, K] bf16; outputs are the same four [B*T] arrays tasu_frame_stats(TASU_INPUT_LOGITS) gives,
 * plus (optional) row_sumexp2 = sum_v exp(2(x - row_max)), from which sum_v p_v^2 = row_sumexp2/row_sumexp^2
 * feeds the LayerNorm fold of single-frame rows without another pass.
 */
int64_t tasu_ctc_head_stats_workspace(int B, int T, int n_prefix);
int tasu_ctc_head_stats(const void* x_bf16, int64_t ldx, const void* w_bf16, int64_t ldw, const float* bias,
                        int B, int T, int n_prefix, int V, int K, int blank_id, int32_t* argmax,
                        float* x_blank, float* row_max, float* row_sumexp, float* row_sumexp2,
                        void* workspace, int64_t workspace_bytes, void* stream);

/* C[M,N] (fp32) = A · B with either operand in either storage order, no epilogue — the backward contractions of the
 * projector without materialised transposes:  a_mn_major = 0: A is [M, K] (K contiguous), 1: A is [K, M] (M contiguous);
 * b_mn_major = 0: B is [N, K] (K contiguous, nn.Linear layout), 1: B is [K, N] (N contiguous).  MN-major operands are
 * fetched as [64 K-rows][64 MN] TMA boxes and consumed through MN-major UMMA shared-memory descriptors.
 *   dW2 = dy^T · h : A = dy [rows, H] (a_mn_major = 1), B = h [rows, Hb] (b_mn_major = 1), K = rows
 *   dh  = dy · W2  : A = dy [rows, H] (0),              B = W2 [H, Hb]  (b_mn_major = 1), K = H            */
int tasu_gemm_bf16_f32(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                       float* C, int64_t ldc, int M, int N, int K, void* stream);

/* EXPERIMENTAL (not yet validated on a GPU, not called by the bridge unless asked to): tasu_gemm_bf16_tn with a
 * stream-K tail.  The full waves of 128x256 tiles run one tile per CTA as usual; the tiles of the ragged last wave are
 * cut along K into contiguous pieces dealt out evenly to the CTAs, partial accumulators go through `workspace`
 * (tasu_gemm_streamk_workspace() bytes, 256-byte aligned, ZERO-FILLED ONCE by the caller, then owned by the library; it
 * must not be shared by launches that can run concurrently) and are added in a fixed order by the CTA that holds the
 * tile's last K-block, which then runs the epilogue.  Same contract as tasu_gemm_bf16_tn otherwise; results are
 * deterministic, but split tiles may differ from the unsplit kernel in the last fp32 bits.  Launched cooperatively.
 * tasu_gemm_streamk_schedule_host: HOST — the pieces {tile, kb0, kb1, kind (0 full, 1 contributed, 2 finishing),
 * n_contrib} CTA `cta` computes after its tiles of the full waves, in processing order; returns their number (0..2). */
int64_t tasu_gemm_streamk_workspace(void);
int tasu_gemm_bf16_tn_streamk(const void* A, int64_t lda, const void* B, int64_t ldb,
                              void* C, int c_dtype, int64_t ldc, int M, int N, int K, int epilogue,
                              const float* bias, const float* row_rstd, const float* row_mean,
                              const float* colsum, const int32_t* m_dev, void* workspace, int64_t workspace_bytes,
                              void* stream);
int tasu_gemm_streamk_schedule_host(int num_tiles, int k_blocks, int grid, int cta, int32_t* pieces_host,
                                    int32_t* dp_tiles_host);

/* Cross-attention projector (projector.py:104-126; call site ps-slm.py:475-480), all heads in one launch:
 *   Z[:, h dp : (h+1) dp] = softmax(Q_h K_h^T) K_h,   Q_h = Q[:, h dp : (h+1) dp],  K_h = table[:, h dp : (h+1) dp]
 * with keys = values = the LLM embedding table (V2 rows).  The probabilities are produced tile by tile in shared memory
 * and consumed by the second contraction on the spot — the [N, V2] probability matrix of the composed path
 * (tasu_gemm_bf16_tn with TASU_EPI_SOFTMAX + tasu_gemm_bf16_f32: 2 x 3 GB of traffic per head at N = 9856) never exists.
 * row_max / row_inv [heads, stat_stride]: per head and row, max_k s and 1 / sum_k exp(s - max) of s = Q_h K_h^T
 * (tasu_ctc_head_stats on the head slices) — or both NULL: the kernel finds the row maxima itself in a first sweep over
 * the keys (scores only, no exponentials) and normalises O by the fp32 sum of the probabilities at the end.
 * Q pre-scaled by 1/sqrt(d).  dp in {64, 128, 192, 256}; Z fp32. */
int tasu_attn_softmax_pv(const void* Q_bf16, int64_t ldq, const void* table_bf16, int64_t ldt, int N, int V2,
                         int heads, int dp, const float* row_max, const float* row_inv, int64_t stat_stride,
                         float* Z, int64_t ldz, void* stream);
/* The same with a workspace for the KEY SPLIT of the self-contained mode (row_max = row_inv = NULL): the (128-row tile,
 * head) items — 616 at N = 9856, 8 heads: 4.16 waves on 148 SMs — are cut into `splits` key ranges each so that the last
 * wave is full; a split leaves its unnormalised output (relative to its own row maxima), the maxima and the sums in the
 * workspace and a merge kernel combines them in split order (deterministic).  tasu_attn_split_plan (HOST) returns the
 * workspace bytes (0 when one split is best) and the number of splits it chose; workspace = NULL or given statistics:
 * exactly tasu_attn_softmax_pv. */
int64_t tasu_attn_split_plan(int N, int V2, int heads, int dp, int* splits_host);
int tasu_attn_softmax_pv_ws(const void* Q_bf16, int64_t ldq, const void* table_bf16, int64_t ldt, int N, int V2,
                            int heads, int dp, const float* row_max, const float* row_inv, int64_t stat_stride,
                            float* Z, int64_t ldz, void* workspace, int64_t workspace_bytes, void* stream);

/* CUDA-core cross-check of the same contract (tests and bring-up only; never on the product path) */
int tasu_gemm_bf16_tn_simt(const void* A, int64_t lda, const void* B, int64_t ldb,
                           void* C, int c_dtype, int64_t ldc, int M, int N, int K, int epilogue,
                           const float* bias, const float* row_rstd, const float* row_mean,
                           const float* colsum, void* stream);

/* ---------------------------------------------------------------------------------------
 * Training side of the linear-silu projector: backward of projector.py:149-151 (autograd in the
 * reference; trained by Multitask/utils/deepspeed_utils.py:235-236).  The big contractions reuse
 * tasu_gemm_bf16_tn; these produce its K-major operands and the LayerNorm-fold algebra:
 *   tasu_transpose_cast  dst[c, r] = bf16(row_scale[r] * src[r, c])       (row_scale may be NULL)
 *   tasu_silu_fwd        h = bf16(silu(z))
 *   tasu_silu_bwd        dz = dh*silu'(z); dzsT[j, n] = bf16(rstd[n]*dz[n, j]); db1[j] = sum_n dz;
 *                        g0[j] = sum_n rstd[n]*mean[n]*dz[n, j]           (db1/g0 zeroed by the call)
 *   tasu_colsum          out[c] = sum_r src[r, c]                          (zeroed by the call)
 *   tasu_linear_silu_wgrad_finish   with G = dzsT · x  ([Hb, V], from tasu_gemm_bf16_tn):
 *                        dW1 = gamma*(G - g0) + beta*db1, dgamma = sum_j W1*(G - g0), dbeta = sum_j W1*db1
 */
/* Backward of the cross-attention projector (projector.py:104-126; trained by autograd in the reference): per head,
 * dS = P ∘ (dP − δ), δ_r = dZ[r,:]·Z[r,:] (= Σ_v P·dP), written as the bf16 K-major A operand of dQ_h = dS · K_h.
 * P = the head's probabilities (tasu_gemm_bf16_tn, EPI_SOFTMAX), dP = dZ_h · K_hᵀ (tasu_gemm_bf16_tn), dZ / Z = the
 * head's d-wide slices of the output gradient / output (row stride z_stride). */
int tasu_attn_score_grad(const void* P_bf16, int64_t p_stride, const float* dP, int64_t dp_stride,
                         const float* dZ, const float* Z, int64_t z_stride, int d, int64_t rows, int V,
                         void* dS_bf16, int64_t ds_stride, void* stream);
int tasu_transpose_cast(const void* src, int src_dtype, int64_t rows, int64_t cols, int64_t src_stride,
                        const float* row_scale, void* dst_bf16, int64_t dst_stride, void* stream);
int tasu_silu_fwd(const float* z, int64_t n, void* h_bf16, void* stream);
int tasu_silu_bwd(const float* dh, const float* z, int64_t N, int Hb, const float* row_rstd,
                  const float* row_mean, void* dzsT_bf16, int64_t t_stride, float* db1, float* g0,
                  void* stream);
int tasu_colsum(const void* src, int src_dtype, int64_t rows, int cols, int64_t src_stride, float* out,
                void* stream);
int tasu_linear_silu_wgrad_finish(const float* G, int64_t g_stride, const float* w1, int64_t w1_stride,
                                  const float* gamma, const float* beta, const float* g0, const float* db1,
                                  int Hb, int V,
                                  float* dw1, int64_t dw1_stride, float* dgamma, float* dbeta, void* stream);

/* fp32-accurate mode on the bf16 tensor cores (the reference computes the projector in fp32,
 * conf/ds_config.json:12-14): x = h1+h2+h3 (three bf16 terms); with A' = [h1|h1|h2|h1|h2|h3] (pattern 0) and
 * B' = [h1|h2|h1|h3|h2|h1] (pattern 1), blocks of pad64(K), ONE tasu_gemm_bf16_tn over K' = 6*pad64(K) gives the
 * fp32 product to ~1e-6.  col_scale (optional, [K]) multiplies columns first (LayerNorm gamma fold); optional
 * LayerNorm statistics (activations) / row sums (folded weights: the colsum of TASU_EPI_LNFOLD*). */
int tasu_split_bf16x3(const void* src, int src_dtype, int64_t rows, int K, int64_t src_stride,
                      const float* col_scale, int pattern, void* dst_bf16, int64_t dst_stride,
                      float* ln_mean, float* ln_rstd, float ln_eps, float* row_sum, void* stream);
/* split-K combine of that mode: out = epilogue(sum_p parts[p]) with round-to-nearest fp32 adds; parts[p] is the
 * TASU_EPI_NONE fp32 output of one K slice ([M, ldp], part_stride elements apart). */
int tasu_sum_epilogue(const float* parts, int n_parts, int64_t part_stride, int M, int N, int64_t ldp,
                      int epilogue, const float* bias, const float* row_rstd, const float* row_mean,
                      const float* colsum, void* out, int out_dtype, int64_t ldo, void* stream);

/* Row softmax with known statistics (tasu_frame_stats, TASU_INPUT_LOGITS): out[r, v] = bf16(exp(x[r,v] - row_max[r]) /
 * row_sumexp[r]), columns V..out_row_stride-1 zeroed — the K-major A operand of the vocabulary-transfer contraction
 * softmax(logits_no_blank) · embed_matrix (ps-slm.py:494-497, :509-511). */
int tasu_softmax_rows(const void* x, int x_dtype, int64_t x_row_stride, int64_t rows, int V, const float* row_max,
                      const float* row_sumexp, void* out_bf16, int64_t out_row_stride, void* stream);

/* ---------------------------------------------------------------------------------------
 * Token-row projector: steps 1a + 3 fused for TEXT-SIMULATED posteriors (ps-slm.py:337-358, :360-409 feeding
 * projector.py:149-151).  Every simulated row is  base*1 + (hot-base)*onehot(tok)  (clean: hot=1, base=0;
 * smoothed: hot=(1-alpha)+alpha/V, base=alpha/V; inserted blank: tok=blank, hot=1, base=0), so
 *   W1*LN(x) + b1 = a*gamma[tok]*W1[:,tok] + e*S + D,  a = rstd*(hot-base), e = rstd*(base-mean),
 *   S = W1*gamma, D = W1*beta + b1                      (closed-form mean/rstd of a two-valued row)
 * — a column gather of the fp32 W1 instead of a 2*V*Hb-flop GEMM row, and a column scatter in the backward:
 *   dW1[j,v] = gamma_v*(P[j,v] + E_j) + beta_v*db1_j,  P[j,v] = sum_{r: tok_r = v} a_r dz[r,j],  E_j = sum_r e_r dz[r,j]
 *   dgamma_v = sum_j W1[j,v]*(P[j,v] + E_j),  dbeta_v = sum_j W1[j,v]*db1_j,  db1_j = sum_r dz[r,j].
 * The [B, L, 25055] posterior (ps-slm.py:346, :403) is never built on this path.
 *   tasu_host_group_tokens   HOST: stable counting sort of the rows by token → uniq[n_uniq], seg_off[n_uniq+1],
 *                            perm[n_rows] (rows of token uniq[u] = perm[seg_off[u] .. seg_off[u+1]), ascending)
 *   tasu_linear_rowdots      S[j], D[j] (fp32), one pass over W1
 *   tasu_tokrow_fwd          z[r,:] (fp32, optional), h[r,:] = bf16(silu(z)), row_a[r], row_e[r]
 *   tasu_tokrow_bwd_rows     dz = dh*silu'(z) on the fly; P [n_uniq, Hb] (compact), db1, E — all sums in a fixed order
 *   tasu_tokrow_wgrad_finish dW1, dgamma, dbeta in one pass over W1; slot_ws = int32[V] scratch
 */
int tasu_host_group_tokens(const int32_t* tok_host, int64_t n_rows, int V, int32_t* uniq_host,
                           int32_t* seg_off_host, int32_t* perm_host, int32_t* n_uniq_host);
/* HOST: all of ctc_pseudo_posterior_noise's decisions (ps-slm.py:380-401, insert_prob = 0) from ONE uniform draw
 * u = torch.rand(sum(L_b) + B) — bit-identical to the reference's per-utterance uniform_()/rand(L_b) stream — written
 * as grouped row descriptors into a (pinned) staging buffer; layout in int32 words with cap = sum(L_b):
 * uniq[cap] | seg_off[cap+1] | perm[cap] | hot[cap] (f32) | base[cap] (f32) | pad to 8 B | lens[B] (int64). */
int tasu_host_sim_token_rows(const float* u_host, const int32_t* tok_in_host, const int64_t* len_in_host, int B,
                             int V, float drop_prob, float smooth_low, float smooth_high, int32_t* stage_host,
                             int64_t stage_words, int64_t* n_rows_host, int32_t* n_uniq_host);
int tasu_linear_rowdots(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                        const float* b1, int N, int K, float* S, float* D, void* stream);
int tasu_tokrow_fwd(const float* w1, int64_t w1_stride, const float* gamma, const float* S, const float* D,
                    const int32_t* uniq, const int32_t* seg_off, const int32_t* perm, const float* hot,
                    const float* base, int n_uniq, int64_t n_rows, int V, int Hb, float ln_eps, float* z,
                    void* h_bf16, float* row_a, float* row_e,
                    const float* colT /*optional [n_uniq, Hb] from tasu_tokrow_cols; NULL: gather W1 columns*/,
                    void* stream);
/* training forward (W1 changes every step): ONE dense pass over W1 gives S, D and the compact gamma-scaled
 * columns colT[u,:] = gamma[uniq[u]]*W1[:,uniq[u]] (every sector of W1 read once instead of one sector per element) */
int64_t tasu_tokrow_cols_workspace(int V, int Hb);
/* row pass on the compact columns: one warp per row, z / h / row_a / row_e as tasu_tokrow_fwd; row_slot_ws = int32[n_rows] */
int tasu_tokrow_rows_fwd(const float* colT, const float* S, const float* D, const int32_t* seg_off,
                         const int32_t* perm, const float* hot, const float* base, int n_uniq, int64_t n_rows,
                         int V, int Hb, float ln_eps, float* z, void* h_bf16, float* row_a, float* row_e,
                         int32_t* row_slot_ws, void* stream);
int tasu_tokrow_cols(const float* w1, int64_t w1_stride, const float* gamma, const float* beta, const float* b1,
                     const int32_t* uniq, int n_uniq, int V, int Hb, float* colT, float* S, float* D,
                     void* workspace, int64_t workspace_bytes, void* stream);
int64_t tasu_tokrow_bwd_workspace(int Hb, int n_uniq);    /* bytes of per-CTA partial sums (deterministic db1 / E) */
int tasu_tokrow_bwd_rows(const float* dh, const float* z, int64_t n_rows, int Hb, const int32_t* seg_off,
                         const int32_t* perm, const float* row_a, const float* row_e, int n_uniq, float* P,
                         float* db1, float* E, void* workspace, int64_t workspace_bytes, void* stream);
int tasu_tokrow_wgrad_finish(const float* P, const int32_t* uniq, int n_uniq, int32_t* slot_ws, const float* w1,
                             int64_t w1_stride, const float* gamma, const float* beta, const float* E,
                             const float* db1, int Hb, int V, float* dw1, int64_t dw1_stride, float* dgamma,
                             float* dbeta, void* stream);

/* Composite calls (one per direction of a text-only training step; the launches go out back to back):
 *   fwd: tasu_linear_rowdots → tasu_tokrow_fwd → bf16 cast of W2 → tasu_gemm_bf16_tn(EPI_BIAS)   → y [n_rows, H]
 *   bwd: db2, dW2 = dy^T·h, dh = dy·W2 (tensor cores) → tasu_tokrow_bwd_rows → tasu_tokrow_wgrad_finish
 * z / h_bf16 / row_a / row_e are the activations saved between the two calls; `workspace` (256-byte aligned,
 * tasu_tokrow_train_workspace bytes) is scratch and may be shared by both calls. */
int64_t tasu_tokrow_train_workspace(int64_t n_rows, int n_uniq, int V, int Hb, int H);
int tasu_tokrow_linear_silu_fwd(const float* w1, int64_t w1_stride, const float* gamma, const float* beta,
                                const float* b1, const float* w2, int64_t w2_stride, const float* b2,
                                const int32_t* uniq, const int32_t* seg_off, const int32_t* perm,
                                const float* hot, const float* base, int n_uniq, int64_t n_rows, int V, int Hb,
                                int H, float ln_eps, float* z, void* h_bf16, float* row_a, float* row_e,
                                void* y, int y_dtype, int64_t ldy, void* workspace, int64_t workspace_bytes,
                                void* stream);
int tasu_tokrow_linear_silu_bwd(const void* dy, int dy_dtype, int64_t ldy, const float* z, const void* h_bf16,
                                const float* row_a, const float* row_e, const float* w1, int64_t w1_stride,
                                const float* gamma, const float* beta, const float* w2, int64_t w2_stride,
                                const int32_t* uniq, const int32_t* seg_off, const int32_t* perm, int n_uniq,
                                int64_t n_rows, int V, int Hb, int H, float* dw1, int64_t dw1_stride,
                                float* dgamma, float* dbeta, float* db1, float* dw2, int64_t dw2_stride,
                                float* db2, int phase /* 0 = all; 1 = W1 half (dW1, dgamma, dbeta, db1) then 2 = W2 half
                                (dW2, db2): lets a data-parallel caller all-reduce dW1 under the W2 half */,
                                void* workspace, int64_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------
 * Step 4 — splice (ps-slm.py:765-871).  Integer plan, then one gather/scatter pass.
 *   input_ids [B,S] int64; attention_mask [B,S] uint8/bool (mask_dtype 0) or int64 (1);
 *   num_audio [n_audio] int64 (compressed lengths), divided by div_k on device
 *   (projector_feature_length = len // k, ps-slm.py:483).
 * tasu_splice_rowstat : per-row counts                                   (:771-772, :788-789)
 * tasu_splice_plan    : placeholders, cumsum, per-token slot ordinals    (:805-812, :842-859)
 * tasu_splice_header  : S', padding side, error words, per-row bases     (:809, :861)
 * (the caller reads the header — S', padding side, error words — and passes them back by value)
 * tasu_splice_scatter : ONE launch — per output row the integer outputs + the source of the row (row map, kept in
 *                       row_src_ws for the backward), then the copy of the row:
 *                       writes inputs_embeds / mask / labels / position_ids / final ids
 *                       (:821-840, :867-871); text rows come from `text_src`:
 *                       text_mode 0 = inputs_embeds [B,S,H]; 1 = embedding table indexed by
 *                       token id (fuses embed_tokens, ps-slm.py:525,654).
 *   audio rows: audio_layout 0 = packed [sum M, H]; 1 = padded [n_audio, audio_max_len, H].
 */
int tasu_splice_rowstat(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                        int B, int S, int64_t speech_id, int32_t* rowstat /*[B,8]*/, void* stream);
int tasu_splice_plan(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                     int B, int S, int64_t speech_id, const int64_t* num_audio, int n_audio,
                     int64_t div_k, int32_t* rowstat, int32_t* new_pos /*[B,S]*/,
                     int32_t* text_prefix /*[B,S]*/, int32_t* slot_ord /*[B,S]*/, void* stream);
int tasu_splice_header(const int32_t* rowstat, const int64_t* num_audio, int n_audio, int64_t div_k,
                       int B, int S, int64_t* header /*[TASU_SH_WORDS]*/, int32_t* slot_base /*[B]*/,
                       int32_t* audio_off /*[n_audio+1]*/, void* stream);
/* tasu_splice_plan + tasu_splice_header in ONE launch (the last CTA of the plan kernel to finish writes the header);
 * `ticket` [1] int32: zero before the first use, handed back zeroed. */
int tasu_splice_plan_header(const int64_t* input_ids, const void* attention_mask, int mask_dtype, int B, int S,
                            int64_t speech_id, const int64_t* num_audio, int n_audio, int64_t div_k,
                            int32_t* rowstat, int32_t* new_pos, int32_t* text_prefix, int32_t* slot_ord,
                            int64_t* header, int32_t* slot_base, int32_t* audio_off, int32_t* ticket, void* stream);
int tasu_splice_scatter(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                        const int64_t* labels, int B, int S, int spliced_len, int H, int64_t speech_id,
                        const void* text_src, int text_mode, int64_t text_row_stride,
                        const void* audio_rows, int audio_layout, int64_t audio_row_stride,
                        int64_t audio_max_len, int n_audio, int emb_dtype,
                        const int32_t* rowstat, const int32_t* new_pos, const int32_t* text_prefix,
                        const int32_t* slot_ord, const int32_t* slot_base, const int32_t* audio_off,
                        int left_padding, int64_t pad_id, int64_t ignore_id,
                        void* out_emb, void* out_mask, int64_t* out_labels, int64_t* out_pos,
                        int64_t* out_ids, int64_t* row_src_ws /*[B*S'] scratch: source of every output row*/,
                        int32_t* audio_dest /*optional: output row of every audio row; caller pre-fills -1*/,
                        void* stream);
/* The same with the audio rows stored in permuted order (grouped kept-frame layout, step 2b'): audio row r of the packed
 * layout (audio_layout 0) is read from row audio_perm[r] of audio_rows (< 0: zero row).  row_src_ws / audio_dest keep
 * referring to the packed row r, so the backward is unchanged.  audio_perm = NULL: tasu_splice_scatter. */
int tasu_splice_scatter_perm(const int64_t* input_ids, const void* attention_mask, int mask_dtype,
                             const int64_t* labels, int B, int S, int spliced_len, int H, int64_t speech_id,
                             const void* text_src, int text_mode, int64_t text_row_stride,
                             const void* audio_rows, const int32_t* audio_perm, int audio_layout,
                             int64_t audio_row_stride, int64_t audio_max_len, int n_audio, int emb_dtype,
                             const int32_t* rowstat, const int32_t* new_pos, const int32_t* text_prefix,
                             const int32_t* slot_ord, const int32_t* slot_base, const int32_t* audio_off,
                             int left_padding, int64_t pad_id, int64_t ignore_id,
                             void* out_emb, void* out_mask, int64_t* out_labels, int64_t* out_pos,
                             int64_t* out_ids, int64_t* row_src_ws, int32_t* audio_dest, void* stream);
/* dst[r,:] = idx[r] >= 0 ? src[idx[r],:] : 0.  With idx = audio_dest of tasu_splice_scatter and src = the gradient of
 * inputs_embeds this is the backward of the audio part of the splice (index_put of ps-slm.py:867-869; training). */
int tasu_gather_rows(const void* src, int dtype, int64_t src_row_stride, const int32_t* idx, int64_t n_rows, int H,
                     void* dst, int64_t dst_row_stride, void* stream);

/* Backward of the text part of the splice (ps-slm.py:833-834 index_put of inputs_embeds): grad_text [text_rows, H]
 * (zero-filled by the call) receives grad_emb[r] at row row_src[r] for every destination row r copied from a text row
 * (row_src = the map tasu_splice_scatter wrote into row_src_ws, text_mode 0: flattened token index b*S + j). */
int tasu_splice_text_grad(const void* grad_emb, int dtype, int64_t grad_row_stride, const int64_t* row_src, int64_t n_rows,
                          int H, void* grad_text, int64_t text_row_stride, int64_t text_rows, void* stream);

/* dst[i] = cast(scale * src[i]), i < n: gradient wire format of the data-parallel all-reduce (fp32 -> bf16 before it,
 * bf16 -> fp32 with the 1/world factor after it; reference: fp32 ZeRO-2 reduce-scatter, conf/ds_config.json:15-21). */
int tasu_flat_scale_cast(const void* src, int src_dtype, void* dst, int dst_dtype, int64_t n, float scale, void* stream);

/* ---------------------------------------------------------------------------------------
 * Cross-rank packing (north star: "NCCL ... only to all-gather per-rank compressed lengths and outputs for packing").
 * After all-gathering the per-rank lengths (all_lens [W, b_max] int64, utterance i = rank i % W, local index i / W —
 * the reference's sample sharding, speech_dataset_large.py:80-91) and the packed rows (flat: rank r's rows start at row
 * r * slab_rows), copy the rows of the selected utterances sel[0..n_sel) (global ids, any order; NULL = 0..n_sel-1) into
 * `out`, contiguously in selection order, and emit their lengths.  `total` [1] = rows written (never more than max_rows).
 * workspace: (W * b_max + 2 * n_sel) int32.  Two launches, nothing is built on or shipped from the host. */
int tasu_packed_select(const void* flat, int dtype, int64_t flat_row_stride, int64_t slab_rows, int H,
                       const int64_t* all_lens, int W, int b_max, int64_t n_global, const int32_t* sel, int n_sel,
                       void* out, int64_t out_row_stride, int64_t max_rows, int64_t* out_lens, int32_t* total,
                       int32_t* workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TASU_BRIDGE_H_ */
