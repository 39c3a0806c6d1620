/*
 * stz.h — C ABI of the B200-native StyleTTS-ZS inference hot path
 * (style-diffusion sampling loop + duration predictor).
 *
 * Reference interface this replaces: NONE EXISTS.  ishine/StyleTTS-ZS publishes no inference
 * code (/root/reference/README.md:15-16, "Inference — Under construction"); the path itself is
 * described only in the abstract (/root/reference/README.md:5).  The boundary is therefore the
 * module API that BASELINE.json's north_star dictates —
 *     sample_style(text_emb, prompt_feats, steps, cfg_scale) / predict_duration(...)
 * — and every entry point below is what a ctypes/cffi binding of that module API binds
 * (SURVEY.md §8b).  INTEGRATION.md shows the Python-side stub.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary.
 *   - all functions return 0 on success or a negative stz_status; none throws.
 *   - "dev" pointers are CUDA device pointers on the handle's device, "host" pointers are
 *     host memory (pinned for best copy throughput).  fp32 unless stated.
 *   - device entry points enqueue on `cuda_stream` (a cudaStream_t passed as void*) and
 *     return without synchronising; the caller owns all in/out buffers and keeps them alive
 *     until the stream is synchronised.  The library owns weights, workspace and CUDA graphs.
 *   - a handle is bound to one device and is not re-entrant (one host thread at a time); different handles may be
 *     driven from different host threads concurrently (every knob is per handle).
 *   - masks are uint8 [B,T] / [B,P], 1 = valid token; text masks must be prefix masks.
 *   - there is no CPU fallback: stz_create fails with STZ_E_DEVICE on anything but an sm_100 (compute capability 10.0)
 *     device — the library carries sm_100a code only.
 */
#ifndef STZ_H_
#define STZ_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STZ_ABI_VERSION 1

typedef enum stz_status {
  STZ_OK = 0,
  STZ_E_ARG = -1,      /* null pointer / bad enum / bad size */
  STZ_E_SHAPE = -2,    /* shape outside what the kernels support */
  STZ_E_DEVICE = -3,   /* no sm_100 device / driver entry point missing */
  STZ_E_CUDA = -4,     /* a CUDA call failed; see stz_last_error */
  STZ_E_NOMEM = -5
} stz_status;

typedef enum stz_sampler_kind {
  STZ_SAMPLER_STUDENT = 0, /* distilled few-step student: Euler on Karras(steps) + terminal 0 */
  STZ_SAMPLER_TEACHER = 1, /* teacher: ADPM2 on Karras(steps+1), 2 denoiser evals per step     */
  STZ_SAMPLER_GUIDED = 2   /* guidance-conditioned student: the distillation bakes classifier-free
                              guidance into the network — `cfg_scale` enters as an embedding of the conditioning vector and
                              ONE conditional branch is evaluated per step (half the denoiser rows of the CFG pair); same
                              Euler schedule and noise as STZ_SAMPLER_STUDENT */
} stz_sampler_kind;

/* Field order mirrors styletts-zs_b200/spec.py:StzConfig. */
typedef struct stz_config {
  int32_t n_style, d_style, d_model, n_heads, d_ff, n_layers, d_text, d_prompt, d_time;
  int32_t d_hid, d_sty_tok, n_sp_heads, n_lstm, max_dur;
  float sigma_data, sigma_max, sigma_min, rho;
} stz_config;

typedef struct stz_handle stz_handle;

int stz_abi_version(void);

/* Flat fp32 weight blob layout (mirrors spec.py:weight_entries). */
size_t stz_weights_nfloats(const stz_config* cfg);
/* Offset (in floats) of a named entry, or -1. */
int64_t stz_weight_offset(const stz_config* cfg, const char* name);

/* Uploads the blob to `device`, converts the denoiser's GEMM weights to bf16, precomputes the
 * null-context K/V.  weights_host: fp32 blob of stz_weights_nfloats(cfg) floats. */
int stz_create(const stz_config* cfg, const float* weights_host, size_t nfloats, int device,
               stz_handle** out);
void stz_destroy(stz_handle* h);
/* Message of the last failure on this handle (h may be NULL: last stz_create failure). */
const char* stz_last_error(const stz_handle* h);

/* sample_style(text_emb, prompt_feats, steps, cfg_scale): the CFG style-diffusion loop.
 *   text_emb_dev   [B,T,d_text]     prompt_feats_dev [B,P,d_prompt]
 *   text_mask_dev  [B,T] u8         prompt_mask_dev  [B,P] u8           (NULL = all valid)
 *   noise_dev      [n_slices,B,K,Ds], n_slices = 1 (student) or steps+1 (teacher)
 *   out_style_dev  [B,K,Ds] */
int stz_sample_style(stz_handle* h, const float* text_emb_dev, const uint8_t* text_mask_dev,
                     const float* prompt_feats_dev, const uint8_t* prompt_mask_dev,
                     const float* noise_dev, int B, int T, int P, int steps, float cfg_scale,
                     int sampler_kind, float* out_style_dev, void* cuda_stream);

/* predict_duration(text_emb, text_mask, style_codes) -> int32 frames per token.
 *   style_dev [B,K,Ds]; out_dur_dev [B,T] int32; out_presum_dev [B,T] fp32 or NULL (the
 *   pre-rounding sum of sigmoids, for diagnostics). */
int stz_predict_duration(stz_handle* h, const float* text_emb_dev, const uint8_t* text_mask_dev,
                         const float* style_dev, int B, int T, int32_t* out_dur_dev,
                         float* out_presum_dev, void* cuda_stream);

/* Length regulator — the step right after predict_duration: expands per-token features by the
 * integer durations, on device (the prefix sum of the durations never leaves the GPU).
 *   feats_dev [B,T,C] fp32 (C % 4 == 0, T <= 1024), dur_dev [B,T] int32 (entries <= 0 contribute no frame),
 *   out_frames_dev [B,F_max,C]: frame f of utterance b = feats[b, tok(f)], tok(f) = first token whose cumulative duration
 *   exceeds f; frames past the utterance's total are zeros; totals beyond F_max are truncated;
 *   out_frame_lens_dev [B] = min(sum of durations, F_max); out_frame_tok_dev [B,F_max] (optional, may be NULL) = tok(f) or -1. */
int stz_regulate_length(stz_handle* h, const float* feats_dev, const int32_t* dur_dev, int B, int T, int C, int F_max,
                        float* out_frames_dev, int32_t* out_frame_lens_dev, int32_t* out_frame_tok_dev, void* cuda_stream);

/* Prosody heads — F0 and energy curves over the length-regulated frames (second half; the
 * "duration/prosody predictor" of BASELINE.json's north_star):
 *     frames = regulate([d_enc | s_tok], dur)   d_enc = the duration encoder's output (input of its final BiLSTM)
 *     y = BiLSTM_pros(frames) (packed by frame count);  z = gelu_tanh([y | s_frame] W_h1^T + b_h1)
 *     f0 = z[:, :d_hid/2] . w_f0 + b_f0,  energy = z[:, d_hid/2:] . w_en + b_en,  0 past the utterance's frame count.
 * Runs predict_duration's forward first; dur_in_dev [B,T] (may be NULL) overrides the predicted durations for the
 * regulator (e.g. ground-truth alignments); out_dur_dev [B,T] (may be NULL) receives the predicted ones.
 *   out_f0_dev, out_energy_dev [B,F_max] fp32; out_frame_lens_dev [B] int32 = min(sum of durations, F_max).  T <= 1024. */
int stz_predict_prosody(stz_handle* h, const float* text_emb_dev, const uint8_t* text_mask_dev, const float* style_dev,
                        const int32_t* dur_in_dev, int B, int T, int F_max, float* out_f0_dev, float* out_energy_dev,
                        int32_t* out_frame_lens_dev, int32_t* out_dur_dev, void* cuda_stream);

/* On-device noise.  After stz_set_noise_seed, a NULL `noise` argument of stz_sample_style /
 * stz_synthesize_host means: draw every noise slice on the device — Philox4x32-10, key = seed, counter = (group of four
 * elements inside the utterance's [K*Ds] slice, utterance lo, slice, utterance hi), Box-Muller built from individually
 * rounded fp32 operations — bit-identical to oracle/philox.py.  Utterance b of a call is global utterance
 * first_utterance + b: an utterance's noise does not depend on the batch or the GPU it is sampled on.  Without a seed a
 * NULL noise argument is STZ_E_ARG. */
int stz_set_noise_seed(stz_handle* h, uint64_t seed, uint64_t first_utterance);

/* The same with explicit global utterance indices (host array, copied): utterance b of the next calls is global utterance
 * utterance_ids[b] — for sharded batches that are not a contiguous range (shard.py's length-sorted round-robin).  The
 * calls that follow must have B == n.  stz_set_noise_seed returns to the contiguous numbering. */
int stz_set_noise_utterances(stz_handle* h, uint64_t seed, const uint64_t* utterance_ids, int n);

/* The generator on its own (unit tests, callers that want the tensor): out_dev [slices, B, n_per_utt] fp32,
 * n_per_utt % 4 == 0. */
int stz_philox_normal(uint64_t seed, uint64_t first_utterance, int slices, int B, int n_per_utt, float* out_dev,
                      int device, void* cuda_stream);

/* Host-buffer form of the whole path (what a non-CUDA caller binds): copies the inputs H2D,
 * runs sample_style and (if out_dur_host != NULL) predict_duration on the sampled codes, copies
 * the results D2H and synchronises.  All pointers are host pointers; `noise` may be NULL after stz_set_noise_seed
 * (no noise H2D at all). */
int stz_synthesize_host(stz_handle* h, const float* text_emb, const uint8_t* text_mask,
                        const float* prompt_feats, const uint8_t* prompt_mask, const float* noise,
                        int B, int T, int P, int steps, float cfg_scale, int sampler_kind,
                        float* out_style, int32_t* out_dur);

/* The same call split for pipelining (serving loops): _submit enqueues the H2D copies (copy stream), the compute
 * (internal stream) and the D2H copies (a third stream) of one batch into `slot` (0 or 1) and returns; _wait blocks until
 * that slot's outputs are in the host buffers.  Each slot has its own device staging, so submitting slot 1 before
 * waiting for slot 0 overlaps batch i+1's input copies with batch i's compute:
 *     submit(0, batch 0); for i: submit((i+1)&1, batch i+1); wait(i&1);
 * Host buffers handed to _submit (inputs and outputs) belong to the library until the slot's _wait returns; submitting
 * a slot that is still in flight first waits for it.  stz_synthesize_host == _submit(0) + _wait(0). */
int stz_synthesize_host_submit(stz_handle* h, int slot, const float* text_emb, const uint8_t* text_mask,
                               const float* prompt_feats, const uint8_t* prompt_mask, const float* noise,
                               int B, int T, int P, int steps, float cfg_scale, int sampler_kind,
                               float* out_style_host, int32_t* out_dur_host);
int stz_synthesize_host_wait(stz_handle* h, int slot);

/* ---- introspection / unit-test entry points (not part of the drop-in surface) ---------- */

/* Host-only (no device): the sigma schedule and the fused sampler-step coefficient tables of one call, as uploaded to the
 * device.  Returns the number of denoiser evaluations E (steps for the student, 2*steps
 * for the ADPM2 teacher) or a negative status; any output pointer may be NULL.
 *   sigma_out [E] fp64; coef_out [E][8] fp32 = (c_x, c_mid, c_F, c_noise, c_in(next sigma), cfg_scale, dest, 0):
 *   dest 0: x' = c_x x + c_mid x_mid + c_F F + c_noise noise, dest 1: x_mid = c_x x + c_F F;  tfeat_out [E][d_time] fp32;
 *   init_out [2] fp64 = (sigma_0, c_in(sigma_0)). */
int stz_debug_plan(const stz_config* cfg, int steps, int sampler_kind, float cfg_scale, double* sigma_out,
                   float* coef_out, float* tfeat_out, double* init_out);

/* Number of kernels launched by this handle since creation (graph nodes count per replay). */
int64_t stz_launch_count(const stz_handle* h);

/* The evaluation loop is captured as one CUDA graph per (B, text-length bucket, P, evaluations, sampler, masks).  Text
 * lengths are rounded up to 32, 64, 96, 128, 192, 256, 384, 512, then multiples of 256, with the extra key positions
 * masked, so a serving loop with free-form T captures at most ~10 graphs per (B, P, steps); at most "max_graphs" (32)
 * are kept, least recently used evicted.  Returns the number of cached graphs; *captures_total (may be NULL) = captures
 * since creation (a capture synchronises the stream: a warmed-up loop must not add any). */
int stz_graph_count(const stz_handle* h, int64_t* captures_total);

/* Sizes the workspace for the largest call that will be made (batch, text tokens, prompt tokens, sampler steps and kind),
 * so that no later call reallocates it: a reallocation synchronises the device and drops every captured graph. */
int stz_reserve(stz_handle* h, int max_B, int max_T, int max_P, int max_steps, int sampler_kind);

/* Knobs of one handle (product defaults first; the alternatives are cross-checks kept for tests and tuning):
 *   "use_graph"      1 | 0        evaluation loop as one CUDA graph per shape bucket | eager launches
 *   "t_buckets"      1 | 0        text-length buckets (see stz_graph_count) | exact T
 *   "max_graphs"     32           cached graphs before least-recently-used eviction
 *   "use_pdl"        1 | 0        programmatic dependent launch
 *   "gemm_impl"      0 | 1        persistent tcgen05 GEMM | SIMT cross-check kernel
 *   "gemm_bn"        0 | 128/192/256   tile width heuristic | forced
 *   "fuse_ln"        3 | 0 | 4    residual GEMM + AdaLN in one cluster-of-two kernel where it measured faster (37 .. 74 row
 *                                 tiles: one wave of CTA pairs), GEMM + ln_mod kernels otherwise | never fused | fused at any size
 *   "gln_tile_rows"  0 | 8..128   rows per CTA pair of that kernel: heuristic (96 where one wave still fits, else 128) | forced
 *   "attn_impl"      0 | 2        tcgen05 + TMA attention (resident keys; streaming over 128-key blocks for long text; the
 *                                 mma.sync streaming kernel beyond their shapes: P > 127, K > 64) | always the mma.sync kernel
 *   "attn_ctas"      0 | 2 | 4    resident-key attention kernel: by unit count (4 CTAs/SM with one unit in flight each where all
 *                                 units then fit one wave, else 2 persistent CTAs/SM with prefetch) | forced
 *   "attn_box2"      1 | 0        one TMA box per attention operand (both CFG branches, permuted tensor map) | one box per branch
 *   "chains"         1 | 2..8     utterance chains on parallel graph branches
 *   "lstm_impl"      0 | 1 | 3    tcgen05 cluster recurrence, W_hh in tensor memory | generic kernel | tcgen05, W_hh in shared memory
 *   "lstm_nb"        0 | 8|16|24  sequences per 8-CTA cluster of that kernel: 8 for batches <= 16, else the one of 16 / 24 with
 *                                 fewer waves x step time (15 clusters are co-resident) | forced
 *   "pred_gemm_impl" 0 | 1        split-bf16 tcgen05 predictor GEMMs | fp32 CUDA-core GEMMs
 *   "profile"        0 | 1        see stz_profile_read;   "ablate" (bit mask): tools/ablate.py timing attribution only
 * Knobs are per handle (two handles on two host threads do not interact).  Returns STZ_E_ARG for unknown keys. */
int stz_set_option(stz_handle* h, const char* key, int value);
/* Reads a knob back; also the read-only "last_fuse_mode" (0 | 3: what the last stz_sample_style dispatched for the residual
 * GEMMs — tests assert that the benched configuration runs the fused kernel) and "last_T" (its text-length bucket). */
int stz_get_option(const stz_handle* h, const char* key, int* value);

/* Profile mode (set_option "profile" = 1; setting it also clears the records): sample_style /
 * predict_duration run eagerly (no CUDA graph) with a CUDA-event pair around every kernel launch.
 * stz_profile_read synchronises and returns, for one kernel class, the summed event time (ms), the
 * summed algorithmic work (flops for classes 0,1,3,4; bytes for 2,5) and the launch count.
 * Classes: 0 tcgen05 GEMM, 1 fused attention, 2 LayerNorm+modulate, 3 fp32 CUDA-core GEMM,
 * 4 LSTM recurrence, 5 predictor elementwise, 6 other. */
int stz_profile_read(stz_handle* h, int kernel_class, double* ms, double* work, int64_t* launches);

/* Own bounds checking (compute-sanitizer is not available on every pool): set_option("guard_bytes", n > 0) re-plans the
 * library's workspace arenas with an n-byte poisoned gap after every internal buffer; stz_debug_check_guards synchronises and
 * counts the gap bytes that no longer hold the pattern — *bad_bytes == 0 means no kernel wrote outside its buffers.  Returns
 * the number of gaps checked (0 when guards are off or nothing is allocated yet) or a negative status. */
int stz_debug_check_guards(stz_handle* h, long long* bad_bytes);

/* Copies the fp32 residual stream h [B*K*2, d_model] (row = (b*K+k)*2 + branch) into
 * tap_dev after (eval, layer, stage) during eager (use_graph=0) runs; stage 0/1/2 = after the
 * self-attention / cross-attention / FFN sub-layer, layer == n_layers -> the guided F.
 * tap_dev = NULL disables. */
int stz_debug_set_tap(stz_handle* h, int eval, int layer, int stage, float* tap_dev);

/* Debug timeline of the tcgen05 BiLSTM recurrence kernel: int64 [64 steps][8] (see predictor_tc.cuh; tools/lstm_trace.py). */
int stz_debug_set_lstm_trace(stz_handle* h, long long* trace_dev);

/* Debug timeline of the tcgen05 attention kernel: trace_dev = int64 [CTAs][16] device buffer (NULL disables); thread 0
 * of every CTA stores clock64() at its phase boundaries (tools/att_trace.py). */
int stz_debug_set_att_trace(stz_handle* h, long long* trace_dev);

/* Debug timeline of the product GEMM kernel: trace_dev = int64 [CTAs][64] device buffer (NULL disables); see gemm2.cuh
 * for the slot meanings (tools/gemm_trace.py). */
int stz_debug_set_gemm_trace(stz_handle* h, long long* trace_dev);

/* Isolated-kernel microbenchmark (bench.py, sub-field of the roofline): `iters` back-to-back launches of the product GEMM
 * kernel for one shape (C[M,N] = A[M,K] W[N,K]^T, epi 2 = bf16 out, 3 = GELU bf16 out, 4 = gated residual reduce-add,
 * 6 = the fused gated residual + AdaLN kernel, N = d_model) on the handle's internal stream, timed with a CUDA-event pair;
 * *avg_us = mean microseconds per launch.  Random operands (uniform, W scaled by 1/sqrt(K)), L2-warm. */
int stz_bench_gemm(stz_handle* h, int M, int N, int K, int epi, int iters, double* avg_us);

/* Max co-resident 8-CTA clusters of the BiLSTM recurrence kernel on the current device (diagnostic). */
int stz_debug_max_lstm_clusters(void);

/* C[M,N] = A[M,K] · W[N,K]^T + bias, bf16 operands (device), fp32 out.  impl as "gemm_impl". */
int stz_op_gemm_bf16(const void* A_bf16_dev, const void* W_bf16_dev, const float* bias_dev,
                     float* C_dev, int M, int N, int K, int impl, int device, void* cuda_stream);

/* Unit-test entry for the fused attention kernels (bf16 device buffers).  Self-attention when kv_text == NULL:
 * qkv [2*B*K, 3*d_model] in the denoiser's row layout (row = (b*K + k)*2 + branch), ldq = 3*d_model.  Otherwise
 * cross-attention: queries = first d_model columns of qkv (row stride ldq); keys [text ; prompt (conditional rows only)
 * | null prompt (unconditional rows only)], each K/V buffer holding K | V (2*d_model columns) per row.
 * out [2*B*K, d_model].  impl: 0 tcgen05 + TMA (resident / streaming by shape), 2 mma.sync streaming. */
int stz_op_attention(stz_handle* h, const void* qkv_dev, int ldq, const void* kv_text_dev, const void* kv_prompt_dev,
                     const void* kv_null_dev, const uint8_t* text_mask_dev, const uint8_t* prompt_mask_dev, int B,
                     int T, int P, void* out_dev, int impl, void* cuda_stream);

/* Unit-test entries for the fused epilogues of the product GEMM kernel (device buffers; A [M,K], W [N,K] bf16).
 *   epi 0: out fp32 [M,N] = A W^T + bias                 epi 1: ... + pos[(r / 2) % n_style]  (pos [n_style, N])
 *   epi 2: out bf16 [M,N] = A W^T + bias                 epi 3: out bf16 = gelu_tanh(A W^T + bias)
 *   epi 4: out fp32 [M,N] += gate[seq(r)] * (A W^T + bias), gate[s] = mod[s * n_mod + gate_off ...], seq(r) =
 *          (r / (2 n_style)) * 2 + (r & 1): TMA reduce-add into `out` */
int stz_op_gemm_epi(stz_handle* h, const void* A_dev, const void* W_dev, const float* bias_dev, int M, int N, int K, int epi,
                    void* out_dev, const float* mod_dev, int n_mod, int gate_off, const float* pos_dev, void* cuda_stream);

/* The output projection with the fused sampler epilogue: F = A W^T + bias for the row pair
 * (2j: conditional, 2j+1: unconditional), Fg = F_u + w (F_c - F_u), x' = c_x x[j] + c_mid x_mid[j] + c_F Fg + c_noise noise[j]
 * written to x (dest 0) or x_mid (dest 1), xin rows 2j and 2j+1 = split-bf16 [hi | lo | hi] of c_in x' (row stride 3N).
 * coef_dev: the 8 floats of one stz_debug_plan row.  x, x_mid, noise [M/2, N] fp32; tap (optional) receives Fg. */
int stz_op_gemm_sampler(stz_handle* h, const void* A_dev, const void* W_dev, const float* bias_dev, int M, int N, int K,
                        float* x_dev, float* xmid_dev, const float* noise_dev, const float* coef_dev, void* xin_out_dev,
                        float* tap_dev, void* cuda_stream);

/* The fused residual GEMM + AdaLN kernel (gemmln3_kernel; N = d_model = 512, K a multiple of 256):
 *   mode 0: h' = h + gate[seq] * (A W^T + bias)      mode 1: h' = A W^T + bias + pos[(r / 2) % n_style]
 *   u = bf16(LN(h') * (1 + scale[seq]) + shift[seq])  ([hi | lo | hi] with row stride 3 * 512 if split3)
 * h_dev [M,512] fp32 is read (mode 0) and overwritten with h'; gate / shift / scale rows at mod[s * n_mod + *_off]. */
int stz_op_gemm_ln(stz_handle* h, const void* A_dev, const void* W_dev, const float* bias_dev, int M, int K, int mode,
                   float* h_dev, const float* mod_dev, int n_mod, int gate_off, int shift_off, int scale_off,
                   const float* pos_dev, int split3, void* u_out_dev, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* STZ_H_ */
