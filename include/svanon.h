/*
 * svanon_b200 -- C ABI of the B200-native streaming voice-conversion hot path.
 *
 * Drop-in boundary for the per-chunk loop of Plachtaa/StreamVoiceAnon
 * (`InferenceWrapper.process_one_chunk`, evaluations/infer_arvc.py:492-596).  The reference has no FFI: its
 * "plugin API" is hydra `_target_` instantiation plus Python method calls on three model objects.  Each entry
 * point below states the reference method it replaces (file:line, relative to the reference root); the Python
 * shims in streamvoiceanon_b200/ bind them with ctypes and re-expose the reference's method names
 * (INTEGRATION.md shows the binding a maintainer adds).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; `svanon_last_error()` (thread-local) explains.
 *  - data pointers may be HOST or DEVICE memory (detected with cudaPointerGetAttributes).  Device pointers are
 *    consumed / produced stream-ordered on `cuda_stream` (a cudaStream_t, NULL = legacy default stream) with no
 *    host synchronisation; host pointers are copied in/out inside the call, and calls that write to host
 *    memory return after the copy has completed.
 *  - activations are fp32.  Shapes use the reference's names; [a][b] is row-major with b contiguous.
 *  - one engine per process / GPU; the library may be called from any thread, one call at a time per engine.
 */
#ifndef SVANON_H
#define SVANON_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct svanon_engine svanon_engine;
typedef struct svanon_stream svanon_stream;

#define SVANON_MODEL_AR 0        /* modules.arvc_wrapper.ARVCWrapper            (dual_ar_delay_0_8.pth)           */
#define SVANON_MODEL_TOKENIZER 1 /* firefly_encoder.FireflyArchitecture         (asr_s2s_bsq_8192_causal_...pth)  */
#define SVANON_MODEL_VOCODER 2   /* firefly.FireflyArchitecture (decode path)   (firefly-gan-vq-fsq-...pth)       */

const char* svanon_last_error(void);
int64_t svanon_kernel_launches(void); /* kernels launched by this library so far (process-wide) */

/* ---- engine + weights --------------------------------------------------------------------------------------
 * Replaces model construction + `load_state_dict(strict=False)` (evaluations/infer_arvc.py:51-65,67-96,160-165).
 * Tensors are passed one by one under their reference state-dict key, fp32.  The vocoder may be given in
 * weight-norm form (`...parametrizations.weight.original0/1`); it is folded at finalize, which is what
 * `remove_parametrizations()` (infer_arvc.py:94) does.  Three derived buffers must be supplied too, because the
 * reference builds them with torch ops whose last-bit results the engine must share:
 *   AR:        "decoder.model.freqs_cis" [2048][32][2], "decoder.model.fast_freqs_cis" [8][32][2]
 *              (precompute_freqs_cis, dual_ar_stream.py:993-1001: bf16-rounded, widened to fp32)
 *   tokenizer: "quantizer.pre_module.freqs_cis" [2048][32][2] (persistent buffer of the checkpoint),
 *              "spec_transform.fb" [1025][160] (slaney mel filterbank, spectrogram.py:93-106)
 */
int svanon_engine_create(int device, svanon_engine** out);
void svanon_engine_destroy(svanon_engine* e);
int svanon_load_tensor(svanon_engine* e, int model, const char* name, const float* data, int rank,
                       const int64_t* shape);
int svanon_finalize_weights(svanon_engine* e, int model);

/* ---- stage E: content tokenizer ---------------------------------------------------------------------------
 * Replaces `FireflyArchitecture.encode(audios, audio_lengths)` (modules/vqgan/modules/firefly_encoder.py:553-566)
 * for one full-length utterance or streaming window: log-mel -> ConvNeXt -> 2x down-sample -> windowed
 * transformer -> 13-bit BSQ ids.  wave [n_samples] at 44.1 kHz; ids_out [svanon_enc_num_ids(n_samples)] int64. */
int svanon_enc_num_ids(int64_t n_samples);
int svanon_enc_encode(svanon_engine* e, const float* wave, int64_t n_samples, int64_t* ids_out, void* cuda_stream);

/* ---- stage V: vocoder -------------------------------------------------------------------------------------
 * `svanon_voc_quantizer_decode` replaces `DownsampleFiniteScalarQuantize.decode` (modules/vqgan/modules/fsq.py:
 * 112-116): codes [8][T] int64 -> z [4T][512] (channels-LAST; the reference returns the transpose [512][4T]).
 * `svanon_voc_head` replaces `HiFiGANGenerator.forward` (modules/vqgan/modules/firefly.py:280-293):
 * z [L][512] -> wave [512 L].  `svanon_voc_decode` is `code2wav_fn` (evaluations/infer_arvc.py:173-176). */
int svanon_voc_quantizer_decode(svanon_engine* e, const int64_t* codes, int T, float* z_out, void* cuda_stream);
int svanon_voc_head(svanon_engine* e, const float* z, int L, float* wave_out, void* cuda_stream);
int svanon_voc_decode(svanon_engine* e, const int64_t* codes, int T, float* wave_out, void* cuda_stream);

/* prompt path: `wav2target_fn` (evaluations/infer_arvc.py:168-171) = FireflyArchitecture.encode of
 * the vocoder (modules/vqgan/modules/firefly.py:561-574) + DownsampleFiniteScalarQuantize.encode (fsq.py:106-110):
 * reference wave -> the 8 FSQ codec ids per frame that prefill_prompt consumes.  Needs the checkpoint's `backbone.*`,
 * `quantizer.downsample.*` and `quantizer.residual_fsq.rvqs.*.project_in.*` tensors and "spec_transform.fb" (as for the
 * tokenizer).  waves [n_utt][n_samples] full-length rows; codes_out [n_utt][8][n_samples / 2048] int32. */
int svanon_voc_encode(svanon_engine* e, const float* waves, int n_utt, int64_t n_samples, int32_t* codes_out,
                      void* cuda_stream);

/* ---- stage A: dual-AR decode ------------------------------------------------------------------------------
 * A stream owns what the reference keeps inside one ARVCWrapper/DualARWrapper instance: the slow/fast KV
 * caches (`setup_caches`, infer_arvc.py:55-59, dual_ar_stream.py:225-243,459-475), cached positions,
 * `cached_new_audio_emb` and `cached_ref_emb` (dual_ar_stream.py:775-796,808-815,834-836). */
int svanon_stream_create(svanon_engine* e, int max_seq_len, svanon_stream** out);
void svanon_stream_destroy(svanon_stream* s);
/* DualARWrapper.set_delay, dual_ar_stream.py:630-637 (0..8) */
int svanon_ar_set_delay(svanon_stream* s, int delay);
/* sampling: temperature/top_p defaults 0.7/0.7 (dual_ar_stream.py:1103-1104); seed drives the built-in
 * counter-based Exp(1) generator used when no noise tape is passed */
int svanon_ar_set_sampling(svanon_stream* s, float temperature, float top_p, uint64_t seed);
/* ARVCWrapper.prefill_prompt, arvc_wrapper.py:100-112 -> dual_ar_stream.py:764-796.
 * ref_content [T] int64, ref_audio [8][T] int32, style [192], timbre [32][128] */
int svanon_ar_prefill_prompt(svanon_stream* s, const int64_t* ref_content, const int32_t* ref_audio, int T,
                             const float* style, const float* timbre, void* cuda_stream);
/* ARVCWrapper.prefill_src_condition4delay, arvc_wrapper.py:114-119 -> dual_ar_stream.py:798-815; n == delay */
int svanon_ar_prefill_delay(svanon_stream* s, const int64_t* src_content, int n, void* cuda_stream);
/* ARVCWrapper.decode_one, arvc_wrapper.py:121-126 -> dual_ar_stream.py:817-837 -> decode_one_token_ar :1168-1219.
 * content_id: one int64.  noise: NULL or the Exp(1) tape of this step for the 8 codebook samplers, [8][1000]
 * (slot 0, the discarded 8192-way token head, is never drawn).  codes_out [8] int32; *last_pos (host int) is the
 * value the reference returns as `kv_pos[-1]`. */
int svanon_ar_decode_one(svanon_stream* s, const int64_t* content_id, const float* noise, int32_t* codes_out,
                         int32_t* last_pos, void* cuda_stream);
/* the same step for n in {1,2,4} independent streams in ONE kernel launch (weights are read once).
 * content_ids [n], noise NULL or [n][8][1000], codes_out [n][8] */
int svanon_ar_decode_batch(svanon_stream* const* streams, int n, const int64_t* content_ids, const float* noise,
                           int32_t* codes_out, void* cuda_stream);
/* ARVCWrapper.generate (offline), arvc_wrapper.py:82-98 -> dual_ar_stream.py:698-762.
 * ref_content [Tr], ref_audio [8][Tr] int32, src_content [Ts]; noise NULL or [Ts][8][1000]; codes_out [8][Ts] */
int svanon_ar_generate(svanon_stream* s, const int64_t* ref_content, const int32_t* ref_audio, int Tr,
                       const int64_t* src_content, int Ts, const float* style, const float* timbre,
                       const float* noise, int32_t* codes_out, void* cuda_stream);

/* Per-call sampling arguments of `generate(..., temperature=, top_p=)` (modules/arvc_wrapper.py:82-98): the reference
 * samples the first frame with its defaults (its prefill call passes no kwargs, modules/dual_ar_stream.py:723) and every
 * later frame with the caller's (:745-752).  Set before svanon_ar_generate; negative values clear (all frames use
 * svanon_ar_set_sampling's).  `repetition_penalty` has no effect in the reference (previous_tokens is always None). */
int svanon_ar_set_generate_sampling(svanon_stream* s, float temperature, float top_p);

/* BASELINE config 3 (batched offline conversion; the reference runs its utterances one `generate` after the other, batch
 * size 1: evaluations/infer_arvc.py:56,350-357): `generate` for n utterances with ONE pass over the weights per frame.
 * Every utterance has its own stream (same delay and max_seq_len), prompt, length and noise tape; utterance k produces
 * exactly what svanon_ar_generate produces for it alone.  Arrays of n DEVICE pointers: ref_content[k] [Tr[k]],
 * ref_audio[k] [8][Tr[k]] int32, src_content[k] [Ts[k]], style[k] [192], timbre[k] [32][128], noise NULL or noise[k] NULL or
 * [Ts[k]][8][1000], codes_out[k] [8][Ts[k]] int32.  Shorter utterances leave the lock-step batch when they are done. */
int svanon_ar_generate_many(svanon_stream* const* streams, int n, const int64_t* const* ref_content,
                            const int32_t* const* ref_audio, const int* Tr, const int64_t* const* src_content, const int* Ts,
                            const float* const* style, const float* const* timbre, const float* const* noise,
                            int32_t* const* codes_out, void* cuda_stream);

int svanon_ar_position(const svanon_stream* s); /* next free sequence position */
/* test hook: capture the logits of the next decode steps (slow 8192-way head, pre-norm hidden state, 8 fast
 * heads) of stream 0 of each launch; read them back with svanon_ar_read_debug (host pointers, may be NULL) */
int svanon_ar_debug_logits(svanon_engine* e, int enable);
/* GEMM back end of the encoder / prefill / vocoder projections (process-wide): 1 = fp32 on CUDA cores,
 * 2 (default) = tcgen05 tensor cores with a 3xTF32 split (fp32-grade products, accumulator in TMEM) for M >= 32, N >= 64,
 * 0 = the register double-buffered CUDA-core kernel only.  `svanon_debug_gemm` runs one C = act(A W^T + bias)
 * (A [M][K], W [N][K], row-major fp32, K % 16 == 0; act 0 none / 1 GELU) through the selected back end (tests). */
int svanon_set_gemm_mode(int mode);
/* Single-stream stages as persistent chain kernels (process-wide; the default, 1, can be changed with SVANON_CHAIN in the
 * environment).  Bit 0: the window encoder's assemble + transformer + BSQ run as ONE cooperative launch that walks a
 * device-side op list with grid barriers (csrc/chain.cu) instead of 75 kernel launches (measured 3.41 -> 3.30 ms per chunk).
 * Bit 1: the conv stack behind the mel filterbank joins the same launch (116 more phases; correct, measured SLOWER than the
 * per-op kernels -- profiles/r2q_* -- so it is off).  0 = one kernel launch per op.  Same arithmetic either way (3xTF32
 * tensor-core products, fp32 elsewhere); the tests hold all three to the same fixtures. */
int svanon_set_chain_mode(int mode);
/* test hooks of the chain kernel (device pointers).  svanon_debug_chain_gemm: C = act(A W^T + bias) (A [M][K], W [N][K], M <= 384,
 * K % 32 == 0, N % 16 == 0; act 0 none / 1 GELU) as `repeat` x [GEMM phase, element-wise phase] of one chain launch.
 * svanon_debug_enc_transformer: WindowLimitedTransformer + BSQ (windowed_transformer.py:337-354, bsq.py:330-369) of one window
 * xt [S][512], S <= 128, through the chain (use_chain 1) or the per-op path (0); keep > 0: only the last `keep` rows are produced
 * (the streaming loop, infer_arvc.py:506-518); hidden_out = their final-norm rows [rows][512], ids_out [S]. */
int svanon_debug_chain_gemm(svanon_engine* e, const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                            int act, int repeat, void* cuda_stream);
int svanon_debug_enc_transformer(svanon_engine* e, const float* xt, int S, int keep, int use_chain, float* hidden_out,
                                 int64_t* ids_out, void* cuda_stream);
/* Precision of the tensor-core GEMMs (process-wide): 0 (default) = PARITY mode, fp32-grade products through the 3xTF32
 * split -- token ids bit-exact against the fp32 reference; 1 = PERF mode at the reference's own GPU precision (fp16 autocast
 * for linears and convs, evaluations/infer_arvc.py:493): one kind::f16 tcgen05 pass with fp32 accumulation -- activations are
 * rounded to fp16 on the way into the tensor core (they stay fp32 in HBM), weights come from fp16 copies made once per weight on
 * first use.  Ids are no longer bit-exact: bench.py reports the id-agreement rate and the teacher-forced logit error of this
 * mode beside the parity-mode numbers, never instead of them.  GEMMs that run on CUDA cores (M < 32, thin channels) and the
 * single-stream persistent decode kernel are unaffected. */
int svanon_set_precision(int mode);
/* Wide GEMMs (M >= 4096 rows: the many-stream batches) on CTA pairs -- tcgen05.mma.cta_group::2, operands by tensor-map TMA,
 * persistent tile loop with two TMEM accumulators (csrc/gemm_pair.cu).  Same arithmetic and results as the single-CTA kernel,
 * bit for bit.  -1 = environment (SVANON_GEMM_PAIR, default on), 0 = off, 1 = on, 2 = on with explicitly masked hi terms (A/B
 * check of the "kind::tf32 truncates" assumption the kernel rests on). */
int svanon_set_gemm_pair(int mode);
/* test / measurement hook: the lo term of A (A - trunc_tf32(A), device memory, [M][K] compact) for the following
 * svanon_debug_gemm calls, as the engine's row-wise kernels write it beside their results; null (default): the pair kernel
 * runs its own split pass in front of the GEMM */
int svanon_debug_gemm_alo(const float* a_lo);
/* test hook for the fused GEMM forms of the encoder transformer (device pointers only, synchronous): W2 != NULL: C [M][N] =
 * silu(A W^T) * (A W2^T) (SwiGLU gate, windowed_transformer.py FeedForward); rope_table != NULL: C [M][N] = A W^T with the
 * interleaved-pair RoPE of apply_rotary_emb (windowed_transformer.py:368-380) on its first rope_cols columns, position of row m
 * = m % rope_seg_rows (rope_seg_rows > 0) or m.  The pair kernel does both in its epilogue, the other back ends through the
 * row-wise kernels: the results are the same bits. */
int svanon_debug_gemm_fused(svanon_engine* e, const float* A, const float* W, const float* W2, const float* rope_table,
                            int rope_cols, int rope_seg_rows, float* C, int M, int N, int K, void* cuda_stream);
/* number of GEMM launches the pair kernel has taken in this process (tests: the path under test really ran) */
long long svanon_gemm_pair_launches(void);
/* programmatic dependent launch of the GEMM kernels (default on): a GEMM's launch and weight-only prologue overlap
 * the tail of the kernel before it; it blocks in griddepcontrol.wait before touching activations */
int svanon_set_pdl(int enable);
int svanon_debug_gemm(svanon_engine* e, const float* A, const float* W, const float* bias, float* C, int M, int N,
                      int K, int act, void* cuda_stream);
/* test hook: 1 = svanon_debug_gemm treats W like an engine weight -- the tensor-core kernel then takes its B operand by TMA
 * bulk copies from a pre-split (hi / lo or fp16), pre-tiled copy of W made for the call and dropped after it; 0 (default) =
 * W is caller memory and goes through the producers' register path; 2 = like 1 but the copy is kept (micro-benchmarks that
 * call again with the same, unchanged W) */
int svanon_debug_gemm_weights_static(int enable);

/* test hook: the general form of the GEMM contract the conv layers use (common.cuh GemmParams) --
 *   C[m][c_col0 + n] = bias[n] + sum_t sum_k A[(a_row0 + m * a_row_step + tap_off[t]) * lda + k] * W[t][n][k]
 * A [a_rows][lda] (rows may overlap: lda < K), W [taps][N][K], C [M][ldc]; device pointers only.  Lets a test hold any
 * single descriptor (row-offset taps of either sign, strided rows, column slices) to an fp64 product. */
int svanon_debug_gemm_taps(svanon_engine* e, const float* A, int a_rows, int lda, int a_row0, int a_row_step, const float* W,
                           int taps, const int* tap_off, const float* bias, float* C, int ldc, int c_col0, int M, int N,
                           int K, void* cuda_stream);

/* measurement aid (bench.py `roofline`): while enabled, EVERY GEMM launch of the library is bracketed by CUDA events on
 * the stream it is launched on; svanon_gemm_timing_read synchronises and returns, per back end -- [0] tcgen05 3xTF32
 * (gemm_tc.cu), [1] pipelined CUDA-core (gemm_pipe.cu), [2] register-tiled CUDA-core (gemm.cu), [3] thin-channel direct
 * conv (conv_small.cu) -- the summed launch durations (ms), the executed 2*M*N*K*taps work (GFLOP, fp32-grade products) and
 * the launch count since it was enabled.  The events break programmatic-dependent-launch overlap, so timed passes are
 * slower than production passes; enable != 0 also clears the counters. */
int svanon_gemm_timing(svanon_engine* e, int enable);
int svanon_gemm_timing_read(svanon_engine* e, double* ms /*[4]*/, double* gflop /*[4]*/, int64_t* launches /*[4]*/);

/* batch-1 decode kernel variant: 1 (default) = weights staged through shared memory with TMA bulk copies, grid
 * barriers between phases; 0 = weights loaded straight from global memory + grid barriers (what the 2- and 4-stream
 * launches use).  (Two more variants -- a barrier-free flag-in-data exchange and a per-CTA epoch-word barrier -- were
 * measured slower in round 1, profiles/README.md, and were removed from the product; git history keeps them.) */
int svanon_ar_set_kernel_variant(svanon_engine* e, int variant);
/* measurement aid: in-kernel timeline of the batch-1 decode kernel (variant 1).  enable != 0 zeroes and arms 8 cycle
 * counters that thread 0 of CTA 0 accumulates per category between markers: [0] activation load + norm, [1] wait for the
 * staged weights, [2] dot products + result stores, [3] grid barrier, [4] attention, [5] sampler, [6] other; cycles_out
 * (host, may be NULL) receives the counters accumulated since they were armed; enable == 0 disarms. */
int svanon_ar_profile(svanon_engine* e, int enable, uint64_t* cycles_out);
/* measurement aid: `iters` back-to-back grid barriers of the persistent decode kernels (one CTA per SM), optionally with
 * the publish -> barrier -> read-everybody round trip of a real phase; *ms_out = device time of the whole launch */
int svanon_debug_grid_barrier(svanon_engine* e, int iters, int exchange, float* ms_out);
int svanon_ar_read_debug(svanon_engine* e, float* slow_logits /*[8192]*/, float* hidden /*[768]*/,
                         float* fast_logits /*[8][1000]*/);

/* ---- the per-chunk loop -----------------------------------------------------------------------------------
 * `svanon_stream_set_prompt` = the tail of InferenceWrapper.prefill_prompt (evaluations/infer_arvc.py:468-489)
 * after the setup-path encoders: keeps the prompt truncated to max_prompt_frames for window padding and
 * re-prompting, sets the delay and prefills the KV cache with the full prompt.
 * `svanon_stream_setup` = InferenceWrapper.setup_stream_caches (:443-460).
 * `svanon_stream_process_chunk` = InferenceWrapper.process_one_chunk (:492-596): wave ring update, window
 * re-encode (E), warm-up phases, `chunk` decode steps (A), re-prompt when pos//2 >= max_seq_frames, vocoder on
 * the last decode_window_frames frames (V), tail select.  wave_chunk / wave_out [chunk*2048]; noise NULL or
 * [chunk][8][1000]. */
int svanon_stream_set_prompt(svanon_stream* s, const int64_t* ref_content, const int32_t* ref_audio, int T,
                             const float* style, const float* timbre, int max_prompt_frames, int delay,
                             void* cuda_stream);
int svanon_stream_setup(svanon_stream* s, int encode_window_frames, int decode_window_frames, int max_seq_frames,
                        int buffer_frames, int decode_chunk_frames);
int svanon_stream_process_chunk(svanon_stream* s, const float* wave_chunk, int n_samples, const float* noise,
                                float* wave_out, void* cuda_stream);
/* vocoder evaluation inside the loop: 1 (default) = incremental -- only the new frames are computed, every causal
 * conv reads its left context from per-stream history (equal to the window recompute whenever the window leaves
 * >= 15 frames of history); 0 = recompute decode_window_frames frames per chunk exactly like
 * the reference does.  Call before the first chunk. */
int svanon_stream_set_vocoder_mode(svanon_stream* s, int incremental);
/* window re-encode inside the loop: 1 (default) = ring-buffer state -- the tokenizer's conv-stack outputs (the
 * transformer inputs) of the previous window are kept per stream; when the window slides only its first and last
 * 40 + chunk frames go through the conv stack again (the first ones see the window-start zero padding exactly as in
 * the reference's recompute, the last ones contain the new frames), the attention transformer then runs over the
 * whole window as in the reference.  Same function of the same samples as 0 = re-encode the whole window every chunk
 * (evaluations/infer_arvc.py:495-508); used when encode_window_frames >= 2 * (40 + chunk) + 8.  With >= 8 streams side by
 * side (or mode 2: always) the newest frames do not take a 41-frame span either: they continue from PER-LAYER conv
 * history (the newest 6 rows of every causal-conv input of the stack), 4 mel rows per frame.
 * Mode 3 (opt-in, call before the first chunk) is a DIFFERENT function: the stateful encoder of svanon_enc_push_chunk --
 * ids equal to the reference's OFFLINE encode() of the stream so far, not to its 128-frame window re-encode (which restarts
 * from zero padding every chunk; the two agree on most frames but not bit for bit, SURVEY.md finding 4).  0.23 instead of 15-29
 * GFLOP per frame: the setting for stream counts per GPU, not for parity with the reference's streaming loop. */
int svanon_stream_set_encoder_mode(svanon_stream* s, int incremental);
/* per-stage device time (CUDA events on the launching stream) of the last non-warm-up chunk:
 * ms[0] = E (window encode), ms[1] = A (decode steps), ms[2] = V (vocoder) */
int svanon_stream_set_timing(svanon_stream* s, int enable);
int svanon_stream_last_timing(svanon_stream* s, float* ms);
/* copies of the loop's histories (host pointers): src_content_codes [<= cap] and pred_codes [8][n] (int64) */
int svanon_stream_history(svanon_stream* s, int64_t* src_content, int* n_src, int64_t* pred_codes, int* n_pred,
                          int cap);

/* ---- host audio boundary ---------------------------------------------------------------
 * Sample-rate conversion with torchaudio.functional.resample semantics (polyphase windowed sinc): the 16 kHz / device
 * rate <-> 44.1 kHz step the reference does on the host (`librosa.load(path, sr=self.sr)`, evaluations/infer_arvc.py:
 * 274-278; `torchaudio.functional.resample` in real-time-gui.py).  `kernel` [new][2*width + orig] is the filter bank for
 * the gcd-reduced ratio orig:new (streamvoiceanon_b200.audio builds it with the torchaudio formula); output sample
 * f*new + ph = sum_k kernel[ph][k] * wave[f*orig + k - width]; n_out <= ceil(new * n_in / orig). */
int svanon_resample(svanon_engine* e, const float* wave, int64_t n_in, const float* kernel, int orig, int new_rate, int width,
                    float* out, int64_t n_out, void* cuda_stream);

/* Speaker-embedding anonymisation: `InferenceWrapper.apply_noise_mixing`,
 * evaluations/infer_arvc.py:228-232, applied to style_vectors [1,192] and timbre_latents [1,32,128] before
 * prefill_prompt (:419-421): out = alpha * x + (1 - alpha) * (noise * std(x) + mean(x)), mean / unbiased std over all n
 * elements.  `noise` holds the n standard-normal draws the reference takes from torch's global generator
 * (`torch.randn_like`); the caller supplies them (the shim draws them the same way), so results are reproducible against
 * the reference under a shared seed.  x, noise and out may be host or device pointers; out may alias x. */
int svanon_noise_mix(svanon_engine* e, const float* x, const float* noise, int64_t n, float alpha, float* out,
                     void* cuda_stream);

/* ---- speaker encoders of the prompt path -------------------------------------------------
 * Two more model ids for svanon_load_tensor / svanon_finalize_weights:
 *   3 = style encoder: `modules.campplus.DTDNN.CAMPPlus(feat_dim=80, embedding_size=192)` (configs/hydra_arcs/sv/
 *       campplus.yaml; construction + load_state_dict at evaluations/infer_arvc.py:98-108), reference key names with the
 *       `xvector.dense.*` -> `dense.*` rename CAMPPlus.load_state_dict applies (DTDNN.py:114-130), plus two derived
 *       buffers of torchaudio.compliance.kaldi: "fbank.window" [400] (povey) and "fbank.mel_banks" [80][257];
 *   4 = timbre encoder: `modules.bicodec_speaker_encoder.speaker_encoder.SpeakerEncoder` (configs/hydra_arcs/sv/
 *       sparktts_speaker_encoder.yaml; infer_arvc.py:110-121), the tensors `tokenize_wav` touches (speaker_encoder.layer*,
 *       speaker_encoder.conv, perceiver_sampler.*, quantizer.project_in/out), plus the derived buffers of its
 *       MelSpectrogram: "mel.window" [1024] (hann(640) centred) and "mel.fb" [513][128] (slaney).
 * All waves are 16 kHz mono fp32; pointers may be host or device memory.
 *
 * svanon_kaldi_fbank      replaces `torchaudio.compliance.kaldi.fbank(wave, num_mel_bins=80, dither=0,
 *                         sample_frequency=16000)` (call site infer_arvc.py:186-191): wave [n] -> feat_out [m][80],
 *                         m = 1 + (n - 400) / 160.
 * svanon_campplus_forward replaces `CAMPPlus.forward(x, x_lens)` for one row (modules/campplus/DTDNN.py:132-138; call site
 *                         infer_arvc.py:210): feat [n_frames][80], valid_len = x_lens of the row (rows that count in the
 *                         statistics pooling, after the stride-2 TDNN) -> out [192].
 * svanon_style_vector     replaces `InferenceWrapper.calculate_style_vec` for one row (infer_arvc.py:179-211): fbank, minus
 *                         its time mean, lens = frames / 2, CAMPPlus -> out [192].
 * svanon_timbre_latent    replaces `InferenceWrapper.calculate_timbre_latent` / `SpeakerEncoder.tokenize_wav` for one row
 *                         (infer_arvc.py:213-223, speaker_encoder.py:136-144): wave [n_samples] whose first wave_len
 *                         samples are valid (a zero-padded batch row; wave_len = n_samples otherwise) -> latents_out
 *                         [32][128] (the `zq.mT` the caller keeps) and, if not null, the 32 FSQ indices. */
int svanon_kaldi_fbank(svanon_engine* e, const float* wave16k, int64_t n_samples, float* feat_out, void* cuda_stream);
int svanon_campplus_forward(svanon_engine* e, const float* feat, int64_t n_frames, int valid_len, float* out,
                            void* cuda_stream);
int svanon_style_vector(svanon_engine* e, const float* wave16k, int64_t n_samples, float* out, void* cuda_stream);
int svanon_timbre_latent(svanon_engine* e, const float* wave16k, int64_t n_samples, int64_t wave_len, float* latents_out,
                         int32_t* indices_out, void* cuda_stream);

/* ---- stateful stage entries (`enc_push_chunk`, `voc_push_frames`) -------------------------------
 * An encoder stream holds, for n_streams streams side by side, what an incremental FireflyArchitecture.encode needs: the
 * 1536-sample STFT look-back, the newest 6 rows in front of every causal conv of the stack, and a 520-slot K/V ring per
 * transformer layer.  `svanon_enc_push_chunk` takes the NEW samples only -- wave [n_streams][n_samples_per_stream], 1..8
 * whole frames of 2048 samples -- and writes the ids of the new frames, ids_out [n_streams][frames].  Parity target: the
 * ids of the reference's offline `encode()` (modules/vqgan/modules/firefly_encoder.py:553-566) on each stream's whole
 * prefix (causal-prefix equality; tests/test_gpu_stateful.py), for any stream length (RoPE positions are taken relative to
 * a base that moves every 8192 frames; the first 8703 frames use the offline encode's absolute positions).
 * `svanon_enc_stream_reset` starts new utterances (zero left context).
 *
 * A vocoder stream is the incremental vocoder of the loop on its own: per-stream causal-conv history of every layer of
 * `quantizer.upsample` + `HiFiGANGenerator` (firefly.py:280-293).  `svanon_voc_push_frames` takes frames_per_push new
 * code frames per stream -- codes [n_streams][8][frames_per_push] int64 -- and writes their samples, wave_out
 * [n_streams][frames_per_push * 2048]; a fresh / reset stream starts from zero history, i.e. pushing an utterance frame by
 * frame reproduces `svanon_voc_decode` of the whole utterance (every conv is causal). */
typedef struct svanon_enc_stream svanon_enc_stream;
typedef struct svanon_voc_stream svanon_voc_stream;
int svanon_enc_stream_create(svanon_engine* e, int n_streams, svanon_enc_stream** out);
void svanon_enc_stream_destroy(svanon_enc_stream* s);
int svanon_enc_stream_reset(svanon_enc_stream* s, void* cuda_stream);
int64_t svanon_enc_stream_position(const svanon_enc_stream* s);   /* content frames pushed so far */
int svanon_enc_push_chunk(svanon_enc_stream* s, const float* wave, int n_samples_per_stream, int64_t* ids_out, void* cuda_stream);
int svanon_voc_stream_create(svanon_engine* e, int n_streams, int frames_per_push, svanon_voc_stream** out);
void svanon_voc_stream_destroy(svanon_voc_stream* s);
int svanon_voc_stream_reset(svanon_voc_stream* s, void* cuda_stream);
int svanon_voc_push_frames(svanon_voc_stream* s, const int64_t* codes, float* wave_out, void* cuda_stream);

/* ---- many concurrent streams in lock-step --------------------------------------------------------------------
 * The reference is strictly batch-1 (max_batch_size=1, evaluations/infer_arvc.py:56; `x.view(1, 1, -1)`,
 * modules/dual_ar_stream.py:544): N concurrent utterances are N sequential calls.  These entry points run the same
 * per-stream arithmetic for N independent streams with ONE pass over the weights per step (BASELINE configs 3-4);
 * every stream produces what it would produce alone.
 *
 * `svanon_enc_encode_batch`: FireflyArchitecture.encode (firefly_encoder.py:553-566) for n_utt utterances of the
 * same length; waves [n_utt][n_samples], ids_out [n_utt][svanon_enc_num_ids(n_samples)].
 * `svanon_ar_decode_many`: ARVCWrapper.decode_one (arvc_wrapper.py:121-126) for any number of streams: the frame
 * as tensor-core GEMMs over all streams + per-stream KV-cache attention and samplers (svanon_ar_decode_batch keeps
 * the single persistent kernel, 1/2/4 streams).  content_ids [n], noise NULL or [n][8][1000], codes_out [n][8].
 * A svanon_batch is InferenceWrapper.process_one_chunk (infer_arvc.py:492-596) for N streams that started
 * together: create the streams, give each its prompt (svanon_stream_set_prompt, same delay), then
 * `svanon_batch_setup` (= setup_stream_caches, :443-460, for all of them; the vocoder runs incrementally, so
 * decode_window_frames - decode_chunk_frames must be >= 15) and one `svanon_batch_process_chunk` per chunk:
 * wave_chunks / wave_out [n][chunk*2048], noise NULL or [n][chunk][8][1000].  Streams re-prompt individually.
 * The batch does not own the streams; destroy it before them. */
typedef struct svanon_batch svanon_batch;
int svanon_enc_encode_batch(svanon_engine* e, const float* waves, int n_utt, int64_t n_samples, int64_t* ids_out,
                            void* cuda_stream);
int svanon_ar_decode_many(svanon_stream* const* streams, int n, const int64_t* content_ids, const float* noise,
                          int32_t* codes_out, void* cuda_stream);
int svanon_batch_create(svanon_engine* e, svanon_stream* const* streams, int n, svanon_batch** out);
void svanon_batch_destroy(svanon_batch* b);
int svanon_batch_setup(svanon_batch* b, int encode_window_frames, int decode_window_frames, int max_seq_frames,
                       int buffer_frames, int decode_chunk_frames);
int svanon_batch_process_chunk(svanon_batch* b, const float* wave_chunks, int n_samples_per_stream, const float* noise,
                               float* wave_out, void* cuda_stream);
/* decode path of the batch: 0 (default) = persistent kernel for 1/2/4 streams, many-stream kernels otherwise;
 * 1 = always the many-stream kernels */
int svanon_batch_set_ar_path(svanon_batch* b, int path);
/* Cohort merging (heterogeneous stream phases): two batches with the same settings that are both past
 * their warm-up chunks become ONE batch -- the members of `a` followed by the members of `b` -- which moves every member's
 * state (wave ring, encoder window / stateful-encoder state, vocoder histories) into side-by-side buffers; from the next
 * chunk on all of them share one pass over the weights.  Every stream keeps producing exactly what it produces alone.
 * `a` and `b` are left without members: destroy them.  wave_chunks / wave_out rows of the merged batch: a's streams, then b's. */
int svanon_batch_merge(svanon_batch* a, svanon_batch* b, svanon_batch** out, void* cuda_stream);
/* The members keep[0 .. n_keep) of `a` (indices into a's member order, strictly increasing) as a new lock-step batch, each with
 * the state it had (the moves of svanon_batch_merge, gathered per member): what a server calls once streams of a cohort have
 * left, so that the following chunks compute for the remaining streams only.  `a` must be past its warm-up chunks and is left
 * without members: destroy it.  The streams left out are no longer advanced by any batch; the caller owns them (destroy
 * them, or give them a new prompt and set them up again before further use).  No reference counterpart (the reference is
 * one stream per process, evaluations/infer_arvc.py:56). */
int svanon_batch_select(svanon_batch* a, const int* keep, int n_keep, svanon_batch** out, void* cuda_stream);
int svanon_batch_set_encoder_mode(svanon_batch* b, int incremental);   /* as svanon_stream_set_encoder_mode */
/* per-stage device time of the last non-warm-up chunk, as svanon_stream_last_timing */
int svanon_batch_set_timing(svanon_batch* b, int enable);
int svanon_batch_last_timing(svanon_batch* b, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* SVANON_H */
