// Host-side engine: weight store, workspace, per-stream state and the three stage drivers.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "ar_decode.cuh"
#include "common.cuh"

namespace svanon {

// MODEL_STYLE = CAMPPlus (style vector), MODEL_TIMBRE = BiCodec speaker encoder (timbre latents): the prompt path's two
// speaker encoders (speaker.hpp)
enum Model : int { MODEL_AR = 0, MODEL_TOKENIZER = 1, MODEL_VOCODER = 2, MODEL_STYLE = 3, MODEL_TIMBRE = 4, MODEL_COUNT = 5 };

struct Tensor {
  float* data = nullptr;
  std::vector<long long> shape;
  long long numel() const {
    long long n = 1;
    for (auto s : shape) n *= s;
    return n;
  }
};

// Bump allocator over one device arena; reset at the start of every stage call.  Growing reallocates
// (device sync) and only happens the first time a larger problem size is seen.
struct Workspace {
  char* base = nullptr;
  size_t cap = 0, off = 0;
  void ensure(size_t bytes);
  void reset() { off = 0; }
  float* alloc_f(long long n) { return reinterpret_cast<float*>(alloc_bytes((size_t)n * sizeof(float))); }
  void* alloc_bytes(size_t bytes);
  ~Workspace();
};

struct ConvNextW {
  const float *gamma, *dw_w /*[7][C]*/, *dw_b, *ln_w, *ln_b, *pw1_w, *pw1_b, *pw2_w, *pw2_b;
  int C;
};

// The conv stack shared by the two FireflyArchitecture encoders (content tokenizer and the vocoder's own encoder):
// ConvNeXtEncoder (stem + 4 stages) + the quantizer's 2 x [causal conv k2 s2 + ConvNeXt] down-sampling.
struct ConvStackW {
  const float *stem_w = nullptr, *stem_b = nullptr, *stem_ln_w = nullptr, *stem_ln_b = nullptr;
  const float *mid_ln_w[3] = {}, *mid_ln_b[3] = {}, *mid_w[3] = {}, *mid_b[3] = {};
  std::vector<ConvNextW> blocks[4];
  const float *bb_norm_w = nullptr, *bb_norm_b = nullptr;
  const float *down_w[2] = {}, *down_b[2] = {};
  ConvNextW down_block[2];
  bool ready = false;
};

struct EncLayerW {
  const float *attn_norm, *wqkv, *wo, *ffn_norm, *w1, *w3, *w2, *ls_attn, *ls_ffn;
};

struct ResConvW {
  const float* w;   // d == 1: [C][k][C]  (single GEMM over k overlapping rows); d > 1: [k][C][C] (k taps)
  const float* b;
  int k, d;
};

struct Engine;

// A channels-last activation buffer that keeps `margin` rows of history in front of the `rows` new rows of a
// step: [margin + rows][C].  After every step the newest `margin` rows are moved to the front.
struct SBuf {
  float* base = nullptr;
  int margin = 0, rows = 0, C = 0;
  long long seg = 0;               // floats between the buffers of consecutive streams
  float* data() const { return base + (long long)margin * C; }
};

struct ShiftDesc {
  float* base;
  int margin_floats;
  int shift_floats;
  long long seg;
};

// Persistent state of the incremental vocoder for one stream (with >= 15 frames of true
// history the per-frame result equals the reference's 64-frame window recompute).
struct VocState {
  int c = 0;                    // code frames per step
  int B = 1;                    // streams side by side (every SBuf holds B segments)
  float* arena = nullptr;
  size_t arena_floats = 0;
  SBuf u1, u2, p0, c0;
  struct Level {
    SBuf x;                     // transposed-conv output, read by conv1 of iteration 0 of all branches
    SBuf t[3][3];               // conv1 outputs  [branch][iteration]
    SBuf r[3][2];               // residual stream after iterations 0 and 1 [branch][iteration]
    SBuf next;                  // ParallelBlock mean = next level's (or conv_post's) input
  } lv[5];
  ShiftDesc* desc_dev = nullptr;
  int n_desc = 0;
  int primed_frames = 0;
  ~VocState();
};

// Per-stream descriptor of the many-stream decode path (ar_batch.cu), rebuilt by the host every step
struct ArBatchSlot {
  float *kc, *vc, *fkc, *fvc, *x_audio;
  const long long* content_id;
  const float* cond_row;
  const float* noise;
  int* out_codes;
  int* pred_hist;               // [8][pred_ld] history to append this frame's codes to (column pred_col), or null
  long long pred_ld;
  int pred_col;
  int pos;
  unsigned step;
  unsigned long long seed;
  float temperature, top_p;     // this stream's sampling arguments
};

struct ArBatchWork {
  static constexpr int RING = 4;
  int cap = 0;
  float *x = nullptr, *nrm = nullptr, *qkv = nullptr, *y = nullptr, *h13 = nullptr, *g = nullptr, *xf = nullptr,
        *logits = nullptr;
  ArBatchSlot* slots_dev = nullptr;
  ArBatchSlot* slots_host[RING] = {};     // pinned staging ring
  cudaEvent_t ev[RING] = {};
  bool ev_pending[RING] = {};
  int cur = 0;
  void ensure(int B);
  ~ArBatchWork();
};

// Receptive field, in content frames, of one transformer-input token of the tokenizer's conv stack (39, SURVEY.md
// section 8a-E) plus one frame of slack.
constexpr int ENC_RF = 40;

// Causal-conv history of the tokenizer's conv stack for B streams side by side: the newest 6 rows of every buffer a
// k = 7 causal conv reads (the mel rows in front of the stem, the input of each of the 18 + 2 ConvNeXt blocks).
// Steady-state INPUT rows of every causal layer of the conv stack for the last `frames_cap` content frames of B streams (rings
// indexed by absolute row number modulo the capacity; mel-rate layers hold 4 rows per frame, the two down-sampled levels 2 and
// 1).  Filled by the passes that run with true left context (the per-layer-history pass of the newest frames, and the
// first full-window pass); read by the window-start pass of Engine::enc_window_step, which then recomputes only the rows
// the zero padding at the window start can reach (Engine::enc_conv_stack_head).
struct ConvStackRings {
  float* arena = nullptr;
  float* mel = nullptr;                 // stem input          [B][4 F][160]
  float* blk[20] = {};                  // ConvNeXt block inputs [B][4 F][C_j] (18 blocks), [B][2 F][512], [B][F][512]
  float* ds_in[2] = {};                 // inputs of the two stride-2 convs [B][4 F][512], [B][2 F][512]
  int B = 0, frames_cap = 0;            // F (power of two)
  long long frames = 0;                 // content frames appended so far = absolute frame number of the next append
  long long filled_from = 0;            // rings hold steady-state rows for absolute frames >= filled_from (contaminated rows of
                                        // the first pass excluded by construction, see enc_conv_stack_head)
  void alloc(int n_streams, int window_frames);
  void release() { if (arena) cudaFree(arena); arena = nullptr; B = 0; frames_cap = 0; frames = 0; }
  ~ConvStackRings() { release(); }
};

struct ConvStackHist {
  static constexpr int N_BLK = 20;
  float* arena = nullptr;
  float* mel = nullptr;                 // [B][6][160]
  float* blk[N_BLK] = {};               // [B][6][C_j]
  int B = 0;
  ConvStackRings* rings = nullptr;      // optional: every pass that updates the history also appends its rows here
  void alloc(int n_streams);
  ~ConvStackHist() {
    if (arena) cudaFree(arena);
  }
};

// Ring-buffer state of the streaming window encoder (Engine::enc_window_step): the transformer inputs of the last
// window, for B streams side by side.
struct EncWindowState {
  float* xt[2] = {nullptr, nullptr};   // [B][S][512], double-buffered
  int cur = 0, B = 0, S = 0;
  bool valid = false;
  bool enabled = true;
  ConvStackHist hist;                  // per-layer conv history of the newest frames (many-stream mode)
  bool hist_valid = false;
  ConvStackRings rings;                // steady-state layer inputs of the last window (triangular window-start pass)
  bool use_rings = default_use_rings();
  static bool default_use_rings() {
    const char* e = getenv("SVANON_ENC_HEAD_TRI");            // 0: the window-start span always goes through the whole conv stack
    return !e || atoi(e) != 0;
  }
  int tail_hist_min_streams = default_tail_hist_min();   // conv history for the newest frames from this many streams (0: never)
  static int default_tail_hist_min() {
    const char* e = getenv("SVANON_ENC_HIST_MIN_STREAMS");     // tuning knob
    return e ? atoi(e) : 8;
  }
  ~EncWindowState() {
    for (auto p : xt)
      if (p) cudaFree(p);
  }
};

// Stateful content encoder (enc_stream.cu): B streams side by side that are fed their NEW samples only.
constexpr int ENC_RING = 520;                 // K/V ring slots per (layer, stream, head): 512-token window + <= 8 tokens of a push
constexpr int ENC_POS_PERIOD = 8192;          // RoPE position base moves in steps of this many frames
constexpr int ENC_ROPE_STREAM_ROWS = ENC_POS_PERIOD + 1024;
constexpr int ENC_STREAM_WAVE = (N_FFT - HOP) + 8 * SAMPLES_PER_FRAME;   // wave staging per stream: look-back + one push
struct EncStream {
  int B = 0;
  long long pos = 0;                          // content frames pushed so far (all B streams advance together)
  ConvStackHist hist;                         // per-layer causal-conv history
  float* wave = nullptr;                      // [B][ENC_STREAM_WAVE]
  float *kc = nullptr, *vc = nullptr;         // [ENC_LAYERS][B][ENC_HEADS][ENC_RING][64], keys UNROTATED
  long long* off_dev = nullptr;               // [B] member position = pos + off (non-zero only after a cohort merge)
  std::vector<long long> off;
  EncStream() = default;
  EncStream(const EncStream&) = delete;
  EncStream& operator=(const EncStream&) = delete;
  ~EncStream();
};

std::vector<SBuf*> voc_state_bufs(VocState& vs);                                        // voc_stream.cu
void voc_state_concat(VocState& dst, VocState& a, VocState& b, cudaStream_t st);       // dst <- a's streams, then b's
void voc_state_select(VocState& dst, VocState& a, const int* keep, int n, cudaStream_t st);   // dst <- a's streams keep[]

constexpr int HIST_CAP = 4096;     // columns kept of src_content_codes / pred_codes (the reference trims to 2048)

struct Stream {
  Engine* eng = nullptr;
  int max_seq = AR_MAX_SEQ;
  int delay = 0;
  float temperature = 0.7f, top_p = 0.7f;
  float gen_temperature = -1.f, gen_top_p = -1.f;   // svanon_ar_set_generate_sampling: frames >= 1 of generate (< 0: unset)
  // ---- AR state (device)
  float *kc = nullptr, *vc = nullptr, *fkc = nullptr, *fvc = nullptr;
  float* x_audio = nullptr;        // [768]     cached_new_audio_emb
  float* ref_emb_tail = nullptr;   // [8][768]  cached_ref_emb (last `delay` prompt frames)
  float* spk_rows = nullptr;       // [33][768] speaker condition rows of the current prompt
  int* codes_dev = nullptr;        // [8]       codes of the last decode step
  long long* content_id_dev = nullptr;
  float* noise_dev = nullptr;      // [8][8][1000] staging for host-side noise tapes (up to 8 frames per chunk)
  const long long* step_content_id = nullptr;   // per-step inputs of the next decode launch
  const float* step_noise = nullptr;
  const float* step_cond_row = nullptr;
  int* step_pred_hist = nullptr;                // many-stream path: append the step's codes here (column step_pred_col)
  int step_pred_col = 0;
  int pos_next = 0;                // next free sequence position (== cached_kv_pos[-1] + 1)
  unsigned step = 0;               // decode_one_token_ar calls so far (prefills included)
  unsigned long long seed = 0;
  // ---- per-chunk loop state (InferenceWrapper.setup_stream_caches, infer_arvc.py:443-460), all on the device
  int enc_win = 0, dec_win = 0, max_seq_frames = 0, buffer_frames = 0, chunk = 1;
  float *wave_ring = nullptr, *wave_ring_tmp = nullptr;     // [enc_win*2048]
  long long* src_hist = nullptr;   // [HIST_CAP]     src_content_codes
  int n_src = 0;
  int* pred_hist = nullptr;        // [8][HIST_CAP]  pred_codes
  int n_pred = 0;
  long long* ref_content_dev = nullptr;   // [ref_frames]   prompt truncated to max_prompt_frames
  int* ref_audio_dev = nullptr;           // [8][ref_frames]
  int ref_frames = 0;
  float *style_dev = nullptr, *timbre_dev = nullptr;
  bool delay_prefilled = false;
  long long* ids_win_dev = nullptr;       // [enc_win]
  long long* codes_win_dev = nullptr;     // [8][dec_win]
  float* wave_win_dev = nullptr;          // [dec_win*2048]
  EncWindowState enc_state;               // conv-stack outputs of the last window (Engine::enc_window_step)
  EncStream enc_stream;                   // encoder mode 3: stateful encoder (offline-encode semantics), enc_stream.cu
  bool enc_stateful = false;
  VocState voc;                           // incremental vocoder state (used when dec_win >= 16)
  int voc_mode = 1;                       // 1: incremental when possible, 0: always recompute the window
  bool voc_incremental = false;
  int voc_fed = 0;                        // pred frames already pushed through the incremental vocoder
  // optional per-stage device timing of the last processed chunk (bench.py): events E0,E1=A0,A1,V0,V1
  bool timing = false;
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
  ~Stream();
};

struct Engine {
  int device = 0;
  int num_sms = 148;
  std::unordered_map<std::string, Tensor> w[MODEL_COUNT];
  bool finalized[MODEL_COUNT] = {false, false, false, false, false};
  std::vector<float*> owned;                   // packed weights built at finalize
  Workspace ws;
  cudaStream_t own_stream = nullptr;
  // lo terms (x - trunc_tf32(x)) of the activations that feed the pair GEMM kernel (gemm_pair.cu), written by their producers:
  // grow-only buffers, one per role (0: norm / dwconv output, 1: hidden of the MLP, 2: attention output).  null while a stream
  // capture is running and the buffer would have to grow -- the GEMM then splits A itself or stays on the single-CTA kernel.
  struct LoScratch { float* p = nullptr; size_t cap = 0; } lo_scr[3];
  float* lo_scratch(int slot, size_t n_floats, cudaStream_t st);

  // ---- AR
  ArDecodeArgs ar{};
  const float *ctx_w = nullptr, *ctx_b = nullptr, *style_w = nullptr, *style_b = nullptr;
  const float *w4s = nullptr, *w4e = nullptr;
  float *ar_x = nullptr, *ar_h = nullptr, *ar_q = nullptr, *ar_g = nullptr, *ar_part = nullptr, *ar_logits = nullptr;
  unsigned* ar_barrier = nullptr;
  // last encoder layer for the kept tokens only (enc_transformer_bsq); SVANON_ENC_TAIL_ONLY=0 runs it for all rows
  bool enc_tail_only = [] { const char* e = getenv("SVANON_ENC_TAIL_ONLY"); return !e || atoi(e) != 0; }();
  float *dbg_slow_logits = nullptr, *dbg_hidden = nullptr, *dbg_fast_logits = nullptr;
  bool debug_logits = false;
  unsigned long long* ar_prof = nullptr;       // in-kernel timeline counters (svanon_ar_profile), null = off
  int ar_variant = 1;                          // batch-1 decode kernel: 0 direct loads, 1 TMA-staged weights

  // ---- tokenizer
  const float *dft_w = nullptr, *fb_t = nullptr;      // windowed DFT basis and mel filterbank (same LogMelSpectrogram in both)
  ConvStackW tok_cs;                                  // tokenizer conv stack
  EncLayerW enc_layers[ENC_LAYERS];
  const float *enc_norm_w = nullptr, *enc_rope = nullptr, *bsq_w = nullptr, *bsq_b = nullptr;
  const float* enc_rope_stream = nullptr;             // [ENC_ROPE_STREAM_ROWS][32][2], same formula (optional: stateful encoder)

  // ---- vocoder
  const float *fsq_w = nullptr, *fsq_b = nullptr;
  const float *up_w[2] = {}, *up_b[2] = {};
  ConvNextW up_block[2];
  const float *pre_w = nullptr, *pre_b = nullptr;
  const float *ups_w[5] = {}, *ups_b[5] = {};
  ResConvW res1[5][3][3], res2[5][3][3];
  const float *post_w = nullptr, *post_b = nullptr;
  ConvStackW voc_cs;                                  // the vocoder's own encoder (prompt path), optional
  const float *fsq_in_w = nullptr, *fsq_in_b = nullptr;   // [8][4][64], [8][4]

  ~Engine();

  const Tensor& get(int model, const std::string& name) const;
  bool has(int model, const std::string& name) const { return w[model].count(name) != 0; }
  float* upload(const std::vector<float>& host);
  float* dev_alloc(long long n_floats);

  void load_tensor(int model, const std::string& name, const float* data, int rank, const long long* shape);
  void finalize(int model);
  void finalize_ar();
  void finalize_tokenizer();
  void finalize_vocoder();
  void finalize_style();       // speaker.cu
  void finalize_timbre();

  // stage drivers (all device pointers, stream-ordered, no host sync)
  int enc_num_ids(long long n_samples) const { return (int)(((n_samples / HOP) / 2) / 2); }
  void enc_encode(const float* wave_dev /*[B][n]*/, int B, long long n_samples, long long* ids_dev /*[B][S]*/, cudaStream_t st);
  // hist_mode 0: zero left context (a window / utterance start); 1: the same, and the newest 6 rows of every causal
  // conv input are captured into `hist`; 2: left context = `hist` (continuation of the streams captured there: the wave
  // pointers must have 1536 real samples in front), `hist` is advanced
  void enc_conv_stack(const ConvStackW& w, const float* const* src, const long long* pitch, int nsrc, int per_src,
                      long long n, float* xt, cudaStream_t st, ConvStackHist* hist = nullptr, int hist_mode = 0,
                      float* mel_dst = nullptr, long long mel_dst_seg = 0);
  void pack_conv_stack(int model, ConvStackW& cs);
  void build_spectrogram_consts(int model);
  // FireflyArchitecture.encode of the vocoder (firefly.py:561-574): wave [B][n] -> codec ids int32 [B][8][n/2048]
  void voc_encode(const float* wave_dev, int B, long long n, int* codes_dev, cudaStream_t st);
  // hidden_out (test hook): the final-norm output of the rows whose ids are produced ([B * S] or [B * keep_last] rows x 512)
  void enc_transformer_bsq(float* xt, int B, int S, long long* ids_dev, cudaStream_t st, int keep_last = 0,
                           float* hidden_out = nullptr);
  // the window re-encode of the streaming loop with the conv-stack outputs kept between chunks (wave_ring [B][S*2048])
  void enc_window_step(EncWindowState& state, const float* wave_ring, int B, int S, int c, long long* ids_dev,
                       cudaStream_t st);
  // one stream: window assemble + transformer + BSQ as one persistent chain launch (enc_chain.cu); false = not applicable
  // span_src / span_pitch given (the two wave spans, as for enc_conv_stack): the conv stack from the stem on runs inside the
  // chain too and `spans` is not read
  bool enc_window_chain(const float* spans, const float* prev, float* next, int S, int c, int Ls, long long* ids,
                        cudaStream_t st, const float* const* span_src = nullptr, const long long* span_pitch = nullptr);
  std::shared_ptr<struct EncChains> enc_chains;
  // test hooks (svanon_debug_enc_transformer, svanon_debug_chain_gemm): device pointers
  void debug_enc_transformer(const float* xt, int S, int keep, bool use_chain, float* hidden_out, long long* ids_out, cudaStream_t st);
  void debug_chain_gemm(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int act, int repeat,
                        cudaStream_t st);
  // stateful encoder (enc_stream.cu): c new frames per stream -> their ids (ids[b * ids_ld + j])
  void enc_stream_init(EncStream& es, int B);
  void enc_stream_reset(EncStream& es, cudaStream_t st);
  void enc_push(EncStream& es, const float* wave_chunk, long long pitch, int c, long long* ids, long long ids_ld, cudaStream_t st);
  void voc_quantizer_decode(const long long* codes_dev, long long ld, int T, float* z_dev /*[4T][512]*/, cudaStream_t st);
  void voc_head(const float* z_dev /*[L][512]*/, int L, float* wave_dev /*[512 L]*/, cudaStream_t st);
  void voc_decode(const long long* codes_dev, long long ld, int T, float* wave_dev, cudaStream_t st);
  // out == nullptr: in place on x
  // window-start span of B streams with only the rows the zero padding can reach recomputed (ConvStackRings): wave = the
  // windows' first samples (stream b at wave + b * pitch), abs_frame0 = absolute frame number of the window start; writes the
  // ENC_RF - 1 transformer inputs per stream that differ from the steady state to xt_out [B][ENC_RF - 1][512]
  void enc_conv_stack_head(const ConvStackW& w, const float* wave, long long pitch, int B, const ConvStackRings& rings,
                           long long abs_frame0, float* xt_out, long long out_seg, cudaStream_t st);
  // the window-start pass and the per-layer-history pass of the c newest frames in the same launches (engine.cu)
  void enc_conv_stack_merged(const ConvStackW& w, const float* wave_ring, long long pitch, long long nw, int B, int c, ConvStackHist& hist,
                             ConvStackRings& rings, long long abs_frame0, float* xt_head, long long head_seg, float* xt_tail,
                             cudaStream_t st);
  void convnext(const ConvNextW& w, float* x, int rows, float* tmp, float* hid, cudaStream_t st, float* out = nullptr,
                int seg_rows = 0, long long x_seg = 0, long long out_seg = 0);
  // stateful (incremental) vocoder, voc_stream.cu
  void voc_state_init(VocState& vs, int frames_per_step, int n_streams = 1);
  void voc_state_reset(VocState& vs, cudaStream_t st);
  // codes: stream b's [8][c] block starts b * codes_seg after `codes` (row stride ld); wave_out [B][c*2048]
  void voc_step(VocState& vs, const long long* codes, long long ld, float* wave_out, cudaStream_t st,
                long long codes_seg = 0);

  // speaker encoders of the prompt path (speaker.cu / speaker.hpp); 16 kHz waves, device pointers
  std::shared_ptr<void> style_net, timbre_net;
  void kaldi_fbank(const float* wave, long long n, float* feat /*[frames][80]*/, cudaStream_t st);
  void campplus_forward(const float* feat /*[T][80]*/, long long T, int len, float* out /*[192]*/, cudaStream_t st);
  void style_vector(const float* wave, long long n, float* out /*[192]*/, cudaStream_t st);
  void timbre_latent(const float* wave, long long n, long long wave_len, float* out /*[32][128]*/, int* indices /*[32] or null*/,
                     cudaStream_t st);

  // AR
  void ar_forward_tokens(Stream& s, float* x /*[M][768]*/, int M, int pos0, cudaStream_t st);
  void ar_forward_tokens_many(Stream* const* ss, const int* M, const int* pos0, int n, float* x, cudaStream_t st);
  int ar_prompt_rows(Stream& s, const long long* ref_content, const int* ref_audio, int T, const float* style,
                     const float* timbre, float* x, cudaStream_t st);
  int ar_delay_rows(Stream& s, const long long* src_content, float* x, cudaStream_t st);
  void reprompt_many(Stream* const* streams, int n, Workspace& staging, cudaStream_t st);
  void ar_prefill_prompt(Stream& s, const long long* ref_content, const int* ref_audio, int T, const float* style,
                         const float* timbre, cudaStream_t st);
  void ar_prefill_delay(Stream& s, const long long* src_content, int n, cudaStream_t st);
  void reprompt(Stream& s, Workspace& staging, cudaStream_t st);
  void ar_decode_step(Stream* const* streams, int batch, cudaStream_t st);
  // any number of streams: the frame as a sequence of GEMM / attention / sampler kernels over all streams (ar_batch.cu)
  void ar_decode_step_gemm(Stream* const* streams, int batch, cudaStream_t st);
  ArBatchWork arb;
};

}  // namespace svanon
