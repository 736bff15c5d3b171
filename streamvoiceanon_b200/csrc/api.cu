// C ABI (include/svanon.h): argument marshalling (host or device pointers), stream lifecycle and the
// per-chunk loop.  All model work is in engine.cu / ar_decode.cu.
#include "api_common.hpp"

namespace svanon {
extern bool g_gemm_use_pipe;
extern bool g_gemm_use_tc;
extern bool g_use_pdl;
extern bool g_use_conv_small;
extern bool g_use_chain, g_chain_conv;
#ifdef SVANON_TC_PROF
void tc_prof_dump();
#endif
}
namespace svanon {
thread_local std::string g_api_err;
}
static const float* g_debug_gemm_alo = nullptr;   // svanon_debug_gemm_alo: lo term of A for the next svanon_debug_gemm calls
static int g_debug_gemm_static = 0;         // svanon_debug_gemm_weights_static: 0 off, 1 static + dropped after the call, 2 static + kept

namespace {

void set_prompt_copy(Stream& s, const long long* ref_content, const int* ref_audio, int T, int keep, const float* style,
                     const float* timbre, cudaStream_t st) {
  // InferenceWrapper.prefill_prompt keeps `[:max_prompt_frames]` copies for window padding and re-prompting
  // (infer_arvc.py:469-473)
  if (s.ref_content_dev) cudaFree(s.ref_content_dev);
  if (s.ref_audio_dev) cudaFree(s.ref_audio_dev);
  s.ref_content_dev = dmalloc<long long>(keep);
  s.ref_audio_dev = dmalloc<int>((size_t)8 * keep);
  s.ref_frames = keep;
  SV_CUDA(cudaMemcpyAsync(s.ref_content_dev, ref_content, (size_t)keep * sizeof(long long), cudaMemcpyDeviceToDevice, st));
  SV_CUDA(cudaMemcpy2DAsync(s.ref_audio_dev, (size_t)keep * sizeof(int), ref_audio, (size_t)T * sizeof(int),
                            (size_t)keep * sizeof(int), 8, cudaMemcpyDeviceToDevice, st));
  if (!s.style_dev) s.style_dev = dmalloc<float>(192);
  if (!s.timbre_dev) s.timbre_dev = dmalloc<float>(32 * 128);
  SV_CUDA(cudaMemcpyAsync(s.style_dev, style, 192 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SV_CUDA(cudaMemcpyAsync(s.timbre_dev, timbre, 32 * 128 * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

void decode_frames(Stream& s, const long long* content_ids_dev, int n, const float* noise_dev, cudaStream_t st) {
  Engine& e = *s.eng;
  for (int i = 0; i < n; ++i) {
    if (s.n_pred >= HIST_CAP) {     // keep the newest half (the reference keeps at most 2048 entries, :593-594)
      const int keep = HIST_CAP / 2;
      int* tmp = (int*)e.ws.base;   // workspace is free between stages
      SV_CUDA(cudaMemcpy2DAsync(tmp, keep * sizeof(int), s.pred_hist + (s.n_pred - keep), HIST_CAP * sizeof(int),
                                keep * sizeof(int), 8, cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpy2DAsync(s.pred_hist, HIST_CAP * sizeof(int), tmp, keep * sizeof(int), keep * sizeof(int), 8,
                                cudaMemcpyDeviceToDevice, st));
      s.n_pred = keep;
    }
    s.step_content_id = content_ids_dev + i;
    s.step_cond_row = nullptr;
    s.step_noise = noise_dev ? noise_dev + (size_t)i * 8 * AR_CB_SIZE : nullptr;
    Stream* one = &s;
    e.ar_decode_step(&one, 1, st);
    launch_append_codes(s.codes_dev, s.pred_hist, HIST_CAP, s.n_pred, st);
    s.n_pred += 1;
  }
}

}  // namespace

extern "C" {

const char* svanon_last_error(void) { return g_api_err.c_str(); }
int64_t svanon_kernel_launches(void) { return g_kernel_launches; }

int svanon_engine_create(int device, svanon_engine** out) {
  return guarded([&] {
    SV_CHECK(out, "null out pointer");
    int count = 0;
    SV_CUDA(cudaGetDeviceCount(&count));
    SV_CHECK(device >= 0 && device < count, "no such CUDA device");
    SV_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    SV_CUDA(cudaGetDeviceProperties(&prop, device));
    SV_CHECK(prop.major == 10, "svanon_b200 is built for sm_100a (B200) only");
    auto* h = new svanon_engine();
    h->eng.device = device;
    h->eng.num_sms = prop.multiProcessorCount;
    *out = h;
  });
}

void svanon_engine_destroy(svanon_engine* e) { delete e; }

int svanon_load_tensor(svanon_engine* e, int model, const char* name, const float* data, int rank, const int64_t* shape) {
  return guarded([&] {
    SV_CHECK(e && name && data && shape && rank >= 1 && rank <= 4, "bad arguments");
    SV_CUDA(cudaSetDevice(e->eng.device));
    long long shp[4];
    for (int i = 0; i < rank; ++i) shp[i] = shape[i];
    e->eng.load_tensor(model, name, data, rank, shp);
  });
}

int svanon_finalize_weights(svanon_engine* e, int model) {
  return guarded([&] {
    SV_CHECK(e, "null engine");
    SV_CUDA(cudaSetDevice(e->eng.device));
    e->eng.finalize(model);
    SV_CUDA(cudaDeviceSynchronize());
  });
}

int svanon_enc_num_ids(int64_t n_samples) { return (int)(((n_samples / HOP) / 2) / 2); }

int svanon_enc_encode(svanon_engine* e, const float* wave, int64_t n, int64_t* ids_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && wave && ids_out, "null argument");
    Args a(e, stream, (size_t)n * 4 + 65536);
    const float* w = a.in(wave, (size_t)n);
    long long* ids = (long long*)a.out(ids_out, (size_t)svanon_enc_num_ids(n));
    e->eng.enc_encode(w, 1, n, ids, a.st);
    a.finish();
  });
}

int svanon_voc_quantizer_decode(svanon_engine* e, const int64_t* codes, int T, float* z_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && codes && z_out, "null argument");
    Args a(e, stream, (size_t)T * (64 + 4 * 512 * 4));
    const long long* c = (const long long*)a.in(codes, (size_t)8 * T);
    float* z = a.out(z_out, (size_t)4 * T * 512);
    e->eng.voc_quantizer_decode(c, T, T, z, a.st);
    a.finish();
  });
}

int svanon_voc_head(svanon_engine* e, const float* z, int L, float* wave_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && z && wave_out, "null argument");
    Args a(e, stream, (size_t)L * 512 * 4 * 2);
    const float* zd = a.in(z, (size_t)L * 512);
    float* w = a.out(wave_out, (size_t)L * 512);
    e->eng.voc_head(zd, L, w, a.st);
    a.finish();
  });
}

int svanon_voc_decode(svanon_engine* e, const int64_t* codes, int T, float* wave_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && codes && wave_out, "null argument");
    Args a(e, stream, (size_t)T * (64 + 2048 * 4));
    const long long* c = (const long long*)a.in(codes, (size_t)8 * T);
    float* w = a.out(wave_out, (size_t)T * SAMPLES_PER_FRAME);
    e->eng.voc_decode(c, T, T, w, a.st);
    a.finish();
  });
}

int svanon_stream_create(svanon_engine* e, int max_seq_len, svanon_stream** out) {
  return guarded([&] {
    SV_CHECK(e && out, "null argument");
    SV_CHECK(max_seq_len >= 64 && max_seq_len <= AR_MAX_SEQ, "max_seq_len must be in [64, 2048]");
    SV_CUDA(cudaSetDevice(e->eng.device));
    auto* h = new svanon_stream();
    h->owner = e;
    Stream& s = h->st;
    s.eng = &e->eng;
    s.max_seq = (max_seq_len + 7) / 8 * 8;   // find_multiple(max_seq_len, 8), dual_ar_stream.py:232
    const size_t kv = (size_t)AR_LAYERS * AR_HEADS * s.max_seq * HEAD_DIM;
    const size_t fkv = (size_t)AR_FAST_LAYERS * AR_HEADS * AR_CODEBOOKS * HEAD_DIM;
    s.kc = dmalloc<float>(kv); s.vc = dmalloc<float>(kv);
    s.fkc = dmalloc<float>(fkv); s.fvc = dmalloc<float>(fkv);
    SV_CUDA(cudaMemset(s.kc, 0, kv * 4)); SV_CUDA(cudaMemset(s.vc, 0, kv * 4));
    SV_CUDA(cudaMemset(s.fkc, 0, fkv * 4)); SV_CUDA(cudaMemset(s.fvc, 0, fkv * 4));
    s.x_audio = dmalloc<float>(AR_DIM);
    s.ref_emb_tail = dmalloc<float>(AR_MAX_DELAY * AR_DIM);
    s.spk_rows = dmalloc<float>(AR_SPK_TOKENS * AR_DIM);
    s.codes_dev = dmalloc<int>(8);
    s.content_id_dev = dmalloc<long long>(8);
    s.noise_dev = dmalloc<float>((size_t)8 * 8 * AR_CB_SIZE);
    s.src_hist = dmalloc<long long>(HIST_CAP);
    s.pred_hist = dmalloc<int>((size_t)8 * HIST_CAP);
    *out = h;
  });
}

void svanon_stream_destroy(svanon_stream* s) { delete s; }

int svanon_ar_set_delay(svanon_stream* s, int delay) {
  return guarded([&] {
    SV_CHECK(s, "null stream");
    SV_CHECK(delay >= 0 && delay <= AR_MAX_DELAY, "delay must be in [0, 8]");
    s->st.delay = delay;
  });
}

int svanon_ar_set_sampling(svanon_stream* s, float temperature, float top_p, uint64_t seed) {
  return guarded([&] {
    SV_CHECK(s, "null stream");
    s->st.temperature = temperature; s->st.top_p = top_p; s->st.seed = seed;
  });
}

int svanon_ar_set_generate_sampling(svanon_stream* s, float temperature, float top_p) {
  return guarded([&] {
    SV_CHECK(s, "null stream");
    SV_CHECK((temperature < 0.f && top_p < 0.f) || (temperature >= 0.f && top_p >= 0.f), "set both or clear both (negative)");
    s->st.gen_temperature = temperature; s->st.gen_top_p = top_p;
  });
}

int svanon_ar_prefill_prompt(svanon_stream* s, const int64_t* ref_content, const int32_t* ref_audio, int T,
                             const float* style, const float* timbre, void* stream) {
  return guarded([&] {
    SV_CHECK(s && ref_content && ref_audio && style && timbre, "null argument");
    Args a(s->owner, stream, (size_t)T * 48 + 32768);
    const long long* rc = (const long long*)a.in(ref_content, (size_t)T);
    const int* ra = a.in(ref_audio, (size_t)8 * T);
    const float* sv = a.in(style, 192);
    const float* tl = a.in(timbre, 32 * 128);
    s->st.eng->ar_prefill_prompt(s->st, rc, ra, T, sv, tl, a.st);
    a.finish();
  });
}

int svanon_ar_prefill_delay(svanon_stream* s, const int64_t* src_content, int n, void* stream) {
  return guarded([&] {
    SV_CHECK(s && src_content, "null argument");
    Args a(s->owner, stream, 4096);
    const long long* sc = (const long long*)a.in(src_content, (size_t)n);
    s->st.eng->ar_prefill_delay(s->st, sc, n, a.st);
    a.finish();
  });
}

int svanon_ar_decode_batch(svanon_stream* const* streams, int n, const int64_t* content_ids, const float* noise,
                           int32_t* codes_out, void* stream) {
  return guarded([&] {
    SV_CHECK(streams && n >= 1 && n <= AR_MAX_BATCH && content_ids && codes_out, "bad arguments");
    svanon_engine* h = streams[0]->owner;
    Args a(h, stream, (size_t)n * (8 * AR_CB_SIZE * 4 + 256));
    const long long* ids = (const long long*)a.in(content_ids, (size_t)n);
    const float* nz = a.in(noise, (size_t)n * 8 * AR_CB_SIZE);
    int* out = a.out(codes_out, (size_t)n * 8);
    Stream* ss[AR_MAX_BATCH];
    for (int i = 0; i < n; ++i) {
      SV_CHECK(streams[i] && streams[i]->owner == h, "streams must belong to one engine");
      ss[i] = &streams[i]->st;
      ss[i]->step_content_id = ids + i;
      ss[i]->step_cond_row = nullptr;
      ss[i]->step_noise = nz ? nz + (size_t)i * 8 * AR_CB_SIZE : nullptr;
    }
    h->eng.ar_decode_step(ss, n, a.st);
    for (int i = 0; i < n; ++i)
      SV_CUDA(cudaMemcpyAsync(out + i * 8, ss[i]->codes_dev, 8 * sizeof(int), cudaMemcpyDeviceToDevice, a.st));
    a.finish();
  });
}

int svanon_ar_decode_one(svanon_stream* s, const int64_t* content_id, const float* noise, int32_t* codes_out,
                         int32_t* last_pos, void* stream) {
  if (!s) { g_api_err = "null stream"; return 1; }
  svanon_stream* one = s;
  const int rc = svanon_ar_decode_batch(&one, 1, content_id, noise, codes_out, stream);
  if (rc == 0 && last_pos) *last_pos = s->st.pos_next - 1;
  return rc;
}

namespace {
// Prompt part of offline generate for one utterance (dual_ar_stream.py:709-722), shared by svanon_ar_generate and
// svanon_ar_generate_many: 33 speaker rows, then (cond_t, audio'_t) for t < Tr + d with cond = [ref_cond, src_cond[:d]]
// and audio' = [wait4start[:d], embed(ref_audio)], then remaining[0] -- up to, not including, the first decode step
// (all but the last two tokens go through the multi-token path; the last two are a normal decode step).
void generate_prefill(Engine& e, Stream& s, const long long* rc, const int* ra, int Tr, const long long* sc, int Ts,
                      const float* sv, const float* tl, cudaStream_t st) {
  const int d = s.delay;
  const int n_pairs = Tr + d;
  const int n_tok = AR_SPK_TOKENS + 2 * n_pairs + 1;
  SV_CHECK(n_tok + 2 * (Ts - 1) <= s.max_seq, "utterance does not fit the KV cache (max_seq_len)");
  e.ws.ensure(((size_t)(n_tok + 8) * 12000 + (1u << 20)) * sizeof(float));
  e.ws.reset();
  float* x = e.ws.alloc_f((long long)(n_tok + 2) * AR_DIM);
  GemmParams p;
  p.A = tl; p.W = e.ctx_w; p.C = x; p.bias = e.ctx_b; p.M = 32; p.N = AR_DIM; p.K = 128; p.lda = 128; p.ldc = AR_DIM;
  launch_gemm(p, st);
  GemmParams q;
  q.A = sv; q.W = e.style_w; q.C = x + 32 * AR_DIM; q.bias = e.style_b; q.M = 1; q.N = AR_DIM; q.K = 192; q.lda = 192;
  q.ldc = AR_DIM;
  launch_gemm(q, st);
  float* seq = x + AR_SPK_TOKENS * AR_DIM;
  launch_gather_rows(e.ar.cond_emb, rc, seq, Tr, AR_DIM, 2 * AR_DIM, st);
  if (d > 0) {
    launch_gather_rows(e.ar.cond_emb, sc, seq + (long long)2 * Tr * AR_DIM, d, AR_DIM, 2 * AR_DIM, st);
    launch_copy_rows(e.w4s, AR_DIM, seq + AR_DIM, 2 * AR_DIM, d, AR_DIM, st);
  }
  launch_embed_codes(e.ar.codebook_emb, ra, Tr, seq + (long long)(2 * d + 1) * AR_DIM, Tr, 2 * AR_DIM, st);
  const int n_pre = n_tok - 2;
  launch_copy_rows(seq + (long long)(2 * n_pairs - 1) * AR_DIM, AR_DIM, s.x_audio, AR_DIM, 1, AR_DIM, st);
  e.ar_forward_tokens(s, x, n_pre, 0, st);
  s.pos_next = n_pre;
}
}  // namespace

int svanon_ar_generate(svanon_stream* sh, const int64_t* ref_content, const int32_t* ref_audio, int Tr,
                       const int64_t* src_content, int Ts, const float* style, const float* timbre, const float* noise,
                       int32_t* codes_out, void* stream) {
  return guarded([&] {
    SV_CHECK(sh && ref_content && ref_audio && src_content && style && timbre && codes_out, "null argument");
    Stream& s = sh->st;
    Engine& e = *s.eng;
    const int d = s.delay;
    SV_CHECK(Tr >= 1 && Ts >= 1 && Ts >= d, "generate needs Tr >= 1 and Ts >= delay");
    Args a(sh->owner, stream, (size_t)(Tr + Ts) * 64 + (size_t)Ts * 8 * AR_CB_SIZE * 4 + 65536);
    const long long* rc = (const long long*)a.in(ref_content, (size_t)Tr);
    const int* ra = a.in(ref_audio, (size_t)8 * Tr);
    const long long* sc = (const long long*)a.in(src_content, (size_t)Ts);
    const float* sv = a.in(style, 192);
    const float* tl = a.in(timbre, 32 * 128);
    const float* nz = a.in(noise, (size_t)Ts * 8 * AR_CB_SIZE);
    int* out = a.out(codes_out, (size_t)8 * Ts);
    cudaStream_t st = a.st;
    generate_prefill(e, s, rc, ra, Tr, sc, Ts, sv, tl, st);
    // per-call sampling arguments apply from the SECOND frame on: the reference's prefill call passes none
    // (dual_ar_stream.py:723), the loop passes the caller's (:745-752)
    struct Restore {
      Stream& s; float t, p;
      ~Restore() { s.temperature = t; s.top_p = p; }
    } restore{s, s.temperature, s.top_p};
    for (int i = 0; i < Ts; ++i) {
      if (i == 1 && s.gen_temperature >= 0.f) { s.temperature = s.gen_temperature; s.top_p = s.gen_top_p; }
      // remaining = [src_cond[d:], wait4end[:d]]  (dual_ar_stream.py:716)
      const int j = d + i;
      if (j < Ts) { s.step_content_id = sc + j; s.step_cond_row = nullptr; }
      else { s.step_content_id = nullptr; s.step_cond_row = e.w4e + (long long)(j - Ts) * AR_DIM; }
      s.step_noise = nz ? nz + (size_t)i * 8 * AR_CB_SIZE : nullptr;
      Stream* one = &s;
      e.ar_decode_step(&one, 1, st);
      SV_CUDA(cudaMemcpy2DAsync(out + i, (size_t)Ts * sizeof(int), s.codes_dev, sizeof(int), sizeof(int), 8,
                                cudaMemcpyDeviceToDevice, st));
    }
    a.finish();
  });
}

int svanon_ar_generate_many(svanon_stream* const* streams, int n, const int64_t* const* ref_content,
                            const int32_t* const* ref_audio, const int* Tr, const int64_t* const* src_content, const int* Ts,
                            const float* const* style, const float* const* timbre, const float* const* noise,
                            int32_t* const* codes_out, void* stream) {
  return guarded([&] {
    SV_CHECK(streams && n >= 1 && ref_content && ref_audio && Tr && src_content && Ts && style && timbre && codes_out,
             "null argument");
    svanon_engine* h = streams[0] ? streams[0]->owner : nullptr;
    SV_CHECK(h, "null stream");
    SV_CUDA(cudaSetDevice(h->eng.device));
    Engine& e = h->eng;
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<Stream*> ss(n);
    int max_ts = 0;
    for (int k = 0; k < n; ++k) {
      SV_CHECK(streams[k] && streams[k]->owner == h, "streams must belong to one engine");
      for (int j = 0; j < k; ++j) SV_CHECK(streams[j] != streams[k], "a stream may appear only once in a call");
      SV_CHECK(ref_content[k] && ref_audio[k] && src_content[k] && style[k] && timbre[k] && codes_out[k], "null utterance buffer");
      SV_CHECK(on_device(ref_content[k]) && on_device(ref_audio[k]) && on_device(src_content[k]) && on_device(style[k]) &&
                   on_device(timbre[k]) && on_device(codes_out[k]) && (!noise || !noise[k] || on_device(noise[k])),
               "per-utterance buffers must be device pointers");
      Stream& s = streams[k]->st;
      SV_CHECK(s.delay == streams[0]->st.delay && s.max_seq == streams[0]->st.max_seq, "utterances of one call share delay and max_seq_len");
      SV_CHECK(Tr[k] >= 1 && Ts[k] >= 1 && Ts[k] >= s.delay && Ts[k] <= HIST_CAP, "generate needs Tr >= 1 and delay <= Ts <= 4096");
      ss[k] = &s;
      max_ts = std::max(max_ts, Ts[k]);
    }
    for (int k = 0; k < n; ++k)
      generate_prefill(e, *ss[k], (const long long*)ref_content[k], ref_audio[k], Tr[k], (const long long*)src_content[k], Ts[k],
                       style[k], timbre[k], st);
    const int d = ss[0]->delay;
    std::vector<Stream*> active;
    active.reserve(n);
    // per-call sampling arguments (svanon_ar_set_generate_sampling) apply from the SECOND frame on, per utterance
    struct Saved { float t, p; };
    std::vector<Saved> saved(n);
    for (int k = 0; k < n; ++k) saved[k] = {ss[k]->temperature, ss[k]->top_p};
    struct Restore {
      std::vector<Stream*>& ss; std::vector<Saved>& sv;
      ~Restore() { for (size_t k = 0; k < ss.size(); ++k) { ss[k]->temperature = sv[k].t; ss[k]->top_p = sv[k].p; } }
    } restore{ss, saved};
    for (int i = 0; i < max_ts; ++i) {
      if (i == 1)
        for (Stream* s : ss)
          if (s->gen_temperature >= 0.f) { s->temperature = s->gen_temperature; s->top_p = s->gen_top_p; }
      active.clear();
      for (int k = 0; k < n; ++k) {
        if (i >= Ts[k]) continue;                     // shorter utterances leave the lock-step batch when they are done
        Stream& s = *ss[k];
        const long long* sc = (const long long*)src_content[k];
        const int j = d + i;                          // remaining = [src_cond[d:], wait4end[:d]]  (dual_ar_stream.py:716)
        if (j < Ts[k]) { s.step_content_id = sc + j; s.step_cond_row = nullptr; }
        else { s.step_content_id = nullptr; s.step_cond_row = e.w4e + (long long)(j - Ts[k]) * AR_DIM; }
        s.step_noise = (noise && noise[k]) ? noise[k] + (size_t)i * 8 * AR_CB_SIZE : nullptr;
        s.step_pred_hist = s.pred_hist;               // the step writes its 8 codes into column i of the history
        s.step_pred_col = i;
        active.push_back(&s);
      }
      e.ar_decode_step_gemm(active.data(), (int)active.size(), st);
    }
    for (int k = 0; k < n; ++k) {
      Stream& s = *ss[k];
      SV_CUDA(cudaMemcpy2DAsync(codes_out[k], (size_t)Ts[k] * sizeof(int), s.pred_hist, (size_t)HIST_CAP * sizeof(int),
                                (size_t)Ts[k] * sizeof(int), 8, cudaMemcpyDeviceToDevice, st));
      s.step_pred_hist = nullptr;
      s.step_content_id = nullptr;
      s.step_cond_row = nullptr;
      s.step_noise = nullptr;
    }
  });
}

int svanon_ar_position(const svanon_stream* s) { return s ? s->st.pos_next : -1; }

int svanon_ar_debug_logits(svanon_engine* e, int enable) {
  return guarded([&] {
    SV_CHECK(e, "null engine");
    e->eng.debug_logits = enable != 0;
  });
}

int svanon_set_pdl(int enable) {
  g_use_pdl = enable != 0;
  return 0;
}

int svanon_set_chain_mode(int mode) {
  g_use_chain = (mode & 1) != 0;
  g_chain_conv = (mode & 2) != 0;
  return 0;
}

int svanon_debug_chain_gemm(svanon_engine* e, const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                            int act, int repeat, void* stream) {
  return guarded([&] {
    SV_CHECK(e && A && W && C && M > 0 && N > 0 && K > 0, "bad arguments");
    SV_CHECK(on_device(A) && on_device(W) && on_device(C) && (!bias || on_device(bias)), "device pointers only");
    SV_CUDA(cudaSetDevice(e->eng.device));
    e->eng.debug_chain_gemm(A, W, bias, C, M, N, K, act, repeat, (cudaStream_t)stream);
  });
}

int svanon_debug_enc_transformer(svanon_engine* e, const float* xt, int S, int keep, int use_chain, float* hidden_out,
                                 int64_t* ids_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && xt && ids_out, "bad arguments");
    SV_CHECK(on_device(xt) && on_device(ids_out) && (!hidden_out || on_device(hidden_out)), "device pointers only");
    SV_CUDA(cudaSetDevice(e->eng.device));
    e->eng.debug_enc_transformer(xt, S, keep, use_chain != 0, hidden_out, reinterpret_cast<long long*>(ids_out), (cudaStream_t)stream);
  });
}

int svanon_set_gemm_mode(int mode) {
  return guarded([&] {
    SV_CHECK(mode >= 0 && mode <= 2, "gemm mode: 0 = fp32 CUDA-core (register double-buffer only), 1 = fp32 CUDA-core, 2 = tcgen05 3xTF32");
    g_gemm_use_pipe = mode >= 1;
    g_gemm_use_tc = mode == 2;
    g_gemm_pair_allowed = mode == 2;
    g_use_conv_small = mode >= 1;
  });
}

int svanon_debug_gemm_weights_static(int enable) {
  g_debug_gemm_static = enable;
  return 0;
}

int svanon_debug_gemm_alo(const float* a_lo) {
  g_debug_gemm_alo = a_lo;
  return 0;
}

int svanon_set_gemm_pair(int mode) {
  return guarded([&] {
    SV_CHECK(mode >= -1 && mode <= 2, "gemm pair mode: -1 environment, 0 off, 1 on, 2 on with masked hi copies");
    g_gemm_pair_mode = mode;
  });
}

long long svanon_gemm_pair_launches(void) { return g_gemm_pair_launches; }

int svanon_set_precision(int mode) {
  return guarded([&] {
    SV_CHECK(mode == 0 || mode == 1, "precision: 0 = fp32-grade (3xTF32 split, parity mode), 1 = fp16 single-pass tensor-core GEMMs (perf mode)");
    g_gemm_half = mode == 1;
  });
}

int svanon_debug_gemm(svanon_engine* e, const float* A, const float* W, const float* bias, float* C, int M, int N, int K,
                      int act, void* stream) {
  return guarded([&] {
    SV_CHECK(e && A && W && C && M > 0 && N > 0 && K > 0 && K % 16 == 0, "bad arguments");
    Args a(e, stream, ((size_t)M * K + (size_t)N * K + (size_t)M * N + N) * 4 + 65536);
    GemmParams p;
    p.A = a.in(A, (size_t)M * K); p.W = a.in(W, (size_t)N * K); p.bias = a.in(bias, (size_t)N);
    p.C = a.out(C, (size_t)M * N);
    p.M = M; p.N = N; p.K = K; p.lda = K; p.ldc = N; p.act = act;
    p.w_static = g_debug_gemm_static != 0; // caller memory: normally never cached as a converted weight copy
    if (g_debug_gemm_alo) { SV_CHECK(on_device(g_debug_gemm_alo) && on_device(A), "svanon_debug_gemm_alo: device pointers only"); p.Alo = g_debug_gemm_alo; }
    launch_gemm(p, a.st);
    a.finish();
    if (g_debug_gemm_static == 1) gemm_forget_weights(p.W);
#ifdef SVANON_TC_PROF
    tc_prof_dump();
#endif
  });
}

int svanon_debug_gemm_fused(svanon_engine* e, const float* A, const float* W, const float* W2, const float* rope_table,
                            int rope_cols, int rope_seg_rows, float* C, int M, int N, int K, void* stream) {
  return guarded([&] {
    SV_CHECK(e && A && W && C && M > 0 && N > 0 && K > 0 && K % 16 == 0, "bad arguments");
    SV_CHECK(on_device(A) && on_device(W) && on_device(C) && (!W2 || on_device(W2)) && (!rope_table || on_device(rope_table)),
             "device pointers only");
    SV_CHECK((W2 != nullptr) != (rope_table != nullptr), "exactly one fused form: W2 (SwiGLU gate) or rope_table (RoPE on q | k)");
    SV_CUDA(cudaSetDevice(e->eng.device));
    cudaStream_t st = (cudaStream_t)stream;
    GemmParams p;
    p.A = A; p.W = W; p.C = C; p.M = M; p.N = N; p.K = K; p.lda = K; p.ldc = N;
    p.w_static = g_debug_gemm_static != 0;
    p.Alo = g_debug_gemm_alo;
    float* tmp = nullptr;
    if (W2) {
      SV_CUDA(cudaMalloc(&tmp, (size_t)M * 2 * N * sizeof(float)));
      p.W2 = W2; p.dual_tmp = tmp;
    } else {
      p.rope_table = rope_table; p.rope_cols = rope_cols; p.rope_seg_rows = rope_seg_rows;
    }
    launch_gemm(p, st);
    SV_CUDA(cudaStreamSynchronize(st));
    if (tmp) cudaFree(tmp);
    if (g_debug_gemm_static == 1) { gemm_forget_weights(p.W); if (W2) gemm_forget_weights(W2); }
  });
}

int svanon_debug_gemm_taps(svanon_engine* e, const float* A, int a_rows, int lda, int a_row0, int a_row_step, const float* W,
                           int taps, const int* tap_off, const float* bias, float* C, int ldc, int c_col0, int M, int N, int K,
                           void* stream) {
  return guarded([&] {
    SV_CHECK(e && A && W && C && tap_off && M > 0 && N > 0 && K > 0 && taps >= 1 && taps <= MAX_TAPS, "bad arguments");
    SV_CHECK(on_device(A) && on_device(W) && on_device(C) && (!bias || on_device(bias)), "device pointers only");
    SV_CHECK(a_row_step >= 1 && c_col0 >= 0 && c_col0 + N <= ldc, "bad layout");
    for (int t = 0; t < taps; ++t) {
      const long long lo = (long long)a_row0 + tap_off[t], hi = (long long)a_row0 + (long long)(M - 1) * a_row_step + tap_off[t];
      SV_CHECK(lo >= 0 && hi * lda + K <= (long long)a_rows * lda, "a tap reads outside the A buffer");
    }
    SV_CUDA(cudaSetDevice(e->eng.device));
    GemmParams p;
    p.A = A + (long long)a_row0 * lda; p.W = W; p.bias = bias; p.C = C + c_col0;
    p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldc = ldc; p.a_row_step = a_row_step; p.taps = taps;
    for (int t = 0; t < taps; ++t) p.tap_off[t] = tap_off[t];
    p.w_static = g_debug_gemm_static != 0;     // svanon_debug_gemm_weights_static: W treated like an engine weight
    launch_gemm(p, (cudaStream_t)stream);
    if (g_debug_gemm_static == 1) {
      SV_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
      gemm_forget_weights(p.W);
    }
  });
}

int svanon_gemm_timing(svanon_engine* e, int enable) {
  return guarded([&] {
    SV_CHECK(e, "null engine");
    SV_CUDA(cudaSetDevice(e->eng.device));
    gemm_timing_enable(enable != 0);
  });
}

int svanon_gemm_timing_read(svanon_engine* e, double* ms, double* gflop, int64_t* launches) {
  return guarded([&] {
    SV_CHECK(e && ms && gflop && launches, "null argument");
    SV_CUDA(cudaSetDevice(e->eng.device));
    long long n[GEMM_BACKENDS];
    gemm_timing_read(ms, gflop, n);
    for (int i = 0; i < GEMM_BACKENDS; ++i) launches[i] = n[i];
  });
}

int svanon_ar_set_kernel_variant(svanon_engine* e, int variant) {
  return guarded([&] {
    SV_CHECK(e, "null engine");
    SV_CHECK(variant == 0 || variant == 1, "variant: 0 direct loads, 1 TMA-staged weights");
    e->eng.ar_variant = variant;
  });
}

int svanon_ar_profile(svanon_engine* e, int enable, uint64_t* cycles_out) {
  return guarded([&] {
    SV_CHECK(e, "null engine");
    Engine& eng = e->eng;
    SV_CUDA(cudaSetDevice(eng.device));
    SV_CUDA(cudaDeviceSynchronize());
    if (cycles_out && eng.ar_prof)
      SV_CUDA(cudaMemcpy(cycles_out, eng.ar_prof, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    if (enable) {
      if (!eng.ar_prof) SV_CUDA(cudaMalloc(&eng.ar_prof, 8 * sizeof(unsigned long long)));
      SV_CUDA(cudaMemset(eng.ar_prof, 0, 8 * sizeof(unsigned long long)));
    } else if (eng.ar_prof) {
      cudaFree(eng.ar_prof);
      eng.ar_prof = nullptr;
    }
  });
}

int svanon_debug_grid_barrier(svanon_engine* e, int iters, int exchange, float* ms_out) {
  return guarded([&] {
    SV_CHECK(e && ms_out && iters > 0, "bad arguments");
    SV_CHECK(e->eng.finalized[MODEL_AR], "AR weights not finalized");
    SV_CUDA(cudaSetDevice(e->eng.device));
    SV_CUDA(cudaDeviceSynchronize());
    *ms_out = grid_barrier_probe(e->eng.ar_barrier, iters, e->eng.ar_g, exchange, e->eng.num_sms, nullptr);
  });
}

int svanon_ar_read_debug(svanon_engine* e, float* slow_logits, float* hidden, float* fast_logits) {
  return guarded([&] {
    SV_CHECK(e && e->eng.finalized[MODEL_AR], "AR weights not finalized");
    SV_CUDA(cudaSetDevice(e->eng.device));
    SV_CUDA(cudaDeviceSynchronize());
    if (slow_logits) SV_CUDA(cudaMemcpy(slow_logits, e->eng.dbg_slow_logits, AR_VOCAB * 4, cudaMemcpyDeviceToHost));
    if (hidden) SV_CUDA(cudaMemcpy(hidden, e->eng.dbg_hidden, AR_DIM * 4, cudaMemcpyDeviceToHost));
    if (fast_logits) SV_CUDA(cudaMemcpy(fast_logits, e->eng.dbg_fast_logits, 8 * AR_CB_SIZE * 4, cudaMemcpyDeviceToHost));
  });
}

// ----------------------------------------------------------------------------------------------- per-chunk loop
int svanon_stream_set_prompt(svanon_stream* sh, const int64_t* ref_content, const int32_t* ref_audio, int T,
                             const float* style, const float* timbre, int max_prompt_frames, int delay, void* stream) {
  return guarded([&] {
    SV_CHECK(sh && ref_content && ref_audio && style && timbre, "null argument");
    SV_CHECK(delay >= 0 && delay <= AR_MAX_DELAY, "delay must be in [0, 8]");
    Stream& s = sh->st;
    Args a(sh->owner, stream, (size_t)T * 48 + 32768);
    const long long* rc = (const long long*)a.in(ref_content, (size_t)T);
    const int* ra = a.in(ref_audio, (size_t)8 * T);
    const float* sv = a.in(style, 192);
    const float* tl = a.in(timbre, 32 * 128);
    s.delay = delay;
    set_prompt_copy(s, rc, ra, T, std::min(T, max_prompt_frames), sv, tl, a.st);
    s.eng->ar_prefill_prompt(s, rc, ra, T, sv, tl, a.st);   // NOTE: the reference prefills the UNtruncated prompt
    a.finish();
  });
}

int svanon_stream_setup(svanon_stream* sh, int enc_win, int dec_win, int max_seq_frames, int buffer_frames, int chunk) {
  return guarded([&] {
    SV_CHECK(sh, "null stream");
    SV_CHECK(enc_win >= 1 && enc_win <= 2048 && dec_win >= 1 && dec_win <= 1024, "window sizes out of range");
    SV_CHECK(chunk >= 1 && chunk <= 8 && chunk <= enc_win && chunk <= dec_win, "decode_chunk_frames must be in [1, 8]");
    SV_CHECK(buffer_frames >= 0 && buffer_frames < HIST_CAP / 2, "buffer_frames out of range");
    Stream& s = sh->st;
    SV_CUDA(cudaSetDevice(s.eng->device));
    SV_CUDA(cudaDeviceSynchronize());
    for (void* p : {(void*)s.wave_ring, (void*)s.wave_ring_tmp, (void*)s.ids_win_dev, (void*)s.codes_win_dev, (void*)s.wave_win_dev})
      if (p) cudaFree(p);
    s.enc_win = enc_win; s.dec_win = dec_win; s.max_seq_frames = max_seq_frames; s.buffer_frames = buffer_frames;
    s.chunk = chunk;
    const size_t nw = (size_t)enc_win * SAMPLES_PER_FRAME;
    s.wave_ring = dmalloc<float>(nw);
    s.wave_ring_tmp = dmalloc<float>(nw);
    SV_CUDA(cudaMemset(s.wave_ring, 0, nw * 4));
    s.ids_win_dev = dmalloc<long long>(enc_win);
    s.codes_win_dev = dmalloc<long long>((size_t)8 * dec_win);
    s.wave_win_dev = dmalloc<float>((size_t)dec_win * SAMPLES_PER_FRAME);
    s.n_src = 0; s.n_pred = 0; s.delay_prefilled = false;
    s.enc_state.valid = false;
    if (s.enc_stream.wave) s.eng->enc_stream_reset(s.enc_stream, nullptr);
    // incremental vocoder when the window leaves >= 15 frames of history in front of the new chunk
    s.voc_incremental = s.voc_mode != 0 && (dec_win - chunk >= 15) && (std::min(dec_win - chunk, 24) / chunk * chunk >= 15);
    s.voc_fed = 0;
    if (s.voc_incremental) s.eng->voc_state_init(s.voc, chunk);
  });
}

int svanon_stream_process_chunk(svanon_stream* sh, const float* wave_chunk, int n, const float* noise, float* wave_out,
                                void* stream) {
  return guarded([&] {
    SV_CHECK(sh && wave_chunk && wave_out, "null argument");
    Stream& s = sh->st;
    Engine& e = *s.eng;
    SV_CHECK(s.enc_win > 0, "svanon_stream_setup has not been called");
    SV_CHECK(s.ref_frames > 0, "svanon_stream_set_prompt has not been called");
    SV_CHECK(n == s.chunk * SAMPLES_PER_FRAME, "chunk must hold decode_chunk_frames * 2048 samples");
    NvtxRange nvtx_("svanon_stream_process_chunk");
    Args a(sh->owner, stream, (size_t)n * 8 + (size_t)s.chunk * 8 * AR_CB_SIZE * 4 +
                                  (size_t)(s.ref_frames + s.buffer_frames) * 64 + 65536);
    cudaStream_t st = a.st;
    const float* wc = a.in(wave_chunk, (size_t)n);
    const float* nz = a.in(noise, (size_t)s.chunk * 8 * AR_CB_SIZE);
    float* out = a.out(wave_out, (size_t)n);
    const size_t nw = (size_t)s.enc_win * SAMPLES_PER_FRAME;
    // 1. wave ring: shift left by n, append the chunk (infer_arvc.py:495-496)
    SV_CUDA(cudaMemcpyAsync(s.wave_ring_tmp, s.wave_ring + n, (nw - n) * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SV_CUDA(cudaMemcpyAsync(s.wave_ring_tmp + (nw - n), wc, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    std::swap(s.wave_ring, s.wave_ring_tmp);
    // 2. E: re-encode the whole window, keep the last `chunk` ids (:505-518)
    s.ev_valid = false;
    if (s.timing) SV_CUDA(cudaEventRecord(s.ev[0], st));
    if (s.enc_stateful) {
      if (s.enc_stream.B != 1) e.enc_stream_init(s.enc_stream, 1);
      e.enc_push(s.enc_stream, wc, n, s.chunk, s.ids_win_dev + (s.enc_win - s.chunk), s.enc_win, st);
    } else if (s.enc_state.enabled) e.enc_window_step(s.enc_state, s.wave_ring, 1, s.enc_win, s.chunk, s.ids_win_dev, st);
    else e.enc_encode(s.wave_ring, 1, (long long)nw, s.ids_win_dev, st);
    if (s.timing) SV_CUDA(cudaEventRecord(s.ev[1], st));
    if (s.n_src + s.chunk > HIST_CAP) {
      const int keep = HIST_CAP / 2;
      long long* tmp = (long long*)e.ws.base;
      SV_CUDA(cudaMemcpyAsync(tmp, s.src_hist + (s.n_src - keep), keep * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpyAsync(s.src_hist, tmp, keep * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      s.n_src = keep;
    }
    SV_CUDA(cudaMemcpyAsync(s.src_hist + s.n_src, s.ids_win_dev + (s.enc_win - s.chunk), s.chunk * sizeof(long long),
                            cudaMemcpyDeviceToDevice, st));
    s.n_src += s.chunk;
    // 3./4. warm-up phases (:519-525)
    bool silent = false;
    if (s.n_src < s.delay) {
      silent = true;
    } else if (!s.delay_prefilled && s.delay != 0) {
      e.ar_prefill_delay(s, s.src_hist + (s.n_src - s.delay), s.delay, st);
      silent = true;
    }
    if (silent) {
      SV_CUDA(cudaMemsetAsync(out, 0, (size_t)n * sizeof(float), st));
      a.finish();
      return;
    }
    // 5. A: `chunk` decode steps (:534-538)
    decode_frames(s, s.src_hist + (s.n_src - s.chunk), s.chunk, nz, st);
    if (s.timing) SV_CUDA(cudaEventRecord(s.ev[2], st));
    const int current_pos = s.pos_next - 1;
    // 6. re-prompt (:547-564)
    if (current_pos / 2 >= s.max_seq_frames) e.reprompt(s, a.h->staging, st);
    // 7. V.  Reference: vocoder over the last decode_window_frames frames, left-padded with the prompt's tail,
    //    keep the last chunk (:567-583,596).  With >= 15 frames of true history the incremental vocoder produces
    //    the same samples from the `chunk` new frames only (voc_stream.cu); smaller windows are recomputed.
    const int have = std::min(s.n_pred, s.dec_win);
    const int pad = s.dec_win - have;
    SV_CHECK(pad <= s.ref_frames, "prompt shorter than the vocoder window padding needs (the reference fails here too)");
    if (s.voc_incremental) {
      const int c = s.chunk;
      if (s.voc_fed == 0) {
        // prime with the newest prompt frames the reference would have put in front of the first window
        const int k = std::min(pad, 24) / c * c;
        SV_CHECK(k >= 15, "internal: not enough padding frames to prime the incremental vocoder");
        for (int f = 0; f < k; f += c) {
          launch_concat_cols(s.ref_audio_dev + (s.ref_frames - k + f), s.ref_frames, c, s.pred_hist, HIST_CAP, 0,
                             s.codes_win_dev, c, 8, true, st);
          e.voc_step(s.voc, s.codes_win_dev, c, s.wave_win_dev, st);
        }
      }
      launch_concat_cols(s.pred_hist + (s.n_pred - c), HIST_CAP, c, s.pred_hist, HIST_CAP, 0, s.codes_win_dev, c, 8, true, st);
      if (s.timing) SV_CUDA(cudaEventRecord(s.ev[3], st));
      e.voc_step(s.voc, s.codes_win_dev, c, out, st);
      if (s.timing) { SV_CUDA(cudaEventRecord(s.ev[4], st)); s.ev_valid = true; }
      s.voc_fed += c;
    } else {
      launch_concat_cols(s.ref_audio_dev + (s.ref_frames - pad), s.ref_frames, pad, s.pred_hist + (s.n_pred - have),
                         HIST_CAP, have, s.codes_win_dev, s.dec_win, 8, true, st);
      if (s.timing) SV_CUDA(cudaEventRecord(s.ev[3], st));
      e.voc_decode(s.codes_win_dev, s.dec_win, s.dec_win, s.wave_win_dev, st);
      if (s.timing) { SV_CUDA(cudaEventRecord(s.ev[4], st)); s.ev_valid = true; }
      // 8. tail select (:596)
      SV_CUDA(cudaMemcpyAsync(out, s.wave_win_dev + ((size_t)s.dec_win * SAMPLES_PER_FRAME - n), (size_t)n * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
    }
    a.finish();
  });
}

int svanon_stream_set_vocoder_mode(svanon_stream* sh, int incremental) {
  return guarded([&] {
    SV_CHECK(sh, "null stream");
    SV_CHECK(sh->st.enc_win == 0 || sh->st.n_pred == 0, "set the vocoder mode before streaming starts");
    sh->st.voc_mode = incremental ? 1 : 0;
    if (sh->st.enc_win > 0) {
      Stream& s = sh->st;
      s.voc_incremental = s.voc_mode != 0 && (s.dec_win - s.chunk >= 15) &&
                          (std::min(s.dec_win - s.chunk, 24) / s.chunk * s.chunk >= 15);
      if (s.voc_incremental && !s.voc.arena) s.eng->voc_state_init(s.voc, s.chunk);
    }
  });
}

int svanon_stream_set_encoder_mode(svanon_stream* sh, int incremental) {
  return guarded([&] {
    SV_CHECK(sh, "null stream");
    SV_CHECK(incremental >= 0 && incremental <= 3, "encoder mode: 0 full re-encode, 1 ring-buffer state (auto), 2 + conv history, "
                                                   "3 stateful (offline-encode semantics)");
    SV_CHECK(incremental != 3 || sh->st.n_src == 0, "switch to the stateful encoder before the first chunk");
    sh->st.enc_state.enabled = incremental != 0;
    sh->st.enc_state.tail_hist_min_streams = incremental == 2 ? 1 : 8;
    sh->st.enc_state.valid = false;
    sh->st.enc_stateful = incremental == 3;
  });
}

int svanon_stream_set_timing(svanon_stream* sh, int enable) {
  return guarded([&] {
    SV_CHECK(sh, "null stream");
    Stream& s = sh->st;
    SV_CUDA(cudaSetDevice(s.eng->device));
    if (enable)
      for (auto& e : s.ev)
        if (!e) SV_CUDA(cudaEventCreate(&e));
    s.timing = enable != 0;
    s.ev_valid = false;
  });
}

int svanon_stream_last_timing(svanon_stream* sh, float* ms) {
  return guarded([&] {
    SV_CHECK(sh && ms, "null argument");
    Stream& s = sh->st;
    SV_CHECK(s.timing && s.ev_valid, "no timed chunk available (enable timing, then process a non-warm-up chunk)");
    SV_CUDA(cudaEventSynchronize(s.ev[4]));
    SV_CUDA(cudaEventElapsedTime(&ms[0], s.ev[0], s.ev[1]));
    SV_CUDA(cudaEventElapsedTime(&ms[1], s.ev[1], s.ev[2]));
    SV_CUDA(cudaEventElapsedTime(&ms[2], s.ev[3], s.ev[4]));
  });
}

int svanon_stream_history(svanon_stream* sh, int64_t* src_content, int* n_src, int64_t* pred_codes, int* n_pred, int cap) {
  return guarded([&] {
    SV_CHECK(sh && n_src && n_pred, "null argument");
    Stream& s = sh->st;
    SV_CUDA(cudaSetDevice(s.eng->device));
    SV_CUDA(cudaDeviceSynchronize());
    const int ns = std::min(s.n_src, cap), np = std::min(s.n_pred, cap);
    *n_src = ns; *n_pred = np;
    if (src_content && ns > 0)
      SV_CUDA(cudaMemcpy(src_content, s.src_hist + (s.n_src - ns), (size_t)ns * sizeof(long long), cudaMemcpyDeviceToHost));
    if (pred_codes && np > 0) {
      std::vector<int> tmp((size_t)8 * np);
      SV_CUDA(cudaMemcpy2D(tmp.data(), (size_t)np * sizeof(int), s.pred_hist + (s.n_pred - np), HIST_CAP * sizeof(int),
                           (size_t)np * sizeof(int), 8, cudaMemcpyDeviceToHost));
      for (size_t i = 0; i < tmp.size(); ++i) pred_codes[i] = tmp[i];
    }
  });
}

}  // extern "C"
