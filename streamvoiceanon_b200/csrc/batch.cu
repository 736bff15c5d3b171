// Many concurrent streams in lock-step (BASELINE configs 3-4: "concurrent streams / GPU at RTF < 1").
//
// The reference is strictly batch-1 (max_batch_size=1, infer_arvc.py:56): N concurrent utterances are N sequential
// `process_one_chunk` calls.  A svanon_batch advances N streams by one chunk with ONE pass over the weights:
//   E  every stream's 128-frame window side by side through the same GEMMs (M = N * 512 rows),
//   A  the many-stream decode path of ar_batch.cu (1, 2 or 4 streams: the persistent kernel),
//   V  the incremental vocoder with N conv histories side by side,
// while each stream keeps its own prompt, KV cache, histories, sampler state and re-prompt schedule, so every
// stream produces exactly what it would produce alone (tests/test_gpu_batch.py).
#include "api_common.hpp"

#include <climits>

struct svanon_batch {
  svanon_engine* owner = nullptr;
  std::vector<svanon_stream*> streams;
  int enc_win = 0, dec_win = 0, max_seq_frames = 0, buffer_frames = 0, chunk = 1, delay = 0;
  float *wave_ring = nullptr, *wave_ring_tmp = nullptr;   // [n][enc_win*2048]
  long long* ids_win = nullptr;                           // [n][enc_win]
  long long* codes_win = nullptr;                         // [n][8][chunk]
  long long* step_ids = nullptr;                          // [chunk][n] content ids of this chunk, step-major
  VocState voc;
  EncWindowState enc_state;
  EncStream enc_stream;                                   // encoder mode 3 (stateful, offline-encode semantics)
  bool enc_stateful = false;
  struct SlotPtrs {                                       // static per-stream pointers (device table)
    long long* src_hist;
    int* pred_hist;
    const int* ref_audio;
    int ref_frames;
    int src_off, pred_off;                                // member's history column = batch counter + offset (merged cohorts)
  };
  std::vector<int> src_off, pred_off;                     // host copies of the offsets (0 unless the batch came from a merge)
  SlotPtrs* ptrs_dev = nullptr;
  int n_src = 0, n_pred = 0, voc_fed = 0;
  bool delay_prefilled = false;
  int ar_path = 0;                                        // 0 auto, 1 always the many-stream GEMM path
  bool timing = false, ev_valid = false;
  cudaEvent_t ev[5] = {};
  ~svanon_batch() {
    for (void* p : {(void*)wave_ring, (void*)wave_ring_tmp, (void*)ids_win, (void*)codes_win, (void*)step_ids, (void*)ptrs_dev})
      if (p) cudaFree(p);
    for (auto& e : ev)
      if (e) cudaEventDestroy(e);
  }
};

namespace {

using SlotPtrs = svanon_batch::SlotPtrs;

// src_content_codes[b][col0 + j] = ids_win[b][enc_win - chunk + j];  step_ids[j][b] = the same id
__global__ void batch_take_ids_kernel(const SlotPtrs* __restrict__ ptrs, const long long* __restrict__ ids_win, int enc_win,
                                      int chunk, int col0, long long* __restrict__ step_ids, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * chunk) return;
  const int b = i / chunk, j = i % chunk;
  const long long id = ids_win[(long long)b * enc_win + enc_win - chunk + j];
  ptrs[b].src_hist[col0 + ptrs[b].src_off + j] = id;
  step_ids[(long long)j * n + b] = id;
}

// codes_win[b][k][j] = from_ref ? ref_audio[b][k][ref_frames - back + j] : pred_hist[b][k][col0 + j]
__global__ void batch_gather_codes_kernel(const SlotPtrs* __restrict__ ptrs, long long* __restrict__ codes_win, int chunk,
                                          int from_ref, int back, int col0, int n) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 8 * chunk) return;
  const int b = i / (8 * chunk), k = (i / chunk) % 8, j = i % chunk;
  const SlotPtrs& p = ptrs[b];
  const int v = from_ref ? p.ref_audio[(long long)k * p.ref_frames + (p.ref_frames - back + j)]
                         : p.pred_hist[(long long)k * HIST_CAP + col0 + p.pred_off + j];
  codes_win[i] = v;
}

// (re)builds the device table of per-member pointers and history offsets
void upload_slot_table(svanon_batch* b, cudaStream_t st) {
  const int n = (int)b->streams.size();
  std::vector<SlotPtrs> ptrs(n);
  for (int i = 0; i < n; ++i) {
    Stream& s = b->streams[i]->st;
    ptrs[i] = {s.src_hist, s.pred_hist, s.ref_audio_dev, s.ref_frames, b->src_off[i], b->pred_off[i]};
  }
  if (!b->ptrs_dev) b->ptrs_dev = dmalloc<SlotPtrs>(n);
  SV_CUDA(cudaStreamSynchronize(st));                    // rare (setup, merge, history compaction): a blocking copy is fine
  SV_CUDA(cudaMemcpy(b->ptrs_dev, ptrs.data(), (size_t)n * sizeof(SlotPtrs), cudaMemcpyHostToDevice));
}

// Every member keeps the newest HIST_CAP / 2 columns of its source-id (src = true) or codec-id history; afterwards all
// members sit at column HIST_CAP / 2 again, so the per-member offsets of a merged batch drop to zero.
void compact_histories(svanon_batch* b, bool src, cudaStream_t st) {
  Engine& e = b->owner->eng;
  const int keep = HIST_CAP / 2;
  const int n = (int)b->streams.size();
  for (int i = 0; i < n; ++i) {
    Stream& s = b->streams[i]->st;
    if (src) {
      const int have = b->n_src + b->src_off[i];
      long long* tmp = (long long*)e.ws.base;
      SV_CUDA(cudaMemcpyAsync(tmp, s.src_hist + (have - keep), keep * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpyAsync(s.src_hist, tmp, keep * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      b->src_off[i] = 0;
    } else {
      const int have = b->n_pred + b->pred_off[i];
      int* tmp = (int*)e.ws.base;
      SV_CUDA(cudaMemcpy2DAsync(tmp, keep * sizeof(int), s.pred_hist + (have - keep), HIST_CAP * sizeof(int), keep * sizeof(int), 8,
                                cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpy2DAsync(s.pred_hist, HIST_CAP * sizeof(int), tmp, keep * sizeof(int), keep * sizeof(int), 8,
                                cudaMemcpyDeviceToDevice, st));
      b->pred_off[i] = 0;
    }
  }
  if (src) b->n_src = keep; else b->n_pred = keep;
  upload_slot_table(b, st);
}

}  // namespace

extern "C" {

int svanon_enc_encode_batch(svanon_engine* e, const float* waves, int n_utt, int64_t n_samples, int64_t* ids_out,
                            void* stream) {
  return guarded([&] {
    SV_CHECK(e && waves && ids_out && n_utt >= 1, "bad arguments");
    const size_t S = (size_t)svanon_enc_num_ids(n_samples);
    Args a(e, stream, ((size_t)n_samples * 4 + S * 8) * n_utt + 65536);
    const float* w = a.in(waves, (size_t)n_samples * n_utt);
    long long* ids = (long long*)a.out(ids_out, S * n_utt);
    e->eng.enc_encode(w, n_utt, n_samples, ids, a.st);
    a.finish();
  });
}

int svanon_voc_encode(svanon_engine* e, const float* waves, int n_utt, int64_t n_samples, int32_t* codes_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && waves && codes_out && n_utt >= 1, "bad arguments");
    const size_t T = (size_t)(n_samples / SAMPLES_PER_FRAME);
    Args a(e, stream, ((size_t)n_samples * 4 + T * 8 * 4) * n_utt + 65536);
    const float* w = a.in(waves, (size_t)n_samples * n_utt);
    int* codes = a.out(codes_out, T * 8 * n_utt);
    e->eng.voc_encode(w, n_utt, n_samples, codes, a.st);
    a.finish();
  });
}

int svanon_resample(svanon_engine* e, const float* wave, int64_t n_in, const float* kernel, int orig, int nw, int width,
                    float* out, int64_t n_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && wave && kernel && out, "null argument");
    SV_CHECK(orig >= 1 && nw >= 1 && width >= 0 && n_in >= 1, "bad resampling ratio");
    const int taps = 2 * width + orig;
    const long long full = (n_in / orig + 1) * (long long)nw;
    SV_CHECK(n_out >= 0 && n_out <= full, "n_out exceeds what the input yields (ceil(new * n_in / orig))");
    Args a(e, stream, ((size_t)n_in + (size_t)nw * taps + (size_t)n_out) * 4 + 65536);
    const float* x = a.in(wave, (size_t)n_in);
    const float* k = a.in(kernel, (size_t)nw * taps);
    float* o = a.out(out, (size_t)n_out);
    launch_resample(x, n_in, k, orig, nw, width, taps, o, n_out, a.st);
    a.finish();
  });
}

int svanon_noise_mix(svanon_engine* e, const float* x, const float* noise, int64_t n, float alpha, float* out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && x && noise && out, "null argument");
    SV_CHECK(n >= 1 && n <= (1ll << 24), "element count out of range (1 .. 2^24)");
    Args a(e, stream, (size_t)n * 12 + 65536);
    const float* xd = a.in(x, (size_t)n);
    const float* nd = a.in(noise, (size_t)n);
    float* o = a.out(out, (size_t)n);
    launch_noise_mix(xd, nd, n, alpha, o, a.st);
    a.finish();
  });
}

int svanon_ar_decode_many(svanon_stream* const* streams, int n, const int64_t* content_ids, const float* noise,
                          int32_t* codes_out, void* stream) {
  return guarded([&] {
    SV_CHECK(streams && n >= 1 && content_ids && codes_out, "bad arguments");
    svanon_engine* h = streams[0]->owner;
    Args a(h, stream, (size_t)n * (8 * AR_CB_SIZE * 4 + 256) + 65536);
    const long long* ids = (const long long*)a.in(content_ids, (size_t)n);
    const float* nz = a.in(noise, (size_t)n * 8 * AR_CB_SIZE);
    int* out = a.out(codes_out, (size_t)n * 8);
    std::vector<Stream*> ss(n);
    for (int i = 0; i < n; ++i) {
      SV_CHECK(streams[i] && streams[i]->owner == h, "streams must belong to one engine");
      ss[i] = &streams[i]->st;
      ss[i]->step_content_id = ids + i;
      ss[i]->step_cond_row = nullptr;
      ss[i]->step_noise = nz ? nz + (size_t)i * 8 * AR_CB_SIZE : nullptr;
      ss[i]->step_pred_hist = nullptr;
    }
    h->eng.ar_decode_step_gemm(ss.data(), n, a.st);
    for (int i = 0; i < n; ++i)
      SV_CUDA(cudaMemcpyAsync(out + i * 8, ss[i]->codes_dev, 8 * sizeof(int), cudaMemcpyDeviceToDevice, a.st));
    a.finish();
  });
}

int svanon_batch_create(svanon_engine* e, svanon_stream* const* streams, int n, svanon_batch** out) {
  return guarded([&] {
    SV_CHECK(e && streams && out && n >= 1, "bad arguments");
    auto* b = new svanon_batch();
    b->owner = e;
    for (int i = 0; i < n; ++i) {
      SV_CHECK(streams[i] && streams[i]->owner == e, "streams must belong to the batch's engine");
      for (int j = 0; j < i; ++j) SV_CHECK(streams[j] != streams[i], "a stream may appear only once in a batch");
      b->streams.push_back(streams[i]);
    }
    *out = b;
  });
}

void svanon_batch_destroy(svanon_batch* b) { delete b; }

int svanon_batch_set_ar_path(svanon_batch* b, int path) {
  return guarded([&] {
    SV_CHECK(b && (path == 0 || path == 1), "path: 0 auto (persistent kernel for 1/2/4 streams), 1 many-stream kernels");
    b->ar_path = path;
  });
}

int svanon_batch_setup(svanon_batch* b, int enc_win, int dec_win, int max_seq_frames, int buffer_frames, int chunk) {
  return guarded([&] {
    SV_CHECK(b, "null batch");
    SV_CHECK(enc_win >= 1 && enc_win <= 2048 && dec_win >= 1 && dec_win <= 1024, "window sizes out of range");
    SV_CHECK(chunk >= 1 && chunk <= 8 && chunk <= enc_win && chunk <= dec_win, "decode_chunk_frames must be in [1, 8]");
    SV_CHECK(buffer_frames >= 0 && buffer_frames < HIST_CAP / 2, "buffer_frames out of range");
    SV_CHECK((dec_win - chunk >= 15) && (std::min(dec_win - chunk, 24) / chunk * chunk >= 15),
             "batched streaming uses the incremental vocoder: decode_window_frames must leave >= 15 frames of history");
    Engine& e = b->owner->eng;
    SV_CUDA(cudaSetDevice(e.device));
    SV_CUDA(cudaDeviceSynchronize());
    const int n = (int)b->streams.size();
    for (void* p : {(void*)b->wave_ring, (void*)b->wave_ring_tmp, (void*)b->ids_win, (void*)b->codes_win, (void*)b->step_ids,
                    (void*)b->ptrs_dev})
      if (p) cudaFree(p);
    b->enc_win = enc_win; b->dec_win = dec_win; b->max_seq_frames = max_seq_frames; b->buffer_frames = buffer_frames;
    b->chunk = chunk;
    b->delay = b->streams[0]->st.delay;
    const size_t nw = (size_t)enc_win * SAMPLES_PER_FRAME;
    b->wave_ring = dmalloc<float>(nw * n);
    b->wave_ring_tmp = dmalloc<float>(nw * n);
    SV_CUDA(cudaMemset(b->wave_ring, 0, nw * n * 4));
    b->ids_win = dmalloc<long long>((size_t)enc_win * n);
    b->codes_win = dmalloc<long long>((size_t)8 * chunk * n);
    b->step_ids = dmalloc<long long>((size_t)chunk * n);
    std::vector<SlotPtrs> ptrs(n);
    for (int i = 0; i < n; ++i) {
      Stream& s = b->streams[i]->st;
      SV_CHECK(s.ref_frames > 0, "svanon_stream_set_prompt must be called on every stream before svanon_batch_setup");
      SV_CHECK(s.delay == b->delay, "all streams of a batch must use the same delay");
      const int pad = dec_win;       // first window is all padding
      SV_CHECK(std::min(pad, 24) <= s.ref_frames, "prompt shorter than the vocoder window padding needs");
      s.enc_win = 0;                 // the stream is driven by the batch now, not by svanon_stream_process_chunk
      s.dec_win = dec_win; s.max_seq_frames = max_seq_frames; s.buffer_frames = buffer_frames;
      s.chunk = chunk; s.n_src = 0; s.n_pred = 0; s.delay_prefilled = false;
      ptrs[i] = {s.src_hist, s.pred_hist, s.ref_audio_dev, s.ref_frames, 0, 0};
    }
    b->src_off.assign(n, 0);
    b->pred_off.assign(n, 0);
    b->ptrs_dev = dmalloc<SlotPtrs>(n);
    SV_CUDA(cudaMemcpy(b->ptrs_dev, ptrs.data(), (size_t)n * sizeof(SlotPtrs), cudaMemcpyHostToDevice));
    b->n_src = 0; b->n_pred = 0; b->voc_fed = 0; b->delay_prefilled = false;
    b->enc_state.valid = false;
    if (b->enc_stream.wave) e.enc_stream_reset(b->enc_stream, nullptr);
    e.voc_state_init(b->voc, chunk, n);
  });
}

int svanon_batch_process_chunk(svanon_batch* b, const float* wave_chunks, int n_samples, const float* noise, float* wave_out,
                               void* stream) {
  return guarded([&] {
    SV_CHECK(b && wave_chunks && wave_out, "null argument");
    SV_CHECK(b->enc_win > 0, "svanon_batch_setup has not been called");
    SV_CHECK(n_samples == b->chunk * SAMPLES_PER_FRAME, "every stream's chunk must hold decode_chunk_frames * 2048 samples");
    NvtxRange nvtx_("svanon_batch_process_chunk");
    Engine& e = b->owner->eng;
    const int n = (int)b->streams.size();
    const int c = b->chunk;
    const size_t per_noise = (size_t)c * 8 * AR_CB_SIZE;
    size_t prompt_budget = 0;
    for (auto* sh : b->streams) prompt_budget += (size_t)(sh->st.ref_frames + b->buffer_frames) * 64 + 1024;
    Args a(b->owner, stream, ((size_t)n_samples * 8 + per_noise * 4) * n + prompt_budget + 65536);
    cudaStream_t st = a.st;
    const float* wc = a.in(wave_chunks, (size_t)n_samples * n);
    const float* nz = a.in(noise, per_noise * n);
    float* out = a.out(wave_out, (size_t)n_samples * n);
    const size_t nw = (size_t)b->enc_win * SAMPLES_PER_FRAME;
    // 1. wave rings: shift left by the chunk, append (infer_arvc.py:495-496), all streams at once
    SV_CUDA(cudaMemcpy2DAsync(b->wave_ring_tmp, nw * 4, b->wave_ring + n_samples, nw * 4, (nw - n_samples) * 4, n,
                              cudaMemcpyDeviceToDevice, st));
    SV_CUDA(cudaMemcpy2DAsync(b->wave_ring_tmp + (nw - n_samples), nw * 4, wc, (size_t)n_samples * 4, (size_t)n_samples * 4, n,
                              cudaMemcpyDeviceToDevice, st));
    std::swap(b->wave_ring, b->wave_ring_tmp);
    // 2. E: all windows side by side, keep the last `chunk` ids of each (:505-518)
    b->ev_valid = false;
    if (b->timing) SV_CUDA(cudaEventRecord(b->ev[0], st));
    if (b->enc_stateful) {
      if (b->enc_stream.B != n) e.enc_stream_init(b->enc_stream, n);
      e.enc_push(b->enc_stream, wc, n_samples, c, b->ids_win + (b->enc_win - c), b->enc_win, st);
    } else if (b->enc_state.enabled) e.enc_window_step(b->enc_state, b->wave_ring, n, b->enc_win, c, b->ids_win, st);
    else e.enc_encode(b->wave_ring, n, (long long)nw, b->ids_win, st);
    if (b->timing) SV_CUDA(cudaEventRecord(b->ev[1], st));
    int max_src_off = 0, max_pred_off = 0;
    for (int i = 0; i < n; ++i) { max_src_off = std::max(max_src_off, b->src_off[i]); max_pred_off = std::max(max_pred_off, b->pred_off[i]); }
    if (b->n_src + max_src_off + c > HIST_CAP) compact_histories(b, true, st);
    launch_pdl(batch_take_ids_kernel, dim3((n * c + 127) / 128), dim3(128), 0, st, (const SlotPtrs*)b->ptrs_dev,
               (const long long*)b->ids_win, b->enc_win, c, b->n_src, b->step_ids, n);
    SV_LAUNCHED();
    b->n_src += c;
    for (int i = 0; i < n; ++i) b->streams[i]->st.n_src = b->n_src + b->src_off[i];
    // 3./4. warm-up phases (:519-525) -- the streams started together and share the delay
    bool silent = false;
    if (b->n_src < b->delay) {
      silent = true;
    } else if (!b->delay_prefilled && b->delay != 0) {
      for (auto* sh : b->streams) e.ar_prefill_delay(sh->st, sh->st.src_hist + (sh->st.n_src - b->delay), b->delay, st);
      b->delay_prefilled = true;
      silent = true;
    }
    if (silent) {
      SV_CUDA(cudaMemsetAsync(out, 0, (size_t)n_samples * n * sizeof(float), st));
      a.finish();
      return;
    }
    // 5. A: `chunk` decode steps for all streams (:534-538)
    std::vector<Stream*> ss(n);
    for (int i = 0; i < n; ++i) ss[i] = &b->streams[i]->st;
    const bool persistent = b->ar_path == 0 && (n == 1 || n == 2 || n == 4);
    for (int j = 0; j < c; ++j) {
      if (b->n_pred + max_pred_off >= HIST_CAP) {     // keep the newest half (the reference keeps at most 2048 entries, :593-594)
        compact_histories(b, false, st);
        max_pred_off = 0;
      }
      for (int i = 0; i < n; ++i) {
        Stream& s = *ss[i];
        s.step_content_id = b->step_ids + (size_t)j * n + i;
        s.step_cond_row = nullptr;
        s.step_noise = nz ? nz + (size_t)i * per_noise + (size_t)j * 8 * AR_CB_SIZE : nullptr;
        s.step_pred_hist = s.pred_hist;
        s.step_pred_col = b->n_pred + b->pred_off[i];
      }
      if (persistent) {
        e.ar_decode_step(ss.data(), n, st);
        for (int i = 0; i < n; ++i) launch_append_codes(ss[i]->codes_dev, ss[i]->pred_hist, HIST_CAP, b->n_pred + b->pred_off[i], st);
      } else {
        e.ar_decode_step_gemm(ss.data(), n, st);
      }
      b->n_pred += 1;
      for (int i = 0; i < n; ++i) ss[i]->n_pred = b->n_pred + b->pred_off[i];
    }
    if (b->timing) SV_CUDA(cudaEventRecord(b->ev[2], st));
    // 6. re-prompt (:547-564): per stream, on its own schedule (prompt lengths differ)
    {
      std::vector<Stream*> due;
      for (Stream* s : ss)
        if ((s->pos_next - 1) / 2 >= b->max_seq_frames) due.push_back(s);
      if (!due.empty()) e.reprompt_many(due.data(), (int)due.size(), a.h->staging, st);     // one pass over the weights for all
    }
    // 7. V: incremental vocoder on the new frames of every stream; primed with the prompt's newest frames, which
    //    is what the reference puts in front of the first windows (:567-571)
    if (b->voc_fed == 0) {
      const int k = std::min(b->dec_win - c, 24) / c * c;
      for (int f = 0; f < k; f += c) {
        launch_pdl(batch_gather_codes_kernel, dim3((n * 8 * c + 127) / 128), dim3(128), 0, st, (const SlotPtrs*)b->ptrs_dev,
                   b->codes_win, c, 1, k - f, 0, n);
        SV_LAUNCHED();
        e.voc_step(b->voc, b->codes_win, c, out, st, (long long)8 * c);
      }
    }
    launch_pdl(batch_gather_codes_kernel, dim3((n * 8 * c + 127) / 128), dim3(128), 0, st, (const SlotPtrs*)b->ptrs_dev,
               b->codes_win, c, 0, 0, b->n_pred - c, n);
    SV_LAUNCHED();
    if (b->timing) SV_CUDA(cudaEventRecord(b->ev[3], st));
    e.voc_step(b->voc, b->codes_win, c, out, st, (long long)8 * c);
    if (b->timing) { SV_CUDA(cudaEventRecord(b->ev[4], st)); b->ev_valid = true; }
    b->voc_fed += c;
    a.finish();
  });
}

// Two cohorts that have both left their warm-up (delay prefilled, vocoder primed) become ONE lock-step batch: the members
// of `a` followed by the members of `b`, each with the state it had -- wave ring, encoder window state (or stateful-encoder
// state), vocoder histories -- so every stream keeps producing exactly what it produces alone, and the step after the merge
// makes one pass over the weights for everybody.  `a` and `b` are left empty (destroy them).  The members' history columns
// differ (the cohorts started at different chunks): the merged batch counts like `a` and carries a per-member offset.
int svanon_batch_merge(svanon_batch* a, svanon_batch* b, svanon_batch** out, void* stream) {
  return guarded([&] {
    SV_CHECK(a && b && out && a != b, "bad arguments");
    SV_CHECK(a->owner == b->owner, "batches of different engines");
    SV_CHECK(a->enc_win > 0 && b->enc_win > 0, "svanon_batch_setup has not been called on both batches");
    SV_CHECK(a->enc_win == b->enc_win && a->dec_win == b->dec_win && a->max_seq_frames == b->max_seq_frames &&
                 a->buffer_frames == b->buffer_frames && a->chunk == b->chunk && a->delay == b->delay,
             "batches with different stream settings cannot merge");
    SV_CHECK(a->enc_stateful == b->enc_stateful && a->enc_state.enabled == b->enc_state.enabled &&
                 a->enc_state.tail_hist_min_streams == b->enc_state.tail_hist_min_streams, "batches with different encoder modes");
    SV_CHECK(a->voc_fed > 0 && b->voc_fed > 0 && (a->delay == 0 || (a->delay_prefilled && b->delay_prefilled)),
             "both batches must be past their warm-up chunks (delay prefilled, first frame decoded)");
    Engine& e = a->owner->eng;
    SV_CUDA(cudaSetDevice(e.device));
    cudaStream_t st = (cudaStream_t)stream;
    const int na = (int)a->streams.size(), nb = (int)b->streams.size(), n = na + nb;
    auto* c = new svanon_batch();
    std::unique_ptr<svanon_batch> guard(c);
    c->owner = a->owner;
    c->streams = a->streams;
    c->streams.insert(c->streams.end(), b->streams.begin(), b->streams.end());
    c->enc_win = a->enc_win; c->dec_win = a->dec_win; c->max_seq_frames = a->max_seq_frames; c->buffer_frames = a->buffer_frames;
    c->chunk = a->chunk; c->delay = a->delay; c->ar_path = a->ar_path;
    c->n_src = a->n_src; c->n_pred = a->n_pred; c->voc_fed = std::min(a->voc_fed, b->voc_fed); c->delay_prefilled = true;
    c->src_off = a->src_off; c->pred_off = a->pred_off;
    for (int i = 0; i < nb; ++i) {
      c->src_off.push_back(b->n_src + b->src_off[i] - a->n_src);
      c->pred_off.push_back(b->n_pred + b->pred_off[i] - a->n_pred);
    }
    int lo_src = 0, lo_pred = 0;                      // keep every offset >= 0: count like the member that is furthest behind
    for (int i = 0; i < n; ++i) { lo_src = std::min(lo_src, c->src_off[i]); lo_pred = std::min(lo_pred, c->pred_off[i]); }
    c->n_src += lo_src; c->n_pred += lo_pred;
    for (int i = 0; i < n; ++i) { c->src_off[i] -= lo_src; c->pred_off[i] -= lo_pred; }
    const size_t nw = (size_t)c->enc_win * SAMPLES_PER_FRAME;
    c->wave_ring = dmalloc<float>(nw * n);
    c->wave_ring_tmp = dmalloc<float>(nw * n);
    c->ids_win = dmalloc<long long>((size_t)c->enc_win * n);
    c->codes_win = dmalloc<long long>((size_t)8 * c->chunk * n);
    c->step_ids = dmalloc<long long>((size_t)c->chunk * n);
    auto cat = [&](float* dst, const float* pa, const float* pb, size_t per) {      // [na][per] ++ [nb][per]
      SV_CUDA(cudaMemcpyAsync(dst, pa, per * na * sizeof(float), cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpyAsync(dst + per * na, pb, per * nb * sizeof(float), cudaMemcpyDeviceToDevice, st));
    };
    cat(c->wave_ring, a->wave_ring, b->wave_ring, nw);
    // ---- encoder state
    c->enc_stateful = a->enc_stateful;
    c->enc_state.enabled = a->enc_state.enabled;
    c->enc_state.tail_hist_min_streams = a->enc_state.tail_hist_min_streams;
    if (c->enc_stateful) {
      SV_CHECK(a->enc_stream.wave && b->enc_stream.wave && a->enc_stream.pos > 0 && b->enc_stream.pos > 0, "stateful encoder state missing");
      e.enc_stream_init(c->enc_stream, n);
      EncStream &ea = a->enc_stream, &eb = b->enc_stream, &ec = c->enc_stream;
      cat(ec.wave, ea.wave, eb.wave, ENC_STREAM_WAVE);
      const size_t per_kv = (size_t)ENC_HEADS * ENC_RING * HEAD_DIM;
      for (int l = 0; l < ENC_LAYERS; ++l) {
        cat(ec.kc + (size_t)l * n * per_kv, ea.kc + (size_t)l * na * per_kv, eb.kc + (size_t)l * nb * per_kv, per_kv);
        cat(ec.vc + (size_t)l * n * per_kv, ea.vc + (size_t)l * na * per_kv, eb.vc + (size_t)l * nb * per_kv, per_kv);
      }
      // positions: the merged state counts like `a`; b's members carry the difference (ring slot = position % ring size)
      ec.pos = ea.pos;
      for (int i = 0; i < na; ++i) ec.off[i] = ea.off[i];
      for (int i = 0; i < nb; ++i) ec.off[na + i] = eb.pos + eb.off[i] - ea.pos;
      SV_CUDA(cudaMemcpyAsync(ec.off_dev, ec.off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, st));
      const int dims[4] = {128, 256, 384, 512}, depths[4] = {3, 3, 9, 3};
      cat(ec.hist.mel, ea.hist.mel, eb.hist.mel, (size_t)6 * N_MELS);
      int j = 0;
      for (int s4 = 0; s4 < 4; ++s4)
        for (int d = 0; d < depths[s4]; ++d, ++j) cat(ec.hist.blk[j], ea.hist.blk[j], eb.hist.blk[j], (size_t)6 * dims[s4]);
      for (int d = 0; d < 2; ++d, ++j) cat(ec.hist.blk[j], ea.hist.blk[j], eb.hist.blk[j], (size_t)6 * 512);
    } else if (a->enc_state.valid && b->enc_state.valid && a->enc_state.S == b->enc_state.S &&
               a->enc_state.hist_valid == b->enc_state.hist_valid) {
      EncWindowState &sa = a->enc_state, &sb = b->enc_state, &sc = c->enc_state;
      const int S = sa.S;
      for (auto& p : sc.xt) SV_CUDA(cudaMalloc(&p, (size_t)n * S * ENC_DIM * sizeof(float)));
      sc.B = n; sc.S = S; sc.cur = 0; sc.valid = true;
      cat(sc.xt[0], sa.xt[sa.cur], sb.xt[sb.cur], (size_t)S * ENC_DIM);
      sc.hist_valid = sa.hist_valid;
      if (sc.hist_valid) {
        sc.hist.alloc(n);
        const int dims[4] = {128, 256, 384, 512}, depths[4] = {3, 3, 9, 3};
        cat(sc.hist.mel, sa.hist.mel, sb.hist.mel, (size_t)6 * N_MELS);
        int j = 0;
        for (int s4 = 0; s4 < 4; ++s4)
          for (int d = 0; d < depths[s4]; ++d, ++j) cat(sc.hist.blk[j], sa.hist.blk[j], sb.hist.blk[j], (size_t)6 * dims[s4]);
        for (int d = 0; d < 2; ++d, ++j) cat(sc.hist.blk[j], sa.hist.blk[j], sb.hist.blk[j], (size_t)6 * 512);
      }
    } else {
      c->enc_state.valid = false;                      // the next step re-encodes the whole window once (same ids)
    }
    // ---- vocoder histories
    e.voc_state_init(c->voc, c->chunk, n);
    voc_state_concat(c->voc, a->voc, b->voc, st);
    if (a->timing || b->timing) {
      for (auto& ev : c->ev) SV_CUDA(cudaEventCreate(&ev));
      c->timing = true;
    }
    upload_slot_table(c, st);                         // synchronises: the copies above are complete
    for (int i = 0; i < n; ++i) {
      Stream& s = c->streams[i]->st;
      s.n_src = c->n_src + c->src_off[i];
      s.n_pred = c->n_pred + c->pred_off[i];
    }
    a->streams.clear(); a->enc_win = 0;
    b->streams.clear(); b->enc_win = 0;
    *out = guard.release();
  });
}

// The members keep[0 .. n_keep) of `a` (indices into a's member order, strictly increasing) as a batch of their own, each with
// the state it had: what a server does when streams of a cohort have left, so that the steps afterwards compute for the
// remaining streams only.  Same state moves as svanon_batch_merge, gathered per member instead of concatenated; the rings of
// steady-state layer inputs (ConvStackRings) start empty in the new batch, as after a merge.  `a` is left without members
// (destroy it); the streams that were left out belong to no batch any more (their batch-level state -- wave ring, encoder and
// vocoder state -- is gone with `a`).
int svanon_batch_select(svanon_batch* a, const int* keep, int n_keep, svanon_batch** out, void* stream) {
  return guarded([&] {
    SV_CHECK(a && keep && out && n_keep >= 1, "bad arguments");
    SV_CHECK(a->enc_win > 0, "svanon_batch_setup has not been called");
    const int na = (int)a->streams.size(), n = n_keep;
    SV_CHECK(n <= na, "more members to keep than the batch has");
    for (int i = 0; i < n; ++i)
      SV_CHECK(keep[i] >= 0 && keep[i] < na && (i == 0 || keep[i] > keep[i - 1]), "keep: strictly increasing member indices");
    SV_CHECK(a->voc_fed > 0 && (a->delay == 0 || a->delay_prefilled),
             "the batch must be past its warm-up chunks (delay prefilled, first frame decoded)");
    Engine& e = a->owner->eng;
    SV_CUDA(cudaSetDevice(e.device));
    cudaStream_t st = (cudaStream_t)stream;
    auto* c = new svanon_batch();
    std::unique_ptr<svanon_batch> guard(c);
    c->owner = a->owner;
    for (int i = 0; i < n; ++i) c->streams.push_back(a->streams[keep[i]]);
    c->enc_win = a->enc_win; c->dec_win = a->dec_win; c->max_seq_frames = a->max_seq_frames; c->buffer_frames = a->buffer_frames;
    c->chunk = a->chunk; c->delay = a->delay; c->ar_path = a->ar_path;
    c->n_src = a->n_src; c->n_pred = a->n_pred; c->voc_fed = a->voc_fed; c->delay_prefilled = true;
    int lo_src = INT_MAX, lo_pred = INT_MAX;          // count like the kept member that is furthest behind
    for (int i = 0; i < n; ++i) { lo_src = std::min(lo_src, a->src_off[keep[i]]); lo_pred = std::min(lo_pred, a->pred_off[keep[i]]); }
    c->n_src += lo_src; c->n_pred += lo_pred;
    for (int i = 0; i < n; ++i) {
      c->src_off.push_back(a->src_off[keep[i]] - lo_src);
      c->pred_off.push_back(a->pred_off[keep[i]] - lo_pred);
    }
    const size_t nw = (size_t)c->enc_win * SAMPLES_PER_FRAME;
    c->wave_ring = dmalloc<float>(nw * n);
    c->wave_ring_tmp = dmalloc<float>(nw * n);
    c->ids_win = dmalloc<long long>((size_t)c->enc_win * n);
    c->codes_win = dmalloc<long long>((size_t)8 * c->chunk * n);
    c->step_ids = dmalloc<long long>((size_t)c->chunk * n);
    auto take = [&](float* dst, const float* src, size_t per) {      // dst [n][per] <- rows keep[] of src; runs travel in one copy
      for (int i = 0; i < n;) {
        int j = i + 1;
        while (j < n && keep[j] == keep[j - 1] + 1) ++j;
        SV_CUDA(cudaMemcpyAsync(dst + per * i, src + per * keep[i], per * (size_t)(j - i) * sizeof(float), cudaMemcpyDeviceToDevice, st));
        i = j;
      }
    };
    take(c->wave_ring, a->wave_ring, nw);
    // ---- encoder state
    c->enc_stateful = a->enc_stateful;
    c->enc_state.enabled = a->enc_state.enabled;
    c->enc_state.tail_hist_min_streams = a->enc_state.tail_hist_min_streams;
    const int dims[4] = {128, 256, 384, 512}, depths[4] = {3, 3, 9, 3};
    auto take_hist = [&](ConvStackHist& hc, ConvStackHist& ha) {
      take(hc.mel, ha.mel, (size_t)6 * N_MELS);
      int j = 0;
      for (int s4 = 0; s4 < 4; ++s4)
        for (int d = 0; d < depths[s4]; ++d, ++j) take(hc.blk[j], ha.blk[j], (size_t)6 * dims[s4]);
      for (int d = 0; d < 2; ++d, ++j) take(hc.blk[j], ha.blk[j], (size_t)6 * 512);
    };
    if (c->enc_stateful) {
      SV_CHECK(a->enc_stream.wave && a->enc_stream.pos > 0 && a->enc_stream.B == na, "stateful encoder state missing");
      e.enc_stream_init(c->enc_stream, n);
      EncStream &ea = a->enc_stream, &ec = c->enc_stream;
      take(ec.wave, ea.wave, ENC_STREAM_WAVE);
      const size_t per_kv = (size_t)ENC_HEADS * ENC_RING * HEAD_DIM;
      for (int l = 0; l < ENC_LAYERS; ++l) {
        take(ec.kc + (size_t)l * n * per_kv, ea.kc + (size_t)l * na * per_kv, per_kv);
        take(ec.vc + (size_t)l * n * per_kv, ea.vc + (size_t)l * na * per_kv, per_kv);
      }
      ec.pos = ea.pos;
      for (int i = 0; i < n; ++i) ec.off[i] = ea.off[keep[i]];
      SV_CUDA(cudaMemcpyAsync(ec.off_dev, ec.off.data(), (size_t)n * sizeof(long long), cudaMemcpyHostToDevice, st));
      take_hist(ec.hist, ea.hist);
    } else if (a->enc_state.valid && a->enc_state.B == na) {
      EncWindowState &sa = a->enc_state, &sc = c->enc_state;
      const int S = sa.S;
      for (auto& p : sc.xt) SV_CUDA(cudaMalloc(&p, (size_t)n * S * ENC_DIM * sizeof(float)));
      sc.B = n; sc.S = S; sc.cur = 0; sc.valid = true;
      take(sc.xt[0], sa.xt[sa.cur], (size_t)S * ENC_DIM);
      // fewer members than the conv-history mode wants: the next step runs the tail span instead, the history is not needed
      sc.hist_valid = sa.hist_valid && sa.hist.B == na && sc.tail_hist_min_streams > 0 && n >= sc.tail_hist_min_streams;
      if (sc.hist_valid) {
        sc.hist.alloc(n);
        take_hist(sc.hist, sa.hist);
      }
    } else {
      c->enc_state.valid = false;                      // the next step re-encodes the whole window once (same ids)
    }
    // ---- vocoder histories
    e.voc_state_init(c->voc, c->chunk, n);
    voc_state_select(c->voc, a->voc, keep, n, st);
    if (a->timing) {
      for (auto& ev : c->ev) SV_CUDA(cudaEventCreate(&ev));
      c->timing = true;
    }
    upload_slot_table(c, st);                         // synchronises: the copies above are complete
    for (int i = 0; i < n; ++i) {
      Stream& s = c->streams[i]->st;
      s.n_src = c->n_src + c->src_off[i];
      s.n_pred = c->n_pred + c->pred_off[i];
    }
    a->streams.clear(); a->enc_win = 0;
    *out = guard.release();
  });
}

int svanon_batch_set_encoder_mode(svanon_batch* b, int incremental) {
  return guarded([&] {
    SV_CHECK(b, "null batch");
    SV_CHECK(incremental >= 0 && incremental <= 3, "encoder mode: 0 full re-encode, 1 ring-buffer state (auto), 2 + conv history, "
                                                   "3 stateful (offline-encode semantics)");
    SV_CHECK(incremental != 3 || b->n_src == 0, "switch to the stateful encoder before the first chunk");
    b->enc_state.enabled = incremental != 0;
    b->enc_state.tail_hist_min_streams = incremental == 2 ? 1 : 8;
    b->enc_state.valid = false;
    b->enc_stateful = incremental == 3;
  });
}

int svanon_batch_set_timing(svanon_batch* b, int enable) {
  return guarded([&] {
    SV_CHECK(b, "null batch");
    SV_CUDA(cudaSetDevice(b->owner->eng.device));
    if (enable)
      for (auto& e : b->ev)
        if (!e) SV_CUDA(cudaEventCreate(&e));
    b->timing = enable != 0;
    b->ev_valid = false;
  });
}

int svanon_batch_last_timing(svanon_batch* b, float* ms) {
  return guarded([&] {
    SV_CHECK(b && ms, "null argument");
    SV_CHECK(b->timing && b->ev_valid, "no timed chunk available (enable timing, then process a non-warm-up chunk)");
    SV_CUDA(cudaEventSynchronize(b->ev[4]));
    SV_CUDA(cudaEventElapsedTime(&ms[0], b->ev[0], b->ev[1]));
    SV_CUDA(cudaEventElapsedTime(&ms[1], b->ev[1], b->ev[2]));
    SV_CUDA(cudaEventElapsedTime(&ms[2], b->ev[3], b->ev[4]));
  });
}

}  // extern "C"
