// Stateful content encoder (`enc_push_chunk`; section 7 step 6b) and the stand-alone stateful vocoder
// entry (`voc_push_frames`).
//
// `svanon_enc_push_chunk` feeds the NEW samples of a stream only and returns the ids of the new content frames.  Its
// parity target is the reference's OFFLINE `FireflyArchitecture.encode()` (modules/vqgan/modules/firefly_encoder.py:
// 553-566) on the stream's whole prefix: every conv of the tokenizer is left-pad-only causal and the window transformer is
// causal with a 512-token look-back (windowed_transformer.py:291-317), so token t of the offline encode depends on samples
// [0, (t + 1) * 2048) only and can be produced incrementally from
//   * per-layer causal-conv history (ConvStackHist: the newest 6 rows in front of the stem and of each of the 18 + 2
//     ConvNeXt blocks; the 1536-sample STFT look-back is kept as a wave tail), and
//   * a per-stream K/V ring of the 8 transformer layers (520 slots: the 512-token window plus the <= 8 tokens of a push).
// 0.23 GFLOP per frame instead of the 29 GFLOP window re-encode.  This is NOT what the reference's streaming loop computes
// (it re-encodes a 128-frame window that starts from zero padding every chunk, SURVEY finding 4): the loop keeps the
// window semantics by default and offers this as encoder mode 3, documented as "offline-encode semantics".
//
// Keys are stored UNROTATED; RoPE is applied to q and to every key when it is read, with positions taken relative to a
// base that moves in steps of 8192 frames: for the first 8192 + 511 frames (6.7 minutes) that is exactly the offline
// encode's absolute position (the reference's own table stops at 2048 frames = 95 s, where its `encode()` fails); beyond,
// positions wrap consistently for a query and all its keys, so a stream can run for any length.
#include "api_common.hpp"

namespace svanon {

namespace {

constexpr int ER_WARPS = 4;

__device__ __forceinline__ float er_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ring[l][b][h][slot][64] <- k / v of the push's rows (row r = b * c + j sits at position pos0 + j)
__global__ void __launch_bounds__(256) enc_ring_append_kernel(const float* __restrict__ qkv, float* __restrict__ kc,
                                                              float* __restrict__ vc, int c, long long pos0,
                                                              const long long* __restrict__ off) {
  pdl_trigger();
  pdl_wait();
  const int r = blockIdx.x, b = r / c, j = r % c;
  const int slot = (int)((pos0 + off[b] + j) % ENC_RING);
  const float* row = qkv + (long long)r * 3 * ENC_DIM;
  for (int i = threadIdx.x; i < ENC_DIM; i += blockDim.x) {
    const int h = i / HEAD_DIM, d = i % HEAD_DIM;
    const long long o = (((long long)b * ENC_HEADS + h) * ENC_RING + slot) * HEAD_DIM + d;
    kc[o] = row[ENC_DIM + i];
    vc[o] = row[2 * ENC_DIM + i];
  }
}

// Attention of one (head, row) over the ring: keys at positions max(0, p - 511) .. p, RoPE on the fly.
// grid (ENC_HEADS, rows); 4 warps walk the keys (lane = interleaved pair), online softmax per warp, merged in warp order.
__global__ void __launch_bounds__(ER_WARPS * 32) enc_attn_ring_kernel(const float* __restrict__ qkv, const float* __restrict__ kc,
                                                                      const float* __restrict__ vc, const float* __restrict__ rope,
                                                                      float* __restrict__ y, int c, long long pos0,
                                                                      const long long* __restrict__ off) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[ER_WARPS][2 + HEAD_DIM];
  const int h = blockIdx.x, r = blockIdx.y, b = r / c, j = r % c;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long p = pos0 + off[b] + j;
  const long long lo = p >= ENC_WINDOW ? p - (ENC_WINDOW - 1) : 0;
  const long long base = lo >= ENC_POS_PERIOD ? (lo / ENC_POS_PERIOD) * ENC_POS_PERIOD : 0;
  const float* kh = kc + ((long long)b * ENC_HEADS + h) * ENC_RING * HEAD_DIM;
  const float* vh = vc + ((long long)b * ENC_HEADS + h) * ENC_RING * HEAD_DIM;
  const float2 qv = *(reinterpret_cast<const float2*>(qkv + (long long)r * 3 * ENC_DIM + h * HEAD_DIM) + lane);
  const float2 qcs = __ldg(reinterpret_cast<const float2*>(rope + ((p - base) * (HEAD_DIM / 2) + lane) * 2));
  const float2 q = make_float2(qv.x * qcs.x - qv.y * qcs.y, qv.y * qcs.x + qv.x * qcs.y);
  float m = -INFINITY, l = 0.f;
  float2 acc = make_float2(0.f, 0.f);
  for (long long key = lo + warp; key <= p; key += ER_WARPS) {
    const int slot = (int)(key % ENC_RING);
    const float2 kr = __ldcg(reinterpret_cast<const float2*>(kh + slot * HEAD_DIM) + lane);
    const float2 vv = __ldcg(reinterpret_cast<const float2*>(vh + slot * HEAD_DIM) + lane);
    const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((key - base) * (HEAD_DIM / 2) + lane) * 2));
    const float2 k = make_float2(kr.x * cs.x - kr.y * cs.y, kr.y * cs.x + kr.x * cs.y);
    const float s = er_warp_sum(q.x * k.x + q.y * k.y) * 0.125f;
    const float mn = fmaxf(m, s);
    const float corr = expf(m - mn), pr = expf(s - mn);
    l = l * corr + pr;
    acc.x = acc.x * corr + pr * vv.x;
    acc.y = acc.y * corr + pr * vv.y;
    m = mn;
  }
  if (lane == 0) { sm[warp][0] = m; sm[warp][1] = l; }
  sm[warp][2 + 2 * lane] = acc.x;
  sm[warp][3 + 2 * lane] = acc.y;
  __syncthreads();
  if (warp == 0) {
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < ER_WARPS; ++w) mm = fmaxf(mm, sm[w][0]);
    float ll = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
    for (int w = 0; w < ER_WARPS; ++w) {
      const float cw = (sm[w][0] == -INFINITY) ? 0.f : expf(sm[w][0] - mm);
      ll += sm[w][1] * cw; ax += sm[w][2 + 2 * lane] * cw; ay += sm[w][3 + 2 * lane] * cw;
    }
    const float inv = 1.f / ll;
    *(reinterpret_cast<float2*>(y + (long long)r * ENC_DIM + h * HEAD_DIM) + lane) = make_float2(ax * inv, ay * inv);
  }
}

}  // namespace

EncStream::~EncStream() {
  for (void* p : {(void*)wave, (void*)kc, (void*)vc, (void*)off_dev})
    if (p) cudaFree(p);
}

void Engine::enc_stream_init(EncStream& es, int B) {
  SV_CHECK(finalized[MODEL_TOKENIZER], "tokenizer weights not finalized");
  SV_CHECK(enc_rope_stream, "the tokenizer was loaded without its long RoPE table (quantizer.pre_module.freqs_cis_stream)");
  SV_CHECK(B >= 1, "streams per encoder state");
  for (void* p : {(void*)es.wave, (void*)es.kc, (void*)es.vc, (void*)es.off_dev})
    if (p) cudaFree(p);
  es.B = B;
  es.pos = 0;
  es.off.assign(B, 0);
  SV_CUDA(cudaMalloc(&es.off_dev, (size_t)B * sizeof(long long)));
  SV_CUDA(cudaMemset(es.off_dev, 0, (size_t)B * sizeof(long long)));
  const size_t nw = (size_t)B * ENC_STREAM_WAVE;
  const size_t nkv = (size_t)ENC_LAYERS * B * ENC_HEADS * ENC_RING * HEAD_DIM;
  SV_CUDA(cudaMalloc(&es.wave, nw * sizeof(float)));
  SV_CUDA(cudaMalloc(&es.kc, nkv * sizeof(float)));
  SV_CUDA(cudaMalloc(&es.vc, nkv * sizeof(float)));
  SV_CUDA(cudaMemset(es.wave, 0, nw * sizeof(float)));
  SV_CUDA(cudaMemset(es.kc, 0, nkv * sizeof(float)));
  SV_CUDA(cudaMemset(es.vc, 0, nkv * sizeof(float)));
  es.hist.alloc(B);
}

void Engine::enc_stream_reset(EncStream& es, cudaStream_t st) {
  SV_CHECK(es.wave, "encoder stream not initialised");
  es.pos = 0;                                   // the next push starts an utterance: zero left context everywhere
  es.off.assign(es.B, 0);
  SV_CUDA(cudaMemsetAsync(es.off_dev, 0, (size_t)es.B * sizeof(long long), st));
  SV_CUDA(cudaMemsetAsync(es.wave, 0, (size_t)es.B * ENC_STREAM_WAVE * sizeof(float), st));
}

// c new content frames per stream: wave_chunk rows [c * 2048] (row b at wave_chunk + b * pitch) -> ids[b * ids_ld + j]
void Engine::enc_push(EncStream& es, const float* wave_chunk, long long pitch, int c, long long* ids, long long ids_ld,
                      cudaStream_t st) {
  NvtxRange nvtx_("svanon:E push (stateful)");
  SV_CHECK(es.wave, "encoder stream not initialised");
  SV_CHECK(c >= 1 && c <= 8, "1..8 content frames per push");
  const int B = es.B, M = B * c;
  const long long n = (long long)c * SAMPLES_PER_FRAME;
  const int LEAD = N_FFT - HOP;                 // 1536 samples of STFT look-back
  // wave staging per stream: [tail of the previous push (1536) | new samples]
  SV_CUDA(cudaMemcpy2DAsync(es.wave + LEAD, (size_t)ENC_STREAM_WAVE * sizeof(float), wave_chunk, (size_t)pitch * sizeof(float),
                            (size_t)n * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  ws.ensure(((size_t)(n / HOP) * 14000 * B + (size_t)n * B + (size_t)M * (ENC_DIM * 8 + ENC_INTER * 3) + (4u << 20)) * sizeof(float));
  ws.reset();
  float* xt = ws.alloc_f((long long)M * ENC_DIM);
  const float* src = es.wave + LEAD;
  const long long wpitch = ENC_STREAM_WAVE;
  enc_conv_stack(tok_cs, &src, &wpitch, 1, B, n, xt, st, &es.hist, es.pos == 0 ? 1 : 2);
  // keep the newest 1536 samples in front for the next push (staged through the workspace: the ranges overlap)
  {
    float* tail = ws.alloc_f((long long)B * LEAD);
    SV_CUDA(cudaMemcpy2DAsync(tail, (size_t)LEAD * sizeof(float), es.wave + n, (size_t)ENC_STREAM_WAVE * sizeof(float),
                              (size_t)LEAD * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    SV_CUDA(cudaMemcpy2DAsync(es.wave, (size_t)ENC_STREAM_WAVE * sizeof(float), tail, (size_t)LEAD * sizeof(float),
                              (size_t)LEAD * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
  }
  // WindowLimitedTransformer (windowed_transformer.py:337-354) for the new tokens against the K/V ring
  float* nrm = ws.alloc_f((long long)M * ENC_DIM);
  float* qkv = ws.alloc_f((long long)M * 3 * ENC_DIM);
  float* y = ws.alloc_f((long long)M * ENC_DIM);
  float* h13 = ws.alloc_f((long long)M * 2 * ENC_INTER);
  float* gbuf = ws.alloc_f((long long)M * ENC_INTER);
  const size_t layer_kv = (size_t)B * ENC_HEADS * ENC_RING * HEAD_DIM;
  for (int l = 0; l < ENC_LAYERS; ++l) {
    const EncLayerW& L = enc_layers[l];
    launch_rmsnorm(xt, nrm, L.attn_norm, M, ENC_DIM, 1e-5f, st);
    GemmParams p;
    p.A = nrm; p.W = L.wqkv; p.C = qkv; p.M = M; p.N = 3 * ENC_DIM; p.K = ENC_DIM; p.lda = ENC_DIM; p.ldc = 3 * ENC_DIM;
    launch_gemm(p, st);
    launch_pdl(enc_ring_append_kernel, dim3(M), dim3(256), 0, st, (const float*)qkv, es.kc + l * layer_kv, es.vc + l * layer_kv, c,
               es.pos, (const long long*)es.off_dev);
    SV_LAUNCHED();
    launch_pdl(enc_attn_ring_kernel, dim3(ENC_HEADS, M), dim3(ER_WARPS * 32), 0, st, (const float*)qkv,
               (const float*)(es.kc + l * layer_kv), (const float*)(es.vc + l * layer_kv), enc_rope_stream, y, c, es.pos,
               (const long long*)es.off_dev);
    SV_LAUNCHED();
    GemmParams po;
    po.A = y; po.W = L.wo; po.C = xt; po.gamma = L.ls_attn; po.residual = xt; po.M = M; po.N = ENC_DIM; po.K = ENC_DIM;
    po.lda = ENC_DIM; po.ldc = ENC_DIM; po.ldr = ENC_DIM;
    launch_gemm(po, st);
    launch_rmsnorm(xt, nrm, L.ffn_norm, M, ENC_DIM, 1e-5f, st);
    GemmParams p1;
    p1.A = nrm; p1.W = L.w1; p1.C = h13; p1.M = M; p1.N = ENC_INTER; p1.K = ENC_DIM; p1.lda = ENC_DIM; p1.ldc = 2 * ENC_INTER;
    GemmParams p13[2] = {p1, p1};
    p13[1].W = L.w3; p13[1].C = h13 + ENC_INTER;
    launch_gemm(p13, 2, st);
    launch_silu_mul(h13, gbuf, M, ENC_INTER, st);
    GemmParams p2;
    p2.A = gbuf; p2.W = L.w2; p2.C = xt; p2.gamma = L.ls_ffn; p2.residual = xt; p2.M = M; p2.N = ENC_DIM; p2.K = ENC_INTER;
    p2.lda = ENC_INTER; p2.ldc = ENC_DIM; p2.ldr = ENC_DIM;
    launch_gemm(p2, st);
  }
  launch_rmsnorm(xt, nrm, enc_norm_w, M, ENC_DIM, 1e-5f, st);
  long long* ids_tmp = reinterpret_cast<long long*>(ws.alloc_f((long long)M * 2 + 4));
  launch_bsq(nrm, bsq_w, bsq_b, ids_tmp, M, st);
  SV_CUDA(cudaMemcpy2DAsync(ids, (size_t)ids_ld * sizeof(long long), ids_tmp, (size_t)c * sizeof(long long),
                            (size_t)c * sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
  es.pos += c;
}

}  // namespace svanon

// ------------------------------------------------------------------------------------------------------- C ABI
struct svanon_enc_stream {
  svanon_engine* owner = nullptr;
  EncStream st;
};
struct svanon_voc_stream {
  svanon_engine* owner = nullptr;
  VocState st;
  long long* codes_dev = nullptr;      // [n][8][c] int64 staging
  ~svanon_voc_stream() {
    if (codes_dev) cudaFree(codes_dev);
  }
};

extern "C" {

int svanon_enc_stream_create(svanon_engine* e, int n_streams, svanon_enc_stream** out) {
  return guarded([&] {
    SV_CHECK(e && out && n_streams >= 1, "bad arguments");
    SV_CUDA(cudaSetDevice(e->eng.device));
    auto* h = new svanon_enc_stream();
    h->owner = e;
    try {
      e->eng.enc_stream_init(h->st, n_streams);
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

void svanon_enc_stream_destroy(svanon_enc_stream* s) { delete s; }

int svanon_enc_stream_reset(svanon_enc_stream* s, void* stream) {
  return guarded([&] {
    SV_CHECK(s, "null encoder stream");
    SV_CUDA(cudaSetDevice(s->owner->eng.device));
    s->owner->eng.enc_stream_reset(s->st, (cudaStream_t)stream);
  });
}

int64_t svanon_enc_stream_position(const svanon_enc_stream* s) { return s ? s->st.pos : -1; }

int svanon_enc_push_chunk(svanon_enc_stream* s, const float* wave, int n_samples_per_stream, int64_t* ids_out, void* stream) {
  return guarded([&] {
    SV_CHECK(s && wave && ids_out, "null argument");
    SV_CHECK(n_samples_per_stream >= SAMPLES_PER_FRAME && n_samples_per_stream % SAMPLES_PER_FRAME == 0 &&
                 n_samples_per_stream <= 8 * SAMPLES_PER_FRAME, "a push holds 1..8 whole content frames (2048 samples each) per stream");
    const int c = n_samples_per_stream / SAMPLES_PER_FRAME, B = s->st.B;
    Args a(s->owner, stream, ((size_t)n_samples_per_stream * 4 + (size_t)c * 8) * B + 65536);
    const float* w = a.in(wave, (size_t)n_samples_per_stream * B);
    long long* ids = (long long*)a.out(ids_out, (size_t)c * B);
    s->owner->eng.enc_push(s->st, w, n_samples_per_stream, c, ids, c, a.st);
    a.finish();
  });
}

int svanon_voc_stream_create(svanon_engine* e, int n_streams, int frames_per_push, svanon_voc_stream** out) {
  return guarded([&] {
    SV_CHECK(e && out && n_streams >= 1, "bad arguments");
    SV_CHECK(frames_per_push >= 1 && frames_per_push <= 8, "1..8 code frames per push");
    SV_CUDA(cudaSetDevice(e->eng.device));
    SV_CHECK(e->eng.finalized[MODEL_VOCODER], "vocoder weights not finalized");
    auto* h = new svanon_voc_stream();
    h->owner = e;
    try {
      e->eng.voc_state_init(h->st, frames_per_push, n_streams);
      SV_CUDA(cudaMalloc(&h->codes_dev, (size_t)n_streams * 8 * frames_per_push * sizeof(long long)));
    } catch (...) {
      delete h;
      throw;
    }
    *out = h;
  });
}

void svanon_voc_stream_destroy(svanon_voc_stream* s) { delete s; }

int svanon_voc_stream_reset(svanon_voc_stream* s, void* stream) {
  return guarded([&] {
    SV_CHECK(s, "null vocoder stream");
    SV_CUDA(cudaSetDevice(s->owner->eng.device));
    s->owner->eng.voc_state_reset(s->st, (cudaStream_t)stream);
  });
}

int svanon_voc_push_frames(svanon_voc_stream* s, const int64_t* codes, float* wave_out, void* stream) {
  return guarded([&] {
    SV_CHECK(s && codes && wave_out, "null argument");
    const int c = s->st.c, B = s->st.B;
    Args a(s->owner, stream, ((size_t)8 * c * 8 + (size_t)c * SAMPLES_PER_FRAME * 4) * B + 65536);
    const long long* cd = (const long long*)a.in(codes, (size_t)B * 8 * c);
    float* w = a.out(wave_out, (size_t)B * c * SAMPLES_PER_FRAME);
    s->owner->eng.voc_step(s->st, cd, c, w, a.st, (long long)8 * c);
    a.finish();
  });
}

}  // extern "C"
