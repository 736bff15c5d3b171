// Stage A for MANY concurrent streams: one frame of the dual-AR decode (DualARWrapper.decode_one,
// dual_ar_stream.py:817-837 -> decode_one_token_ar :1168-1219) for B independent streams in lock-step.
//
// The persistent kernel of ar_decode.cu is a GEMV machine: right for 1-4 streams, where the frame is bound by
// weight bandwidth and by the latency of its 200 dependent phases.  With tens of streams the projections are real
// GEMMs (M = 2B rows in the slow stack, B rows in the fast stack) and belong on the tensor cores, so this path
// runs the frame as a sequence of kernels over all streams: the weights are read once per frame for everybody
// (tcgen05 3xTF32 GEMMs of gemm_tc.cu, fp32-grade products -> the same token ids), while everything that is per
// stream -- RoPE position, KV-cache append, attention over that stream's own cache, sampler noise -- is driven by a
// per-slot descriptor table in device memory.  Sampling reuses the exact sampler of the persistent kernel.
#include "ar_decode_common.cuh"
#include "engine.hpp"

namespace svanon {

void launch_arb_attn_slow_tma(const ArBatchSlot* slots, const float* qkv, float* y, const float* rope, int layer, int max_seq,
                              int B, bool kv_half, cudaStream_t st);      // ar_attn_tma.cu

using namespace ardec;

namespace {

// x[2b] = cached_new_audio_emb, x[2b+1] = embedding[content_id] (or an explicit row)
__global__ void __launch_bounds__(256) arb_gather_kernel(const ArBatchSlot* __restrict__ slots, const float* __restrict__ cond_emb,
                                                         float* __restrict__ x) {
  pdl_trigger();
  pdl_wait();
  const int m = blockIdx.x, b = m >> 1, j = m & 1;
  const ArBatchSlot& s = slots[b];
  const float* src = j == 0 ? s.x_audio : (s.cond_row ? s.cond_row : cond_emb + (*s.content_id) * D);
  for (int c = threadIdx.x; c < D; c += blockDim.x) x[(long long)m * D + c] = src[c];
}

// Slow-stack attention of one (stream, head), fused with what precedes it in Attention.forward
// (dual_ar_stream.py:895-936): RoPE on q and k of the two new tokens, KV-cache append, then attention of both tokens
// over that stream's valid cache prefix.  grid (H, B); every lane owns the interleaved pair (2*lane, 2*lane+1) of
// the head, 8 warps walk the keys; token 0 sits at pos (keys <= pos), token 1 at pos+1.
constexpr int ATT_WARPS = 8;
__global__ void __launch_bounds__(ATT_WARPS * 32) arb_attn_slow_kernel(const ArBatchSlot* __restrict__ slots,
                                                                       const float* __restrict__ qkv, float* __restrict__ y,
                                                                       const float* __restrict__ rope, int layer, int max_seq) {
  pdl_trigger();
  pdl_wait();
  __shared__ float sm[ATT_WARPS * 2 * PART];
  const int h = blockIdx.x, b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ArBatchSlot& s = slots[b];
  const int pos = s.pos;
  const int nkeys = pos + 2;
  float* kc = s.kc + ((long long)layer * H + h) * max_seq * HEAD_DIM;
  float* vc = s.vc + ((long long)layer * H + h) * max_seq * HEAD_DIM;
  float2 q[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float* row = qkv + (long long)(2 * b + j) * 3 * D + h * HEAD_DIM;
    const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((long long)(pos + j) * (HEAD_DIM / 2) + lane) * 2));
    const float2 qv = *(reinterpret_cast<const float2*>(row) + lane);
    q[j] = make_float2(qv.x * cs.x - qv.y * cs.y, qv.y * cs.x + qv.x * cs.y);
    if (warp == 0) {
      const float2 kv = *(reinterpret_cast<const float2*>(row + D) + lane);
      const float2 vv = *(reinterpret_cast<const float2*>(row + 2 * D) + lane);
      *(reinterpret_cast<float2*>(kc + (long long)(pos + j) * HEAD_DIM) + lane) =
          make_float2(kv.x * cs.x - kv.y * cs.y, kv.y * cs.x + kv.x * cs.y);
      *(reinterpret_cast<float2*>(vc + (long long)(pos + j) * HEAD_DIM) + lane) = vv;
    }
  }
  if (warp == 0) __threadfence();
  __syncthreads();                       // the two new keys are in the cache before anybody walks it
  const float2 q0 = q[0], q1 = q[1];
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  for (int key = warp; key < nkeys; key += ATT_WARPS) {
    const float2 kv = __ldcg(reinterpret_cast<const float2*>(kc + (long long)key * HEAD_DIM) + lane);
    const float2 vv = __ldcg(reinterpret_cast<const float2*>(vc + (long long)key * HEAD_DIM) + lane);
    const float s0 = warp_sum(q0.x * kv.x + q0.y * kv.y) * 0.125f;
    const float s1 = warp_sum(q1.x * kv.x + q1.y * kv.y) * 0.125f;
    if (key <= pos) {
      const float mn = fmaxf(m0, s0);
      const float c = expf(m0 - mn), p = expf(s0 - mn);
      l0 = l0 * c + p; a0.x = a0.x * c + p * vv.x; a0.y = a0.y * c + p * vv.y; m0 = mn;
    }
    {
      const float mn = fmaxf(m1, s1);
      const float c = expf(m1 - mn), p = expf(s1 - mn);
      l1 = l1 * c + p; a1.x = a1.x * c + p * vv.x; a1.y = a1.y * c + p * vv.y; m1 = mn;
    }
  }
  float* mine0 = sm + (warp * 2 + 0) * PART;
  float* mine1 = sm + (warp * 2 + 1) * PART;
  if (lane == 0) { mine0[0] = m0; mine0[1] = l0; mine1[0] = m1; mine1[1] = l1; }
  mine0[2 + 2 * lane] = a0.x; mine0[3 + 2 * lane] = a0.y;
  mine1[2 + 2 * lane] = a1.x; mine1[3 + 2 * lane] = a1.y;
  __syncthreads();
  if (warp < 2) {
    const int tkn = warp;
    float mm = -INFINITY;
    for (int ww = 0; ww < ATT_WARPS; ++ww) mm = fmaxf(mm, sm[(ww * 2 + tkn) * PART]);
    float ll = 0.f, ax = 0.f, ay = 0.f;
    for (int ww = 0; ww < ATT_WARPS; ++ww) {
      const float* pp = sm + (ww * 2 + tkn) * PART;
      const float c = (pp[0] == -INFINITY) ? 0.f : expf(pp[0] - mm);
      ll += pp[1] * c; ax += pp[2 + 2 * lane] * c; ay += pp[3 + 2 * lane] * c;
    }
    const float inv = 1.f / ll;
    *(reinterpret_cast<float2*>(y + (long long)(2 * b + tkn) * D + h * HEAD_DIM) + lane) = make_float2(ax * inv, ay * inv);
  }
}

// Fast-stack attention, fused the same way: RoPE at position cb, append into the 8-slot cache, attention over the
// <= 8 keys.  One warp per (stream, head).
__global__ void __launch_bounds__(128) arb_attn_fast_kernel(const ArBatchSlot* __restrict__ slots, const float* __restrict__ qkv,
                                                            float* __restrict__ y, const float* __restrict__ rope, int layer,
                                                            int cb, int n_items) {
  pdl_trigger();
  pdl_wait();
  const int item = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (item >= n_items) return;
  const int b = item / H, h = item % H;
  const ArBatchSlot& s = slots[b];
  float* kc = s.fkc + ((long long)layer * H + h) * AR_CODEBOOKS * HEAD_DIM;
  float* vc = s.fvc + ((long long)layer * H + h) * AR_CODEBOOKS * HEAD_DIM;
  const float* row = qkv + (long long)b * 3 * D + h * HEAD_DIM;
  const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((long long)cb * (HEAD_DIM / 2) + lane) * 2));
  const float2 qr = *(reinterpret_cast<const float2*>(row) + lane);
  const float2 kr = *(reinterpret_cast<const float2*>(row + D) + lane);
  const float2 v_new = *(reinterpret_cast<const float2*>(row + 2 * D) + lane);
  const float2 qv = make_float2(qr.x * cs.x - qr.y * cs.y, qr.y * cs.x + qr.x * cs.y);
  const float2 k_new = make_float2(kr.x * cs.x - kr.y * cs.y, kr.y * cs.x + kr.x * cs.y);
  *(reinterpret_cast<float2*>(kc + cb * HEAD_DIM) + lane) = k_new;
  *(reinterpret_cast<float2*>(vc + cb * HEAD_DIM) + lane) = v_new;
  float sc[AR_CODEBOOKS];
  float mx = -INFINITY;
#pragma unroll
  for (int key = 0; key < AR_CODEBOOKS; ++key) {
    float v = -INFINITY;
    if (key <= cb) {
      const float2 kv = key == cb ? k_new : __ldcg(reinterpret_cast<const float2*>(kc + key * HEAD_DIM) + lane);
      v = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
    }
    sc[key] = v;
    mx = fmaxf(mx, v);
  }
  float l = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
  for (int key = 0; key < AR_CODEBOOKS; ++key) {
    if (key <= cb) {
      const float p = expf(sc[key] - mx);
      const float2 vv = key == cb ? v_new : __ldcg(reinterpret_cast<const float2*>(vc + key * HEAD_DIM) + lane);
      l += p; ax += p * vv.x; ay += p * vv.y;
    }
  }
  const float inv = 1.f / l;
  *(reinterpret_cast<float2*>(y + (long long)b * D + h * HEAD_DIM) + lane) = make_float2(ax * inv, ay * inv);
}

// top-p / temperature sampler of codebook `cb` for every stream (one CTA each), then the next fast input row
__global__ void __launch_bounds__(NT) arb_sample_kernel(const ArBatchSlot* __restrict__ slots, const float* __restrict__ logits,
                                                        const float* __restrict__ fast_emb, float* __restrict__ xf, int cb) {
  pdl_trigger();
  pdl_wait();
  __shared__ SampleSmem ssm;
  const int b = blockIdx.x;
  const ArBatchSlot& s = slots[b];
  const float* noise = s.noise ? s.noise + cb * AR_CB_SIZE : nullptr;
  const int tok = sample_topp(logits + (long long)b * 1024, noise, s.seed, s.step, cb + 1, s.temperature, s.top_p, ssm);
  if (threadIdx.x == 0) s.out_codes[cb] = tok;
  for (int i = threadIdx.x; i < D; i += NT) xf[(long long)b * D + i] = __ldg(fast_emb + (long long)tok * D + i);
}

// cached_new_audio_emb = embed(pred codes) (dual_ar_stream.py:245-255, 834) and the pred_codes history column
__global__ void __launch_bounds__(256) arb_finish_kernel(const ArBatchSlot* __restrict__ slots, const float* __restrict__ codebook_emb) {
  pdl_trigger();
  pdl_wait();
  const ArBatchSlot& s = slots[blockIdx.x];
  int code[AR_CODEBOOKS];
#pragma unroll
  for (int k = 0; k < AR_CODEBOOKS; ++k) code[k] = s.out_codes[k];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < AR_CODEBOOKS; ++k) v += __ldg(codebook_emb + ((long long)code[k] + k * AR_CB_SIZE) * D + c);
    s.x_audio[c] = v;
  }
  if (s.pred_hist && threadIdx.x < AR_CODEBOOKS) s.pred_hist[(long long)threadIdx.x * s.pred_ld + s.pred_col] = code[threadIdx.x];
}

// 1 (default): the TMA-staged warp-specialised kernel of ar_attn_tma.cu; 0: the warp-per-key kernel above (kept for
// A/B measurements: SVANON_ATTN_TMA=0)
bool use_attn_tma() {
  static const bool on = [] { const char* e = getenv("SVANON_ATTN_TMA"); return !e || atoi(e) != 0; }();
  return on;
}

void gemm(const float* A, long long lda, const float* W, float* C, long long ldc, const float* residual, int M, int N, int K,
          cudaStream_t st) {
  GemmParams p;
  p.A = A; p.W = W; p.C = C; p.residual = residual; p.M = M; p.N = N; p.K = K; p.lda = lda; p.ldc = ldc; p.ldr = ldc;
  launch_gemm(p, st);
}

}  // namespace

ArBatchWork::~ArBatchWork() {
  for (void* p : {(void*)x, (void*)nrm, (void*)qkv, (void*)y, (void*)h13, (void*)g, (void*)xf, (void*)logits, (void*)slots_dev})
    if (p) cudaFree(p);
  for (int i = 0; i < RING; ++i) {
    if (slots_host[i]) cudaFreeHost(slots_host[i]);
    if (ev[i]) cudaEventDestroy(ev[i]);
  }
}

void ArBatchWork::ensure(int B) {
  if (B <= cap) return;
  SV_CUDA(cudaDeviceSynchronize());
  for (void* p : {(void*)x, (void*)nrm, (void*)qkv, (void*)y, (void*)h13, (void*)g, (void*)xf, (void*)logits, (void*)slots_dev})
    if (p) cudaFree(p);
  for (int i = 0; i < RING; ++i)
    if (slots_host[i]) { cudaFreeHost(slots_host[i]); slots_host[i] = nullptr; }
  const size_t M = (size_t)2 * B;
  auto fa = [&](float*& p, size_t n) { p = nullptr; SV_CUDA(cudaMalloc(&p, n * sizeof(float))); };
  fa(x, M * AR_DIM); fa(nrm, M * AR_DIM); fa(qkv, M * 3 * AR_DIM); fa(y, M * AR_DIM); fa(h13, M * 2 * AR_INTER);
  fa(g, M * AR_INTER); fa(xf, (size_t)B * AR_DIM); fa(logits, (size_t)B * 1024);
  slots_dev = nullptr;
  SV_CUDA(cudaMalloc(&slots_dev, (size_t)B * sizeof(ArBatchSlot)));
  for (int i = 0; i < RING; ++i) {
    SV_CUDA(cudaMallocHost(&slots_host[i], (size_t)B * sizeof(ArBatchSlot)));
    if (!ev[i]) SV_CUDA(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    ev_pending[i] = false;
  }
  cap = B;
}

// One decode step for `batch` streams (any count) on the GEMM path.
void Engine::ar_decode_step_gemm(Stream* const* streams, int batch, cudaStream_t st) {
  NvtxRange nvtx_("svanon:A decode (many streams)");
  SV_CHECK(finalized[MODEL_AR], "AR weights not finalized");
  SV_CHECK(batch >= 1, "empty decode batch");
  arb.ensure(batch);
  const int B = batch;
  const int max_seq = streams[0]->max_seq;
  // ---- per-slot descriptors: pinned ring buffer -> device table (stream-ordered, no host sync in steady state)
  const int slot = arb.cur;
  arb.cur = (arb.cur + 1) % ArBatchWork::RING;
  if (arb.ev_pending[slot]) SV_CUDA(cudaEventSynchronize(arb.ev[slot]));
  ArBatchSlot* hs = arb.slots_host[slot];
  for (int b = 0; b < B; ++b) {
    Stream& s = *streams[b];
    SV_CHECK(s.pos_next + 2 <= s.max_seq, "KV cache full: re-prompt before decoding further");
    SV_CHECK(s.max_seq == max_seq, "batched streams must share max_seq_len");
    SV_CHECK(s.step_content_id || s.step_cond_row, "decode step without a content id");
    ArBatchSlot& d = hs[b];
    d.kc = s.kc; d.vc = s.vc; d.fkc = s.fkc; d.fvc = s.fvc; d.x_audio = s.x_audio;
    d.content_id = s.step_content_id; d.cond_row = s.step_cond_row; d.noise = s.step_noise; d.out_codes = s.codes_dev;
    d.pos = s.pos_next; d.step = s.step; d.seed = s.seed;
    d.temperature = s.temperature; d.top_p = s.top_p;
    d.pred_hist = s.step_pred_hist; d.pred_ld = HIST_CAP; d.pred_col = s.step_pred_col;
  }
  SV_CUDA(cudaMemcpyAsync(arb.slots_dev, hs, (size_t)B * sizeof(ArBatchSlot), cudaMemcpyHostToDevice, st));
  SV_CUDA(cudaEventRecord(arb.ev[slot], st));
  arb.ev_pending[slot] = true;
  const ArBatchSlot* sd = arb.slots_dev;

  auto layer = [&](const ArLayerWeights& w, float* xr, int M, int li, bool fast, int cb) {
    launch_rmsnorm(xr, arb.nrm, w.attn_norm, M, AR_DIM, AR_NORM_EPS, st);
    gemm(arb.nrm, AR_DIM, w.wqkv, arb.qkv, 3 * AR_DIM, nullptr, M, 3 * AR_DIM, AR_DIM, st);
    if (fast) {
      launch_pdl(arb_attn_fast_kernel, dim3((B * AR_HEADS + 3) / 4), dim3(128), 0, st, sd, (const float*)arb.qkv, arb.y,
                 ar.fast_rope, li, cb, B * AR_HEADS);
    } else if (use_attn_tma()) {
      launch_arb_attn_slow_tma(sd, arb.qkv, arb.y, ar.rope, li, max_seq, B, false, st);
    } else {
      launch_pdl(arb_attn_slow_kernel, dim3(AR_HEADS, B), dim3(ATT_WARPS * 32), 0, st, sd, (const float*)arb.qkv, arb.y,
                 ar.rope, li, max_seq);
    }
    SV_LAUNCHED();
    gemm(arb.y, AR_DIM, w.wo, xr, AR_DIM, xr, M, AR_DIM, AR_DIM, st);
    launch_rmsnorm(xr, arb.nrm, w.ffn_norm, M, AR_DIM, AR_NORM_EPS, st);
    {
      GemmParams p13[2];                     // w1 and w3 side by side in one launch
      for (int i = 0; i < 2; ++i) {
        GemmParams& p = p13[i];
        p.A = arb.nrm; p.W = i ? w.w3 : w.w1; p.C = arb.h13 + i * AR_INTER; p.M = M; p.N = AR_INTER; p.K = AR_DIM;
        p.lda = AR_DIM; p.ldc = 2 * AR_INTER;
      }
      launch_gemm(p13, 2, st);
    }
    launch_silu_mul(arb.h13, arb.g, M, AR_INTER, st);
    gemm(arb.g, AR_INTER, w.w2, xr, AR_DIM, xr, M, AR_DIM, AR_INTER, st);
  };

  launch_pdl(arb_gather_kernel, dim3(2 * B), dim3(256), 0, st, sd, ar.cond_emb, arb.x);
  SV_LAUNCHED();
  for (int l = 0; l < AR_LAYERS; ++l) layer(ar.slow[l], arb.x, 2 * B, l, false, 0);
  // hidden state handed to the fast transformer = PRE-norm residual of the last token (dual_ar_stream.py:354-355)
  launch_copy_rows(arb.x + AR_DIM, 2 * AR_DIM, arb.xf, AR_DIM, B, AR_DIM, st);
  for (int cb = 0; cb < AR_CODEBOOKS; ++cb) {
    for (int l = 0; l < AR_FAST_LAYERS; ++l) layer(ar.fast[l], arb.xf, B, l, true, cb);
    launch_rmsnorm(arb.xf, arb.nrm, ar.fast_norm_w, B, AR_DIM, AR_NORM_EPS, st);
    gemm(arb.nrm, AR_DIM, ar.fast_output_w, arb.logits, 1024, nullptr, B, AR_CB_SIZE, AR_DIM, st);
    if (debug_logits)        // test / evaluation hook: the fast-head logits of stream 0 (svanon_ar_read_debug)
      SV_CUDA(cudaMemcpyAsync(dbg_fast_logits + (size_t)cb * AR_CB_SIZE, arb.logits, AR_CB_SIZE * sizeof(float),
                              cudaMemcpyDeviceToDevice, st));
    launch_pdl(arb_sample_kernel, dim3(B), dim3(NT), 0, st, sd, (const float*)arb.logits, ar.fast_emb, arb.xf, cb);
    SV_LAUNCHED();
  }
  launch_pdl(arb_finish_kernel, dim3(B), dim3(256), 0, st, sd, ar.codebook_emb);
  SV_LAUNCHED();
  for (int b = 0; b < B; ++b) {
    streams[b]->pos_next += 2;
    streams[b]->step += 1;
  }
}

}  // namespace svanon
