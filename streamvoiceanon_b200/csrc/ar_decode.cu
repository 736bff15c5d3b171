// Stage A: one frame of the dual-AR decode as ONE persistent cooperative kernel.
//
// Reference path: DualARWrapper.decode_one (dual_ar_stream.py:817-837) -> decode_one_token_ar (:1168-1219):
// 12 "slow" layers over the 2 new tokens [cached_new_audio_emb, embedding[content_id]] with KV-cache append,
// then 8 sequential codebook steps of the 4-layer "fast" transformer, each followed by top-p/temperature
// sampling (:1099-1132) with Gumbel-style argmax(p/q) (:1092-1096), then embed(pred codes) (:245-255,834).
// That is 44 dependent layer evaluations + 8 samplers per 46 ms frame: at batch 1 the job is latency- and
// weight-bandwidth-bound (every layer is a 30.7 MB fp32 GEMV).  The whole frame therefore runs inside one
// launch: one CTA per SM, every projection a warp-per-row GEMV streamed straight from HBM/L2 with 128-bit
// no-allocate loads, phases separated by a hand-rolled grid barrier (5 per slow layer, 4 per fast layer).
// The discarded 8192-way token head (:1183-1192) is skipped unless debug logits are requested; the noise
// tape is slot-indexed so skipping it does not shift the stream (SURVEY.md finding 5).
//
// B independent streams can share one launch (weights are read once for all of them): the slow phases then
// see M = 2B token rows and the fast phases M = B.
#include "ar_decode_common.cuh"

namespace svanon {

using namespace ardec;

namespace {


// ---------------------------------------------------------------- one transformer layer, M rows
// x: residual stream rows [M][D] (global scratch).  ROWS_PER_STREAM = 2 (slow) or 1 (fast).
template <int B, int RPS, bool FAST>
__device__ __forceinline__ void layer(const ArDecodeArgs& a, const ArLayerWeights& w, int layer_idx, int cb,
                                      float* smem, unsigned nblocks) {
  constexpr int M = B * RPS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = warp * gridDim.x + blockIdx.x;
  const int total_warps = NW * gridDim.x;
  float* xs = smem;

  // ---- phase 1: attention_norm + wqkv (+RoPE) ; q -> scratch, k/v -> cache
  load_rmsnorm<M>(a.x, w.attn_norm, xs, nullptr);
  for (int pair = gwarp; pair < 3 * D / 2; pair += total_warps) {
    const float* rows[2] = {w.wqkv + (long long)(2 * pair) * D, w.wqkv + (long long)(2 * pair + 1) * D};
    float o[2][M];
    warp_rows_dot<2, M>(rows, xs, D, o);
    if (lane == 0) {
      const int r = 2 * pair;
      const int sec = r / D;                 // 0 q, 1 k, 2 v
      const int c = r % D;
      const int h = c / HEAD_DIM, d = c % HEAD_DIM;
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const int b = m / RPS, j = m % RPS;
        const int pos = FAST ? cb : a.s[b].pos + j;
        float v0 = o[0][m], v1 = o[1][m];
        if (sec < 2) {
          const float* tab = (FAST ? a.fast_rope : a.rope) + ((long long)pos * (HEAD_DIM / 2) + d / 2) * 2;
          const float cs = __ldg(tab), sn = __ldg(tab + 1);
          const float r0 = v0 * cs - v1 * sn, r1 = v1 * cs + v0 * sn;
          v0 = r0; v1 = r1;
        }
        if (sec == 0) {
          a.q[m * D + c] = v0; a.q[m * D + c + 1] = v1;
        } else {
          float* base;
          if (FAST) base = (sec == 1 ? a.s[b].fkc : a.s[b].fvc) + (((long long)layer_idx * H + h) * AR_CODEBOOKS + pos) * HEAD_DIM;
          else base = (sec == 1 ? a.s[b].kc : a.s[b].vc) + (((long long)layer_idx * H + h) * a.max_seq + pos) * HEAD_DIM;
          base[d] = v0; base[d + 1] = v1;
        }
      }
    }
  }
  grid_sync(a.barrier, nblocks);

  float* ys = smem;                          // [M][D] attention output
  if (!FAST) {
    // ---- phase 2: split-KV attention partials.  work item = (stream, head, split)
    const int nitems = B * H * a.nsplit;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int b = item / (H * a.nsplit), h = (item / a.nsplit) % H, sp = item % a.nsplit;
      const int pos = a.s[b].pos;
      const int nkeys = pos + 2;
      const int chunk = (nkeys + a.nsplit - 1) / a.nsplit;
      const int k_begin = sp * chunk, k_end = min(nkeys, k_begin + chunk);
      const float* kc = a.s[b].kc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      const float* vc = a.s[b].vc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      // each lane owns dims (2*lane, 2*lane+1); a warp walks keys k_begin+warp, +NW, ...
      float2 q0 = __ldcg(reinterpret_cast<const float2*>(a.q + (b * 2 + 0) * D + h * HEAD_DIM) + lane);
      float2 q1 = __ldcg(reinterpret_cast<const float2*>(a.q + (b * 2 + 1) * D + h * HEAD_DIM) + lane);
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      for (int key = k_begin + warp; key < k_end; key += NW) {
        const float2 kv = __ldcg(reinterpret_cast<const float2*>(kc + (long long)key * HEAD_DIM) + lane);
        const float2 vv = __ldcg(reinterpret_cast<const float2*>(vc + (long long)key * HEAD_DIM) + lane);
        float s0 = warp_sum(q0.x * kv.x + q0.y * kv.y) * 0.125f;
        float s1 = warp_sum(q1.x * kv.x + q1.y * kv.y) * 0.125f;
        if (key <= pos) {                    // token 0 sits at `pos`, token 1 at pos+1
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
      // combine the NW warps of this CTA through shared memory
      float* sm = smem;                      // [NW][2][PART]
      __syncthreads();
      float* mine0 = sm + (warp * 2 + 0) * PART;
      float* mine1 = sm + (warp * 2 + 1) * PART;
      if (lane == 0) { mine0[0] = m0; mine0[1] = l0; mine1[0] = m1; mine1[1] = l1; }
      mine0[2 + 2 * lane] = a0.x; mine0[3 + 2 * lane] = a0.y;
      mine1[2 + 2 * lane] = a1.x; mine1[3 + 2 * lane] = a1.y;
      __syncthreads();
      if (warp < 2) {
        const int tkn = warp;
        float mm = -INFINITY;
        for (int ww = 0; ww < NW; ++ww) mm = fmaxf(mm, sm[(ww * 2 + tkn) * PART]);
        float ll = 0.f, ax = 0.f, ay = 0.f;
        for (int ww = 0; ww < NW; ++ww) {
          const float* pp = sm + (ww * 2 + tkn) * PART;
          const float c = (pp[0] == -INFINITY) ? 0.f : expf(pp[0] - mm);
          ll += pp[1] * c; ax += pp[2 + 2 * lane] * c; ay += pp[3 + 2 * lane] * c;
        }
        float* dst = a.part + ((((long long)b * H + h) * a.nsplit + sp) * 2 + tkn) * PART;
        if (lane == 0) { dst[0] = mm; dst[1] = ll; }
        dst[2 + 2 * lane] = ax; dst[3 + 2 * lane] = ay;
      }
      __syncthreads();
    }
    grid_sync(a.barrier, nblocks);
    // ---- phase 3a: every CTA merges the split partials of all (stream, head, token) into ys
    for (int it = warp; it < B * H * 2; it += NW) {
      const int b = it / (H * 2), h = (it / 2) % H, tkn = it % 2;
      const float* base = a.part + (((long long)b * H + h) * a.nsplit * 2 + tkn) * PART;
      float mm = -INFINITY;
      for (int sp = 0; sp < a.nsplit; ++sp) mm = fmaxf(mm, __ldcg(base + (long long)sp * 2 * PART));
      float ll = 0.f, ax = 0.f, ay = 0.f;
      for (int sp = 0; sp < a.nsplit; ++sp) {
        const float* pp = base + (long long)sp * 2 * PART;
        const float pm = __ldcg(pp);
        const float c = (pm == -INFINITY) ? 0.f : expf(pm - mm);
        ll += __ldcg(pp + 1) * c;
        const float2 av = __ldcg(reinterpret_cast<const float2*>(pp + 2) + lane);
        ax += av.x * c; ay += av.y * c;
      }
      const float inv = 1.f / ll;
      ys[(b * 2 + tkn) * D + h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[(b * 2 + tkn) * D + h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  } else {
    // ---- fast path: <= 8 keys, every CTA recomputes the attention of all heads (no extra barrier)
    for (int it = warp; it < B * H; it += NW) {
      const int b = it / H, h = it % H;
      const float2 qv = __ldcg(reinterpret_cast<const float2*>(a.q + b * D + h * HEAD_DIM) + lane);
      const float* kc = a.s[b].fkc + ((long long)layer_idx * H + h) * AR_CODEBOOKS * HEAD_DIM;
      const float* vc = a.s[b].fvc + ((long long)layer_idx * H + h) * AR_CODEBOOKS * HEAD_DIM;
      float sc[AR_CODEBOOKS];
      float mx = -INFINITY;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        float s = -INFINITY;
        if (key <= cb) {
          const float2 kv = __ldcg(reinterpret_cast<const float2*>(kc + key * HEAD_DIM) + lane);
          s = warp_sum(qv.x * kv.x + qv.y * kv.y) * 0.125f;
        }
        sc[key] = s;
        mx = fmaxf(mx, s);
      }
      float l = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        if (key <= cb) {
          const float p = expf(sc[key] - mx);
          const float2 vv = __ldcg(reinterpret_cast<const float2*>(vc + key * HEAD_DIM) + lane);
          l += p; ax += p * vv.x; ay += p * vv.y;
        }
      }
      const float inv = 1.f / l;
      ys[b * D + h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[b * D + h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  }
  // ---- phase 3b: wo + residual -> h
  for (int row = gwarp; row < D; row += total_warps) {
    const float* rows[1] = {w.wo + (long long)row * D};
    float o[1][M];
    warp_rows_dot<1, M>(rows, ys, D, o);
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) a.h[m * D + row] = __ldcg(a.x + m * D + row) + o[0][m];
    }
  }
  grid_sync(a.barrier, nblocks);

  // ---- phase 4: ffn_norm + silu(w1 h) * (w3 h) -> g
  load_rmsnorm<M>(a.h, w.ffn_norm, xs, nullptr);
  for (int row = gwarp; row < I; row += total_warps) {
    const float* rows[2] = {w.w1 + (long long)row * D, w.w3 + (long long)row * D};
    float o[2][M];
    warp_rows_dot<2, M>(rows, xs, D, o);
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float u = o[0][m];
        a.g[m * I + row] = (u / (1.f + expf(-u))) * o[1][m];
      }
    }
  }
  grid_sync(a.barrier, nblocks);

  // ---- phase 5: w2 + residual -> x
  for (int i = threadIdx.x; i < M * I; i += NT) xs[i] = __ldcg(a.g + i);
  __syncthreads();
  for (int row = gwarp; row < D; row += total_warps) {
    const float* rows[1] = {w.w2 + (long long)row * I};
    float o[1][M];
    warp_rows_dot<1, M>(rows, xs, I, o);
    if (lane == 0) {
#pragma unroll
      for (int m = 0; m < M; ++m) a.x[m * D + row] = __ldcg(a.h + m * D + row) + o[0][m];
    }
  }
  grid_sync(a.barrier, nblocks);
}

template <int B>
__global__ void __launch_bounds__(NT, 1) ar_decode_kernel(const ArDecodeArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ SampleSmem ssm;
  const unsigned nblocks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = warp * gridDim.x + blockIdx.x;
  const int total_warps = NW * gridDim.x;
  const int gtid = blockIdx.x * NT + threadIdx.x;

  grid_sync_init(a.barrier);
  // ---- phase 0: assemble the 2 input rows per stream: [cached_new_audio_emb, embedding[content_id]]
  for (int i = gtid; i < B * 2 * D; i += NT * gridDim.x) {
    const int b = i / (2 * D), j = (i / D) % 2, c = i % D;
    float v;
    if (j == 0) v = __ldcg(a.s[b].x_audio + c);
    else if (a.s[b].cond_row) v = __ldcg(a.s[b].cond_row + c);
    else v = __ldg(a.cond_emb + (*a.s[b].content_id) * D + c);
    a.x[i] = v;
  }
  grid_sync(a.barrier, nblocks);

  for (int l = 0; l < AR_LAYERS; ++l) layer<B, 2, false>(a, a.slow[l], l, 0, smem, nblocks);

  // hidden state handed to the fast transformer = PRE-norm residual of the last token (dual_ar_stream.py:354-355)
  if (a.dbg_slow_logits) {
    // optional (tests): the discarded 8192-way token head of stream 0
    float* xs = smem;
    load_rmsnorm<1>(a.x + D, a.norm_w, xs, nullptr);
    for (int row = gwarp; row < AR_VOCAB; row += total_warps) {
      const float* rows[1] = {a.output_w + (long long)row * D};
      float o[1][1];
      warp_rows_dot<1, 1>(rows, xs, D, o);
      if (lane == 0) a.dbg_slow_logits[row] = o[0][0];
    }
    if (a.dbg_hidden) for (int i = gtid; i < D; i += NT * gridDim.x) a.dbg_hidden[i] = __ldcg(a.x + D + i);
    __syncthreads();
  }
  // compact: xf[b] = x[b*2+1]   (written into h, then swapped in by using h as the fast residual stream)
  for (int i = gtid; i < B * D; i += NT * gridDim.x) a.h[i] = __ldcg(a.x + ((i / D) * 2 + 1) * D + (i % D));
  grid_sync(a.barrier, nblocks);
  for (int i = gtid; i < B * D; i += NT * gridDim.x) a.x[i] = __ldcg(a.h + i);
  grid_sync(a.barrier, nblocks);

  for (int cb = 0; cb < AR_CODEBOOKS; ++cb) {
    for (int l = 0; l < AR_FAST_LAYERS; ++l) layer<B, 1, true>(a, a.fast[l], l, cb, smem, nblocks);
    // fast_norm + fast_output -> logits
    float* xs = smem;
    load_rmsnorm<B>(a.x, a.fast_norm_w, xs, nullptr);
    for (int row = gwarp; row < AR_CB_SIZE; row += total_warps) {
      const float* rows[1] = {a.fast_output_w + (long long)row * D};
      float o[1][B];
      warp_rows_dot<1, B>(rows, xs, D, o);
      if (lane == 0) {
#pragma unroll
        for (int b = 0; b < B; ++b) a.logits[b * 1024 + row] = o[0][b];
      }
    }
    grid_sync(a.barrier, nblocks);
    // sampling: CTA b samples stream b, writes the code and the next fast input row
    if (blockIdx.x < B) {
      const int b = blockIdx.x;
      if (a.dbg_fast_logits && b == 0)
        for (int i = threadIdx.x; i < AR_CB_SIZE; i += NT) a.dbg_fast_logits[cb * AR_CB_SIZE + i] = __ldcg(a.logits + i);
      const float* noise = a.s[b].noise ? a.s[b].noise + cb * AR_CB_SIZE : nullptr;
      const int tok = sample_topp(a.logits + b * 1024, noise, a.s[b].seed, a.s[b].step, cb + 1, a.s[b].temperature,
                                  a.s[b].top_p, ssm);
      if (threadIdx.x == 0) a.s[b].out_codes[cb] = tok;
      for (int i = threadIdx.x; i < D; i += NT) a.x[b * D + i] = __ldg(a.fast_emb + (long long)tok * D + i);
    }
    grid_sync(a.barrier, nblocks);
  }

  // ---- cached_new_audio_emb = embed(pred codes)  (dual_ar_stream.py:245-255, 834)
  for (int i = gtid; i < B * D; i += NT * gridDim.x) {
    const int b = i / D, c = i % D;
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < AR_CODEBOOKS; ++k) {
      const int code = __ldcg(a.s[b].out_codes + k);
      s += __ldg(a.codebook_emb + ((long long)code + k * AR_CB_SIZE) * D + c);
    }
    a.s[b].x_audio[c] = s;
  }
  grid_sync_finish(a.barrier);
}

template <int B>
void launch_b(const ArDecodeArgs& args, int grid, cudaStream_t st) {
  const size_t smem = (size_t)2 * B * AR_INTER * sizeof(float) > (size_t)NW * 2 * PART * sizeof(float)
                          ? (size_t)2 * B * AR_INTER * sizeof(float)
                          : (size_t)NW * 2 * PART * sizeof(float);
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(ar_decode_kernel<B>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  void* kargs[] = {(void*)&args};
  SV_CUDA(cudaLaunchCooperativeKernel((void*)ar_decode_kernel<B>, dim3(grid), dim3(NT), kargs, smem, st));
  ++g_kernel_launches;
}

}  // namespace

// Measurement aid: `iters` back-to-back grid barriers (optionally with the store -> fence -> barrier -> L2 load round
// trip a real phase has) in the launch configuration of the decode kernels.
__global__ void __launch_bounds__(NT, 1) grid_barrier_probe_kernel(unsigned* bar, int iters, float* scratch, int exchange) {
  const unsigned nblocks = gridDim.x;
  grid_sync_init(bar);
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    if (exchange) {
      // every CTA publishes 6 floats, then reads everybody's (a 768-float activation vector, as between two phases)
      if (threadIdx.x < 6) scratch[(i & 1) * 1024 + blockIdx.x * 6 + threadIdx.x] = acc + (float)i;
    }
    grid_sync(bar, nblocks);
    if (exchange) {
      for (int j = threadIdx.x; j < (int)nblocks * 6; j += NT) acc += __ldcg(scratch + (i & 1) * 1024 + j);
    }
  }
  if (exchange && acc == 12345.678f) scratch[2047] = acc;
  grid_sync_finish(bar);
}

float grid_barrier_probe(unsigned* bar, int iters, float* scratch, int exchange, int grid, cudaStream_t st) {
  void* kargs[] = {(void*)&bar, (void*)&iters, (void*)&scratch, (void*)&exchange};
  cudaEvent_t e0, e1;
  SV_CUDA(cudaEventCreate(&e0));
  SV_CUDA(cudaEventCreate(&e1));
  SV_CUDA(cudaLaunchCooperativeKernel((void*)grid_barrier_probe_kernel, dim3(grid), dim3(NT), kargs, 0, st));
  SV_CUDA(cudaEventRecord(e0, st));
  SV_CUDA(cudaLaunchCooperativeKernel((void*)grid_barrier_probe_kernel, dim3(grid), dim3(NT), kargs, 0, st));
  SV_CUDA(cudaEventRecord(e1, st));
  SV_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  SV_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return ms;
}

int ar_decode_max_batch() { return 4; }

void launch_ar_decode(const ArDecodeArgs& args, int batch, int grid, cudaStream_t st) {
  switch (batch) {
    case 1: launch_b<1>(args, grid, st); break;
    case 2: launch_b<2>(args, grid, st); break;
    case 4: launch_b<4>(args, grid, st); break;
    default: SV_CHECK(false, "ar decode batch must be 1, 2 or 4");
  }
}

}  // namespace svanon
