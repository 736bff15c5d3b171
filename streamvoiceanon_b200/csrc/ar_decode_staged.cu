// Stage A, batch 1: the persistent decode kernel with TMA-staged weights.
//
// Same phase structure as ar_decode.cu (one cooperative launch per frame, grid barrier between dependent
// GEMV phases), but no phase ever waits for HBM: every CTA owns a fixed, CONTIGUOUS slice of each weight matrix
// (<= 96 KB), and while it computes weight-phase w it already has the slice of phase w+1 in flight as
// `cp.async.bulk` (TMA 1-D bulk copy, SASS UBLKCP) global->shared copies that complete on an mbarrier.
// The DRAM/L2 latency of the next 30 MB of weights therefore overlaps the current phase's math and the grid
// barrier; the math itself reads weights from shared memory.  Slow-layer weights are streamed once per frame
// (L2 evict_first), fast-layer weights are re-used 8x per frame and across frames (L2 evict_last).
//
//   weight-phase schedule (184 per frame): 12 x [wqkv, wo, w1|w3, w2]  then 8 x (4 x [wqkv, wo, w1|w3, w2] + fast_output)
#include "ar_decode_common.cuh"

namespace svanon {

using namespace ardec;

namespace {

constexpr int WBUF_BYTES = 98304;           // 96 KB: the largest per-CTA slice (w1|w3: 16 row pairs x 6 KB)
constexpr int XS_FLOATS = 2 * AR_INTER;     // activations of up to 2 rows x 2304
constexpr int N_SLOW_WP = AR_LAYERS * 4;
constexpr int WP_PER_CB = AR_FAST_LAYERS * 4 + 1;
constexpr int N_WP = N_SLOW_WP + AR_CODEBOOKS * WP_PER_CB;

enum Kind : int { K_QKV = 0, K_WO = 1, K_W13 = 2, K_W2 = 3, K_LOGITS = 4 };

struct Slice {
  const float* src[2];
  int bytes_per_unit;     // per region
  int regions;
  int u0, u1;             // this CTA's unit range
  bool fast;
};

__device__ __forceinline__ int* slice_table() {
  __shared__ int t[10];
  return t;
}
__device__ __forceinline__ void slice_table_init() {
  if (threadIdx.x < 5) {
    const int units[5] = {3 * D / 2, D, I, D, AR_CB_SIZE};        // K_QKV (row pairs), K_WO, K_W13, K_W2, K_LOGITS
    const int U = units[threadIdx.x];
    slice_table()[2 * threadIdx.x] = (int)((long long)U * blockIdx.x / gridDim.x);
    slice_table()[2 * threadIdx.x + 1] = (int)((long long)U * (blockIdx.x + 1) / gridDim.x);
  }
}

__device__ __forceinline__ Slice slice_of(const ArDecodeArgs& a, int wp) {
  Slice s;
  int kind;
  const ArLayerWeights* lw;
  if (wp < N_SLOW_WP) {
    lw = &a.slow[wp >> 2];
    kind = wp & 3;
    s.fast = false;
  } else {
    const int r = (wp - N_SLOW_WP) % WP_PER_CB;
    s.fast = true;
    if (r == WP_PER_CB - 1) { kind = K_LOGITS; lw = &a.fast[0]; }
    else { lw = &a.fast[r >> 2]; kind = r & 3; }
  }
  s.regions = 1;
  s.src[1] = nullptr;
  switch (kind) {
    case K_QKV: s.src[0] = lw->wqkv; s.bytes_per_unit = 2 * D * 4; break;
    case K_WO: s.src[0] = lw->wo; s.bytes_per_unit = D * 4; break;
    case K_W13: s.src[0] = lw->w1; s.src[1] = lw->w3; s.bytes_per_unit = D * 4; s.regions = 2; break;
    case K_W2: s.src[0] = lw->w2; s.bytes_per_unit = I * 4; break;
    default: s.src[0] = a.fast_output_w; s.bytes_per_unit = D * 4; break;
  }
  // this CTA's unit range of each kind never changes: computed once per launch (slice_table_init), no divisions here
  s.u0 = slice_table()[2 * kind];
  s.u1 = slice_table()[2 * kind + 1];
  return s;
}

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni WAIT_DONE;\n"
      "bra.uni WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                             unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

struct Stage {
  unsigned char* wbuf;            // 2 x WBUF_BYTES
  unsigned long long* mbar;       // 2
  unsigned long long pol_stream, pol_keep;
  int wp;                         // next weight phase to be consumed (uniform across the CTA)
};

// Optional in-kernel timeline (svanon_ar_profile): thread 0 of CTA 0 accumulates SM clock cycles per category between
// markers; off (null pointer) in production.
enum ProfCat : int { P_ACT = 0, P_WAIT = 1, P_DOT = 2, P_SYNC = 3, P_ATTN = 4, P_SAMPLE = 5, P_MISC = 6, P_N = 8 };
struct Prof {
  unsigned long long* acc;
  long long last;
  __device__ __forceinline__ void start(unsigned long long* out) {
    acc = (blockIdx.x == 0 && threadIdx.x == 0) ? out : nullptr;
    if (acc) last = clock64();
  }
  __device__ __forceinline__ void tick(int cat) {
    if (acc) {
      const long long t = clock64();
      acc[cat] += (unsigned long long)(t - last);
      last = t;
    }
  }
};

// issue the bulk copies of weight phase `wp` into buffer wp&1 (one thread)
__device__ __forceinline__ void issue(const ArDecodeArgs& a, Stage& sg, int wp) {
  if (wp >= N_WP) return;
  const Slice s = slice_of(a, wp);
  const int n = s.u1 - s.u0;
  if (n <= 0) return;                     // nothing to load: consumers skip the wait as well
  unsigned char* dst = sg.wbuf + (size_t)(wp & 1) * WBUF_BYTES;
  unsigned long long* bar = sg.mbar + (wp & 1);
  const unsigned region_bytes = (unsigned)n * s.bytes_per_unit;
  mbar_expect_tx(bar, region_bytes * s.regions);
  const unsigned long long pol = s.fast ? sg.pol_keep : sg.pol_stream;
  for (int r = 0; r < s.regions; ++r) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(s.src[r]) + (size_t)s.u0 * s.bytes_per_unit;
    // split into <= 32 KB pieces so several copies are in flight
    for (unsigned off = 0; off < region_bytes; off += 32768) {
      const unsigned len = min(32768u, region_bytes - off);
      tma_bulk_g2s(dst + (size_t)r * region_bytes + off, src + off, len, bar, pol);
    }
  }
}

// start of a weight phase: prefetch the next one, wait for this one; returns the slice and its smem base
__device__ __forceinline__ const float* begin_phase(const ArDecodeArgs& a, Stage& sg, Slice& s) {
  const int wp = sg.wp;
  if (threadIdx.x == NT - 32) issue(a, sg, wp + 1);      // lane 0 of the last warp: it has no rows in most phases
  s = slice_of(a, wp);
  if (s.u1 > s.u0) mbar_wait(sg.mbar + (wp & 1), (wp >> 1) & 1);
  sg.wp = wp + 1;
  return reinterpret_cast<const float*>(sg.wbuf + (size_t)(wp & 1) * WBUF_BYTES);
}

// dot products of NR weight rows held in SHARED memory with M activation vectors in shared memory
template <int NR, int M>
__device__ __forceinline__ void warp_rows_dot_s(const float* const* w, const float* xs, int K, float (&out)[NR][M]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) out[r][m] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 768) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      float4 wv[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) wv[r] = *reinterpret_cast<const float4*>(w[r] + k0 + (lane + 32 * i) * 4);
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k0 + (lane + 32 * i) * 4);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          out[r][m] = fmaf(wv[r].x, xv.x, out[r][m]);
          out[r][m] = fmaf(wv[r].y, xv.y, out[r][m]);
          out[r][m] = fmaf(wv[r].z, xv.z, out[r][m]);
          out[r][m] = fmaf(wv[r].w, xv.w, out[r][m]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) out[r][m] = warp_sum(out[r][m]);
}

// one transformer layer for stream 0; M = 2 (slow) or 1 (fast) rows
template <int M, bool FAST>
__device__ __forceinline__ void layer_staged(const ArDecodeArgs& a, const ArLayerWeights& w, int layer_idx, int cb,
                                             float* xs, Stage& sg, unsigned nblocks, Prof& pf, const float* x_in) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ArStreamDev& st = a.s[0];
  Slice s;

  // ---- phase 1: attention_norm + wqkv (+RoPE); q -> scratch, k/v -> cache
  load_rmsnorm<M>(x_in, w.attn_norm, xs, nullptr);      // x_in == a.x except right after a sampler (see the codebook loop)
  pf.tick(P_ACT);
  {
    const float* wb = begin_phase(a, sg, s);
    pf.tick(P_WAIT);
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[2] = {wb + (size_t)(u - s.u0) * 2 * D, wb + (size_t)(u - s.u0) * 2 * D + D};
      const int r = 2 * u;
      const int sec = r / D, c = r % D;
      const int h = c / HEAD_DIM, d = c % HEAD_DIM;
      // the RoPE entries are requested BEFORE the dot products: their L2 latency hides behind the math
      float2 rope_cs[M];
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const int pos = FAST ? cb : st.pos + m;
        rope_cs[m] = (lane == 0 && sec < 2)
                         ? __ldg(reinterpret_cast<const float2*>((FAST ? a.fast_rope : a.rope) + ((long long)pos * (HEAD_DIM / 2) + d / 2) * 2))
                         : make_float2(1.f, 0.f);
      }
      float o[2][M];
      warp_rows_dot_s<2, M>(rows, xs, D, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const int pos = FAST ? cb : st.pos + m;
          float v0 = o[0][m], v1 = o[1][m];
          if (sec < 2) {
            const float cs = rope_cs[m].x, sn = rope_cs[m].y;
            const float r0 = v0 * cs - v1 * sn, r1 = v1 * cs + v0 * sn;
            v0 = r0; v1 = r1;
          }
          if (sec == 0) {
            a.q[m * D + c] = v0; a.q[m * D + c + 1] = v1;
          } else {
            float* base;
            if (FAST) base = (sec == 1 ? st.fkc : st.fvc) + (((long long)layer_idx * H + h) * AR_CODEBOOKS + pos) * HEAD_DIM;
            else base = (sec == 1 ? st.kc : st.vc) + (((long long)layer_idx * H + h) * a.max_seq + pos) * HEAD_DIM;
            base[d] = v0; base[d + 1] = v1;
          }
        }
      }
    }
  }
  pf.tick(P_DOT);
  grid_sync(a.barrier, nblocks);
  pf.tick(P_SYNC);

  float* ys = xs;                          // [M][D] attention output
  if (!FAST) {
    // ---- phase 2: split-KV attention partials, work item = (head, split)
    const int nitems = H * a.nsplit;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int h = item / a.nsplit, sp = item % a.nsplit;
      const int pos = st.pos;
      const int nkeys = pos + 2;
      const int chunk = (nkeys + a.nsplit - 1) / a.nsplit;
      const int k_begin = sp * chunk, k_end = min(nkeys, k_begin + chunk);
      const float* kc = st.kc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      const float* vc = st.vc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      const float2 q0 = __ldcg(reinterpret_cast<const float2*>(a.q + h * HEAD_DIM) + lane);
      const float2 q1 = __ldcg(reinterpret_cast<const float2*>(a.q + D + h * HEAD_DIM) + lane);
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      for (int key = k_begin + warp; key < k_end; key += NW) {
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
      float* sm = xs;                      // [NW][2][PART]
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
        float* dst = a.part + (((long long)h * a.nsplit + sp) * 2 + tkn) * PART;
        if (lane == 0) { dst[0] = mm; dst[1] = ll; }
        dst[2 + 2 * lane] = ax; dst[3 + 2 * lane] = ay;
      }
      __syncthreads();
    }
    pf.tick(P_ATTN);
    grid_sync(a.barrier, nblocks);
    pf.tick(P_SYNC);
    // ---- phase 3a: every CTA merges the split partials of all (head, token) into ys
    // (all loads of an item are issued before any is used: lane sp fetches split sp's (m, l), the accumulator
    // slices are fetched in one fully unrolled, predicated batch -- one L2 round trip instead of one per split)
    for (int it = warp; it < H * 2; it += NW) {
      const int h = it / 2, tkn = it % 2;
      const float* base = a.part + ((long long)h * a.nsplit * 2 + tkn) * PART;
      float pm = -INFINITY, pl = 0.f;
      if (lane < a.nsplit) {
        const float2 ml = __ldcg(reinterpret_cast<const float2*>(base + (long long)lane * 2 * PART));
        pm = ml.x; pl = ml.y;
      }
      float2 av[16];
#pragma unroll
      for (int sp = 0; sp < 16; ++sp)
        av[sp] = sp < a.nsplit ? __ldcg(reinterpret_cast<const float2*>(base + (long long)sp * 2 * PART + 2) + lane)
                               : make_float2(0.f, 0.f);
      const float mm = warp_max(pm);
      const float cl = (pm == -INFINITY) ? 0.f : expf(pm - mm);      // this lane's split: rescale factor
      float ll = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
      for (int sp = 0; sp < 16; ++sp) {                               // fixed split order: deterministic sums
        const float c = __shfl_sync(0xffffffffu, cl, sp);
        const float l_sp = __shfl_sync(0xffffffffu, pl, sp);
        if (sp < a.nsplit) { ll += l_sp * c; ax += av[sp].x * c; ay += av[sp].y * c; }
      }
      const float inv = 1.f / ll;
      ys[tkn * D + h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[tkn * D + h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  } else {
    // ---- fast path: <= 8 keys, every CTA recomputes the attention of all heads
    for (int h = warp; h < H; h += NW) {
      const float2 qv = __ldcg(reinterpret_cast<const float2*>(a.q + h * HEAD_DIM) + lane);
      const float* kc = st.fkc + ((long long)layer_idx * H + h) * AR_CODEBOOKS * HEAD_DIM;
      const float* vc = st.fvc + ((long long)layer_idx * H + h) * AR_CODEBOOKS * HEAD_DIM;
      // q, all keys and all values are requested before anything is used: one L2 round trip instead of three
      float2 kv[AR_CODEBOOKS], vv[AR_CODEBOOKS];
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        kv[key] = vv[key] = make_float2(0.f, 0.f);
        if (key <= cb) {
          kv[key] = __ldcg(reinterpret_cast<const float2*>(kc + key * HEAD_DIM) + lane);
          vv[key] = __ldcg(reinterpret_cast<const float2*>(vc + key * HEAD_DIM) + lane);
        }
      }
      float sc[AR_CODEBOOKS];
      float mx = -INFINITY;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        float sv = -INFINITY;
        if (key <= cb) sv = warp_sum(qv.x * kv[key].x + qv.y * kv[key].y) * 0.125f;
        sc[key] = sv;
        mx = fmaxf(mx, sv);
      }
      float l = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        if (key <= cb) {
          const float p = expf(sc[key] - mx);
          l += p; ax += p * vv[key].x; ay += p * vv[key].y;
        }
      }
      const float inv = 1.f / l;
      ys[h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  }
  // ---- phase 3b: wo + residual -> h
  pf.tick(P_ATTN);
  {
    const float* wb = begin_phase(a, sg, s);
    pf.tick(P_WAIT);
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[1] = {wb + (size_t)(u - s.u0) * D};
      float res[M];                          // residual requested before the dot products (latency hidden)
#pragma unroll
      for (int m = 0; m < M; ++m) res[m] = lane == 0 ? __ldcg(a.x + m * D + u) : 0.f;
      float o[1][M];
      warp_rows_dot_s<1, M>(rows, ys, D, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) a.h[m * D + u] = res[m] + o[0][m];
      }
    }
  }
  pf.tick(P_DOT);
  grid_sync(a.barrier, nblocks);
  pf.tick(P_SYNC);

  // ---- phase 4: ffn_norm + silu(w1 h) * (w3 h) -> g
  load_rmsnorm<M>(a.h, w.ffn_norm, xs, nullptr);
  pf.tick(P_ACT);
  {
    const float* wb = begin_phase(a, sg, s);
    pf.tick(P_WAIT);
    const int n = s.u1 - s.u0;
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[2] = {wb + (size_t)(u - s.u0) * D, wb + (size_t)n * D + (size_t)(u - s.u0) * D};
      float o[2][M];
      warp_rows_dot_s<2, M>(rows, xs, D, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float t = o[0][m];
          a.g[m * I + u] = (t / (1.f + expf(-t))) * o[1][m];
        }
      }
    }
  }
  pf.tick(P_DOT);
  grid_sync(a.barrier, nblocks);
  pf.tick(P_SYNC);

  // ---- phase 5: w2 + residual -> x
  for (int i = threadIdx.x; i < M * I; i += NT) xs[i] = __ldcg(a.g + i);
  __syncthreads();
  pf.tick(P_ACT);
  {
    const float* wb = begin_phase(a, sg, s);
    pf.tick(P_WAIT);
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[1] = {wb + (size_t)(u - s.u0) * I};
      float res[M];
#pragma unroll
      for (int m = 0; m < M; ++m) res[m] = lane == 0 ? __ldcg(a.h + m * D + u) : 0.f;
      float o[1][M];
      warp_rows_dot_s<1, M>(rows, xs, I, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) a.x[m * D + u] = res[m] + o[0][m];
      }
    }
  }
  pf.tick(P_DOT);
  grid_sync(a.barrier, nblocks);
  pf.tick(P_SYNC);
}

__global__ void __launch_bounds__(NT, 1) ar_decode_staged_kernel(const ArDecodeArgs a) {
  extern __shared__ __align__(128) unsigned char dsmem[];
  Stage sg;
  sg.wbuf = dsmem;
  float* xs = reinterpret_cast<float*>(dsmem + 2 * WBUF_BYTES);
  sg.mbar = reinterpret_cast<unsigned long long*>(dsmem + 2 * WBUF_BYTES + XS_FLOATS * sizeof(float));
  SampleSmem& ssm = *reinterpret_cast<SampleSmem*>(xs);       // the sampler runs while xs is idle
  static_assert(sizeof(SampleSmem) <= XS_FLOATS * sizeof(float), "sampler scratch must fit the activation buffer");
  sg.wp = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(sg.pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.L2::evict_first.b64 %0, %1;" : "=l"(sg.pol_keep) : "f"(a.keep_fraction));
  const unsigned nblocks = gridDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = warp * gridDim.x + blockIdx.x;
  const int total_warps = NW * gridDim.x;
  const int gtid = blockIdx.x * NT + threadIdx.x;
  const ArStreamDev& st = a.s[0];

  slice_table_init();
  if (threadIdx.x == 0) {
    mbar_init(sg.mbar + 0, 1);
    mbar_init(sg.mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == NT - 32) issue(a, sg, 0);
  grid_sync_init(a.barrier);
  Prof pf;
  pf.start(a.prof);

  // ---- phase 0: the 2 input rows [cached_new_audio_emb, embedding[content_id]]
  for (int i = gtid; i < 2 * D; i += NT * gridDim.x) {
    const int j = i / D, c = i % D;
    float v;
    if (j == 0) v = __ldcg(st.x_audio + c);
    else if (st.cond_row) v = __ldcg(st.cond_row + c);
    else v = __ldg(a.cond_emb + (*st.content_id) * D + c);
    a.x[i] = v;
  }
  grid_sync(a.barrier, nblocks);

  pf.tick(P_MISC);
  for (int l = 0; l < AR_LAYERS; ++l) layer_staged<2, false>(a, a.slow[l], l, 0, xs, sg, nblocks, pf, a.x);

  if (a.dbg_slow_logits) {
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
  // fast residual stream starts from the PRE-norm hidden state of the last token (dual_ar_stream.py:354-355)
  // (row 0 of x, the first token's residual, is dead after the slow stack: the hidden row moves there directly)
  for (int i = gtid; i < D; i += NT * gridDim.x) a.x[i] = __ldcg(a.x + D + i);
  grid_sync(a.barrier, nblocks);

  pf.tick(P_MISC);
  // Every CTA runs every sampler itself: same logits, same noise -> same token everywhere, so no barrier is needed to
  // publish it.  CTA 0 additionally stores the code and the token's embedding row into x (the residual the next wo
  // phase reads, ordered by the QKV-phase barrier); the next layer-0 norm reads the embedding row directly.
  __shared__ int s_toks[AR_CODEBOOKS];
  const float* x_in = a.x;
  for (int cb = 0; cb < AR_CODEBOOKS; ++cb) {
    for (int l = 0; l < AR_FAST_LAYERS; ++l) layer_staged<1, true>(a, a.fast[l], l, cb, xs, sg, nblocks, pf, l == 0 ? x_in : a.x);
    load_rmsnorm<1>(a.x, a.fast_norm_w, xs, nullptr);
    pf.tick(P_ACT);
    {
      Slice s;
      const float* wb = begin_phase(a, sg, s);
      pf.tick(P_WAIT);
      for (int u = s.u0 + warp; u < s.u1; u += NW) {
        const float* rows[1] = {wb + (size_t)(u - s.u0) * D};
        float o[1][1];
        warp_rows_dot_s<1, 1>(rows, xs, D, o);
        if (lane == 0) a.logits[u] = o[0][0];
      }
    }
    pf.tick(P_DOT);
    grid_sync(a.barrier, nblocks);
    pf.tick(P_SYNC);
    if (blockIdx.x == 0 && a.dbg_fast_logits)
      for (int i = threadIdx.x; i < AR_CB_SIZE; i += NT) a.dbg_fast_logits[cb * AR_CB_SIZE + i] = __ldcg(a.logits + i);
    const float* noise = st.noise ? st.noise + cb * AR_CB_SIZE : nullptr;
    const int tok = sample_topp(a.logits, noise, st.seed, st.step, cb + 1, st.temperature, st.top_p, ssm);
    if (threadIdx.x == 0) s_toks[cb] = tok;
    x_in = a.fast_emb + (long long)tok * D;
    if (blockIdx.x == 0) {
      if (threadIdx.x == 0) st.out_codes[cb] = tok;
      for (int i = threadIdx.x; i < D; i += NT) a.x[i] = __ldg(x_in + i);
    }
    pf.tick(P_SAMPLE);
  }

  // ---- cached_new_audio_emb = embed(pred codes)  (dual_ar_stream.py:245-255, 834)
  __syncthreads();
  for (int c = gtid; c < D; c += NT * gridDim.x) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < AR_CODEBOOKS; ++k) s += __ldg(a.codebook_emb + ((long long)s_toks[k] + k * AR_CB_SIZE) * D + c);
    st.x_audio[c] = s;
  }
  pf.tick(P_MISC);
  grid_sync_finish(a.barrier);
}

}  // namespace

bool ar_decode_staged_supported(int grid) {
  // every per-CTA slice must fit one 96 KB staging buffer: w1|w3 needs ceil(2304/grid) * 6 KB
  return ((AR_INTER + grid - 1) / grid) * 2 * AR_DIM * 4 <= WBUF_BYTES;
}

void launch_ar_decode_staged(const ArDecodeArgs& args, int grid, cudaStream_t st) {
  const size_t smem = (size_t)2 * WBUF_BYTES + XS_FLOATS * sizeof(float) + 64;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(ar_decode_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  void* kargs[] = {(void*)&args};
  SV_CUDA(cudaLaunchCooperativeKernel((void*)ar_decode_staged_kernel, dim3(grid), dim3(NT), kargs, smem, st));
  ++g_kernel_launches;
}

}  // namespace svanon
