// Stage A, batch 1: persistent decode kernel with TMA-staged weights and NO grid barriers.
//
// ar_decode_staged.cu spends ~70 % of its time in ~207 grid barriers: every phase ends with "store outputs, fence,
// arrive, poll a counter" and the next phase starts with "reload the activations from L2" -- two dependent L2 round
// trips plus a fence per phase.  Here every activation that crosses CTAs travels as a self-validating 8-byte word
// {fp32 value, 32-bit tag} (the flag-in-data idea of NCCL's LL protocol): producers write `st.volatile.v2`,
// consumers spin on `ld.volatile.v2` until the tag equals the tag of the phase that must have produced the value.
// One L2 round trip per phase, no fences, no atomics, no barrier.  tag = (launch epoch << 12) | phase index, so a
// stale word can never be mistaken for a fresh one.  A buffer written in phase p is only rewritten >= 2 all-to-all
// phases later, by which time every CTA has provably finished reading it.
//
// The KV cache itself stays plain fp32 for later launches; the K/V rows produced in this launch are ADDITIONALLY
// published in tagged form for the attention phase of the same launch (slow: 2 new rows per layer; fast: the 8-slot
// cache of the frame).  Weight staging (cp.async.bulk + mbarrier, one phase ahead) is as in ar_decode_staged.cu.
#include "ar_decode_common.cuh"

namespace svanon {

using namespace ardec;

namespace {

constexpr int WBUF_BYTES = 98304;
constexpr int XS_FLOATS = 2 * AR_INTER;
constexpr int N_SLOW_WP = AR_LAYERS * 4;
constexpr int WP_PER_CB = AR_FAST_LAYERS * 4 + 1;
constexpr int N_WP = N_SLOW_WP + AR_CODEBOOKS * WP_PER_CB;

enum Kind : int { K_QKV = 0, K_WO = 1, K_W13 = 2, K_W2 = 3, K_LOGITS = 4 };

struct Slice {
  const float* src[2];
  int bytes_per_unit;
  int regions;
  int u0, u1;
  bool fast;
};

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
  int U;
  s.regions = 1;
  s.src[1] = nullptr;
  switch (kind) {
    case K_QKV: U = 3 * D / 2; s.src[0] = lw->wqkv; s.bytes_per_unit = 2 * D * 4; break;
    case K_WO: U = D; s.src[0] = lw->wo; s.bytes_per_unit = D * 4; break;
    case K_W13: U = I; s.src[0] = lw->w1; s.src[1] = lw->w3; s.bytes_per_unit = D * 4; s.regions = 2; break;
    case K_W2: U = D; s.src[0] = lw->w2; s.bytes_per_unit = I * 4; break;
    default: U = AR_CB_SIZE; s.src[0] = a.fast_output_w; s.bytes_per_unit = D * 4; break;
  }
  s.u0 = (int)((long long)U * blockIdx.x / gridDim.x);
  s.u1 = (int)((long long)U * (blockIdx.x + 1) / gridDim.x);
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
      "LL_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni LL_WAIT_DONE;\n"
      "bra.uni LL_WAIT_LOOP;\n"
      "LL_WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                             unsigned long long policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// ---- tagged 8-byte words
__device__ __forceinline__ void ll_store(uint2* p, float v, unsigned tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
// Watchdog: a protocol bug must not wedge the GPU.  After ~4M unsuccessful polls (seconds) the launch is
// declared dead: the flag makes every later poll return immediately and the host reports the failure.
__device__ int g_ll_abort = 0;

__device__ __forceinline__ float ll_load(const uint2* p, unsigned tag) {
  unsigned v, t;
  unsigned spins = 0;
  while (true) {
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(t) : "l"(p) : "memory");
    if (t == tag) break;
    if ((++spins & 0x3FFFu) == 0) {
      if (*reinterpret_cast<volatile int*>(&g_ll_abort) != 0 || spins > (1u << 22)) {
        g_ll_abort = 1;
        break;
      }
    }
  }
  return __uint_as_float(v);
}

// N tagged words at base[i*stride]: all loads are issued back to back (they overlap in flight) and only then
// checked; a miss re-polls the whole batch.  One L2 round trip when the data is already there.
template <int N>
__device__ __forceinline__ void ll_load_n(const uint2* base, int stride, unsigned tag, float (&out)[N]) {
  unsigned v[N], t[N];
  unsigned spins = 0;
  while (true) {
#pragma unroll
    for (int i = 0; i < N; ++i)
      asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[i]), "=r"(t[i]) : "l"(base + (size_t)i * stride));
    bool ok = true;
#pragma unroll
    for (int i = 0; i < N; ++i) ok = ok && (t[i] == tag);
    if (ok) break;
    if ((++spins & 0x3FFu) == 0) {
      if (*reinterpret_cast<volatile int*>(&g_ll_abort) != 0 || spins > (1u << 20)) {
        g_ll_abort = 1;
        break;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < N; ++i) out[i] = __uint_as_float(v[i]);
}

struct Stage {
  unsigned char* wbuf;
  unsigned long long* mbar;
  unsigned long long pol_stream, pol_keep;
  int wp;
};

__device__ __forceinline__ void issue(const ArDecodeArgs& a, Stage& sg, int wp) {
  if (wp >= N_WP) return;
  const Slice s = slice_of(a, wp);
  const int n = s.u1 - s.u0;
  if (n <= 0) return;
  unsigned char* dst = sg.wbuf + (size_t)(wp & 1) * WBUF_BYTES;
  unsigned long long* bar = sg.mbar + (wp & 1);
  const unsigned region_bytes = (unsigned)n * s.bytes_per_unit;
  mbar_expect_tx(bar, region_bytes * s.regions);
  const unsigned long long pol = s.fast ? sg.pol_keep : sg.pol_stream;
  for (int r = 0; r < s.regions; ++r) {
    const unsigned char* src = reinterpret_cast<const unsigned char*>(s.src[r]) + (size_t)s.u0 * s.bytes_per_unit;
    for (unsigned off = 0; off < region_bytes; off += 32768) {
      const unsigned len = min(32768u, region_bytes - off);
      tma_bulk_g2s(dst + (size_t)r * region_bytes + off, src + off, len, bar, pol);
    }
  }
}

// All threads of the CTA have passed a __syncthreads since they last read buffer (wp+1)&1 (every weight phase
// ends with one before the next phase's activations are loaded), so refilling it here is safe.
__device__ __forceinline__ const float* begin_phase(const ArDecodeArgs& a, Stage& sg, Slice& s) {
  const int wp = sg.wp;
  if (threadIdx.x == 0) issue(a, sg, wp + 1);
  s = slice_of(a, wp);
  if (s.u1 > s.u0) mbar_wait(sg.mbar + (wp & 1), (wp >> 1) & 1);
  sg.wp = wp + 1;
  return reinterpret_cast<const float*>(sg.wbuf + (size_t)(wp & 1) * WBUF_BYTES);
}

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

// RMSNorm of M tagged rows into shared memory (polls until the rows have been published)
template <int M>
__device__ __forceinline__ void load_rmsnorm_ll(const uint2* x, unsigned tag, const float* w, float* xs) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();                         // previous users of xs are done
  for (int m = warp; m < M; m += NW) {
    float v[D / 32];
    float s = 0.f;
    ll_load_n<D / 32>(x + m * D + lane, 32, tag, v);
#pragma unroll
    for (int i = 0; i < D / 32; ++i) s = fmaf(v[i], v[i], s);
    s = warp_sum(s);
    const float inv = rsqrtf(s / D + AR_NORM_EPS);
#pragma unroll
    for (int i = 0; i < D / 32; ++i) xs[m * D + lane + 32 * i] = v[i] * inv * __ldg(w + lane + 32 * i);
  }
  __syncthreads();
}

struct Tags {
  unsigned base;     // epoch << 12
  unsigned ph;       // next phase index (uniform across the grid: same static schedule everywhere)
  unsigned x, h, q, g, part, logits;
  __device__ __forceinline__ unsigned next() { return base | (++ph); }
};

// tagged scratch, carved out of a.ll (uint2 words)
struct LL {
  uint2 *x, *h, *q, *knew, *vnew, *g, *part, *logits, *fkv;
};
__device__ __forceinline__ LL carve(uint2* p) {
  LL l;
  l.x = p; p += 2 * D;
  l.h = p; p += 2 * D;
  l.q = p; p += 2 * D;
  l.knew = p; p += 2 * D;
  l.vnew = p; p += 2 * D;
  l.g = p; p += 2 * I;
  l.part = p; p += H * 16 * 2 * PART;
  l.logits = p; p += 1024;
  l.fkv = p;                                 // [4 layers][8 slots][2 (k,v)][768]
  return l;
}

template <int M, bool FAST>
__device__ __forceinline__ void layer_ll(const ArDecodeArgs& a, const ArLayerWeights& w, int layer_idx, int cb,
                                         float* xs, Stage& sg, Tags& tg, const LL& ll, const uint2* x_src) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const ArStreamDev& st = a.s[0];
  Slice s;
  const unsigned tag_x_in = tg.x;

  // ---- phase 1: attention_norm + wqkv (+RoPE); q and the new k/v rows are published tagged, k/v also go to the cache
  load_rmsnorm_ll<M>(x_src, tag_x_in, w.attn_norm, xs);
  {
    const float* wb = begin_phase(a, sg, s);
    const unsigned tag = tg.next();
    tg.q = tag;
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[2] = {wb + (size_t)(u - s.u0) * 2 * D, wb + (size_t)(u - s.u0) * 2 * D + D};
      float o[2][M];
      warp_rows_dot_s<2, M>(rows, xs, D, o);
      if (lane == 0) {
        const int r = 2 * u;
        const int sec = r / D, c = r % D;
        const int h = c / HEAD_DIM, d = c % HEAD_DIM;
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const int pos = FAST ? cb : st.pos + m;
          float v0 = o[0][m], v1 = o[1][m];
          if (sec < 2) {
            const float* tab = (FAST ? a.fast_rope : a.rope) + ((long long)pos * (HEAD_DIM / 2) + d / 2) * 2;
            const float cs = __ldg(tab), sn = __ldg(tab + 1);
            const float r0 = v0 * cs - v1 * sn, r1 = v1 * cs + v0 * sn;
            v0 = r0; v1 = r1;
          }
          if (sec == 0) {
            ll_store(ll.q + m * D + c, v0, tag);
            ll_store(ll.q + m * D + c + 1, v1, tag);
          } else if (FAST) {
            // fast cache of this frame lives only in tagged form: slot tag identifies (epoch, layer, slot)
            uint2* dst = ll.fkv + (((size_t)layer_idx * AR_CODEBOOKS + pos) * 2 + (sec - 1)) * D + c;
            const unsigned ftag = tg.base | (unsigned)(2048 + layer_idx * AR_CODEBOOKS + pos);
            ll_store(dst, v0, ftag);
            ll_store(dst + 1, v1, ftag);
          } else {
            float* base = (sec == 1 ? st.kc : st.vc) + (((long long)layer_idx * H + h) * a.max_seq + pos) * HEAD_DIM;
            base[d] = v0; base[d + 1] = v1;
            uint2* dst = (sec == 1 ? ll.knew : ll.vnew) + m * D + c;
            ll_store(dst, v0, tag);
            ll_store(dst + 1, v1, tag);
          }
        }
      }
    }
  }

  float* ys = xs;
  if (!FAST) {
    // ---- phase 2: split-KV attention partials, work item = (head, split); keys < pos come from the cache
    //      (written by earlier launches), keys pos and pos+1 from the tagged rows of phase 1
    const unsigned tag_part = tg.next();
    tg.part = tag_part;
    const int nitems = H * a.nsplit;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int h = item / a.nsplit, sp = item % a.nsplit;
      const int pos = st.pos;
      const int nkeys = pos + 2;
      const int chunk = (nkeys + a.nsplit - 1) / a.nsplit;
      const int k_begin = sp * chunk, k_end = min(nkeys, k_begin + chunk);
      const float* kc = st.kc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      const float* vc = st.vc + ((long long)layer_idx * H + h) * a.max_seq * HEAD_DIM;
      float2 q0, q1;
      {
        float t0[2], t1[2];
        ll_load_n<2>(ll.q + h * HEAD_DIM + 2 * lane, 1, tg.q, t0);
        ll_load_n<2>(ll.q + D + h * HEAD_DIM + 2 * lane, 1, tg.q, t1);
        q0 = make_float2(t0[0], t0[1]);
        q1 = make_float2(t1[0], t1[1]);
      }
      float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
      float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
      for (int key = k_begin + warp; key < k_end; key += NW) {
        float2 kv, vv;
        if (key < pos) {
          kv = __ldcg(reinterpret_cast<const float2*>(kc + (long long)key * HEAD_DIM) + lane);
          vv = __ldcg(reinterpret_cast<const float2*>(vc + (long long)key * HEAD_DIM) + lane);
        } else {
          const int mrow = key - pos;
          float tk[2], tv[2];
          ll_load_n<2>(ll.knew + mrow * D + h * HEAD_DIM + 2 * lane, 1, tg.q, tk);
          ll_load_n<2>(ll.vnew + mrow * D + h * HEAD_DIM + 2 * lane, 1, tg.q, tv);
          kv = make_float2(tk[0], tk[1]);
          vv = make_float2(tv[0], tv[1]);
        }
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
        float lsum = 0.f, ax = 0.f, ay = 0.f;
        for (int ww = 0; ww < NW; ++ww) {
          const float* pp = sm + (ww * 2 + tkn) * PART;
          const float c = (pp[0] == -INFINITY) ? 0.f : expf(pp[0] - mm);
          lsum += pp[1] * c; ax += pp[2 + 2 * lane] * c; ay += pp[3 + 2 * lane] * c;
        }
        uint2* dst = ll.part + (((size_t)h * a.nsplit + sp) * 2 + tkn) * PART;
        if (lane == 0) { ll_store(dst, mm, tag_part); ll_store(dst + 1, lsum, tag_part); }
        ll_store(dst + 2 + 2 * lane, ax, tag_part);
        ll_store(dst + 3 + 2 * lane, ay, tag_part);
      }
      __syncthreads();
    }
    // ---- phase 3a: every CTA merges the partials of all (head, token) into ys
    __syncthreads();
    for (int it = warp; it < H * 2; it += NW) {
      const int h = it / 2, tkn = it % 2;
      const uint2* base = ll.part + ((size_t)h * a.nsplit * 2 + tkn) * PART;
      // four strided batches over the splits: m, l, acc.x, acc.y  (unused splits of the 16-wide batch are clamped)
      float pm[16], pl[16], pax[16], pay[16];
      {
        const int ns = a.nsplit;
        const int stride = 2 * PART;
        // clamp: entries >= ns re-read split ns-1 (valid tag) and are ignored below
        const uint2* b0 = base;
        unsigned spins = 0;
        while (true) {
          unsigned vm[16], tm[16], vl[16], tl[16], vx[16], tx[16], vy[16], ty[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const uint2* pp = b0 + (size_t)min(i, ns - 1) * stride;
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vm[i]), "=r"(tm[i]) : "l"(pp));
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vl[i]), "=r"(tl[i]) : "l"(pp + 1));
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vx[i]), "=r"(tx[i]) : "l"(pp + 2 + 2 * lane));
            asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(vy[i]), "=r"(ty[i]) : "l"(pp + 3 + 2 * lane));
          }
          bool ok = true;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            ok = ok && tm[i] == tag_part && tl[i] == tag_part && tx[i] == tag_part && ty[i] == tag_part;
          if (ok) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              pm[i] = __uint_as_float(vm[i]); pl[i] = __uint_as_float(vl[i]);
              pax[i] = __uint_as_float(vx[i]); pay[i] = __uint_as_float(vy[i]);
            }
            break;
          }
          if ((++spins & 0x3FFu) == 0) {
            if (*reinterpret_cast<volatile int*>(&g_ll_abort) != 0 || spins > (1u << 20)) { g_ll_abort = 1; break; }
          }
        }
      }
      float mm = -INFINITY;
#pragma unroll
      for (int sp = 0; sp < 16; ++sp) if (sp < a.nsplit) mm = fmaxf(mm, pm[sp]);
      float lsum = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
      for (int sp = 0; sp < 16; ++sp) {
        if (sp < a.nsplit) {
          const float c = (pm[sp] == -INFINITY) ? 0.f : expf(pm[sp] - mm);
          lsum += pl[sp] * c; ax += pax[sp] * c; ay += pay[sp] * c;
        }
      }
      const float inv = 1.f / lsum;
      ys[tkn * D + h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[tkn * D + h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  } else {
    // ---- fast path: <= 8 keys from the tagged per-frame cache; every CTA recomputes all heads
    __syncthreads();
    for (int h = warp; h < H; h += NW) {
      float2 qv;
      {
        float t0[2];
        ll_load_n<2>(ll.q + h * HEAD_DIM + 2 * lane, 1, tg.q, t0);
        qv = make_float2(t0[0], t0[1]);
      }
      float sc[AR_CODEBOOKS];
      float2 vvs[AR_CODEBOOKS];
      float2 kks[AR_CODEBOOKS];
      // keys 0..cb-1 were published in earlier codebook steps (valid long ago): issue all their loads at once
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        kks[key] = make_float2(0.f, 0.f);
        vvs[key] = make_float2(0.f, 0.f);
        if (key <= cb) {
          const uint2* kp = ll.fkv + (((size_t)layer_idx * AR_CODEBOOKS + key) * 2 + 0) * D + h * HEAD_DIM + 2 * lane;
          const unsigned ftag = tg.base | (unsigned)(2048 + layer_idx * AR_CODEBOOKS + key);
          float tk[2], tv[2];
          ll_load_n<2>(kp, 1, ftag, tk);
          ll_load_n<2>(kp + D, 1, ftag, tv);
          kks[key] = make_float2(tk[0], tk[1]);
          vvs[key] = make_float2(tv[0], tv[1]);
        }
      }
      float mx = -INFINITY;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        float sv = -INFINITY;
        if (key <= cb) sv = warp_sum(qv.x * kks[key].x + qv.y * kks[key].y) * 0.125f;
        sc[key] = sv;
        mx = fmaxf(mx, sv);
      }
      float l = 0.f, ax = 0.f, ay = 0.f;
#pragma unroll
      for (int key = 0; key < AR_CODEBOOKS; ++key) {
        if (key <= cb) {
          const float p = expf(sc[key] - mx);
          l += p; ax += p * vvs[key].x; ay += p * vvs[key].y;
        }
      }
      const float inv = 1.f / l;
      ys[h * HEAD_DIM + 2 * lane] = ax * inv;
      ys[h * HEAD_DIM + 2 * lane + 1] = ay * inv;
    }
    __syncthreads();
  }
  // ---- phase 3b: wo + residual -> h
  {
    const float* wb = begin_phase(a, sg, s);
    const unsigned tag = tg.next();
    tg.h = tag;
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      // the residual was published a phase ago: start its load now, use it after the dot product
      unsigned rv[M], rt[M];
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(rv[m]), "=r"(rt[m]) : "l"(x_src + m * D + u));
      }
      const float* rows[1] = {wb + (size_t)(u - s.u0) * D};
      float o[1][M];
      warp_rows_dot_s<1, M>(rows, ys, D, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float r = (rt[m] == tag_x_in) ? __uint_as_float(rv[m]) : ll_load(x_src + m * D + u, tag_x_in);
          ll_store(ll.h + m * D + u, r + o[0][m], tag);
        }
      }
    }
  }

  // ---- phase 4: ffn_norm + silu(w1 h) * (w3 h) -> g
  load_rmsnorm_ll<M>(ll.h, tg.h, w.ffn_norm, xs);
  {
    const float* wb = begin_phase(a, sg, s);
    const unsigned tag = tg.next();
    tg.g = tag;
    const int n = s.u1 - s.u0;
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      const float* rows[2] = {wb + (size_t)(u - s.u0) * D, wb + (size_t)n * D + (size_t)(u - s.u0) * D};
      float o[2][M];
      warp_rows_dot_s<2, M>(rows, xs, D, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float t = o[0][m];
          ll_store(ll.g + m * I + u, (t / (1.f + expf(-t))) * o[1][m], tag);
        }
      }
    }
  }

  // ---- phase 5: w2 + residual -> x (row 0.. of the x buffer)
  __syncthreads();
  {
    constexpr int PER = (M * I + NT - 1) / NT;
    constexpr int FULL = (M * I) / NT;               // batches that are complete for every thread
    float gv[FULL > 0 ? FULL : 1];
    ll_load_n<FULL>(ll.g + threadIdx.x, NT, tg.g, gv);
#pragma unroll
    for (int k = 0; k < FULL; ++k) xs[threadIdx.x + k * NT] = gv[k];
    if (PER > FULL) {
      const int i = threadIdx.x + FULL * NT;
      if (i < M * I) xs[i] = ll_load(ll.g + i, tg.g);
    }
  }
  __syncthreads();
  {
    const float* wb = begin_phase(a, sg, s);
    const unsigned tag = tg.next();
    for (int u = s.u0 + warp; u < s.u1; u += NW) {
      unsigned rv[M], rt[M];
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m)
          asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(rv[m]), "=r"(rt[m]) : "l"(ll.h + m * D + u));
      }
      const float* rows[1] = {wb + (size_t)(u - s.u0) * I};
      float o[1][M];
      warp_rows_dot_s<1, M>(rows, xs, I, o);
      if (lane == 0) {
#pragma unroll
        for (int m = 0; m < M; ++m) {
          const float r = (rt[m] == tg.h) ? __uint_as_float(rv[m]) : ll_load(ll.h + m * D + u, tg.h);
          ll_store(ll.x + m * D + u, r + o[0][m], tag);
        }
      }
    }
    tg.x = tag;
  }
}

__global__ void __launch_bounds__(NT, 1) ar_decode_ll_kernel(const ArDecodeArgs a) {
  extern __shared__ __align__(128) unsigned char dsmem[];
  Stage sg;
  sg.wbuf = dsmem;
  float* xs = reinterpret_cast<float*>(dsmem + 2 * WBUF_BYTES);
  sg.mbar = reinterpret_cast<unsigned long long*>(dsmem + 2 * WBUF_BYTES + XS_FLOATS * sizeof(float));
  SampleSmem& ssm = *reinterpret_cast<SampleSmem*>(xs);
  static_assert(sizeof(SampleSmem) <= XS_FLOATS * sizeof(float), "sampler scratch must fit the activation buffer");
  sg.wp = 0;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(sg.pol_stream));
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(sg.pol_keep));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gwarp = warp * gridDim.x + blockIdx.x;
  const int total_warps = NW * gridDim.x;
  const int gtid = blockIdx.x * NT + threadIdx.x;
  const ArStreamDev& st = a.s[0];
  const LL ll = carve(reinterpret_cast<uint2*>(a.ll));
  Tags tg;
  tg.base = a.epoch << 12;
  tg.ph = 0;

  if (threadIdx.x == 0) {
    mbar_init(sg.mbar + 0, 1);
    mbar_init(sg.mbar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) issue(a, sg, 0);

  // ---- phase 0: publish the 2 input rows [cached_new_audio_emb, embedding[content_id]]
  {
    const unsigned tag = tg.next();
    for (int i = gtid; i < 2 * D; i += NT * gridDim.x) {
      const int j = i / D, c = i % D;
      float v;
      if (j == 0) v = __ldcg(st.x_audio + c);
      else if (st.cond_row) v = __ldcg(st.cond_row + c);
      else v = __ldg(a.cond_emb + (*st.content_id) * D + c);
      ll_store(ll.x + i, v, tag);
    }
    tg.x = tag;
  }

  for (int l = 0; l < AR_LAYERS; ++l) layer_ll<2, false>(a, a.slow[l], l, 0, xs, sg, tg, ll, ll.x);

  if (a.dbg_slow_logits) {
    load_rmsnorm_ll<1>(ll.x + D, tg.x, a.norm_w, xs);
    for (int row = gwarp; row < AR_VOCAB; row += total_warps) {
      const float* rows[1] = {a.output_w + (long long)row * D};
      float o[1][1];
      warp_rows_dot<1, 1>(rows, xs, D, o);
      if (lane == 0) a.dbg_slow_logits[row] = o[0][0];
    }
    if (a.dbg_hidden && blockIdx.x == 0)
      for (int i = threadIdx.x; i < D; i += NT) a.dbg_hidden[i] = ll_load(ll.x + D + i, tg.x);
    __syncthreads();
  }

  // the fast stack starts from the PRE-norm hidden state of the last slow token = row 1 of x
  for (int cb = 0; cb < AR_CODEBOOKS; ++cb) {
    for (int l = 0; l < AR_FAST_LAYERS; ++l)
      layer_ll<1, true>(a, a.fast[l], l, cb, xs, sg, tg, ll, (cb == 0 && l == 0) ? ll.x + D : ll.x);
    load_rmsnorm_ll<1>(ll.x, tg.x, a.fast_norm_w, xs);
    {
      Slice s;
      const float* wb = begin_phase(a, sg, s);
      const unsigned tag = tg.next();
      tg.logits = tag;
      for (int u = s.u0 + warp; u < s.u1; u += NW) {
        const float* rows[1] = {wb + (size_t)(u - s.u0) * D};
        float o[1][1];
        warp_rows_dot_s<1, 1>(rows, xs, D, o);
        if (lane == 0) ll_store(ll.logits + u, o[0][0], tag);
      }
    }
    const unsigned tag_x = tg.next();
    if (blockIdx.x == 0) {
      __syncthreads();
      {
        float lv[2] = {0.f, 0.f};
        if (threadIdx.x + NT < AR_CB_SIZE) ll_load_n<2>(ll.logits + threadIdx.x, NT, tg.logits, lv);
        else lv[0] = ll_load(ll.logits + threadIdx.x, tg.logits);
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const int i = threadIdx.x + k * NT;
          if (i < AR_CB_SIZE) {
            a.logits[i] = lv[k];
            if (a.dbg_fast_logits) a.dbg_fast_logits[cb * AR_CB_SIZE + i] = lv[k];
          }
        }
      }
      __syncthreads();
      const float* noise = st.noise ? st.noise + cb * AR_CB_SIZE : nullptr;
      const int tok = sample_topp(a.logits, noise, st.seed, st.step, cb + 1, a.temperature, a.top_p, ssm);
      if (threadIdx.x == 0) st.out_codes[cb] = tok;
      for (int i = threadIdx.x; i < D; i += NT) ll_store(ll.x + i, __ldg(a.fast_emb + (long long)tok * D + i), tag_x);
      if (cb == AR_CODEBOOKS - 1) {
        // cached_new_audio_emb = embed(pred codes)  (dual_ar_stream.py:245-255, 834); CTA 0 knows all 8 codes
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += NT) {
          float s = 0.f;
#pragma unroll
          for (int k = 0; k < AR_CODEBOOKS; ++k) {
            const int code = st.out_codes[k];
            s += __ldg(a.codebook_emb + ((long long)code + k * AR_CB_SIZE) * D + c);
          }
          st.x_audio[c] = s;
        }
      }
    }
    tg.x = tag_x;
  }
}

}  // namespace

int ar_decode_ll_aborted() {
  int v = 0;
  SV_CUDA(cudaMemcpyFromSymbol(&v, g_ll_abort, sizeof(int)));
  return v;
}

size_t ar_decode_ll_scratch_words() {
  return (size_t)5 * 2 * AR_DIM + 2 * AR_INTER + (size_t)AR_HEADS * 16 * 2 * (2 + HEAD_DIM) + 1024 +
         (size_t)AR_FAST_LAYERS * AR_CODEBOOKS * 2 * AR_DIM;
}

void launch_ar_decode_ll(const ArDecodeArgs& args, int grid, cudaStream_t st) {
  const size_t smem = (size_t)2 * WBUF_BYTES + XS_FLOATS * sizeof(float) + 64;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(ar_decode_ll_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  void* kargs[] = {(void*)&args};
  SV_CUDA(cudaLaunchCooperativeKernel((void*)ar_decode_ll_kernel, dim3(grid), dim3(NT), kargs, smem, st));
  ++g_kernel_launches;
}

}  // namespace svanon
