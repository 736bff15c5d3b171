// Device helpers shared by the two AR decode kernels (ar_decode.cu: direct global weight loads, any batch;
// ar_decode_staged.cu: TMA bulk-copy weight staging, batch 1).
#pragma once
#include "ar_decode.cuh"

namespace svanon {
namespace ardec {

constexpr int NT = 512;
constexpr int NW = NT / 32;
constexpr int D = AR_DIM;
constexpr int I = AR_INTER;
constexpr int H = AR_HEADS;
constexpr int PART = 2 + HEAD_DIM;     // (m, l, acc[64]) per (stream, head, split, token)

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Grid barrier; co-residency is guaranteed by the cooperative launch.  One monotonically increasing arrival counter
// (bar[0]): arriving is a fire-and-forget `red.release`, waiting an acquire-poll until the counter reaches
// base + k * nblocks.  bar[1] carries the counter across launches.  (A per-CTA epoch-word barrier and a barrier-free
// flag-in-data exchange were measured slower in round 1 -- profiles/README.md -- and live in git history only.)
__device__ __forceinline__ unsigned& grid_target() {
  __shared__ unsigned t;
  return t;
}
__device__ __forceinline__ void grid_sync_init(unsigned* bar) {
  if (threadIdx.x == 0) {
    unsigned base;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(base) : "l"(bar + 1) : "memory");
    grid_target() = base;
  }
  __syncthreads();
}
__device__ __forceinline__ void grid_sync(unsigned* bar, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = grid_target() + nblocks;
    grid_target() = t;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
    } while ((int)(v - t) < 0);
  }
  __syncthreads();
}
// every CTA has passed the last barrier's arrive before any CTA can get here, so the counter is final
__device__ __forceinline__ void grid_sync_finish(unsigned* bar) {
  if (blockIdx.x == 0 && threadIdx.x == 0) bar[1] = grid_target();
}

// dot products of NR weight rows (length K, row-major, streamed) with M activation vectors held in shared
// memory ([M][K]); every lane ends up with all NR*M sums.
template <int NR, int M>
__device__ __forceinline__ void warp_rows_dot(const float* const* w, const float* xs, int K, float (&out)[NR][M]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) out[r][m] = 0.f;
  // K is a multiple of 768 = 6 * 128
  for (int k0 = 0; k0 < K; k0 += 768) {
    float4 wv[NR][6];
#pragma unroll
    for (int r = 0; r < NR; ++r)
#pragma unroll
      for (int i = 0; i < 6; ++i) wv[r][i] = ld_stream(w[r] + k0 + (lane + 32 * i) * 4);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int m = 0; m < M; ++m) {
        const float4 xv = *reinterpret_cast<const float4*>(xs + m * K + k0 + (lane + 32 * i) * 4);
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          out[r][m] = fmaf(wv[r][i].x, xv.x, out[r][m]);
          out[r][m] = fmaf(wv[r][i].y, xv.y, out[r][m]);
          out[r][m] = fmaf(wv[r][i].z, xv.z, out[r][m]);
          out[r][m] = fmaf(wv[r][i].w, xv.w, out[r][m]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NR; ++r)
#pragma unroll
    for (int m = 0; m < M; ++m) out[r][m] = warp_sum(out[r][m]);
}

// RMSNorm of M rows (global, written by other CTAs -> L1-bypassing loads) into shared memory.
template <int M>
__device__ __forceinline__ void load_rmsnorm(const float* x, const float* w, float* xs, float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int m = warp; m < M; m += NW) {
    float v[D / 32];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < D / 32; ++i) {
      v[i] = __ldcg(x + m * D + lane + 32 * i);
      s = fmaf(v[i], v[i], s);
    }
    s = warp_sum(s);
    const float inv = rsqrtf(s / D + AR_NORM_EPS);
#pragma unroll
    for (int i = 0; i < D / 32; ++i) xs[m * D + lane + 32 * i] = v[i] * inv * __ldg(w + lane + 32 * i);
  }
  (void)red;
  __syncthreads();
}

// ---------------------------------------------------------------- sampler (one CTA)
// logits_to_probs + multinomial_sample_one_no_sync, dual_ar_stream.py:1092-1132, V = 1000 padded to 1024.
struct SampleSmem {
  float key[1024];
  int idx[1024];
  double scan[NT];
  float redf[NW];
  int redi[NW];
  float bc[2];
};

__device__ __forceinline__ unsigned philox_round_mix(unsigned long long seed, unsigned a, unsigned b, unsigned c) {
  // counter-based generator for the production (no tape) path: Philox-2x32-like mixing, 10 rounds
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
  unsigned x0 = a ^ (c * 0x9E3779B9u), x1 = b;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p = (unsigned long long)0xD256D193u * x0;
    x0 = ((unsigned)(p >> 32)) ^ x1 ^ k0;
    x1 = (unsigned)p;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
    x0 ^= k1;
  }
  return x0;
}

static __device__ int sample_topp(const float* logits_g, const float* noise, unsigned long long seed, unsigned step,
                           unsigned slot, float temperature, float top_p, SampleSmem& s) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 1024; i += NT) {
    s.key[i] = (i < AR_CB_SIZE) ? __ldcg(logits_g + i) : -INFINITY;
    s.idx[i] = i;
  }
  __syncthreads();
  // bitonic sort, descending by key (ties: lower index first)
  for (int k = 2; k <= 1024; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < 512; t += NT) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const bool desc = ((lo & k) == 0);
        const float a = s.key[lo], b = s.key[hi];
        const int ia = s.idx[lo], ib = s.idx[hi];
        const bool a_first = (a > b) || (a == b && ia < ib);      // a should precede b in descending order
        if (a_first != desc) {
          s.key[lo] = b; s.key[hi] = a;
          s.idx[lo] = ib; s.idx[hi] = ia;
        }
      }
      __syncthreads();
    }
  }
  // softmax over the sorted logits, cumulative sum (fp64 accumulate like ATen's CPU cumsum), top-p mask
  const float mx = s.key[0];
  const int i0 = tid * 2;
  const float e0 = expf(s.key[i0] - mx), e1 = expf(s.key[i0 + 1] - mx);     // exp(-inf) = 0 for the padding
  float part = e0 + e1;
  part = warp_sum(part);
  if ((tid & 31) == 0) s.redf[tid >> 5] = part;
  __syncthreads();
  if (tid < 32) {
    float t = (tid < NW) ? s.redf[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) s.bc[0] = t;
  }
  __syncthreads();
  const float inv_sum = 1.f / s.bc[0];
  const float p0 = e0 * inv_sum, p1 = e1 * inv_sum;
  s.scan[tid] = (double)p0 + (double)p1;
  __syncthreads();
  for (int off = 1; off < NT; off <<= 1) {                 // Hillis-Steele inclusive scan over pair sums
    double v = s.scan[tid];
    if (tid >= off) v += s.scan[tid - off];
    __syncthreads();
    s.scan[tid] = v;
    __syncthreads();
  }
  const double before = (tid == 0) ? 0.0 : s.scan[tid - 1];
  const float c0 = (float)(before + (double)p0);
  const float c1 = (float)(before + (double)p0 + (double)p1);
  const bool keep0 = (i0 == 0) || !(c0 > top_p);
  const bool keep1 = !(c1 > top_p);
  // second softmax over the kept logits / T (the normaliser is shared; argmax(p/q) is taken on p/q itself)
  const float tdiv = fmaxf(temperature, 1e-5f);
  const float mx2 = mx / tdiv;
  const float f0 = (keep0 && i0 < AR_CB_SIZE) ? expf(s.key[i0] / tdiv - mx2) : 0.f;
  const float f1 = (keep1 && i0 + 1 < AR_CB_SIZE) ? expf(s.key[i0 + 1] / tdiv - mx2) : 0.f;
  float part2 = warp_sum(f0 + f1);
  __syncthreads();
  if ((tid & 31) == 0) s.redf[tid >> 5] = part2;
  __syncthreads();
  if (tid < 32) {
    float t = (tid < NW) ? s.redf[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) s.bc[1] = t;
  }
  __syncthreads();
  const float inv2 = 1.f / s.bc[1];
  float best = -1.f;
  int besti = 0x7fffffff;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int i = i0 + u;
    const float f = u ? f1 : f0;
    const int orig = s.idx[i];
    if (orig < AR_CB_SIZE) {
      float q;
      if (noise) {
        q = __ldg(noise + orig);
      } else {
        const unsigned r = philox_round_mix(seed, step, slot, (unsigned)orig);
        q = -__logf(1.f - (r >> 8) * (1.f / 16777216.f) * 0.99999994f);
        q = fmaxf(q, 1e-30f);
      }
      const float val = (f * inv2) / q;
      if (val > best || (val == best && orig < besti)) { best = val; besti = orig; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
  }
  if ((tid & 31) == 0) { s.redf[tid >> 5] = best; s.redi[tid >> 5] = besti; }
  __syncthreads();
  if (tid < 32) {
    float b = (tid < NW) ? s.redf[tid] : -2.f;
    int bi = (tid < NW) ? s.redi[tid] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, b, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
    }
    if (tid == 0) s.redi[0] = bi;
  }
  __syncthreads();
  const int tok = s.redi[0];
  __syncthreads();
  return tok;
}


}  // namespace ardec
}  // namespace svanon
