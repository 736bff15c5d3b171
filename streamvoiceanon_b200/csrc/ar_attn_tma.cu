// Stage A, many streams: slow-stack attention of one (stream, head) as ONE warp-specialised kernel that fuses everything
// Attention.forward does around SDPA (modules/dual_ar_stream.py:895-936) -- RoPE on q and k of the frame's two tokens,
// KV-cache append (KVCache.update, :132-150), attention of both tokens over the stream's valid cache prefix -- and
// streams the cache through shared memory with TMA bulk copies:
//
//   warp 0 (producer)   one elected lane walks the cache prefix [0, pos) in tiles of 128 keys: first the K tiles, then
//                       the V tiles, each ONE contiguous `cp.async.bulk` (the cache is [layer][head][max_seq][64], so a
//                       head's keys are adjacent rows) into a 4-stage ring, completing on the stage's `full` mbarrier;
//                       it runs up to four tiles (128 KB) ahead of the math, bounded by the `empty` mbarriers.
//   warps 1-4 (consumers)
//       prologue        RoPE of the two new tokens' q and k, append of k / v to the cache (plain stores; the two new
//                       keys are used from shared memory, never re-read through the async proxy)
//       pass 1 (K)      thread = key: each lane dots its key row with both queries, reading the row and the queries in
//                       a per-lane ROTATED order of 16-byte chunks, which makes the row-per-lane access conflict-free;
//                       scores (x 1/8) go to a shared score strip [2][max_seq]
//       softmax         block max, exp and sum over the strip (fixed order), probabilities left unnormalised
//       pass 2 (V)      thread = two output dims: each warp folds its 32 keys of the tile into a [2][64] partial with
//                       the probabilities broadcast from the strip; partials are added in warp order and scaled by 1/sum
//
// Why two passes instead of an online softmax: with <= 2048 keys the score strip is 16 KB, K and V are read exactly once
// either way, and the accumulators never need rescaling; the producer keeps streaming V tiles while the softmax runs.
// Why CUDA cores: two query rows per head (M = 2 of a 64- or 128-row MMA tile) and 0.5 FMA per byte of K/V -- the kernel
// is bound by how fast the cache leaves HBM; profiles/ holds the ncu capture (DRAM throughput vs tensor-pipe idle).
//
// Template parameter KV: float (parity build: fp32 cache) or __half (perf mode: the reference's own GPU cache type,
// evaluations/infer_arvc.py:55-59), rows of 64 elements either way.
#include <cuda_fp16.h>

#include "ar_decode_common.cuh"
#include "engine.hpp"

namespace svanon {

using namespace ardec;

namespace {

constexpr int A2_TILE = 128;                 // keys per tile
constexpr int A2_STAGES = 4;
constexpr int A2_CONSUMERS = 4;              // consumer warps: 32 keys of a tile each
constexpr int A2_THREADS = (A2_CONSUMERS + 1) * 32;
constexpr int A2_STRIP = AR_MAX_SEQ + 8;     // score strip per query

__device__ __forceinline__ unsigned a2_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void a2_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void a2_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void a2_mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a2_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "A2_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni A2_WAIT_DONE;\n"
      "bra.uni A2_WAIT_LOOP;\n"
      "A2_WAIT_DONE:\n"
      "}\n" ::"r"(a2_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void a2_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(a2_smem_u32(dst)), "l"(src), "r"(bytes), "r"(a2_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void a2_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(A2_CONSUMERS * 32) : "memory"); }

// 4 consecutive elements of a cache row as floats
__device__ __forceinline__ float4 a2_ld4(const float* row, int chunk4) { return *reinterpret_cast<const float4*>(row + 4 * chunk4); }
__device__ __forceinline__ float4 a2_ld4(const __half* row, int chunk4) {
  const uint2 raw = *reinterpret_cast<const uint2*>(row + 4 * chunk4);
  const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
  const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ float2 a2_ld2(const float* row, int pair) { return *reinterpret_cast<const float2*>(row + 2 * pair); }
__device__ __forceinline__ float2 a2_ld2(const __half* row, int pair) {
  return __half22float2(*reinterpret_cast<const __half2*>(row + 2 * pair));
}
__device__ __forceinline__ void a2_st2(float* row, int pair, float2 v) { *reinterpret_cast<float2*>(row + 2 * pair) = v; }
__device__ __forceinline__ void a2_st2(__half* row, int pair, float2 v) {
  *reinterpret_cast<__half2*>(row + 2 * pair) = __floats2half2_rn(v.x, v.y);
}

template <typename KV>
struct A2Smem {
  KV tiles[A2_STAGES][A2_TILE * HEAD_DIM];
  float sc[2][A2_STRIP];
  float q[2][HEAD_DIM];
  float knew[2][HEAD_DIM];
  float vnew[2][HEAD_DIM];
  float part[A2_CONSUMERS][2][HEAD_DIM];
  float red[2][A2_CONSUMERS];
  unsigned long long full[A2_STAGES], empty[A2_STAGES];
};

template <typename KV>
__global__ void __launch_bounds__(A2_THREADS, 1) arb_attn_slow_tma_kernel(const ArBatchSlot* __restrict__ slots,
                                                                          const float* __restrict__ qkv, float* __restrict__ y,
                                                                          const float* __restrict__ rope, int layer, int max_seq) {
  extern __shared__ unsigned char a2_raw[];
  A2Smem<KV>& sm = *reinterpret_cast<A2Smem<KV>*>((reinterpret_cast<uintptr_t>(a2_raw) + 127) & ~(uintptr_t)127);
  const int h = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < A2_STAGES; ++s) { a2_mbar_init(&sm.full[s], 1); a2_mbar_init(&sm.empty[s], A2_CONSUMERS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  pdl_wait();                              // qkv comes from the GEMM in front; the cache from earlier frames
  __syncthreads();
  const ArBatchSlot& s = slots[b];
  const int pos = s.pos;                   // keys [0, pos) are in the cache; this frame adds pos and pos + 1
  const int nT = (pos + A2_TILE - 1) / A2_TILE;
  KV* kc = reinterpret_cast<KV*>(s.kc) + ((long long)layer * H + h) * max_seq * HEAD_DIM;
  KV* vc = reinterpret_cast<KV*>(s.vc) + ((long long)layer * H + h) * max_seq * HEAD_DIM;

  if (warp == 0) {
    // ================================================================= producer
    if (lane == 0) {
      for (int t = 0; t < 2 * nT; ++t) {
        const int st = t % A2_STAGES;
        a2_mbar_wait(&sm.empty[st], ((t / A2_STAGES) & 1) ^ 1);      // first round passes on the fresh barrier
        const int tt = t < nT ? t : t - nT;
        const int rows = min(A2_TILE, pos - tt * A2_TILE);
        const unsigned bytes = (unsigned)rows * HEAD_DIM * sizeof(KV);
        a2_mbar_expect_tx(&sm.full[st], bytes);
        a2_bulk_g2s(sm.tiles[st], (t < nT ? kc : vc) + (long long)tt * A2_TILE * HEAD_DIM, bytes, &sm.full[st]);
      }
    }
    return;
  }
  // =================================================================== consumers
  const int cw = warp - 1;
  if (cw < 2) {
    // RoPE (interleaved pairs, lane owns (2 lane, 2 lane + 1)) + cache append of token cw
    const int j = cw;
    const float* row = qkv + (long long)(2 * b + j) * 3 * D + h * HEAD_DIM;
    const float2 cs = __ldg(reinterpret_cast<const float2*>(rope + ((long long)(pos + j) * (HEAD_DIM / 2) + lane) * 2));
    const float2 qv = *(reinterpret_cast<const float2*>(row) + lane);
    const float2 kv = *(reinterpret_cast<const float2*>(row + D) + lane);
    const float2 vv = *(reinterpret_cast<const float2*>(row + 2 * D) + lane);
    const float2 qr = make_float2(qv.x * cs.x - qv.y * cs.y, qv.y * cs.x + qv.x * cs.y);
    float2 kr = make_float2(kv.x * cs.x - kv.y * cs.y, kv.y * cs.x + kv.x * cs.y);
    float2 vr = vv;
    a2_st2(kc + (long long)(pos + j) * HEAD_DIM, lane, kr);
    a2_st2(vc + (long long)(pos + j) * HEAD_DIM, lane, vr);
    if (sizeof(KV) == 2) {                 // what later frames will read back from the cache is what this frame uses
      kr = __half22float2(__floats2half2_rn(kr.x, kr.y));
      vr = __half22float2(__floats2half2_rn(vr.x, vr.y));
    }
    sm.q[j][2 * lane] = qr.x; sm.q[j][2 * lane + 1] = qr.y;
    sm.knew[j][2 * lane] = kr.x; sm.knew[j][2 * lane + 1] = kr.y;
    sm.vnew[j][2 * lane] = vr.x; sm.vnew[j][2 * lane + 1] = vr.y;
  }
  a2_consumer_sync();
  // ---- pass 1: scores of the cached keys, thread = key
  for (int t = 0; t < nT; ++t) {
    const int st = t % A2_STAGES;
    a2_mbar_wait(&sm.full[st], (t / A2_STAGES) & 1);
    const int key = cw * 32 + lane;
    const int g = t * A2_TILE + key;
    if (g < pos) {
      const KV* krow = sm.tiles[st] + key * HEAD_DIM;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int i = 0; i < HEAD_DIM / 4; ++i) {
        const int c = (i + lane) & (HEAD_DIM / 4 - 1);           // rotated chunk order: conflict-free row-per-lane reads
        const float4 k4 = a2_ld4(krow, c);
        const float4 q0 = *reinterpret_cast<const float4*>(&sm.q[0][4 * c]);
        const float4 q1 = *reinterpret_cast<const float4*>(&sm.q[1][4 * c]);
        s0 = fmaf(q0.x, k4.x, s0); s0 = fmaf(q0.y, k4.y, s0); s0 = fmaf(q0.z, k4.z, s0); s0 = fmaf(q0.w, k4.w, s0);
        s1 = fmaf(q1.x, k4.x, s1); s1 = fmaf(q1.y, k4.y, s1); s1 = fmaf(q1.z, k4.z, s1); s1 = fmaf(q1.w, k4.w, s1);
      }
      sm.sc[0][g] = s0 * 0.125f;
      sm.sc[1][g] = s1 * 0.125f;
    }
    __syncwarp();
    if (lane == 0) a2_mbar_arrive(&sm.empty[st]);
  }
  // the frame's own two keys: token 0 sees key pos, token 1 sees pos and pos + 1
  if (cw == 0) {
    const float2 q0 = a2_ld2(sm.q[0], lane), q1 = a2_ld2(sm.q[1], lane);
    const float2 k0 = a2_ld2(sm.knew[0], lane), k1 = a2_ld2(sm.knew[1], lane);
    const float d00 = warp_sum(q0.x * k0.x + q0.y * k0.y) * 0.125f;
    const float d10 = warp_sum(q1.x * k0.x + q1.y * k0.y) * 0.125f;
    const float d11 = warp_sum(q1.x * k1.x + q1.y * k1.y) * 0.125f;
    if (lane == 0) { sm.sc[0][pos] = d00; sm.sc[1][pos] = d10; sm.sc[1][pos + 1] = d11; }
  }
  a2_consumer_sync();
  // ---- softmax over the strip (n0 = pos + 1 entries for token 0, pos + 2 for token 1), probabilities unnormalised
  const int ct = tid - 32;                 // 0 .. 127 among the consumers
  float inv[2];
#pragma unroll
  for (int qi = 0; qi < 2; ++qi) {
    const int n = pos + 1 + qi;
    float m = -INFINITY;
    for (int i = ct; i < n; i += A2_CONSUMERS * 32) m = fmaxf(m, sm.sc[qi][i]);
    m = warp_max(m);
    if (lane == 0) sm.red[qi][cw] = m;
    a2_consumer_sync();
    m = sm.red[qi][0];
#pragma unroll
    for (int w = 1; w < A2_CONSUMERS; ++w) m = fmaxf(m, sm.red[qi][w]);
    a2_consumer_sync();
    float l = 0.f;
    for (int i = ct; i < n; i += A2_CONSUMERS * 32) {
      const float e = expf(sm.sc[qi][i] - m);
      sm.sc[qi][i] = e;
      l += e;
    }
    l = warp_sum(l);
    if (lane == 0) sm.red[qi][cw] = l;
    a2_consumer_sync();
    float tot = sm.red[qi][0];
#pragma unroll
    for (int w = 1; w < A2_CONSUMERS; ++w) tot += sm.red[qi][w];
    inv[qi] = 1.f / tot;
    a2_consumer_sync();
  }
  // ---- pass 2: P V, thread = output dims (2 lane, 2 lane + 1), warp = 32 keys of the tile
  float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
  for (int t = nT; t < 2 * nT; ++t) {
    const int st = t % A2_STAGES;
    a2_mbar_wait(&sm.full[st], (t / A2_STAGES) & 1);
    const int g0 = (t - nT) * A2_TILE + cw * 32;
    const int cnt = min(32, pos - g0);
    const KV* vbase = sm.tiles[st] + (cw * 32) * HEAD_DIM;
#pragma unroll 4
    for (int k = 0; k < cnt; ++k) {
      const float2 v = a2_ld2(vbase + k * HEAD_DIM, lane);
      const float p0 = sm.sc[0][g0 + k], p1 = sm.sc[1][g0 + k];
      a0.x = fmaf(p0, v.x, a0.x); a0.y = fmaf(p0, v.y, a0.y);
      a1.x = fmaf(p1, v.x, a1.x); a1.y = fmaf(p1, v.y, a1.y);
    }
    __syncwarp();
    if (lane == 0) a2_mbar_arrive(&sm.empty[st]);
  }
  if (cw == 0) {
    const float2 v0 = a2_ld2(sm.vnew[0], lane), v1 = a2_ld2(sm.vnew[1], lane);
    const float p00 = sm.sc[0][pos], p10 = sm.sc[1][pos], p11 = sm.sc[1][pos + 1];
    a0.x = fmaf(p00, v0.x, a0.x); a0.y = fmaf(p00, v0.y, a0.y);
    a1.x = fmaf(p10, v0.x, a1.x); a1.y = fmaf(p10, v0.y, a1.y);
    a1.x = fmaf(p11, v1.x, a1.x); a1.y = fmaf(p11, v1.y, a1.y);
  }
  a2_st2(sm.part[cw][0], lane, a0);
  a2_st2(sm.part[cw][1], lane, a1);
  a2_consumer_sync();
  if (cw < 2) {
    float2 r = a2_ld2(sm.part[0][cw], lane);
#pragma unroll
    for (int w = 1; w < A2_CONSUMERS; ++w) {
      const float2 o = a2_ld2(sm.part[w][cw], lane);
      r.x += o.x; r.y += o.y;
    }
    const float scale = cw == 0 ? inv[0] : inv[1];
    r.x *= scale; r.y *= scale;
    *(reinterpret_cast<float2*>(y + (long long)(2 * b + cw) * D + h * HEAD_DIM) + lane) = r;
  }
}

template <typename KV>
void launch_tma(const ArBatchSlot* slots, const float* qkv, float* y, const float* rope, int layer, int max_seq, int B,
                cudaStream_t st) {
  constexpr size_t SMEM = sizeof(A2Smem<KV>) + 128;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(arb_attn_slow_tma_kernel<KV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  launch_pdl(arb_attn_slow_tma_kernel<KV>, dim3(AR_HEADS, B), dim3(A2_THREADS), SMEM, st, slots, qkv, y, rope, layer, max_seq);
}

}  // namespace

// kv_half: the stream's slow cache holds __half rows (perf mode) instead of float rows
void launch_arb_attn_slow_tma(const ArBatchSlot* slots, const float* qkv, float* y, const float* rope, int layer, int max_seq,
                              int B, bool kv_half, cudaStream_t st) {
  SV_CHECK(max_seq <= AR_MAX_SEQ, "max_seq exceeds the attention kernel's score strip");
  if (kv_half) launch_tma<__half>(slots, qkv, y, rope, layer, max_seq, B, st);
  else launch_tma<float>(slots, qkv, y, rope, layer, max_seq, B, st);
}

}  // namespace svanon
