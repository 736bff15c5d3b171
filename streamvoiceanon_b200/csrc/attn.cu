// Multi-token causal (optionally window-limited) attention in fp32, one warp per query row.
// Used by the content encoder's WindowLimitedTransformer (windowed_transformer.py:163-194, mask :291-303)
// and by the AR prompt prefill (dual_ar_stream.py:895-936 with the causal_mask rows of :333).
// Keys/values are staged through shared memory in tiles of 32 keys shared by the 8 queries of a CTA;
// softmax is the usual running max / running sum formulation in fp32.
#include <cstdlib>

#include "common.cuh"

namespace svanon {

namespace {

constexpr int QW = 8;        // queries (warps) per CTA
constexpr int KT = 32;       // keys per tile

__global__ void __launch_bounds__(QW * 32)
attention_kernel(const float* __restrict__ q, long long q_ld, const float* __restrict__ k, const float* __restrict__ v,
                 long long kv_head_stride, long long kv_row_stride, float* __restrict__ out, long long out_ld, int nq,
                 int qpos0, int window) {
  pdl_trigger();
  pdl_wait();
  __shared__ float Ks[KT][HEAD_DIM + 1];
  __shared__ float Vs[KT][HEAD_DIM];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  // independent streams side by side: segment `seg` owns query rows [seg*nq, (seg+1)*nq) and the same key rows
  const int bps = (nq + QW - 1) / QW;
  const int seg = blockIdx.x / bps;
  const int qi0 = (blockIdx.x - seg * bps) * QW;
  q += (long long)seg * nq * q_ld;
  k += (long long)seg * nq * kv_row_stride;
  v += (long long)seg * nq * kv_row_stride;
  out += (long long)seg * nq * out_ld;
  const int qi = qi0 + warp;
  const bool active = qi < nq;
  const int pos = qpos0 + qi;
  const int lo = max(0, pos - window + 1);

  float qr[HEAD_DIM];
  if (active) {
    const float* qp = q + (long long)qi * q_ld + h * HEAD_DIM;
#pragma unroll
    for (int d = 0; d < HEAD_DIM; ++d) qr[d] = __ldg(qp + d) * 0.125f;   // 1/sqrt(64)
  } else {
#pragma unroll
    for (int d = 0; d < HEAD_DIM; ++d) qr[d] = 0.f;
  }

  // key range needed by this CTA
  const int last_q = min(qi0 + QW, nq) - 1;
  const int k_hi = qpos0 + last_q;                         // inclusive
  const int k_lo = max(0, qpos0 + qi0 - window + 1);
  const float* kb = k + (long long)h * kv_head_stride;
  const float* vb = v + (long long)h * kv_head_stride;

  float m = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
  for (int kt = (k_lo / KT) * KT; kt <= k_hi; kt += KT) {
    __syncthreads();
    for (int i = threadIdx.x; i < KT * HEAD_DIM; i += QW * 32) {
      const int key = i / HEAD_DIM, d = i % HEAD_DIM;
      const int kp = kt + key;
      float kv = 0.f, vv = 0.f;
      if (kp <= k_hi) {
        kv = __ldg(kb + (long long)kp * kv_row_stride + d);
        vv = __ldg(vb + (long long)kp * kv_row_stride + d);
      }
      Ks[key][d] = kv;
      Vs[key][d] = vv;
    }
    __syncthreads();
    if (!active) continue;
    const int kp = kt + lane;
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < HEAD_DIM; ++d) s = fmaf(qr[d], Ks[lane][d], s);
    const bool valid = (kp >= lo) && (kp <= pos);
    s = valid ? s : -INFINITY;
    float tmax = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    if (tmax == -INFINITY) continue;                      // whole tile masked for this query (warp-uniform)
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);                   // m = -inf on the first tile -> 0
    const float p = valid ? expf(s - m_new) : 0.f;
    float psum = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
    l = l * corr + psum;
    acc0 *= corr;
    acc1 *= corr;
#pragma unroll
    for (int j = 0; j < KT; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      acc0 = fmaf(pj, Vs[j][lane], acc0);
      acc1 = fmaf(pj, Vs[j][lane + 32], acc1);
    }
    m = m_new;
  }
  if (active) {
    float* op = out + (long long)qi * out_ld + h * HEAD_DIM;
    const float inv = 1.f / l;
    op[lane] = acc0 * inv;
    op[lane + 32] = acc1 * inv;
  }
}

// Short sequences (all keys of a stream fit shared memory: <= 128 positions, the encoder's streaming window): the
// CTA stages K and V of its head ONCE (66 KB) and its 16 warps then run one query each without any further block
// barrier -- the tiled kernel above pays two barriers and one global round trip per 32 keys, which made it the single
// slowest kernel of the window encode (26.7 us for 17 MFLOP).
constexpr int SQW = 16;      // warps per CTA
constexpr int SMAXK = 128;   // keys held in shared memory
// QPW queries per warp: 1 spreads a single stream over many CTAs (latency); 8 makes one CTA own a whole (stream, head)
// so that K/V are staged once instead of once per 16 queries (many streams: 4.5x less staging traffic).
template <int QPW>
__global__ void __launch_bounds__(SQW * 32)
attention_short_kernel(const float* __restrict__ q, long long q_ld, const float* __restrict__ k, const float* __restrict__ v,
                       long long kv_head_stride, long long kv_row_stride, float* __restrict__ out, long long out_ld, int nq,
                       int qpos0, int window, float* __restrict__ out_lo) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float att_smem[];
  float (*Ks)[HEAD_DIM + 1] = reinterpret_cast<float (*)[HEAD_DIM + 1]>(att_smem);
  float (*Vs)[HEAD_DIM] = reinterpret_cast<float (*)[HEAD_DIM]>(att_smem + SMAXK * (HEAD_DIM + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  constexpr int QPC = SQW * QPW;                            // queries per CTA
  const int bps = (nq + QPC - 1) / QPC;
  const int seg = blockIdx.x / bps;
  const int qi0 = (blockIdx.x - seg * bps) * QPC;
  q += (long long)seg * nq * q_ld;
  k += (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  v += (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  out += (long long)seg * nq * out_ld;
  if (out_lo) out_lo += (long long)seg * nq * out_ld;      // lo term of the result, same layout (common.cuh: tf32_lo)
  const int k_hi = qpos0 + min(qi0 + QPC, nq) - 1;          // newest key any query of this CTA needs (inclusive)
  for (int i = threadIdx.x; i < (k_hi + 1) * (HEAD_DIM / 4); i += SQW * 32) {
    const int key = i / (HEAD_DIM / 4), c = (i % (HEAD_DIM / 4)) * 4;
    const float4 kv = __ldg(reinterpret_cast<const float4*>(k + (long long)key * kv_row_stride + c));
    const float4 vv = __ldg(reinterpret_cast<const float4*>(v + (long long)key * kv_row_stride + c));
    Ks[key][c] = kv.x; Ks[key][c + 1] = kv.y; Ks[key][c + 2] = kv.z; Ks[key][c + 3] = kv.w;
    *reinterpret_cast<float4*>(&Vs[key][c]) = vv;
  }
  __syncthreads();
  for (int jq = 0; jq < QPW; ++jq) {
  const int qi = qi0 + warp + SQW * jq;                     // interleaved: every warp gets early and late queries
  if (qi >= nq) break;
  const int pos = qpos0 + qi;
  const int lo = max(0, pos - window + 1);
  float qr[HEAD_DIM];
  {
    const float* qp = q + (long long)qi * q_ld + h * HEAD_DIM;
#pragma unroll
    for (int d = 0; d < HEAD_DIM; ++d) qr[d] = __ldg(qp + d) * 0.125f;   // 1/sqrt(64)
  }
  float m = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
  for (int kt = (lo / KT) * KT; kt <= pos; kt += KT) {
    const int kp = kt + lane;
    const bool valid = (kp >= lo) && (kp <= pos);
    float s = 0.f;
    if (kp <= k_hi) {
#pragma unroll
      for (int d = 0; d < HEAD_DIM; ++d) s = fmaf(qr[d], Ks[kp][d], s);
    }
    s = valid ? s : -INFINITY;
    float tmax = s;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
    if (tmax == -INFINITY) continue;
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);
    const float p = valid ? expf(s - m_new) : 0.f;
    float psum = p;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
    l = l * corr + psum;
    acc0 *= corr;
    acc1 *= corr;
    const int jn = min(KT, pos - kt + 1);
    for (int j = 0; j < jn; ++j) {
      const float pj = __shfl_sync(0xffffffffu, p, j);
      acc0 = fmaf(pj, Vs[kt + j][lane], acc0);
      acc1 = fmaf(pj, Vs[kt + j][lane + 32], acc1);
    }
    m = m_new;
  }
  float* op = out + (long long)qi * out_ld + h * HEAD_DIM;
  const float inv = 1.f / l;
  op[lane] = acc0 * inv;
  op[lane + 32] = acc1 * inv;
  if (out_lo) {
    float* lp = out_lo + (long long)qi * out_ld + h * HEAD_DIM;
    lp[lane] = tf32_lo(acc0 * inv);
    lp[lane + 32] = tf32_lo(acc1 * inv);
  }
  }
}

// Many streams, short sequences (<= 128 positions): TWO THREADS PER QUERY, each owning 32 of the head's 64 dimensions (every
// other 16-byte chunk, so that the pair's two addresses of a 128-bit shared-memory load are adjacent: one wavefront, no conflict).
// attention_short_kernel gives every key a lane and pays one shared-memory load per FMA in the q . k phase plus shuffle
// reductions per 32 keys (288 us per layer at 128 streams: 7.4 TFLOP/s).  Here the CTA of a (stream, head) stages K and V once
// and a lane pair keeps its query (2 x 32 registers) and output accumulator (2 x 32): a key or value row is read with broadcast
// 128-bit loads (two addresses per warp instruction) for 32 FMAs of each of the warp's 16 queries -- FMA-bound instead of
// shared-memory-bound -- the two half dot products meet in one shuffle, and max / sum / rescale are thread-local (flash-style
// over tiles of 16 keys).  Causal: warp w of 8 owns queries 16 w .. 16 w + 15 and walks exactly w + 1 key tiles; odd CTAs map
// their warps in reverse so that the long warps of co-resident CTAs do not share a scheduler.  (A first version with one thread
// per query needed 185 registers: 8 warps per SM, FMA pipe 20 % busy, 235 us -- profiles/r2zf_*.)
constexpr int RQ_KT = 16;
constexpr int RQ_THREADS = 2 * SMAXK;
__global__ void __launch_bounds__(RQ_THREADS, 2)
attention_rowq_kernel(const float* __restrict__ q, long long q_ld, const float* __restrict__ k, const float* __restrict__ v,
                      long long kv_head_stride, long long kv_row_stride, float* __restrict__ out, long long out_ld, int nq,
                      int qpos0, int window, float* __restrict__ out_lo) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float att_smem[];
  float (*Ks)[HEAD_DIM] = reinterpret_cast<float (*)[HEAD_DIM]>(att_smem);
  float (*Vs)[HEAD_DIM] = reinterpret_cast<float (*)[HEAD_DIM]>(att_smem + SMAXK * HEAD_DIM);
  constexpr int HD = HEAD_DIM / 2;
  const int h = blockIdx.y, seg = blockIdx.x;
  q += (long long)seg * nq * q_ld + h * HEAD_DIM;
  k += (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  v += (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  out += (long long)seg * nq * out_ld + h * HEAD_DIM;
  if (out_lo) out_lo += (long long)seg * nq * out_ld + h * HEAD_DIM;
  const int n_keys = qpos0 + nq;                              // keys 0 .. n_keys - 1 (<= SMAXK)
  for (int i = threadIdx.x; i < n_keys * (HEAD_DIM / 4); i += RQ_THREADS) {
    const int key = i / (HEAD_DIM / 4), c = (i % (HEAD_DIM / 4)) * 4;
    *reinterpret_cast<float4*>(&Ks[key][c]) = __ldg(reinterpret_cast<const float4*>(k + (long long)key * kv_row_stride + c));
    *reinterpret_cast<float4*>(&Vs[key][c]) = __ldg(reinterpret_cast<const float4*>(v + (long long)key * kv_row_stride + c));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = ((blockIdx.x + blockIdx.y) & 1) ? (RQ_THREADS / 32 - 1 - warp) : warp;
  const int qi = wq * 16 + (lane >> 1);
  const int d0 = (lane & 1) * 4;                              // this thread's dimensions: 8 i + d0 .. + 3, i = 0..7
  const bool q_ok = qi < nq;
  const int pos = qpos0 + (q_ok ? qi : 0);
  const int lo = max(0, pos - window + 1);
  float qr[HD], acc[HD];
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 t = q_ok ? __ldg(reinterpret_cast<const float4*>(q + (long long)qi * q_ld + d0 + 2 * d)) : make_float4(0.f, 0.f, 0.f, 0.f);
    qr[d] = t.x * 0.125f; qr[d + 1] = t.y * 0.125f; qr[d + 2] = t.z * 0.125f; qr[d + 3] = t.w * 0.125f;      // 1/sqrt(64)
    acc[d] = acc[d + 1] = acc[d + 2] = acc[d + 3] = 0.f;
  }
  float m = -INFINITY, l = 0.f;
  // the warp walks the key tiles any of its queries needs (warp-uniform bounds; every lane takes part in the shuffles)
  const int w_hi = min(n_keys - 1, qpos0 + min(wq * 16 + 15, nq - 1));
  const int w_lo = max(0, qpos0 + wq * 16 - window + 1);
  for (int kt = (w_lo / RQ_KT) * RQ_KT; kt <= w_hi; kt += RQ_KT) {
    float sc[RQ_KT];
    float tmax = -INFINITY;
#pragma unroll
    for (int j = 0; j < RQ_KT; ++j) {
      const int kp = kt + j;
      float s0 = 0.f, s1 = 0.f;
      if (kp < n_keys) {                                      // warp-uniform
#pragma unroll
        for (int d = 0; d < HD; d += 8) {
          const float4 a = *reinterpret_cast<const float4*>(&Ks[kp][d0 + 2 * d]);
          const float4 b = *reinterpret_cast<const float4*>(&Ks[kp][d0 + 2 * d + 8]);
          s0 = fmaf(qr[d], a.x, s0); s0 = fmaf(qr[d + 1], a.y, s0); s0 = fmaf(qr[d + 2], a.z, s0); s0 = fmaf(qr[d + 3], a.w, s0);
          s1 = fmaf(qr[d + 4], b.x, s1); s1 = fmaf(qr[d + 5], b.y, s1); s1 = fmaf(qr[d + 6], b.z, s1); s1 = fmaf(qr[d + 7], b.w, s1);
        }
      }
      float sp = s0 + s1;
      // both lanes of the pair end up with the same bits: (low half) + (high half)
      const float other = __shfl_xor_sync(0xffffffffu, sp, 1);
      sp = (lane & 1) ? other + sp : sp + other;
      const bool valid = q_ok && kp >= lo && kp <= pos;
      sc[j] = valid ? sp : -INFINITY;
      tmax = fmaxf(tmax, sc[j]);
    }
    if (tmax == -INFINITY) continue;                          // pair-uniform; no shuffles below
    const float m_new = fmaxf(m, tmax);
    const float corr = expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] *= corr;
#pragma unroll
    for (int j = 0; j < RQ_KT; ++j) {
      const int kp = kt + j;
      const float pj = expf(sc[j] - m_new);                   // 0 for masked keys
      l += pj;
      if (kp < n_keys) {
#pragma unroll
        for (int d = 0; d < HD; d += 4) {
          const float4 vv = *reinterpret_cast<const float4*>(&Vs[kp][d0 + 2 * d]);
          acc[d] = fmaf(pj, vv.x, acc[d]); acc[d + 1] = fmaf(pj, vv.y, acc[d + 1]);
          acc[d + 2] = fmaf(pj, vv.z, acc[d + 2]); acc[d + 3] = fmaf(pj, vv.w, acc[d + 3]);
        }
      }
    }
    m = m_new;
  }
  if (!q_ok) return;
  const float inv = 1.f / l;
  float* op = out + (long long)qi * out_ld + d0;
  float* lp = out_lo ? out_lo + (long long)qi * out_ld + d0 : nullptr;
#pragma unroll
  for (int d = 0; d < HD; d += 4) {
    const float4 o = make_float4(acc[d] * inv, acc[d + 1] * inv, acc[d + 2] * inv, acc[d + 3] * inv);
    *reinterpret_cast<float4*>(op + 2 * d) = o;
    if (lp) *reinterpret_cast<float4*>(lp + 2 * d) = make_float4(tf32_lo(o.x), tf32_lo(o.y), tf32_lo(o.z), tf32_lo(o.w));
  }
}

// The last c queries of each segment only; thread t owns key t of the segment (nq <= 128).
__global__ void __launch_bounds__(ATT_TAIL_MAX_KEYS)
attention_tail_kernel(const float* __restrict__ q, long long q_ld, const float* __restrict__ k, const float* __restrict__ v,
                      long long kv_head_stride, long long kv_row_stride, float* __restrict__ out, long long out_ld, int nq,
                      int c, int window) {
  pdl_trigger();
  pdl_wait();
  __shared__ float qs[HEAD_DIM];
  __shared__ float ps[ATT_TAIL_MAX_KEYS];
  __shared__ float red[2][ATT_TAIL_MAX_KEYS / 32];
  __shared__ float osum[2][HEAD_DIM];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int h = blockIdx.y;
  const int seg = blockIdx.x / c, j = blockIdx.x - seg * c;
  const int pos = nq - c + j;
  const int lo = max(0, pos - window + 1);
  const float* kb = k + (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  const float* vb = v + (long long)seg * nq * kv_row_stride + (long long)h * kv_head_stride;
  if (t < HEAD_DIM) qs[t] = __ldg(q + ((long long)seg * nq + pos) * q_ld + h * HEAD_DIM + t) * 0.125f;   // 1/sqrt(64)
  __syncthreads();
  const bool valid = t >= lo && t <= pos;
  float s = -INFINITY;
  if (valid) {
    const float4* kr = reinterpret_cast<const float4*>(kb + (long long)t * kv_row_stride);
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < HEAD_DIM / 4; ++d) {
      const float4 kv = __ldg(kr + d);
      a = fmaf(qs[4 * d], kv.x, a);
      a = fmaf(qs[4 * d + 1], kv.y, a);
      a = fmaf(qs[4 * d + 2], kv.z, a);
      a = fmaf(qs[4 * d + 3], kv.w, a);
    }
    s = a;
  }
  float m = s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[0][warp] = m;
  __syncthreads();
  m = red[0][0];
#pragma unroll
  for (int w = 1; w < ATT_TAIL_MAX_KEYS / 32; ++w) m = fmaxf(m, red[0][w]);
  const float p = valid ? expf(s - m) : 0.f;
  ps[t] = p;
  float l = p;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  if (lane == 0) red[1][warp] = l;
  __syncthreads();
  l = red[1][0];
#pragma unroll
  for (int w = 1; w < ATT_TAIL_MAX_KEYS / 32; ++w) l += red[1][w];
  // P.V: thread = (half of the keys, output dim)
  const int d = t & (HEAD_DIM - 1), half = t >> 6;
  const int k0 = max(lo, half * (ATT_TAIL_MAX_KEYS / 2)), k1 = min(pos, half * (ATT_TAIL_MAX_KEYS / 2) + ATT_TAIL_MAX_KEYS / 2 - 1);
  float acc = 0.f;
  for (int key = k0; key <= k1; ++key) acc = fmaf(ps[key], __ldg(vb + (long long)key * kv_row_stride + d), acc);
  osum[half][d] = acc;
  __syncthreads();
  if (t < HEAD_DIM) out[((long long)seg * c + j) * out_ld + h * HEAD_DIM + t] = (osum[0][t] + osum[1][t]) / l;
}

__global__ void kv_append_kernel(const float* __restrict__ qkv, int heads, float* __restrict__ kc, float* __restrict__ vc,
                                 int max_seq, int pos0) {
  pdl_trigger();
  pdl_wait();
  const int row = blockIdx.x;
  const int D = heads * HEAD_DIM;
  const float* kr = qkv + (long long)row * 3 * D + D;
  const float* vr = kr + D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    const int h = c / HEAD_DIM, d = c % HEAD_DIM;
    const long long dst = ((long long)h * max_seq + pos0 + row) * HEAD_DIM + d;
    kc[dst] = kr[c];
    vc[dst] = vr[c];
  }
}

}  // namespace

void launch_attention(const float* q, long long q_ld, const float* k, const float* v, long long kv_head_stride,
                      long long kv_row_stride, float* out, long long out_ld, int nq, int qpos0, int heads, int window,
                      cudaStream_t st, int nseg, float* out_lo) {
  if (nq <= 0 || nseg <= 0) return;
  if (qpos0 + nq <= SMAXK && (kv_row_stride % 4) == 0) {        // the whole key range of a stream fits shared memory
    constexpr size_t SMEM = (size_t)SMAXK * (2 * HEAD_DIM + 1) * sizeof(float);
    static bool configured = false;
    if (!configured) {
      SV_CUDA(cudaFuncSetAttribute(attention_short_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
      SV_CUDA(cudaFuncSetAttribute(attention_short_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
      configured = true;
    }
    static const bool rowq = [] {
      const char* e = getenv("SVANON_ATTN_ROWQ");          // 0: many streams keep the lane-per-key kernel (A/B)
      return !e || atoi(e) != 0;
    }();
    const bool al16 = ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
                        reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(out_lo)) & 15) == 0 &&
                      q_ld % 4 == 0 && out_ld % 4 == 0 && kv_head_stride % 4 == 0;
    if (nseg >= 8 && rowq && al16) {
      constexpr size_t SMEM_RQ = (size_t)SMAXK * 2 * HEAD_DIM * sizeof(float);
      static bool configured_rq = false;
      if (!configured_rq) {
        SV_CUDA(cudaFuncSetAttribute(attention_rowq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_RQ));
        configured_rq = true;
      }
      launch_pdl(attention_rowq_kernel, dim3(nseg, heads), dim3(RQ_THREADS), SMEM_RQ, st, q, q_ld, k, v, kv_head_stride, kv_row_stride, out,
                 out_ld, nq, qpos0, window, out_lo);
    } else if (nseg >= 8)
      launch_pdl(attention_short_kernel<8>, dim3((nq + SQW * 8 - 1) / (SQW * 8) * nseg, heads), dim3(SQW * 32), SMEM, st, q, q_ld,
                 k, v, kv_head_stride, kv_row_stride, out, out_ld, nq, qpos0, window, out_lo);
    else
      launch_pdl(attention_short_kernel<1>, dim3((nq + SQW - 1) / SQW * nseg, heads), dim3(SQW * 32), SMEM, st, q, q_ld, k, v,
                 kv_head_stride, kv_row_stride, out, out_ld, nq, qpos0, window, out_lo);
    SV_LAUNCHED();
    return;
  }
  SV_CHECK(out_lo == nullptr, "attention: the lo output exists on the short-sequence kernel only");
  dim3 grid((nq + QW - 1) / QW * nseg, heads);
  launch_pdl(attention_kernel, dim3(grid), dim3(QW * 32), 0, st, q, q_ld, k, v, kv_head_stride, kv_row_stride, out, out_ld, nq, qpos0,
                                             window);
  SV_LAUNCHED();
}

void launch_attention_tail(const float* q, long long q_ld, const float* k, const float* v, long long kv_head_stride,
                           long long kv_row_stride, float* out, long long out_ld, int nq, int c, int heads, int window,
                           cudaStream_t st, int nseg) {
  SV_CHECK(nq >= 1 && nq <= ATT_TAIL_MAX_KEYS && c >= 1 && c <= nq && (kv_row_stride % 4) == 0, "attention_tail: unsupported shape");
  launch_pdl(attention_tail_kernel, dim3(nseg * c, heads), dim3(ATT_TAIL_MAX_KEYS), 0, st, q, q_ld, k, v, kv_head_stride,
             kv_row_stride, out, out_ld, nq, c, window);
  SV_LAUNCHED();
}

void launch_kv_append(const float* qkv, int rows, int heads, float* kc, float* vc, int max_seq, int pos0,
                      cudaStream_t st) {
  if (rows <= 0) return;
  launch_pdl(kv_append_kernel, dim3(rows), dim3(256), 0, st, qkv, heads, kc, vc, max_seq, pos0);
  SV_LAUNCHED();
}

}  // namespace svanon
