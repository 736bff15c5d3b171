// Multi-token causal (optionally window-limited) attention in fp32, one warp per query row.
// Used by the content encoder's WindowLimitedTransformer (windowed_transformer.py:163-194, mask :291-303)
// and by the AR prompt prefill (dual_ar_stream.py:895-936 with the causal_mask rows of :333).
// Keys/values are staged through shared memory in tiles of 32 keys shared by the 8 queries of a CTA;
// softmax is the usual running max / running sum formulation in fp32.
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
    if (nseg >= 8)
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
