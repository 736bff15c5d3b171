// Direct causal conv for the two thinnest HiFi-GAN levels (C = 32 and C = 16 channels, ResBlock1 convs of
// firefly.py:149-219): out[m][co] = bias[co] + sum_tap sum_ci W[tap][co][ci] * silu(x[m + off_tap][ci]) (+ residual).
//
// (Input rows are staged with a pitch of C + 4 floats: the 4 / 8 rows a warp reads in one instruction then fall in different
// banks -- with pitch C every input load was a 4-way / 8-way conflict; V 6.5 -> 5.8 ms at 128 streams, profiles/r2zi_*.)
// As GEMMs these problems have N = K-per-tap = 16 or 32: far too thin for the tensor cores and latency-bound in the
// generic cp.async pipeline (11 taps = 11 dependent pipeline steps, ~24 us per launch for 35 MFLOP).  Here a CTA owns
// TR consecutive output rows of one problem: it stages the SiLU'd input rows (with their causal halo) and the whole
// weight tensor (<= 45 KB) in shared memory once, then every thread produces 4 output channels of one row from
// registers.  Same GemmParams contract as gemm.cu (up to 3 problems per launch, streams side by side).
#include <algorithm>
#include <cstdlib>

#include "common.cuh"

namespace svanon {

namespace {

struct ConvBatch {
  GemmParams p[3];
};

__device__ __forceinline__ float silu_acc(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// C channels in and out, TR output rows per CTA.  A thread owns RB = 2 rows x 4 output channels {cq, cq+Q, cq+2Q, cq+3Q}
// (strided, so that the Q threads of a row read CONSECUTIVE float4 weights: conflict-free); blockDim = TR/2 * C/4.
template <int C, int TR>
__global__ void __launch_bounds__(TR * C / 8) conv_small_kernel(const ConvBatch batch) {
  extern __shared__ __align__(16) float smem[];
  constexpr int Q = C / 4;                       // channel quads
  const GemmParams& p = batch.p[blockIdx.z];
  const int ktaps = p.taps > 1 ? p.taps : p.K / C;          // dilation 1: one "tap" of k*C overlapping columns
  const int halo = -p.tap_off[0];                           // the oldest row any tap reaches
  float* w_s = smem;                                        // [ktaps][Q (ci quad)][C (co)][4 (ci)]
  constexpr int XP = C + 4;                                 // row pitch of x_s: the 4 (C = 32) / 8 (C = 16) rows a warp reads at once
                                                            // must not share banks (pitch C: 4-way / 8-way conflicts on every load)
  float* x_s = smem + ktaps * C * C;                        // [halo + TR][XP], SiLU applied
  const int tid = threadIdx.x;
  pdl_trigger();
  // weights: global layout dilation 1: W[co][tap*C + ci]; dilated: W[(tap*C + co)*C + ci]
  for (int i = tid; i < ktaps * C * Q; i += blockDim.x) {
    const int q = i % Q, co = (i / Q) % C, tap = i / (Q * C);
    const float* src = p.taps > 1 ? p.W + ((long long)tap * C + co) * C + q * 4 : p.W + (long long)co * p.K + tap * C + q * 4;
    *reinterpret_cast<float4*>(w_s + ((tap * Q + q) * C + co) * 4) = __ldg(reinterpret_cast<const float4*>(src));
  }
  pdl_wait();
  const int m0 = blockIdx.x * TR;                            // TR divides the rows of a stream: a tile never straddles two
  const float* a0 = p.A + gemm_a_row(p, m0);
  for (int i = tid; i < (halo + TR) * Q; i += blockDim.x) {
    const int q = i % Q, r = i / Q;
    float4 v = *reinterpret_cast<const float4*>(a0 + (long long)(r - halo) * p.lda + q * 4);
    v.x = silu_acc(v.x); v.y = silu_acc(v.y); v.z = silu_acc(v.z); v.w = silu_acc(v.w);
    *reinterpret_cast<float4*>(x_s + r * XP + q * 4) = v;
  }
  __syncthreads();
  const int cq = tid % Q, rp = tid / Q;                      // rows m0 + 2*rp, +1; output channels cq + Q*j
  float acc[2][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[0][j] = acc[1][j] = p.bias ? __ldg(p.bias + cq + Q * j) : 0.f;
  const int first_off = p.taps > 1 ? 0 : p.tap_off[0];
  for (int tap = 0; tap < ktaps; ++tap) {
    const int off = p.taps > 1 ? p.tap_off[tap] : first_off + tap;
    const float* xr = x_s + (halo + 2 * rp + off) * XP;
    const float* wt = w_s + tap * Q * C * 4 + cq * 4;
#pragma unroll
    for (int q = 0; q < Q; ++q) {
      const float4 x0 = *reinterpret_cast<const float4*>(xr + q * 4);
      const float4 x1 = *reinterpret_cast<const float4*>(xr + XP + q * 4);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(wt + (q * C + Q * j) * 4);
        acc[0][j] = fmaf(x0.x, wv.x, acc[0][j]); acc[1][j] = fmaf(x1.x, wv.x, acc[1][j]);
        acc[0][j] = fmaf(x0.y, wv.y, acc[0][j]); acc[1][j] = fmaf(x1.y, wv.y, acc[1][j]);
        acc[0][j] = fmaf(x0.z, wv.z, acc[0][j]); acc[1][j] = fmaf(x1.z, wv.z, acc[1][j]);
        acc[0][j] = fmaf(x0.w, wv.w, acc[0][j]); acc[1][j] = fmaf(x1.w, wv.w, acc[1][j]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int m = m0 + 2 * rp + r;
    if (m >= p.M) continue;
    const float* res = p.residual ? p.residual + gemm_r_row(p, m) : nullptr;
    float* dst = p.C + gemm_c_row(p, m);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = cq + Q * j;
      dst[co] = acc[r][j] + (res ? res[co] : 0.f);
    }
  }
}

template <int C, int TR>
void launch_cfg(const ConvBatch& b, int count, size_t smem, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(conv_small_kernel<C, TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    configured = true;
  }
  launch_pdl(conv_small_kernel<C, TR>, dim3((b.p[0].M + TR - 1) / TR, 1, count), dim3(TR * C / 8), smem, st, b);
}

}  // namespace

// Returns false when the problem is not one of the thin causal convs this kernel is for.
bool launch_conv_small(const GemmParams* ps, int count, cudaStream_t st) {
  const GemmParams& p0 = ps[0];
  const int C = p0.N;
  if (C != 16 && C != 32) return false;
  static const bool big_tiles = [] {
    const char* e = getenv("SVANON_CONV_SMALL_BIG");        // 0: 64 / 128 rows per CTA also at many streams (A/B).  Default: twice the
    return !e || atoi(e) != 0;                              // rows from M = 8192 on -- the weight tensor is staged half as often
  }();
  const int TR = (C == 16 ? 128 : 64) * ((big_tiles && p0.M >= 8192) ? 2 : 1);
  static const long long max_m = [] {
    const char* e = getenv("SVANON_CONV_SMALL_MAX_M");      // tuning knob: largest M that takes this kernel (above it the
    return e ? atoll(e) : (1LL << 40);                       // tensor-core kernel's thin 128 x N tile takes over: measured slower)
  }();
  if (p0.M > max_m) return false;
  ConvBatch b;
  size_t smem = 0;
  for (int i = 0; i < count; ++i) {
    const GemmParams& p = ps[i];
    const int kt = p.taps > 1 ? p.taps : p.K / C;
    if (p.N != C || p.lda != C || p.a_row_step != 1 || p.prologue != PRO_SILU || p.act != ACT_NONE || p.gamma || p.accumulate ||
        p.out_scale != 1.f || (p.taps > 1 ? p.K != C : p.K % C != 0) || kt < 1 || kt > MAX_TAPS || p.M != p0.M || p.M % TR != 0 ||
        (p.seg_rows > 0 && p.seg_rows % TR != 0))
      return false;
    const int halo = -p.tap_off[0];
    if (halo < 0) return false;
    for (int t = 0; t < (p.taps > 1 ? p.taps : 1); ++t)
      if (p.tap_off[t] > 0 || p.tap_off[t] < -halo) return false;
    if (p.taps == 1 && p.tap_off[0] + kt - 1 > 0) return false;
    smem = std::max(smem, ((size_t)kt * C * C + (size_t)(halo + TR) * (C + 4)) * sizeof(float));
    b.p[i] = p;
  }
  for (int i = count; i < 3; ++i) b.p[i] = ps[0];
  if (smem > 96 * 1024) return false;
  if (C == 16) { if (TR == 256) launch_cfg<16, 256>(b, count, smem, st); else launch_cfg<16, 128>(b, count, smem, st); }
  else { if (TR == 128) launch_cfg<32, 128>(b, count, smem, st); else launch_cfg<32, 64>(b, count, smem, st); }
  return true;
}

}  // namespace svanon
