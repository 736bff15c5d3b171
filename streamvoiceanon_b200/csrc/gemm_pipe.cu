// Latency-oriented fp32 multi-tap GEMM for the SMALL problems of the streaming path (per-frame vocoder convs:
// M = 32..2048 rows, N = 16..256; encoder window: M = 128..512) -- same contract as gemm.cu.
//
// The register double-buffer of gemm.cu keeps one 16-deep K-slab in flight, so a CTA that owns a long K range
// pays one L2/DRAM round trip per slab.  Here the operands stream through a 4-stage cp.async (LDGSTS) pipeline of
// 32-deep slabs: 96 K-elements per row are in flight while one slab is multiplied.  Tiles sit in shared memory as
// [row][32 floats] (128 B rows, K-major) with the 16-byte chunks XOR-swizzled by (row & 7) -- the same layout TMA
// produces with SWIZZLE_128B -- so the float4-along-K reads of 8 different rows hit 8 different bank groups.
// Split-K over a thread-block cluster with a DSMEM reduction in the rank-0 CTA is inherited from gemm.cu.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace svanon {

namespace {

constexpr int PK = 32;         // K-slab depth (floats) = one 128-byte row
constexpr int STAGES = 4;

struct PipeBatch {
  GemmParams p[3];
  int split;
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// element (row, chunk c of 4 floats) of a swizzled [rows][32] tile
__device__ __forceinline__ int swz(int row, int c) { return row * PK + ((c ^ (row & 7)) << 2); }

template <int BM, int BN, int TM, int TN, bool SPLIT>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_pipe_kernel(const PipeBatch batch) {
  constexpr int NTX = BN / TN, NTY = BM / TM, NT = NTX * NTY;
  constexpr int STAGE_FLOATS = (BM + BN) * PK;
  constexpr int A_CH = BM * 8, B_CH = BN * 8;                 // 16-byte chunks per stage
  static_assert(BM * BN <= STAGES * STAGE_FLOATS, "partial tile must fit the pipeline shared memory");
  extern __shared__ __align__(128) float smem[];
  pdl_trigger();

  const int split = SPLIT ? batch.split : 1;
  const int zb = SPLIT ? blockIdx.z / split : blockIdx.z;
  const int rank = SPLIT ? blockIdx.z % split : 0;
  const GemmParams& p = batch.p[zb];

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kSlabs = (p.K + PK - 1) / PK;
  const int total = kSlabs * p.taps;
  const int it_begin = (int)((long long)total * rank / split);
  const int it_end = (int)((long long)total * (rank + 1) / split);
  const int n_it = it_end - it_begin;
  // weights do not depend on the kernel in front: pull this CTA's slice of weight row n0 + tid into L2 while it finishes
  if (gridDim.x * gridDim.y * gridDim.z <= 160 && tid < BN && n0 + tid < p.N && n_it > 0) {
    int t = it_begin / kSlabs;
    int s0 = it_begin - t * kSlabs;
    int left = n_it;
    while (left > 0) {
      const int s1 = min(kSlabs, s0 + left);
      const int k0 = s0 * PK, k1 = min(p.K, s1 * PK);
      if (k1 > k0) l2_prefetch_bulk(p.W + ((long long)t * p.N + n0 + tid) * p.K + k0, (unsigned)(k1 - k0) * 4u);
      left -= s1 - s0;
      s0 = 0;
      ++t;
    }
  }
  pdl_wait();

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  constexpr int A_PER = A_CH / NT;
  static_assert(A_CH % NT == 0, "A chunks must divide evenly over the threads");
  long long a_row[A_PER];                // start of this thread's A rows (-1: past M); fixed over the K loop
#pragma unroll
  for (int j = 0; j < A_PER; ++j) {
    const int m = m0 + ((tid + j * NT) >> 3);
    a_row[j] = (m < p.M) ? gemm_a_row(p, m) : -1;
  }

  auto issue = [&](int it, int stage) {
    float* As = smem + stage * STAGE_FLOATS;
    float* Bs = As + BM * PK;
    const int t = it / kSlabs;
    const int k0 = (it - t * kSlabs) * PK;
    const long long off = p.tap_off[t];
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int i = tid + j * NT;
      const int row = i >> 3, c = i & 7;
      const int k = k0 + c * 4;
      const bool ok = (a_row[j] >= 0) && (k < p.K);
      const float* src = ok ? p.A + a_row[j] + off * p.lda + k : p.A;
      cp_async16(As + swz(row, c), src, ok ? 16 : 0);
    }
    for (int i = tid; i < B_CH; i += NT) {
      const int row = i >> 3, c = i & 7;
      const int n = n0 + row;
      const int k = k0 + c * 4;
      const bool ok = (n < p.N) && (k < p.K);
      const float* src = ok ? p.W + ((long long)t * p.N + n) * p.K + k : p.W;
      cp_async16(Bs + swz(row, c), src, ok ? 16 : 0);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < n_it) issue(it_begin + s, s);
    cp_async_commit();
  }

  for (int li = 0; li < n_it; ++li) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();                                   // slab li has landed; everyone is done with slab li-1
    if (li + STAGES - 1 < n_it) issue(it_begin + li + STAGES - 1, (li + STAGES - 1) % STAGES);
    cp_async_commit();
    float* As = smem + (li % STAGES) * STAGE_FLOATS;
    const float* Bs = As + BM * PK;
    if (p.prologue == PRO_SILU) {
      for (int i = tid; i < A_CH; i += NT) {
        float4* q = reinterpret_cast<float4*>(As) + i;
        float4 v = *q;
        v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
        *q = v;
      }
      __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float4 a4[TM], b4[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a4[i] = *reinterpret_cast<const float4*>(As + swz(ty + i * NTY, c));
#pragma unroll
      for (int j = 0; j < TN; ++j) b4[j] = *reinterpret_cast<const float4*>(Bs + swz(tx + j * NTX, c));
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc[i][j] = fmaf(a4[i].x, b4[j].x, acc[i][j]);
          acc[i][j] = fmaf(a4[i].y, b4[j].y, acc[i][j]);
          acc[i][j] = fmaf(a4[i].z, b4[j].z, acc[i][j]);
          acc[i][j] = fmaf(a4[i].w, b4[j].w, acc[i][j]);
        }
    }
  }
  cp_async_wait<0>();

  if (SPLIT) {
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();
    if (rank != 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) smem[(i * TN + j) * NT + tid] = acc[i][j];
    }
    cluster.sync();
    if (rank == 0) {
      for (int r = 1; r < split; ++r) {
        const float* remote = cluster.map_shared_rank(smem, r);
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] += remote[(i * TN + j) * NT + tid];
      }
    }
    cluster.sync();
    if (rank != 0) return;
  }

  // ---------------------------------------------------------------- epilogue (thread owns rows ty+i*NTY, cols tx+j*NTX)
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty + i * NTY;
    if (m >= p.M) continue;
    const long long c_row = gemm_c_row(p, m);
    const long long r_row = p.residual ? gemm_r_row(p, m) : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx + j * NTX;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += __ldg(p.bias + n);
      if (p.act == ACT_GELU) v = gelu_erf(v);
      else if (p.act == ACT_LOGCLAMP) v = logf(fmaxf(v, 1e-5f));
      if (p.gamma) v *= __ldg(p.gamma + n);
      if (p.residual) v += __ldg(p.residual + r_row + n);
      v *= p.out_scale;
      float* dst = p.C + c_row + n;
      *dst = p.accumulate ? *dst + v : v;
    }
  }
}

template <int BM, int BN, int TM, int TN>
void launch_pipe_cfg(PipeBatch& b, int count, int split, cudaStream_t st) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr size_t SMEM = (size_t)STAGES * (BM + BN) * PK * sizeof(float);
  const GemmParams& p = b.p[0];
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, count * split);
  b.split = split;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(gemm_pipe_kernel<BM, BN, TM, TN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    SV_CUDA(cudaFuncSetAttribute(gemm_pipe_kernel<BM, BN, TM, TN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 1;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = split;
  cfg.attrs = attr;
  cfg.numAttrs = split > 1 ? 2 : 1;
  if (split > 1) SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_pipe_kernel<BM, BN, TM, TN, true>, b));
  else SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_pipe_kernel<BM, BN, TM, TN, false>, b));
}

}  // namespace

// Returns false when the problem should go to the big-tile kernel of gemm.cu instead.
bool launch_gemm_pipe(const GemmParams* ps, int count, cudaStream_t st) {
  const GemmParams& p = ps[0];
  auto ctas = [&](int bm, int bn) { return (long long)((p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn) * count; };
  if (p.N > 64 && ctas(128, 128) >= 2 * 148) return false;          // large: throughput kernel
  if (p.N <= 64 && p.N > 32 && p.M >= 8192) return false;
  PipeBatch b;
  int min_slabs = 1 << 30;
  for (int i = 0; i < count; ++i) {
    b.p[i] = ps[i];
    min_slabs = std::min(min_slabs, (ps[i].K + PK - 1) / PK * ps[i].taps);
  }
  for (int i = count; i < 3; ++i) b.p[i] = ps[0];
  auto pick_split = [&](long long n_ctas) {
    int s = 1;
    while (s < 8 && n_ctas * s < 148 && min_slabs / (s * 2) >= 2) s *= 2;
    return s;
  };
  if (p.N <= 16) {
    if (p.M >= 4096) launch_pipe_cfg<128, 16, 4, 2>(b, count, pick_split(ctas(128, 16)), st);
    else launch_pipe_cfg<64, 16, 2, 2>(b, count, pick_split(ctas(64, 16)), st);
  } else if (p.N <= 32) {
    if (p.M >= 4096) launch_pipe_cfg<128, 32, 4, 4>(b, count, pick_split(ctas(128, 32)), st);
    else launch_pipe_cfg<64, 32, 4, 2>(b, count, pick_split(ctas(64, 32)), st);
  } else if (p.M <= 32) {
    launch_pipe_cfg<32, 64, 2, 4>(b, count, pick_split(ctas(32, 64)), st);
  } else if (p.M > 128 && p.N >= 128) {
    // encoder-window sized problems (M = 256..512): 8x4 register tiles keep the FMA pipe, not LDS, the limiter
    launch_pipe_cfg<128, 64, 8, 4>(b, count, pick_split(ctas(128, 64)), st);
  } else {
    launch_pipe_cfg<64, 64, 4, 4>(b, count, pick_split(ctas(64, 64)), st);
  }
  return true;
}

}  // namespace svanon
