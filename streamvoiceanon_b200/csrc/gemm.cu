// fp32 multi-tap GEMM on CUDA cores (parity mode: bit-level agreement of token ids with the fp32
// oracle needs fp32 products and fp32 accumulation; see DESIGN.md "precision").
//
//   C[m,n] (op)= epi( sum_t sum_k pro(A[(m*a_row_step + tap_off[t])*lda + k]) * W[t][n][k] )
//
// One kernel covers: Linear layers, 1x1 convs, the DFT/mel matmuls (overlapping frame rows, lda = hop),
// causal dense convs in channels-last layout (dilation 1 -> a single GEMM over k overlapping rows,
// dilation > 1 -> one tap per kernel element), strided down-sampling convs (a_row_step = stride) and
// transposed convs (N = stride*C_out).  Up to three same-shape problems run side by side (the three
// ResBlock1 branches of a HiFi-GAN ParallelBlock).
//
// Tiling: BMxBN output tile per CTA, BK = 16, register micro-tiles of (4*MG)x(4*NG), global->register
// prefetch of the next K-slab overlapped with the FMAs of the current one (double-buffered smem).
//
// Small problems (the per-frame stateful vocoder, the 512-row encoder window) would leave most of the 148 SMs
// idle, so the K loop can be split over a thread-block CLUSTER of S CTAs (S <= 8): every CTA accumulates its
// K-slice in registers, parks the partial tile in its own shared memory, and the cluster's rank-0 CTA sums the
// partials in fixed rank order through distributed shared memory before running the epilogue -- one launch, no
// global scratch, deterministic summation order.
#include <cooperative_groups.h>

#include <algorithm>

#include <vector>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace svanon {

namespace {

constexpr int BK = 16;

struct GemmBatch {
  GemmParams p[3];
  int split;   // cluster size along z (1 = no split-K)
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }

template <int BM, int BN, int MG, int NG, bool SPLIT>
__global__ void __launch_bounds__((BM / (4 * MG)) * (BN / (4 * NG)))
gemm_kernel(const GemmBatch batch) {
  pdl_trigger();
  constexpr int TM = 4 * MG, TN = 4 * NG;
  constexpr int NTX = BN / TN, NTY = BM / TM, NT = NTX * NTY;
  constexpr int A_F4 = BM * BK / 4, B_F4 = BN * BK / 4;
  constexpr int A_LD = (A_F4 + NT - 1) / NT, B_LD = (B_F4 + NT - 1) / NT;
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int SMEM_FLOATS = 2 * BK * (LDA_S + LDB_S);
  static_assert(!SPLIT || BM * BN <= SMEM_FLOATS, "partial tile must fit the staging shared memory");

  const int split = SPLIT ? batch.split : 1;
  const int zb = SPLIT ? blockIdx.z / split : blockIdx.z;       // problem index
  const int rank = SPLIT ? blockIdx.z % split : 0;              // == cluster rank (cluster dims (1,1,split))
  const GemmParams& p = batch.p[zb];
  __shared__ __align__(16) float smem[SMEM_FLOATS];
  float (*As)[BK][LDA_S] = reinterpret_cast<float (*)[BK][LDA_S]>(smem);
  float (*Bs)[BK][LDB_S] = reinterpret_cast<float (*)[BK][LDB_S]>(smem + 2 * BK * LDA_S);

  const int tid = threadIdx.x;
  const int tx = tid % NTX, ty = tid / NTX;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kIters = p.K / BK;
  const int total = kIters * p.taps;
  const int it_begin = (int)((long long)total * rank / split);
  const int it_end = (int)((long long)total * (rank + 1) / split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float4 ra[A_LD], rb[B_LD];
  long long a_row[A_LD];                 // start of this thread's A rows (-1: past M); fixed over the K loop
#pragma unroll
  for (int l = 0; l < A_LD; ++l) {
    const int m = m0 + ((tid + l * NT) >> 2);
    a_row[l] = (m < p.M) ? gemm_a_row(p, m) : -1;
  }

  auto load_tiles = [&](int it) {
    const int t = it / kIters;
    const int k0 = (it - t * kIters) * BK;
    const long long off = p.tap_off[t];
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int i = tid + l * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_F4 % NT == 0 || i < A_F4) {
        const int kq = i & 3;
        if (a_row[l] >= 0) {
          const float* src = p.A + a_row[l] + off * p.lda + k0 + kq * 4;
          v = __ldg(reinterpret_cast<const float4*>(src));
          if (p.prologue == PRO_SILU) {
            v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w);
          }
        }
      }
      ra[l] = v;
    }
#pragma unroll
    for (int l = 0; l < B_LD; ++l) {
      const int i = tid + l * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (B_F4 % NT == 0 || i < B_F4) {
        const int row = i >> 2, kq = i & 3;
        const int n = n0 + row;
        if (n < p.N) {
          const float* src = p.W + ((long long)t * p.N + n) * p.K + k0 + kq * 4;
          v = __ldg(reinterpret_cast<const float4*>(src));
        }
      }
      rb[l] = v;
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int l = 0; l < A_LD; ++l) {
      const int i = tid + l * NT;
      if (A_F4 % NT == 0 || i < A_F4) {
        const int row = i >> 2, kq = i & 3;
        As[buf][kq * 4 + 0][row] = ra[l].x;
        As[buf][kq * 4 + 1][row] = ra[l].y;
        As[buf][kq * 4 + 2][row] = ra[l].z;
        As[buf][kq * 4 + 3][row] = ra[l].w;
      }
    }
#pragma unroll
    for (int l = 0; l < B_LD; ++l) {
      const int i = tid + l * NT;
      if (B_F4 % NT == 0 || i < B_F4) {
        const int row = i >> 2, kq = i & 3;
        Bs[buf][kq * 4 + 0][row] = rb[l].x;
        Bs[buf][kq * 4 + 1][row] = rb[l].y;
        Bs[buf][kq * 4 + 2][row] = rb[l].z;
        Bs[buf][kq * 4 + 3][row] = rb[l].w;
      }
    }
  };

  if (it_begin < it_end) {
    load_tiles(it_begin);
    store_tiles(0);
    __syncthreads();
    for (int it = it_begin; it < it_end; ++it) {
      const int buf = (it - it_begin) & 1;
      if (it + 1 < it_end) load_tiles(it + 1);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int g = 0; g < MG; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(&As[buf][k][g * (BM / MG) + ty * 4]);
          a[g * 4 + 0] = v.x; a[g * 4 + 1] = v.y; a[g * 4 + 2] = v.z; a[g * 4 + 3] = v.w;
        }
#pragma unroll
        for (int g = 0; g < NG; ++g) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[buf][k][g * (BN / NG) + tx * 4]);
          b[g * 4 + 0] = v.x; b[g * 4 + 1] = v.y; b[g * 4 + 2] = v.z; b[g * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      if (it + 1 < it_end) {
        store_tiles(buf ^ 1);
        __syncthreads();
      }
    }
  }

  if (SPLIT) {
    // ---- split-K reduction over the cluster through distributed shared memory
    cg::cluster_group cluster = cg::this_cluster();
    __syncthreads();                                   // everyone is done reading As/Bs
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
    cluster.sync();                                    // keep the partials alive until rank 0 has read them
    if (rank != 0) return;
  }

  // ---------------------------------------------------------------- epilogue
#pragma unroll
  for (int gi = 0; gi < MG; ++gi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + gi * (BM / MG) + ty * 4 + i;
      if (m >= p.M) continue;
      const long long c_row = gemm_c_row(p, m);
      const long long r_row = p.residual ? gemm_r_row(p, m) : 0;
#pragma unroll
      for (int gj = 0; gj < NG; ++gj) {
        const int nb = n0 + gj * (BN / NG) + tx * 4;
        float y[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int n = nb + j;
          float v = acc[gi * 4 + i][gj * 4 + j];
          if (n < p.N) {
            if (p.bias) v += __ldg(p.bias + n);
            if (p.act == ACT_GELU) v = gelu_erf(v);
            else if (p.act == ACT_LOGCLAMP) v = logf(fmaxf(v, 1e-5f));
            if (p.gamma) v *= __ldg(p.gamma + n);
            if (p.residual) v += __ldg(p.residual + r_row + n);
            v *= p.out_scale;
          }
          y[j] = v;
        }
        float* dst = p.C + c_row + nb;
        if (nb + 3 < p.N && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
          float4 o = make_float4(y[0], y[1], y[2], y[3]);
          if (p.accumulate) {
            const float4 c = *reinterpret_cast<const float4*>(dst);
            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
          }
          *reinterpret_cast<float4*>(dst) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb + j < p.N) dst[j] = p.accumulate ? dst[j] + y[j] : y[j];
        }
      }
    }
  }
}

template <int BM, int BN, int MG, int NG>
void launch_cfg(GemmBatch& b, int count, int split, cudaStream_t st) {
  constexpr int NT = (BM / (4 * MG)) * (BN / (4 * NG));
  constexpr bool CAN_SPLIT = BM * BN <= 2 * BK * (BM + 4 + BN + 4);
  const GemmParams& p = b.p[0];
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, count);
  if constexpr (CAN_SPLIT) {
    if (split > 1) {
      b.split = split;
      grid.z = count * split;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = grid;
      cfg.blockDim = dim3(NT);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 1;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = split;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BM, BN, MG, NG, true>, b));
      return;
    }
  }
  b.split = 1;
  gemm_kernel<BM, BN, MG, NG, false><<<grid, NT, 0, st>>>(b);
}

}  // namespace

bool g_use_pdl = true;      // programmatic dependent launch for the GEMM kernels (gemm_pipe.cu, gemm_tc.cu)
bool g_gemm_use_pipe = true;
bool g_gemm_use_tc = true;
bool launch_gemm_pipe(const GemmParams* ps, int count, cudaStream_t st);   // gemm_pipe.cu
bool launch_gemm_tc(const GemmParams* ps, int count, cudaStream_t st);     // gemm_tc.cu
bool launch_conv_small(const GemmParams* ps, int count, cudaStream_t st);  // conv_small.cu
bool launch_gemm_pair(const GemmParams* ps, int count, cudaStream_t st);   // gemm_pair.cu
bool launch_gemm_pair_taps(const GemmParams* ps, int count, cudaStream_t st);   // gemm_pair.cu: convs with taps
bool g_use_conv_small = true;

// ---- measurement aid (svanon_gemm_timing): every GEMM launch bracketed by CUDA events on its own stream, summed per
// back end.  The events sit between consecutive kernels, so programmatic dependent launch cannot overlap a GEMM's
// prologue with its predecessor while this is on: bench.py uses it in a separate pass, never for `value`.
namespace {
struct GemmTiming {
  bool on = false;
  struct Rec { cudaEvent_t e0, e1; int backend; double flop; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    SV_CUDA(cudaEventCreate(&e));
    return e;
  }
} g_gemm_timing;
double gemm_flop(const GemmParams* ps, int count) {
  double f = 0;
  for (int i = 0; i < count; ++i) f += 2.0 * ps[i].M * ps[i].N * ps[i].K * ps[i].taps;
  return f;
}
}  // namespace

void gemm_timing_enable(bool on) {
  SV_CUDA(cudaDeviceSynchronize());
  for (auto& r : g_gemm_timing.recs) { g_gemm_timing.pool.push_back(r.e0); g_gemm_timing.pool.push_back(r.e1); }
  g_gemm_timing.recs.clear();
  g_gemm_timing.on = on;
}
void gemm_timing_read(double* ms, double* gflop, long long* launches) {
  SV_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < GEMM_BACKENDS; ++i) { ms[i] = 0; gflop[i] = 0; launches[i] = 0; }
  for (auto& r : g_gemm_timing.recs) {
    float t = 0.f;
    SV_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms[r.backend] += t; gflop[r.backend] += r.flop * 1e-9; launches[r.backend] += 1;
  }
}

// a kernel that runs several GEMMs inside one launch (chain.cu) brackets itself: counted as one tensor-core launch
bool gemm_timing_on() { return g_gemm_timing.on; }
void gemm_timing_external(cudaStream_t st, bool begin, double flop) {
  static cudaEvent_t e0 = nullptr;
  if (begin) {
    e0 = g_gemm_timing.get();
    SV_CUDA(cudaEventRecord(e0, st));
    return;
  }
  GemmTiming::Rec r{e0, g_gemm_timing.get(), GEMM_BACKEND_TC, flop};
  SV_CUDA(cudaEventRecord(r.e1, st));
  g_gemm_timing.recs.push_back(r);
}

static void launch_gemm_dispatch(const GemmParams* ps, int count, cudaStream_t st, int* backend);

// Back ends other than the pair kernel compute the fused forms of GemmParams (SwiGLU gate, RoPE) through the row-wise kernels.
static void launch_gemm_fused(const GemmParams* ps, int count, cudaStream_t st, int* backend) {
  const GemmParams& p = ps[0];
  if (!p.W2 && !p.rope_table) { launch_gemm_dispatch(ps, count, st, backend); return; }
  SV_CHECK(count == 1, "fused GEMM forms take one problem per launch");
  if (g_gemm_use_tc && launch_gemm_pair(ps, 1, st)) {
    *backend = GEMM_BACKEND_TC;
    SV_LAUNCHED();
    return;
  }
  if (p.W2) {
    SV_CHECK(p.dual_tmp && !p.bias && !p.gamma && !p.residual && !p.accumulate && p.seg_rows == 0 && p.ldc == p.N,
             "SwiGLU GEMM: plain output, no epilogue terms, scratch for the two products");
    GemmParams q[2] = {p, p};
    for (auto& g : q) { g.W2 = nullptr; g.dual_tmp = nullptr; g.Clo = nullptr; g.ldc = 2 * p.N; }
    q[0].C = p.dual_tmp;
    q[1].C = p.dual_tmp + p.N;
    q[1].W = p.W2;
    launch_gemm_dispatch(q, 2, st, backend);
    launch_silu_mul(p.dual_tmp, p.C, p.M, p.N, st, p.Clo);
    return;
  }
  GemmParams q = p;
  q.rope_table = nullptr;
  launch_gemm_dispatch(&q, 1, st, backend);
  SV_CHECK(p.seg_rows == 0 && p.ldc == 3LL * (p.rope_cols / 2), "RoPE GEMM: plain fused qkv output");
  launch_rope_qk(p.C, p.rope_table, p.M, p.rope_cols / (2 * HEAD_DIM), p.rope_pos0, st, p.rope_seg_rows);
}

void launch_gemm(const GemmParams* ps, int count, cudaStream_t st) {
  if (!g_gemm_timing.on) {
    int backend;
    launch_gemm_fused(ps, count, st, &backend);
    return;
  }
  GemmTiming::Rec r{g_gemm_timing.get(), g_gemm_timing.get(), 0, gemm_flop(ps, count) * (ps[0].W2 ? 2.0 : 1.0)};
  SV_CUDA(cudaEventRecord(r.e0, st));
  launch_gemm_fused(ps, count, st, &r.backend);
  SV_CUDA(cudaEventRecord(r.e1, st));
  g_gemm_timing.recs.push_back(r);
}

static void launch_gemm_dispatch(const GemmParams* ps, int count, cudaStream_t st, int* backend) {
  *backend = GEMM_BACKEND_FP32;
  SV_CHECK(count >= 1 && count <= 3, "gemm batch count");
  GemmBatch b;
  int min_iters = 1 << 30;
  for (int i = 0; i < count; ++i) {
    b.p[i] = ps[i];
    SV_CHECK(ps[i].K % BK == 0 && ps[i].K > 0, "gemm K must be a positive multiple of 16");
    SV_CHECK(ps[i].lda % 4 == 0, "gemm lda must be a multiple of 4");
    SV_CHECK(ps[i].taps >= 1 && ps[i].taps <= MAX_TAPS, "gemm taps");
    SV_CHECK(ps[i].M == ps[0].M && ps[i].N == ps[0].N, "batched gemm problems must share M and N");
    min_iters = std::min(min_iters, ps[i].K / BK * ps[i].taps);
  }
  for (int i = count; i < 3; ++i) b.p[i] = ps[0];
  const GemmParams& p = ps[0];
  if (p.M <= 0 || p.N <= 0) return;
  // the 32-channel HiFi-GAN level at many streams: the pair kernel's conv form (tensor cores, L2-bound) before the CUDA-core kernel
  static const bool conv32_pair = [] {
    const char* e = getenv("SVANON_CONV32_PAIR");          // 0: the 32-channel level stays on conv_small at every stream count
    return !e || atoi(e) != 0;
  }();
  static const bool conv16_pair = [] {
    const char* e = getenv("SVANON_CONV16_PAIR");          // 1: the 16-channel level on the pair kernel too (BN = 16, 64-byte swizzle
    return e && atoi(e) != 0;                              // rows).  Correct, measured SLOWER than conv_small (V 4.16 vs 3.99 ms at 128
                                                           // streams, profiles/r2zza_*: 16 KB of A per 128 x 16 x 16 slab): off by default
  }();
  if (g_gemm_use_tc && ((conv32_pair && p.N == 32 && p.M >= 32768) || (conv16_pair && p.N == 16 && p.M >= 65536)) &&
      launch_gemm_pair_taps(ps, count, st)) {
    *backend = GEMM_BACKEND_TC;
    SV_LAUNCHED();
    return;
  }
  if (g_use_conv_small && launch_conv_small(ps, count, st)) {     // thin causal convs (HiFi-GAN levels with 16/32 channels)
    *backend = GEMM_BACKEND_CONV_SMALL;
    SV_LAUNCHED();
    return;
  }
  if (g_gemm_use_tc && launch_gemm_pair(ps, count, st)) {       // tcgen05 3xTF32 on CTA pairs, operands by tensor-map TMA (M >= 4096)
    *backend = GEMM_BACKEND_TC;
    SV_LAUNCHED();
    return;
  }
  if (g_gemm_use_tc && launch_gemm_pair_taps(ps, count, st)) {  // the same kernel for causal convs with taps / SiLU input (M >= 4096)
    *backend = GEMM_BACKEND_TC;
    SV_LAUNCHED();
    return;
  }
  if (g_gemm_use_tc && launch_gemm_tc(ps, count, st)) {         // tcgen05 3xTF32 (M >= 32, N >= 64)
    *backend = GEMM_BACKEND_TC;
    SV_LAUNCHED();
    return;
  }
  if (g_gemm_use_pipe && launch_gemm_pipe(ps, count, st)) {     // small / latency-bound problems
    *backend = GEMM_BACKEND_PIPE;
    SV_LAUNCHED();
    return;
  }
  auto ctas = [&](int bm, int bn) { return (long long)((p.M + bm - 1) / bm) * ((p.N + bn - 1) / bn) * count; };
  // split-K over a cluster when the plain grid cannot fill the 148 SMs twice and each slice keeps >= 4 K-slabs
  auto pick_split = [&](long long n_ctas) {
    int s = 1;
    while (s < 8 && n_ctas * s < 2 * 148 && min_iters / (s * 2) >= 4) s *= 2;
    return s;
  };
  if (p.N <= 16) {
    launch_cfg<128, 16, 1, 1>(b, count, pick_split(ctas(128, 16)), st);
  } else if (p.N <= 32) {
    launch_cfg<128, 32, 2, 1>(b, count, pick_split(ctas(128, 32)), st);
  } else if (p.N <= 64 && p.M >= 2048) {
    launch_cfg<128, 64, 2, 1>(b, count, 1, st);
  } else if (ctas(128, 128) >= 2 * 148) {
    launch_cfg<128, 128, 2, 2>(b, count, 1, st);
  } else if (p.M > 32) {
    launch_cfg<64, 64, 1, 1>(b, count, pick_split(ctas(64, 64)), st);
  } else {
    launch_cfg<32, 64, 1, 1>(b, count, pick_split(ctas(32, 64)), st);
  }
  SV_LAUNCHED();
}

}  // namespace svanon
