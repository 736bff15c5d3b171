// Tensor-core GEMM for the dense projections of the encoder window and the AR prompt prefill: tcgen05.mma
// (5th-gen tensor cores, accumulator in TMEM) with fp32-grade products through a 3xTF32 split.
//
// Why a split: token ids are decided by sign(proj) (BSQ) and argmax(p/q) (sampler), so parity with the fp32
// reference needs ~fp32 products; a single TF32 pass (10-bit mantissa) flips ids.  Every fp32 operand x is split
// into hi = tf32(x) and lo = x - hi (exact in fp32; the tensor core truncates lo to TF32 again), and
//     x*y ~= hi_x*hi_y + hi_x*lo_y + lo_x*hi_y          (dropped terms <= 2^-20 |x*y|)
// is accumulated in fp32 in TMEM: three kind::tf32 MMAs per K-step.
//
// Data path: 128-byte-swizzled K-major tiles ([row][32 floats], 16-byte chunks XORed with row&7 == UMMA/TMA
// SWIZZLE_128B).  The operands have to pass through registers anyway (the hi/lo split is arithmetic), so the 16
// producer warps load their chunks global -> registers several K-slabs ahead, split, and store hi and lo tiles
// straight into a free pipeline stage, fence the generic->async proxy, and one thread issues the 12 MMAs of the slab
// (4 K-steps x 3 products) and commits them to an mbarrier that frees the stage.  CTA tile 128 x BN, one CTA per SM,
// accumulator = BN TMEM columns, epilogue reads TMEM with tcgen05.ld.  Split-K over a thread-block cluster with a
// DSMEM reduction distributed over the ranks.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>
#include <unordered_map>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace svanon {

namespace {

constexpr int TK = 32;                 // K-slab = one 128-byte swizzle row
constexpr int TBM = 128;

struct TcBatch {
  GemmParams p[3];
  int split;
  int wprefetch;     // weight-slice L2 prefetch before griddepcontrol.wait (SVANON_TC_WPREFETCH=0 switches it off)
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ int swz(int row, int c) { return row * TK + ((c ^ (row & 7)) << 2); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni TC_WAIT_DONE;\n"
      "bra.uni TC_WAIT_LOOP;\n"
      "TC_WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp: start>>4 [0,14),
// LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) = 2)
__device__ __forceinline__ unsigned long long umma_desc(const void* tile) {
  const unsigned long long addr = smem_u32(tile);
  return ((addr & 0x3FFFFull) >> 4) | (1ull << 16) | ((1024ull >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor: D=f32 (bit 4), A=B=tf32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ unsigned umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TBM >> 4) << 24);
}
// kind::f16 instruction descriptor: D=f32 (bit 4), A=B=f16 (format 0 at [7,10) and [10,13)), K-major, N>>3, M>>4
__device__ __forceinline__ unsigned umma_idesc_f16(int n) {
  return (1u << 4) | ((unsigned)(n >> 3) << 17) | ((unsigned)(TBM >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                         unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float (&v)[16]) {
  unsigned r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

#ifdef SVANON_TC_PROF
// tuning build only (-DSVANON_TC_PROF, tools/profile_gemm_timeline.py): clock64 marks of thread 0 (producer) and the MMA
// thread of CTA (0,0,0), printed by svanon_debug_gemm
__device__ long long g_tc_prof[16];
#define TC_MARK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (tid == 0 || tid == TC_PRODUCERS)) g_tc_prof[(i) + (tid == 0 ? 0 : 8)] = clock64(); } while (0)
#else
#define TC_MARK(i) do {} while (0)
#endif

constexpr int TC_PRODUCER_WARPS = 16;              // warps 0-15: global -> registers -> hi/lo split -> swizzled tiles; epilogue
constexpr int TC_PRODUCERS = TC_PRODUCER_WARPS * 32;
constexpr int TC_THREADS = TC_PRODUCERS + 32;      // warp 16: MMA issuer (one elected lane)
constexpr int TC_DEPTH_SMALL = 4;                  // ... on the 128 x 64 tile (see the kernel)
constexpr int TC_DEPTH = 2;                        // K-slabs a producer thread keeps in flight in registers (2, 3 and 4
                                                   // measure the same: the loop is not bound by load latency)

__device__ __forceinline__ float4 ldg_nc(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }
// Warp-specialised.  16 producer warps stream the K-slabs: every thread keeps TC_DEPTH slabs of its own 16-byte chunks
// in flight in REGISTERS (plain 128-bit no-allocate loads; the weight chunks of the first slabs are requested before
// griddepcontrol.wait), splits a slab into hi/lo TF32 terms straight into the 128-byte-swizzled tiles of a free
// pipeline stage and arrives on full[stage].  One thread of warp 16 waits on full[stage], issues the slab's 12
// tcgen05.mma and commits them to empty[stage].  The first kernel version used 4 producer warps and was bound by
// their instruction latency (ncu: 1.3 warps per scheduler, 9.4 cycles per issued instruction, ~2.7 us per slab);
// with 16 warps a slab costs each thread 3-4 chunks.  After the last slab all 16 warps run the epilogue
// (TMEM -> registers -> global).  Split-K over a thread-block cluster: every rank parks its partial tile in its own
// shared memory and then reduces (in fixed rank order) and writes ONE column slice of the tile, so the DSMEM reads are
// spread over all ranks instead of being serialised in rank 0.
//
// HALF = true is the PERF MODE of svanon_set_precision (the reference's own GPU precision: fp16 autocast,
// evaluations/infer_arvc.py:493): one kind::f16 pass instead of the 3xTF32 split.  A K-slab is still one 128-byte swizzle
// row per tile row -- 64 halves instead of 32 floats -- so the tile geometry, descriptors and the pipeline are the same;
// activations (fp32 in HBM) are rounded to fp16 by the producers, weights come from an fp16 copy made once per weight
// (GemmParams::Wh), there is one A tile and one B tile per stage, and a slab costs 4 MMAs instead of 12.
template <int BN, int STAGES, bool SPLIT, bool HALF>
__global__ void __launch_bounds__(TC_THREADS, 1) gemm_tc_kernel(const TcBatch batch) {
  constexpr int A_FLOATS = TBM * TK, B_FLOATS = BN * TK;              // tile sizes in 4-byte words (128 bytes per row)
  constexpr int STAGE_FLOATS = HALF ? (A_FLOATS + B_FLOATS) : (2 * A_FLOATS + 2 * B_FLOATS);   // A_hi | A_lo | B_hi | B_lo
  constexpr int TKE = HALF ? 2 * TK : TK;                             // K elements per slab
  constexpr int AV = HALF ? 2 : 1;                                    // float4 loads per 16-byte A chunk
  // 16-byte chunks per thread.  Thin tiles (BN = 16 / 32: the low-channel HiFi-GAN levels at many streams) have fewer B chunks
  // than producer threads: they are only launched with the B operand arriving by TMA, the register path stays compiled
  // for one (guarded) chunk.
  constexpr int A_PER = TBM * 8 / TC_PRODUCERS, B_PER = (BN * 8 >= TC_PRODUCERS) ? BN * 8 / TC_PRODUCERS : 1;
  constexpr int TM_COLS = BN < 32 ? 32 : BN;                          // tensor-memory allocation granule
  // K-slabs a producer thread keeps in flight in registers.  Wide tiles are MMA-bound (2, 3 and 4 measure the same); the
  // 128 x 64 tile of the small, latency-bound problems pays one L2 round trip per D slabs (0.35-0.5 us per slab at D = 2:
  // profiles/README.md timeline), and has the registers for 4.  fp16 slabs are twice as deep in K; 2 of them spill at BN = 256.
  constexpr int D = (HALF && BN == 256) ? 1 : (BN == 64 && !HALF && TC_DEPTH_SMALL > TC_DEPTH ? TC_DEPTH_SMALL : TC_DEPTH);
  static_assert(TBM * BN <= STAGES * STAGE_FLOATS, "partial tile must fit the pipeline shared memory");
  static_assert(A_PER >= 1 && B_PER >= 1, "tile too small for the producer count");
  extern __shared__ unsigned char dsmem_raw[];
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES], acc_bar;
  __shared__ unsigned tmem_holder;
  __shared__ float s_bias[BN], s_gamma[BN];

  const int split = SPLIT ? batch.split : 1;
  const int zb = SPLIT ? blockIdx.z / split : blockIdx.z;
  const int rank = SPLIT ? blockIdx.z % split : 0;
  const GemmParams& p = batch.p[zb];
  const int tid = threadIdx.x, lane = tid & 31;
  // broadcast from lane 0: the compiler then knows the warp index -- and every role branch on it -- is warp-uniform
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * BN;
  const int kSlabs = (p.K + TKE - 1) / TKE;
  const int total = kSlabs * p.taps;
  const int it_begin = (int)((long long)total * rank / split);
  const int it_end = (int)((long long)total * (rank + 1) / split);
  const int n_it = it_end - it_begin;

  // B operand by TMA: the weights exist pre-split (hi / lo, or fp16) and pre-tiled in HBM -- per (tap, K-slab) the rows of
  // all output channels as 128-byte swizzled tile rows -- so the BN rows of a slab are ONE contiguous block per term and
  // arrive with one cp.async.bulk each, completing on the stage's `full` barrier; the producer warps only handle A.
  const bool tma_b = p.Wt[0] != nullptr;
  TC_MARK(0);
  pdl_trigger();
  if (tma_b && batch.wprefetch && tid < n_it) {
    // this CTA's weight blocks: one per slab (and term), contiguous
    const int it = it_begin + tid;
    const int rows = min(BN, p.wt_npad - n0);
    const long long off = ((long long)it * p.wt_npad + n0) * 128;
    l2_prefetch_bulk(reinterpret_cast<const char*>(p.Wt[0]) + off, (unsigned)rows * 128u);
    if (!HALF) l2_prefetch_bulk(reinterpret_cast<const char*>(p.Wt[1]) + off, (unsigned)rows * 128u);
  }
  if (!tma_b && batch.wprefetch && tid < BN && n0 + tid < p.N && n_it > 0) {
    // this CTA's slice of weight row n0 + tid: slabs [it_begin, it_end) = per tap a contiguous K range
    const unsigned esz = HALF ? 2u : 4u;
    const char* wbase = HALF ? reinterpret_cast<const char*>(p.Wh) : reinterpret_cast<const char*>(p.W);
    int t = it_begin / kSlabs;
    int s0 = it_begin - t * kSlabs;
    int left = n_it;
    while (left > 0) {
      const int s1 = min(kSlabs, s0 + left);
      const int k0 = s0 * TKE, k1 = min(p.K, s1 * TKE);
      if (k1 > k0)
        l2_prefetch_bulk(wbase + (((long long)t * p.N + n0 + tid) * p.K + k0) * esz, (unsigned)(k1 - k0) * esz);
      left -= s1 - s0;
      s0 = 0;
      ++t;
    }
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], TC_PRODUCER_WARPS + (tma_b ? 1 : 0)); mbar_init(&empty_bar[s], 1); }
    mbar_init(&acc_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(TM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < BN; i += TC_THREADS) {
    const int n = n0 + i;
    s_bias[i] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.f;
    s_gamma[i] = (p.gamma && n < p.N) ? __ldg(p.gamma + n) : 1.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_d = tmem_holder;
  TC_MARK(1);

  if (warp < TC_PRODUCER_WARPS) {
    // =========================================================== producers
    // fixed per-thread chunk assignment: chunk i = tid + j*512 -> row = i>>3, c = i&7.  All loop state is advanced
    // incrementally (no division, 32-bit shared-memory addresses): the loop is issue-bound, every instruction counts.
    const float* a_ptr[A_PER];           // row base + this thread's chunk; null: row past M
    unsigned a_soff[A_PER];
#pragma unroll
    for (int j = 0; j < A_PER; ++j) {
      const int i = tid + j * TC_PRODUCERS, row = i >> 3, c = i & 7;
      const int m = m0 + row;
      a_ptr[j] = (m < p.M) ? p.A + gemm_a_row(p, m) + c * (HALF ? 8 : 4) : nullptr;
      a_soff[j] = (unsigned)swz(row, c) * 4u;
    }
    const float* b_ptr[B_PER];            // HALF: points into the fp16 copy (addressed in 4-byte words: 2 halves each)
    unsigned b_soff[B_PER];
#pragma unroll
    for (int j = 0; j < B_PER; ++j) {
      const int i = tid + j * TC_PRODUCERS, row = i >> 3, c = i & 7;
      const int n = (i < BN * 8) ? n0 + row : p.N;                    // thin tiles: chunks past the tile are nobody's
      if (HALF) b_ptr[j] = (n < p.N) ? reinterpret_cast<const float*>(p.Wh) + (((long long)n * p.K) >> 1) + c * 4 : nullptr;
      else b_ptr[j] = (n < p.N) ? p.W + (long long)n * p.K + c * 4 : nullptr;
      b_soff[j] = (unsigned)swz(row, c) * 4u;
    }
    const int c4 = (tid & 7) * (HALF ? 8 : 4);         // K offset of this thread's chunks inside a slab
    const bool silu = p.prologue == PRO_SILU;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const long long tap_stride = (long long)p.N * p.K;   // in K elements
    // load cursor: tap ld_t, K offset ld_k (elements) of the next slab to request
    int ld_t = it_begin / kSlabs;
    int ld_k = (it_begin - ld_t * kSlabs) * TKE;
    long long ld_a = (long long)p.tap_off[ld_t] * p.lda, ld_b = (long long)ld_t * tap_stride;
    float4 ra[D][A_PER * AV], rb[D][B_PER];
    auto load_b = [&](float4 (&dst)[B_PER]) {
      if (tma_b) return;
      const bool k_ok = (ld_k + c4) < p.K;
#pragma unroll
      for (int j = 0; j < B_PER; ++j)       // HALF: offsets in 4-byte words of the fp16 copy (2 elements each)
        dst[j] = (b_ptr[j] && k_ok) ? ldg_nc(b_ptr[j] + (HALF ? ((ld_b + ld_k) >> 1) : (ld_b + ld_k))) : zero4;
    };
    auto load_a = [&](float4 (&dst)[A_PER * AV]) {
      const bool k_ok = (ld_k + c4) < p.K;
#pragma unroll
      for (int j = 0; j < A_PER; ++j) {
#pragma unroll
        for (int v = 0; v < AV; ++v) dst[j * AV + v] = (a_ptr[j] && k_ok) ? ldg_nc(a_ptr[j] + ld_a + ld_k + 4 * v) : zero4;
      }
    };
    auto advance = [&] {
      ld_k += TKE;
      if (ld_k >= kSlabs * TKE) {
        ld_k = 0;
        ++ld_t;
        ld_a = (long long)p.tap_off[ld_t < p.taps ? ld_t : 0] * p.lda;
        ld_b += tap_stride;
      }
    };
    // Weights do not depend on the previous kernel: request them before waiting for it.  The cursor is replayed for
    // the activation loads of the same slabs afterwards.
    {
      const int t0 = ld_t, k0 = ld_k;
      const long long a0 = ld_a, b0 = ld_b;
#pragma unroll
      for (int d = 0; d < D; ++d)
        if (d < n_it) { load_b(rb[d]); advance(); }
      pdl_wait();
      TC_MARK(2);
      ld_t = t0; ld_k = k0; ld_a = a0; ld_b = b0;
#pragma unroll
      for (int d = 0; d < D; ++d)
        if (d < n_it) { load_a(ra[d]); advance(); }
    }
    const unsigned smem_base = smem_u32(smem);
    const unsigned full_base = smem_u32(&full_bar[0]), empty_base = smem_u32(&empty_bar[0]);
    unsigned st_stage = 0, st_parity = 1;              // parity to wait for on empty[stage]
    auto sts4 = [](unsigned addr, float4 v) {
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    };
    auto split_sts = [&](unsigned hi_addr, unsigned lo_delta, float4 v) {
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
      sts4(hi_addr, h);
      sts4(hi_addr + lo_delta, l);
    };
    for (int li0 = 0; li0 < n_it; li0 += D) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const int li = li0 + d;
        if (li < n_it) {
          // MMAs that read this stage are done.  One lane per warp polls / arrives: 512 threads hammering one
          // mbarrier word serialise in the shared-memory pipe that the tile stores need.
          if (lane == 0) {
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "TC_PW_LOOP:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra.uni TC_PW_DONE;\n"
                "bra.uni TC_PW_LOOP;\n"
                "TC_PW_DONE:\n"
                "}\n" ::"r"(empty_base + st_stage * 8u), "r"(st_parity) : "memory");
          }
          __syncwarp();
          const unsigned a_stage = smem_base + st_stage * (unsigned)(STAGE_FLOATS * 4);
          const unsigned b_stage = a_stage + 2u * A_FLOATS * 4u;
          if (tma_b && warp == 0 && elect_one()) {         // (uniform branch + elected lane: no per-copy R2UR loop; measured neutral)
            const int rows = min(BN, p.wt_npad - n0);
            const unsigned bytes = (unsigned)rows * 128u;
            const long long off = ((long long)(it_begin + li) * p.wt_npad + n0) * 128;
            const unsigned bar = full_base + st_stage * 8u;
            const unsigned b0 = HALF ? a_stage + (unsigned)A_FLOATS * 4u : b_stage;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(HALF ? bytes : 2u * bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(b0), "l"(reinterpret_cast<const char*>(p.Wt[0]) + off), "r"(bytes), "r"(bar) : "memory");
            if (!HALF)
              asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                           ::"r"(b0 + (unsigned)B_FLOATS * 4u), "l"(reinterpret_cast<const char*>(p.Wt[1]) + off), "r"(bytes), "r"(bar) : "memory");
          }
          if constexpr (HALF) {
            const unsigned b_stage_h = a_stage + (unsigned)A_FLOATS * 4u;
#pragma unroll
            for (int j = 0; j < A_PER; ++j) {
              float4 v0 = ra[d][j * AV], v1 = ra[d][j * AV + AV - 1];
              if (silu) {
                v0.x = silu_fast(v0.x); v0.y = silu_fast(v0.y); v0.z = silu_fast(v0.z); v0.w = silu_fast(v0.w);
                v1.x = silu_fast(v1.x); v1.y = silu_fast(v1.y); v1.z = silu_fast(v1.z); v1.w = silu_fast(v1.w);
              }
              const __half2 h0 = __floats2half2_rn(v0.x, v0.y), h1 = __floats2half2_rn(v0.z, v0.w);
              const __half2 h2 = __floats2half2_rn(v1.x, v1.y), h3 = __floats2half2_rn(v1.z, v1.w);
              float4 pk;
              pk.x = __uint_as_float(*reinterpret_cast<const unsigned*>(&h0));
              pk.y = __uint_as_float(*reinterpret_cast<const unsigned*>(&h1));
              pk.z = __uint_as_float(*reinterpret_cast<const unsigned*>(&h2));
              pk.w = __uint_as_float(*reinterpret_cast<const unsigned*>(&h3));
              sts4(a_stage + a_soff[j], pk);
            }
            if (!tma_b) {
#pragma unroll
              for (int j = 0; j < B_PER; ++j) sts4(b_stage_h + b_soff[j], rb[d][j]);
            }
          } else {
#pragma unroll
            for (int j = 0; j < A_PER; ++j) {
              float4 v = ra[d][j];
              if (silu) { v.x = silu_fast(v.x); v.y = silu_fast(v.y); v.z = silu_fast(v.z); v.w = silu_fast(v.w); }
              split_sts(a_stage + a_soff[j], A_FLOATS * 4u, v);
            }
            if (!tma_b) {
#pragma unroll
              for (int j = 0; j < B_PER; ++j) split_sts(b_stage + b_soff[j], B_FLOATS * 4u, rb[d][j]);
            }
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic writes -> visible to the MMA
          __syncwarp();
          if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full_base + st_stage * 8u) : "memory");
          if (++st_stage == STAGES) { st_stage = 0; st_parity ^= 1u; }
          if (li + D < n_it) { load_b(rb[d]); load_a(ra[d]); advance(); }
        }
      }
    }
    TC_MARK(3);
  } else {
    pdl_wait();
    TC_MARK(2);
    {
      // =========================================================== MMA issuer
      // The whole warp walks the loop and ONE ELECTED lane issues: with everything the descriptors are built from
      // warp-uniform (the accumulator address is broadcast from lane 0), they live in uniform registers and the slab's MMAs
      // issue back to back.  Issued from inside an `if (lane == 0)` branch the compiler wrapped EVERY tcgen05.mma in an
      // ELECT / 7 x R2UR.BROADCAST / BRA.U.ANY loop: ~45 ns per instruction, 0.55 us per 12-MMA slab (chain-kernel trace,
      // profiles/r2m_*), which is what bounded the small tiles.
      const unsigned tmem_u = (unsigned)__shfl_sync(0xffffffffu, (int)tmem_d, 0);
      const unsigned idesc = HALF ? umma_idesc_f16(BN) : umma_idesc(BN);
      for (int li = 0; li < n_it; ++li) {
        const int stage = li % STAGES;
        mbar_wait(&full_bar[stage], (li / STAGES) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* As = smem + stage * STAGE_FLOATS;
        if (elect_one()) {
          if constexpr (HALF) {
            const unsigned long long a_d = umma_desc(As), b_d = umma_desc(As + A_FLOATS);
#pragma unroll
            for (int kk = 0; kk < TK / 8; ++kk) {
              const unsigned long long adv = (unsigned long long)(kk * 32 >> 4);   // 16 halves = 32 bytes per K-step
              umma_f16(tmem_u, a_d + adv, b_d + adv, idesc, (li > 0 || kk > 0) ? 1u : 0u);
            }
          } else {
            float* Bs = As + 2 * A_FLOATS;
            const unsigned long long a_hi = umma_desc(As), a_lo = umma_desc(As + A_FLOATS);
            const unsigned long long b_hi = umma_desc(Bs), b_lo = umma_desc(Bs + B_FLOATS);
#pragma unroll
            for (int kk = 0; kk < TK / 8; ++kk) {
              const unsigned long long adv = (unsigned long long)(kk * 32 >> 4);   // 8 tf32 = 32 bytes per K-step
              umma_tf32(tmem_u, a_hi + adv, b_lo + adv, idesc, (li > 0 || kk > 0) ? 1u : 0u);
              umma_tf32(tmem_u, a_lo + adv, b_hi + adv, idesc, 1u);
              umma_tf32(tmem_u, a_hi + adv, b_hi + adv, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&acc_bar);
      __syncwarp();
      TC_MARK(3);
    }
  }
  __syncwarp();

  // ---------------------------------------------------------------- TMEM -> registers -> (split-K reduce) -> global
  // warp w reads TMEM lanes (w & 3) * 32 .. +31 (the hardware's lane quadrant of a warp) and the column group w >> 2
  constexpr int CG = BN >= 64 ? BN / 4 : 16;        // columns per column group (thin tiles: BN / 16 groups of 16, the other warps idle)
  const int quad = warp & 3, grp = warp >> 2;
  cg::cluster_group cluster = cg::this_cluster();
  if (warp < TC_PRODUCER_WARPS) {
    if (n_it > 0 && lane == 0) mbar_wait(&acc_bar, 0);
    __syncwarp();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC_MARK(4);
  }
  auto finish = [&](float (&v)[16], int m, int col0) {      // bias/act/gamma/residual/scale + store of 16 columns
    const int n_base = n0 + col0;
    const long long c_row = gemm_c_row(p, m);
    float* dst = p.C + c_row + n_base;
    const float* res = p.residual ? p.residual + gemm_r_row(p, m) + n_base : nullptr;
    const bool vec = (n_base + 15 < p.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                     (!res || (reinterpret_cast<uintptr_t>(res) & 15) == 0) && !p.accumulate;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float y = v[j] + s_bias[col0 + j];
      if (p.act == ACT_GELU) y = gelu_erf(y);
      else if (p.act == ACT_LOGCLAMP) y = logf(fmaxf(y, 1e-5f));
      v[j] = y * s_gamma[col0 + j];
    }
    if (vec) {
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (res) {
          const float4 r4 = __ldg(reinterpret_cast<const float4*>(res + j));
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale;
        *reinterpret_cast<float4*>(dst + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        if (n_base + j < p.N) {
          float y = v[j];
          if (res) y += __ldg(res + j);
          y *= p.out_scale;
          dst[j] = p.accumulate ? dst[j] + y : y;
        }
      }
    }
  };
  if (!SPLIT) {
    // A thread holds ROWS of the accumulator (TMEM lane = row), so direct stores would write 32 rows x 16 bytes per
    // instruction (32 sectors, half filled; the same for the residual loads): ~9 us per 128 x 256 tile, more than its
    // math at K = 512.  Every warp therefore transposes its 32 x CG block through its own slice of the (now idle)
    // pipeline shared memory and then reads the residual and writes C in whole row segments.
    if constexpr (BN >= 128) {
      if (warp < TC_PRODUCER_WARPS) {
        constexpr int SP = CG + 4;                      // padded row pitch of the staging block (floats)
        static_assert(TC_PRODUCER_WARPS * 32 * SP <= STAGES * STAGE_FLOATS, "staging must fit the pipeline shared memory");
        float* stg = smem + warp * (32 * SP);
        const int cg0 = grp * CG;
#pragma unroll
        for (int c0 = 0; c0 < CG; c0 += 16) {
          float v[16];
          if (n_it > 0) tmem_ld16(tmem_d + ((unsigned)(quad * 32) << 16) + cg0 + c0, v);
          else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = 0.f;
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float y = v[j] + s_bias[cg0 + c0 + j];
            if (p.act == ACT_GELU) y = gelu_erf(y);
            else if (p.act == ACT_LOGCLAMP) y = logf(fmaxf(y, 1e-5f));
            v[j] = y * s_gamma[cg0 + c0 + j];
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(stg + lane * SP + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        constexpr int LPR = CG / 4;                     // lanes per row (one float4 each)
        constexpr int RPI = 32 / LPR;                   // rows per instruction
        const int rr = lane / LPR, c = (lane % LPR) * 4;
        const int n_base = n0 + cg0 + c;
#pragma unroll 4
        for (int r0 = 0; r0 < 32; r0 += RPI) {
          const int r = r0 + rr;
          const int m = m0 + quad * 32 + r;
          if (m >= p.M || n_base >= p.N) continue;
          float4 o = *reinterpret_cast<const float4*>(stg + r * SP + c);
          float* dst = p.C + gemm_c_row(p, m) + n_base;
          const float* res = p.residual ? p.residual + gemm_r_row(p, m) + n_base : nullptr;
          const bool vec = (n_base + 3 < p.N) && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) &&
                           (!res || (reinterpret_cast<uintptr_t>(res) & 15) == 0) && !p.accumulate;
          if (vec) {
            if (res) {
              const float4 r4 = __ldg(reinterpret_cast<const float4*>(res));
              o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
            }
            o.x *= p.out_scale; o.y *= p.out_scale; o.z *= p.out_scale; o.w *= p.out_scale;
            *reinterpret_cast<float4*>(dst) = o;
          } else {
            const float e[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              if (n_base + j < p.N) {
                float y = e[j];
                if (res) y += __ldg(res + j);
                y *= p.out_scale;
                dst[j] = p.accumulate ? dst[j] + y : y;
              }
            }
          }
        }
      }
    } else {
      // 128 x 64 tile: a thread's 16 columns are two whole sectors already
      if (warp < TC_PRODUCER_WARPS) {
        const int row = quad * 32 + lane;
        const int m = m0 + row;
        if (grp * CG < BN) {
#pragma unroll
          for (int c0 = grp * CG; c0 < (grp + 1) * CG; c0 += 16) {
            float v[16];
            if (n_it > 0) tmem_ld16(tmem_d + ((unsigned)(quad * 32) << 16) + c0, v);
            else {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = 0.f;
            }
            if (m < p.M) finish(v, m, c0);
          }
        }
      }
    }
  } else {
    // every rank parks its partial tile column-major ([BN][128] floats) in its own shared memory
    if (warp < TC_PRODUCER_WARPS && grp * CG < BN) {
      const int row = quad * 32 + lane;
#pragma unroll
      for (int c0 = grp * CG; c0 < (grp + 1) * CG; c0 += 16) {
        float v[16];
        if (n_it > 0) tmem_ld16(tmem_d + ((unsigned)(quad * 32) << 16) + c0, v);
        else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = 0.f;
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) smem[(c0 + j) * TBM + row] = v[j];
      }
    }
    TC_MARK(7);
    cluster.sync();
    // rank r owns the 16-column groups r, r + split, ...: 128 rows x 16 columns = 512 (row, 4-column) items per group
    if (warp < TC_PRODUCER_WARPS) {
      const int row = tid & (TBM - 1), sub = tid >> 7;          // sub 0..3 -> columns sub*4 .. +3 of the group
      const int m = m0 + row;
      for (int g16 = rank; g16 < BN / 16; g16 += split) {
        const int col0 = g16 * 16 + sub * 4;
        float acc4[4] = {0.f, 0.f, 0.f, 0.f};
        for (int r = 0; r < split; ++r) {
          const float* part = cluster.map_shared_rank(smem, r);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc4[j] += part[(col0 + j) * TBM + row];
        }
        if (m < p.M) {
          const int n_base = n0 + col0;
          float* dst = p.C + gemm_c_row(p, m) + n_base;
          const float* res = p.residual ? p.residual + gemm_r_row(p, m) + n_base : nullptr;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (n_base + j < p.N) {
              float y = acc4[j] + s_bias[col0 + j];
              if (p.act == ACT_GELU) y = gelu_erf(y);
              else if (p.act == ACT_LOGCLAMP) y = logf(fmaxf(y, 1e-5f));
              y *= s_gamma[col0 + j];
              if (res) y += __ldg(res + j);
              y *= p.out_scale;
              dst[j] = p.accumulate ? dst[j] + y : y;
            }
          }
        }
      }
    }
  }
  TC_MARK(5);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (SPLIT) cluster.sync();
  else __syncthreads();
  TC_MARK(6);
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TM_COLS) : "memory");
}

template <int BN, int STAGES, bool HALF = false>
void launch_tc_cfg(TcBatch& b, int count, int split, cudaStream_t st) {
  constexpr size_t SMEM = (size_t)STAGES * (HALF ? 1 : 2) * (TBM * TK + BN * TK) * sizeof(float) + 1024;
  const GemmParams& p = b.p[0];
  dim3 grid((p.N + BN - 1) / BN, (p.M + TBM - 1) / TBM, count * split);
  b.split = split;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, true, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    SV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, false, HALF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(TC_THREADS);
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
  if (split > 1) SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, true, HALF>, b));
  else SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, false, HALF>, b));
}

}  // namespace

#ifdef SVANON_TC_PROF
void tc_prof_dump() {
  long long h[16];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_tc_prof, sizeof(h));
  const char* names[8] = {"entry", "prologue done (bars, TMEM alloc, bias)", "griddepcontrol.wait done", "main loop done", "accumulator ready",
                          "epilogue done", "final sync done", "partials parked (split)"};
  for (int w = 0; w < 2; ++w) {
    fprintf(stderr, "%s:", w ? "  mma thread" : "  producer t0");
    for (int i = 1; i < 8; ++i)
      if (h[i + 8 * w] && h[8 * w]) fprintf(stderr, "  [%s] +%.2f us", names[i], (h[i + 8 * w] - h[8 * w]) / 1965.0);
    fprintf(stderr, "\n");
  }
}
#endif

// ---- perf mode: fp16 copies of the (immutable) engine weights, made once per weight pointer on first use
bool g_gemm_half = false;

namespace {
__global__ void f32_to_f16_kernel(const float* __restrict__ src, __half* __restrict__ dst, size_t n) {
  const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<__half2*>(dst + i) = __floats2half2_rn(v.x, v.y);
    *reinterpret_cast<__half2*>(dst + i + 2) = __floats2half2_rn(v.z, v.w);
  } else {
    for (size_t j = i; j < n; ++j) dst[j] = __float2half_rn(src[j]);
  }
}
struct HalfCopy { __half* data; size_t n; };
std::unordered_map<const float*, HalfCopy>& half_registry() {
  static std::unordered_map<const float*, HalfCopy> r;
  return r;
}
}  // namespace

// fp16 copy of a weight tensor (keyed on its pointer; engine weights never change after finalize).  The first request
// allocates and converts on `st`; while a stream capture is running no new copy can be made (null: the caller falls back
// to the fp32-grade path for this launch).
const __half* gemm_half_weights(const float* W, size_t n, cudaStream_t st) {
  auto& reg = half_registry();
  auto it = reg.find(W);
  if (it != reg.end() && it->second.n >= n) return it->second.data;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  if (it != reg.end()) { cudaFree(it->second.data); reg.erase(it); }
  __half* d = nullptr;
  SV_CUDA(cudaMalloc(&d, (n + 8) * sizeof(__half)));
  f32_to_f16_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(W, d, n);
  SV_CUDA(cudaGetLastError());
  SV_CUDA(cudaStreamSynchronize(st));       // first use only: the copy is complete before any other stream can see it
  reg.emplace(W, HalfCopy{d, n});
  return d;
}

// caller-owned weights (svanon_debug_gemm*): converted on every call into a grow-only scratch, never cached
const __half* gemm_half_scratch(const float* W, size_t n, cudaStream_t st) {
  static __half* buf = nullptr;
  static size_t cap = 0;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  if (n + 8 > cap) {
    SV_CUDA(cudaDeviceSynchronize());
    if (buf) cudaFree(buf);
    buf = nullptr;
    SV_CUDA(cudaMalloc(&buf, (n + 8) * sizeof(__half)));
    cap = n + 8;
  }
  f32_to_f16_kernel<<<(unsigned)((n / 4 + 255) / 256 + 1), 256, 0, st>>>(W, buf, n);
  SV_CUDA(cudaGetLastError());
  return buf;
}

// ---- pre-tiled weights for the TMA B path.  Layout per term (hi, lo, or the single fp16 term): [tap][K-slab][n_pad rows]
// [128 bytes], a row = the slab's K elements of output channel n (32 floats / 64 halves) with its eight 16-byte chunks
// XOR-swizzled by (n & 7) -- exactly the bytes a tile row holds in shared memory -- zero-filled past N and past K.
namespace {
template <bool HALF>
__global__ void tile_weights_kernel(const float* __restrict__ W, unsigned char* __restrict__ t0, unsigned char* __restrict__ t1,
                                    int taps, int N, int K, int n_pad, int kSlabs) {
  constexpr int TKE = HALF ? 64 : 32, E = HALF ? 8 : 4;       // K elements per slab / per 16-byte chunk
  const long long chunk = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)taps * kSlabs * n_pad * 8;
  if (chunk >= total) return;
  const int c = (int)(chunk & 7);
  const long long row = chunk >> 3;                            // (t * kSlabs + s) * n_pad + n
  const int n = (int)(row % n_pad);
  const long long ts = row / n_pad;
  const int s = (int)(ts % kSlabs), t = (int)(ts / kSlabs);
  const int k = s * TKE + c * E;
  float v[8];
#pragma unroll
  for (int i = 0; i < E; ++i) v[i] = (n < N && k + i < K) ? W[((long long)t * N + n) * K + k + i] : 0.f;
  const long long dst = row * 128 + ((c ^ (n & 7)) << 4);
  if (HALF) {
    __half2 h[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(t0 + dst) = *reinterpret_cast<const uint4*>(h);
  } else {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xFFFFE000u); lo.x = v[0] - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xFFFFE000u); lo.y = v[1] - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xFFFFE000u); lo.z = v[2] - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xFFFFE000u); lo.w = v[3] - hi.w;
    *reinterpret_cast<float4*>(t0 + dst) = hi;
    *reinterpret_cast<float4*>(t1 + dst) = lo;
  }
}
struct TiledCopy { unsigned char *t0, *t1; int n_pad; size_t n; };
std::unordered_map<const float*, TiledCopy>& tiled_registry(bool half) {
  static std::unordered_map<const float*, TiledCopy> r[2];
  return r[half ? 1 : 0];
}
}  // namespace

// Pre-tiled copy of an (immutable) engine weight, made on first use; false while a stream capture is running or when the
// feature is off (the launch then takes the register path).
bool gemm_tiled_weights(const float* W, int taps, int N, int K, bool half, cudaStream_t st, const void** t0, const void** t1,
                        int* n_pad_out) {
  auto& reg = tiled_registry(half);
  const size_t n_el = (size_t)taps * N * K;
  auto it = reg.find(W);
  if (it != reg.end() && it->second.n == n_el) {
    *t0 = it->second.t0; *t1 = it->second.t1; *n_pad_out = it->second.n_pad;
    return true;
  }
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;
  }
  if (it != reg.end()) { cudaFree(it->second.t0); if (it->second.t1) cudaFree(it->second.t1); reg.erase(it); }
  const int tke = half ? 64 : 32;
  const int kSlabs = (K + tke - 1) / tke;
  const int n_pad = (N + 63) / 64 * 64;
  const size_t bytes = (size_t)taps * kSlabs * n_pad * 128;
  TiledCopy tc{nullptr, nullptr, n_pad, n_el};
  SV_CUDA(cudaMalloc(&tc.t0, bytes));
  if (!half) SV_CUDA(cudaMalloc(&tc.t1, bytes));
  const long long chunks = (long long)taps * kSlabs * n_pad * 8;
  const unsigned blocks = (unsigned)((chunks + 255) / 256);
  if (half) tile_weights_kernel<true><<<blocks, 256, 0, st>>>(W, tc.t0, tc.t1, taps, N, K, n_pad, kSlabs);
  else tile_weights_kernel<false><<<blocks, 256, 0, st>>>(W, tc.t0, tc.t1, taps, N, K, n_pad, kSlabs);
  SV_CUDA(cudaGetLastError());
  SV_CUDA(cudaStreamSynchronize(st));       // first use only
  reg.emplace(W, tc);
  *t0 = tc.t0; *t1 = tc.t1; *n_pad_out = n_pad;
  return true;
}

// drops the cached copies of one weight pointer (test hook: svanon_debug_gemm with static-weight treatment)
void gemm_forget_weights(const float* W) {
  cudaDeviceSynchronize();
  gemm_pair_forget_weights(W);
  auto ih = half_registry().find(W);
  if (ih != half_registry().end()) { cudaFree(ih->second.data); half_registry().erase(ih); }
  for (int h = 0; h < 2; ++h) {
    auto& reg = tiled_registry(h != 0);
    auto it = reg.find(W);
    if (it != reg.end()) { cudaFree(it->second.t0); if (it->second.t1) cudaFree(it->second.t1); reg.erase(it); }
  }
}

void gemm_half_release() {
  gemm_pair_release();
  for (auto& kv : half_registry()) cudaFree(kv.second.data);
  half_registry().clear();
  for (int h = 0; h < 2; ++h) {
    for (auto& kv : tiled_registry(h != 0)) { cudaFree(kv.second.t0); if (kv.second.t1) cudaFree(kv.second.t1); }
    tiled_registry(h != 0).clear();
  }
}

// Returns false when the problem is not a good fit (M < 32: latency kernels; N < 64).  Defaults measured on the streaming
// loop: M >= 32 (3.79 -> 3.72 ms per chunk vs M >= 96), split-K clusters of at most 4 (8 is no faster).
bool launch_gemm_tc(const GemmParams* ps, int count, cudaStream_t st) {
  const GemmParams& p = ps[0];
  static const int wprefetch = [] {
    const char* e = getenv("SVANON_TC_WPREFETCH");          // tuning knob: 0 = no weight-slice L2 prefetch
    return (e && atoi(e) == 0) ? 0 : 1;
  }();
  static const int min_m = [] {
    const char* e = getenv("SVANON_TC_MIN_M");          // tuning knob: smallest M that goes to the tensor cores
    return e ? atoi(e) : 32;
  }();
  // thin outputs (N = 16 / 32: the two lowest HiFi-GAN levels): a 128 x N tile with the B operand by TMA exists and is
  // correct (tests/test_gpu_gemm.py), but measured SLOWER than the shared-memory direct conv kernel (conv_small.cu) at 128
  // streams -- V 7.41 vs 6.61 ms per step, profiles/r2h_batch128_*.json -- so conv_small keeps precedence in launch_gemm (gemm.cu); this
  // tile only takes the thin shapes conv_small declines (SVANON_CONV_SMALL_MAX_M=8191 routes the vocoder levels here for A/B)
  const bool thin = (p.N == 16 || p.N == 32) && p.M >= 8192;
  if (p.M < min_m || (p.N < 64 && !thin)) return false;
  TcBatch b;
  // Weight-slice L2 prefetch pays where a GEMM is one latency-bound wave of CTAs (single stream: 3.54 -> 3.47 ms per chunk);
  // on multi-wave grids the extra L2 fill traffic costs more than it hides (128 streams: 34.2 -> 35.5 ms per step).
  b.wprefetch = (wprefetch && (long long)((p.M + TBM - 1) / TBM) * ((p.N + 63) / 64) * count <= 148) ? 1 : 0;
  int min_slabs = 1 << 30;
  for (int i = 0; i < count; ++i) {
    b.p[i] = ps[i];
    min_slabs = std::min(min_slabs, (ps[i].K + TK - 1) / TK * ps[i].taps);
  }
  for (int i = count; i < 3; ++i) b.p[i] = ps[0];
  auto ctas = [&](int bn) { return (long long)((p.M + TBM - 1) / TBM) * ((p.N + bn - 1) / bn) * count; };
  static const int max_split = [] {
    const char* e = getenv("SVANON_TC_MAX_SPLIT");      // tuning knob: largest split-K cluster
    return e ? atoi(e) : 4;
  }();
  auto pick_split = [&](long long n_ctas) {
    int s = 1;
    while (s < max_split && n_ctas * s * 2 <= 160 && min_slabs / (s * 2) >= 2) s *= 2;
    return s;
  };
  static const int max_bn = [] {
    const char* e = getenv("SVANON_TC_MAX_BN");         // tuning knob: widest CTA tile (64 / 128 / 256)
    return e ? atoi(e) : 256;
  }();
  // The producers' load/store pipe is the busiest unit (hi/lo tile stores + operand loads, ncu: L1/TEX 51 %, tensor pipe
  // 26 % on the 128 x 128 tile) and every tile pays a fixed prologue/epilogue: a wider tile amortises both over more MMA
  // work.  128 x 256 when the grid still fills the GPU and the padded width does not waste more than 128 x 128 tiles
  // would.  (Switching the weight path off gains 11-16 % on large shapes; the A operand in tensor memory and a persistent
  // kernel with dedicated epilogue warps were measured slower -- profiles/README.md.)
  auto padded = [&](int bn) { return (long long)((p.N + bn - 1) / bn) * bn; };
  // perf mode (svanon_set_precision 1): every problem of the batch needs its fp16 weight copy; K slabs hold 64 elements
  static const bool tma_weights = [] {
    const char* e = getenv("SVANON_TC_TMA_WEIGHTS");     // 1 (default): B operand by bulk copies from pre-tiled weights; 0: register path
    return !e || atoi(e) != 0;
  }();
  bool half = g_gemm_half;
  bool all_static = tma_weights;
  for (int i = 0; i < count; ++i) all_static = all_static && ps[i].w_static;
  if (half && all_static) {          // fp16, pre-tiled
    bool ok = true;
    for (int i = 0; i < count && ok; ++i)
      ok = gemm_tiled_weights(ps[i].W, ps[i].taps, ps[i].N, ps[i].K, true, st, &b.p[i].Wt[0], &b.p[i].Wt[1], &b.p[i].wt_npad);
    if (!ok) for (int i = 0; i < count; ++i) b.p[i].Wt[0] = b.p[i].Wt[1] = nullptr;
    for (int i = count; i < 3; ++i) { b.p[i].Wt[0] = b.p[0].Wt[0]; b.p[i].Wt[1] = b.p[0].Wt[1]; b.p[i].wt_npad = b.p[0].wt_npad; }
    half = ok;
    if (!ok) all_static = false;
  }
  if (half && !all_static) {
    for (int i = 0; i < count && half; ++i) {
      const size_t nw = (size_t)ps[i].taps * ps[i].N * ps[i].K;
      b.p[i].Wh = ps[i].w_static ? gemm_half_weights(ps[i].W, nw, st) : (count == 1 ? gemm_half_scratch(ps[i].W, nw, st) : nullptr);
      half = b.p[i].Wh != nullptr;
    }
  }
  if (!half && all_static) {         // parity mode, pre-split + pre-tiled
    bool ok = true;
    for (int i = 0; i < count && ok; ++i)
      ok = gemm_tiled_weights(ps[i].W, ps[i].taps, ps[i].N, ps[i].K, false, st, &b.p[i].Wt[0], &b.p[i].Wt[1], &b.p[i].wt_npad);
    if (!ok) for (int i = 0; i < count; ++i) b.p[i].Wt[0] = b.p[i].Wt[1] = nullptr;
    for (int i = count; i < 3; ++i) { b.p[i].Wt[0] = b.p[0].Wt[0]; b.p[i].Wt[1] = b.p[0].Wt[1]; b.p[i].wt_npad = b.p[0].wt_npad; }
  }
  // (A persistent variant of the wide tiles -- 8 producer + 8 epilogue warps, two TMEM accumulators, the epilogue of tile i over
  // the MMAs of tile i + 1 -- was built and measured in round 2: correct, but SLOWER, 227 vs 211 us on 16384 x 2048 x 512 and 201
  // vs 164 us on 16384 x 512 x 2048: with half the producer warps the main loop fell from 1.28 to 1.57 us per slab, more than the
  // hidden epilogue gave back.  profiles/r2f_*; the kernel lives in git history, commit "persistent wide-tile kernel".)
  if (thin) {
    if (b.p[0].Wt[0] == nullptr) return false;          // thin tiles need the pre-tiled weights (TMA B path)
    if (half) { if (p.N == 32) launch_tc_cfg<32, 4, true>(b, count, 1, st); else launch_tc_cfg<16, 4, true>(b, count, 1, st); }
    else { if (p.N == 32) launch_tc_cfg<32, 4>(b, count, 1, st); else launch_tc_cfg<16, 4>(b, count, 1, st); }
    return true;
  }
  if (half) {
    min_slabs = 1 << 30;
    for (int i = 0; i < count; ++i) min_slabs = std::min(min_slabs, (ps[i].K + 2 * TK - 1) / (2 * TK) * ps[i].taps);
    if (max_bn >= 256 && ctas(256) >= 120 && p.N >= 256 && padded(256) * 100 <= padded(128) * 107)
      launch_tc_cfg<256, 4, true>(b, count, 1, st);
    else if (max_bn >= 128 && ctas(128) >= 120 && p.N >= 128) launch_tc_cfg<128, 4, true>(b, count, 1, st);
    else launch_tc_cfg<64, 4, true>(b, count, pick_split(ctas(64)), st);
    return true;
  }
  if (max_bn >= 256 && ctas(256) >= 120 && p.N >= 256 && padded(256) * 100 <= padded(128) * 107)
    launch_tc_cfg<256, 2>(b, count, 1, st);
  else if (max_bn >= 128 && ctas(128) >= 120 && p.N >= 128) launch_tc_cfg<128, 3>(b, count, 1, st);
  else launch_tc_cfg<64, 4>(b, count, pick_split(ctas(64)), st);
  return true;
}

}  // namespace svanon
