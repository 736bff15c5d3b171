// Wide GEMMs of the many-stream batches (M = thousands of rows: 128 streams x 128 window tokens, 128 x 164 conv-stack rows) on
// CTA PAIRS: tcgen05.mma.cta_group::2, 256 x BN tile per pair, both operands by tensor-map TMA, two TMEM accumulators.
//
// Same arithmetic as gemm_tc.cu (fp32-grade products through the 3xTF32 split, same term order per K-step, same epilogue
// order), different data path.  gemm_tc.cu streams A global -> registers -> hi/lo split -> st.shared with 16 producer warps:
// at large M its main loop runs within 15 % of the MMA rate, but every 128 x 256 tile pays prologue, first-load latency,
// accumulator drain and epilogue serially (33 us on the SM for 18 us of MMAs at K = 512, profiles/README.md), and a persistent
// variant could not hide the epilogue because the producer warps need the registers and the issue slots.  Here nothing
// passes through registers on the way in:
//   * hi terms are the RAW fp32 arrays: kind::tf32 reads the upper 19 bits of each 32-bit container, i.e. it truncates exactly
//     like the explicit `x & 0xFFFFE000` of gemm_tc.cu (held bit-for-bit by tests/test_gpu_gemm.py::test_pair_gemm_bitwise);
//   * lo terms (x - trunc(x), exact in fp32) exist as arrays of their own: the weights' once per weight (registry), the
//     activations' written by whoever produced the activation (GemmParams::Alo) or by one split pass in front of the GEMM;
//   * one warp per CTA issues cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 32 floats x 128 rows: the K-major UMMA layout)
//     for its 128 rows of A (hi, lo) and its HALF of the B tile (hi, lo) -- the pair shares B through cta_group::2, so a
//     128 x 256 x 32 slab of MMA work costs each SM 64 KB of L2 reads instead of 96 KB (the L2 fabric caps at ~43 B/clk/SM,
//     B300_MICROARCH.md; 96 KB per 1.12 us slab would sit exactly on it) and three 64 KB stages fit beside the epilogue;
//   * the loads of both CTAs complete on the LEADER's `full` barrier (.cta_group::2 form of the tensor copy), the leader's
//     elected lane issues the slab's 12 MMAs for both SMs and commits them with .multicast::cluster to both CTAs' `empty`
//     barriers;
//   * persistent: a pair walks tiles t = pair, pair + 74, ...; accumulator (t & 1) of 2 x BN TMEM columns, so the 16 epilogue
//     warps of both CTAs drain tile i (tcgen05.ld -> bias / GELU / layer-scale / residual -> global, optionally the lo term
//     of the result for the next GEMM) while the tensor cores run tile i + 1.
//   * conv form (TAPS instantiation, launch_gemm_pair_taps at the end of this file): causal convs as GEMMs over taps -- the A maps
//     are 3-D (channel, row in segment incl. left context, segment) over SiLU'd hi / lo operand arrays made by one pass per
//     distinct input, tap t's rows are taken at row offset tap_off[t], its weights at a row / column offset of the weight map.
// Restrictions (everything else stays on gemm_tc.cu): unit row step, N % 128 == 0 (conv form: also 64), K (per tap) % 32 == 0,
// no accumulate, engine-owned weights, parity mode, M >= 4096; the plain form also wants contiguous A rows and no prologue.
#include <cooperative_groups.h>
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <unordered_map>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace svanon {

namespace {

constexpr int PK = 32;                  // K-slab = one 128-byte swizzle row
constexpr int PBM = 128;                // rows per CTA (the pair's tile has 256)
constexpr int P_EPI_WARPS = 16;
constexpr int P_WARP_TMA = 16, P_WARP_MMA = 17;
constexpr int P_THREADS = 18 * 32;

struct alignas(64) PairProblem {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
  CUtensorMap b2_hi, b2_lo;             // SwiGLU form: the second weight matrix (boxes of 64 rows, like b_hi / b_lo in that form)
  GemmParams p;
  float* C_lo;                          // optional: lo term of the result, [M][N] compact
  // conv form (TAPS instantiation): A maps are 3-D (channel, row in segment incl. `reach` rows of left context, segment); K-slab s
  // belongs to tap s / spt, whose A rows start toff[tap] (>= 0, relative to the first context row) and whose weights start
  // b_col_tap * tap columns / b_row_tap * tap rows into the 2-D weight map
  int k_slabs, spt, a_seg_rows, a_box_rows, b_col_tap, b_row_tap;
  int toff[MAX_TAPS];
};
struct PairArgs {
  PairProblem prob[3];
  int count, tiles_m, tiles_n, k_slabs;
  int dual;                             // SwiGLU form (GemmParams::W2): see the kernel
};

__device__ __forceinline__ float gelu_erf_p(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ unsigned smem_u32p(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void pmbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32p(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void pmbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PW_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra.uni PW_DONE;\n"
      "bra.uni PW_LOOP;\n"
      "PW_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ bool pelect_one() {
  unsigned pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(pred));
  return pred != 0;
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (same as gemm_tc.cu)
// ROWB = 128: SWIZZLE_128B tiles (rows of 32 floats, 8-row groups 1024 bytes apart); ROWB = 64: SWIZZLE_64B tiles (rows of 16 floats,
// 8-row groups 512 bytes apart: the 16-channel convs, whose taps are 16 floats wide)
template <unsigned ROWB = 128>
__device__ __forceinline__ unsigned long long pumma_desc(unsigned addr) {
  constexpr unsigned long long layout = ROWB == 128 ? 2ull : 4ull, sbo = 8ull * ROWB;
  return (((unsigned long long)addr & 0x3FFFFull) >> 4) | (1ull << 16) | ((sbo >> 4) << 32) | (1ull << 46) | (layout << 61);
}
// instruction descriptor: D=f32, A=B=tf32, K-major, N>>3 at [17,23), M>>4 at [24,29) with M = 256 (the pair)
__device__ __forceinline__ unsigned pumma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
}
__device__ __forceinline__ void pumma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                           unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (count 1) on the barrier at this offset in BOTH CTAs of the pair once all prior MMAs of this thread are complete
__device__ __forceinline__ void pumma_commit_both(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
// tensor copy global -> this CTA's shared memory; completion bytes go to the barrier at cluster address `bar` (the leader's)
__device__ __forceinline__ void tma_load_2d_pair(unsigned dst, const CUtensorMap* map, int c0, int c1, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(unsigned dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
__device__ __forceinline__ void ptmem_ld16(unsigned taddr, float (&v)[16]) {
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

template <int BN, bool TAPS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(P_THREADS, 1) gemm_pair_kernel(const __grid_constant__ PairArgs args) {
  constexpr int BH = BN / 2;                                  // B rows (output columns) this CTA stages
  constexpr int KB = BN == 16 ? 16 : PK;                      // floats per K-slab (BN = 16: the 16-channel convs, 64-byte swizzle rows)
  constexpr unsigned ROWB = KB * 4u;
  constexpr unsigned A_BYTES = PBM * ROWB, B_BYTES = BH * ROWB;
  constexpr unsigned STAGE_BYTES = 2u * A_BYTES + 2u * B_BYTES;      // A_hi | A_lo | B_hi | B_lo: 64 KB (BN 256) / 48 KB (BN 128)
  constexpr int STAGES = BN == 256 ? 3 : (BN == 32 ? 6 : (BN == 16 ? 8 : 4));
  constexpr int CG = BN >= 64 ? BN / 4 : 16;                  // accumulator columns per epilogue warp (BN = 32: two warp groups idle)
  extern __shared__ unsigned char dsmem_raw[];
  __shared__ unsigned long long full_bar[STAGES], empty_bar[STAGES], acc_full[2], acc_empty[2];
  __shared__ unsigned tmem_holder;
  const unsigned smem_base = (smem_u32p(dsmem_raw) + 1023u) & ~1023u;

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);     // warp-uniform for the compiler
  unsigned rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int tiles_per_problem = args.tiles_m * args.tiles_n;
  const int total_tiles = tiles_per_problem * args.count;
  const int k_slabs = args.k_slabs;

  pdl_trigger();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { pmbar_init(&full_bar[s], 1); pmbar_init(&empty_bar[s], 1); }
#pragma unroll
    for (int a = 0; a < 2; ++a) { pmbar_init(&acc_full[a], 1); pmbar_init(&acc_empty[a], 2 * P_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32p(&tmem_holder)), "r"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cg::this_cluster().sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_holder;
  pdl_wait();

  if (warp == P_WARP_TMA) {
    // ===================================================== loads (both CTAs): own A rows, own half of the B tile
    unsigned stage = 0, parity = 1;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int z = tile / tiles_per_problem, r = tile - z * tiles_per_problem;
      const int mt = r / args.tiles_n, nt = r - mt * args.tiles_n;
      const PairProblem& pr = args.prob[z];
      // SwiGLU form: a pair tile is 128 OUTPUT columns; this CTA's B half = 64 rows of W | the same 64 rows of W2
      const int row0 = mt * 256 + (int)rank * PBM, n0 = args.dual ? nt * (BN / 2) + (int)rank * (BH / 2) : nt * BN + (int)rank * BH;
      const int n_slabs = TAPS ? pr.k_slabs : k_slabs;
      // conv form: this CTA's 128 rows are 128 / a_box_rows runs of rows inside one segment each
      const int seg0 = TAPS ? row0 / pr.a_seg_rows : 0, r_in0 = TAPS ? row0 - seg0 * pr.a_seg_rows : 0;
      for (int s = 0; s < n_slabs; ++s) {
        pmbar_wait(smem_u32p(&empty_bar[stage]), parity);
        if (pelect_one()) {
          const unsigned dst = smem_base + stage * STAGE_BYTES;
          const unsigned bar_local = smem_u32p(&full_bar[stage]);
          const unsigned bar = mapa_u32(bar_local, 0);
          if (rank == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_local), "r"(2u * STAGE_BYTES) : "memory");
          if constexpr (TAPS) {
            const int tap = s / pr.spt, kc = (s - tap * pr.spt) * KB;
            const int roff = pr.toff[tap];
            int seg = seg0, r_in = r_in0;
            for (int sub = 0; sub < PBM; sub += pr.a_box_rows) {
              tma_load_3d_pair(dst + (unsigned)sub * ROWB, &pr.a_hi, kc, r_in + roff, seg, bar);
              tma_load_3d_pair(dst + A_BYTES + (unsigned)sub * ROWB, &pr.a_lo, kc, r_in + roff, seg, bar);
              r_in += pr.a_box_rows;
              if (r_in >= pr.a_seg_rows) { r_in = 0; ++seg; }
            }
            tma_load_2d_pair(dst + 2u * A_BYTES, &pr.b_hi, kc + tap * pr.b_col_tap, n0 + tap * pr.b_row_tap, bar);
            tma_load_2d_pair(dst + 2u * A_BYTES + B_BYTES, &pr.b_lo, kc + tap * pr.b_col_tap, n0 + tap * pr.b_row_tap, bar);
          } else {
            tma_load_2d_pair(dst, &pr.a_hi, s * PK, row0, bar);
            tma_load_2d_pair(dst + A_BYTES, &pr.a_lo, s * PK, row0, bar);
            tma_load_2d_pair(dst + 2u * A_BYTES, &pr.b_hi, s * PK, n0, bar);
            tma_load_2d_pair(dst + 2u * A_BYTES + B_BYTES, &pr.b_lo, s * PK, n0, bar);
            if (args.dual) {
              tma_load_2d_pair(dst + 2u * A_BYTES + B_BYTES / 2u, &pr.b2_hi, s * PK, n0, bar);
              tma_load_2d_pair(dst + 2u * A_BYTES + B_BYTES + B_BYTES / 2u, &pr.b2_lo, s * PK, n0, bar);
            }
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; parity ^= 1u; }
      }
    }
    // tail: the last multicast commits have landed on this CTA's `empty` barriers before it may exit
    for (int s = 0; s < STAGES; ++s) {
      pmbar_wait(smem_u32p(&empty_bar[stage]), parity);
      if (++stage == STAGES) { stage = 0; parity ^= 1u; }
    }
  } else if (warp == P_WARP_MMA) {
    // ===================================================== MMA issue (leader CTA only), one elected lane, uniform registers
    if (rank == 0) {
      const unsigned tmem_u = (unsigned)__shfl_sync(0xffffffffu, (int)tmem_base, 0);
      const unsigned idesc = pumma_idesc(BN);
      unsigned stage = 0, parity = 0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const unsigned acc = (unsigned)it & 1u;
        pmbar_wait(smem_u32p(&acc_empty[acc]), (((unsigned)it >> 1) & 1u) ^ 1u);     // both CTAs drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d_tmem = tmem_u + acc * BN;
        const int n_slabs = TAPS ? args.prob[tile / tiles_per_problem].k_slabs : k_slabs;
        for (int s = 0; s < n_slabs; ++s) {
          pmbar_wait(smem_u32p(&full_bar[stage]), parity);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (pelect_one()) {
            const unsigned a0 = smem_base + stage * STAGE_BYTES;
            const unsigned long long a_hi = pumma_desc<ROWB>(a0), a_lo = pumma_desc<ROWB>(a0 + A_BYTES);
            const unsigned long long b_hi = pumma_desc<ROWB>(a0 + 2u * A_BYTES), b_lo = pumma_desc<ROWB>(a0 + 2u * A_BYTES + B_BYTES);
#pragma unroll
            for (int kk = 0; kk < KB / 8; ++kk) {
              const unsigned long long adv = (unsigned long long)(kk * 32 >> 4);   // 8 tf32 = 32 bytes per K-step
              pumma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, (s > 0 || kk > 0) ? 1u : 0u);
              pumma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
              pumma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
            }
            pumma_commit_both(smem_u32p(&empty_bar[stage]));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; parity ^= 1u; }
        }
        if (pelect_one()) pumma_commit_both(smem_u32p(&acc_full[acc]));
        __syncwarp();
      }
    }
  } else {
    // ===================================================== epilogue (both CTAs): TMEM lanes = this CTA's 128 rows, all BN columns
    const int quad = warp & 3, grp = warp >> 2;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int z = tile / tiles_per_problem, r = tile - z * tiles_per_problem;
      const int mt = r / args.tiles_n, nt = r - mt * args.tiles_n;
      const PairProblem& pr = args.prob[z];
      const GemmParams& p = pr.p;
      const unsigned acc = (unsigned)it & 1u;
      if (lane == 0) pmbar_wait(smem_u32p(&acc_full[acc]), ((unsigned)it >> 1) & 1u);
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int m = mt * 256 + (int)rank * PBM + quad * 32 + lane;
      const bool row_ok = m < p.M;
      const long long c_row = row_ok ? gemm_c_row(p, m) : 0;
      const long long r_row = (row_ok && p.residual) ? gemm_r_row(p, m) : 0;
      const int act = p.act;
      const float out_scale = p.out_scale;
      if (grp * CG >= BN) {
        // BN = 32: this warp group has no columns; it hands the accumulator back at once
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0)
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32p(&acc_empty[acc]), 0)) : "memory");
        continue;
      }
      if (args.dual) {
        // accumulator columns: [0,64) = A W^T for output columns 0..63 of the tile, [64,128) = A W2^T for the same columns,
        // [128,192) / [192,256) the same for output columns 64..127 (the peer CTA's B half).  Warp group g owns output
        // columns [32 g, 32 g + 32).
        const unsigned col1 = (unsigned)((grp >> 1) * (BN / 2) + (grp & 1) * 32);
#pragma unroll 1
        for (int c0 = 0; c0 < 32; c0 += 16) {
          float v1[16], v3[16];
          ptmem_ld16(tmem_base + acc * BN + ((unsigned)(quad * 32) << 16) + col1 + (unsigned)c0, v1);
          ptmem_ld16(tmem_base + acc * BN + ((unsigned)(quad * 32) << 16) + col1 + (unsigned)(BN / 4 + c0), v3);
          if (c0 == 16) {
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0)
              asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32p(&acc_empty[acc]), 0)) : "memory");
          }
          if (!row_ok) continue;
          const int n_base = nt * (BN / 2) + grp * 32 + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            float o[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float a = v1[j + e], b = v3[j + e];
              o[e] = (a / (1.f + expf(-a))) * b;                       // silu_mul_kernel's expression
            }
            *reinterpret_cast<float4*>(p.C + c_row + n_base + j) = make_float4(o[0], o[1], o[2], o[3]);
            if (pr.C_lo)
              *reinterpret_cast<float4*>(pr.C_lo + (long long)m * p.N + n_base + j) =
                  make_float4(tf32_lo(o[0]), tf32_lo(o[1]), tf32_lo(o[2]), tf32_lo(o[3]));
          }
        }
        continue;
      }
      const float* rope = p.rope_table;
      const long long rope_row = rope ? (long long)(p.rope_pos0 + (p.rope_seg_rows > 0 ? m % p.rope_seg_rows : m)) * HEAD_DIM : 0;
#pragma unroll 1
      for (int c0 = grp * CG; c0 < (grp + 1) * CG; c0 += 16) {
        float v[16];
        ptmem_ld16(tmem_base + acc * BN + ((unsigned)(quad * 32) << 16) + (unsigned)c0, v);
        if (c0 + 16 == (grp + 1) * CG) {
          // the accumulator is in registers: hand it back to the MMA warp of the leader
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0)
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(smem_u32p(&acc_empty[acc]), 0)) : "memory");
        }
        if (!row_ok) continue;
        const int n_base = nt * BN + c0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          const float4 b4 = p.bias ? __ldg(reinterpret_cast<const float4*>(p.bias + n_base + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 g4 = p.gamma ? __ldg(reinterpret_cast<const float4*>(p.gamma + n_base + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
          float y[4] = {v[j] + b4.x, v[j + 1] + b4.y, v[j + 2] + b4.z, v[j + 3] + b4.w};
          const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (act == ACT_GELU) y[e] = gelu_erf_p(y[e]);
            else if (act == ACT_LOGCLAMP) y[e] = logf(fmaxf(y[e], 1e-5f));
            y[e] *= g[e];
          }
          float4 o = make_float4(y[0], y[1], y[2], y[3]);
          if (p.residual) {
            const float4 r4 = __ldg(reinterpret_cast<const float4*>(p.residual + r_row + n_base + j));
            o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
          }
          o.x *= out_scale; o.y *= out_scale; o.z *= out_scale; o.w *= out_scale;
          if (rope && n_base < p.rope_cols) {
            // rope_qk_kernel's arithmetic on the pairs (2i, 2i + 1) of a head; table entry [pos][i] = (cos, sin)
            const float4 cs = __ldg(reinterpret_cast<const float4*>(rope + rope_row + ((n_base + j) & (HEAD_DIM - 1))));
            const float x0 = o.x, x1 = o.y, x2 = o.z, x3 = o.w;
            rope_pair(x0, x1, cs.x, cs.y, o.x, o.y);
            rope_pair(x2, x3, cs.z, cs.w, o.z, o.w);
          }
          *reinterpret_cast<float4*>(p.C + c_row + n_base + j) = o;
          if (pr.C_lo)
            *reinterpret_cast<float4*>(pr.C_lo + (long long)m * p.N + n_base + j) =
                make_float4(tf32_lo(o.x), tf32_lo(o.y), tf32_lo(o.z), tf32_lo(o.w));
        }
      }
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cg::this_cluster().sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * BN) : "memory");
}

// lo = x - trunc_tf32(x) of a [M][K] array with row pitch lda -> compact [M][K]; optionally the masked hi term as well
__global__ void split_lo_kernel(const float* __restrict__ A, long long lda, float* __restrict__ lo, float* __restrict__ hi, long long M, int K) {
  pdl_trigger();
  pdl_wait();
  const int k4 = K >> 2;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * k4) return;
  const long long m = i / k4;
  const int k = (int)(i - m * k4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(A + m * lda + k);
  float4 h, l;
  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
  *reinterpret_cast<float4*>(lo + m * K + k) = l;
  if (hi) *reinterpret_cast<float4*>(hi + m * K + k) = h;
}

// ---- host side
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    SV_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres));
    SV_CHECK(f != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available in this driver");
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

struct MapKey {
  const void* ptr; long long rows, ld; int K, box_rows;      // box_rows carries the slab width too (+ 1 << 20 for 16-float slabs)
  bool operator==(const MapKey& o) const { return ptr == o.ptr && rows == o.rows && ld == o.ld && K == o.K && box_rows == o.box_rows; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h = h * 1000003u ^ std::hash<long long>()(k.rows);
    h = h * 1000003u ^ std::hash<long long>()(k.ld);
    h = h * 1000003u ^ (size_t)(k.K * 131 + k.box_rows);
    return h;
  }
};
// [rows][K] fp32, row pitch ld floats -> boxes of 32 floats x box_rows rows, 128-byte swizzle, zero fill out of bounds
const CUtensorMap& tensor_map_2d(const float* base, long long rows, int K, long long ld, int box_rows, int kb = PK) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  const MapKey key{base, rows, ld, K, box_rows + (kb == PK ? 0 : (1 << 20))};
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  if (cache.size() > 4096) cache.clear();
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, kb == PK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  return cache.emplace(key, m).first->second;
}

// lo (and, for the A/B switch, masked hi) copies of engine weights, made once per weight pointer
struct LoCopy { float* lo; float* hi; size_t n; };
std::unordered_map<const float*, LoCopy>& lo_registry() {
  static std::unordered_map<const float*, LoCopy> r;
  return r;
}
bool capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return true;
  }
  return false;
}
void launch_split(const float* A, long long lda, float* lo, float* hi, long long M, int K, cudaStream_t st) {
  const long long n4 = M * (K / 4);
  launch_pdl(split_lo_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, st, A, lda, lo, hi, M, K);
  SV_LAUNCHED();
}
const LoCopy* weight_lo(const float* W, int N, int K, bool want_hi, cudaStream_t st) {
  auto& reg = lo_registry();
  const size_t n = (size_t)N * K;
  auto it = reg.find(W);
  if (it != reg.end() && it->second.n == n && (!want_hi || it->second.hi)) return &it->second;
  if (capturing(st)) return nullptr;
  if (it != reg.end()) { cudaFree(it->second.lo); if (it->second.hi) cudaFree(it->second.hi); reg.erase(it); }
  LoCopy c{nullptr, nullptr, n};
  SV_CUDA(cudaMalloc(&c.lo, n * sizeof(float)));
  if (want_hi) SV_CUDA(cudaMalloc(&c.hi, n * sizeof(float)));
  launch_split(W, K, c.lo, c.hi, N, K, st);
  SV_CUDA(cudaStreamSynchronize(st));      // first use only
  return &reg.emplace(W, c).first->second;
}

// grow-only scratch for the activation lo terms nobody produced (two slots: the two problems of a launch may have different A)
struct Scratch { float* p = nullptr; size_t cap = 0; };
float* scratch_floats(Scratch& s, size_t n, cudaStream_t st) {
  if (n <= s.cap) return s.p;
  if (capturing(st)) return nullptr;
  SV_CUDA(cudaDeviceSynchronize());
  if (s.p) cudaFree(s.p);
  s.p = nullptr; s.cap = 0;
  SV_CUDA(cudaMalloc(&s.p, n * sizeof(float)));
  s.cap = n;
  return s.p;
}
struct ScratchSet { Scratch lo[2], hi[2]; };
std::unordered_map<cudaStream_t, ScratchSet>& scratch_registry() {      // per stream: launches on one stream are ordered
  static std::unordered_map<cudaStream_t, ScratchSet> r;
  return r;
}

// [nseg][rows][C] fp32 compact -> boxes of 32 floats x box_rows rows of one segment, 128-byte swizzle, zero fill out of bounds
CUtensorMap tensor_map_3d(const float* base, int C, int rows, int nseg, int box_rows, int kb = PK) {
  CUtensorMap m;
  const cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)nseg};
  const cuuint64_t strides[2] = {(cuuint64_t)C * sizeof(float), (cuuint64_t)rows * C * sizeof(float)};
  const cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = encode_tiled_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, kb == PK ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled (3-D) failed");
  return m;
}

// conv form: the GEMM's A rows with their left context, SiLU applied if the GEMM asks for it (the prologue gemm_tc.cu applies in its
// producers, once per TAP there), as hi (the value itself) and lo (its TF32 remainder) arrays [nseg][reach + rows][C] compact
struct ConvOperandJob { const float* A; float* hi; float* lo; long long lda, a_seg; int reach, silu; };
struct ConvOperandArgs { ConvOperandJob job[3]; int rows, C; };
// (one launch for the up to three distinct inputs of a conv launch: blockIdx.z = input)
__global__ void conv_operand_kernel(const ConvOperandArgs args) {
  pdl_trigger();
  pdl_wait();
  const ConvOperandJob& j = args.job[blockIdx.z];
  const int C = args.C, rows = args.rows, reach = j.reach;
  const int c4 = C >> 2;
  const long long per_seg = (long long)(reach + rows) * c4;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= per_seg) return;
  const int r = (int)(i / c4), c = (int)(i - (long long)r * c4) * 4;
  const long long b = blockIdx.y;
  float4 v = *reinterpret_cast<const float4*>(j.A + b * j.a_seg + (long long)(r - reach) * j.lda + c);
  if (j.silu) {
    v.x = __fdividef(v.x, 1.f + __expf(-v.x)); v.y = __fdividef(v.y, 1.f + __expf(-v.y));
    v.z = __fdividef(v.z, 1.f + __expf(-v.z)); v.w = __fdividef(v.w, 1.f + __expf(-v.w));
  }
  const long long o = (b * (reach + rows) + r) * C + c;
  *reinterpret_cast<float4*>(j.hi + o) = v;
  *reinterpret_cast<float4*>(j.lo + o) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
}
struct ConvScratch { Scratch hi[3], lo[3]; };
std::unordered_map<cudaStream_t, ConvScratch>& conv_scratch_registry() {
  static std::unordered_map<cudaStream_t, ConvScratch> r;
  return r;
}

template <int BN, bool TAPS = false>
void launch_pair_cfg(const PairArgs& a, cudaStream_t st) {
  constexpr int STAGES = BN == 256 ? 3 : (BN == 32 ? 6 : (BN == 16 ? 8 : 4));
  constexpr size_t ROWB = BN == 16 ? 64 : 128;
  constexpr size_t SMEM = (size_t)STAGES * (2 * PBM * ROWB + 2 * (BN / 2) * ROWB) + 1024;
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(gemm_pair_kernel<BN, TAPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM));
    configured = true;
  }
  static const int n_sm = [] {
    int dev = 0, n = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n;
  }();
  const int total = a.tiles_m * a.tiles_n * a.count;
  const int pairs = std::min(n_sm / 2, total);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SV_CUDA(cudaLaunchKernelEx(&cfg, gemm_pair_kernel<BN, TAPS>, a));
}

}  // namespace

bool g_gemm_pair_allowed = true;         // cleared by svanon_set_gemm_mode(< 2): the CUDA-core modes take every GEMM
long long g_gemm_pair_launches = 0;      // launches this kernel took (svanon_gemm_pair_launches: tests check the path was taken)
int g_gemm_pair_mode = -1;      // -1: environment (SVANON_GEMM_PAIR, default on); 0 off; 1 on; 2 on with explicitly masked hi copies

void gemm_pair_release() {
  cudaDeviceSynchronize();
  for (auto& kv : lo_registry()) { cudaFree(kv.second.lo); if (kv.second.hi) cudaFree(kv.second.hi); }
  lo_registry().clear();
  for (auto& kv : scratch_registry())
    for (int i = 0; i < 2; ++i) {
      if (kv.second.lo[i].p) cudaFree(kv.second.lo[i].p);
      if (kv.second.hi[i].p) cudaFree(kv.second.hi[i].p);
    }
  scratch_registry().clear();
  for (auto& kv : conv_scratch_registry())
    for (int i = 0; i < 3; ++i) {
      if (kv.second.hi[i].p) cudaFree(kv.second.hi[i].p);
      if (kv.second.lo[i].p) cudaFree(kv.second.lo[i].p);
    }
  conv_scratch_registry().clear();
}
void gemm_pair_forget_weights(const float* W) {
  auto& reg = lo_registry();
  auto it = reg.find(W);
  if (it == reg.end()) return;
  cudaDeviceSynchronize();
  cudaFree(it->second.lo);
  if (it->second.hi) cudaFree(it->second.hi);
  reg.erase(it);
}

namespace {
int pair_mode() {
  static const int env_mode = [] {
    const char* e = getenv("SVANON_GEMM_PAIR");          // 0: off; 1 (default): on; 2: on, hi terms from explicitly masked copies
    return e ? atoi(e) : 1;
  }();
  return g_gemm_pair_mode >= 0 ? g_gemm_pair_mode : env_mode;
}
}  // namespace

// Would launch_gemm hand this launch to the pair kernel?  Callers use it to decide whether producing the lo term of an
// activation (GemmParams::Alo / Clo, the `*_lo` outputs of the row-wise kernels) is worth the extra store.
bool gemm_pair_eligible(const GemmParams* ps, int count) {
  static const int min_m = [] {
    const char* e = getenv("SVANON_GEMM_PAIR_MIN_M");    // tuning knob: smallest M that takes the pair kernel
    return e ? atoi(e) : 4096;
  }();
  static const int min_tiles = [] {
    const char* e = getenv("SVANON_GEMM_PAIR_MIN_TILES"); // tuning knob: fewest 256 x BN tiles worth 74 persistent pairs
    return e ? atoi(e) : 48;
  }();
  if (pair_mode() == 0 || g_gemm_half || !g_gemm_pair_allowed || count < 1 || count > 2) return false;
  const GemmParams& p0 = ps[0];
  if (p0.M < min_m || p0.N % 128 != 0 || p0.K % PK != 0) return false;
  for (int i = 0; i < count; ++i) {
    const GemmParams& p = ps[i];
    if (p.M != p0.M || p.N != p0.N || p.K != p0.K) return false;
    if (p.W2 && (count != 1 || p.bias || p.gamma || p.residual || p.act != ACT_NONE || p.out_scale != 1.f || p.seg_rows != 0 ||
                 p.ldc != p.N || p.rope_table || (reinterpret_cast<uintptr_t>(p.W2) & 15)))
      return false;
    if (p.rope_table && (count != 1 || p.rope_cols % 16 != 0 || (reinterpret_cast<uintptr_t>(p.rope_table) & 15))) return false;
    if (p.taps != 1 || p.a_row_step != 1 || p.tap_off[0] != 0 || p.prologue != PRO_NONE || p.accumulate || !p.w_static) return false;
    if (p.seg_rows > 0 && p.a_seg != (long long)p.seg_rows * p.lda) return false;      // A rows must be one plain array
    if (p.lda % 4 != 0 || p.ldc % 4 != 0 || (p.residual && p.ldr % 4 != 0)) return false;
    if (p.seg_rows > 0 && (p.c_seg % 4 != 0 || (p.residual && p.r_seg % 4 != 0))) return false;
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    if (!al16(p.A) || !al16(p.W) || !al16(p.C) || !al16(p.bias) || !al16(p.gamma) || !al16(p.residual) || !al16(p.Alo) || !al16(p.Clo))
      return false;
  }
  const int BN = (p0.W2 || p0.N % 256 == 0) ? 256 : 128;
  const int tile_n = p0.W2 ? 128 : BN;                       // output columns per pair tile
  return (long long)((p0.M + 255) / 256) * (p0.N / tile_n) * count >= min_tiles;
}

// Returns false when the launch is not one this kernel takes (the caller continues with gemm_tc.cu).
bool launch_gemm_pair(const GemmParams* ps, int count, cudaStream_t st) {
  if (!gemm_pair_eligible(ps, count)) return false;
  const int mode = pair_mode();
  const GemmParams& p0 = ps[0];
  const bool dual = p0.W2 != nullptr;
  const int BN = (dual || p0.N % 256 == 0) ? 256 : 128;
  PairArgs a;
  a.count = count;
  a.dual = dual ? 1 : 0;
  a.tiles_m = (p0.M + 255) / 256;
  a.tiles_n = p0.N / (dual ? 128 : BN);
  a.k_slabs = p0.K / PK;
  const bool mask_hi = mode == 2;
  for (int i = 0; i < count; ++i) {
    const GemmParams& p = ps[i];
    PairProblem& pr = a.prob[i];
    pr.p = p;
    pr.C_lo = p.Clo;
    const LoCopy* w = weight_lo(p.W, p.N, p.K, mask_hi, st);
    if (!w) return false;
    // activation terms: the producer's lo array, or one split pass here (shared when both problems read the same A)
    const float* a_hi = p.A;
    long long a_hi_ld = p.lda;
    const float* a_lo = p.Alo;
    if (i == 1 && ps[1].A == ps[0].A && ps[1].lda == ps[0].lda && ps[1].Alo == ps[0].Alo) {
      pr.a_hi = a.prob[0].a_hi;
      pr.a_lo = a.prob[0].a_lo;
    } else {
      if (!a_lo || mask_hi) {
        ScratchSet& ss = scratch_registry()[st];
        float* lo = scratch_floats(ss.lo[i], (size_t)p.M * p.K, st);
        float* hi = mask_hi ? scratch_floats(ss.hi[i], (size_t)p.M * p.K, st) : nullptr;
        if (!lo || (mask_hi && !hi)) return false;
        launch_split(p.A, p.lda, lo, hi, p.M, p.K, st);
        a_lo = lo;
        if (mask_hi) { a_hi = hi; a_hi_ld = p.K; }
      }
      pr.a_hi = tensor_map_2d(a_hi, p.M, p.K, a_hi_ld, PBM);
      pr.a_lo = tensor_map_2d(a_lo, p.M, p.K, p.K, PBM);
    }
    const int b_box = dual ? BN / 4 : BN / 2;
    pr.b_hi = tensor_map_2d(mask_hi ? w->hi : p.W, p.N, p.K, p.K, b_box);
    pr.b_lo = tensor_map_2d(w->lo, p.N, p.K, p.K, b_box);
    pr.b2_hi = pr.b_hi;
    pr.b2_lo = pr.b_lo;
    if (dual) {
      const LoCopy* w2 = weight_lo(p.W2, p.N, p.K, mask_hi, st);
      if (!w2) return false;
      pr.b2_hi = tensor_map_2d(mask_hi ? w2->hi : p.W2, p.N, p.K, p.K, b_box);
      pr.b2_lo = tensor_map_2d(w2->lo, p.N, p.K, p.K, b_box);
    }
  }
  for (int i = count; i < 3; ++i) a.prob[i] = a.prob[0];
  if (BN == 256) launch_pair_cfg<256>(a, st);
  else launch_pair_cfg<128>(a, st);
  ++g_gemm_pair_launches;
  return true;
}

// ---- conv form: causal convs with left context as GEMMs over taps (the ResBlock convs of the wide HiFi-GAN levels at many streams:
// C = N = 128 / 256 channels, 3 / 7 / 11 taps, up to three problems per launch, SiLU on the input; the transposed up-sampling convs).
// One pass turns each distinct input into SiLU'd hi / lo arrays (gemm_tc.cu applies the SiLU in its producers, once per tap), the
// kernel then takes tap t's rows through a 3-D tensor map at row offset tap_off[t]; same term and K order as gemm_tc.cu: same bits.
bool launch_gemm_pair_taps(const GemmParams* ps, int count, cudaStream_t st) {
  static const bool on = [] {
    const char* e = getenv("SVANON_GEMM_PAIR_TAPS");     // 0: convs with taps stay on the single-CTA kernel
    return !e || atoi(e) != 0;
  }();
  static const int min_m = [] {
    const char* e = getenv("SVANON_GEMM_PAIR_MIN_M");
    return e ? atoi(e) : 4096;
  }();
  static const int min_tiles = [] {
    const char* e = getenv("SVANON_GEMM_PAIR_MIN_TILES");
    return e ? atoi(e) : 48;
  }();
  if (!on || pair_mode() == 0 || g_gemm_half || !g_gemm_pair_allowed || count < 1 || count > 3) return false;
  const GemmParams& p0 = ps[0];
  if (p0.M < min_m || (p0.N % 128 != 0 && p0.N != 64 && p0.N != 32 && p0.N != 16)) return false;
  const int C = (int)p0.lda;                                   // channels per tap = A row length
  const int kb = (p0.N == 16) ? 16 : PK;                       // 16 channels: K-slabs of 16 floats (64-byte swizzle rows)
  if (C % kb != 0 || C < kb || (p0.N == 16 && C != 16)) return false;
  const int seg_rows = p0.seg_rows;
  if (seg_rows > 0 && !(seg_rows % PBM == 0 || (PBM % seg_rows == 0 && seg_rows % 8 == 0))) return false;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  int ktaps[3], reach[3];
  bool any_taps = false;
  for (int i = 0; i < count; ++i) {
    const GemmParams& p = ps[i];
    if (p.M != p0.M || p.N != p0.N || p.lda != C || p.seg_rows != seg_rows) return false;
    if (p.a_row_step != 1 || p.accumulate || !p.w_static || p.W2 || p.rope_table || p.Alo || p.Clo) return false;
    if (p.prologue != PRO_NONE && p.prologue != PRO_SILU) return false;
    if (p.taps > 1 ? p.K != C : (p.K % C != 0)) return false;
    ktaps[i] = p.taps > 1 ? p.taps : p.K / C;
    if (ktaps[i] < 1 || ktaps[i] > MAX_TAPS) return false;
    int lo_off = 0;
    for (int t = 0; t < ktaps[i]; ++t) {
      const int off = p.taps > 1 ? p.tap_off[t] : p.tap_off[0] + t;
      if (off > 0) return false;
      lo_off = std::min(lo_off, off);
    }
    reach[i] = -lo_off;
    if (ktaps[i] > 1 || p.prologue == PRO_SILU || reach[i] > 0) any_taps = true;
    if (p.ldc % 4 != 0 || (p.residual && p.ldr % 4 != 0) || (seg_rows > 0 && (p.a_seg % 4 != 0 || p.c_seg % 4 != 0 || (p.residual && p.r_seg % 4 != 0))))
      return false;
    if (!al16(p.A) || !al16(p.W) || !al16(p.C) || !al16(p.bias) || !al16(p.gamma) || !al16(p.residual)) return false;
  }
  if (!any_taps) return false;                                 // plain GEMMs take the plain path
  const int BN = (p0.N % 256 == 0) ? 256 : (p0.N % 128 == 0 ? 128 : p0.N);
  PairArgs a;
  a.count = count;
  a.dual = 0;
  a.tiles_m = (p0.M + 255) / 256;
  a.tiles_n = p0.N / BN;
  a.k_slabs = 0;
  if ((long long)a.tiles_m * a.tiles_n * count < min_tiles) return false;
  if (capturing(st)) return false;
  const int nseg = seg_rows > 0 ? p0.M / seg_rows : 1;
  const int rows = seg_rows > 0 ? seg_rows : p0.M;             // rows per segment
  if (seg_rows > 0 && p0.M % seg_rows != 0) return false;
  ConvScratch& cs = conv_scratch_registry()[st];
  ConvOperandArgs oa;
  oa.rows = rows; oa.C = C;
  int n_jobs = 0, max_reach = 0;
  for (int i = 0; i < count; ++i) {
    const GemmParams& p = ps[i];
    PairProblem& pr = a.prob[i];
    pr.p = p;
    pr.C_lo = nullptr;
    // weights: dilated convs [k][N][C] (tap = N rows further), dilation 1 [N][k C] (tap = C columns further)
    const int w_rows = p.taps > 1 ? ktaps[i] * p.N : p.N, w_cols = p.taps > 1 ? C : ktaps[i] * C;
    const LoCopy* w = weight_lo(p.W, w_rows, w_cols, false, st);
    if (!w) return false;
    pr.b_hi = tensor_map_2d(p.W, w_rows, w_cols, w_cols, BN / 2, kb);
    pr.b_lo = tensor_map_2d(w->lo, w_rows, w_cols, w_cols, BN / 2, kb);
    pr.b2_hi = pr.b_hi;
    pr.b2_lo = pr.b_lo;
    pr.b_row_tap = p.taps > 1 ? p.N : 0;
    pr.b_col_tap = p.taps > 1 ? 0 : C;
    pr.spt = C / kb;
    pr.k_slabs = ktaps[i] * pr.spt;
    pr.a_seg_rows = seg_rows > 0 ? seg_rows : (p0.M + 255) / 256 * 256;
    pr.a_box_rows = std::min(PBM, rows >= PBM ? PBM : rows);
    for (int t = 0; t < MAX_TAPS; ++t) pr.toff[t] = 0;
    for (int t = 0; t < ktaps[i]; ++t) pr.toff[t] = (p.taps > 1 ? p.tap_off[t] : p.tap_off[0] + t) + reach[i];
    // operand arrays: shared with an earlier problem of the launch that reads the same rows the same way
    int same = -1;
    for (int j = 0; j < i; ++j)
      if (ps[j].A == p.A && ps[j].a_seg == p.a_seg && reach[j] == reach[i] && ps[j].prologue == p.prologue) { same = j; break; }
    if (same >= 0) {
      pr.a_hi = a.prob[same].a_hi;
      pr.a_lo = a.prob[same].a_lo;
      continue;
    }
    const size_t n = (size_t)nseg * (reach[i] + rows) * C;
    float* hi = scratch_floats(cs.hi[i], n, st);
    float* lo = scratch_floats(cs.lo[i], n, st);
    if (!hi || !lo) return false;
    oa.job[n_jobs++] = ConvOperandJob{p.A, hi, lo, p.lda, seg_rows > 0 ? p.a_seg : 0LL, reach[i], p.prologue == PRO_SILU ? 1 : 0};
    max_reach = std::max(max_reach, reach[i]);
    pr.a_hi = tensor_map_3d(hi, C, reach[i] + rows, nseg, pr.a_box_rows, kb);
    pr.a_lo = tensor_map_3d(lo, C, reach[i] + rows, nseg, pr.a_box_rows, kb);
  }
  if (n_jobs > 0) {
    for (int i = n_jobs; i < 3; ++i) oa.job[i] = oa.job[0];
    const long long per_seg = (long long)(max_reach + rows) * (C / 4);
    launch_pdl(conv_operand_kernel, dim3((unsigned)((per_seg + 255) / 256), nseg, n_jobs), dim3(256), 0, st, oa);
    SV_LAUNCHED();
  }
  for (int i = count; i < 3; ++i) a.prob[i] = a.prob[0];
  for (int i = 0; i < 3; ++i) a.prob[i].p.prologue = PRO_NONE;  // applied by conv_operand_kernel
  if (BN == 256) launch_pair_cfg<256, true>(a, st);
  else if (BN == 128) launch_pair_cfg<128, true>(a, st);
  else if (BN == 64) launch_pair_cfg<64, true>(a, st);         // 64-channel level: L2-bound (40 KB of operands per 0.28 us slab)
  else if (BN == 32) launch_pair_cfg<32, true>(a, st);         // 32-channel level: L2-bound too, still ~3x the CUDA-core conv kernel
  else launch_pair_cfg<16, true>(a, st);                       // 16-channel level: 64-byte swizzle rows, K-slabs of one tap
  ++g_gemm_pair_launches;
  return true;
}

}  // namespace svanon
