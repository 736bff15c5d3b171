// Persistent chain kernel (chain.cuh): op-list interpreter with tcgen05 GEMM phases and row-wise phases, one cooperative
// launch per stage of the single-stream chunk.
#include "chain.cuh"

#include <algorithm>
#include <cstdlib>

namespace svanon {

bool gemm_tiled_weights(const float* W, int taps, int N, int K, bool half, cudaStream_t st, const void** t0, const void** t1,
                        int* n_pad_out);                       // gemm_tc.cu
bool gemm_timing_on();                                         // gemm.cu
void gemm_timing_external(cudaStream_t st, bool begin, double flop);

namespace {

// 8 worker warps + 1 control warp: 9 warps keep the register cap at 168 per thread (17 warps: 96, and the interpreter's
// long-lived state then spills -- the producers' in-flight loads and the attention's staging registers went to local memory)
constexpr int CH_WW = 8;                        // worker warps: A producers, epilogue, row-wise phases
constexpr int CH_WORKERS = CH_WW * 32;
constexpr int CH_THREADS = CH_WORKERS + 32;     // warp CH_WW: weight TMA + MMA issue (one lane), op-descriptor fetch
constexpr int CH_STAGES = 3;
constexpr int CH_A_STAGE_BYTES = 2 * 128 * 128; // hi | lo tiles of 128 rows x 128 bytes
constexpr int CH_RING_BYTES = CH_STAGES * CH_A_STAGE_BYTES;
constexpr int CH_SMEM_BYTES = CH_RING_BYTES + 2 * CHAIN_B_BYTES + 1024;
constexpr int CH_TMEM_COLS = 256;
constexpr int CH_DEPTH = 3;                     // (slab, M tile) iterations a producer thread keeps in flight in registers

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- bounded waits: a protocol bug must trap (the launch fails), never hang the GPU
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned n = 0;
  while (!mbar_try(bar, parity)) {
    if (++n > (1u << 24)) __trap();
  }
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

// Grid barrier (co-residency by the cooperative launch): monotonic arrival counter bar[0], bar[1] carries it across launches
// -- the scheme of ar_decode_common.cuh with a bounded poll.
struct GridBar {
  unsigned* bar;
  unsigned target;
};
__device__ __forceinline__ void grid_sync(GridBar& g, unsigned nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    g.target += nblocks;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(g.bar) : "memory");
    unsigned v, n = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(g.bar) : "memory");
      if (++n > (1u << 26)) __trap();
    } while ((int)(v - g.target) < 0);
  }
  __syncthreads();
}

// ---- tcgen05 (same conventions as gemm_tc.cu: K-major SWIZZLE_128B tiles, kind::tf32, M = 128)
__device__ __forceinline__ unsigned long long umma_desc(unsigned saddr) {
  return ((unsigned long long)(saddr & 0x3FFFFu) >> 4) | (1ull << 16) | ((1024ull >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ unsigned umma_idesc(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(128 >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc,
                                          unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(unsigned bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// one lane of a converged warp (the compiler keeps the warp's control flow uniform around it)
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
__device__ __forceinline__ int bcast0(int v) { return __shfl_sync(0xffffffffu, v, 0); }

template <class T>
__device__ __forceinline__ T* dynp(T* p, const ChainDyn& d) {
  const uintptr_t v = reinterpret_cast<uintptr_t>(p);
  return (v >= 1 && v <= (uintptr_t)CHAIN_DYN) ? reinterpret_cast<T*>(const_cast<void*>(d.p[v - 1])) : p;
}

// value of 4 consecutive columns of row `row` of a pending sum
__device__ __forceinline__ float4 pend4(const ChainPend& in, const float* res, long long row, int col) {
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  if (in.P) {
    const float* p = in.P + row * in.ldp + col;
#pragma unroll 8
    for (int k = 0; k < in.ks; ++k) {
      const float4 v = ldcg4(p + (long long)k * in.ks_stride);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  if (in.bias) { const float4 b = ldg4(in.bias + col); s.x += b.x; s.y += b.y; s.z += b.z; s.w += b.w; }
  if (in.gamma) { const float4 g = ldg4(in.gamma + col); s.x *= g.x; s.y *= g.y; s.z *= g.z; s.w *= g.w; }
  if (res) { const float4 r = ldcg4(res + row * in.ldr + col); s.x += r.x; s.y += r.y; s.z += r.z; s.w += r.w; }
  return s;
}

struct ChainArgs {
  const ChainOp* ops;
  const int* gemm_ops;
  const ChainWJob* wjobs;        // [n_gemm][grid]: this CTA's weight block of every GEMM (host-made: one 32-byte load)
  int n_ops, n_gemm;
  unsigned* barrier;
  unsigned long long* prof;      // tuning aid (SVANON_CHAIN_PROF): [op][cta 0 / last cta][start, work done, barrier done] in ns
  int trace_op;                  // SVANON_CHAIN_TRACE: per-slab pipeline timestamps of this GEMM op in CTA 0, after the op table
  ChainDyn dyn;
};
__device__ __forceinline__ unsigned long long gtime_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// issue the bulk copies of this CTA's weight block (descriptor `w`) into the B buffer at `bbuf` (one thread)
__device__ __forceinline__ bool row_map(const ChainOp& op, int r, long long& ro) {
  ro = r;
  if (op.period <= 0) return true;
  const int seg = r / op.period, t = r - seg * op.period - op.margin;
  if (t < 0) return false;
  if (op.y_period > 0) ro = (long long)seg * op.y_period + op.y_margin + t;
  return true;
}
__device__ __forceinline__ void prefetch_weights(const ChainWJob& w, unsigned bbuf, unsigned bar) {
  if (w.n_sl <= 0) return;
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((unsigned)w.n_sl * 2u * w.bytes) : "memory");
  for (int s = 0; s < w.n_sl; ++s) {
    const long long off = (long long)s * w.slab_stride;
    const unsigned dst = bbuf + (unsigned)s * 2u * w.bytes;
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(w.w0 + off), "r"(w.bytes), "r"(bar) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst + w.bytes), "l"(w.w1 + off), "r"(w.bytes), "r"(bar) : "memory");
  }
}

__global__ void __launch_bounds__(CH_THREADS, 1) chain_kernel(const ChainArgs a) {
  extern __shared__ unsigned char ch_smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(ch_smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ unsigned long long full_bar[CH_STAGES], empty_bar[CH_STAGES], acc_bar, bfull_bar[2], bfree_bar[2];
  __shared__ __align__(16) ChainWJob s_wj[4];      // job descriptors of the GEMMs in flight: GEMM q in slot q & 3 (cp.async targets)
  __shared__ unsigned tmem_holder;
  // the op descriptors are double-buffered: warp 16 fetches op i + 1 while phase i runs (the list is static)
  constexpr int OP_INTS = (int)(sizeof(ChainOp) / 4);
  static_assert(sizeof(ChainOp) % 16 == 0 && sizeof(ChainOp) <= 512, "ChainOp travels as one 16-byte cp.async per lane");
  __shared__ __align__(16) unsigned char s_op_raw[2][sizeof(ChainOp)];

  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform for the compiler (role branches stay uniform)
  const unsigned ring = smem_u32(smem);
  const unsigned bbase = ring + CH_RING_BYTES;
  const unsigned nblocks = gridDim.x;

  GridBar gb{a.barrier, 0};
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < CH_STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), CH_WW); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&acc_bar), 1);
    mbar_init(smem_u32(&bfull_bar[0]), 1);
    mbar_init(smem_u32(&bfull_bar[1]), 1);
    mbar_init(smem_u32(&bfree_bar[0]), 1);
    mbar_init(smem_u32(&bfree_bar[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gb.target) : "l"(a.barrier + 1) : "memory");
  }
  {
    const int* src = reinterpret_cast<const int*>(a.ops);
    int* dst = reinterpret_cast<int*>(s_op_raw[0]);
    for (int i = tid; i < OP_INTS; i += CH_THREADS) dst[i] = __ldg(src + i);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_holder)), "r"(CH_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem_base = tmem_holder;

  // pipeline state, per role (every thread of a role advances it identically)
  unsigned st_stage = 0, st_parity = 1;      // producers: wait on empty[stage]
  unsigned mm_stage = 0, mm_parity = 0;      // MMA thread: wait on full[stage]
  unsigned acc_parity = 0;                   // everyone: accumulator-ready barrier
  unsigned b_par0 = 0, b_par1 = 0;           // control warp: weight buffers (landed)
  unsigned f_par0 = 0, f_par1 = 0;           // control thread: weight buffers (MMAs that read them are done)
  const bool ctl = (warp == CH_WW && lane == 0);
  // warp-uniform copies for the control warp's MMA loop
  const unsigned ring_u = (unsigned)bcast0((int)ring), bbase_u = (unsigned)bcast0((int)bbase), tmem_u = (unsigned)bcast0((int)tmem_base);

  // the weight blocks of the first two GEMMs
  const ChainWJob* my_wjobs = a.wjobs + (size_t)blockIdx.x * a.n_gemm;      // GEMM q: my_wjobs[q]
  if (ctl) {
    // the op list and this CTA's descriptors are cold (the chunk streams ~0.8 GB through L2 between two launches): pull them
    // into L2 once, so that the per-phase descriptor fetches below are L2 hits
    l2_prefetch_bulk(a.ops, (unsigned)a.n_ops * (unsigned)sizeof(ChainOp));
    if (a.n_gemm > 0) l2_prefetch_bulk(my_wjobs, (unsigned)a.n_gemm * (unsigned)sizeof(ChainWJob));
  }
  if (warp == CH_WW && lane < 8 && (lane >> 2) < a.n_gemm)          // descriptors of the first two GEMMs
    reinterpret_cast<int4*>(s_wj)[lane] = __ldg(reinterpret_cast<const int4*>(my_wjobs) + lane);
  __syncwarp();
  if (ctl) {
    for (int q = 0; q < 2 && q < a.n_gemm; ++q)
      prefetch_weights(s_wj[q], bbase + (unsigned)q * CHAIN_B_BYTES, smem_u32(&bfull_bar[q]));
  }
  __syncthreads();

  for (int oi = 0; oi < a.n_ops; ++oi) {
    const ChainOp& op = *reinterpret_cast<const ChainOp*>(s_op_raw[oi & 1]);
    // (cp.async: no register holds the descriptor, so nothing waits for it until the end of the phase)
    const bool fetch_next = warp == CH_WW && oi + 1 < a.n_ops;
    if (fetch_next && lane * 16 < (int)sizeof(ChainOp)) {
      const unsigned dst = smem_u32(s_op_raw[(oi + 1) & 1]) + (unsigned)lane * 16u;
      const unsigned char* src = reinterpret_cast<const unsigned char*>(a.ops + oi + 1) + lane * 16;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    const bool prof = a.prof && tid == 0 && (blockIdx.x == 0 || blockIdx.x == nblocks - 1);
    unsigned long long* pslot = a.prof + (size_t)oi * 16 + (blockIdx.x == 0 ? 0 : 3);
    unsigned long long* gmark = (a.prof && blockIdx.x == 0) ? a.prof + (size_t)oi * 16 + 6 : nullptr;
    if (prof) pslot[0] = gtime_ns();
    unsigned long long* trace = (a.prof && blockIdx.x == 0 && oi == a.trace_op) ? a.prof + (size_t)a.n_ops * 16 : nullptr;

    if (op.kind == CH_GEMM) {
      // ================================================================================ GEMM phase
      const ChainWJob& wj = s_wj[op.gemm_seq & 3];
      const int n_sl = wj.n_sl;
      const bool has_job = n_sl > 0;
      const int buf = op.gemm_seq & 1;
      const bool more = op.gemm_seq + 2 < a.n_gemm;
      if (warp == CH_WW && lane < 4 && more) {     // descriptor of the next-but-one GEMM (its weight block goes into this buffer)
        const unsigned dst = smem_u32(&s_wj[(op.gemm_seq + 2) & 3]) + (unsigned)lane * 16u;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(my_wjobs + op.gemm_seq + 2) + lane * 16;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
      }
      if (has_job) {
        const int ks = wj.ks;
        const int BN = op.BN, n0 = wj.n0;
        const int m_tiles = (op.M + 127) >> 7;
        const int n_it = n_sl * m_tiles;
        if (warp < CH_WW) {
          // ---------------------------------------------------------------- producers: A rows -> hi/lo swizzled tiles
          const int c = tid & 7;
          const int row0 = tid >> 3;                       // chunk rows row0 + 32 j
          unsigned soff[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int row = row0 + 32 * j;
            soff[j] = (unsigned)(row * 128 + ((c ^ (row & 7)) << 4));
          }
          const float* abase = op.A + wj.a_off + c * 4;
          int ld_sl = 0, ld_mt = 0;
          float4 ra[CH_DEPTH][4];
          auto load = [&](float4 (&dst)[4]) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int m = ld_mt * 128 + row0 + 32 * j;
              dst[j] = (m < op.M) ? ldcg4(abase + (long long)m * op.a_row_stride + ld_sl * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (++ld_mt == m_tiles) { ld_mt = 0; ++ld_sl; }
          };
#pragma unroll
          for (int d = 0; d < CH_DEPTH; ++d)
            if (d < n_it) load(ra[d]);
          if (gmark && tid == 0) { gmark[0] = gtime_ns(); if (n_it > 0 && ra[0][0].x == 12345.678f) gmark[0] = 0; gmark[1] = gtime_ns(); }
          const unsigned full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
          for (int it0 = 0; it0 < n_it; it0 += CH_DEPTH) {
#pragma unroll
            for (int d = 0; d < CH_DEPTH; ++d) {
              const int it = it0 + d;
              if (it < n_it) {
                if (lane == 0) mbar_wait(empty0 + st_stage * 8u, st_parity);
                __syncwarp();
                if (trace && tid == 0 && it < 16) trace[it * 4] = gtime_ns();
                const unsigned a_stage = ring + st_stage * (unsigned)CH_A_STAGE_BYTES;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float4 v = ra[d][j];
                  float4 h, l;
                  h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
                  h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
                  h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
                  h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_stage + soff[j]), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
                  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(a_stage + soff[j] + 16384u), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(full0 + st_stage * 8u) : "memory");
                if (trace && tid == 0 && it < 16) trace[it * 4 + 1] = gtime_ns();
                if (++st_stage == CH_STAGES) { st_stage = 0; st_parity ^= 1u; }
                if (it + CH_DEPTH < n_it) load(ra[d]);
              }
            }
          }
        } else {
          // ---------------------------------------------------------------- MMA issue: the whole control warp walks the loop
          // (values broadcast from lane 0 are warp-uniform for the compiler: descriptors stay in uniform registers instead of a
          // per-instruction R2UR loop -- 12 tcgen05.mma per slab cost 0.55 us that way), one elected lane issues
          if (gmark && lane == 0) gmark[2] = gtime_ns();
          mbar_wait(smem_u32(&bfull_bar[buf]), buf ? b_par1 : b_par0);
          if (buf) b_par1 ^= 1u; else b_par0 ^= 1u;
          if (gmark && lane == 0) gmark[3] = gtime_ns();
          const int BNu = bcast0(BN), n_itu = bcast0(n_it), m_tilesu = bcast0(m_tiles), cont = bcast0(op.acc_cont), keep = bcast0(op.acc_keep);
          const unsigned idesc = umma_idesc(BNu);
          const unsigned bblk = bbase_u + (unsigned)bcast0(buf) * CHAIN_B_BYTES;
          const unsigned bterm = (unsigned)BNu * 128u;
          int sl = 0, mt = 0;
          for (int it = 0; it < n_itu; ++it) {
            mbar_wait(smem_u32(&full_bar[mm_stage]), mm_parity);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (trace && lane == 0 && it < 16) trace[it * 4 + 2] = gtime_ns();
            const unsigned a_stage = ring_u + mm_stage * (unsigned)CH_A_STAGE_BYTES;
            const unsigned long long a_hi = umma_desc(a_stage), a_lo = umma_desc(a_stage + 16384u);
            const unsigned long long b_hi = umma_desc(bblk + (unsigned)sl * 2u * bterm), b_lo = umma_desc(bblk + (unsigned)sl * 2u * bterm + bterm);
            const unsigned d_tmem = tmem_u + (unsigned)(mt * BNu);
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const unsigned long long adv = (unsigned long long)(kk * 2);      // 8 tf32 = 32 bytes per K-step
                umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, (sl > 0 || kk > 0 || cont) ? 1u : 0u);
                umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, 1u);
                umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              }
              umma_commit(smem_u32(&empty_bar[mm_stage]));
            }
            __syncwarp();
            if (trace && lane == 0 && it < 16) trace[it * 4 + 3] = gtime_ns();
            if (++mm_stage == CH_STAGES) { mm_stage = 0; mm_parity ^= 1u; }
            if (++mt == m_tilesu) { mt = 0; ++sl; }
          }
          if (elect_one()) {
            umma_commit(smem_u32(&bfree_bar[buf]));
            if (!keep) umma_commit(smem_u32(&acc_bar));
          }
          __syncwarp();
          if (gmark && lane == 0) gmark[4] = gtime_ns();
        }
        __syncwarp();
        // ------------------------------------------------------------------ accumulators -> result (or K-slice partial) rows
        if (warp < CH_WW && !op.acc_keep) {
          if (lane == 0) mbar_wait(smem_u32(&acc_bar), acc_parity);
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (gmark && tid == 0) gmark[5] = gtime_ns();
          const int quad = warp & 3, grp = warp >> 2;
          // A thread holds ROWS of the accumulator (TMEM lane = row): stored straight from registers, one instruction would
          // touch 32 different 128-byte lines (32 LSU wavefronts for 512 bytes -- 1 us of the phase for a 128 x 64 tile).  Every
          // warp therefore passes its 32-row block through its own slice of the (now idle) A ring and writes whole row segments.
          float* stg = reinterpret_cast<float*>(smem) + warp * (32 * 36);
          if (op.epi == EPI_SILU_MUL) {
            // the tile holds 16 columns of h1 next to the same 16 columns of h3 (interleaved weight rows): y = silu(h1) * h3
            const int pairs = BN >> 5;
            for (int mt = 0; mt < m_tiles; ++mt)
            for (int pr = grp; pr < pairs; pr += CH_WW / 4) {
              float v1[16], v3[16];
              const unsigned tad = tmem_base + ((unsigned)(quad * 32) << 16) + (unsigned)(mt * BN + pr * 32);
              tmem_ld16(tad, v1);
              tmem_ld16(tad + 16u, v3);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float4 o;
                o.x = (v1[4 * j] / (1.f + expf(-v1[4 * j]))) * v3[4 * j];
                o.y = (v1[4 * j + 1] / (1.f + expf(-v1[4 * j + 1]))) * v3[4 * j + 1];
                o.z = (v1[4 * j + 2] / (1.f + expf(-v1[4 * j + 2]))) * v3[4 * j + 2];
                o.w = (v1[4 * j + 3] / (1.f + expf(-v1[4 * j + 3]))) * v3[4 * j + 3];
                *reinterpret_cast<float4*>(stg + lane * 20 + 4 * j) = o;
              }
              __syncwarp();
              const int rr = lane >> 2, c = (lane & 3) * 4;
              const int n = (n0 >> 1) + pr * 16 + c;
#pragma unroll
              for (int r0 = 0; r0 < 32; r0 += 8) {
                const int r = r0 + rr, m = mt * 128 + quad * 32 + r;
                if (m < op.M) *reinterpret_cast<float4*>(op.y + (long long)m * op.ldy + n) = *reinterpret_cast<const float4*>(stg + r * 20 + c);
              }
              __syncwarp();
            }
          } else {
            const int cgs = BN >> 4;                         // 16-column groups per M tile
            const int W = cgs >= 4 ? 2 : 1;                  // groups per warp and pass (BN = 64: two halves of 32 columns)
            const int per_pass = W * (CH_WW / 4);            // groups the warps of one lane quadrant cover per pass
            const int SP = W * 16 + 4;
            const int lpr_sh = W == 2 ? 3 : 2;               // lanes per row segment: 8 or 4
            float* dbase = op.epi == EPI_PARTIAL ? op.Pout + (long long)ks * op.pout_ks_stride : op.y;
            const long long dld = op.epi == EPI_PARTIAL ? op.ldp_out : op.ldy;
            const float* eres = dynp(op.in.res, a.dyn);
            for (int mt = 0; mt < m_tiles; ++mt)
            for (int cb = grp * W; cb < cgs; cb += per_pass) {
              const int m = mt * 128 + quad * 32 + lane;
#pragma unroll
              for (int g = 0; g < 2; ++g) {
                if (g < W && cb + g < cgs) {
                  const int cg = cb + g;
                  float v[16];
                  tmem_ld16(tmem_base + ((unsigned)(quad * 32) << 16) + (unsigned)(mt * BN + cg * 16), v);
                  const int n = n0 + cg * 16;
                  if (op.epi != EPI_PARTIAL && m < op.M && n < op.N) {
                    // direct result: y = act(res + gamma * (acc + bias)), or RoPE on the q / k columns of a qkv row
                    if (op.in.bias) {
#pragma unroll
                      for (int j = 0; j < 16; ++j) v[j] += __ldg(op.in.bias + n + j);
                    }
                    if (op.in.gamma) {
#pragma unroll
                      for (int j = 0; j < 16; ++j) v[j] *= __ldg(op.in.gamma + n + j);
                    }
                    if (eres) {
                      const float4* rp = reinterpret_cast<const float4*>(eres + (long long)m * op.in.ldr + n);
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        const float4 r = __ldcg(rp + j);
                        v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                      }
                    }
                    if (op.epi == EPI_ROPE && n < 2 * op.heads * HEAD_DIM) {
                      const float4* cs = reinterpret_cast<const float4*>(op.table + ((long long)(op.q_first + m) * (HEAD_DIM / 2) + ((n & (HEAD_DIM - 1)) >> 1)) * 2);
#pragma unroll
                      for (int j = 0; j < 4; ++j) {
                        const float4 t = __ldg(cs + j);        // (cos, sin) of pairs 2 j and 2 j + 1
                        const float x0 = v[4 * j], x1 = v[4 * j + 1], x2 = v[4 * j + 2], x3 = v[4 * j + 3];
                        v[4 * j] = x0 * t.x - x1 * t.y; v[4 * j + 1] = x1 * t.x + x0 * t.y;
                        v[4 * j + 2] = x2 * t.z - x3 * t.w; v[4 * j + 3] = x3 * t.z + x2 * t.w;
                      }
                    }
                    if (op.act == CHA_GELU) {
#pragma unroll
                      for (int j = 0; j < 16; ++j) v[j] = gelu_erf(v[j]);
                    }
                  }
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    *reinterpret_cast<float4*>(stg + lane * SP + g * 16 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
              }
              __syncwarp();
              const int rr = lane >> lpr_sh, c = (lane & ((1 << lpr_sh) - 1)) * 4;
              const int n = n0 + cb * 16 + c;
              const int rpi = 32 >> lpr_sh;
              if (c < (min(W, cgs - cb) << 4) && n < op.N) {
                for (int r0 = 0; r0 < 32; r0 += rpi) {
                  const int r = r0 + rr, mm = mt * 128 + quad * 32 + r;
                  if (mm < op.M) *reinterpret_cast<float4*>(dbase + (long long)mm * dld + n) = *reinterpret_cast<const float4*>(stg + r * SP + c);
                }
              }
              __syncwarp();
            }
          }
        }
        if (!op.acc_keep) acc_parity ^= 1u;
      }
      if (gmark && tid == 0) gmark[6] = gtime_ns();
      // once the MMAs that read this GEMM's weight buffer are done, fetch the block of the next-but-one GEMM into it
      if (ctl) {
        if (has_job) {
          mbar_wait(smem_u32(&bfree_bar[buf]), buf ? f_par1 : f_par0);
          if (buf) f_par1 ^= 1u; else f_par0 ^= 1u;
        }
      }
      if (warp == CH_WW && more) {
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        if (lane == 0) prefetch_weights(s_wj[(op.gemm_seq + 2) & 3], bbase + (unsigned)buf * CHAIN_B_BYTES, smem_u32(&bfull_bar[buf]));
      }
      if (gmark && ctl) gmark[7] = gtime_ns();
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (warp < CH_WW) {
      const ChainPend& in = op.in;
      const float* res = dynp(in.res, a.dyn);
      if (op.kind == CH_NORM || op.kind == CH_BSQ) {
        // ============================================================================== row per CTA: norm (+ BSQ)
        // thread = (float4 column cg, K-slice group kg): the K-slice partials of the row are summed two ways in parallel
        // and folded in fixed order (deterministic); warps 0-3 then hold the row and do the norm
        const int C = op.N, n4 = C >> 2;                   // C <= 512
        float4* red = reinterpret_cast<float4*>(smem);     // [2][128]
        float* scr = reinterpret_cast<float*>(smem + 4 * 128 * 16);   // [8 + 13 * 4]
        const int cg = tid & 127, kg = tid >> 7, col = cg * 4;
        const bool active = cg < n4;
        float* xout = dynp(op.xout, a.dyn);
        float* xout2 = dynp(op.xout2, a.dyn);
        const float* prev = dynp(op.prev, a.dyn);
        for (int r = blockIdx.x; r < op.M; r += (int)nblocks) {
          long long ro;
          if (!row_map(op, r, ro)) continue;               // (CTA-uniform)
          long long srow = r;
          bool from_prev = false;
          if (op.asm_S > 0) {
            if (r < op.asm_rf) srow = r;
            else if (r < op.asm_S - op.asm_c) { from_prev = true; srow = r + op.asm_c; }
            else srow = 2 * op.asm_Ls - (op.asm_S - r);
          }
          float4 part = make_float4(0.f, 0.f, 0.f, 0.f);
          if (active) {
            if (from_prev) {
              if (kg == 0) part = ldcg4(prev + srow * C + col);
            } else if (in.P) {
              const float* p = in.P + srow * in.ldp + col;
#pragma unroll 8
              for (int k = kg; k < in.ks; k += 2) {
                const float4 v = ldcg4(p + (long long)k * in.ks_stride);
                part.x += v.x; part.y += v.y; part.z += v.z; part.w += v.w;
              }
            }
          }
          red[kg * 128 + cg] = part;
          asm volatile("bar.sync 1, %0;" ::"n"(CH_WORKERS) : "memory");
          if (tid < 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            float s = 0.f;
            if (active) {
              const float4 r0 = red[cg], r1 = red[128 + cg];
              v.x = r0.x + r1.x; v.y = r0.y + r1.y; v.z = r0.z + r1.z; v.w = r0.w + r1.w;
              if (!from_prev) {
                if (in.bias) { const float4 bb = ldg4(in.bias + col); v.x += bb.x; v.y += bb.y; v.z += bb.z; v.w += bb.w; }
                if (in.gamma) { const float4 g = ldg4(in.gamma + col); v.x *= g.x; v.y *= g.y; v.z *= g.z; v.w *= g.w; }
                if (res) { const float4 rr = ldcg4(res + srow * in.ldr + col); v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w; }
              }
              if (xout) *reinterpret_cast<float4*>(xout + (long long)r * op.ldx + col) = v;
              if (xout2) *reinterpret_cast<float4*>(xout2 + (long long)r * op.ldx2 + col) = v;
              s = (op.norm == CHN_RMS) ? v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w : v.x + v.y + v.z + v.w;
            }
            s = warp_sum(s);
            if (lane == 0) scr[warp] = s;
            asm volatile("bar.sync 2, 128;" ::: "memory");
            const float tot = ((scr[0] + scr[1]) + scr[2]) + scr[3];
            float mean = 0.f, inv;
            if (op.norm == CHN_RMS) {
              inv = rsqrtf(tot / C + op.eps);
            } else {
              mean = tot / C;
              float q = 0.f;
              if (active) {
                const float dx = v.x - mean, dy = v.y - mean, dz = v.z - mean, dw = v.w - mean;
                q = dx * dx + dy * dy + dz * dz + dw * dw;
              }
              q = warp_sum(q);
              if (lane == 0) scr[4 + warp] = q;
              asm volatile("bar.sync 2, 128;" ::: "memory");
              inv = 1.f / sqrtf((((scr[4] + scr[5]) + scr[6]) + scr[7]) / C + op.eps);
            }
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) {
              const float4 w = ldg4(op.w + col);
              o.x = (v.x - mean) * inv * w.x; o.y = (v.y - mean) * inv * w.y;
              o.z = (v.z - mean) * inv * w.z; o.w = (v.w - mean) * inv * w.w;
              if (op.b) { const float4 bb = ldg4(op.b + col); o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w; }
              if (op.y) *reinterpret_cast<float4*>(op.y + ro * op.ldy + col) = o;
            }
            if (op.kind == CH_BSQ) {
              // LFQ ids (bsq.py:330-369): bit_i = (proj_i > 0), id = sum bit_i << (12 - i)
              for (int bit = 0; bit < BSQ_BITS; ++bit) {
                float d = 0.f;
                if (active) {
                  const float4 w = ldg4(op.table + bit * C + col);
                  d = fmaf(o.x, w.x, d); d = fmaf(o.y, w.y, d); d = fmaf(o.z, w.z, d); d = fmaf(o.w, w.w, d);
                }
                d = warp_sum(d);
                if (lane == 0) scr[8 + bit * 4 + warp] = d;
              }
              asm volatile("bar.sync 2, 128;" ::: "memory");
              if (tid == 0) {
                long long id = 0;
                for (int bit = 0; bit < BSQ_BITS; ++bit) {
                  const float d = (((scr[8 + bit * 4] + scr[9 + bit * 4]) + scr[10 + bit * 4]) + scr[11 + bit * 4]) + __ldg(op.table_b + bit);
                  if (d > 0.f) id |= 1LL << (BSQ_BITS - 1 - bit);
                }
                long long* ids = dynp(op.ids, a.dyn);
                ids[op.q_first + r] = id;
              }
            }
          }
          asm volatile("bar.sync 1, %0;" ::"n"(CH_WORKERS) : "memory");
        }
      } else if (op.kind == CH_DWLN) {
        // ============================================================================== depthwise causal conv k = 7 + LayerNorm
        const int C = op.N, nv = C >> 7;
        for (int r = blockIdx.x + (int)nblocks * warp; r < op.M; r += (int)nblocks * CH_WW) {
          long long ro;
          if (!row_map(op, r, ro)) continue;
          const int seg0 = op.seg_rows > 0 ? (r / op.seg_rows) * op.seg_rows : 0;
          float4 acc[4];
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) acc[i] = ldg4(op.dw_b + (lane + 32 * i) * 4);
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            const int rr = r - 6 + j;
            if (rr >= seg0) {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (i < nv) {
                  const int col = (lane + 32 * i) * 4;
                  const float4 x = ldcg4(res + (long long)rr * in.ldr + col);
                  const float4 w = ldg4(op.dw_w + j * C + col);
                  acc[i].x = fmaf(w.x, x.x, acc[i].x); acc[i].y = fmaf(w.y, x.y, acc[i].y);
                  acc[i].z = fmaf(w.z, x.z, acc[i].z); acc[i].w = fmaf(w.w, x.w, acc[i].w);
                }
            }
          }
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) s += acc[i].x + acc[i].y + acc[i].z + acc[i].w;
          const float mean = warp_sum(s) / C;
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) {
              const float dx = acc[i].x - mean, dy = acc[i].y - mean, dz = acc[i].z - mean, dw = acc[i].w - mean;
              q += dx * dx + dy * dy + dz * dz + dw * dw;
            }
          const float inv = 1.f / sqrtf(warp_sum(q) / C + op.eps);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i < nv) {
              const int col = (lane + 32 * i) * 4;
              const float4 w = ldg4(op.w + col), bb = ldg4(op.b + col);
              float4 o;
              o.x = (acc[i].x - mean) * inv * w.x + bb.x; o.y = (acc[i].y - mean) * inv * w.y + bb.y;
              o.z = (acc[i].z - mean) * inv * w.z + bb.z; o.w = (acc[i].w - mean) * inv * w.w + bb.w;
              *reinterpret_cast<float4*>(op.y + ro * op.ldy + col) = o;
            }
        }
      } else if (op.kind == CH_ACT) {
        // ============================================================================== element-wise: y = act(value)
        const int n4 = op.N >> 2;
        const int total = op.M * n4;                       // < 2^31: M <= 384 rows
        float* yb = dynp(op.y, a.dyn);
        for (int idx = blockIdx.x * CH_WORKERS + tid; idx < total; idx += (int)nblocks * CH_WORKERS) {
          const int r = idx / n4;
          const int col = (idx - r * n4) * 4;
          long long ro;
          if (!row_map(op, r, ro)) continue;
          float4 o;
          if (op.act == CHA_SILU_MUL) {
            const float4 h1 = pend4(in, res, r, col), h3 = pend4(in, res, r, op.N + col);
            o.x = (h1.x / (1.f + expf(-h1.x))) * h3.x; o.y = (h1.y / (1.f + expf(-h1.y))) * h3.y;
            o.z = (h1.z / (1.f + expf(-h1.z))) * h3.z; o.w = (h1.w / (1.f + expf(-h1.w))) * h3.w;
          } else {
            o = pend4(in, res, r, col);
            if (op.act == CHA_GELU) { o.x = gelu_erf(o.x); o.y = gelu_erf(o.y); o.z = gelu_erf(o.z); o.w = gelu_erf(o.w); }
          }
          *reinterpret_cast<float4*>(yb + ro * op.ldy + col) = o;
        }
      } else if (op.kind == CH_QKV_ROPE) {
        // ============================================================================== qkv = value; RoPE on q and k
        const int D = op.heads * HEAD_DIM;                 // q | k | v blocks of D columns
        const int n4 = (3 * D) >> 2;
        const long long total = (long long)op.M * n4;
        for (long long idx = (long long)blockIdx.x * CH_WORKERS + tid; idx < total; idx += (long long)nblocks * CH_WORKERS) {
          const long long r = idx / n4;
          const int col = (int)(idx - r * n4) * 4;
          float4 x = pend4(in, res, r, col);
          if (col < 2 * D) {
            const int i0 = (col & (HEAD_DIM - 1)) >> 1;    // pair index of (x.x, x.y); (x.z, x.w) is pair i0 + 1
            const float4 cs = ldg4(op.table + ((long long)(op.q_first + r) * (HEAD_DIM / 2) + i0) * 2);
            float4 o;
            o.x = x.x * cs.x - x.y * cs.y; o.y = x.y * cs.x + x.x * cs.y;
            o.z = x.z * cs.z - x.w * cs.w; o.w = x.w * cs.z + x.z * cs.w;
            x = o;
          }
          *reinterpret_cast<float4*>(op.y + r * op.ldy + col) = x;
        }
      } else if (op.kind == CH_ATTN) {
        // ============================================================================== causal window attention, <= 128 keys
        // job = (head, block of 8 queries): K and V rows of the head staged once in shared memory (every thread issues its
        // loads before it stores: two L2 round trips), one warp per query
        float (*Ks)[HEAD_DIM + 1] = reinterpret_cast<float (*)[HEAD_DIM + 1]>(smem);
        float (*Vs)[HEAD_DIM] = reinterpret_cast<float (*)[HEAD_DIM]>(smem + 128 * (HEAD_DIM + 1) * 4);
        float4* Qs = reinterpret_cast<float4*>(smem + 128 * (HEAD_DIM + 1) * 4 + 128 * HEAD_DIM * 4);
        const int qblocks = (op.nq + CH_WW - 1) / CH_WW;
        const int D = op.heads * HEAD_DIM;
        const long long ld = 3 * D;
        for (int job = blockIdx.x; job < op.heads * qblocks; job += (int)nblocks) {
          const int h = job % op.heads, qb = job / op.heads;
          const int q_lo = op.q_first + qb * CH_WW, q_hi = min(op.q_first + op.nq, q_lo + CH_WW);   // [q_lo, q_hi)
          const int k_hi = q_hi - 1;                        // newest key needed
          const int k_lo = max(0, q_lo - op.window + 1);
          const float* kb = op.A + D + h * HEAD_DIM;
          const float* vb = op.A + 2 * D + h * HEAD_DIM;
          const int items = (k_hi - k_lo + 1) * (HEAD_DIM / 4);          // <= 2048 = 8 per thread
          const int qi = q_lo + warp;
          const bool has_q = qi < q_hi;
          // the 8 query rows travel through shared memory as well (one float4 per thread of warps 0-3, same round trip)
          float4 qv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (tid < CH_WW * (HEAD_DIM / 4) && q_lo + tid / (HEAD_DIM / 4) < q_hi)
            qv = ldcg4(op.A + (long long)(q_lo + tid / (HEAD_DIM / 4)) * ld + h * HEAD_DIM + (tid % (HEAD_DIM / 4)) * 4);
#pragma unroll
          for (int j0 = 0; j0 < 8; j0 += 4) {
            float4 kr[4], vr[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = min(tid + (j0 + j) * CH_WORKERS, items - 1);   // (clamped: the store below is predicated)
              const int key = k_lo + i / (HEAD_DIM / 4), c4 = (i % (HEAD_DIM / 4)) * 4;
              kr[j] = ldcg4(kb + (long long)key * ld + c4);
              vr[j] = ldcg4(vb + (long long)key * ld + c4);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int i = tid + (j0 + j) * CH_WORKERS;
              if (i < items) {
                const int key = k_lo + i / (HEAD_DIM / 4), c4 = (i % (HEAD_DIM / 4)) * 4;
                Ks[key][c4] = kr[j].x; Ks[key][c4 + 1] = kr[j].y; Ks[key][c4 + 2] = kr[j].z; Ks[key][c4 + 3] = kr[j].w;
                *reinterpret_cast<float4*>(&Vs[key][c4]) = vr[j];
              }
            }
          }
          if (tid < CH_WW * (HEAD_DIM / 4)) Qs[tid] = qv;
          asm volatile("bar.sync 1, %0;" ::"n"(CH_WORKERS) : "memory");
          if (has_q) {
            float qr[HEAD_DIM];
#pragma unroll
            for (int d4 = 0; d4 < HEAD_DIM / 4; ++d4) {
              const float4 t = Qs[warp * (HEAD_DIM / 4) + d4];
              qr[4 * d4] = t.x * 0.125f; qr[4 * d4 + 1] = t.y * 0.125f; qr[4 * d4 + 2] = t.z * 0.125f; qr[4 * d4 + 3] = t.w * 0.125f;
            }
            const int pos = qi;
            const int lo = max(0, pos - op.window + 1);
            float m = -INFINITY, l = 0.f, acc0 = 0.f, acc1 = 0.f;
            for (int kt = (lo / 32) * 32; kt <= pos; kt += 32) {
              const int kp = kt + lane;
              const bool valid = (kp >= lo) && (kp <= pos);
              float s = 0.f;
              if (valid) {
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
              const int jn = min(32, pos - kt + 1);
              for (int j = 0; j < jn; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p, j);
                acc0 = fmaf(pj, Vs[kt + j][lane], acc0);
                acc1 = fmaf(pj, Vs[kt + j][lane + 32], acc1);
              }
              m = m_new;
            }
            float* out = op.y + (long long)qi * op.ldy + h * HEAD_DIM;
            const float inv = 1.f / l;
            out[lane] = acc0 * inv;
            out[lane + 32] = acc1 * inv;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(CH_WORKERS) : "memory");
        }
      }
    }
    if (fetch_next) asm volatile("cp.async.wait_all;" ::: "memory");
    if (a.prof) __syncthreads();
    if (prof) pslot[1] = gtime_ns();
    if (oi + 1 < a.n_ops) {
      if (op.no_grid_sync) {               // (read before anyone can overwrite the op: the barrier below comes first)
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      } else {
        grid_sync(gb, nblocks);
      }
    }
    if (prof) pslot[2] = gtime_ns();
  }
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(CH_TMEM_COLS) : "memory");
  // every CTA has passed the last barrier's arrive before any CTA gets here: the counter is final
  if (blockIdx.x == 0 && tid == 0) a.barrier[1] = gb.target;
}

}  // namespace

Chain::~Chain() {
  if (ops_dev) cudaFree(ops_dev);
  if (gemm_ops_dev) cudaFree(gemm_ops_dev);
  if (wjobs_dev) cudaFree(wjobs_dev);
}

void Chain::upload(int grid_) {
  grid = grid_;
  std::vector<int> gi;
  gemm_flop = 0;
  for (size_t i = 0; i < ops.size(); ++i)
    if (ops[i].kind == CH_GEMM) {
      ops[i].gemm_seq = (int)gi.size();
      gi.push_back((int)i);
      gemm_flop += 2.0 * ops[i].M * ops[i].N * ops[i].K;
    }
  n_gemm = (int)gi.size();
  if (ops_dev) cudaFree(ops_dev);
  if (gemm_ops_dev) cudaFree(gemm_ops_dev);
  ops_dev = nullptr;
  gemm_ops_dev = nullptr;
  SV_CUDA(cudaMalloc(&ops_dev, ops.size() * sizeof(ChainOp)));
  SV_CUDA(cudaMemcpy(ops_dev, ops.data(), ops.size() * sizeof(ChainOp), cudaMemcpyHostToDevice));
  SV_CUDA(cudaMalloc(&gemm_ops_dev, (gi.size() + 1) * sizeof(int)));
  if (!gi.empty()) SV_CUDA(cudaMemcpy(gemm_ops_dev, gi.data(), gi.size() * sizeof(int), cudaMemcpyHostToDevice));
  // per GEMM and CTA: where the CTA's weight block lives (the kernel's control thread reads one descriptor per GEMM)
  std::vector<ChainWJob> wj((size_t)(n_gemm + 1) * grid);      // [cta][q]
  for (int q = 0; q < n_gemm; ++q) {
    const ChainOp& g = ops[gi[q]];
    SV_CHECK(g.n_tiles * g.ksplit <= grid && ((g.M + 127) / 128) * g.BN <= CH_TMEM_COLS, "chain: GEMM phase does not fit");
    SV_CHECK(g.epi == EPI_PARTIAL || g.ksplit == 1, "chain: a direct epilogue needs the whole K range in one job");
    SV_CHECK(g.epi != EPI_SILU_MUL || g.BN % 32 == 0, "chain: silu-mul tiles hold h1 | h3 column pairs");
    SV_CHECK(!g.acc_cont || (q > 0 && ops[gi[q - 1]].acc_keep && ops[gi[q - 1]].BN == g.BN && ops[gi[q - 1]].n_tiles == g.n_tiles &&
                             ops[gi[q - 1]].ksplit == g.ksplit && ops[gi[q - 1]].M == g.M), "chain: accumulator continuation");
    for (int cta = 0; cta < grid; ++cta) {
      ChainWJob& w = wj[(size_t)cta * n_gemm + q];
      w = ChainWJob{nullptr, nullptr, 0u, 0, 0, 0, 0, 0, {0, 0, 0, 0}};
      if (cta >= g.n_tiles * g.ksplit) continue;
      const int nt = cta % g.n_tiles, ks = cta / g.n_tiles;
      const int s0 = g.slabs * ks / g.ksplit, s1 = g.slabs * (ks + 1) / g.ksplit;
      SV_CHECK((s1 - s0) * 2 * g.BN * 128 <= CHAIN_B_BYTES, "chain: weight block exceeds the buffer");
      const long long off = ((long long)(g.slab_lo + s0) * g.wt_npad + (long long)nt * g.BN) * 128;
      w.w0 = g.Wt0 + off;
      w.w1 = g.Wt1 + off;
      w.bytes = (unsigned)g.BN * 128u;
      w.n_sl = s1 - s0;
      w.slab_stride = (long long)g.wt_npad * 128;
      w.a_off = (long long)(g.slab_lo + s0) * 32;
      w.n0 = nt * g.BN;
      w.ks = ks;
    }
  }
  if (wjobs_dev) cudaFree(wjobs_dev);
  wjobs_dev = nullptr;
  SV_CUDA(cudaMalloc(&wjobs_dev, wj.size() * sizeof(ChainWJob)));
  SV_CUDA(cudaMemcpy(wjobs_dev, wj.data(), wj.size() * sizeof(ChainWJob), cudaMemcpyHostToDevice));
  uploaded = true;
}

// (BN, ksplit) that minimises the bytes the busiest CTA moves on the critical path: its A rows (every M tile of the K slice),
// its partial tile, and its share of the consumer's partial reads; the weight block is prefetched a GEMM ahead and counts a
// quarter.  Limits: jobs <= grid, weight block <= CHAIN_B_BYTES, M tiles x BN <= the TMEM allocation.
bool chain_gemm_config(int M, int N, int K, int grid, int* BN_out, int* ks_out) {
  if (K <= 0 || K % 32 != 0 || N % 16 != 0 || M <= 0) return false;
  const int slabs = K / 32, mt = (M + 127) / 128;
  if (mt > CHAIN_MAX_MTILES) return false;
  double best = 1e30;
  bool found = false;
  for (int bn : {64, 32, 16}) {
    if (mt * bn > CH_TMEM_COLS) continue;
    const int nt = (N + bn - 1) / bn;
    for (int s = 1; s <= slabs && s <= 32; ++s) {
      if ((long long)nt * s > grid) break;
      const int sl = (slabs + s - 1) / s;
      if (sl * 2 * bn * 128 > CHAIN_B_BYTES) continue;
      const double a_bytes = (double)mt * 128 * sl * 128, b_bytes = (double)sl * bn * 256, out = (double)M * bn * 4;
      const double cons = (double)s * M * N * 4 / grid;
      const double cost = a_bytes + 0.25 * b_bytes + out + cons;
      if (cost < best) { best = cost; *BN_out = bn; *ks_out = s; found = true; }
    }
  }
  return found;
}

size_t chain_partial_floats(int M, int N, int K, int grid) {
  int bn = 0, ks = 0;
  SV_CHECK(chain_gemm_config(M, N, K, grid, &bn, &ks), "chain: GEMM shape does not fit a phase");
  return (size_t)ks * M * N;
}

void chain_set_gemm_tiled(ChainOp& op, const float* A, long long a_row_stride, const float* W, int M, int N, int K, int bn, int ks,
                          int slab_lo, int slabs, int grid, cudaStream_t st) {
  SV_CHECK(K % 32 == 0 && N % 16 == 0 && slab_lo >= 0 && slabs >= 1 && slab_lo + slabs <= K / 32, "chain: GEMM shape");
  SV_CHECK(a_row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "chain: A rows must be 16-byte aligned");
  const void *t0 = nullptr, *t1 = nullptr;
  int n_pad = 0;
  SV_CHECK(gemm_tiled_weights(W, 1, N, K, false, st, &t0, &t1, &n_pad), "chain: pre-tiled weights unavailable (stream capture?)");
  op.kind = CH_GEMM;
  op.M = M; op.N = N; op.K = slabs * 32;
  op.A = A; op.a_row_stride = a_row_stride;
  op.Wt0 = static_cast<const unsigned char*>(t0);
  op.Wt1 = static_cast<const unsigned char*>(t1);
  op.wt_npad = n_pad;
  op.BN = bn; op.n_tiles = (N + bn - 1) / bn; op.ksplit = ks; op.slabs = slabs; op.slab_lo = slab_lo;
  SV_CHECK(op.n_tiles * bn <= n_pad && op.n_tiles * ks <= grid, "chain: tiling does not fit");
  op.Pout = nullptr; op.ldp_out = N; op.pout_ks_stride = (long long)M * N;
}

void chain_set_gemm(ChainOp& op, const float* A, long long a_row_stride, const float* W, int M, int N, int K, float* P, int grid,
                    cudaStream_t st) {
  int bn = 0, ks = 0;
  SV_CHECK(chain_gemm_config(M, N, K, grid, &bn, &ks), "chain: GEMM shape does not fit a phase");
  SV_CHECK(a_row_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "chain: A rows must be 16-byte aligned");
  const void *t0 = nullptr, *t1 = nullptr;
  int n_pad = 0;
  SV_CHECK(gemm_tiled_weights(W, 1, N, K, false, st, &t0, &t1, &n_pad), "chain: pre-tiled weights unavailable (stream capture?)");
  op.kind = CH_GEMM;
  op.M = M; op.N = N; op.K = K;
  op.A = A; op.a_row_stride = a_row_stride;
  op.Wt0 = static_cast<const unsigned char*>(t0);
  op.Wt1 = static_cast<const unsigned char*>(t1);
  op.wt_npad = n_pad;
  op.BN = bn; op.n_tiles = (N + bn - 1) / bn; op.ksplit = ks; op.slabs = K / 32;
  SV_CHECK(op.n_tiles * bn <= n_pad, "chain: weight tile rows past the padded copy");
  op.Pout = P; op.ldp_out = N; op.pout_ks_stride = (long long)M * N;
}

static int chain_env_mode() {
  const char* e = getenv("SVANON_CHAIN");                // svanon_set_chain_mode's argument; default 1
  return e ? atoi(e) : 1;
}
bool g_use_chain = (chain_env_mode() & 1) != 0;          // bit 0: the encoder's transformer half as a chain launch
bool g_chain_conv = (chain_env_mode() & 2) != 0;         // bit 1: the conv stack inside the chain as well (measured slower: off)

bool chain_supported(int grid) { return g_use_chain && grid >= 100 && !g_gemm_half; }

void launch_chain(Chain& c, const ChainDyn& dyn, unsigned* barrier, int grid, cudaStream_t st) {
  SV_CHECK(c.uploaded && !c.ops.empty() && c.grid == grid, "chain not built for this grid");
  static bool configured = false;
  if (!configured) {
    SV_CUDA(cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES));
    configured = true;
  }
  ChainArgs args;
  args.ops = c.ops_dev;
  args.gemm_ops = c.gemm_ops_dev;
  args.wjobs = c.wjobs_dev;
  args.n_ops = (int)c.ops.size();
  args.n_gemm = c.n_gemm;
  args.barrier = barrier;
  args.dyn = dyn;
  static const int prof_at = [] { const char* e = getenv("SVANON_CHAIN_PROF"); return e ? atoi(e) : 0; }();   // print the n-th launch
  static int n_launch = 0;
  args.prof = nullptr;
  const bool do_prof = prof_at > 0 && ++n_launch == prof_at;
  static const int trace_op = [] { const char* e = getenv("SVANON_CHAIN_TRACE"); return e ? atoi(e) : -1; }();
  args.trace_op = trace_op;
  if (do_prof) {
    SV_CUDA(cudaMalloc(&args.prof, (c.ops.size() * 16 + 64) * sizeof(unsigned long long)));
    SV_CUDA(cudaMemset(args.prof, 0, (c.ops.size() * 16 + 64) * sizeof(unsigned long long)));
  }
  void* kargs[] = {(void*)&args};
  const bool timing = gemm_timing_on();
  if (timing) gemm_timing_external(st, true, 0);
  SV_CUDA(cudaLaunchCooperativeKernel((void*)chain_kernel, dim3(grid), dim3(CH_THREADS), kargs, CH_SMEM_BYTES, st));
  ++g_kernel_launches;
  if (timing) gemm_timing_external(st, false, c.gemm_flop);
  if (do_prof) {
    SV_CUDA(cudaStreamSynchronize(st));
    std::vector<unsigned long long> h(c.ops.size() * 16 + 64);
    SV_CUDA(cudaMemcpy(h.data(), args.prof, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    cudaFree(args.prof);
    const char* names[8] = {"?", "GEMM", "NORM", "DWLN", "ACT", "QKV_ROPE", "ATTN", "BSQ"};
    const unsigned long long t0 = h[0];
    if (trace_op >= 0 && trace_op < (int)c.ops.size()) {
      const unsigned long long* tr = h.data() + c.ops.size() * 16;
      const unsigned long long b = h[(size_t)trace_op * 16];
      fprintf(stderr, "chain trace of op %d (us after the phase start): slab | ring slot free, stored + arrived | operands seen by the MMA thread, MMAs issued + committed\n", trace_op);
      for (int it = 0; it < 16; ++it)
        if (tr[it * 4 + 1])
          fprintf(stderr, "  %2d | %6.2f %6.2f | %6.2f %6.2f\n", it, (double)(long long)(tr[it * 4] - b) * 1e-3, (double)(long long)(tr[it * 4 + 1] - b) * 1e-3,
                  (double)(long long)(tr[it * 4 + 2] - b) * 1e-3, (double)(long long)(tr[it * 4 + 3] - b) * 1e-3);
    }
    for (size_t i = 0; i < c.ops.size(); ++i) {
      const ChainOp& o = c.ops[i];
      fprintf(stderr, "chain op %3zu %-8s M=%4d N=%4d K=%4d bn=%2d nt=%3d ks=%2d | cta0: start %8.2f us  work %6.2f  barrier %6.2f | last cta: work %6.2f  barrier %6.2f\n",
              i, names[o.kind & 7], o.M, o.N, o.K, o.BN, o.n_tiles, o.ksplit, (h[i * 16] - t0) * 1e-3, (h[i * 16 + 1] - h[i * 16]) * 1e-3,
              (h[i * 16 + 2] - h[i * 16 + 1]) * 1e-3, (h[i * 16 + 4] - h[i * 16 + 3]) * 1e-3, (h[i * 16 + 5] - h[i * 16 + 4]) * 1e-3);
      if (o.kind == CH_GEMM) {
        const char* mn[8] = {"loads issued", "first load back", "mma: start", "mma: weights in", "mma: issued", "acc ready", "epilogue done", "prefetch issued"};
        fprintf(stderr, "      ");
        for (int k = 0; k < 8; ++k)
          if (h[i * 16 + 6 + k]) fprintf(stderr, " [%s +%.2f]", mn[k], (double)(long long)(h[i * 16 + 6 + k] - h[i * 16]) * 1e-3);
        fprintf(stderr, "\n");
      }
    }
  }
}

}  // namespace svanon
