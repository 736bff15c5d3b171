// Shared declarations for the svanon_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

namespace svanon {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

#define SV_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess)                                                                 \
      throw ::svanon::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " + \
                            __FILE__ + ":" + std::to_string(__LINE__));                    \
  } while (0)

#define SV_CHECK(cond, msg)                                                            \
  do {                                                                                 \
    if (!(cond))                                                                       \
      throw ::svanon::Error(std::string(msg) + " (" #cond ") at " + __FILE__ + ":" + \
                            std::to_string(__LINE__));                                 \
  } while (0)

// NVTX range over a host-side scope: the stages of the per-chunk loop show up by name ("svanon:E window", "svanon:A
// decode", "svanon:V step", ...) on an Nsight Systems / ncu --nvtx timeline.  Costs a few ns without a profiler attached.
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
extern long long g_kernel_launches;
#define SV_LAUNCHED()                 \
  do {                                \
    ++::svanon::g_kernel_launches;    \
    SV_CUDA(cudaGetLastError());      \
  } while (0)

// Programmatic dependent launch: every kernel lets its successor start launching right away; kernels launched with
// the programmatic-serialization attribute (the GEMMs) block in pdl_wait() until their predecessors have completed
// and flushed, after doing the part of their prologue that only touches weights.
extern bool g_use_pdl;
#ifdef __CUDACC__
// launch with the programmatic-stream-serialization attribute (the kernel must call pdl_wait() before it touches
// anything a predecessor wrote)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_use_pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SV_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
}
// L2 prefetch of a contiguous global range (16-byte aligned, size a multiple of 16): one bulk instruction, no destination.
// Used for WEIGHTS before griddepcontrol.wait: in the per-chunk chain every GEMM's weights are cold (E + A + V stream
// ~0.8 GB of fp32 weights through the 126 MB L2 per chunk), and with programmatic dependent launch the kernel starts
// while its predecessor still runs -- so its weight slice can travel DRAM -> L2 during that time instead of inside the
// main loop, whose per-slab cost is otherwise one DRAM round trip per prefetch depth.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---------------------------------------------------------------- model constants
// configs/hydra_arcs/vc/firefly_arvc_bsq_8192_delay0_8.yaml
constexpr int AR_DIM = 768;
constexpr int AR_HEADS = 12;
constexpr int HEAD_DIM = 64;
constexpr int AR_INTER = 2304;
constexpr int AR_LAYERS = 12;
constexpr int AR_FAST_LAYERS = 4;
constexpr int AR_CODEBOOKS = 8;
constexpr int AR_CB_SIZE = 1000;
constexpr int AR_VOCAB = 8192;
constexpr int AR_MAX_SEQ = 2048;
constexpr int AR_SPK_TOKENS = 33;   // 32 timbre tokens + 1 style token (arvc_wrapper.py:108-109)
constexpr int AR_MAX_DELAY = 8;
constexpr float AR_NORM_EPS = 1e-5f;
// configs/hydra_arcs/speech_tokenizers/causal-encoder-lfq-8192.yaml
constexpr int N_FFT = 2048;
constexpr int HOP = 512;
constexpr int N_FREQ = 1025;
constexpr int N_FREQ_PAD = 1040;    // K of the mel GEMM, padded to a multiple of 16
constexpr int N_MELS = 160;
constexpr int ENC_DIM = 512;
constexpr int ENC_HEADS = 8;
constexpr int ENC_INTER = 1536;
constexpr int ENC_LAYERS = 8;
constexpr int ENC_WINDOW = 512;
constexpr int BSQ_BITS = 13;
constexpr int SAMPLES_PER_FRAME = 2048;

// ---------------------------------------------------------------- generic GEMM
// C[m, n] (op)= epilogue( sum_t sum_k pro(A[(m * a_row_step + tap_off[t]) * lda + k]) * W[t][n][k] )
// A rows may overlap (lda < K) and may start before the buffer's row 0 (history / zero margin rows).
enum Prologue : int { PRO_NONE = 0, PRO_SILU = 1 };
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_LOGCLAMP = 2 };

constexpr int MAX_TAPS = 11;

struct GemmParams {
  const float* A = nullptr;
  const float* W = nullptr;        // [taps][N][K]
  float* C = nullptr;
  const float* bias = nullptr;     // [N] or null
  const float* gamma = nullptr;    // [N] or null: y *= gamma[n] (after activation)
  const float* residual = nullptr; // [M][ldr] or null: y += residual
  int M = 0, N = 0, K = 0;
  long long lda = 0, ldc = 0, ldr = 0;
  int a_row_step = 1;
  int taps = 1;
  int tap_off[MAX_TAPS] = {0};     // row offset (in A rows) of each tap
  int prologue = PRO_NONE;
  int act = ACT_NONE;
  float out_scale = 1.f;           // y *= out_scale (applied last, before accumulate)
  int accumulate = 0;              // C += y instead of C = y
  // Independent streams side by side: with seg_rows > 0 output row m belongs to segment b = m / seg_rows (row
  // t = m % seg_rows of that stream) and the A / C / residual rows of segment b start b * {a,c,r}_seg floats after
  // the base pointer (each stream's buffer keeps its own causal margin rows in front).  seg_rows == 0: one stream.
  int seg_rows = 0;
  long long a_seg = 0, c_seg = 0, r_seg = 0;
  // perf mode (svanon_set_precision): fp16 copy of W, filled in by the tensor-core launcher from its registry.  w_static =
  // false keeps a launch out of that registry (W is caller memory that may change: svanon_debug_gemm*).
  const void* Wh = nullptr;
  bool w_static = true;
  // pre-tiled weight copies for the TMA B path of the tensor-core kernel (gemm_tc.cu): Wt[0] = hi term (or the fp16 copy),
  // Wt[1] = lo term (null in perf mode), rows padded to wt_npad per (tap, K-slab); filled in by the launcher
  const void* Wt[2] = {nullptr, nullptr};
  int wt_npad = 0;
  // pair kernel (gemm_pair.cu): lo terms x - trunc_tf32(x) as arrays of their own.  Alo: the lo term of A, compact [M][K], written
  // by whoever produced A (null: one split pass in front of the GEMM); Clo: if set, the kernel also writes the lo term of its
  // result, compact [M][N], for the GEMM that consumes C next.  Kernels other than the pair kernel ignore both.
  const float* Alo = nullptr;
  float* Clo = nullptr;
  // Fused forms (launch_gemm computes them with any back end; the pair kernel does them in its epilogue, the others through the
  // row-wise kernels afterwards):
  //   W2 != null: SwiGLU gate, C[m, n] = silu(sum_k A W[n]) * (sum_k A W2[n]) -- W2 has W's shape, no bias / gamma / residual;
  //     dual_tmp [M][2N] is scratch for back ends that need the two products in memory.
  //   rope_table != null: interleaved-pair RoPE on the first rope_cols columns of C (q | k of a fused qkv row), position of row
  //     m = rope_pos0 + (rope_seg_rows > 0 ? m % rope_seg_rows : m); table as launch_rope_qk takes it.
  const float* W2 = nullptr;
  float* dual_tmp = nullptr;
  const float* rope_table = nullptr;
  int rope_cols = 0, rope_seg_rows = 0, rope_pos0 = 0;
};

#ifdef __CUDACC__
__device__ __forceinline__ long long gemm_a_row(const GemmParams& p, int m) {
  if (p.seg_rows > 0) {
    const int b = m / p.seg_rows;
    return (long long)b * p.a_seg + (long long)(m - b * p.seg_rows) * p.a_row_step * p.lda;
  }
  return (long long)m * p.a_row_step * p.lda;
}
__device__ __forceinline__ long long gemm_c_row(const GemmParams& p, int m) {
  if (p.seg_rows > 0) {
    const int b = m / p.seg_rows;
    return (long long)b * p.c_seg + (long long)(m - b * p.seg_rows) * p.ldc;
  }
  return (long long)m * p.ldc;
}
__device__ __forceinline__ long long gemm_r_row(const GemmParams& p, int m) {
  if (p.seg_rows > 0) {
    const int b = m / p.seg_rows;
    return (long long)b * p.r_seg + (long long)(m - b * p.seg_rows) * p.ldr;
  }
  return (long long)m * p.ldr;
}
// lo term of the 3xTF32 split: x - trunc_tf32(x) (exact in fp32).  Row-wise kernels that feed a wide GEMM write it beside their
// result (`*_lo` arguments, same compact layout) so that the pair kernel (gemm_pair.cu) takes both terms by TMA.
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
// interleaved-pair rotation of apply_rotary_emb with the rounding points spelled out (one product rounded, the other fused),
// so that every kernel that applies it -- rope_qk_kernel, the pair GEMM's epilogue -- produces the same bits
__device__ __forceinline__ void rope_pair(float x0, float x1, float c, float s, float& y0, float& y1) {
  y0 = __fmaf_rn(x0, c, -__fmul_rn(x1, s));
  y1 = __fmaf_rn(x0, s, __fmul_rn(x1, c));
}
// row `row` of a buffer made of segments of seg_rows rows that start seg_stride floats apart (0 rows: plain)
__device__ __forceinline__ long long seg_row_off(long long row, int seg_rows, long long seg_stride, long long ld) {
  if (seg_rows > 0) {
    const long long b = row / seg_rows;
    return b * seg_stride + (row - b * seg_rows) * ld;
  }
  return row * ld;
}
#endif

// perf mode switch of the tensor-core GEMM (gemm_tc.cu): single-pass fp16 MMAs instead of the 3xTF32 split
extern bool g_gemm_half;
void gemm_half_release();
void gemm_forget_weights(const float* W);
// pair kernel switch (gemm_pair.cu): -1 environment (SVANON_GEMM_PAIR, default on), 0 off, 1 on, 2 on with masked hi copies
extern int g_gemm_pair_mode;
extern long long g_gemm_pair_launches;
extern bool g_gemm_pair_allowed;
bool gemm_pair_eligible(const GemmParams* ps, int count);
void gemm_pair_release();
void gemm_pair_forget_weights(const float* W);

// measurement aid behind svanon_gemm_timing: per back end, summed event-timed launch durations and executed flops
enum GemmBackend : int { GEMM_BACKEND_TC = 0, GEMM_BACKEND_PIPE = 1, GEMM_BACKEND_FP32 = 2, GEMM_BACKEND_CONV_SMALL = 3, GEMM_BACKENDS = 4 };
void gemm_timing_enable(bool on);
void gemm_timing_read(double* ms /*[4]*/, double* gflop /*[4]*/, long long* launches /*[4]*/);

// up to 3 independent problems of identical shape run as blockIdx.z
void launch_gemm(const GemmParams* p, int count, cudaStream_t st);
inline void launch_gemm(const GemmParams& p, cudaStream_t st) { launch_gemm(&p, 1, st); }

// ---------------------------------------------------------------- misc kernels (kernels_misc.cu)
// Row-wise kernels take an optional segment description for side-by-side streams: rows are numbered over all
// streams, stream b = row / seg_rows, and its rows start b * {x,y}_seg floats after the base (seg_rows 0: plain).
void launch_layernorm(const float* x, float* y, const float* w, const float* b, int rows, int C, float eps,
                      cudaStream_t st, int seg_rows = 0, long long x_seg = 0, long long y_seg = 0);
// depthwise causal conv k=7 (rows before 0 are read from the buffer margin) + LayerNorm over C, channels-last
void launch_dwconv7_ln(const float* x, float* y, const float* dw_w /*[7][C]*/, const float* dw_b, const float* ln_w,
                       const float* ln_b, int rows, int C, float eps, cudaStream_t st, int seg_rows = 0,
                       long long x_seg = 0, float* y_lo = nullptr);
void launch_rmsnorm(const float* x, float* y, const float* w, int rows, int C, float eps, cudaStream_t st,
                    long long x_ld = 0, float* y_lo = nullptr);
// interleaved-pair RoPE on q and k inside a fused [rows, 3*H*64] qkv buffer; table [pos][32][2] (bf16-rounded fp32)
void launch_rope_qk(float* qkv, const float* table, int rows, int heads, int pos0, cudaStream_t st, int seg_rows = 0);
void launch_silu_mul(const float* h13 /*[rows][2*I]*/, float* out /*[rows][I]*/, int rows, int I, cudaStream_t st,
                     float* out_lo = nullptr);
void launch_magnitude(const float* spec /*[T][ld_in] re|im*/, float* mag /*[T][N_FREQ_PAD]*/, int T, int ld_in,
                      cudaStream_t st);
void launch_bsq(const float* z /*[T][512]*/, const float* w /*[13][512]*/, const float* b, long long* ids, int T,
                cudaStream_t st);
void launch_fsq_lookup(const long long* codes /*[8][T] (stride ld)*/, long long ld, const float* w /*[8][64][4]*/,
                       const float* b /*[8][64]*/, float* z /*[T][512]*/, int T, cudaStream_t st, int seg_rows = 0,
                       long long codes_seg = 0);
void launch_fsq_encode(const float* z /*[B*T][512]*/, const float* w /*[8][4][64]*/, const float* b /*[8][4]*/,
                       int* codes /*[B][8][T]*/, int B, int T, cudaStream_t st);
void launch_conv_post(const float* x /*[L][16] with 12 margin rows*/, const float* w /*[13][16]*/, const float* b,
                      float* out, int L, cudaStream_t st, int seg_rows = 0, long long x_seg = 0);
void launch_resample(const float* x, long long n_in, const float* kern /*[new][taps]*/, int orig, int nw, int width, int taps,
                     float* out, long long n_out, cudaStream_t st);
// out = alpha * x + (1 - alpha) * (noise * std(x) + mean(x)), std unbiased over all n elements (apply_noise_mixing,
// evaluations/infer_arvc.py:228-232); statistics accumulated in fp64 by one CTA (n is 192 or 4096 in the reference)
void launch_noise_mix(const float* x, const float* noise, long long n, float alpha, float* out, cudaStream_t st);
void launch_gather_rows(const float* table, const long long* idx, float* out, int rows, int C, long long out_ld,
                        cudaStream_t st);
// out[t] = sum_i table[codes[i][t] + i*1000]  (BaseTransformer.embed)
void launch_embed_codes(const float* table, const int* codes /*[8][T] stride ld*/, long long ld, float* out, int T,
                        long long out_ld, cudaStream_t st);
void launch_copy_rows(const float* src, long long src_ld, float* dst, long long dst_ld, int rows, int C,
                      cudaStream_t st);
void launch_scale_add3(const float* a, const float* b, const float* c, float* out, long long n, float s, cudaStream_t st,
                       long long seg_n = 0, long long out_seg = 0);
void launch_fill(float* p, long long n, float v, cudaStream_t st, int nseg = 1, long long seg_stride = 0);
void launch_concat_cols(const int* a, long long a_ld, int a_n, const int* b, long long b_ld, int b_n, void* out,
                        long long out_ld, int rows, bool out_i64, cudaStream_t st);
void launch_append_codes(const int* codes8, int* hist, long long ld, int col, cudaStream_t st);
void launch_i64_to_i32(const long long* in, int* out, long long n, cudaStream_t st);

// ---------------------------------------------------------------- attention (attn.cu)
// Multi-query causal attention, one warp per query.  q rows [nq][q_ld] (head h at +h*64); keys/values addressed as
// base + h*head_stride + key*row_stride.  Query i sits at absolute position qpos0+i and attends keys
// max(0, pos-window+1) .. pos.
void launch_attention(const float* q, long long q_ld, const float* k, const float* v, long long kv_head_stride,
                      long long kv_row_stride, float* out, long long out_ld, int nq, int qpos0, int heads, int window,
                      cudaStream_t st, int nseg = 1, float* out_lo = nullptr);
// Only the last `c` queries of every segment (positions nq-c .. nq-1, nq <= 128 keys): one CTA per (query, head), one
// thread per key.  out is compact: row seg*c + j.  Used by the last layer of the streaming window encode, whose other
// rows nobody reads (infer_arvc.py:506-518 keeps the last `chunk` ids).
constexpr int ATT_TAIL_MAX_KEYS = 128;
void launch_attention_tail(const float* q, long long q_ld, const float* k, const float* v, long long kv_head_stride,
                           long long kv_row_stride, float* out, long long out_ld, int nq, int c, int heads, int window,
                           cudaStream_t st, int nseg);
// scatter k,v of a fused qkv buffer into a [H][max_seq][64] cache at positions pos0..pos0+rows-1
void launch_kv_append(const float* qkv, int rows, int heads, float* kc, float* vc, int max_seq, int pos0,
                      cudaStream_t st);

}  // namespace svanon
