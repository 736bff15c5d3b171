// Persistent "chain" kernel: one cooperative launch (one CTA per SM) walks a device-side list of dependent ops -- tcgen05
// GEMM phases and row-wise phases -- with a grid barrier between them, instead of one kernel launch per op.
//
// Why: the single-stream chunk is ~210 dependent kernels; a small GEMM launch is one latency-bound wave of CTAs that lives
// 9-13 us for 2-4 us of MMAs (prologue, griddepcontrol.wait, split-K cluster barriers, exit, launch gap: DESIGN.md section
// 7.1).  Inside one kernel the per-op fixed cost is one grid barrier (~1.8 us), TMEM and the mbarriers are set up once, and
// -- because the op list is static -- every CTA fetches the WEIGHT block of its next-but-one GEMM job into shared memory
// (cp.async.bulk from the pre-split, pre-tiled weight copies of gemm_tc.cu) while the current ops run, so a GEMM phase only
// waits for its activations.
//
// GEMM phase: job = (N tile of BN columns, K slice); the job's B block (all its K-slabs, hi and lo terms) is resident in
// shared memory; the 16 producer warps stream the A rows of ALL M tiles (<= 3 x 128 rows) of the K slice through a 3-stage
// ring (global -> registers -> hi/lo TF32 split -> 128-byte-swizzled tile), one thread issues the 3xTF32 tcgen05.mma into one
// TMEM accumulator per M tile, and the epilogue writes the job's PARTIAL tile (fp32) to P[k-slice][M][N] in global memory.
// The consumer of a GEMM is always a row-wise phase that sums the K-slice partials in fixed order (deterministic) while it
// applies bias / layer-scale / residual / activation / norm -- the split-K reduction costs no phase of its own.
#pragma once
#include <vector>

#include "common.cuh"

namespace svanon {

enum ChainKind : int { CH_GEMM = 1, CH_NORM = 2, CH_DWLN = 3, CH_ACT = 4, CH_QKV_ROPE = 5, CH_ATTN = 6, CH_BSQ = 7 };
enum ChainNorm : int { CHN_RMS = 0, CHN_LN = 1 };
enum ChainAct : int { CHA_NONE = 0, CHA_GELU = 1, CHA_SILU_MUL = 2 };
// what a GEMM phase does with its accumulators.  EPI_PARTIAL: K-slice partials to Pout (summed by the consuming row phase);
// the others need ksplit == 1 and write finished values: EPI_DIRECT y = act(res + gamma * (acc + bias)) (terms of `in`),
// EPI_ROPE the same plus RoPE on the q | k columns of a qkv row (table, q_first, heads), EPI_SILU_MUL y[:, j] =
// silu(acc[:, h1 col j]) * acc[:, h3 col j] for weights whose rows interleave 16 rows of w1 with the same 16 rows of w3.
enum ChainEpi : int { EPI_PARTIAL = 0, EPI_DIRECT = 1, EPI_ROPE = 2, EPI_SILU_MUL = 3 };

constexpr int CHAIN_DYN = 8;             // per-launch pointers: a pointer field holding 1..CHAIN_DYN means dyn[value - 1]
inline const float* chain_dyn(int slot) { return reinterpret_cast<const float*>((uintptr_t)(slot + 1)); }

// value(r, c) = res[r * ldr + c] + gamma[c] * (bias[c] + sum_{k < ks} P[k * ks_stride + r * ldp + c])   (null terms drop out)
struct ChainPend {
  const float* P = nullptr;
  const float* bias = nullptr;
  const float* gamma = nullptr;
  const float* res = nullptr;            // may be a dyn slot
  long long ks_stride = 0;
  int ks = 0, ldp = 0, ldr = 0, pad_ = 0;
};

struct alignas(16) ChainOp {
  int kind = 0;
  int M = 0, N = 0, K = 0;               // rows, columns (row ops: N = C), reduction length (GEMM)
  ChainPend in;                          // row ops: the input value
  // ---- CH_GEMM: Pout[ks][m][n] = sum over the K slice of A[m * a_row_stride + k] * W[n][k]
  const float* A = nullptr;
  const unsigned char* Wt0 = nullptr;    // pre-tiled hi / lo weight terms (gemm_tiled_weights)
  const unsigned char* Wt1 = nullptr;
  float* Pout = nullptr;
  long long a_row_stride = 0;            // floats between consecutive output rows' A rows (a_row_step * lda)
  long long pout_ks_stride = 0;
  int wt_npad = 0, BN = 0, n_tiles = 0, ksplit = 0, slabs = 0, gemm_seq = -1, ldp_out = 0;
  int no_grid_sync = 0;                  // the next op does not read this op's result (independent GEMM): CTA barrier only
  int epi = 0;                           // ChainEpi
  int slab_lo = 0;                       // first K-slab of W (and of the A rows) this op covers; `slabs` counts from there
  int acc_keep = 0;                      // no epilogue: the next GEMM op (same tiling) continues these accumulators ...
  int acc_cont = 0;                      // ... and sets this: first MMA accumulates (a K range too long for one weight block)
  // ---- row ops
  float* xout = nullptr;                 // optional: the materialised input value (may be a dyn slot)
  float* xout2 = nullptr;                // optional second copy (may be a dyn slot)
  float* y = nullptr;                    // the op's result
  const float* w = nullptr;              // norm weight
  const float* b = nullptr;              // norm bias (LN)
  int ldx = 0, ldx2 = 0, ldy = 0, norm = 0, act = 0;
  float eps = 0.f;
  // CH_DWLN: depthwise causal conv k = 7 over the rows of `in.res` (plain rows; rows before a segment start read as zero)
  const float* dw_w = nullptr;           // [7][C]
  const float* dw_b = nullptr;
  int seg_rows = 0;                      // rows per independent segment (0: one segment)
  // row phases over buffers of several segments with zero MARGIN rows in front of each (the causal convs' left context):
  // period > 0: row r belongs to segment r / period at t = r % period - margin; rows with t < 0 are margin rows and are left
  // alone; y_period > 0: the result goes to row seg * y_period + y_margin + t of y (another segment layout)
  int period = 0, margin = 0, y_period = 0, y_margin = 0;
  // CH_NORM with asm_S > 0: the window assemble of the streaming encoder (Engine::enc_window_step): row p of the value is
  // row p of `in` (p < asm_rf), prev[p + asm_c] (p < S - c), or row 2 * Ls - (S - p) of `in` (the tail span)
  int asm_S = 0, asm_Ls = 0, asm_rf = 0, asm_c = 0, pad1_ = 0;
  const float* prev = nullptr;           // may be a dyn slot
  // CH_ATTN: queries q_first .. q_first + nq - 1 of `A` = qkv [rows][3 * heads * 64] (RoPE applied), causal, window
  int heads = 0, q_first = 0, nq = 0, window = 0;
  // CH_QKV_ROPE / CH_BSQ
  const float* table = nullptr;          // RoPE table [pos][32][2]; BSQ: projection weights [13][512]
  const float* table_b = nullptr;        // BSQ bias [13]
  long long* ids = nullptr;              // may be a dyn slot
};

// this CTA's weight block of one GEMM phase (made by Chain::upload): n_sl K-slabs, per slab one `bytes` block of each term
// ... and the job's geometry, so that no CTA divides anything at the start of a GEMM phase
struct ChainWJob {
  const unsigned char* w0;
  const unsigned char* w1;
  unsigned bytes;
  int n_sl;                              // 0: this CTA has no job in the GEMM
  long long slab_stride;
  long long a_off;                       // floats from op.A to the job's first K-slab
  int n0;                                // first output column
  int ks;                                // K-slice index (partial buffer)
  int pad_[4];
};
static_assert(sizeof(ChainWJob) == 64, "ChainWJob is read as four 16-byte words");

struct ChainDyn {
  const void* p[CHAIN_DYN];
};

// A built chain: the op list on the device plus its GEMM index.
struct Chain {
  std::vector<ChainOp> ops;              // host copy
  ChainOp* ops_dev = nullptr;
  int* gemm_ops_dev = nullptr;           // op index of the q-th GEMM
  ChainWJob* wjobs_dev = nullptr;        // [n_gemm][grid]
  int n_gemm = 0, grid = 0;
  double gemm_flop = 0;                  // 2 M N K summed over the GEMM ops
  bool uploaded = false;
  ~Chain();
  void upload(int grid);
};

constexpr int CHAIN_B_BYTES = 64 * 1024;          // weight block of one GEMM job (all its K-slabs, hi + lo)
constexpr int CHAIN_MAX_MTILES = 3;

// Picks (BN, ksplit) for an M x N x K GEMM phase on `grid` CTAs; false if the shape does not fit the phase's limits.
bool chain_gemm_config(int M, int N, int K, int grid, int* BN, int* ksplit);
// Fills the GEMM fields of `op` (tiling, pre-tiled weights made on first use, partial buffer); P must hold ksplit * M * N floats.
void chain_set_gemm(ChainOp& op, const float* A, long long a_row_stride, const float* W, int M, int N, int K, float* P, int grid,
                    cudaStream_t st);
// the same with the tiling given: BN columns per job, `ksplit` K slices over the slabs [slab_lo, slab_lo + slabs) of W [N][K]
void chain_set_gemm_tiled(ChainOp& op, const float* A, long long a_row_stride, const float* W, int M, int N, int K, int BN,
                          int ksplit, int slab_lo, int slabs, int grid, cudaStream_t st);
size_t chain_partial_floats(int M, int N, int K, int grid);
extern bool g_use_chain, g_chain_conv;            // svanon_set_chain_mode / SVANON_CHAIN
bool chain_supported(int grid);
void launch_chain(Chain& c, const ChainDyn& dyn, unsigned* barrier, int grid, cudaStream_t st);

}  // namespace svanon
