// Single-stream window encoder as ONE persistent chain launch (chain.cuh): the window assemble, the 8-layer windowed RoPE
// transformer (windowed_transformer.py:337-354) and the BSQ ids (bsq.py:330-369) of Engine::enc_window_step -- 75 kernel
// launches of the per-op path -- as a list of tcgen05 GEMM phases and row-wise phases separated by grid barriers.
// Same arithmetic as enc_transformer_bsq (engine.cu): 3xTF32 products, fp32 everything else; K-slice partials are summed in
// fixed order by the consuming row phase.
#include <cstdlib>
#include <map>
#include <memory>
#include <tuple>

#include "chain.cuh"
#include "engine.hpp"

namespace svanon {

struct EncChain {
  Chain chain;
  float *x = nullptr, *nrm = nullptr, *qkv = nullptr, *y = nullptr, *g = nullptr, *P = nullptr;
  std::vector<float*> bufs;
  ~EncChain() {
    for (auto p : bufs) cudaFree(p);
  }
  float* alloc(size_t n) {
    float* p = nullptr;
    SV_CUDA(cudaMalloc(&p, n * sizeof(float)));
    bufs.push_back(p);
    return p;
  }
};

struct EncChains {
  std::map<std::tuple<int, int, int, int>, std::unique_ptr<EncChain>> by_cfg;   // (S, c, Ls, tail_only)
  unsigned* barrier = nullptr;
  ~EncChains() {
    if (barrier) cudaFree(barrier);
  }
};

enum { DYN_SPANS = 0, DYN_PREV = 1, DYN_NEXT = 2, DYN_IDS = 3 };

static std::unique_ptr<EncChain> build_enc_chain(Engine& e, EncChains& set, int S, int c, int Ls, bool tail_only, cudaStream_t st) {
  auto ec = std::make_unique<EncChain>();
  const int grid = e.num_sms;
  const int D = ENC_DIM, I = ENC_INTER;
  ec->x = ec->alloc((size_t)S * D);
  ec->nrm = ec->alloc((size_t)S * D);
  ec->qkv = ec->alloc((size_t)S * 3 * D);
  ec->y = ec->alloc((size_t)S * D);
  ec->g = ec->alloc((size_t)S * I);
  const int R = tail_only ? c : S;
  size_t pf = 0;
  pf = std::max(pf, chain_partial_floats(S, 3 * D, D, grid));
  for (int rows : {S, R}) {
    pf = std::max(pf, chain_partial_floats(rows, D, D, grid));
    pf = std::max(pf, 2 * chain_partial_floats(rows, I, D, grid));
    pf = std::max(pf, chain_partial_floats(rows, D, I, grid));
  }
  ec->P = ec->alloc(pf);
  auto& ops = ec->chain.ops;
  static const bool fused = [] { const char* v = getenv("SVANON_CHAIN_FUSE"); return !v || atoi(v) != 0; }();   // A/B knob
  auto pend = [&](const ChainOp& gemm) {
    ChainPend p;
    p.P = gemm.Pout; p.ks = gemm.ksplit; p.ks_stride = gemm.pout_ks_stride; p.ldp = gemm.ldp_out;
    return p;
  };
  {
    ChainOp o;                            // window assemble + attention norm of layer 0
    o.kind = CH_NORM; o.M = S; o.N = D; o.norm = CHN_RMS; o.eps = 1e-5f; o.w = e.enc_layers[0].attn_norm;
    o.in.res = chain_dyn(DYN_SPANS); o.in.ldr = D;
    o.prev = chain_dyn(DYN_PREV);
    o.asm_S = S; o.asm_Ls = Ls; o.asm_rf = ENC_RF; o.asm_c = c;
    o.xout = const_cast<float*>(chain_dyn(DYN_NEXT)); o.ldx = D;
    o.xout2 = ec->x; o.ldx2 = D;
    o.y = ec->nrm; o.ldy = D;
    ops.push_back(o);
  }
  for (int l = 0; l < ENC_LAYERS; ++l) {
    const EncLayerW& L = e.enc_layers[l];
    const bool tail = tail_only && l == ENC_LAYERS - 1;
    const int rows = tail ? c : S, off = tail ? S - c : 0;
    if (fused) {
      // qkv with the whole K range per job (16 columns x 512: one 64 KB weight block): RoPE in the GEMM's epilogue
      ChainOp gq;
      chain_set_gemm_tiled(gq, ec->nrm, D, L.wqkv, S, 3 * D, D, 16, 1, 0, D / 32, grid, st);
      gq.epi = EPI_ROPE; gq.heads = ENC_HEADS; gq.table = e.enc_rope; gq.q_first = 0; gq.y = ec->qkv; gq.ldy = 3 * D;
      ops.push_back(gq);
    } else {
      ChainOp gq;
      chain_set_gemm(gq, ec->nrm, D, L.wqkv, S, 3 * D, D, ec->P, grid, st);
      ops.push_back(gq);
      ChainOp o;
      o.kind = CH_QKV_ROPE; o.M = S; o.N = 3 * D; o.in = pend(gq); o.heads = ENC_HEADS; o.table = e.enc_rope; o.q_first = 0;
      o.y = ec->qkv; o.ldy = 3 * D;
      ops.push_back(o);
    }
    {
      ChainOp o;
      o.kind = CH_ATTN; o.A = ec->qkv; o.heads = ENC_HEADS; o.q_first = off; o.nq = rows; o.window = ENC_WINDOW;
      o.y = ec->y; o.ldy = D;
      ops.push_back(o);
    }
    ChainOp go;
    chain_set_gemm(go, ec->y + (size_t)off * D, D, L.wo, rows, D, D, ec->P, grid, st);
    ops.push_back(go);
    {
      ChainOp o;                          // x += ls_attn * wo(...); ffn norm
      o.kind = CH_NORM; o.M = rows; o.N = D; o.norm = CHN_RMS; o.eps = 1e-5f; o.w = L.ffn_norm;
      o.in = pend(go); o.in.gamma = L.ls_attn; o.in.res = ec->x + (size_t)off * D; o.in.ldr = D;
      o.xout = ec->x + (size_t)off * D; o.ldx = D;
      o.y = ec->nrm; o.ldy = D;
      ops.push_back(o);
    }
    // w1 and w3: two independent GEMM phases with no grid barrier in between (their weights together exceed one CTA's weight
    // buffer), partials side by side in P[ks][rows][2 I]
    if (fused) {
      // w1 | w3 as ONE weight matrix whose rows interleave 16 rows of w1 with the same 16 rows of w3: a 32-column tile holds
      // h1 and h3 of 16 hidden units, so silu(h1) * h3 happens in the epilogue.  The K range is two weight blocks long: two
      // GEMM ops on the same accumulators, no barrier in between.
      float* w13 = ec->alloc((size_t)2 * I * D);
      SV_CUDA(cudaMemcpy2DAsync(w13, (size_t)32 * D * sizeof(float), L.w1, (size_t)16 * D * sizeof(float),
                                (size_t)16 * D * sizeof(float), I / 16, cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpy2DAsync(w13 + (size_t)16 * D, (size_t)32 * D * sizeof(float), L.w3, (size_t)16 * D * sizeof(float),
                                (size_t)16 * D * sizeof(float), I / 16, cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaStreamSynchronize(st));
      ChainOp ga, gb;
      chain_set_gemm_tiled(ga, ec->nrm, D, w13, rows, 2 * I, D, 32, 1, 0, D / 64, grid, st);
      chain_set_gemm_tiled(gb, ec->nrm, D, w13, rows, 2 * I, D, 32, 1, D / 64, D / 64, grid, st);
      ga.acc_keep = 1; ga.no_grid_sync = 1;
      gb.acc_cont = 1; gb.epi = EPI_SILU_MUL; gb.y = ec->g; gb.ldy = I;
      ops.push_back(ga);
      ops.push_back(gb);
    } else {
    ChainOp g1, g3;
    chain_set_gemm(g1, ec->nrm, D, L.w1, rows, I, D, ec->P, grid, st);
    chain_set_gemm(g3, ec->nrm, D, L.w3, rows, I, D, ec->P + I, grid, st);
    g1.ldp_out = g3.ldp_out = 2 * I;
    g1.pout_ks_stride = g3.pout_ks_stride = (long long)rows * 2 * I;
    g1.no_grid_sync = 1;
    ops.push_back(g1);
    ops.push_back(g3);
    {
      ChainOp o;
      o.kind = CH_ACT; o.act = CHA_SILU_MUL; o.M = rows; o.N = I; o.in = pend(g1); o.y = ec->g; o.ldy = I;
      ops.push_back(o);
    }
    }
    ChainOp g2;
    chain_set_gemm(g2, ec->g, I, L.w2, rows, D, I, ec->P, grid, st);
    ops.push_back(g2);
    {
      ChainOp o;                          // x += ls_ffn * w2(...); next layer's attention norm, or the final norm + BSQ
      const bool last = l == ENC_LAYERS - 1;
      o.kind = last ? CH_BSQ : CH_NORM; o.M = rows; o.N = D; o.norm = CHN_RMS; o.eps = 1e-5f;
      o.w = last ? e.enc_norm_w : e.enc_layers[l + 1].attn_norm;
      o.in = pend(g2); o.in.gamma = L.ls_ffn; o.in.res = ec->x + (size_t)off * D; o.in.ldr = D;
      if (!last) { o.xout = ec->x; o.ldx = D; o.y = ec->nrm; o.ldy = D; }
      else { o.y = ec->nrm; o.ldy = D;   // the final-norm rows (read back by the test hook)
             o.table = e.bsq_w; o.table_b = e.bsq_b; o.ids = reinterpret_cast<long long*>(const_cast<float*>(chain_dyn(DYN_IDS))); o.q_first = off; }
      ops.push_back(o);
    }
  }
  ec->chain.upload(grid);
  return ec;
}

// The transformer half of Engine::enc_window_step for one stream: spans [2 * Ls][512] (conv-stack outputs of the window's
// first and last Ls frames), prev / next = the window state of the previous / this chunk, ids [S] (the last c, or all S, are
// written).  False: not applicable (the caller runs the per-op path).
bool Engine::enc_window_chain(const float* spans, const float* prev, float* next, int S, int c, int Ls, long long* ids,
                              cudaStream_t st) {
  if (!chain_supported(num_sms) || S > 128 || c < 1 || c > S) return false;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return false;                        // a cooperative launch cannot be captured
  }
  if (!enc_chains) {
    enc_chains = std::make_shared<EncChains>();
    SV_CUDA(cudaMalloc(&enc_chains->barrier, 64 * sizeof(unsigned)));
    SV_CUDA(cudaMemset(enc_chains->barrier, 0, 64 * sizeof(unsigned)));
  }
  const bool tail_only = enc_tail_only && c < S;
  auto key = std::make_tuple(S, c, Ls, tail_only ? 1 : 0);
  auto& slot = enc_chains->by_cfg[key];
  if (!slot) slot = build_enc_chain(*this, *enc_chains, S, c, Ls, tail_only, st);
  ChainDyn dyn{};
  dyn.p[DYN_SPANS] = spans;
  dyn.p[DYN_PREV] = prev;
  dyn.p[DYN_NEXT] = next;
  dyn.p[DYN_IDS] = ids;
  launch_chain(slot->chain, dyn, enc_chains->barrier, num_sms, st);
  return true;
}

// Test hook: the transformer + BSQ of one window xt [S][512] through the chain (use_chain) or the per-op path; hidden_out = the
// final-norm rows whose ids are produced (keep > 0: the last `keep` rows), ids_out [S] (only those columns are written).
void Engine::debug_enc_transformer(const float* xt, int S, int keep, bool use_chain, float* hidden_out, long long* ids_out,
                                   cudaStream_t st) {
  SV_CHECK(finalized[MODEL_TOKENIZER], "tokenizer weights not finalized");
  SV_CHECK(S >= 1 && S <= 128 && keep >= 0 && keep <= S, "debug_enc_transformer: S <= 128");
  const int c = keep > 0 ? keep : S;
  const bool tail = enc_tail_only && keep > 0 && keep < S;
  const int rows = tail ? c : S;
  ws.ensure(((size_t)S * 16000 + (4u << 20)) * sizeof(float));
  ws.reset();
  if (!use_chain) {
    float* x = ws.alloc_f((long long)S * ENC_DIM);
    SV_CUDA(cudaMemcpyAsync(x, xt, (size_t)S * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
    enc_transformer_bsq(x, 1, S, ids_out, st, keep, hidden_out);
    return;
  }
  // a (spans, prev) pair whose assembled window is xt: head rows = xt[0, rf), prev[p + c] = xt[p], tail rows = xt[S - c, S)
  const int Ls = ENC_RF + c;
  float* spans = ws.alloc_f((long long)2 * Ls * ENC_DIM);
  float* prev = ws.alloc_f((long long)(S + c) * ENC_DIM);
  float* next = ws.alloc_f((long long)S * ENC_DIM);
  SV_CUDA(cudaMemsetAsync(spans, 0, (size_t)2 * Ls * ENC_DIM * sizeof(float), st));
  const int rf = std::min(ENC_RF, S);
  SV_CUDA(cudaMemcpyAsync(spans, xt, (size_t)rf * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SV_CUDA(cudaMemcpyAsync(prev + (size_t)c * ENC_DIM, xt, (size_t)S * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  SV_CUDA(cudaMemcpyAsync(spans + (size_t)(2 * Ls - c) * ENC_DIM, xt + (size_t)(S - c) * ENC_DIM, (size_t)c * ENC_DIM * sizeof(float),
                          cudaMemcpyDeviceToDevice, st));
  const bool was = g_use_chain;
  g_use_chain = true;
  const bool ok = enc_window_chain(spans, prev, next, S, c, Ls, ids_out, st);
  g_use_chain = was;
  SV_CHECK(ok, "chain not applicable here");
  if (hidden_out) {
    auto& ec = enc_chains->by_cfg[std::make_tuple(S, c, Ls, tail ? 1 : 0)];
    SV_CUDA(cudaMemcpyAsync(hidden_out, ec->nrm, (size_t)rows * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  (void)next;
}

// Test hook: C = act(A W^T + bias) as `repeat` x [GEMM phase, element-wise phase] of one chain launch (A [M][K], W [N][K]).
void Engine::debug_chain_gemm(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, int act, int repeat,
                              cudaStream_t st) {
  SV_CHECK(repeat >= 1 && repeat <= 16, "repeat");
  if (!enc_chains) {
    enc_chains = std::make_shared<EncChains>();
    SV_CUDA(cudaMalloc(&enc_chains->barrier, 64 * sizeof(unsigned)));
    SV_CUDA(cudaMemset(enc_chains->barrier, 0, 64 * sizeof(unsigned)));
  }
  Chain ch;
  float* P = nullptr;
  SV_CUDA(cudaMalloc(&P, chain_partial_floats(M, N, K, num_sms) * sizeof(float)));
  for (int r = 0; r < repeat; ++r) {
    ChainOp g;
    chain_set_gemm(g, A, K, W, M, N, K, P, num_sms, st);
    ch.ops.push_back(g);
    ChainOp o;
    o.kind = CH_ACT; o.act = act ? CHA_GELU : CHA_NONE; o.M = M; o.N = N;
    o.in.P = g.Pout; o.in.ks = g.ksplit; o.in.ks_stride = g.pout_ks_stride; o.in.ldp = g.ldp_out; o.in.bias = bias;
    o.y = C; o.ldy = N;
    ch.ops.push_back(o);
  }
  ch.upload(num_sms);
  ChainDyn dyn{};
  launch_chain(ch, dyn, enc_chains->barrier, num_sms, st);
  SV_CUDA(cudaStreamSynchronize(st));
  cudaFree(P);
  gemm_forget_weights(W);
}

}  // namespace svanon
