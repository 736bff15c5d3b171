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
  // conv-stack half (with_conv): log-mel rows of the two spans [6 pad rows | 2 x (6 margin + T) rows][160] -- written by the
  // per-op front (DFT, magnitude, mel filterbank) -- and the spans the transformer half assembles the window from
  float *mel = nullptr, *spans = nullptr;
  bool with_conv = false;
  int T = 0;
  std::vector<float*> bufs;
  ~EncChain() {
    for (auto p : bufs) cudaFree(p);
  }
  float* alloc(size_t n) {
    float* p = nullptr;
    SV_CUDA(cudaMalloc(&p, n * sizeof(float)));
    SV_CUDA(cudaMemset(p, 0, n * sizeof(float)));          // margin rows are never written: they stay zero
    bufs.push_back(p);
    return p;
  }
};

struct EncChains {
  std::map<std::tuple<int, int, int, int>, std::unique_ptr<EncChain>> by_cfg;   // (S, c, Ls, tail_only + 2 * with_conv)
  unsigned* barrier = nullptr;
  ~EncChains() {
    if (barrier) cudaFree(barrier);
  }
};

enum { DYN_SPANS = 0, DYN_PREV = 1, DYN_NEXT = 2, DYN_IDS = 3 };

// The conv-stack half of the window encoder for the two spans (head: the window's first Ls frames, tail: its last Ls frames):
// ConvNeXtEncoder (firefly.py:506-517: stem, 4 stages of ConvNeXt blocks, LayerNorms) and the quantizer's two down-sampling
// blocks (bsq_no_upsample.py:46-60) -- Engine::enc_conv_stack from the stem on -- as chain phases.
//
// Layout: every activation buffer holds BOTH spans with their 6 zero margin rows (the causal left context) as ONE matrix of
// 2 x (6 + rows) rows; GEMM phases run over all those rows (margin rows compute values nobody reads), row phases skip the
// margin rows when they write into a margin-carrying buffer (ChainOp::period / margin), so the margins stay zero.  The
// stride-2 convs read two consecutive rows per output row: their inputs are written with a period of twice the output's
// (12 margin rows), which makes "input row = 2 x output row" hold across the segment boundary.
static void build_conv_ops(Engine& e, EncChain& ec, const ConvStackW& w, int T, int grid, cudaStream_t st) {
  auto& ops = ec.chain.ops;
  const int MARG = 6;
  const int P0 = MARG + T, M0 = 2 * P0;                       // stage rows
  const int dims[4] = {128, 256, 384, 512};
  auto pend = [&](const ChainOp& gemm, const float* bias) {
    ChainPend p;
    p.P = gemm.Pout; p.ks = gemm.ksplit; p.ks_stride = gemm.pout_ks_stride; p.ldp = gemm.ldp_out; p.bias = bias;
    return p;
  };
  // partial-sum buffer: the largest ksplit * M * N of the GEMM phases below
  size_t pf = chain_partial_floats(M0, 128, 7 * N_MELS, grid);
  for (int s = 0; s < 4; ++s) {
    pf = std::max(pf, chain_partial_floats(M0, 4 * dims[s], dims[s], grid));
    pf = std::max(pf, chain_partial_floats(M0, dims[s], 4 * dims[s], grid));
    if (s > 0) pf = std::max(pf, chain_partial_floats(M0, dims[s], dims[s - 1], grid));
  }
  pf = std::max(pf, chain_partial_floats(M0, 512, 1024, grid));
  float* P = ec.alloc(pf);
  float* tmp = ec.alloc((size_t)M0 * 512);
  float* hid = ec.alloc((size_t)M0 * 2048);
  auto gemm = [&](const float* A, long long a_stride, const float* W, int M, int N, int K) {
    ChainOp g;
    chain_set_gemm(g, A, a_stride, W, M, N, K, P, grid, st);
    ops.push_back(g);
    return g;
  };
  // ConvNeXtBlock.forward (firefly.py:421-440) on x (period / margin layout), result into y (its own layout; null = in place)
  auto convnext = [&](const ConvNextW& cw, float* x, int M, int period, float* y, int y_period, int y_margin) {
    const int C = cw.C;
    ChainOp d;
    d.kind = CH_DWLN; d.M = M; d.N = C; d.in.res = x; d.in.ldr = C; d.dw_w = cw.dw_w; d.dw_b = cw.dw_b; d.w = cw.ln_w; d.b = cw.ln_b;
    d.eps = 1e-6f; d.seg_rows = period; d.period = period; d.margin = MARG; d.y = tmp; d.ldy = C;
    ops.push_back(d);
    const ChainOp g1 = gemm(tmp, C, cw.pw1_w, M, 4 * C, C);
    ChainOp a1;
    a1.kind = CH_ACT; a1.act = CHA_GELU; a1.M = M; a1.N = 4 * C; a1.in = pend(g1, cw.pw1_b); a1.y = hid; a1.ldy = 4 * C;
    ops.push_back(a1);
    const ChainOp g2 = gemm(hid, 4 * C, cw.pw2_w, M, C, 4 * C);
    ChainOp a2;                                               // x + gamma * (pw2 + bias)
    a2.kind = CH_ACT; a2.act = CHA_NONE; a2.M = M; a2.N = C; a2.in = pend(g2, cw.pw2_b); a2.in.gamma = cw.gamma; a2.in.res = x; a2.in.ldr = C;
    a2.period = period; a2.margin = MARG; a2.y = y ? y : x; a2.ldy = C;
    if (y) { a2.y_period = y_period; a2.y_margin = y_margin; }
    ops.push_back(a2);
  };
  // stem: causal conv k = 7 over the mel rows as one GEMM over 7 overlapping rows, then LayerNorm
  float* x = ec.alloc((size_t)M0 * dims[0]);
  {
    const ChainOp g = gemm(ec.mel, N_MELS, w.stem_w, M0, dims[0], 7 * N_MELS);
    ChainOp n;
    n.kind = CH_NORM; n.norm = CHN_LN; n.M = M0; n.N = dims[0]; n.eps = 1e-6f; n.w = w.stem_ln_w; n.b = w.stem_ln_b;
    n.in = pend(g, w.stem_b); n.period = P0; n.margin = MARG; n.y = x; n.ldy = dims[0];
    ops.push_back(n);
  }
  for (int s = 0; s < 4; ++s) {
    const int C = dims[s];
    if (s > 0) {
      const int Cp = dims[s - 1];
      ChainOp n;                                              // LayerNorm + 1 x 1 conv into the next stage's buffer
      n.kind = CH_NORM; n.norm = CHN_LN; n.M = M0; n.N = Cp; n.eps = 1e-6f; n.w = w.mid_ln_w[s - 1]; n.b = w.mid_ln_b[s - 1];
      n.in.res = x; n.in.ldr = Cp; n.y = tmp; n.ldy = Cp;
      ops.push_back(n);
      const ChainOp g = gemm(tmp, Cp, w.mid_w[s - 1], M0, C, Cp);
      float* xn = ec.alloc((size_t)M0 * C);
      ChainOp a;
      a.kind = CH_ACT; a.act = CHA_NONE; a.M = M0; a.N = C; a.in = pend(g, w.mid_b[s - 1]); a.period = P0; a.margin = MARG; a.y = xn; a.ldy = C;
      ops.push_back(a);
      x = xn;
    }
    for (const auto& blk : w.blocks[s]) convnext(blk, x, M0, P0, nullptr, 0, 0);
  }
  // final LayerNorm into the first stride-2 conv's input layout, then 2 x [conv k2 s2 + ConvNeXt]
  int rows = T;
  float* cur = ec.alloc((size_t)2 * (2 * MARG + rows) * 512);
  {
    ChainOp n;
    n.kind = CH_NORM; n.norm = CHN_LN; n.M = M0; n.N = 512; n.eps = 1e-6f; n.w = w.bb_norm_w; n.b = w.bb_norm_b;
    n.in.res = x; n.in.ldr = 512; n.period = P0; n.margin = MARG; n.y_period = 2 * MARG + rows; n.y_margin = 2 * MARG;
    n.y = cur; n.ldy = 512;
    ops.push_back(n);
  }
  for (int i = 0; i < 2; ++i) {
    const int r2 = rows / 2, Pd = MARG + r2, Md = 2 * Pd;
    const ChainOp g = gemm(cur, 1024, w.down_w[i], Md, 512, 1024);
    float* dn = ec.alloc((size_t)Md * 512);
    ChainOp a;
    a.kind = CH_ACT; a.act = CHA_NONE; a.M = Md; a.N = 512; a.in = pend(g, w.down_b[i]); a.period = Pd; a.margin = MARG; a.y = dn; a.ldy = 512;
    ops.push_back(a);
    if (i == 0) {
      float* nxt = ec.alloc((size_t)2 * (2 * MARG + r2) * 512);
      convnext(w.down_block[i], dn, Md, Pd, nxt, 2 * MARG + r2, 2 * MARG);
      cur = nxt;
    } else {
      convnext(w.down_block[i], dn, Md, Pd, ec.spans, r2, 0);   // plain [2][Ls][512]
    }
    rows = r2;
  }
}

static std::unique_ptr<EncChain> build_enc_chain(Engine& e, EncChains& set, int S, int c, int Ls, bool tail_only, bool with_conv,
                                                 cudaStream_t st) {
  auto ec = std::make_unique<EncChain>();
  const int grid = e.num_sms;
  const int D = ENC_DIM, I = ENC_INTER;
  ec->with_conv = with_conv;
  if (with_conv) {
    ec->T = 4 * Ls;
    ec->mel = ec->alloc((size_t)(6 + 2 * (6 + ec->T)) * N_MELS);
    ec->spans = ec->alloc((size_t)2 * Ls * D);
    build_conv_ops(e, *ec, e.tok_cs, ec->T, grid, st);
  }
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
    o.in.res = with_conv ? ec->spans : chain_dyn(DYN_SPANS); o.in.ldr = D;
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
                              cudaStream_t st, const float* const* span_src, const long long* span_pitch) {
  if (!chain_supported(num_sms) || S > 128 || c < 1 || c > S) return false;
  // with the wave spans given, the conv stack runs inside the chain as well (two spans of 4 Ls mel rows: <= 384 rows in all)
  const bool with_conv = g_chain_conv && span_src && 2 * (6 + 4 * Ls) <= 128 * CHAIN_MAX_MTILES;
  if (!with_conv && !spans) return false;
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
  auto key = std::make_tuple(S, c, Ls, (tail_only ? 1 : 0) + (with_conv ? 2 : 0));
  auto& slot = enc_chains->by_cfg[key];
  if (!slot) slot = build_enc_chain(*this, *enc_chains, S, c, Ls, tail_only, with_conv, st);
  if (with_conv) {
    // per-op front: left-padded frames -> DFT GEMM -> magnitude -> mel filterbank + log, straight into the chain's mel rows
    const int T = slot->T;
    enc_conv_stack(tok_cs, span_src, span_pitch, 2, 1, (long long)T * HOP, nullptr, st, nullptr, 0,
                   slot->mel + (size_t)(6 + 6) * N_MELS, (long long)(6 + T) * N_MELS);
  }
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
  const bool ok = enc_window_chain(spans, prev, next, S, c, Ls, ids_out, st, nullptr, nullptr);
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
