// Engine: weight store + packing, workspace, and the E / V / A(prefill) stage drivers.
#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace svanon {

long long g_kernel_launches = 0;

// ------------------------------------------------------------------------------------------ workspace
void Workspace::ensure(size_t bytes) {
  if (bytes <= cap) return;
  SV_CUDA(cudaDeviceSynchronize());
  if (base) SV_CUDA(cudaFree(base));
  base = nullptr;
  cap = 0;
  const size_t want = (bytes + (64u << 20)) & ~((size_t)(1u << 20) - 1);
  SV_CUDA(cudaMalloc(&base, want));
  cap = want;
}
void* Workspace::alloc_bytes(size_t bytes) {
  const size_t a = (off + 255) & ~(size_t)255;
  SV_CHECK(a + bytes <= cap, "workspace overflow");
  off = a + bytes;
  return base + a;
}
Workspace::~Workspace() {
  if (base) cudaFree(base);
}

Stream::~Stream() {
  for (void* p : {(void*)kc, (void*)vc, (void*)fkc, (void*)fvc, (void*)x_audio, (void*)ref_emb_tail, (void*)spk_rows,
                  (void*)codes_dev, (void*)content_id_dev, (void*)noise_dev, (void*)wave_ring, (void*)wave_ring_tmp,
                  (void*)src_hist, (void*)pred_hist, (void*)ref_content_dev, (void*)ref_audio_dev, (void*)style_dev,
                  (void*)timbre_dev, (void*)ids_win_dev, (void*)codes_win_dev, (void*)wave_win_dev})
    if (p) cudaFree(p);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
}

Engine::~Engine() {
  for (int m = 0; m < MODEL_COUNT; ++m)
    for (auto& kv : w[m]) cudaFree(kv.second.data);
  for (float* p : owned) cudaFree(p);
  for (void* p : {(void*)ar_x, (void*)ar_h, (void*)ar_q, (void*)ar_g, (void*)ar_part, (void*)ar_logits,
                  (void*)ar_barrier, (void*)dbg_slow_logits, (void*)dbg_hidden, (void*)dbg_fast_logits})
    if (p) cudaFree(p);
  for (auto& l : lo_scr) if (l.p) cudaFree(l.p);
  if (own_stream) cudaStreamDestroy(own_stream);
}

float* Engine::lo_scratch(int slot, size_t n, cudaStream_t st) {
  LoScratch& s = lo_scr[slot];
  if (n <= s.cap) return s.p;
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  SV_CUDA(cudaDeviceSynchronize());
  if (s.p) cudaFree(s.p);
  s.p = nullptr;
  s.cap = 0;
  SV_CUDA(cudaMalloc(&s.p, (n + n / 8) * sizeof(float)));
  s.cap = n + n / 8;
  return s.p;
}

const Tensor& Engine::get(int model, const std::string& name) const {
  auto it = w[model].find(name);
  if (it == w[model].end()) throw Error("missing tensor '" + name + "' in model " + std::to_string(model));
  return it->second;
}

float* Engine::dev_alloc(long long n) {
  float* p = nullptr;
  SV_CUDA(cudaMalloc(&p, (size_t)std::max<long long>(n, 1) * sizeof(float)));
  owned.push_back(p);
  return p;
}

float* Engine::upload(const std::vector<float>& host) {
  float* p = dev_alloc((long long)host.size());
  SV_CUDA(cudaMemcpy(p, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
  return p;
}

void Engine::load_tensor(int model, const std::string& name, const float* data, int rank, const long long* shape) {
  SV_CHECK(model >= 0 && model < MODEL_COUNT, "bad model id");
  SV_CHECK(!finalized[model], "model already finalized");
  Tensor t;
  t.shape.assign(shape, shape + rank);
  const long long n = t.numel();
  SV_CHECK(n > 0, "empty tensor");
  SV_CUDA(cudaMalloc(&t.data, (size_t)n * sizeof(float)));
  SV_CUDA(cudaMemcpy(t.data, data, (size_t)n * sizeof(float), cudaMemcpyDefault));
  auto it = w[model].find(name);
  if (it != w[model].end()) {
    cudaFree(it->second.data);
    w[model].erase(it);
  }
  w[model].emplace(name, std::move(t));
}

namespace {

std::vector<float> to_host(const Tensor& t) {
  std::vector<float> h((size_t)t.numel());
  SV_CUDA(cudaMemcpy(h.data(), t.data, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
  return h;
}

void expect_shape(const Tensor& t, std::initializer_list<long long> shp, const std::string& name) {
  if (t.shape != std::vector<long long>(shp)) throw Error("tensor '" + name + "' has an unexpected shape");
}

}  // namespace

void Engine::finalize(int model) {
  SV_CHECK(model >= 0 && model < MODEL_COUNT, "bad model id");
  if (finalized[model]) return;
  if (model == MODEL_AR) finalize_ar();
  else if (model == MODEL_TOKENIZER) finalize_tokenizer();
  else if (model == MODEL_VOCODER) finalize_vocoder();
  else if (model == MODEL_STYLE) finalize_style();
  else finalize_timbre();
  finalized[model] = true;
}

// ------------------------------------------------------------------------------------------ AR weights
void Engine::finalize_ar() {
  auto g = [&](const std::string& n) { return get(MODEL_AR, n).data; };
  auto layer = [&](const std::string& p) {
    ArLayerWeights lw;
    expect_shape(get(MODEL_AR, p + ".attention.wqkv.weight"), {3 * AR_DIM, AR_DIM}, p);
    expect_shape(get(MODEL_AR, p + ".feed_forward.w2.weight"), {AR_DIM, AR_INTER}, p);
    lw.attn_norm = g(p + ".attention_norm.weight");
    lw.wqkv = g(p + ".attention.wqkv.weight");
    lw.wo = g(p + ".attention.wo.weight");
    lw.ffn_norm = g(p + ".ffn_norm.weight");
    lw.w1 = g(p + ".feed_forward.w1.weight");
    lw.w3 = g(p + ".feed_forward.w3.weight");
    lw.w2 = g(p + ".feed_forward.w2.weight");
    return lw;
  };
  for (int i = 0; i < AR_LAYERS; ++i) ar.slow[i] = layer("decoder.model.layers." + std::to_string(i));
  for (int i = 0; i < AR_FAST_LAYERS; ++i) ar.fast[i] = layer("decoder.model.fast_layers." + std::to_string(i));
  ar.norm_w = g("decoder.model.norm.weight");
  ar.output_w = g("decoder.model.output.weight");
  ar.fast_norm_w = g("decoder.model.fast_norm.weight");
  ar.fast_output_w = g("decoder.model.fast_output.weight");
  ar.fast_emb = g("decoder.model.fast_embeddings.weight");
  ar.codebook_emb = g("decoder.model.codebook_embeddings.weight");
  ar.cond_emb = g("embedding.weight");
  expect_shape(get(MODEL_AR, "decoder.model.freqs_cis"), {AR_MAX_SEQ, HEAD_DIM / 2, 2}, "decoder.model.freqs_cis");
  expect_shape(get(MODEL_AR, "decoder.model.fast_freqs_cis"), {AR_CODEBOOKS, HEAD_DIM / 2, 2}, "fast_freqs_cis");
  ar.rope = g("decoder.model.freqs_cis");
  ar.fast_rope = g("decoder.model.fast_freqs_cis");
  ctx_w = g("context_in.weight"); ctx_b = g("context_in.bias");
  style_w = g("style_in.weight"); style_b = g("style_in.bias");
  w4s = g("decoder.wait4start_embedding.weight");
  w4e = g("decoder.wait4end_embedding.weight");

  const int B = AR_MAX_BATCH;
  ar_x = nullptr;
  SV_CUDA(cudaMalloc(&ar_x, (size_t)2 * B * AR_DIM * 4));
  SV_CUDA(cudaMalloc(&ar_h, (size_t)2 * B * AR_DIM * 4));
  SV_CUDA(cudaMalloc(&ar_q, (size_t)2 * B * AR_DIM * 4));
  SV_CUDA(cudaMalloc(&ar_g, (size_t)2 * B * AR_INTER * 4));
  SV_CUDA(cudaMalloc(&ar_part, (size_t)B * AR_HEADS * 16 * 2 * (2 + HEAD_DIM) * 4));
  SV_CUDA(cudaMalloc(&ar_logits, (size_t)B * 1024 * 4));
  SV_CUDA(cudaMalloc(&ar_barrier, 1024 * sizeof(unsigned)));
  SV_CUDA(cudaMemset(ar_barrier, 0, 1024 * sizeof(unsigned)));
  SV_CUDA(cudaMalloc(&dbg_slow_logits, AR_VOCAB * 4));
  SV_CUDA(cudaMalloc(&dbg_hidden, AR_DIM * 4));
  SV_CUDA(cudaMalloc(&dbg_fast_logits, AR_CODEBOOKS * AR_CB_SIZE * 4));
  ar.x = ar_x; ar.h = ar_h; ar.q = ar_q; ar.g = ar_g; ar.part = ar_part; ar.logits = ar_logits;
  ar.barrier = ar_barrier;
}

// ------------------------------------------------------------------------------------------ shared packers
namespace {

// Conv1d weight [Co][Ci][k] -> [Co][k][Ci]  (one GEMM row = k consecutive channels-last input rows)
std::vector<float> pack_conv_rows(const std::vector<float>& w, int Co, int Ci, int k) {
  std::vector<float> o((size_t)Co * k * Ci);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int j = 0; j < k; ++j) o[((size_t)co * k + j) * Ci + ci] = w[((size_t)co * Ci + ci) * k + j];
  return o;
}
// Conv1d weight [Co][Ci][k] -> [k][Co][Ci]  (one GEMM tap per kernel element)
std::vector<float> pack_conv_taps(const std::vector<float>& w, int Co, int Ci, int k) {
  std::vector<float> o((size_t)Co * k * Ci);
  for (int co = 0; co < Co; ++co)
    for (int ci = 0; ci < Ci; ++ci)
      for (int j = 0; j < k; ++j) o[((size_t)j * Co + co) * Ci + ci] = w[((size_t)co * Ci + ci) * k + j];
  return o;
}
// depthwise [C][1][7] -> [7][C]
std::vector<float> pack_dw(const std::vector<float>& w, int C) {
  std::vector<float> o((size_t)7 * C);
  for (int c = 0; c < C; ++c)
    for (int j = 0; j < 7; ++j) o[(size_t)j * C + c] = w[(size_t)c * 7 + j];
  return o;
}
// ConvTranspose1d weight [Ci][Co][k], stride s.
//  k == 2s: out[t*s+r][co] = sum_ci x[t-1][ci] w[ci][co][r+s] + x[t][ci] w[ci][co][r]   (firefly.py:114-138)
//           -> W'[n = r*Co+co][tap*Ci+ci], tap 0 = row t-1, tap 1 = row t
//  k == s : out[t*s+r][co] = sum_ci x[t][ci] w[ci][co][r] -> W'[n][ci]
std::vector<float> pack_tconv(const std::vector<float>& w, int Ci, int Co, int k, int s) {
  const int taps = k / s;
  std::vector<float> o((size_t)s * Co * taps * Ci);
  for (int r = 0; r < s; ++r)
    for (int co = 0; co < Co; ++co)
      for (int ci = 0; ci < Ci; ++ci) {
        const size_t n = (size_t)r * Co + co;
        if (taps == 2) {
          o[n * 2 * Ci + ci] = w[((size_t)ci * Co + co) * k + r + s];
          o[n * 2 * Ci + Ci + ci] = w[((size_t)ci * Co + co) * k + r];
        } else {
          o[n * Ci + ci] = w[((size_t)ci * Co + co) * k + r];
        }
      }
  return o;
}
std::vector<float> tile_bias(const std::vector<float>& b, int s) {
  std::vector<float> o;
  for (int r = 0; r < s; ++r) o.insert(o.end(), b.begin(), b.end());
  return o;
}

}  // namespace

static ConvNextW pack_convnext(Engine& e, int model, const std::string& p, int C) {
  ConvNextW cw;
  cw.C = C;
  cw.gamma = e.get(model, p + ".gamma").data;
  cw.dw_w = e.upload(pack_dw(to_host(e.get(model, p + ".dwconv.conv.weight")), C));
  cw.dw_b = e.get(model, p + ".dwconv.conv.bias").data;
  cw.ln_w = e.get(model, p + ".norm.weight").data;
  cw.ln_b = e.get(model, p + ".norm.bias").data;
  expect_shape(e.get(model, p + ".pwconv1.weight"), {4 * C, C}, p);
  cw.pw1_w = e.get(model, p + ".pwconv1.weight").data;
  cw.pw1_b = e.get(model, p + ".pwconv1.bias").data;
  cw.pw2_w = e.get(model, p + ".pwconv2.weight").data;
  cw.pw2_b = e.get(model, p + ".pwconv2.bias").data;
  return cw;
}

// ------------------------------------------------------------------------------------------ tokenizer weights
// windowed DFT basis (rows 0..1024 real, 1025..2049 imaginary; torch.hann_window(2048) is periodic) and the slaney mel
// filterbank, shared by the two LogMelSpectrogram instances (identical parameters in both YAMLs)
void Engine::build_spectrogram_consts(int M) {
  if (!dft_w) {
    std::vector<float> dft((size_t)2 * N_FREQ * N_FFT);
    const double two_pi = 6.283185307179586476925286766559;
    std::vector<double> win(N_FFT);
    for (int n = 0; n < N_FFT; ++n) win[n] = 0.5 - 0.5 * std::cos(two_pi * n / N_FFT);
    for (int f = 0; f < N_FREQ; ++f)
      for (int n = 0; n < N_FFT; ++n) {
        const long long ph = ((long long)f * n) % N_FFT;
        const double a = two_pi * (double)ph / N_FFT;
        dft[(size_t)f * N_FFT + n] = (float)(win[n] * std::cos(a));
        dft[(size_t)(N_FREQ + f) * N_FFT + n] = (float)(-win[n] * std::sin(a));
      }
    dft_w = upload(dft);
  }
  if (!fb_t) {
    const Tensor& fb = get(M, "spec_transform.fb");
    expect_shape(fb, {N_FREQ, N_MELS}, "spec_transform.fb");
    auto h = to_host(fb);
    std::vector<float> t((size_t)N_MELS * N_FREQ_PAD, 0.f);
    for (int f = 0; f < N_FREQ; ++f)
      for (int m = 0; m < N_MELS; ++m) t[(size_t)m * N_FREQ_PAD + f] = h[(size_t)f * N_MELS + m];
    fb_t = upload(t);
  }
}

// ConvNeXtEncoder (firefly.py:443-517) + quantizer.downsample of model M (same key names in both checkpoints)
void Engine::pack_conv_stack(int M, ConvStackW& cs) {
  auto g = [&](const std::string& n) { return get(M, n).data; };
  const int dims[4] = {128, 256, 384, 512};
  const int depths[4] = {3, 3, 9, 3};
  expect_shape(get(M, "backbone.downsample_layers.0.0.conv.weight"), {128, N_MELS, 7}, "stem");
  cs.stem_w = upload(pack_conv_rows(to_host(get(M, "backbone.downsample_layers.0.0.conv.weight")), 128, N_MELS, 7));
  cs.stem_b = g("backbone.downsample_layers.0.0.conv.bias");
  cs.stem_ln_w = g("backbone.downsample_layers.0.1.weight");
  cs.stem_ln_b = g("backbone.downsample_layers.0.1.bias");
  for (int i = 1; i < 4; ++i) {
    const std::string p = "backbone.downsample_layers." + std::to_string(i);
    cs.mid_ln_w[i - 1] = g(p + ".0.weight");
    cs.mid_ln_b[i - 1] = g(p + ".0.bias");
    expect_shape(get(M, p + ".1.weight"), {dims[i], dims[i - 1], 1}, p);
    cs.mid_w[i - 1] = g(p + ".1.weight");
    cs.mid_b[i - 1] = g(p + ".1.bias");
  }
  for (int s = 0; s < 4; ++s) {
    cs.blocks[s].clear();
    for (int j = 0; j < depths[s]; ++j)
      cs.blocks[s].push_back(pack_convnext(*this, M, "backbone.stages." + std::to_string(s) + "." + std::to_string(j), dims[s]));
  }
  cs.bb_norm_w = g("backbone.norm.weight");
  cs.bb_norm_b = g("backbone.norm.bias");
  for (int i = 0; i < 2; ++i) {
    const std::string p = "quantizer.downsample." + std::to_string(i);
    expect_shape(get(M, p + ".0.conv.weight"), {ENC_DIM, ENC_DIM, 2}, p);
    cs.down_w[i] = upload(pack_conv_rows(to_host(get(M, p + ".0.conv.weight")), ENC_DIM, ENC_DIM, 2));
    cs.down_b[i] = g(p + ".0.conv.bias");
    cs.down_block[i] = pack_convnext(*this, M, p + ".1", ENC_DIM);
  }
  cs.ready = true;
}

void Engine::finalize_tokenizer() {
  const int M = MODEL_TOKENIZER;
  auto g = [&](const std::string& n) { return get(M, n).data; };
  build_spectrogram_consts(M);
  pack_conv_stack(M, tok_cs);
  for (int i = 0; i < ENC_LAYERS; ++i) {
    const std::string p = "quantizer.pre_module.layers." + std::to_string(i);
    EncLayerW& l = enc_layers[i];
    expect_shape(get(M, p + ".attention.wqkv.weight"), {3 * ENC_DIM, ENC_DIM}, p);
    l.attn_norm = g(p + ".attention_norm.weight");
    l.wqkv = g(p + ".attention.wqkv.weight");
    l.wo = g(p + ".attention.wo.weight");
    l.ffn_norm = g(p + ".ffn_norm.weight");
    l.w1 = g(p + ".feed_forward.w1.weight");
    l.w3 = g(p + ".feed_forward.w3.weight");
    l.w2 = g(p + ".feed_forward.w2.weight");
    l.ls_attn = g(p + ".attention_layer_scale.gamma");
    l.ls_ffn = g(p + ".ffn_layer_scale.gamma");
  }
  enc_norm_w = g("quantizer.pre_module.norm.weight");
  expect_shape(get(M, "quantizer.pre_module.freqs_cis"), {2048, HEAD_DIM / 2, 2}, "quantizer.pre_module.freqs_cis");
  enc_rope = g("quantizer.pre_module.freqs_cis");
  if (has(M, "quantizer.pre_module.freqs_cis_stream")) {
    expect_shape(get(M, "quantizer.pre_module.freqs_cis_stream"), {ENC_ROPE_STREAM_ROWS, HEAD_DIM / 2, 2}, "freqs_cis_stream");
    enc_rope_stream = g("quantizer.pre_module.freqs_cis_stream");
  }
  expect_shape(get(M, "quantizer.residual_bsq.rvqs.0.project_in.weight"), {BSQ_BITS, ENC_DIM}, "bsq project_in");
  bsq_w = g("quantizer.residual_bsq.rvqs.0.project_in.weight");
  bsq_b = g("quantizer.residual_bsq.rvqs.0.project_in.bias");
}

// ------------------------------------------------------------------------------------------ vocoder weights
void Engine::finalize_vocoder() {
  const int M = MODEL_VOCODER;
  // fold weight norm: w = g * v / ||v||_(1,2)  (remove_parametrizations, infer_arvc.py:94)
  {
    std::vector<std::string> bases;
    const std::string suf = ".parametrizations.weight.original1";
    for (auto& kv : w[M])
      if (kv.first.size() > suf.size() && kv.first.compare(kv.first.size() - suf.size(), suf.size(), suf) == 0)
        bases.push_back(kv.first.substr(0, kv.first.size() - suf.size()));
    for (auto& base : bases) {
      const Tensor& tv = get(M, base + ".parametrizations.weight.original1");
      const Tensor& tg = get(M, base + ".parametrizations.weight.original0");
      auto v = to_host(tv);
      auto gg = to_host(tg);
      const long long n0 = tv.shape[0], inner = tv.numel() / n0;
      SV_CHECK(tg.numel() == n0, "weight-norm g shape");
      for (long long i = 0; i < n0; ++i) {
        double nrm = 0;
        for (long long j = 0; j < inner; ++j) nrm += (double)v[i * inner + j] * v[i * inner + j];
        const float sc = (float)(gg[i] / std::sqrt(nrm));
        for (long long j = 0; j < inner; ++j) v[i * inner + j] *= sc;
      }
      finalized[M] = false;
      load_tensor(M, base + ".weight", v.data(), (int)tv.shape.size(), tv.shape.data());
    }
  }
  auto g = [&](const std::string& n) { return get(M, n).data; };
  {
    std::vector<float> fw, fbv;
    for (int gi = 0; gi < 8; ++gi) {
      const std::string p = "quantizer.residual_fsq.rvqs." + std::to_string(gi) + ".project_out";
      expect_shape(get(M, p + ".weight"), {64, 4}, p);
      auto a = to_host(get(M, p + ".weight"));
      auto b = to_host(get(M, p + ".bias"));
      fw.insert(fw.end(), a.begin(), a.end());
      fbv.insert(fbv.end(), b.begin(), b.end());
    }
    fsq_w = upload(fw);
    fsq_b = upload(fbv);
  }
  for (int i = 0; i < 2; ++i) {
    const std::string p = "quantizer.upsample." + std::to_string(i);
    expect_shape(get(M, p + ".0.conv.weight"), {512, 512, 2}, p);
    up_w[i] = upload(pack_tconv(to_host(get(M, p + ".0.conv.weight")), 512, 512, 2, 2));
    up_b[i] = upload(tile_bias(to_host(get(M, p + ".0.conv.bias")), 2));
    up_block[i] = pack_convnext(*this, M, p + ".1", 512);
  }
  expect_shape(get(M, "head.conv_pre.conv.weight"), {512, 512, 13}, "conv_pre");
  pre_w = upload(pack_conv_rows(to_host(get(M, "head.conv_pre.conv.weight")), 512, 512, 13));
  pre_b = g("head.conv_pre.conv.bias");
  const int ch[6] = {512, 256, 128, 64, 32, 16};
  const int upk[5] = {16, 16, 4, 4, 4}, ups[5] = {8, 8, 2, 2, 2};
  const int rk[3] = {3, 7, 11}, rd[3] = {1, 3, 5};
  for (int i = 0; i < 5; ++i) {
    const std::string p = "head.ups." + std::to_string(i) + ".conv";
    expect_shape(get(M, p + ".weight"), {ch[i], ch[i + 1], upk[i]}, p);
    ups_w[i] = upload(pack_tconv(to_host(get(M, p + ".weight")), ch[i], ch[i + 1], upk[i], ups[i]));
    ups_b[i] = upload(tile_bias(to_host(get(M, p + ".bias")), ups[i]));
    const int C = ch[i + 1];
    for (int j = 0; j < 3; ++j)
      for (int d = 0; d < 3; ++d)
        for (int which = 0; which < 2; ++which) {
          const std::string q = "head.resblocks." + std::to_string(i) + ".blocks." + std::to_string(j) +
                                (which ? ".convs2." : ".convs1.") + std::to_string(d) + ".conv";
          expect_shape(get(M, q + ".weight"), {C, C, rk[j]}, q);
          ResConvW rw;
          rw.k = rk[j];
          rw.d = rd[d];
          auto hw = to_host(get(M, q + ".weight"));
          rw.w = upload(rd[d] == 1 ? pack_conv_rows(hw, C, C, rk[j]) : pack_conv_taps(hw, C, C, rk[j]));
          rw.b = g(q + ".bias");
          (which ? res2 : res1)[i][j][d] = rw;
        }
  }
  expect_shape(get(M, "head.conv_post.conv.weight"), {1, 16, 13}, "conv_post");
  post_w = upload(pack_conv_rows(to_host(get(M, "head.conv_post.conv.weight")), 1, 16, 13));
  post_b = g("head.conv_post.conv.bias");
  // optional: the vocoder's own encoder (prompt path: reference wave -> codec ids)
  if (has(M, "backbone.norm.weight") && has(M, "quantizer.residual_fsq.rvqs.0.project_in.weight")) {
    build_spectrogram_consts(has(M, "spec_transform.fb") ? M : MODEL_TOKENIZER);
    pack_conv_stack(M, voc_cs);
    std::vector<float> iw, ib;
    for (int gi = 0; gi < 8; ++gi) {
      const std::string p = "quantizer.residual_fsq.rvqs." + std::to_string(gi) + ".project_in";
      expect_shape(get(M, p + ".weight"), {4, 64}, p);
      auto a = to_host(get(M, p + ".weight"));
      auto b = to_host(get(M, p + ".bias"));
      iw.insert(iw.end(), a.begin(), a.end());
      ib.insert(ib.end(), b.begin(), b.end());
    }
    fsq_in_w = upload(iw);
    fsq_in_b = upload(ib);
  }
}

// ------------------------------------------------------------------------------------------ ConvNeXt block
// ConvNeXtBlock.forward (firefly.py:421-440), channels-last, in place on x (every stream's x has >= 6 zero/history
// rows before its row 0).  tmp [rows][C], hid [rows][4C] are plain.  `rows` counts the rows of all streams; with
// seg_rows > 0 stream b's x (and out) rows start b * x_seg (out_seg) floats after the base.
void Engine::convnext(const ConvNextW& cw, float* x, int rows, float* tmp, float* hid, cudaStream_t st, float* out,
                      int seg_rows, long long x_seg, long long out_seg) {
  const int C = cw.C;
  GemmParams p1;
  p1.A = tmp; p1.W = cw.pw1_w; p1.C = hid; p1.bias = cw.pw1_b; p1.M = rows; p1.N = 4 * C; p1.K = C;
  p1.lda = C; p1.ldc = 4 * C; p1.act = ACT_GELU;
  GemmParams p2;
  p2.A = hid; p2.W = cw.pw2_w; p2.C = out ? out : x; p2.bias = cw.pw2_b; p2.gamma = cw.gamma; p2.residual = x;
  p2.M = rows; p2.N = C; p2.K = 4 * C; p2.lda = 4 * C; p2.ldc = C; p2.ldr = C;
  if (seg_rows > 0) {
    p2.seg_rows = seg_rows; p2.a_seg = (long long)seg_rows * 4 * C; p2.c_seg = out ? out_seg : x_seg; p2.r_seg = x_seg;
  }
  // wide batches (pair GEMM kernel): the producers write the lo terms of the GEMM inputs beside their results
  if (gemm_pair_eligible(&p1, 1)) p1.Alo = lo_scratch(0, (size_t)rows * C, st);
  if (p1.Alo && gemm_pair_eligible(&p2, 1)) p1.Clo = lo_scratch(1, (size_t)rows * 4 * C, st);
  p2.Alo = p1.Clo;
  launch_dwconv7_ln(x, tmp, cw.dw_w, cw.dw_b, cw.ln_w, cw.ln_b, rows, C, 1e-6f, st, seg_rows, x_seg, const_cast<float*>(p1.Alo));
  launch_gemm(p1, st);
  launch_gemm(p2, st);
}

// ------------------------------------------------------------------------------------------ stage E
namespace {
// One causal-conv input buffer [6 margin rows | T rows][C] per stream (stride seg), history hist [B][6][C].
// mode 2: margin <- hist (the left context of this step), then hist <- newest 6 rows of [hist | rows];
// mode 1: margin stays zero, hist <- newest 6 rows of [zeros | rows].  grid (ceil(C/128), B).
__global__ void conv_hist_kernel(float* __restrict__ buf, long long seg, float* __restrict__ hist, int T, int C, int mode,
                                 float* __restrict__ ring, int ring_cap, long long abs_row0) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float* b = buf + blockIdx.y * seg;                 // points at the first margin row
  float* h = hist + (long long)blockIdx.y * 6 * C;
  float old[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) old[r] = (mode == 2) ? h[r * C + c] : 0.f;
  if (mode == 2) {
#pragma unroll
    for (int r = 0; r < 6; ++r) b[(long long)r * C + c] = old[r];
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int src = T + r;                           // row index in [hist(6) | rows(T)]
    h[r * C + c] = src < 6 ? old[src] : b[(long long)src * C + c];
  }
  if (ring) {                                        // ConvStackRings: the T rows at absolute rows abs_row0 ..
    float* rg = ring + (long long)blockIdx.y * ring_cap * C;
    for (int t = 0; t < T; ++t) rg[((abs_row0 + t) & (ring_cap - 1)) * C + c] = b[(long long)(6 + t) * C + c];
  }
}

void launch_conv_hist(float* buf_margin, long long seg, float* hist, int B, int T, int C, int mode, cudaStream_t st,
                      float* ring = nullptr, int ring_cap = 0, long long abs_row0 = 0) {
  launch_pdl(conv_hist_kernel, dim3((C + 127) / 128, B), dim3(128), 0, st, buf_margin, seg, hist, T, C, mode, ring, ring_cap, abs_row0);
  SV_LAUNCHED();
}

// rows [0, T) of B side-by-side buffers (stride seg) -> ring rows abs_row0 .. (no margin in front of src)
__global__ void ring_append_kernel(const float* __restrict__ src, long long seg, int T, int C, float* __restrict__ ring, int ring_cap,
                                   long long abs_row0) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* b = src + blockIdx.y * seg;
  float* rg = ring + (long long)blockIdx.y * ring_cap * C;
  for (int t = blockIdx.z; t < T; t += gridDim.z) rg[((abs_row0 + t) & (ring_cap - 1)) * C + c] = b[(long long)t * C + c];
}
void launch_ring_append(const float* src, long long seg, int B, int T, int C, float* ring, int ring_cap, long long abs_row0, cudaStream_t st) {
  launch_pdl(ring_append_kernel, dim3((C + 127) / 128, B, T > 32 ? 32 : 1), dim3(128), 0, st, src, seg, T, C, ring, ring_cap, abs_row0);
  SV_LAUNCHED();
}
// ring rows abs_row0 .. abs_row0 + n - 1 -> rows [0, n) at dst (B buffers, stride seg)
__global__ void ring_fetch_kernel(const float* __restrict__ ring, int ring_cap, long long abs_row0, int n, float* __restrict__ dst,
                                  long long seg, int C) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float* rg = ring + (long long)blockIdx.y * ring_cap * C;
  float* d = dst + blockIdx.y * seg;
  for (int t = 0; t < n; ++t) d[(long long)t * C + c] = rg[((abs_row0 + t) & (ring_cap - 1)) * C + c];
}
void launch_ring_fetch(const float* ring, int ring_cap, long long abs_row0, int n, float* dst, long long seg, int B, int C, cudaStream_t st) {
  launch_pdl(ring_fetch_kernel, dim3((C + 127) / 128, B), dim3(128), 0, st, ring, ring_cap, abs_row0, n, dst, seg, C);
  SV_LAUNCHED();
}
// One causal layer's input of the MERGED pass (Engine::enc_conv_stack_merged), in place, per stream:
//   before: [head_old valid rows | ... | n_new rows of the newest frames at new_src]
//   after:  [head_old | 6 ring rows (steady-state rows behind the head) | 6 history rows (the newest frames' left context) | n_new]
// and, like conv_hist_kernel in mode 2, history <- newest 6 rows of [history | new rows], ring <- new rows.  A thread owns one
// channel of one stream and reads the new rows before it writes anything, so the overlapping ranges are safe.
constexpr int RELAYOUT_MAX_NEW = 16;
__global__ void relayout_kernel(float* __restrict__ x, long long seg, int C, int head_old, int new_src, int n_new,
                                float* __restrict__ ring, int ring_cap, long long abs_head, long long abs_new, float* __restrict__ hist) {
  pdl_trigger();
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float* b = x + blockIdx.y * seg;
  float* rg = ring + (long long)blockIdx.y * ring_cap * C;
  float* h = hist + (long long)blockIdx.y * 6 * C;
  float nw[RELAYOUT_MAX_NEW], old[6];
#pragma unroll
  for (int i = 0; i < RELAYOUT_MAX_NEW; ++i) nw[i] = i < n_new ? b[(long long)(new_src + i) * C + c] : 0.f;
#pragma unroll
  for (int i = 0; i < 6; ++i) old[i] = h[i * C + c];
#pragma unroll
  for (int i = 0; i < 6; ++i) b[(long long)(head_old + i) * C + c] = rg[((abs_head + i) & (ring_cap - 1)) * C + c];
#pragma unroll
  for (int i = 0; i < 6; ++i) b[(long long)(head_old + 6 + i) * C + c] = old[i];
#pragma unroll
  for (int i = 0; i < RELAYOUT_MAX_NEW; ++i)
    if (i < n_new) {
      b[(long long)(head_old + 12 + i) * C + c] = nw[i];
      rg[((abs_new + i) & (ring_cap - 1)) * C + c] = nw[i];
    }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    const int src = n_new + r;                        // row index in [history(6) | new(n_new)]
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 6; ++i) if (src == i) v = old[i];
#pragma unroll
    for (int i = 0; i < RELAYOUT_MAX_NEW; ++i) if (src == 6 + i) v = nw[i];
    h[r * C + c] = v;
  }
}
void launch_relayout(float* x, long long seg, int B, int C, int head_old, int new_src, int n_new, float* ring, int ring_cap,
                     long long abs_head, long long abs_new, float* hist, cudaStream_t st) {
  SV_CHECK(n_new >= 1 && n_new <= RELAYOUT_MAX_NEW, "merged conv pass: at most 4 content frames per chunk");
  launch_pdl(relayout_kernel, dim3((C + 127) / 128, B), dim3(128), 0, st, x, seg, C, head_old, new_src, n_new, ring, ring_cap, abs_head,
             abs_new, hist);
  SV_LAUNCHED();
}
}  // namespace

void ConvStackRings::alloc(int n, int window_frames) {
  int cap = 64;
  while (cap < window_frames + 8) cap *= 2;
  if (arena && B == n && frames_cap == cap) { frames = 0; filled_from = 0; return; }
  release();
  const int dims[4] = {128, 256, 384, 512};
  const int depths[4] = {3, 3, 9, 3};
  size_t per = (size_t)4 * cap * N_MELS;
  for (int s = 0; s < 4; ++s) per += (size_t)4 * cap * dims[s] * depths[s];
  per += (size_t)2 * cap * 512 + (size_t)cap * 512;           // blocks of the two down-sampled levels
  per += (size_t)4 * cap * 512 + (size_t)2 * cap * 512;       // inputs of the two stride-2 convs
  SV_CUDA(cudaMalloc(&arena, per * n * sizeof(float)));
  float* p = arena;
  mel = p; p += (size_t)n * 4 * cap * N_MELS;
  int j = 0;
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < depths[s]; ++d) { blk[j++] = p; p += (size_t)n * 4 * cap * dims[s]; }
  blk[j++] = p; p += (size_t)n * 2 * cap * 512;
  blk[j++] = p; p += (size_t)n * cap * 512;
  ds_in[0] = p; p += (size_t)n * 4 * cap * 512;
  ds_in[1] = p; p += (size_t)n * 2 * cap * 512;
  B = n; frames_cap = cap; frames = 0; filled_from = 0;
}

void ConvStackHist::alloc(int n) {
  if (arena && B == n) return;
  if (arena) cudaFree(arena);
  arena = nullptr;
  const int dims[4] = {128, 256, 384, 512};
  const int depths[4] = {3, 3, 9, 3};
  size_t total = (size_t)n * 6 * N_MELS;
  for (int s = 0; s < 4; ++s) total += (size_t)n * 6 * dims[s] * depths[s];
  total += (size_t)n * 6 * 512 * 2;
  SV_CUDA(cudaMalloc(&arena, total * sizeof(float)));
  SV_CUDA(cudaMemset(arena, 0, total * sizeof(float)));
  float* p = arena;
  mel = p; p += (size_t)n * 6 * N_MELS;
  int j = 0;
  for (int s = 0; s < 4; ++s)
    for (int d = 0; d < depths[s]; ++d) { blk[j++] = p; p += (size_t)n * 6 * dims[s]; }
  for (int d = 0; d < 2; ++d) { blk[j++] = p; p += (size_t)n * 6 * 512; }
  B = n;
}

static size_t enc_ws_floats(int B, long long n) { return ((size_t)(n / HOP) * 14000 + (size_t)n) * B; }

// FireflyArchitecture.encode (firefly_encoder.py:553-566) in two halves.
//
// enc_conv_stack: everything up to the transformer input -- log-mel, ConvNeXtEncoder, the two down-sampling blocks --
// for NS same-length wave segments side by side (segment i = rows of src[i / per_src] with pitch[i / per_src]): wave
// -> xt [NS][n/2048][512].  Streams never mix: every causal conv reads its own segment's zero margin; the GEMMs simply
// see NS times more rows.  Workspace comes from `ws` (caller has sized and reset it).
// mel_dst: front only -- the log-mel rows of segment i go to mel_dst + i * mel_dst_seg and nothing else runs (the chain kernel
// continues from there, enc_chain.cu).
void Engine::enc_conv_stack(const ConvStackW& w, const float* const* src, const long long* pitch, int nsrc, int per_src,
                            long long n, float* xt, cudaStream_t st, ConvStackHist* hist, int hist_mode, float* mel_dst,
                            long long mel_dst_seg) {
  SV_CHECK(w.ready && dft_w && fb_t, "encoder weights not finalized");
  SV_CHECK(hist_mode == 0 || (hist && hist->B == nsrc * per_src), "conv history does not match the stream count");
  const bool cont = hist_mode == 2;        // continuation: left context comes from the history, not from zeros
  const int B = nsrc * per_src;
  const int T = (int)(n / HOP);
  const int T2 = T / 2, S = T2 / 2;
  const int MARG = 6;
  const int BT = B * T;
  const int segT = B > 1 ? T : 0;          // seg_rows of the T-row buffers (0 = plain single stream)
  // 1. left-pad win-hop zeros (spectrogram.py:37-45) and take frames as overlapping GEMM rows (lda = hop)
  const long long wseg = N_FFT - HOP + n;
  float* wpad = ws.alloc_f(wseg * B);
  if (!cont) launch_fill(wpad, N_FFT - HOP, 0.f, st, B, wseg);
  for (int i = 0; i < nsrc; ++i) {
    const int lead = cont ? (N_FFT - HOP) : 0;          // continuation: the 1536 samples in front of the new ones are real
    SV_CUDA(cudaMemcpy2DAsync(wpad + (long long)i * per_src * wseg + (N_FFT - HOP) - lead, (size_t)wseg * sizeof(float),
                              src[i] - lead, (size_t)pitch[i] * sizeof(float), (size_t)(n + lead) * sizeof(float), per_src,
                              cudaMemcpyDeviceToDevice, st));
  }
  const int SPEC_LD = 2052;
  float* spec = ws.alloc_f((long long)BT * SPEC_LD);
  {
    GemmParams p;
    p.A = wpad; p.W = dft_w; p.C = spec; p.M = BT; p.N = 2 * N_FREQ; p.K = N_FFT; p.lda = HOP; p.ldc = SPEC_LD;
    p.seg_rows = segT; p.a_seg = wseg; p.c_seg = (long long)T * SPEC_LD;
    launch_gemm(p, st);
  }
  float* mag = ws.alloc_f((long long)BT * N_FREQ_PAD);
  launch_magnitude(spec, mag, BT, SPEC_LD, st);
  // 2. mel filterbank + log(clamp(., 1e-5))  (spectrogram.py:108-130); 6 zero rows in front for the causal stem
  const long long mel_seg = mel_dst ? mel_dst_seg : (long long)(MARG + T) * N_MELS;
  float* mel_buf = mel_dst ? mel_dst - MARG * N_MELS : ws.alloc_f(mel_seg * B);
  if (!cont && !mel_dst) launch_fill(mel_buf, (long long)MARG * N_MELS, 0.f, st, B, mel_seg);
  float* mel = mel_buf + MARG * N_MELS;
  {
    GemmParams p;
    p.A = mag; p.W = fb_t; p.C = mel; p.M = BT; p.N = N_MELS; p.K = N_FREQ_PAD; p.lda = N_FREQ_PAD; p.ldc = N_MELS;
    p.act = ACT_LOGCLAMP;
    p.seg_rows = segT; p.a_seg = (long long)T * N_FREQ_PAD; p.c_seg = mel_seg;
    launch_gemm(p, st);
  }
  if (mel_dst) return;
  // ConvStackRings (optional): this pass runs with true left context (or is the first full-window pass) -- its layer inputs are
  // the steady-state rows the window-start pass of later chunks reads back
  ConvStackRings* rg = (hist_mode && hist->rings && hist->rings->B == B) ? hist->rings : nullptr;
  const long long abs_row0 = rg ? 4 * rg->frames : 0;
  const int rcap = rg ? rg->frames_cap : 0;
  if (rg) SV_CHECK(T % 4 == 0 && T <= 4 * rcap, "conv-stack rings: whole frames, at most the ring depth");
  if (hist_mode) launch_conv_hist(mel_buf, mel_seg, hist->mel, B, T, N_MELS, hist_mode, st, rg ? rg->mel : nullptr, 4 * rcap, abs_row0);
  int blk_idx = 0;

  // 3. ConvNeXtEncoder (firefly.py:506-517)
  const int dims[4] = {128, 256, 384, 512};
  float* tmp = ws.alloc_f((long long)BT * 512);
  float* hid = ws.alloc_f((long long)BT * 2048);
  float* x = nullptr;
  long long x_seg = 0;
  for (int s = 0; s < 4; ++s) {
    const int C = dims[s];
    const long long xs = (long long)(MARG + T) * C;
    float* xb = ws.alloc_f(xs * B);
    if (!cont) launch_fill(xb, (long long)MARG * C, 0.f, st, B, xs);
    float* xn = xb + MARG * C;
    if (s == 0) {
      GemmParams p;   // stem: causal conv k=7 as one GEMM over 7 overlapping rows
      p.A = mel; p.W = w.stem_w; p.C = tmp; p.bias = w.stem_b; p.M = BT; p.N = C; p.K = 7 * N_MELS; p.lda = N_MELS;
      p.ldc = C; p.tap_off[0] = -6;
      p.seg_rows = segT; p.a_seg = mel_seg; p.c_seg = (long long)T * C;
      launch_gemm(p, st);
      launch_layernorm(tmp, xn, w.stem_ln_w, w.stem_ln_b, BT, C, 1e-6f, st, segT, (long long)T * C, xs);
    } else {
      const int Cp = dims[s - 1];
      launch_layernorm(x, tmp, w.mid_ln_w[s - 1], w.mid_ln_b[s - 1], BT, Cp, 1e-6f, st, segT, x_seg, (long long)T * Cp);
      GemmParams p;
      p.A = tmp; p.W = w.mid_w[s - 1]; p.C = xn; p.bias = w.mid_b[s - 1]; p.M = BT; p.N = C; p.K = Cp; p.lda = Cp; p.ldc = C;
      p.seg_rows = segT; p.a_seg = (long long)T * Cp; p.c_seg = xs;
      launch_gemm(p, st);
    }
    x = xn;
    x_seg = xs;
    for (auto& blk : w.blocks[s]) {
      if (hist_mode) launch_conv_hist(xb, xs, hist->blk[blk_idx], B, T, C, hist_mode, st, rg ? rg->blk[blk_idx] : nullptr, 4 * rcap, abs_row0);
      ++blk_idx;
      convnext(blk, x, BT, tmp, hid, st, nullptr, segT, x_seg, 0);
    }
  }
  float* feat = ws.alloc_f((long long)BT * 512);
  launch_layernorm(x, feat, w.bb_norm_w, w.bb_norm_b, BT, 512, 1e-6f, st, segT, x_seg, (long long)T * 512);
  // 4. DownsampleBinarySphericalQuantize.downsample (bsq_no_upsample.py:46-60): 2 x [conv k2 s2 + ConvNeXt]
  float* cur = feat;
  long long cur_seg = (long long)T * 512;
  int rows = T;
  for (int i = 0; i < 2; ++i) {
    const int r2 = rows / 2;
    const long long ds = (long long)(MARG + r2) * 512;
    float* db = ws.alloc_f(ds * B);
    if (!cont) launch_fill(db, (long long)MARG * 512, 0.f, st, B, ds);
    float* dn = db + MARG * 512;
    if (rg) launch_ring_append(cur, cur_seg, B, rows, 512, rg->ds_in[i], (i == 0 ? 4 : 2) * rcap, abs_row0 >> i, st);
    GemmParams p;
    p.A = cur; p.W = w.down_w[i]; p.C = dn; p.bias = w.down_b[i]; p.M = B * r2; p.N = 512; p.K = 1024; p.lda = 512;
    p.a_row_step = 2; p.ldc = 512;
    p.seg_rows = B > 1 ? r2 : 0; p.a_seg = cur_seg; p.c_seg = ds;
    launch_gemm(p, st);
    if (hist_mode) launch_conv_hist(db, ds, hist->blk[ConvStackHist::N_BLK - 2 + i], B, r2, 512, hist_mode, st,
                                    rg ? rg->blk[ConvStackHist::N_BLK - 2 + i] : nullptr, (i == 0 ? 2 : 1) * rcap, abs_row0 >> (i + 1));
    // the second block writes its result straight into the plain output buffer
    convnext(w.down_block[i], dn, B * r2, tmp, hid, st, i == 1 ? xt : nullptr, B > 1 ? r2 : 0, ds, (long long)r2 * 512);
    cur = dn;
    cur_seg = ds;
    rows = r2;
  }
  (void)S;
}

// The window-start span with only the rows the zero padding can reach.  The reference re-encodes the whole window with zero
// left context (infer_arvc.py:495-508); a row of a causal layer that lies further from the window start than the padding's
// reach is the same function of the same samples as in the pass that ran with true left context, so it is READ BACK from
// ConvStackRings instead of recomputed.  Reach, in rows from the window start: log-mel 3 (STFT left pad), stem 9, ConvNeXt
// block j (1..18) 9 + 6 j, first stride-2 conv ceil(117 / 2) = 59, its block 65, second stride-2 conv 33, its block 39 =
// ENC_RF - 1 transformer inputs.  Block j therefore runs on 9 + 6 j rows per stream instead of the span's 164: its input is
// the fresh output of block j - 1 (9 + 6 (j - 1) rows) followed by 6 ring rows.  45 % of the span's GEMM work remains.
void Engine::enc_conv_stack_head(const ConvStackW& w, const float* wave, long long pitch, int B, const ConvStackRings& rg,
                                 long long abs_frame0, float* xt_out, long long out_seg, cudaStream_t st) {
  SV_CHECK(w.ready && rg.arena && rg.B == B && B >= 1, "conv-stack rings not ready");
  const int MARG = 6;
  const int cap = rg.frames_cap;
  const long long r1 = 4 * abs_frame0, r2a = 2 * abs_frame0, r4 = abs_frame0;      // absolute row of the window start per rate
  // 1. log-mel: rows 0..2 fresh (zero left pad), rows 3..8 from the ring
  const int MEL_ROWS = 9;
  const long long mel_seg = (long long)(MARG + MEL_ROWS) * N_MELS;
  float* mel_buf = ws.alloc_f(mel_seg * B);
  launch_fill(mel_buf, (long long)MARG * N_MELS, 0.f, st, B, mel_seg);
  float* mel = mel_buf + MARG * N_MELS;
  {
    const long long n3 = 3 * HOP;
    enc_conv_stack(w, &wave, &pitch, 1, B, n3, nullptr, st, nullptr, 0, mel, mel_seg);
  }
  launch_ring_fetch(rg.mel, 4 * cap, r1 + 3, 6, mel + 3 * N_MELS, mel_seg, B, N_MELS, st);
  const int dims[4] = {128, 256, 384, 512};
  const int ROWS_MAX = 9 + 6 * 18;               // 117
  float* tmp = ws.alloc_f((long long)B * (ROWS_MAX + 1) * 512);
  float* hid = ws.alloc_f((long long)B * (ROWS_MAX + 1) * 2048);
  float* x = nullptr;
  long long x_seg = 0;
  int rows = MEL_ROWS;                           // fresh rows of the current layer input
  int j = 0;                                     // blocks done
  for (int s = 0; s < 4; ++s) {
    const int C = dims[s];
    const int rows_end = 9 + 6 * (j + (int)w.blocks[s].size());        // rows after the stage's last block
    const long long xs = (long long)(MARG + rows_end) * C;
    float* xb = ws.alloc_f(xs * B);
    launch_fill(xb, (long long)MARG * C, 0.f, st, B, xs);
    float* xn = xb + MARG * C;
    if (s == 0) {
      GemmParams p;   // stem on 9 rows
      p.A = mel; p.W = w.stem_w; p.C = tmp; p.bias = w.stem_b; p.M = B * rows; p.N = C; p.K = 7 * N_MELS; p.lda = N_MELS;
      p.ldc = C; p.tap_off[0] = -6;
      p.seg_rows = rows; p.a_seg = mel_seg; p.c_seg = (long long)rows * C;
      launch_gemm(p, st);
      launch_layernorm(tmp, xn, w.stem_ln_w, w.stem_ln_b, B * rows, C, 1e-6f, st, rows, (long long)rows * C, xs);
    } else {
      const int Cp = dims[s - 1];
      launch_layernorm(x, tmp, w.mid_ln_w[s - 1], w.mid_ln_b[s - 1], B * rows, Cp, 1e-6f, st, rows, x_seg, (long long)rows * Cp);
      GemmParams p;
      p.A = tmp; p.W = w.mid_w[s - 1]; p.C = xn; p.bias = w.mid_b[s - 1]; p.M = B * rows; p.N = C; p.K = Cp; p.lda = Cp; p.ldc = C;
      p.seg_rows = rows; p.a_seg = (long long)rows * Cp; p.c_seg = xs;
      launch_gemm(p, st);
    }
    x = xn;
    x_seg = xs;
    for (auto& blk : w.blocks[s]) {
      // input of block j + 1: fresh rows [0, rows) + 6 steady-state rows
      launch_ring_fetch(rg.blk[j], 4 * cap, r1 + rows, 6, x + (long long)rows * C, xs, B, C, st);
      rows += 6;
      convnext(blk, x, B * rows, tmp, hid, st, nullptr, rows, x_seg, 0);
      ++j;
    }
  }
  // rows == 117 fresh rows of the backbone output; the stride-2 conv pairs rows (2 i, 2 i + 1): row 117 from the ring
  const int R0 = rows + 1;                       // 118
  float* feat = ws.alloc_f((long long)B * R0 * 512);
  launch_layernorm(x, feat, w.bb_norm_w, w.bb_norm_b, B * rows, 512, 1e-6f, st, rows, x_seg, (long long)R0 * 512);
  launch_ring_fetch(rg.ds_in[0], 4 * cap, r1 + rows, 1, feat + (long long)rows * 512, (long long)R0 * 512, B, 512, st);
  float* cur = feat;
  long long cur_seg = (long long)R0 * 512;
  int in_rows = R0;
  for (int i = 0; i < 2; ++i) {
    const int o = in_rows / 2;                   // 59, then 33
    const int o_end = o + 6;                     // rows after the level's block: 65, 39
    const bool last = i == 1;
    // the block's output feeds the next stride-2 conv, which needs one more (ring) row behind it
    const long long ds = (long long)(MARG + o_end + 1) * 512;
    float* db = ws.alloc_f(ds * B);
    launch_fill(db, (long long)MARG * 512, 0.f, st, B, ds);
    float* dn = db + MARG * 512;
    GemmParams p;
    p.A = cur; p.W = w.down_w[i]; p.C = dn; p.bias = w.down_b[i]; p.M = B * o; p.N = 512; p.K = 1024; p.lda = 512;
    p.a_row_step = 2; p.ldc = 512;
    p.seg_rows = o; p.a_seg = cur_seg; p.c_seg = ds;
    launch_gemm(p, st);
    launch_ring_fetch(rg.blk[ConvStackHist::N_BLK - 2 + i], (i == 0 ? 2 : 1) * cap, (i == 0 ? r2a : r4) + o, 6, dn + (long long)o * 512, ds,
                      B, 512, st);
    convnext(w.down_block[i], dn, B * o_end, tmp, hid, st, last ? xt_out : nullptr, o_end, ds, out_seg);
    if (!last) {
      launch_ring_fetch(rg.ds_in[1], 2 * cap, r2a + o_end, 1, dn + (long long)o_end * 512, ds, B, 512, st);
      cur = dn;
      cur_seg = ds;
      in_rows = o_end + 1;                       // 66
    } else {
      SV_CHECK(o_end == ENC_RF - 1, "window-start reach");
    }
  }
}

// The window-start pass (enc_conv_stack_head) and the per-layer-history pass of the c newest frames (enc_conv_stack, hist_mode 2) in
// the SAME launches.  Separately the second is ~80 latency-bound launches on 4 c rows per stream; here every layer's input is
// [6 zero rows | head rows | 6 ring rows | 6 history rows | 4 c new rows] per stream: the causal convs of the new rows see their true
// left context (the history rows), the rows computed at the history positions are junk nobody reads, and one small kernel per
// layer (relayout_kernel) moves the new rows behind the grown head, fetches ring and history rows and appends the new rows to
// both.  The stride-2 convs read row pairs, so their inputs are re-packed as [head rows | new rows] with even counts.
// xt_head [B][ENC_RF - 1][512] with stream pitch head_seg, xt_tail [B][c][512] plain.
void Engine::enc_conv_stack_merged(const ConvStackW& w, const float* wave_ring, long long pitch, long long nw, int B, int c,
                                   ConvStackHist& hist, ConvStackRings& rg, long long abs_frame0, float* xt_head, long long head_seg,
                                   float* xt_tail, cudaStream_t st) {
  SV_CHECK(w.ready && rg.arena && rg.B == B && hist.B == B && c >= 1 && 4 * c <= RELAYOUT_MAX_NEW, "merged conv pass: state not ready");
  const int MARG = 6;
  const int cap = rg.frames_cap;
  const long long r1 = 4 * abs_frame0, r2a = 2 * abs_frame0, r4 = abs_frame0;      // absolute row of the window start per rate
  const long long a1 = 4 * rg.frames, a2 = 2 * rg.frames, a4 = rg.frames;          // absolute row of the first new row per rate
  const int n1 = 4 * c, n2 = 2 * c, n4 = c;
  // 1. log-mel: [6 zero | 3 fresh head rows | 6 ring | 6 history | n1 new]
  const int MEL_ROWS = 9 + 6 + n1;
  const long long mel_seg = (long long)(MARG + MEL_ROWS) * N_MELS;
  float* mel_buf = ws.alloc_f(mel_seg * B);
  launch_fill(mel_buf, (long long)MARG * N_MELS, 0.f, st, B, mel_seg);
  float* mel = mel_buf + MARG * N_MELS;
  {
    const long long n3 = 3 * HOP;
    enc_conv_stack(w, &wave_ring, &pitch, 1, B, n3, nullptr, st, nullptr, 0, mel, mel_seg);
    // newest frames with their true left context (the 1536 samples in front), straight to their final rows
    const long long nt = (long long)c * SAMPLES_PER_FRAME;
    const float* tsrc = wave_ring + (nw - nt);
    ConvStackRings* keep = hist.rings;
    hist.rings = nullptr;                                    // the front only: history and rings are updated by relayout below
    enc_conv_stack(w, &tsrc, &pitch, 1, B, nt, nullptr, st, &hist, 2, mel + 15 * N_MELS, mel_seg);
    hist.rings = keep;
  }
  launch_relayout(mel, mel_seg, B, N_MELS, 3, 15, n1, rg.mel, 4 * cap, r1 + 3, a1, hist.mel, st);
  const int dims[4] = {128, 256, 384, 512};
  const int TAILR = 6 + n1;                      // history + new rows behind the head rows
  const int ROWS_MAX = 117 + TAILR;
  float* tmp = ws.alloc_f((long long)B * (ROWS_MAX + 1) * 512);
  float* hid = ws.alloc_f((long long)B * (ROWS_MAX + 1) * 2048);
  float* x = nullptr;
  long long x_seg = 0;
  int rows = 9;                                  // head rows of the current layer input
  int j = 0;                                     // blocks done
  for (int s = 0; s < 4; ++s) {
    const int C = dims[s];
    const int rows_end = 9 + 6 * (j + (int)w.blocks[s].size());
    const long long xs = (long long)(MARG + rows_end + TAILR) * C;
    float* xb = ws.alloc_f(xs * B);
    launch_fill(xb, (long long)MARG * C, 0.f, st, B, xs);
    float* xn = xb + MARG * C;
    const int seg_rows = rows + TAILR;           // every row of the layout goes through the point-wise layers
    if (s == 0) {
      GemmParams p;
      p.A = mel; p.W = w.stem_w; p.C = tmp; p.bias = w.stem_b; p.M = B * seg_rows; p.N = C; p.K = 7 * N_MELS; p.lda = N_MELS;
      p.ldc = C; p.tap_off[0] = -6;
      p.seg_rows = seg_rows; p.a_seg = mel_seg; p.c_seg = (long long)seg_rows * C;
      launch_gemm(p, st);
      launch_layernorm(tmp, xn, w.stem_ln_w, w.stem_ln_b, B * seg_rows, C, 1e-6f, st, seg_rows, (long long)seg_rows * C, xs);
    } else {
      const int Cp = dims[s - 1];
      launch_layernorm(x, tmp, w.mid_ln_w[s - 1], w.mid_ln_b[s - 1], B * seg_rows, Cp, 1e-6f, st, seg_rows, x_seg, (long long)seg_rows * Cp);
      GemmParams p;
      p.A = tmp; p.W = w.mid_w[s - 1]; p.C = xn; p.bias = w.mid_b[s - 1]; p.M = B * seg_rows; p.N = C; p.K = Cp; p.lda = Cp; p.ldc = C;
      p.seg_rows = seg_rows; p.a_seg = (long long)seg_rows * Cp; p.c_seg = xs;
      launch_gemm(p, st);
    }
    x = xn;
    x_seg = xs;
    for (auto& blk : w.blocks[s]) {
      // [rows | 6 junk | n1 new at rows + 6] -> [rows + 6 | 6 history | n1 new]
      launch_relayout(x, xs, B, C, rows, rows + 6, n1, rg.blk[j], 4 * cap, r1 + rows, a1, hist.blk[j], st);
      rows += 6;
      convnext(blk, x, B * (rows + TAILR), tmp, hid, st, nullptr, rows + TAILR, x_seg, 0);
      ++j;
    }
  }
  // backbone output: head rows [0, 117), new rows at 117 + 6.  Input of the first stride-2 conv: [117 head | ring row | n1 new]
  const int R0 = rows + 1;                       // 118
  const long long f_seg = (long long)(R0 + n1) * 512;
  float* feat = ws.alloc_f(f_seg * B);
  launch_layernorm(x, feat, w.bb_norm_w, w.bb_norm_b, B * rows, 512, 1e-6f, st, rows, x_seg, f_seg);
  launch_layernorm(x + (long long)(rows + 6) * 512, feat + (long long)R0 * 512, w.bb_norm_w, w.bb_norm_b, B * n1, 512, 1e-6f, st, n1, x_seg,
                   f_seg);
  launch_ring_fetch(rg.ds_in[0], 4 * cap, r1 + rows, 1, feat + (long long)rows * 512, f_seg, B, 512, st);
  launch_ring_append(feat + (long long)R0 * 512, f_seg, B, n1, 512, rg.ds_in[0], 4 * cap, a1, st);
  const float* cur = feat;
  long long cur_seg = f_seg;
  int in_head = R0;                              // head rows of the stride-2 conv's input (even)
  int n_in = n1;
  for (int i = 0; i < 2; ++i) {
    const int o = in_head / 2, n_out = n_in / 2; // head rows 59 / 33, new rows 2 c / c
    const int o_end = o + 6;                     // 65, 39
    const bool last = i == 1;
    const long long ds = (long long)(MARG + o_end + 6 + n_out) * 512;
    float* db = ws.alloc_f(ds * B);
    launch_fill(db, (long long)MARG * 512, 0.f, st, B, ds);
    float* dn = db + MARG * 512;
    GemmParams p;
    p.A = cur; p.W = w.down_w[i]; p.C = dn; p.bias = w.down_b[i]; p.M = B * (o + n_out); p.N = 512; p.K = 1024; p.lda = 512;
    p.a_row_step = 2; p.ldc = 512;
    p.seg_rows = o + n_out; p.a_seg = cur_seg; p.c_seg = ds;
    launch_gemm(p, st);
    // [o head | n_out new at o] -> [o + 6 | 6 history | n_out new]
    const int bi = ConvStackHist::N_BLK - 2 + i;
    launch_relayout(dn, ds, B, 512, o, o, n_out, rg.blk[bi], (i == 0 ? 2 : 1) * cap, (i == 0 ? r2a : r4) + o, i == 0 ? a2 : a4, hist.blk[bi], st);
    const int seg_rows = o_end + 6 + n_out;
    if (!last) {
      convnext(w.down_block[i], dn, B * seg_rows, tmp, hid, st, nullptr, seg_rows, ds, 0);
      // input of the second stride-2 conv: [65 head | ring row | n_out new]
      const long long g_seg = (long long)(o_end + 1 + n_out) * 512;
      float* g = ws.alloc_f(g_seg * B);
      SV_CUDA(cudaMemcpy2DAsync(g, (size_t)g_seg * sizeof(float), dn, (size_t)ds * sizeof(float), (size_t)o_end * 512 * sizeof(float), B,
                                cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpy2DAsync(g + (long long)(o_end + 1) * 512, (size_t)g_seg * sizeof(float), dn + (long long)(o_end + 6) * 512,
                                (size_t)ds * sizeof(float), (size_t)n_out * 512 * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
      launch_ring_fetch(rg.ds_in[1], 2 * cap, r2a + o_end, 1, g + (long long)o_end * 512, g_seg, B, 512, st);
      launch_ring_append(g + (long long)(o_end + 1) * 512, g_seg, B, n_out, 512, rg.ds_in[1], 2 * cap, a2, st);
      cur = g;
      cur_seg = g_seg;
      in_head = o_end + 1;                       // 66
      n_in = n_out;
    } else {
      SV_CHECK(o_end == ENC_RF - 1 && n_out == c, "window-start reach");
      const long long o_seg = (long long)seg_rows * 512;
      float* outb = ws.alloc_f(o_seg * B);
      convnext(w.down_block[i], dn, B * seg_rows, tmp, hid, st, outb, seg_rows, ds, o_seg);
      SV_CUDA(cudaMemcpy2DAsync(xt_head, (size_t)head_seg * sizeof(float), outb, (size_t)o_seg * sizeof(float),
                                (size_t)o_end * 512 * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpy2DAsync(xt_tail, (size_t)c * 512 * sizeof(float), outb + (long long)(o_end + 6) * 512, (size_t)o_seg * sizeof(float),
                                (size_t)c * 512 * sizeof(float), B, cudaMemcpyDeviceToDevice, st));
    }
  }
}

// enc_transformer_bsq: WindowLimitedTransformer (windowed_transformer.py:337-354) over S tokens per stream, positions
// 0..S-1 in every stream, in place on xt [B][S][512]; then the 13-bit BSQ ids (bsq.py:330-369).
// keep_last = c > 0: the caller reads the ids of the last c tokens of every stream only (the streaming loop,
// infer_arvc.py:506-518), so the LAST layer runs its queries, output projection and MLP for those rows alone -- keys and
// values of that layer still come from all S tokens.  ids_dev keeps its [B][S] layout; only columns S-c.. are written.
void Engine::enc_transformer_bsq(float* xt, int B, int S, long long* ids_dev, cudaStream_t st, int keep_last, float* hidden_out) {
  const int BS = B * S;
  const bool tail_only = enc_tail_only && keep_last > 0 && keep_last < S && S <= ATT_TAIL_MAX_KEYS;
  float* nrm = ws.alloc_f((long long)BS * ENC_DIM);
  float* qkv = ws.alloc_f((long long)BS * 3 * ENC_DIM);
  float* y = ws.alloc_f((long long)BS * ENC_DIM);
  float* h13 = ws.alloc_f((long long)BS * 2 * ENC_INTER);
  float* gbuf = ws.alloc_f((long long)BS * ENC_INTER);
  // wide batches (pair GEMM kernel, gemm_pair.cu): norm / attention / SiLU-mul write the lo terms of their results as well
  float *nrm_lo = nullptr, *y_lo = nullptr, *g_lo = nullptr;
  {
    GemmParams probe;
    probe.A = nrm; probe.W = enc_layers[0].wqkv; probe.C = qkv; probe.M = BS; probe.N = 3 * ENC_DIM; probe.K = ENC_DIM;
    probe.lda = ENC_DIM; probe.ldc = 3 * ENC_DIM;
    if (gemm_pair_eligible(&probe, 1)) {
      nrm_lo = lo_scratch(0, (size_t)BS * ENC_DIM, st);
      g_lo = lo_scratch(1, (size_t)BS * ENC_INTER, st);
      y_lo = lo_scratch(2, (size_t)BS * ENC_DIM, st);
      if (!nrm_lo || !g_lo || !y_lo) nrm_lo = y_lo = g_lo = nullptr;
    }
  }
  for (int l = 0; l < ENC_LAYERS; ++l) {
    const EncLayerW& L = enc_layers[l];
    launch_rmsnorm(xt, nrm, L.attn_norm, BS, ENC_DIM, 1e-5f, st, 0, nrm_lo);
    GemmParams p;
    p.A = nrm; p.W = L.wqkv; p.C = qkv; p.M = BS; p.N = 3 * ENC_DIM; p.K = ENC_DIM; p.lda = ENC_DIM; p.ldc = 3 * ENC_DIM;
    p.Alo = nrm_lo;
    // RoPE on q | k rides in the GEMM (pair kernel: in its epilogue; other back ends: launch_rope_qk afterwards)
    p.rope_table = enc_rope; p.rope_cols = 2 * ENC_DIM; p.rope_seg_rows = B > 1 ? S : 0; p.rope_pos0 = 0;
    launch_gemm(p, st);
    if (tail_only && l == ENC_LAYERS - 1) {
      const int c = keep_last, R = B * c;
      float* xtail = ws.alloc_f((long long)R * ENC_DIM);
      long long* ids_tail = reinterpret_cast<long long*>(ws.alloc_f((long long)R * 2 + 4));
      launch_attention_tail(qkv, 3 * ENC_DIM, qkv + ENC_DIM, qkv + 2 * ENC_DIM, HEAD_DIM, 3 * ENC_DIM, y, ENC_DIM, S, c,
                            ENC_HEADS, ENC_WINDOW, st, B);
      SV_CUDA(cudaMemcpy2DAsync(xtail, (size_t)c * ENC_DIM * sizeof(float), xt + (long long)(S - c) * ENC_DIM,
                                (size_t)S * ENC_DIM * sizeof(float), (size_t)c * ENC_DIM * sizeof(float), B,
                                cudaMemcpyDeviceToDevice, st));
      GemmParams po;
      po.A = y; po.W = L.wo; po.C = xtail; po.gamma = L.ls_attn; po.residual = xtail; po.M = R; po.N = ENC_DIM; po.K = ENC_DIM;
      po.lda = ENC_DIM; po.ldc = ENC_DIM; po.ldr = ENC_DIM;
      launch_gemm(po, st);
      launch_rmsnorm(xtail, nrm, L.ffn_norm, R, ENC_DIM, 1e-5f, st);
      GemmParams p1;
      p1.A = nrm; p1.W = L.w1; p1.C = h13; p1.M = R; p1.N = ENC_INTER; p1.K = ENC_DIM; p1.lda = ENC_DIM; p1.ldc = 2 * ENC_INTER;
      GemmParams p13[2] = {p1, p1};
      p13[1].W = L.w3; p13[1].C = h13 + ENC_INTER;
      launch_gemm(p13, 2, st);
      launch_silu_mul(h13, gbuf, R, ENC_INTER, st);
      GemmParams p2;
      p2.A = gbuf; p2.W = L.w2; p2.C = xtail; p2.gamma = L.ls_ffn; p2.residual = xtail; p2.M = R; p2.N = ENC_DIM; p2.K = ENC_INTER;
      p2.lda = ENC_INTER; p2.ldc = ENC_DIM; p2.ldr = ENC_DIM;
      launch_gemm(p2, st);
      launch_rmsnorm(xtail, nrm, enc_norm_w, R, ENC_DIM, 1e-5f, st);
      if (hidden_out) SV_CUDA(cudaMemcpyAsync(hidden_out, nrm, (size_t)R * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
      launch_bsq(nrm, bsq_w, bsq_b, ids_tail, R, st);
      SV_CUDA(cudaMemcpy2DAsync(ids_dev + (S - c), (size_t)S * sizeof(long long), ids_tail, (size_t)c * sizeof(long long),
                                (size_t)c * sizeof(long long), B, cudaMemcpyDeviceToDevice, st));
      return;
    }
    // (the lo output exists on the short-sequence attention kernel only: windows of <= 128 tokens, the streaming loop)
    float* y_lo_l = (S <= 128) ? y_lo : nullptr;
    launch_attention(qkv, 3 * ENC_DIM, qkv + ENC_DIM, qkv + 2 * ENC_DIM, HEAD_DIM, 3 * ENC_DIM, y, ENC_DIM, S, 0,
                     ENC_HEADS, ENC_WINDOW, st, B, y_lo_l);
    GemmParams po;
    po.A = y; po.W = L.wo; po.C = xt; po.gamma = L.ls_attn; po.residual = xt; po.M = BS; po.N = ENC_DIM; po.K = ENC_DIM;
    po.lda = ENC_DIM; po.ldc = ENC_DIM; po.ldr = ENC_DIM;
    po.Alo = y_lo_l;
    launch_gemm(po, st);
    launch_rmsnorm(xt, nrm, L.ffn_norm, BS, ENC_DIM, 1e-5f, st, 0, nrm_lo);
    GemmParams p1;                           // SwiGLU gate: silu(x w1^T) * (x w3^T), both products in one launch
    p1.A = nrm; p1.W = L.w1; p1.W2 = L.w3; p1.C = gbuf; p1.M = BS; p1.N = ENC_INTER; p1.K = ENC_DIM; p1.lda = ENC_DIM;
    p1.ldc = ENC_INTER; p1.dual_tmp = h13;
    p1.Alo = nrm_lo; p1.Clo = g_lo;
    launch_gemm(p1, st);
    GemmParams p2;
    p2.A = gbuf; p2.W = L.w2; p2.C = xt; p2.gamma = L.ls_ffn; p2.residual = xt; p2.M = BS; p2.N = ENC_DIM; p2.K = ENC_INTER;
    p2.lda = ENC_INTER; p2.ldc = ENC_DIM; p2.ldr = ENC_DIM;
    p2.Alo = g_lo;
    launch_gemm(p2, st);
  }
  launch_rmsnorm(xt, nrm, enc_norm_w, BS, ENC_DIM, 1e-5f, st);
  if (hidden_out) SV_CUDA(cudaMemcpyAsync(hidden_out, nrm, (size_t)BS * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  launch_bsq(nrm, bsq_w, bsq_b, ids_dev, BS, st);
}


// B full-length utterances / windows of the same length, side by side: wave [B][n] -> ids [B][n/2048].
void Engine::enc_encode(const float* wave, int B, long long n, long long* ids_dev, cudaStream_t st) {
  NvtxRange nvtx_("svanon:E encode");
  SV_CHECK(finalized[MODEL_TOKENIZER], "tokenizer weights not finalized");
  SV_CHECK(B >= 1, "no utterances");
  const int S = (int)(n / HOP) / 4;
  SV_CHECK(S >= 1, "utterance shorter than one content frame (2048 samples)");
  SV_CHECK(S <= 2048, "utterance longer than the tokenizer's RoPE table (2048 content frames)");
  SV_CHECK((long long)B * (n / HOP) < (1 << 30), "batch too large");
  ws.ensure((enc_ws_floats(B, n) + (4u << 20)) * sizeof(float));
  ws.reset();
  float* xt = ws.alloc_f((long long)B * S * ENC_DIM);
  enc_conv_stack(tok_cs, &wave, &n, 1, B, n, xt, st);
  enc_transformer_bsq(xt, B, S, ids_dev, st);
}

namespace {
// xt_new[b][p] = p < rf ? head[b][p] : p < S - c ? prev[b][p + c] : tail[b][Ls - (S - p)]
__global__ void enc_assemble_kernel(const float* __restrict__ spans, const float* __restrict__ prev, float* __restrict__ out,
                                    int B, int S, int Ls, int rf, int c) {
  pdl_trigger();
  pdl_wait();
  const int p = blockIdx.x, b = blockIdx.y;
  const float* src;
  if (p < rf) src = spans + ((long long)b * Ls + p) * ENC_DIM;
  else if (p < S - c) src = prev + ((long long)b * S + p + c) * ENC_DIM;
  else src = spans + ((long long)(B + b) * Ls + (Ls - (S - p))) * ENC_DIM;
  float* dst = out + ((long long)b * S + p) * ENC_DIM;
  for (int i = threadIdx.x; i < ENC_DIM / 4; i += blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}
}  // namespace

// The streaming loop's window re-encode (infer_arvc.py:495-508) with persistent ring-buffer state, same result.
// Every conv of the tokenizer is left-pad-only causal and the receptive field of one transformer-input token is 39
// content frames (stem 6 + 18 x 6 + STFT 3 + down-sampling 39 mel steps), so when the window
// slides by c frames the transformer inputs of window positions >= 39 are the SAME function of the SAME samples as
// one chunk earlier at position + c.  Only two spans go through the conv stack again: the window's first ENC_RF + c
// frames (their tokens see the zero padding at the window start, exactly like in the reference's recompute) and its
// last ENC_RF + c frames (whose last c tokens are the new ones); the tokens in between come from the per-stream
// state.  The attention transformer then runs over all S tokens as in the reference (its inputs at every position
// change with the window start, nothing of it can be kept).  state->xt [B][S][512] double-buffered.
void Engine::enc_window_step(EncWindowState& state, const float* wave_ring, int B, int S, int c, long long* ids_dev,
                             cudaStream_t st) {
  NvtxRange nvtx_("svanon:E window step");
  SV_CHECK(finalized[MODEL_TOKENIZER], "tokenizer weights not finalized");
  const long long nw = (long long)S * SAMPLES_PER_FRAME;
  const int Ls = ENC_RF + c;
  if (state.B != B || state.S != S || !state.xt[0]) {
    for (auto& p : state.xt) {
      if (p) cudaFree(p);
      p = nullptr;
      SV_CUDA(cudaMalloc(&p, (size_t)B * S * ENC_DIM * sizeof(float)));
    }
    state.B = B; state.S = S; state.cur = 0; state.valid = false;
  }
  const bool incremental = state.valid && S >= 2 * Ls + 8;
  // Many streams: the newest frames continue from per-layer conv history (4 mel rows per frame instead of a 41-frame
  // span); few streams: the tail span rides in the same launches as the head span, which is faster than ~120 more
  // (tiny) launches.
  const bool use_hist = state.tail_hist_min_streams > 0 && B >= state.tail_hist_min_streams && S >= 2 * Ls + 8;
  if (!state.valid) state.hist_valid = false;
  float* xt_state = state.xt[state.cur ^ 1];
  // Steady-state layer inputs of the last window (ConvStackRings): with them the window-start pass recomputes only the rows
  // the zero padding reaches.  (Re)allocated empty when the stream count changes (cohort merge); usable once they cover the
  // window again.
  const bool want_rings = use_hist && state.use_rings;
  state.hist.rings = want_rings ? &state.rings : nullptr;
  if (!want_rings && state.rings.arena) state.rings.release();
  if (!incremental || (use_hist && !state.hist_valid)) {
    ws.ensure((enc_ws_floats(B, nw) + (size_t)B * S * ENC_DIM + (4u << 20)) * sizeof(float));
    ws.reset();
    if (use_hist) state.hist.alloc(B);
    if (want_rings) state.rings.alloc(B, S);                  // empty; the pass below appends the window's 4 S rows per layer
    enc_conv_stack(tok_cs, &wave_ring, &nw, 1, B, nw, xt_state, st, use_hist ? &state.hist : nullptr, use_hist ? 1 : 0);
    if (want_rings) state.rings.frames += S;
    state.hist_valid = use_hist;
  } else if (use_hist) {
    const long long nh = (long long)Ls * SAMPLES_PER_FRAME, nt = (long long)c * SAMPLES_PER_FRAME;
    ws.ensure((enc_ws_floats(B, nh) + enc_ws_floats(B, nt) + enc_ws_floats(B, nw) / 3 + (size_t)(2 * B * Ls + B * S) * ENC_DIM +
               (4u << 20)) * sizeof(float));
    ws.reset();
    if (want_rings && (state.rings.B != B || state.rings.frames_cap < S + 8)) state.rings.alloc(B, S);
    // absolute frame number of the window start after this chunk; the rings serve it when they hold every frame from there on
    // (>= 1: the rows of frame 0 of a first full-window pass saw that pass's own zero padding)
    const long long start = want_rings ? state.rings.frames + c - S : -1;
    const bool tri = want_rings && start >= 1;
    float* spans = ws.alloc_f((long long)2 * B * Ls * ENC_DIM);          // [head B x Ls | tail B x Ls (last c rows used)]
    float* tail = ws.alloc_f((long long)B * c * ENC_DIM);
    static const bool merged = [] {
      const char* e = getenv("SVANON_ENC_MERGED");           // 0: window-start pass and newest-frames pass as separate launches
      return !e || atoi(e) != 0;
    }();
    if (tri && merged && 4 * c <= 16) {
      enc_conv_stack_merged(tok_cs, wave_ring, nw, nw, B, c, state.hist, state.rings, start, spans, (long long)Ls * ENC_DIM, tail, st);
    } else {
      if (tri) enc_conv_stack_head(tok_cs, wave_ring, nw, B, state.rings, start, spans, (long long)Ls * ENC_DIM, st);
      else enc_conv_stack(tok_cs, &wave_ring, &nw, 1, B, nh, spans, st);  // window start: zero left context, whole span
      const float* tsrc = wave_ring + (nw - nt);
      enc_conv_stack(tok_cs, &tsrc, &nw, 1, B, nt, tail, st, &state.hist, 2);
    }
    if (want_rings) state.rings.frames += c;
    // place the c new tokens where the assemble kernel expects the end of a tail span
    SV_CUDA(cudaMemcpy2DAsync(spans + ((long long)B * Ls + (Ls - c)) * ENC_DIM, (size_t)Ls * ENC_DIM * sizeof(float), tail,
                              (size_t)c * ENC_DIM * sizeof(float), (size_t)c * ENC_DIM * sizeof(float), B,
                              cudaMemcpyDeviceToDevice, st));
    launch_pdl(enc_assemble_kernel, dim3(S, B), dim3(128), 0, st, (const float*)spans, (const float*)state.xt[state.cur],
               xt_state, B, S, Ls, tri ? ENC_RF - 1 : ENC_RF, c);
    SV_LAUNCHED();
  } else {
    const long long ns = (long long)Ls * SAMPLES_PER_FRAME;
    ws.ensure((enc_ws_floats(2 * B, ns) + enc_ws_floats(B, nw) / 3 + (size_t)(2 * B * Ls + B * S) * ENC_DIM + (4u << 20)) *
              sizeof(float));
    ws.reset();
    float* spans = ws.alloc_f((long long)2 * B * Ls * ENC_DIM);
    const float* src[2] = {wave_ring, wave_ring + (nw - ns)};
    const long long pitch[2] = {nw, nw};
    // one stream: the chain kernel runs everything behind the mel filterbank (conv stack, assemble, transformer, BSQ)
    if (B == 1 && enc_window_chain(nullptr, state.xt[state.cur], xt_state, S, c, Ls, ids_dev, st, src, pitch)) {
      state.cur ^= 1;
      state.valid = true;
      return;
    }
    enc_conv_stack(tok_cs, src, pitch, 2, B, ns, spans, st);
    if (B == 1 && enc_window_chain(spans, state.xt[state.cur], xt_state, S, c, Ls, ids_dev, st)) {
      state.cur ^= 1;                    // the chain assembled the window, ran the transformer and wrote the ids
      state.valid = true;
      return;
    }
    launch_pdl(enc_assemble_kernel, dim3(S, B), dim3(128), 0, st, (const float*)spans, (const float*)state.xt[state.cur],
               xt_state, B, S, Ls, ENC_RF, c);
    SV_LAUNCHED();
  }
  state.cur ^= 1;
  state.valid = true;
  float* xt = ws.alloc_f((long long)B * S * ENC_DIM);
  SV_CUDA(cudaMemcpyAsync(xt, xt_state, (size_t)B * S * ENC_DIM * sizeof(float), cudaMemcpyDeviceToDevice, st));
  enc_transformer_bsq(xt, B, S, ids_dev, st, c);
}

// ------------------------------------------------------------------------------------------ prompt path: wave -> codec ids
// FireflyArchitecture.encode of the vocoder (firefly.py:561-574; `wav2target_fn`, infer_arvc.py:168-171) for B
// full-length rows: the same conv stack as the tokenizer with the vocoder's weights, then the FSQ indices of the 8
// groups (DownsampleFiniteScalarQuantize.encode, fsq.py:106-110).
void Engine::voc_encode(const float* wave, int B, long long n, int* codes_dev, cudaStream_t st) {
  NvtxRange nvtx_("svanon:prompt voc_encode");
  SV_CHECK(finalized[MODEL_VOCODER], "vocoder weights not finalized");
  SV_CHECK(voc_cs.ready, "the vocoder checkpoint was loaded without its encoder (backbone.*, quantizer.downsample.*, project_in)");
  const int S = (int)(n / HOP) / 4;
  SV_CHECK(B >= 1 && S >= 1, "utterance shorter than one frame (2048 samples)");
  ws.ensure((enc_ws_floats(B, n) + (4u << 20)) * sizeof(float));
  ws.reset();
  float* z = ws.alloc_f((long long)B * S * ENC_DIM);
  enc_conv_stack(voc_cs, &wave, &n, 1, B, n, z, st);
  launch_fsq_encode(z, fsq_in_w, fsq_in_b, codes_dev, B, S, st);
}

// ------------------------------------------------------------------------------------------ stage V
static size_t voc_ws_bytes(int T) { return ((size_t)T * 1300000 + (8u << 20)) * sizeof(float); }

// DownsampleFiniteScalarQuantize.decode (fsq.py:112-116): FSQ lookup + 2 x [tconv k2 s2 + ConvNeXt]
static void voc_qdecode_impl(Engine& e, const long long* codes, long long ld, int T, float* z_out, cudaStream_t st) {
  Workspace& ws = e.ws;
  float* z0 = ws.alloc_f((long long)T * 512);
  launch_fsq_lookup(codes, ld, e.fsq_w, e.fsq_b, z0, T, st);
  const int MARG = 6;
  float* tmp = ws.alloc_f((long long)4 * T * 512);
  float* hid = ws.alloc_f((long long)4 * T * 2048);
  float* cur = z0;
  int rows = T;
  for (int i = 0; i < 2; ++i) {
    const int r2 = rows * 2;
    float* xb;
    if (i == 1) {
      xb = z_out;                       // caller provides >= 6 rows of margin in front of z_out
    } else {
      float* buf = ws.alloc_f((long long)(MARG + r2) * 512);
      xb = buf + MARG * 512;
    }
    launch_fill(xb - MARG * 512, (long long)MARG * 512, 0.f, st);
    GemmParams p;
    p.A = cur; p.W = e.up_w[i]; p.C = xb; p.bias = e.up_b[i]; p.M = rows; p.N = 1024; p.K = 512; p.lda = 512; p.ldc = 1024;
    launch_gemm(p, st);
    e.convnext(e.up_block[i], xb, r2, tmp, hid, st);

    cur = xb;
    rows = r2;
  }

}

// HiFiGANGenerator.forward (firefly.py:280-293).  z has >= 12 zero rows in front (conv_pre k = 13).
static void voc_head_impl(Engine& e, const float* z, int L, float* wave, cudaStream_t st) {
  Workspace& ws = e.ws;
  const int ch[6] = {512, 256, 128, 64, 32, 16};
  const int ups[5] = {8, 8, 2, 2, 2};
  // conv_pre -> c0 with 1 margin row (x[t-1] of the first transposed conv)
  float* c0b = ws.alloc_f((long long)(1 + L) * 512);
  launch_fill(c0b, 512, 0.f, st);
  float* cur = c0b + 512;
  {
    GemmParams p;
    p.A = z; p.W = e.pre_w; p.C = cur; p.bias = e.pre_b; p.M = L; p.N = 512; p.K = 13 * 512; p.lda = 512; p.ldc = 512;
    p.tap_off[0] = -12;
    launch_gemm(p, st);
  }

  int rows = L;
  const int RM = 50;                       // largest causal reach of a ResBlock1 conv: (11-1)*5
  for (int i = 0; i < 5; ++i) {
    const int Ci = ch[i], Co = ch[i + 1], s = ups[i];
    const int Lo = rows * s;
    auto with_margin = [&](int margin) {
      float* b = ws.alloc_f((long long)(margin + Lo) * Co);
      launch_fill(b, (long long)margin * Co, 0.f, st);
      return b + (long long)margin * Co;
    };
    float* x = with_margin(RM);
    {
      GemmParams p;                        // SiLU -> FishTransConvNet (k = 2s): rows t-1, t
      p.A = cur; p.W = e.ups_w[i]; p.C = x; p.bias = e.ups_b[i]; p.M = rows; p.N = s * Co; p.K = 2 * Ci; p.lda = Ci;
      p.ldc = (long long)s * Co; p.tap_off[0] = -1; p.prologue = PRO_SILU;
      launch_gemm(p, st);
    }
    float* tmpb[3];
    float* xb[3];
    for (int j = 0; j < 3; ++j) { tmpb[j] = with_margin(RM); xb[j] = with_margin(RM); }

    for (int d = 0; d < 3; ++d) {
      GemmParams p1[3], p2[3];
      for (int j = 0; j < 3; ++j) {
        const ResConvW& w1 = e.res1[i][j][d];
        const ResConvW& w2 = e.res2[i][j][d];
        const float* in = (d == 0) ? x : xb[j];
        auto setup = [&](GemmParams& p, const ResConvW& w, const float* A, float* C, const float* res) {
          p.A = A; p.W = w.w; p.C = C; p.bias = w.b; p.residual = res; p.M = Lo; p.N = Co; p.lda = Co; p.ldc = Co;
          p.ldr = Co; p.prologue = PRO_SILU;
          if (w.d == 1) {
            p.K = w.k * Co; p.taps = 1; p.tap_off[0] = -(w.k - 1);
          } else {
            p.K = Co; p.taps = w.k;
            for (int t = 0; t < w.k; ++t) p.tap_off[t] = -(w.k - 1 - t) * w.d;
          }
        };
        setup(p1[j], w1, in, tmpb[j], nullptr);
        setup(p2[j], w2, tmpb[j], xb[j], in);
      }
      launch_gemm(p1, 3, st);
      launch_gemm(p2, 3, st);

    }
    // ParallelBlock mean (firefly.py:214-215) into the next level's input buffer
    const int next_margin = (i == 4) ? 12 : 1;
    float* nx = with_margin(next_margin);
    launch_scale_add3(xb[0], xb[1], xb[2], nx, (long long)Lo * Co, 1.f / 3.f, st);

    cur = nx;
    rows = Lo;
  }
  launch_conv_post(cur, e.post_w, e.post_b, wave, rows, st);

}

void Engine::voc_quantizer_decode(const long long* codes, long long ld, int T, float* z, cudaStream_t st) {
  SV_CHECK(finalized[MODEL_VOCODER], "vocoder weights not finalized");
  SV_CHECK(T >= 1, "empty code sequence");
  ws.ensure(voc_ws_bytes(T));
  ws.reset();
  float* zb = ws.alloc_f((long long)(6 + 4 * T) * 512);
  voc_qdecode_impl(*this, codes, ld, T, zb + 6 * 512, st);
  SV_CUDA(cudaMemcpyAsync(z, zb + 6 * 512, (size_t)4 * T * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
}

void Engine::voc_head(const float* z, int L, float* wave, cudaStream_t st) {
  SV_CHECK(finalized[MODEL_VOCODER], "vocoder weights not finalized");
  SV_CHECK(L >= 1, "empty feature sequence");
  ws.ensure(voc_ws_bytes((L + 3) / 4));
  ws.reset();
  float* zb = ws.alloc_f((long long)(12 + L) * 512);
  launch_fill(zb, 12 * 512, 0.f, st);
  SV_CUDA(cudaMemcpyAsync(zb + 12 * 512, z, (size_t)L * 512 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  voc_head_impl(*this, zb + 12 * 512, L, wave, st);
}

void Engine::voc_decode(const long long* codes, long long ld, int T, float* wave, cudaStream_t st) {
  NvtxRange nvtx_("svanon:V window decode");
  SV_CHECK(finalized[MODEL_VOCODER], "vocoder weights not finalized");
  SV_CHECK(T >= 1, "empty code sequence");
  ws.ensure(voc_ws_bytes(T));
  ws.reset();
  float* zb = ws.alloc_f((long long)(12 + 4 * T) * 512);
  launch_fill(zb, 12 * 512, 0.f, st);
  voc_qdecode_impl(*this, codes, ld, T, zb + 12 * 512, st);
  voc_head_impl(*this, zb + 12 * 512, 4 * T, wave, st);
}

// ------------------------------------------------------------------------------------------ stage A (multi-token)
// BaseTransformer.forward_generate over new tokens (dual_ar_stream.py:312-356) without the heads, for n streams at once:
// stream i contributes M[i] rows of x (concatenated in stream order) at positions pos0[i].. of ITS sequence.  The dense
// projections run over all rows in one GEMM each (one pass over the weights for everybody); RoPE, KV append and the causal
// attention run per stream on its row slice and its own cache.  x is updated in place to the last layer's residual.
void Engine::ar_forward_tokens_many(Stream* const* ss, const int* M, const int* pos0, int n, float* x, cudaStream_t st) {
  NvtxRange nvtx_("svanon:A prefill tokens");
  int total = 0;
  for (int i = 0; i < n; ++i) {
    SV_CHECK(pos0[i] + M[i] <= ss[i]->max_seq, "sequence position exceeds the KV cache (max_seq_len)");
    total += M[i];
  }
  float* nrm = ws.alloc_f((long long)total * AR_DIM);
  float* qkv = ws.alloc_f((long long)total * 3 * AR_DIM);
  float* y = ws.alloc_f((long long)total * AR_DIM);
  float* h13 = ws.alloc_f((long long)total * 2 * AR_INTER);
  float* gbuf = ws.alloc_f((long long)total * AR_INTER);
  for (int l = 0; l < AR_LAYERS; ++l) {
    const ArLayerWeights& L = ar.slow[l];
    launch_rmsnorm(x, nrm, L.attn_norm, total, AR_DIM, AR_NORM_EPS, st);
    GemmParams p;
    p.A = nrm; p.W = L.wqkv; p.C = qkv; p.M = total; p.N = 3 * AR_DIM; p.K = AR_DIM; p.lda = AR_DIM; p.ldc = 3 * AR_DIM;
    launch_gemm(p, st);
    int r0 = 0;
    for (int i = 0; i < n; ++i) {
      Stream& s = *ss[i];
      const long long layer_stride = (long long)AR_HEADS * s.max_seq * HEAD_DIM;
      float* q_i = qkv + (long long)r0 * 3 * AR_DIM;
      launch_rope_qk(q_i, ar.rope, M[i], AR_HEADS, pos0[i], st);
      launch_kv_append(q_i, M[i], AR_HEADS, s.kc + l * layer_stride, s.vc + l * layer_stride, s.max_seq, pos0[i], st);
      launch_attention(q_i, 3 * AR_DIM, s.kc + l * layer_stride, s.vc + l * layer_stride, (long long)s.max_seq * HEAD_DIM,
                       HEAD_DIM, y + (long long)r0 * AR_DIM, AR_DIM, M[i], pos0[i], AR_HEADS, 1 << 30, st);
      r0 += M[i];
    }
    GemmParams po;
    po.A = y; po.W = L.wo; po.C = x; po.residual = x; po.M = total; po.N = AR_DIM; po.K = AR_DIM; po.lda = AR_DIM;
    po.ldc = AR_DIM; po.ldr = AR_DIM;
    launch_gemm(po, st);
    launch_rmsnorm(x, nrm, L.ffn_norm, total, AR_DIM, AR_NORM_EPS, st);
    GemmParams p1;
    p1.A = nrm; p1.W = L.w1; p1.C = h13; p1.M = total; p1.N = AR_INTER; p1.K = AR_DIM; p1.lda = AR_DIM; p1.ldc = 2 * AR_INTER;
    GemmParams p13[2] = {p1, p1};            // w1 and w3 side by side in one launch
    p13[1].W = L.w3; p13[1].C = h13 + AR_INTER;
    launch_gemm(p13, 2, st);
    launch_silu_mul(h13, gbuf, total, AR_INTER, st);
    GemmParams p2;
    p2.A = gbuf; p2.W = L.w2; p2.C = x; p2.residual = x; p2.M = total; p2.N = AR_DIM; p2.K = AR_INTER; p2.lda = AR_INTER;
    p2.ldc = AR_DIM; p2.ldr = AR_DIM;
    launch_gemm(p2, st);
  }
}

void Engine::ar_forward_tokens(Stream& s, float* x, int M, int pos0, cudaStream_t st) {
  Stream* one = &s;
  ar_forward_tokens_many(&one, &M, &pos0, 1, x, st);
}

// Token rows of DualARWrapper.prefill_prompt (dual_ar_stream.py:764-796) for one stream, written to x (room for
// n_tok + 2 rows): 33 speaker rows, then (condition, audio) pairs with the audio rows delayed by `delay` frames.  Also
// keeps the speaker rows and cached_ref_emb / cached_new_audio_emb on the stream.  Returns n_tok.
int Engine::ar_prompt_rows(Stream& s, const long long* ref_content, const int* ref_audio, int T, const float* style,
                           const float* timbre, float* x, cudaStream_t st) {
  const int d = s.delay;
  SV_CHECK(T >= 1 && T > d, "prompt must be longer than the delay");
  const int n_tok = AR_SPK_TOKENS + 2 * T - (d == 0 ? 1 : 0);
  SV_CHECK(n_tok <= s.max_seq, "prompt does not fit the KV cache");
  // speaker rows: context_in(timbre) 32 tokens, style_in(style) 1 token
  {
    GemmParams p;
    p.A = timbre; p.W = ctx_w; p.C = x; p.bias = ctx_b; p.M = 32; p.N = AR_DIM; p.K = 128; p.lda = 128; p.ldc = AR_DIM;
    launch_gemm(p, st);
    GemmParams q;
    q.A = style; q.W = style_w; q.C = x + 32 * AR_DIM; q.bias = style_b; q.M = 1; q.N = AR_DIM; q.K = 192; q.lda = 192;
    q.ldc = AR_DIM;
    launch_gemm(q, st);
  }
  launch_copy_rows(x, AR_DIM, s.spk_rows, AR_DIM, AR_SPK_TOKENS, AR_DIM, st);
  float* seq = x + AR_SPK_TOKENS * AR_DIM;
  // condition rows (even), audio rows (odd): wait4start[:d] then embed(ref_audio[:, :T-d])
  launch_gather_rows(ar.cond_emb, ref_content, seq, T, AR_DIM, 2 * AR_DIM, st);
  if (d > 0) launch_copy_rows(w4s, AR_DIM, seq + AR_DIM, 2 * AR_DIM, d, AR_DIM, st);
  // (for d == 0 the last audio row is dropped: n_tok already excludes it, the buffer has room for it)
  launch_embed_codes(ar.codebook_emb, ref_audio, T, seq + (long long)(2 * d + 1) * AR_DIM, T - d, 2 * AR_DIM, st);
  // cached_ref_emb = embed(ref_audio[:, T-d:]) ; for d == 0 cached_new_audio_emb = embed(last frame)
  if (d > 0) launch_embed_codes(ar.codebook_emb, ref_audio + (T - d), T, s.ref_emb_tail, d, AR_DIM, st);
  else launch_embed_codes(ar.codebook_emb, ref_audio + (T - 1), T, s.x_audio, 1, AR_DIM, st);
  return n_tok;
}

// Token rows of DualARWrapper.prefill_src_condition4delay (dual_ar_stream.py:798-815): d conditions interleaved with
// cached_ref_emb, 2d - 1 rows; the last audio row becomes cached_new_audio_emb.
int Engine::ar_delay_rows(Stream& s, const long long* src_content, float* x, cudaStream_t st) {
  const int d = s.delay;
  launch_gather_rows(ar.cond_emb, src_content, x, d, AR_DIM, 2 * AR_DIM, st);
  launch_copy_rows(s.ref_emb_tail, AR_DIM, x + AR_DIM, 2 * AR_DIM, d, AR_DIM, st);
  launch_copy_rows(x + (long long)(2 * d - 1) * AR_DIM, AR_DIM, s.x_audio, AR_DIM, 1, AR_DIM, st);
  return 2 * d - 1;
}

// ARVCWrapper.prefill_prompt (arvc_wrapper.py:100-112) -> DualARWrapper.prefill_prompt (dual_ar_stream.py:764-796).
// ref_content [T] int64 (device), ref_audio [8][T] int32 (device), style [192], timbre [32][128] (device).
void Engine::ar_prefill_prompt(Stream& s, const long long* ref_content, const int* ref_audio, int T, const float* style,
                               const float* timbre, cudaStream_t st) {
  SV_CHECK(finalized[MODEL_AR], "AR weights not finalized");
  const int n_max = AR_SPK_TOKENS + 2 * T;
  ws.ensure(((size_t)(n_max + 8) * 12000 + (1u << 20)) * sizeof(float));
  ws.reset();
  float* x = ws.alloc_f((long long)(n_max + 2) * AR_DIM);
  const int n_tok = ar_prompt_rows(s, ref_content, ref_audio, T, style, timbre, x, st);
  ar_forward_tokens(s, x, n_tok, 0, st);
  s.pos_next = n_tok;
  s.step += 1;
  s.delay_prefilled = false;
}

// ARVCWrapper.prefill_src_condition4delay (arvc_wrapper.py:114-119) -> dual_ar_stream.py:798-815
void Engine::ar_prefill_delay(Stream& s, const long long* src_content, int n, cudaStream_t st) {
  NvtxRange nvtx_("svanon:A prefill delay");
  const int d = s.delay;
  SV_CHECK(n == d, "prefill_src_condition4delay expects exactly `delay` content codes");
  SV_CHECK(d > 0, "prefill_src_condition4delay is only defined for delay > 0");
  ws.ensure(((size_t)(2 * d + 8) * 12000 + (1u << 20)) * sizeof(float));
  ws.reset();
  float* x = ws.alloc_f((long long)2 * d * AR_DIM);
  const int rows = ar_delay_rows(s, src_content, x, st);
  ar_forward_tokens(s, x, rows, s.pos_next, st);
  s.pos_next += rows;
  s.step += 1;
  s.delay_prefilled = true;
}

// Re-prompt of the per-chunk loop (infer_arvc.py:547-564) for the n streams that hit their sequence limit in the same
// step: the prompt kept at set_prompt time, extended by the last `buffer_frames` predicted frames and their source
// content ids, is prefilled again from position 0, followed by the delay prefill.  The reference makes two forward calls
// per stream (prefill_prompt, prefill_src_condition4delay); here the 2d - 1 delay tokens ride at the end of the prompt's
// rows -- the same causal attention over the same cache positions -- and all streams of the step share ONE pass over the
// weights (groups of at most REPROMPT_GROUP streams).
void Engine::reprompt_many(Stream* const* streams, int n_all, Workspace& staging, cudaStream_t st) {
  NvtxRange nvtx_("svanon:A re-prompt");
  constexpr int REPROMPT_GROUP = 16;
  for (int g0 = 0; g0 < n_all; g0 += REPROMPT_GROUP) {
    const int n = std::min(REPROMPT_GROUP, n_all - g0);
    Stream* const* ss = streams + g0;
    int M[REPROMPT_GROUP], pos0[REPROMPT_GROUP], Tn[REPROMPT_GROUP];
    size_t rows_max = 0;
    for (int i = 0; i < n; ++i) {
      Stream& s = *ss[i];
      const int buf = std::min(s.buffer_frames, s.n_pred);
      SV_CHECK(s.n_src - s.delay >= buf, "not enough source history for re-prompting");
      Tn[i] = s.ref_frames + buf;
      rows_max += (size_t)AR_SPK_TOKENS + 2 * Tn[i] + 2 * s.delay + 2;
    }
    ws.ensure(((rows_max + 8) * 12000 + (1u << 20)) * sizeof(float));
    ws.reset();
    float* x = ws.alloc_f((long long)rows_max * AR_DIM);
    long long r0 = 0;
    for (int i = 0; i < n; ++i) {
      Stream& s = *ss[i];
      const int buf = Tn[i] - s.ref_frames;
      int* ext_audio = (int*)staging.alloc_bytes((size_t)8 * Tn[i] * sizeof(int));
      long long* ext_content = (long long*)staging.alloc_bytes((size_t)Tn[i] * sizeof(long long));
      launch_concat_cols(s.ref_audio_dev, s.ref_frames, s.ref_frames, s.pred_hist + (s.n_pred - buf), HIST_CAP, buf,
                         ext_audio, Tn[i], 8, false, st);
      SV_CUDA(cudaMemcpyAsync(ext_content, s.ref_content_dev, (size_t)s.ref_frames * sizeof(long long),
                              cudaMemcpyDeviceToDevice, st));
      SV_CUDA(cudaMemcpyAsync(ext_content + s.ref_frames, s.src_hist + (s.n_src - buf - s.delay),
                              (size_t)buf * sizeof(long long), cudaMemcpyDeviceToDevice, st));
      float* xi = x + r0 * AR_DIM;
      int rows = ar_prompt_rows(s, ext_content, ext_audio, Tn[i], s.style_dev, s.timbre_dev, xi, st);
      if (s.delay > 0) rows += ar_delay_rows(s, s.src_hist + (s.n_src - s.delay), xi + (long long)rows * AR_DIM, st);
      M[i] = rows;
      pos0[i] = 0;
      r0 += rows;
    }
    ar_forward_tokens_many(ss, M, pos0, n, x, st);
    for (int i = 0; i < n; ++i) {
      Stream& s = *ss[i];
      s.pos_next = M[i];
      s.step += s.delay > 0 ? 2 : 1;       // prefill_prompt and prefill_src_condition4delay each count as a sampler step
      s.delay_prefilled = s.delay > 0;
    }
  }
}

void Engine::reprompt(Stream& s, Workspace& staging, cudaStream_t st) {
  Stream* one = &s;
  reprompt_many(&one, 1, staging, st);
}

// DualARWrapper.decode_one (dual_ar_stream.py:817-837) for `batch` independent streams in one launch.
void Engine::ar_decode_step(Stream* const* streams, int batch, cudaStream_t st) {
  NvtxRange nvtx_("svanon:A decode (persistent kernel)");
  SV_CHECK(finalized[MODEL_AR], "AR weights not finalized");
  SV_CHECK(batch == 1 || batch == 2 || batch == 4, "decode batch must be 1, 2 or 4");
  ArDecodeArgs a = ar;
  int max_keys = 0;
  for (int b = 0; b < batch; ++b) {
    Stream& s = *streams[b];
    SV_CHECK(s.pos_next + 2 <= s.max_seq, "KV cache full: re-prompt before decoding further");
    SV_CHECK(s.max_seq == streams[0]->max_seq, "batched streams must share max_seq_len");
    ArStreamDev& sd = a.s[b];
    sd.kc = s.kc; sd.vc = s.vc; sd.fkc = s.fkc; sd.fvc = s.fvc; sd.x_audio = s.x_audio;
    sd.content_id = s.step_content_id; sd.cond_row = s.step_cond_row; sd.noise = s.step_noise; sd.out_codes = s.codes_dev;
    SV_CHECK(sd.content_id || sd.cond_row, "decode step without a content id");
    sd.pos = s.pos_next; sd.step = s.step; sd.seed = s.seed;
    sd.temperature = s.temperature; sd.top_p = s.top_p;
    max_keys = std::max(max_keys, s.pos_next + 2);
  }
  a.max_seq = streams[0]->max_seq;
  int nsplit = std::max(1, num_sms / (batch * AR_HEADS));
  nsplit = std::min(nsplit, 16);
  nsplit = std::min(nsplit, std::max(1, max_keys / 16));
  a.nsplit = nsplit;
  a.prof = ar_prof;
  {
    static const float keep = [] {
      const char* e = getenv("SVANON_AR_KEEP_FRACTION");   // tuning knob
      return e ? (float)atof(e) : 0.25f;
    }();
    a.keep_fraction = keep;
  }
  if (debug_logits) { a.dbg_slow_logits = dbg_slow_logits; a.dbg_hidden = dbg_hidden; a.dbg_fast_logits = dbg_fast_logits; }
  else { a.dbg_slow_logits = nullptr; a.dbg_hidden = nullptr; a.dbg_fast_logits = nullptr; }
  if (batch == 1 && ar_variant == 1 && ar_decode_staged_supported(num_sms)) {
    launch_ar_decode_staged(a, num_sms, st);
  } else {
    launch_ar_decode(a, batch, num_sms, st);
  }

  for (int b = 0; b < batch; ++b) {
    streams[b]->pos_next += 2;
    streams[b]->step += 1;
  }
}

}  // namespace svanon
