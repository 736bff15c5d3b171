// CUDA backend of the speaker encoders (speaker.hpp) and their C-ABI entry points (include/svanon.h).
// Functors run as grid-stride `pfor_kernel` launches on the caller's stream, dense layers as launch_gemm (tcgen05 3xTF32
// for M >= 32, N >= 64; the fp32 pipeline kernel for the N = 32 CAM convolutions), buffers come from the engine workspace.
#include "speaker.hpp"

#include "api_common.hpp"

namespace svanon {
namespace {

template <class F>
__global__ void __launch_bounds__(256) pfor_kernel(long long n, F f) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) f(i);
}

struct CudaBK {
  Engine& e;
  int model;
  cudaStream_t st;

  const Tensor& get(const std::string& name, std::initializer_list<long long> shape) {
    const Tensor& t = e.get(model, name);
    if (t.shape != std::vector<long long>(shape)) throw Error("tensor '" + name + "' has an unexpected shape");
    return t;
  }
  std::vector<float> fetch(const std::string& name, std::initializer_list<long long> shape) {
    const Tensor& t = get(name, shape);
    std::vector<float> h((size_t)t.numel());
    SV_CUDA(cudaMemcpy(h.data(), t.data, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    return h;
  }
  const float* dev(const std::string& name, std::initializer_list<long long> shape) { return get(name, shape).data; }
  const float* upload(const std::vector<float>& v) { return e.upload(v); }
  const double* upload_d(const std::vector<double>& v) {
    double* p = nullptr;
    SV_CUDA(cudaMalloc(&p, v.size() * sizeof(double)));
    e.owned.push_back(reinterpret_cast<float*>(p));
    SV_CUDA(cudaMemcpy(p, v.data(), v.size() * sizeof(double), cudaMemcpyHostToDevice));
    return p;
  }
  float* alloc(long long n) { return e.ws.alloc_f(n); }
  void gemm(const GemmParams& p) { launch_gemm(p, st); }
  template <class F>
  void pfor(long long n, const F& f) {
    if (n <= 0) return;
    const long long blocks = (n + 255) / 256;
    const int grid = (int)std::min<long long>(blocks, (long long)e.num_sms * 32);
    pfor_kernel<F><<<grid, 256, 0, st>>>(n, f);
    SV_LAUNCHED();
  }
};

}  // namespace

void Engine::finalize_style() {
  auto net = std::make_shared<spk::StyleNet>();
  CudaBK bk{*this, MODEL_STYLE, nullptr};
  spk::style_finalize(bk, *net);
  style_net = net;
}

void Engine::finalize_timbre() {
  auto net = std::make_shared<spk::TimbreNet>();
  CudaBK bk{*this, MODEL_TIMBRE, nullptr};
  spk::timbre_finalize(bk, *net);
  timbre_net = net;
}

static const spk::StyleNet& style_of(Engine& e) {
  SV_CHECK(e.finalized[MODEL_STYLE] && e.style_net, "style encoder (CAMPPlus) weights not loaded");
  return *static_cast<const spk::StyleNet*>(e.style_net.get());
}

void Engine::kaldi_fbank(const float* wave, long long n, float* feat_rows, cudaStream_t st) {
  const spk::StyleNet& net = style_of(*this);
  ws.ensure(spk::style_ws_floats(n) * sizeof(float));
  ws.reset();
  CudaBK bk{*this, MODEL_STYLE, st};
  spk::kaldi_fbank_rows(bk, net, wave, n, feat_rows);
}

void Engine::campplus_forward(const float* feat_rows, long long T, int len, float* out, cudaStream_t st) {
  const spk::StyleNet& net = style_of(*this);
  SV_CHECK(T >= 4 && T < (1 << 20), "CAMPPlus: between 4 and 2^20 feature frames");
  ws.ensure(spk::style_ws_floats(spk::FB_WIN + (T - 1) * spk::FB_SHIFT) * sizeof(float));
  ws.reset();
  CudaBK bk{*this, MODEL_STYLE, st};
  spk::campplus_forward_rows(bk, net, feat_rows, T, len, out);
}

void Engine::style_vector(const float* wave, long long n, float* out, cudaStream_t st) {
  NvtxRange nvtx_("svanon:prompt style_vector");
  const spk::StyleNet& net = style_of(*this);
  ws.ensure(spk::style_ws_floats(n) * sizeof(float));
  ws.reset();
  CudaBK bk{*this, MODEL_STYLE, st};
  spk::style_forward(bk, net, wave, n, out);
}

void Engine::timbre_latent(const float* wave, long long n, long long wave_len, float* out, int* indices, cudaStream_t st) {
  NvtxRange nvtx_("svanon:prompt timbre_latent");
  SV_CHECK(finalized[MODEL_TIMBRE] && timbre_net, "timbre encoder (BiCodec speaker encoder) weights not loaded");
  const spk::TimbreNet& net = *static_cast<const spk::TimbreNet*>(timbre_net.get());
  SV_CHECK(n >= spk::TM_NFFT, "timbre encoder: the reference wave is shorter than 1024 samples at 16 kHz");
  ws.ensure(spk::timbre_ws_floats(n) * sizeof(float));
  ws.reset();
  CudaBK bk{*this, MODEL_TIMBRE, st};
  spk::timbre_forward(bk, net, wave, n, wave_len, out, indices);
}

}  // namespace svanon

extern "C" {

int svanon_kaldi_fbank(svanon_engine* e, const float* wave16k, int64_t n_samples, float* feat_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && wave16k && feat_out, "null argument");
    const long long T = spk::style_frames(n_samples);
    SV_CHECK(T >= 1, "kaldi fbank: the wave is shorter than one 25 ms frame (400 samples at 16 kHz)");
    Args a(e, stream, ((size_t)n_samples + (size_t)T * spk::FB_MEL) * 4 + 65536);
    const float* w = a.in(wave16k, (size_t)n_samples);
    float* f = a.out(feat_out, (size_t)T * spk::FB_MEL);
    e->eng.kaldi_fbank(w, n_samples, f, a.st);
    a.finish();
  });
}

int svanon_campplus_forward(svanon_engine* e, const float* feat, int64_t n_frames, int valid_len, float* out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && feat && out && n_frames >= 1, "bad arguments");
    Args a(e, stream, ((size_t)n_frames * spk::FB_MEL + 192) * 4 + 65536);
    const float* f = a.in(feat, (size_t)n_frames * spk::FB_MEL);
    float* o = a.out(out, 192);
    e->eng.campplus_forward(f, n_frames, valid_len, o, a.st);
    a.finish();
  });
}

int svanon_style_vector(svanon_engine* e, const float* wave16k, int64_t n_samples, float* out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && wave16k && out && n_samples >= 1, "bad arguments");
    Args a(e, stream, ((size_t)n_samples + 192) * 4 + 65536);
    const float* w = a.in(wave16k, (size_t)n_samples);
    float* o = a.out(out, 192);
    e->eng.style_vector(w, n_samples, o, a.st);
    a.finish();
  });
}

int svanon_timbre_latent(svanon_engine* e, const float* wave16k, int64_t n_samples, int64_t wave_len, float* latents_out,
                         int32_t* indices_out, void* stream) {
  return guarded([&] {
    SV_CHECK(e && wave16k && latents_out && n_samples >= 1, "bad arguments");
    Args a(e, stream, ((size_t)n_samples + 32 * 128 + 32) * 4 + 65536);
    const float* w = a.in(wave16k, (size_t)n_samples);
    float* o = a.out(latents_out, 32 * 128);
    int* idx = indices_out ? a.out(indices_out, 32) : nullptr;
    e->eng.timbre_latent(w, n_samples, wave_len, o, idx, a.st);
    a.finish();
  });
}

}  // extern "C"
