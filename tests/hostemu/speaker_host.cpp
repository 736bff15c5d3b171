// TEST INFRASTRUCTURE ONLY -- a g++ build of streamvoiceanon_b200/csrc/speaker.hpp (the speaker encoders of the prompt
// path) in which every element-parallel functor runs as a host loop and the engine's GEMM is a three-loop restatement of
// the GemmParams contract (common.cuh).  It exists so that the HOST ORCHESTRATION of that file -- buffer shapes, zero
// margins, weight repacking, GEMM descriptors, functor arguments -- is held to the reference-generated fixtures
// (tests/golden/style_vec.npz, timbre_latent.npz) by `pytest -m "not gpu"` on a machine without a GPU.
// It is built by tests/test_speaker_hostemu.py into tests/hostemu/_build/ and loaded by that test alone; the product
// (streamvoiceanon_b200/_lib.py) only ever loads libsvanon_b200.so and has no CPU path.
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../streamvoiceanon_b200/csrc/speaker.hpp"

using namespace svanon;

namespace {

struct HostTensor {
  std::vector<float> data;
  std::vector<long long> shape;
};

struct HostBK {
  std::map<std::string, HostTensor>& w;
  std::vector<std::unique_ptr<std::vector<float>>>& keep_f;
  std::vector<std::unique_ptr<std::vector<double>>>& keep_d;
  std::vector<std::unique_ptr<std::vector<float>>> scratch;
  long long gemms = 0, pfors = 0, gemm_flops = 0;
  bool tf32x3 = std::getenv("HOSTEMU_TF32X3") != nullptr;      // restate the 3xTF32 products of the tensor-core GEMM
  static float head(float v) {
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    std::memcpy(&v, &u, 4);
    return v;
  }

  const HostTensor& get(const std::string& name, std::initializer_list<long long> shape) {
    auto it = w.find(name);
    if (it == w.end()) throw Error("missing tensor '" + name + "'");
    if (it->second.shape != std::vector<long long>(shape)) throw Error("tensor '" + name + "' has an unexpected shape");
    return it->second;
  }
  std::vector<float> fetch(const std::string& name, std::initializer_list<long long> shape) { return get(name, shape).data; }
  const float* dev(const std::string& name, std::initializer_list<long long> shape) { return get(name, shape).data.data(); }
  const float* upload(const std::vector<float>& v) {
    keep_f.emplace_back(new std::vector<float>(v));
    return keep_f.back()->data();
  }
  const double* upload_d(const std::vector<double>& v) {
    keep_d.emplace_back(new std::vector<double>(v));
    return keep_d.back()->data();
  }
  float* alloc(long long n) {
    // poisoned, so that a read of a row nobody wrote shows up as NaN in the result
    scratch.emplace_back(new std::vector<float>((size_t)n + 16, std::nanf("")));
    return scratch.back()->data();
  }
  template <class F>
  void pfor(long long n, const F& f) {
    ++pfors;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) f(i);
  }
  // C[m][n] = bias[n] + sum_t sum_k A[(m * a_row_step + tap_off[t]) * lda + k] * W[t][n][k]   (common.cuh GemmParams)
  void gemm(const GemmParams& p) {
    ++gemms;
    gemm_flops += 2LL * p.M * p.N * p.K * p.taps;
    SV_CHECK(p.K % 16 == 0 && p.K > 0, "gemm K must be a positive multiple of 16");
    SV_CHECK(p.lda % 4 == 0, "gemm lda must be a multiple of 4");
    SV_CHECK(p.taps >= 1 && p.taps <= MAX_TAPS, "gemm taps");
    SV_CHECK(p.prologue == PRO_NONE && p.act == ACT_NONE && !p.gamma && !p.residual && !p.accumulate && p.out_scale == 1.f &&
                 p.seg_rows == 0, "host GEMM restatement: only bias epilogues are used by speaker.hpp");
    SV_CHECK((reinterpret_cast<uintptr_t>(p.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.W) & 15) == 0, "16-byte aligned operands");
#pragma omp parallel for schedule(static)
    for (int m = 0; m < p.M; ++m)
      for (int n = 0; n < p.N; ++n) {
        float acc = 0.f;
        for (int t = 0; t < p.taps; ++t) {
          const float* a = p.A + ((long long)m * p.a_row_step + p.tap_off[t]) * p.lda;
          const float* wr = p.W + ((long long)t * p.N + n) * p.K;
          if (!tf32x3) {
            for (int k = 0; k < p.K; ++k) acc += a[k] * wr[k];
          } else {
            // the tensor-core kernel's arithmetic (gemm_tc.cu): operands split into a 10-bit-mantissa head and the
            // remainder, products hi*hi + hi*lo + lo*hi (lo*lo dropped), fp32 accumulation
            for (int k = 0; k < p.K; ++k) {
              const float ah = head(a[k]), al = a[k] - ah, bh = head(wr[k]), bl = wr[k] - bh;
              acc += ah * bh + ah * bl + al * bh;
            }
          }
        }
        if (p.bias) acc += p.bias[n];
        p.C[(long long)m * p.ldc + n] = acc;
      }
  }
};

struct HostModel {
  std::map<std::string, HostTensor> w;
  std::vector<std::unique_ptr<std::vector<float>>> keep_f;
  std::vector<std::unique_ptr<std::vector<double>>> keep_d;
  spk::StyleNet style;
  spk::TimbreNet timbre;
};

HostModel g_models[2];
std::string g_err;

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

}  // namespace

extern "C" {

const char* hostemu_last_error() { return g_err.c_str(); }

int hostemu_load_tensor(int model, const char* name, const float* data, int rank, const long long* shape) {
  return guarded([&] {
    HostTensor t;
    t.shape.assign(shape, shape + rank);
    long long n = 1;
    for (auto s : t.shape) n *= s;
    t.data.assign(data, data + n);
    g_models[model].w[name] = std::move(t);
  });
}

int hostemu_finalize(int model) {
  return guarded([&] {
    HostModel& m = g_models[model];
    HostBK bk{m.w, m.keep_f, m.keep_d};
    if (model == 0) spk::style_finalize(bk, m.style);
    else spk::timbre_finalize(bk, m.timbre);
  });
}

// counts[0] = GEMM calls, counts[1] = functor launches, counts[2] = GEMM FLOPs of the call
int hostemu_style_vector(const float* wave, long long n, float* out192, long long* counts) {
  return guarded([&] {
    HostModel& m = g_models[0];
    HostBK bk{m.w, m.keep_f, m.keep_d};
    spk::style_forward(bk, m.style, wave, n, out192);
    if (counts) { counts[0] = bk.gemms; counts[1] = bk.pfors; counts[2] = bk.gemm_flops; }
  });
}

int hostemu_kaldi_fbank(const float* wave, long long n, float* feat /*[T][80]*/) {
  return guarded([&] {
    HostModel& m = g_models[0];
    HostBK bk{m.w, m.keep_f, m.keep_d};
    spk::kaldi_fbank_rows(bk, m.style, wave, n, feat);
  });
}

int hostemu_campplus(const float* feat /*[T][80]*/, long long T, int len, float* out192) {
  return guarded([&] {
    HostModel& m = g_models[0];
    HostBK bk{m.w, m.keep_f, m.keep_d};
    spk::campplus_forward_rows(bk, m.style, feat, T, len, out192);
  });
}

int hostemu_timbre_latent(const float* wave, long long n, long long wave_len, float* out, int* indices, float* z, long long* counts) {
  return guarded([&] {
    HostModel& m = g_models[1];
    HostBK bk{m.w, m.keep_f, m.keep_d};
    spk::timbre_forward(bk, m.timbre, wave, n, wave_len, out, indices, z);
    if (counts) { counts[0] = bk.gemms; counts[1] = bk.pfors; counts[2] = bk.gemm_flops; }
  });
}

}  // extern "C"
