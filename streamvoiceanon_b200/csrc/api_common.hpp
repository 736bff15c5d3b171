// Shared between the C-ABI translation units (api.cu, batch.cu): handle types, error capture, argument staging.
#pragma once
#include "../../include/svanon.h"

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "engine.hpp"

using namespace svanon;

struct svanon_engine {
  Engine eng;
  Workspace staging;      // host<->device staging of API arguments
  std::mutex mu;
};
struct svanon_stream {
  Stream st;
  svanon_engine* owner = nullptr;
};

namespace svanon {
extern thread_local std::string g_api_err;
}

namespace {

template <typename F>
int guarded(F&& f) {
  try {
    f();
    return 0;
  } catch (const std::exception& e) {
    svanon::g_api_err = e.what();
    return 1;
  } catch (...) {
    svanon::g_api_err = "unknown error";
    return 1;
  }
}

inline bool on_device(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// Per-call argument staging.  Inputs in host memory are copied to the staging arena; outputs in host memory are
// produced in the arena and copied back (followed by one stream synchronisation) when the scope ends.
struct Args {
  svanon_engine* h;
  cudaStream_t st;
  struct Out { void* host; void* dev; size_t bytes; };
  std::vector<Out> outs;
  Args(svanon_engine* h_, void* stream, size_t budget) : h(h_), st((cudaStream_t)stream) {
    SV_CUDA(cudaSetDevice(h->eng.device));
    h->staging.ensure(budget + (1u << 20));
    h->staging.reset();
  }
  template <typename T>
  const T* in(const T* p, size_t n) {
    if (!p) return nullptr;
    if (on_device(p)) return p;
    T* d = (T*)h->staging.alloc_bytes(n * sizeof(T));
    SV_CUDA(cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, st));
    return d;
  }
  template <typename T>
  T* out(T* p, size_t n) {
    if (on_device(p)) return p;
    T* d = (T*)h->staging.alloc_bytes(n * sizeof(T));
    outs.push_back({(void*)p, (void*)d, n * sizeof(T)});
    return d;
  }
  void finish() {
    if (outs.empty()) return;
    for (auto& o : outs) SV_CUDA(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, st));
    SV_CUDA(cudaStreamSynchronize(st));
    outs.clear();
  }
};

template <typename T>
T* dmalloc(size_t n) {
  T* p = nullptr;
  SV_CUDA(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  return p;
}

}  // namespace
