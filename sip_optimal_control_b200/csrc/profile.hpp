// Per-kernel device timing for bench tooling: when enabled on a handle, every
// launch group of OUR kernels is bracketed by a pair of CUDA events recorded
// on the launching stream.  Disabled (the default) it costs one branch.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace sipoc {

class Profiler {
 public:
  struct Summary {
    std::string name;
    double total_ms = 0.0;
    int64_t launches = 0;
  };

  ~Profiler() { release(); }

  void enable(bool on) { on_ = on; }
  bool enabled() const { return on_; }

  void begin(const char *name, cudaStream_t s) {
    if (!on_) return;
    Span sp;
    sp.name = name;
    if (cudaEventCreate(&sp.a) != cudaSuccess || cudaEventCreate(&sp.b) != cudaSuccess) return;
    cudaEventRecord(sp.a, s);
    spans_.push_back(sp);
    open_ = true;
  }
  void end(cudaStream_t s) {
    if (!on_ || !open_) return;
    cudaEventRecord(spans_.back().b, s);
    open_ = false;
  }

  // Synchronises on every recorded span and folds them by kernel name.
  const std::vector<Summary> &summarise() {
    summary_.clear();
    for (Span &sp : spans_) {
      float ms = 0.f;
      if (cudaEventSynchronize(sp.b) != cudaSuccess) continue;
      if (cudaEventElapsedTime(&ms, sp.a, sp.b) != cudaSuccess) continue;
      Summary *dst = nullptr;
      for (Summary &s : summary_)
        if (s.name == sp.name) dst = &s;
      if (dst == nullptr) {
        summary_.push_back(Summary{sp.name, 0.0, 0});
        dst = &summary_.back();
      }
      dst->total_ms += ms;
      dst->launches += 1;
    }
    return summary_;
  }
  const std::vector<Summary> &summary() const { return summary_; }

  void reset() { release(); }

 private:
  struct Span {
    const char *name;
    cudaEvent_t a = nullptr, b = nullptr;
  };
  void release() {
    for (Span &sp : spans_) {
      if (sp.a) cudaEventDestroy(sp.a);
      if (sp.b) cudaEventDestroy(sp.b);
    }
    spans_.clear();
    summary_.clear();
    open_ = false;
  }
  bool on_ = false, open_ = false;
  std::vector<Span> spans_;
  std::vector<Summary> summary_;
};

// RAII span; `p` may be nullptr.
struct ProfScope {
  Profiler *p;
  cudaStream_t s;
  ProfScope(Profiler *prof, const char *name, cudaStream_t stream) : p(prof), s(stream) {
    if (p) p->begin(name, s);
  }
  ~ProfScope() {
    if (p) p->end(s);
  }
};

}  // namespace sipoc
