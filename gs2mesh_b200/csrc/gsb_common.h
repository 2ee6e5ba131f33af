// Shared host/device helpers for the gs2mesh_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <vector>

#include "gs2mesh_b200.h"

// Opaque handle of include/gs2mesh_b200.h: a copy of the caller's descriptor (owns no device memory).
struct GsbVolume {
  GsbVolumeDesc d;
  uint32_t frame = 0;
  size_t n_bricks = 0;
};

namespace gsb {

constexpr int kTile = 16;          // DGR/cuda_rasterizer/config.h:16-17 (BLOCK_X, BLOCK_Y)
constexpr int kTilePixels = 256;

extern thread_local char g_error[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
  return code;
}

inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define GSB_CUDA_OK(expr)                                                                          \
  do {                                                                                             \
    cudaError_t e__ = (expr);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return ::gsb::fail(GSB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),    \
                         __FILE__, __LINE__);                                                      \
  } while (0)

// Cheap launch check (cudaPeekAtLastError); with `sync` also waits for the stream, like the
// reference's debug mode (auxiliary.h:166-173).
inline int check_launch(const char* what, cudaStream_t s, bool sync) {
  cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess && sync) e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(GSB_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return GSB_OK;
}

// ---------------------------------------------------------------------------------------------
// Optional per-stage device timing (bench.py's roofline numbers): CUDA events recorded on the
// launching stream around each stage when enabled; read back with gsb_profile_collect().
// ---------------------------------------------------------------------------------------------
enum Stage {
  kStPreprocess = 0,
  kStScan,
  kStEmit,
  kStSort,
  kStRanges,
  kStRender,
  kStToU8,
  kStDepthPrep,
  kStMarkBricks,
  kStIntegrate,
  kStCount
};

struct Profiler {
  std::atomic<bool> enabled{false};
  struct Rec {
    int stage;
    cudaEvent_t a, b;
  };
  std::mutex mu;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
};
extern Profiler g_prof;

struct StageTimer {
  cudaStream_t s;
  int stage;
  cudaEvent_t a = nullptr, b = nullptr;
  StageTimer(int stage_, cudaStream_t s_) : s(s_), stage(stage_) {
    if (!g_prof.enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_prof.mu);
    a = g_prof.get();
    b = g_prof.get();
    cudaEventRecord(a, s);
  }
  ~StageTimer() {
    if (!a) return;
    cudaEventRecord(b, s);
    std::lock_guard<std::mutex> lk(g_prof.mu);
    g_prof.recs.push_back({stage, a, b});
  }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Carves consecutive 256-byte aligned regions out of one caller-owned block.
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<char*>(p)) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
  size_t total() const { return align_up(off, 256); }
};

}  // namespace gsb
