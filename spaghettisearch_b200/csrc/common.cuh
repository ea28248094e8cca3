// Shared plumbing for libspaghetti_gpu: error capture, device buffers, the
// engine object.  Nothing here is a hot path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <mutex>
#include <string>
#include <vector>

#include "spaghetti.h"

namespace ss {

void set_error(const char* fmt, ...);

#define SS_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t _err = (call);                                                          \
    if (_err != cudaSuccess) {                                                          \
      ss::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_err)); \
      return _err == cudaErrorMemoryAllocation ? SS_ERR_OOM : SS_ERR_CUDA;              \
    }                                                                                   \
  } while (0)

#define SS_TRY(call)            \
  do {                          \
    int _rc = (call);           \
    if (_rc < 0) return _rc;    \
  } while (0)

#define SS_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      ss::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// cudaMalloc'ed array; freed on reset/destruction.  cudaMalloc (not the async
// pool) so that the block can be shared through CUDA IPC / peer access.
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { reset(); }
  void reset() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  int alloc(size_t count) {
    reset();
    if (count == 0) count = 1;
    cudaError_t err = cudaMalloc((void**)&p, count * sizeof(T));
    if (err != cudaSuccess) {
      p = nullptr;
      set_error("cudaMalloc(%zu bytes) -> %s", count * sizeof(T), cudaGetErrorString(err));
      cudaGetLastError();
      return SS_ERR_OOM;
    }
    n = count;
    return SS_OK;
  }
  // grow-only: keeps the block when it is already large enough (cudaMalloc/cudaFree of
  // gigabyte blocks costs tens to hundreds of milliseconds and synchronises the device)
  int reserve(size_t count) {
    if (p && n >= count) return SS_OK;
    return alloc(count + count / 8);
  }
  size_t bytes() const { return n * sizeof(T); }
};

struct Timer {  // CUDA-event stopwatch on one stream; accumulates milliseconds
  cudaEvent_t a = nullptr, b = nullptr;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pending;
  double total_ms = 0;
};

inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace ss

struct PagerankState;
struct IndexState;
struct CommState;

struct ss_engine {
  int device = 0;
  uint32_t flags = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::mutex mu;  // offline calls are exclusive; ss_score_batch serialises on it too
  PagerankState* pr = nullptr;
  IndexState* idx = nullptr;
  CommState* comm = nullptr;
};

// RAII device selection for entry points (cgo calls arrive on any OS thread).
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

void pagerank_state_free(PagerankState*);
void index_state_free(IndexState*);
void comm_state_free(CommState*);
