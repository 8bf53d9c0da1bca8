// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY (never shipped): a mock of the slice of the CUDA runtime that fss_b200/csrc/host_api.cu uses,
// so that the host-side pipeline logic (worker crew, staging ring, piece / chunk hand-offs, arena pool, adaptive direct
// pieces) runs on the CPU -- under ThreadSanitizer too -- in the GPU-less build container.
//   * a stream is a thread that executes its queue in order, with small random delays (copies and "kernels" really are
//     asynchronous to the submitting thread, so a ring slot overwritten before its copy ran IS caught);
//   * "device memory" is host memory; pinned memory is host memory recorded in a registry (cudaPointerGetAttributes);
//   * the device entry points host_api.cu calls (fssb200_eval, fssb200_eval_packed, ...) are defined by the test as
//     stream operations that DIGEST every key's bytes, so a result is right only if every byte of every key reached the
//     "device" intact, in the format the launch claims.
#pragma once
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <mutex>
#include <random>
#include <thread>

#include "../../include/fssb200.h"

typedef int cudaError_t;
enum : int { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDefault = 4 };
enum : unsigned { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2, cudaMemoryTypeManaged = 3 };
struct cudaPointerAttributes {
  cudaMemoryType type;
};

struct MockStream {
  std::mutex mu;
  std::condition_variable cv, cv_idle;
  std::deque<std::function<void()>> q;
  bool busy = false, stop = false;
  std::thread th;
  std::minstd_rand rng{12345};
  MockStream() : th([this] { loop(); }) {}
  ~MockStream() {
    {
      std::lock_guard<std::mutex> l(mu);
      stop = true;
    }
    cv.notify_all();
    th.join();
  }
  void push(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> l(mu);
      q.push_back(std::move(f));
    }
    cv.notify_all();
  }
  void sync() {
    std::unique_lock<std::mutex> l(mu);
    cv_idle.wait(l, [this] { return q.empty() && !busy; });
  }
  void loop() {
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> l(mu);
        cv.wait(l, [this] { return stop || !q.empty(); });
        if (q.empty()) return;
        f = std::move(q.front());
        q.pop_front();
        busy = true;
      }
      if (rng() % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(rng() % 60));  // "DMA latency"
      f();
      {
        std::lock_guard<std::mutex> l(mu);
        busy = false;
        if (q.empty()) cv_idle.notify_all();
      }
    }
  }
};
typedef MockStream *cudaStream_t;

struct MockEvent {
  std::atomic<uint64_t> recorded{0}, done{0};
};
typedef MockEvent *cudaEvent_t;

namespace mockcuda {
inline std::mutex &reg_mu() { static std::mutex m; return m; }
inline std::map<const uint8_t *, size_t> &pinned() { static std::map<const uint8_t *, size_t> m; return m; }
inline void register_pinned(const void *p, size_t n) {
  std::lock_guard<std::mutex> l(reg_mu());
  pinned()[static_cast<const uint8_t *>(p)] = n;
}
inline void unregister_pinned(const void *p) {
  std::lock_guard<std::mutex> l(reg_mu());
  pinned().erase(static_cast<const uint8_t *>(p));
}
inline bool is_pinned(const void *p) {
  std::lock_guard<std::mutex> l(reg_mu());
  auto &m = pinned();
  auto it = m.upper_bound(static_cast<const uint8_t *>(p));
  if (it == m.begin()) return false;
  --it;
  return static_cast<const uint8_t *>(p) < it->first + it->second;
}
inline MockStream *default_stream() { static MockStream *s = new MockStream(); return s; }
inline MockStream *S(cudaStream_t s) { return s ? s : default_stream(); }
inline std::atomic<long> &live_allocs() { static std::atomic<long> n{0}; return n; }
}  // namespace mockcuda

inline cudaError_t cudaGetDevice(int *d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "mock cuda error"; }
inline cudaError_t cudaMalloc(void *pp, size_t n) {
  void *p = std::malloc(n ? n : 1);
  if (!p) return cudaErrorMemoryAllocation;
  std::memset(p, 0xA5, n);  // stale "device" memory is never zero
  mockcuda::register_pinned(p, n);  // (registry = "not pageable": pinned host memory and device memory)
  *static_cast<void **>(pp) = p;
  ++mockcuda::live_allocs();
  return cudaSuccess;
}
inline cudaError_t cudaFree(void *p) {
  if (p) {
    mockcuda::unregister_pinned(p);
    --mockcuda::live_allocs();
  }
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaHostAlloc(void **pp, size_t n, unsigned) {
  void *p = std::malloc(n ? n : 1);
  if (!p) return cudaErrorMemoryAllocation;
  std::memset(p, 0x5A, n);
  mockcuda::register_pinned(p, n);
  *pp = p;
  ++mockcuda::live_allocs();
  return cudaSuccess;
}
inline cudaError_t cudaFreeHost(void *p) {
  if (p) {
    mockcuda::unregister_pinned(p);
    --mockcuda::live_allocs();
  }
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new MockStream(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t s) { mockcuda::S(s)->sync(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new MockEvent(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s) {
  const uint64_t seq = e->recorded.fetch_add(1, std::memory_order_acq_rel) + 1;
  mockcuda::S(s)->push([e, seq] {
    uint64_t cur = e->done.load(std::memory_order_relaxed);
    while (cur < seq && !e->done.compare_exchange_weak(cur, seq, std::memory_order_release)) {}
  });
  return cudaSuccess;
}
inline cudaError_t cudaEventQuery(cudaEvent_t e) {
  return e->done.load(std::memory_order_acquire) >= e->recorded.load(std::memory_order_acquire) ? cudaSuccess : cudaErrorNotReady;
}
inline cudaError_t cudaEventSynchronize(cudaEvent_t e) {
  while (cudaEventQuery(e) != cudaSuccess) std::this_thread::yield();
  return cudaSuccess;
}
// Pinned or device memory on both sides: asynchronous.  A pageable host buffer on either side: the real runtime is
// synchronous with respect to the host for such copies -- run the copy in stream order and wait for it.
inline cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t s) {
  mockcuda::S(s)->push([dst, src, n] { std::memcpy(dst, src, n); });
  if (!mockcuda::is_pinned(src) || !mockcuda::is_pinned(dst)) mockcuda::S(s)->sync();
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy2DAsync(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t s) {
  mockcuda::S(s)->push([=] {
    for (size_t r = 0; r < height; ++r)
      std::memcpy(static_cast<uint8_t *>(dst) + r * dpitch, static_cast<const uint8_t *>(src) + r * spitch, width);
  });
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void *dst, int v, size_t n, cudaStream_t s) {
  mockcuda::S(s)->push([=] { std::memset(dst, v, n); });
  return cudaSuccess;
}
inline cudaError_t cudaPointerGetAttributes(cudaPointerAttributes *a, const void *p) {
  a->type = mockcuda::is_pinned(p) ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
  return cudaSuccess;
}

namespace fssb200 {
struct KParams {};
}  // namespace fssb200
