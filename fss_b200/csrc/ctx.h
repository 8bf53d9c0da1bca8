// SPDX-License-Identifier: Apache-2.0
//
// ctx.h -- the evaluator context behind `fssb200_ctx*` (include/fssb200.h), shared by api.cu (device-pointer
// entry points), host_api.cu (host-buffer entry points) and multi_api.cu (multi-device entry points).
// A context is immutable after fssb200_ctx_create() except for two relaxed counters / preferences, so any
// number of threads may use one context at the same time (the reference's scheme objects are value types
// without mutable state, dpf.cuh:61-66).
#pragma once
#include <atomic>

#if defined(FSSB200_HOST_MOCK)
// tests/host_emul/host_pipe_mock.cpp: the host-side logic of host_api.cu compiled for the CPU against a mock CUDA runtime
// (worker-thread streams, events, a pinned-memory registry) -- how the lock-free pipeline is exercised, also under
// ThreadSanitizer, in the GPU-less build container.  Never defined in the product build.
#include "mock_cuda.h"
#else
#include "dispatch.h"
#endif

struct fssb200_ctx {
  fssb200_params p;
  fssb200::KParams kp;
  int gk;           // group kind (common.cuh)
  int ncw;
  int mul;
  int sm_count;
  int max_smem_optin;
  uint32_t vmask;
  int point_mode;   // PointMode of the key-major point kernels (kernels.cuh); FSSB200_POINT_MODE overrides
  int gen_mode;     // gen kernels: 1 = Cw tiles written by the TMA unit (CwTileOut), 0 = direct stores; FSSB200_GEN_MODE
  std::atomic<uint64_t> launches{0};
  // preference of the host-buffer entry points: keys per pipeline chunk (0 = library default); set by
  // fssb200_ctx_reserve_host.  The staging memory itself belongs to the per-device arena pool (host_api.cu).
  std::atomic<size_t> host_chunk_keys{0};
  // fssb200_ctx_set_host_mode: 0 = adaptive pack / direct pipeline, 1 = reference layout crosses the link as it is,
  // 2 = every chunk is staged (packed) by the host threads
  std::atomic<int> host_mode{0};
  // statistics of the last fssb200_eval_host call of this context (fssb200_ctx_host_stats)
  std::atomic<uint64_t> last_packed_keys{0}, last_direct_keys{0};
  std::atomic<int> last_pack_threads{0};
};

namespace fssb200 {

#define FSS_CUDA_TRY(expr)                     \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return int(e__);   \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace fssb200
