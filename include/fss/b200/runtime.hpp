// SPDX-License-Identifier: Apache-2.0
// fss/b200/runtime.hpp -- glue between the header-only C++ surface and the C ABI (include/fssb200.h):
// a process-wide cache of evaluator contexts keyed by the full parameter set, and error translation.
// The C ABI never throws; this C++ layer turns non-zero return codes into std::runtime_error (the
// reference only asserts, dpf.cuh:209).
#pragma once
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>
#include <cuda_runtime.h>

#include "../../fssb200.h"

namespace fss::b200 {

inline void Check(int rc, const char *what) {
  if (rc != 0) throw std::runtime_error(std::string(what) + ": " + fssb200_strerror(rc));
}

struct ParamsLess {
  bool operator()(const fssb200_params &a, const fssb200_params &b) const { return std::memcmp(&a, &b, sizeof(a)) < 0; }
};

// Contexts are immutable (round keys + parameters, ~1.5 KB of host memory, no device or pinned memory: the staging
// arenas of the host-array members belong to the library's per-device pool and are sized per call), so one cached
// context per parameter set serves every thread and stream; they live until process exit.
inline fssb200_ctx *ContextFor(fssb200_params p, int device = -1) {  // device < 0: the current device
  static std::mutex mu;
  static std::map<fssb200_params, fssb200_ctx *, ParamsLess> cache;
  int dev = device;
  if (dev < 0) cudaGetDevice(&dev);
  p.device = dev;
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(p);
  if (it != cache.end()) return it->second;
  fssb200_ctx *ctx = nullptr;
  Check(fssb200_ctx_create(&p, &ctx), "fssb200_ctx_create");
  cache.emplace(p, ctx);
  return ctx;
}

template <int in_bits, typename Group, typename Prg, typename In>
fssb200_params MakeParams(int scheme, const Prg &prg, int pred = FSSB200_PRED_LT, const int4 *hash_key = nullptr) {
  static_assert(sizeof(In) == 1 || sizeof(In) == 2 || sizeof(In) == 4 || sizeof(In) == 8 || sizeof(In) == 16);
  fssb200_params p;
  std::memset(&p, 0, sizeof(p));
  p.scheme = scheme;
  p.in_bits = in_bits;
  p.in_bytes = sizeof(In);
  p.group = Group::kFssB200Group;
  p.mod_lo = Group::kFssB200ModLo;
  p.mod_hi = Group::kFssB200ModHi;
  p.prg = Prg::kFssB200Prg;
  p.pred = pred;
  prg.FssB200Key(p.prg_key);
  if (hash_key) std::memcpy(p.hash_key, hash_key, 16);
  return p;
}

// Multi-device members (`EvalBatchMulti`): contexts of one parameter set on `ndev` devices + the per-device error
// codes folded into one exception.
struct MultiCall {
  std::vector<fssb200_ctx *> ctxs;
  std::vector<int> rcs;
  MultiCall(const fssb200_params &p, int ndev, const int *devices) : rcs(size_t(ndev), 0) {
    for (int d = 0; d < ndev; ++d) ctxs.push_back(ContextFor(p, devices ? devices[d] : d));
  }
  void Check(int rc, const char *what) const {
    if (rc == 0) return;
    std::string msg = std::string(what) + ":";
    for (size_t d = 0; d < rcs.size(); ++d)
      if (rcs[d]) msg += " device " + std::to_string(d) + ": " + fssb200_strerror(rcs[d]) + ";";
    throw std::runtime_error(msg);
  }
};

// A 16-byte value the reference passes by value (a seed, an output CW) as a stream-ordered device temp.
struct DeviceBlock {
  int4 *ptr = nullptr;
  cudaStream_t stream;
  DeviceBlock(int4 v, cudaStream_t s) : stream(s) {
    if (cudaMallocAsync(reinterpret_cast<void **>(&ptr), sizeof(int4), s) != cudaSuccess) throw std::bad_alloc();
    cudaMemcpyAsync(ptr, &v, sizeof(int4), cudaMemcpyHostToDevice, s);  // pageable source: staged before return
  }
  ~DeviceBlock() {
    if (ptr) cudaFreeAsync(ptr, stream);
  }
  DeviceBlock(const DeviceBlock &) = delete;
  DeviceBlock &operator=(const DeviceBlock &) = delete;
};

// A device array with blocking transfers, for the single-key members that take host arrays.
template <typename T>
struct DeviceArray {
  T *ptr = nullptr;
  explicit DeviceArray(size_t n) {
    if (cudaMalloc(reinterpret_cast<void **>(&ptr), sizeof(T) * (n ? n : 1)) != cudaSuccess) throw std::bad_alloc();
  }
  ~DeviceArray() {
    if (ptr) cudaFree(ptr);
  }
  void Upload(size_t at, const T *src, size_t n) { cudaMemcpy(ptr + at, src, sizeof(T) * n, cudaMemcpyHostToDevice); }
  void Download(size_t at, T *dst, size_t n) { cudaMemcpy(dst, ptr + at, sizeof(T) * n, cudaMemcpyDeviceToHost); }
  DeviceArray(const DeviceArray &) = delete;
  DeviceArray &operator=(const DeviceArray &) = delete;
};

}  // namespace fss::b200
