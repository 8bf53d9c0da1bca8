// SPDX-License-Identifier: Apache-2.0
// fss/util.cuh -- drop-in for the reference header of the same name (util.cuh:16-64): 128-bit block
// helpers on CUDA `int4`.  Part of the B200 shim: same names and semantics, new code.
#pragma once
#include <cstdint>
#include <thread>
#if defined(_OPENMP)
#include <omp.h>
#endif
#include <cuda_runtime.h>

#if defined(__CUDACC__)
#define FSS_SHIM_HD __host__ __device__ inline
#else
#define FSS_SHIM_HD inline
#endif

namespace fss::util {

FSS_SHIM_HD int4 Xor(int4 a, int4 b) { return int4{a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w}; }
// "Clamping": bit 0 of .w carries the control bit (util.cuh:30-38).
FSS_SHIM_HD int4 SetLsb(int4 v, bool bit) {
  v.w = bit ? (v.w | 1) : (v.w & ~1);
  return v;
}
FSS_SHIM_HD bool GetLsb(int4 v) { return (v.w & 1) != 0; }

// util.cuh:40-45: the recursion depth at which the reference's CPU EvalAll stops spawning OpenMP tasks (par_depth < 0:
// log2 of the thread count, rounded up).  Kept for source compatibility; the evaluator here does not use it.
inline int ResolveParDepth(int par_depth) {
  if (par_depth >= 0) return par_depth;
#if defined(_OPENMP)
  const int threads = omp_get_max_threads();
#else
  const int threads = static_cast<int>(std::thread::hardware_concurrency());
#endif
  int d = 0;
  while ((1 << d) < threads) ++d;
  return d;
}

// Little-endian packing of a domain value into a block (util.cuh:47-64).
template <typename In>
FSS_SHIM_HD int4 Pack(In val) {
  unsigned __int128 v = static_cast<unsigned __int128>(val);
  return int4{static_cast<int>(v & 0xffffffffu), static_cast<int>((v >> 32) & 0xffffffffu),
              static_cast<int>((v >> 64) & 0xffffffffu), static_cast<int>((v >> 96) & 0xffffffffu)};
}

}  // namespace fss::util
