// SPDX-License-Identifier: Apache-2.0
// TEST: a user-defined output group and a user-defined PRG that satisfy the reference's plugin concepts
// (group.cuh:39-45 `Groupable`, prg.cuh:20-23 `Prgable`) and nothing else -- no fss_b200 markers, no base classes.
// This one header is compiled twice: against the REFERENCE's include tree (g++, CPU; oracle/make_golden_plugin.py
// turns its output into tests/golden/plugin_user_v1.txt) and against this repository's include/ (nvcc, B200).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda/std/array>

#if defined(__CUDACC__)
#define UP_HD __host__ __device__
#else
#define UP_HD
#endif

// Z_2^64 x Z_2^63: two independent lanes with wraparound addition.  Lane 1 has 63 bits because bit 0 of `.w` is the
// scheme's clamp bit (group.cuh:28-34): From drops it, Into leaves it 0.
struct PairU64 {
  uint64_t lo = 0, hi = 0;
  static constexpr uint64_t kMask63 = 0x7fffffffffffffffull;
  UP_HD PairU64 operator+(PairU64 r) const {
    PairU64 o;
    o.lo = lo + r.lo;
    o.hi = (hi + r.hi) & kMask63;
    return o;
  }
  UP_HD PairU64 operator-() const {
    PairU64 o;
    o.lo = 0 - lo;
    o.hi = (0 - hi) & kMask63;
    return o;
  }
  UP_HD static PairU64 From(int4 b) {
    PairU64 o;
    o.lo = uint64_t(uint32_t(b.x)) | (uint64_t(uint32_t(b.y)) << 32);
    o.hi = uint64_t(uint32_t(b.z)) | (uint64_t(uint32_t(b.w) >> 1) << 32);
    return o;
  }
  UP_HD int4 Into() const {
    return int4{int(uint32_t(lo)), int(uint32_t(lo >> 32)), int(uint32_t(hi)), int(uint32_t(hi >> 32) << 1)};
  }
};

// A toy length-multiplying function (NOT a cryptographic PRG -- the test only needs determinism and diffusion):
// four rounds of multiply / rotate / xor over the four words, keyed, with the output index mixed in.
template <int mul>
struct MixPrg {
  uint32_t key;
  UP_HD static uint32_t Rot(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
  UP_HD cuda::std::array<int4, mul> Gen(int4 seed) const {
    cuda::std::array<int4, mul> out{};
    for (int i = 0; i < mul; ++i) {
      uint32_t a = uint32_t(seed.x) ^ key, b = uint32_t(seed.y) + 0x9e3779b9u * uint32_t(i + 1), c = uint32_t(seed.z) ^ 0x85ebca6bu,
               d = uint32_t(seed.w) + key;
      for (int r = 0; r < 4; ++r) {
        a = Rot(a * 0xcc9e2d51u + b, 13) ^ d;
        b = Rot(b * 0x1b873593u + c, 17) ^ a;
        c = Rot(c * 0xe6546b64u + d, 5) ^ b;
        d = Rot(d * 0x27d4eb2fu + a, 11) ^ c;
      }
      out[i] = int4{int(a), int(b), int(c), int(d)};
    }
    return out;
  }
};
