// SPDX-License-Identifier: Apache-2.0
// fss/group/uint.cuh -- unsigned integers with addition, optional modulus (reference group/uint.cuh:27-88).
// T from uint8_t to __uint128_t; a 16-byte T needs 0 < mod <= 2^127 because elements are clamped.
#pragma once
#include <cassert>
#include <fss/group.cuh>
#include "../../fssb200.h"

namespace fss::group {

template <typename T, T mod = 0>
  requires((std::is_unsigned_v<T> || std::is_same_v<T, __uint128_t>) && sizeof(T) <= 16 &&
           (sizeof(T) < 16 || (mod > 0 && mod <= static_cast<T>(1) << 127)))
struct Uint {
  T val = 0;
  static constexpr int kFssB200Group = sizeof(T) == 1   ? FSSB200_GROUP_U8
                                       : sizeof(T) == 2 ? FSSB200_GROUP_U16
                                       : sizeof(T) == 4 ? FSSB200_GROUP_U32
                                       : sizeof(T) == 8 ? FSSB200_GROUP_U64
                                                        : FSSB200_GROUP_U128;
  static constexpr uint64_t kFssB200ModLo = static_cast<uint64_t>(static_cast<unsigned __int128>(mod));
  static constexpr uint64_t kFssB200ModHi = static_cast<uint64_t>(static_cast<unsigned __int128>(mod) >> 64);

  Uint() = default;
  FSS_SHIM_HD Uint operator+(Uint rhs) const {
    if constexpr (mod == 0) {
      return Uint(static_cast<T>(val + rhs.val));
    } else {
      const unsigned __int128 s = static_cast<unsigned __int128>(val) + rhs.val;  // both < mod <= 2^127
      return Uint(static_cast<T>(s >= mod ? s - mod : s));
    }
  }
  FSS_SHIM_HD Uint operator-() const {
    if constexpr (mod == 0) return Uint(static_cast<T>(T(0) - val));
    else return Uint(val == 0 ? T(0) : static_cast<T>(mod - val));
  }
  // Little-endian words; a 16-byte T drops the clamp bit with `.w >> 1` (uint.cuh:49-68).
  FSS_SHIM_HD static Uint From(int4 buf) {
    assert((buf.w & 1) == 0);
    const auto u = [](int w) { return static_cast<unsigned __int128>(static_cast<unsigned int>(w)); };
    unsigned __int128 v;
    if constexpr (sizeof(T) < 4) v = u(buf.x) & ((1u << (8 * sizeof(T))) - 1);
    else if constexpr (sizeof(T) == 4) v = u(buf.x);
    else if constexpr (sizeof(T) == 8) v = u(buf.x) | (u(buf.y) << 32);
    else v = u(buf.x) | (u(buf.y) << 32) | (u(buf.z) << 64) | ((u(buf.w) >> 1) << 96);
    if constexpr (mod > 0) v %= mod;
    return Uint(static_cast<T>(v));
  }
  // Upper words zeroed; a 16-byte T re-inserts the clamp bit with `<< 1` (uint.cuh:70-84).
  FSS_SHIM_HD int4 Into() const {
    const unsigned __int128 v = val;
    if constexpr (sizeof(T) <= 4) return int4{static_cast<int>(static_cast<unsigned int>(v)), 0, 0, 0};
    else if constexpr (sizeof(T) == 8)
      return int4{static_cast<int>(static_cast<unsigned int>(v)), static_cast<int>(static_cast<unsigned int>(v >> 32)), 0, 0};
    else
      return int4{static_cast<int>(static_cast<unsigned int>(v)), static_cast<int>(static_cast<unsigned int>(v >> 32)),
                  static_cast<int>(static_cast<unsigned int>(v >> 64)),
                  static_cast<int>(static_cast<unsigned int>(v >> 96) << 1)};
  }

private:
  FSS_SHIM_HD explicit Uint(T v) : val(v) {}
};
static_assert(Groupable<Uint<uint8_t>> && Groupable<Uint<uint64_t>>);
#if !defined(__CUDACC__)  // nvcc's host pass prints 2^127 as a decimal literal gcc then warns about
static_assert(Groupable<Uint<__uint128_t, static_cast<__uint128_t>(1) << 127>>);
#endif

}  // namespace fss::group
