// SPDX-License-Identifier: Apache-2.0
//
// group.cuh -- output-group arithmetic fused into the leaf conversion / value accumulation.
// Restates fss::group::Bytes (group/bytes.cuh:19-43) and fss::group::Uint<T,mod>
// (group/uint.cuh:27-88) with the modulus as a run-time, warp-uniform value.
//
//   From : uint.cuh:49-68 (little-endian words; 16-byte T drops the clamp bit by `.w >> 1`; `% mod`)
//   Into : uint.cuh:70-84 (upper words zeroed; 16-byte T re-inserts the clamp bit by `<< 1`)
//   +    : uint.cuh:33-38,   -x : uint.cuh:40-45
//
// All elements are kept canonical (< mod), so `+` is one add and one conditional subtract.
#pragma once
#include "common.cuh"

namespace fssb200 {

struct GroupArgs {
  uint32_t vmask;   // value mask for the <= 32-bit types (0xff, 0xffff, 0xffffffff)
  uint32_t mod[4];  // little-endian modulus words (kGrp*Mod kinds only)
};

template <int G>
struct Grp;

template <>
struct Grp<kGrpBytes> {
  typedef blk V;
  static FSS_HD V zero(const GroupArgs &) { return zero_blk(); }
  static FSS_HD V from(const GroupArgs &, blk b) { return b; }
  static FSS_HD blk into(const GroupArgs &, V v) { return v; }
  static FSS_HD V add(const GroupArgs &, V a, V b) { return a ^ b; }
  static FSS_HD V neg(const GroupArgs &, V a) { return a; }
  // a + (m ? b : 0), m an all-ones/all-zeros mask
  static FSS_HD V add_masked(const GroupArgs &, V a, uint32_t m, V b) { return xor_masked(a, m, b); }
  // (negf ? -a : a)
  static FSS_HD V cneg(const GroupArgs &, V a, uint32_t) { return a; }
};

// Uint<uint8_t|uint16_t|uint32_t>, wraparound
template <>
struct Grp<kGrpU32> {
  typedef uint32_t V;
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &g, blk b) { return b.x & g.vmask; }
  static FSS_HD blk into(const GroupArgs &, V v) { return make_blk(v, 0, 0, 0); }
  static FSS_HD V add(const GroupArgs &g, V a, V b) { return (a + b) & g.vmask; }
  static FSS_HD V neg(const GroupArgs &g, V a) { return (0u - a) & g.vmask; }
  static FSS_HD V add_masked(const GroupArgs &g, V a, uint32_t m, V b) { return (a + (m & b)) & g.vmask; }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};
template <>
struct Grp<kGrpU8> : Grp<kGrpU32> {};
template <>
struct Grp<kGrpU16> : Grp<kGrpU32> {};

template <>
struct Grp<kGrpU64> {
  typedef uint64_t V;
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &, blk b) { return uint64_t(b.x) | (uint64_t(b.y) << 32); }
  static FSS_HD blk into(const GroupArgs &, V v) { return make_blk(uint32_t(v), uint32_t(v >> 32), 0, 0); }
  static FSS_HD V add(const GroupArgs &, V a, V b) { return a + b; }
  static FSS_HD V neg(const GroupArgs &, V a) { return 0 - a; }
  static FSS_HD V add_masked(const GroupArgs &, V a, uint32_t m, V b) { return a + (b & (0 - uint64_t(m & 1u))); }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};

// Uint<__uint128_t, 2^127>: 127-bit wraparound
template <>
struct Grp<kGrpU127> {
  typedef u128 V;
  static FSS_HD u128 mask() { return (u128(1) << 127) - 1; }
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &, blk b) {
    return u128(b.x) | (u128(b.y) << 32) | (u128(b.z) << 64) | (u128(b.w >> 1) << 96);
  }
  static FSS_HD blk into(const GroupArgs &, V v) {
    return make_blk(uint32_t(v), uint32_t(v >> 32), uint32_t(v >> 64), uint32_t(v >> 96) << 1);
  }
  static FSS_HD V add(const GroupArgs &, V a, V b) { return (a + b) & mask(); }
  static FSS_HD V neg(const GroupArgs &, V a) { return (0 - a) & mask(); }
  static FSS_HD V add_masked(const GroupArgs &, V a, uint32_t m, V b) {
    const uint64_t m64 = 0 - uint64_t(m & 1u);
    const u128 mm = (u128(m64) << 64) | m64;
    return (a + (b & mm)) & mask();
  }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};

// Uint<T <= 32 bit, mod>
template <>
struct Grp<kGrpU32Mod> {
  typedef uint32_t V;
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &g, blk b) { return (b.x & g.vmask) % g.mod[0]; }
  static FSS_HD blk into(const GroupArgs &, V v) { return make_blk(v, 0, 0, 0); }
  static FSS_HD V add(const GroupArgs &g, V a, V b) {
    const uint64_t s = uint64_t(a) + b;
    return uint32_t(s >= g.mod[0] ? s - g.mod[0] : s);
  }
  static FSS_HD V neg(const GroupArgs &g, V a) { return a ? g.mod[0] - a : 0; }
  static FSS_HD V add_masked(const GroupArgs &g, V a, uint32_t m, V b) { return add(g, a, m & b); }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};

template <>
struct Grp<kGrpU64Mod> {
  typedef uint64_t V;
  static FSS_HD uint64_t modv(const GroupArgs &g) { return uint64_t(g.mod[0]) | (uint64_t(g.mod[1]) << 32); }
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &g, blk b) { return (uint64_t(b.x) | (uint64_t(b.y) << 32)) % modv(g); }
  static FSS_HD blk into(const GroupArgs &, V v) { return make_blk(uint32_t(v), uint32_t(v >> 32), 0, 0); }
  static FSS_HD V add(const GroupArgs &g, V a, V b) {
    const uint64_t s = a + b, m = modv(g);
    return (s < a || s >= m) ? s - m : s;
  }
  static FSS_HD V neg(const GroupArgs &g, V a) { return a ? modv(g) - a : 0; }
  static FSS_HD V add_masked(const GroupArgs &g, V a, uint32_t m, V b) { return add(g, a, b & (0 - uint64_t(m & 1u))); }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};

// Uint<__uint128_t, mod>, 0 < mod < 2^127 (mod = 2^127 is kGrpU127)
template <>
struct Grp<kGrpU128Mod> {
  typedef u128 V;
  static FSS_HD u128 modv(const GroupArgs &g) {
    return u128(g.mod[0]) | (u128(g.mod[1]) << 32) | (u128(g.mod[2]) << 64) | (u128(g.mod[3]) << 96);
  }
  // v % m by shift-subtract (v < 2^127, m < 2^127): no library call on the device
  static FSS_HD u128 reduce(u128 v, u128 m) {
    if (v < m) return v;
    int sh = 0;
    u128 d = m;
    while ((d << 1) <= v && !(d >> 126)) { d <<= 1; ++sh; }
    for (; sh >= 0; --sh, d >>= 1)
      if (v >= d) v -= d;
    return v;
  }
  static FSS_HD V zero(const GroupArgs &) { return 0; }
  static FSS_HD V from(const GroupArgs &g, blk b) { return reduce(Grp<kGrpU127>::from(g, b), modv(g)); }
  static FSS_HD blk into(const GroupArgs &g, V v) { return Grp<kGrpU127>::into(g, v); }
  static FSS_HD V add(const GroupArgs &g, V a, V b) {
    const u128 s = a + b, m = modv(g);
    return s >= m ? s - m : s;
  }
  static FSS_HD V neg(const GroupArgs &g, V a) { return a ? modv(g) - a : 0; }
  static FSS_HD V add_masked(const GroupArgs &g, V a, uint32_t m, V b) { return (m & 1u) ? add(g, a, b) : a; }
  static FSS_HD V cneg(const GroupArgs &g, V a, uint32_t negf) { return negf ? neg(g, a) : a; }
};

}  // namespace fssb200
