// SPDX-License-Identifier: Apache-2.0
// fss/prg/aes128_mmo_soft.cuh -- `Aes128Soft<mul>` (reference prg/aes128_mmo_soft.cuh:185-218: software T-table AES that
// the user feeds with a Te0 table and an S-box it has filled through `aes_detail::InitTe0` / `InitSbox`).  Same function
// as Aes128Mmo, same device kernels.  The constructor keeps the reference's argument list; the two tables are accepted
// and ignored (the evaluator carries its own lane-replicated tables in shared memory, fss_b200/csrc/aes.cuh).
// `aes_detail::{Sbox, ComputeTe0, InitTe0, InitSbox}` exist so that code written against the reference compiles and gets
// the values it expects (FIPS-197 S-box; Te0[x] = (2S, S, S, 3S) big-endian packed, aes128_mmo_soft.cuh:55-59); they are
// computed from the field arithmetic here, not from a table.
#pragma once
#include <fss/prg/aes128_mmo_raw.cuh>

namespace fss::prg {

namespace aes_detail {
FSS_SHIM_HD constexpr uint8_t XTime(uint8_t a) { return static_cast<uint8_t>((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); }
FSS_SHIM_HD constexpr uint8_t GfMul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  for (int i = 0; i < 8; ++i) {
    if (b & 1) r ^= a;
    a = XTime(a);
    b >>= 1;
  }
  return r;
}
// FIPS-197 5.1.1: multiplicative inverse in GF(2^8) (x^254), then the affine map
FSS_SHIM_HD constexpr uint8_t Sbox(uint8_t idx) {
  uint8_t inv = 0;
  if (idx) {
    uint8_t p = 1, b = idx;
    for (int e = 254; e; e >>= 1) {
      if (e & 1) p = GfMul(p, b);
      b = GfMul(b, b);
    }
    inv = p;
  }
  uint8_t s = inv, r = inv;
  for (int k = 0; k < 4; ++k) {
    r = static_cast<uint8_t>((r << 1) | (r >> 7));
    s ^= r;
  }
  return static_cast<uint8_t>(s ^ 0x63);
}
FSS_SHIM_HD constexpr uint32_t ComputeTe0(uint8_t idx) {
  const uint8_t s = Sbox(idx), s2 = XTime(s), s3 = static_cast<uint8_t>(s2 ^ s);
  return (uint32_t(s2) << 24) | (uint32_t(s) << 16) | (uint32_t(s) << 8) | uint32_t(s3);
}
FSS_SHIM_HD void InitTe0(uint32_t *dst) {
  for (int i = 0; i < 256; ++i) dst[i] = ComputeTe0(static_cast<uint8_t>(i));
}
FSS_SHIM_HD void InitSbox(uint8_t *dst) {
  for (int i = 0; i < 256; ++i) dst[i] = Sbox(static_cast<uint8_t>(i));
}
static_assert(Sbox(0) == 0x63 && Sbox(1) == 0x7c && Sbox(0x53) == 0xed && ComputeTe0(0) == 0xc66363a5u, "AES tables");
}  // namespace aes_detail

template <int mul>
class Aes128Soft : public Aes128MmoRaw<mul> {
public:
  Aes128Soft(const uint8_t keys[][16], const uint32_t * /*te0*/, const uint8_t * /*sbox*/) : Aes128MmoRaw<mul>(keys) {}
};

static_assert(Prgable<Aes128Mmo<2>, 2> && Prgable<Aes128Mmo<4>, 4> && b200::DevicePrg<Aes128Mmo<1>, 1> &&
              Prgable<Aes128Soft<1>, 1> && b200::DevicePrg<Aes128Soft<2>, 2>);

}  // namespace fss::prg
