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

// The cipher for DEVICE callers (the reference's Aes128Soft is `__host__ __device__`; its src/bench_gpu.cu constructs it
// per thread from keys in constant memory and tables it has put in shared memory, and calls Dpf::Gen / Eval with it inside
// its own kernels).  Big-endian column words, so that the caller's Te0 -- (2S, S, S, 3S) packed most significant byte first,
// ComputeTe0 above -- is the contribution of a row-0 byte and its byte rotations are the other rows'.
FSS_SHIM_HD uint32_t Ror32(uint32_t w, int bits) { return (w >> bits) | (w << (32 - bits)); }
FSS_SHIM_HD uint32_t BigEndian(int w) {
  const uint32_t u = static_cast<uint32_t>(w);
  return (u << 24) | ((u & 0xff00u) << 8) | ((u >> 8) & 0xff00u) | (u >> 24);
}
// 44 round-key words of one 16-byte key (FIPS-197 5.2), S-box lookups through the caller's table
FSS_SHIM_HD void ExpandKey(const uint8_t key[16], const uint8_t *sbox, uint32_t rk[44]) {
  for (int i = 0; i < 4; ++i)
    rk[i] = (uint32_t(key[4 * i]) << 24) | (uint32_t(key[4 * i + 1]) << 16) | (uint32_t(key[4 * i + 2]) << 8) | uint32_t(key[4 * i + 3]);
  uint32_t rcon = 1;
  for (int i = 4; i < 44; ++i) {
    uint32_t t = rk[i - 1];
    if ((i & 3) == 0) {
      t = (t << 8) | (t >> 24);
      t = (uint32_t(sbox[t >> 24]) << 24) | (uint32_t(sbox[(t >> 16) & 0xff]) << 16) | (uint32_t(sbox[(t >> 8) & 0xff]) << 8) |
          uint32_t(sbox[t & 0xff]);
      t ^= rcon << 24;
      rcon = XTime(static_cast<uint8_t>(rcon));
    }
    rk[i] = rk[i - 4] ^ t;
  }
}
// AES_k(block) ^ block (Matyas-Meyer-Oseas), block = the 16 bytes of an int4 in memory order
FSS_SHIM_HD int4 EncryptMmo(const uint32_t rk[44], const uint32_t *te0, const uint8_t *sbox, int4 block) {
  uint32_t s[4] = {BigEndian(block.x) ^ rk[0], BigEndian(block.y) ^ rk[1], BigEndian(block.z) ^ rk[2], BigEndian(block.w) ^ rk[3]};
  for (int round = 1; round < 10; ++round) {
    uint32_t n[4];
    for (int c = 0; c < 4; ++c)
      n[c] = te0[s[c] >> 24] ^ Ror32(te0[(s[(c + 1) & 3] >> 16) & 0xff], 8) ^ Ror32(te0[(s[(c + 2) & 3] >> 8) & 0xff], 16) ^
             Ror32(te0[s[(c + 3) & 3] & 0xff], 24) ^ rk[4 * round + c];
    for (int c = 0; c < 4; ++c) s[c] = n[c];
  }
  uint32_t o[4];
  for (int c = 0; c < 4; ++c)
    o[c] = ((uint32_t(sbox[s[c] >> 24]) << 24) | (uint32_t(sbox[(s[(c + 1) & 3] >> 16) & 0xff]) << 16) |
            (uint32_t(sbox[(s[(c + 2) & 3] >> 8) & 0xff]) << 8) | uint32_t(sbox[s[(c + 3) & 3] & 0xff])) ^
           rk[40 + c];
  return int4{int(BigEndian(int(o[0]))) ^ block.x, int(BigEndian(int(o[1]))) ^ block.y, int(BigEndian(int(o[2]))) ^ block.z,
              int(BigEndian(int(o[3]))) ^ block.w};
}
}  // namespace aes_detail

// Host callers: the same function on the same kernels as Aes128Mmo (the tables are not needed: the evaluator carries its own
// lane-replicated ones in shared memory).  Device callers: the software cipher above on the caller's tables, round keys
// expanded once per object.
template <int mul>
class Aes128Soft {
  uint8_t keys_[mul][16];
  const uint32_t *te0_;
  const uint8_t *sbox_;
  uint32_t rk_[mul][44];

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_AES128_MMO;
  FSS_SHIM_HD Aes128Soft(const uint8_t keys[][16], const uint32_t *te0, const uint8_t *sbox) : te0_(te0), sbox_(sbox) {   // aes128_mmo_soft.cuh:196
    for (int i = 0; i < mul; ++i)
      for (int j = 0; j < 16; ++j) keys_[i][j] = keys[i][j];
#if defined(__CUDA_ARCH__)
    for (int i = 0; i < mul; ++i) aes_detail::ExpandKey(keys_[i], sbox_, rk_[i]);
#else
    for (int i = 0; i < mul; ++i)
      for (int j = 0; j < 44; ++j) rk_[i][j] = 0;
#endif
  }
  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, keys_, 16 * mul); }
  FSS_SHIM_HD cuda::std::array<int4, mul> Gen(int4 seed) const {
#if defined(__CUDA_ARCH__)
    cuda::std::array<int4, mul> out{};
    for (int i = 0; i < mul; ++i) out[i] = aes_detail::EncryptMmo(rk_[i], te0_, sbox_, seed);
    return out;
#else
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
#endif
  }
};

static_assert(Prgable<Aes128Mmo<2>, 2> && Prgable<Aes128Mmo<4>, 4> && b200::DevicePrg<Aes128Mmo<1>, 1> &&
              Prgable<Aes128Soft<1>, 1> && b200::DevicePrg<Aes128Soft<2>, 2>);

}  // namespace fss::prg
