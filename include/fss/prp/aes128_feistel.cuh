// SPDX-License-Identifier: Apache-2.0
// fss/prp/aes128_feistel.cuh -- small-domain PRP: 4-round balanced Feistel network whose round function is AES-128 under
// (seed with round number XORed into its first word), cycle-walked into [0, domain) (reference prp/aes128_feistel.cuh:37-160;
// same class name, default construction, `Permu` signature and -- bit for bit -- the same permutation, so keys hashed by
// either implementation place their points in the same buckets).
//
// Host code by design: it drives the cuckoo hashing of VDMPF key generation and the bucket lookup of BatchEval, a few AES
// blocks per point on values the host holds anyway.  Unlike the reference it does not go through OpenSSL (nothing in this
// library links libcrypto): a compact table-free-at-rest AES-128 (S-box derived once from the field inverse) with the four
// round-key schedules of the current seed cached, so a Feistel round costs one block encryption and no key set-up.
#pragma once
#include <cassert>
#include <cstdint>
#include <cstring>
#include <fss/prp.cuh>

namespace fss::prp {

namespace aes_host {

// GF(2^8) multiply modulo x^8 + x^4 + x^3 + x + 1
constexpr uint8_t GfMul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  for (int i = 0; i < 8; ++i) {
    if (b & 1) r ^= a;
    const bool hi = a & 0x80;
    a = static_cast<uint8_t>(a << 1);
    if (hi) a ^= 0x1b;
    b >>= 1;
  }
  return r;
}

struct Tables {
  uint8_t sbox[256];
  uint32_t t0[256];  // column (2s, s, s, 3s) little-endian: MixColumns(SubBytes) of a byte in row 0
  Tables() : sbox{}, t0{} {
    // inverse by exponentiation (a^254), then the affine map of FIPS-197 section 5.1.1
    for (int v = 0; v < 256; ++v) {
      uint8_t inv = 0;
      if (v) {
        uint8_t p = 1, base = static_cast<uint8_t>(v);
        for (int e = 254; e; e >>= 1) {
          if (e & 1) p = GfMul(p, base);
          base = GfMul(base, base);
        }
        inv = p;
      }
      uint8_t s = inv;
      for (int r = 1; r <= 4; ++r) s ^= static_cast<uint8_t>((inv << r) | (inv >> (8 - r)));
      s ^= 0x63;
      sbox[v] = s;
      t0[v] = uint32_t(GfMul(s, 2)) | (uint32_t(s) << 8) | (uint32_t(s) << 16) | (uint32_t(GfMul(s, 3)) << 24);
    }
  }
};
// (built once at first use: 256 field inversions)
inline const Tables &GetTables() {
  static const Tables t;
  return t;
}

inline uint32_t Rotl8(uint32_t w, int bytes) { return bytes ? (w << (8 * bytes)) | (w >> (32 - 8 * bytes)) : w; }

// 11 round keys as little-endian column words (byte i of the state = byte i of the block, column c = bytes 4c..4c+3)
struct Schedule {
  uint32_t rk[44];
  void Expand(const uint8_t key[16]) {
    const Tables &kTables = GetTables();
    std::memcpy(rk, key, 16);
    uint8_t rcon = 1;
    for (int i = 4; i < 44; ++i) {
      uint32_t t = rk[i - 1];
      if ((i & 3) == 0) {
        t = (t >> 8) | (t << 24);  // RotWord on a little-endian column
        t = uint32_t(kTables.sbox[t & 0xff]) | (uint32_t(kTables.sbox[(t >> 8) & 0xff]) << 8) |
            (uint32_t(kTables.sbox[(t >> 16) & 0xff]) << 16) | (uint32_t(kTables.sbox[t >> 24]) << 24);
        t ^= rcon;
        rcon = GfMul(rcon, 2);
      }
      rk[i] = rk[i - 4] ^ t;
    }
  }
  void Encrypt(const uint8_t in[16], uint8_t out[16]) const {
    const Tables &kTables = GetTables();
    uint32_t s[4], n[4];
    std::memcpy(s, in, 16);
    for (int c = 0; c < 4; ++c) s[c] ^= rk[c];
    for (int round = 1; round < 10; ++round) {
      for (int c = 0; c < 4; ++c)  // ShiftRows: row r of output column c comes from column c + r
        n[c] = kTables.t0[s[c] & 0xff] ^ Rotl8(kTables.t0[(s[(c + 1) & 3] >> 8) & 0xff], 1) ^
               Rotl8(kTables.t0[(s[(c + 2) & 3] >> 16) & 0xff], 2) ^ Rotl8(kTables.t0[s[(c + 3) & 3] >> 24], 3) ^ rk[4 * round + c];
      std::memcpy(s, n, 16);
    }
    for (int c = 0; c < 4; ++c)
      n[c] = (uint32_t(kTables.sbox[s[c] & 0xff]) | (uint32_t(kTables.sbox[(s[(c + 1) & 3] >> 8) & 0xff]) << 8) |
              (uint32_t(kTables.sbox[(s[(c + 2) & 3] >> 16) & 0xff]) << 16) | (uint32_t(kTables.sbox[s[(c + 3) & 3] >> 24]) << 24)) ^
             rk[40 + c];
    std::memcpy(out, n, 16);
  }
};

}  // namespace aes_host

class Aes128Feistel {
  static constexpr int kRounds = 4;
  // round-key schedules of the seed seen last (VDMPF uses one seed, sigma, for a whole Gen / BatchEval)
  int4 cached_seed_{};
  bool cached_ = false;
  aes_host::Schedule sched_[kRounds];

  void UseSeed(int4 seed) {
    if (cached_ && std::memcmp(&seed, &cached_seed_, 16) == 0) return;
    for (int r = 0; r < kRounds; ++r) {
      int4 k = seed;
      k.x ^= r;  // the round number tweaks the key's first word
      uint8_t kb[16];
      std::memcpy(kb, &k, 16);
      sched_[r].Expand(kb);
    }
    cached_seed_ = seed;
    cached_ = true;
  }
  static int BitsFor(__uint128_t domain) {  // ceil(log2(domain))
    int b = 0;
    for (__uint128_t v = domain - 1; v; v >>= 1) ++b;
    return b;
  }

public:
  // One pass of the Feistel network over the 2*half-bit value `v` (little-endian 128-bit blocks in and out of AES).
  __uint128_t Network(__uint128_t v, int half) const {
    const __uint128_t mask = (__uint128_t(1) << half) - 1;
    __uint128_t left = (v >> half) & mask, right = v & mask;
    for (int r = 0; r < kRounds; ++r) {
      uint8_t in[16], out[16];
      std::memcpy(in, &right, 16);
      sched_[r].Encrypt(in, out);
      __uint128_t f;
      std::memcpy(&f, out, 16);
      const __uint128_t next_right = left ^ (f & mask);
      left = right;
      right = next_right;
    }
    return (left << half) | right;
  }

  // The permutation of [0, domain) keyed by `seed`; domain <= 1 maps everything to 0 (aes128_feistel.cuh:125-157).
  __uint128_t Permu(int4 seed, __uint128_t x, __uint128_t domain) {
    assert(x < domain || domain <= 1);
    if (domain <= 1) return 0;
    UseSeed(seed);
    const int half = (BitsFor(domain) + 1) / 2;  // <= 64
    __uint128_t v = x;
    do v = Network(v, half);  // cycle walking: 2^(2 half) < 4 domain, so fewer than 4 passes are expected
    while (v >= domain);
    return v;
  }
};
static_assert(Permutable<Aes128Feistel>);

}  // namespace fss::prp
