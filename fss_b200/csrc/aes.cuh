// SPDX-License-Identifier: Apache-2.0
//
// aes.cuh -- AES-128 Matyas-Meyer-Oseas for sm_100a:  out = AES_k(s) ^ s
// (the function of prg/aes128_mmo.cuh:72-93 == prg/aes128_mmo_soft.cuh:209-217).
//
// Design (B200-first, not the reference's Te0 + S-box + per-thread round keys):
//  * State = the four little-endian words of the block, byte k of word j = AES state row k,
//    column j -- no big-endian byte swaps (the reference does LoadBE32/StoreBE32 per word,
//    aes128_mmo_soft.cuh:113-122).
//  * Four T-tables U0..U3 (U_r[x] = column contribution of S[x] in row r, little-endian), each
//    replicated 32x so that lane l only ever touches bank l: every lookup is one conflict-free
//    shared-memory wavefront regardless of the (pseudo-random) indices.  A single table indexed
//    by random bytes averages ~3.15-way conflicts.
//  * The address of a lookup is ONE instruction, and the 16 address computations of a round are split
//    between the two integer issue pipes (ncu of the first version: ALU pipe 91 % busy, FMA pipe 3 %):
//      - pair {U0,U1}: entry x at A0 + x*256 + t*128 + lane*4 with A0 a 64 KiB-aligned shared-window
//        address:  PRMT(word, laneoff, 0x76k4) drops state byte k into byte 1 of (A0 | lane*4)  [ALU pipe];
//      - U2 at A0 + 64K + x*128 + lane*4, U3 at A0 + 96K + x*128 + lane*4:
//        IDP.4A(word, 128 << 8k, laneoff) = byte_k(word)*128 + laneoff                       [FMA pipe];
//    the table select is the LDS immediate offset.  (Generic code needs shift + mask (+ add) per
//    lookup -- 2-3 integer ops; the reference compiles to 68 integer instructions per round, this is
//    8 PRMT + 8 IDP.4A + 8 LOP3.)
//  * Round keys are warp-uniform constant-bank operands (PrgKeys in the kernel parameter block).
//  * Last round takes S[x] out of the byte lane of the table that already has it in place.
//
// Per block: 160 LDS.32 wavefronts/warp and ~252 integer-pipe instructions/thread
// (+44 when the key is selected per thread), vs. the canonical 444 of SURVEY.md section 8d.
#pragma once
#include "common.cuh"

namespace fssb200 {

// ---- table generation (compile time; FIPS-197 section 5.1.1) ---------------------------------------
struct U0Table {
  uint32_t v[256];
};
constexpr uint8_t gf_xtime(uint8_t a) { return static_cast<uint8_t>((a << 1) ^ ((a & 0x80) ? 0x1b : 0)); }
constexpr uint8_t gf_mul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  for (int i = 0; i < 8; ++i) {
    if (b & 1) r ^= a;
    a = gf_xtime(a);
    b >>= 1;
  }
  return r;
}
constexpr uint8_t sbox_of(int x) {
  uint8_t inv = 0;
  if (x != 0) {
    // x^254 = x^-1 in GF(2^8)
    uint8_t p = 1, b = static_cast<uint8_t>(x);
    for (int e = 254; e; e >>= 1) {
      if (e & 1) p = gf_mul(p, b);
      b = gf_mul(b, b);
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
// U0[x] = bytes (2S, S, S, 3S) in rows 0..3 = little-endian byte 0..3
constexpr U0Table make_u0() {
  U0Table t{};
  for (int x = 0; x < 256; ++x) {
    const uint8_t s = sbox_of(x);
    const uint8_t s2 = gf_xtime(s);
    const uint8_t s3 = static_cast<uint8_t>(s2 ^ s);
    t.v[x] = uint32_t(s2) | (uint32_t(s) << 8) | (uint32_t(s) << 16) | (uint32_t(s3) << 24);
  }
  return t;
}
static_assert(sbox_of(0) == 0x63 && sbox_of(1) == 0x7c && sbox_of(0x53) == 0xed, "AES S-box");

constexpr int kAesTblBytes = 131072;        // 64 KiB pair region {U0,U1} + 32 KiB U2 + 32 KiB U3
constexpr uint32_t kOffU0 = 0, kOffU1 = 128, kOffU2 = 65536, kOffU3 = 98304;
// Byte offset (from A0) of entry x of table t in lane `lane`'s bank.
__host__ __device__ constexpr uint32_t aes_tbl_offset(int t, uint32_t x, uint32_t lane) {
  return t == 0 ? kOffU0 + x * 256u + lane * 4u
       : t == 1 ? kOffU1 + x * 256u + lane * 4u
       : t == 2 ? kOffU2 + x * 128u + lane * 4u
                : kOffU3 + x * 128u + lane * 4u;
}
// U_t[x] from U0[x]: rotate left by 8*t bits.
__host__ __device__ constexpr uint32_t aes_tbl_word(uint32_t u0, int t) { return t == 0 ? u0 : ((u0 << (8 * t)) | (u0 >> (32 - 8 * t))); }

// Host side: key expansion (FIPS-197 5.2) into little-endian words, used by ctx_create.
inline void aes128_expand_le(const uint8_t key[16], uint32_t rk[44]) {
  const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
  uint8_t w[176];
  for (int i = 0; i < 16; ++i) w[i] = key[i];
  for (int i = 4; i < 44; ++i) {
    uint8_t t[4] = {w[4 * i - 4], w[4 * i - 3], w[4 * i - 2], w[4 * i - 1]};
    if (i % 4 == 0) {
      const uint8_t u = t[0];
      t[0] = static_cast<uint8_t>(sbox_of(t[1]) ^ rcon[i / 4 - 1]);
      t[1] = sbox_of(t[2]);
      t[2] = sbox_of(t[3]);
      t[3] = sbox_of(u);
    }
    for (int j = 0; j < 4; ++j) w[4 * i + j] = static_cast<uint8_t>(w[4 * (i - 4) + j] ^ t[j]);
  }
  for (int i = 0; i < 44; ++i)
    rk[i] = uint32_t(w[4 * i]) | (uint32_t(w[4 * i + 1]) << 8) | (uint32_t(w[4 * i + 2]) << 16) |
        (uint32_t(w[4 * i + 3]) << 24);
}

#if defined(__CUDACC__)
__device__ __constant__ U0Table c_u0 = make_u0();
#endif

// ---- lookup primitives -------------------------------------------------------------------------------
#if FSS_DEVICE_CODE
FSS_D uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) { return __byte_perm(a, b, sel); }
FSS_D uint32_t dp4a_u32(uint32_t a, uint32_t b, uint32_t c) { return __dp4a(a, b, c); }
template <uint32_t OFF>
FSS_D uint32_t tlookup(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}
#else
inline uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  const uint64_t v = (uint64_t(b) << 32) | a;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t n = (sel >> (4 * i)) & 0xf;
    uint32_t byte = uint32_t(v >> (8 * (n & 7))) & 0xff;
    if (n & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
inline uint32_t dp4a_u32(uint32_t a, uint32_t b, uint32_t c) {
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
}
// host emulation of the shared-memory table image (tests/host_emul only; A0 = 0 there)
uint8_t *host_aes_tables();
template <uint32_t OFF>
inline uint32_t tlookup(uint32_t addr) {
  return *reinterpret_cast<const uint32_t *>(host_aes_tables() + (addr & 0xffffu) + OFF);
}
#endif

struct AesCtx {
  uint32_t laneoff;  // (A0 | lane*4): byte 1 is zero and receives the table index
};

// U_T[byte K of w] for this lane: one address instruction (PRMT for the pair region, IDP.4A for U2 / U3)
// + one LDS.32 whose immediate selects the table.
template <int T, int K>
FSS_HD uint32_t tl(const AesCtx &c, uint32_t w) {
  if (T == 0) return tlookup<kOffU0>(prmt(w, c.laneoff, 0x7604u | (uint32_t(K) << 4)));
  if (T == 1) return tlookup<kOffU1>(prmt(w, c.laneoff, 0x7604u | (uint32_t(K) << 4)));
  if (T == 2) return tlookup<kOffU2>(dp4a_u32(w, 128u << (8 * K), c.laneoff));
  return tlookup<kOffU3>(dp4a_u32(w, 128u << (8 * K), c.laneoff));
}

// Round-key providers.  r = round 0..10, j = word 0..3.
struct KeyFixed {  // compile-time / warp-uniform key index
  const uint32_t *rk;
  FSS_HD uint32_t operator()(int r, int j) const { return rk[4 * r + j]; }
};
struct KeySelect {  // per-thread choice between rk[2p] (mask 0) and rk[2p+1] (mask ~0)
  const uint32_t *rk;
  const uint32_t *rkd;
  uint32_t mask;
  FSS_HD uint32_t operator()(int r, int j) const { return rk[4 * r + j] ^ (mask & rkd[4 * r + j]); }
};

// out = AES_k(s) ^ s.
template <class Key>
FSS_HD blk aes128_mmo(const AesCtx &c, const Key &key, const blk s) {
  uint32_t a0 = s.x ^ key(0, 0), a1 = s.y ^ key(0, 1), a2 = s.z ^ key(0, 2), a3 = s.w ^ key(0, 3);
#pragma unroll
  for (int r = 1; r <= 9; ++r) {
    const uint32_t t0 = tl<0, 0>(c, a0) ^ tl<1, 1>(c, a1) ^ tl<2, 2>(c, a2) ^ tl<3, 3>(c, a3) ^ key(r, 0);
    const uint32_t t1 = tl<0, 0>(c, a1) ^ tl<1, 1>(c, a2) ^ tl<2, 2>(c, a3) ^ tl<3, 3>(c, a0) ^ key(r, 1);
    const uint32_t t2 = tl<0, 0>(c, a2) ^ tl<1, 1>(c, a3) ^ tl<2, 2>(c, a0) ^ tl<3, 3>(c, a1) ^ key(r, 2);
    const uint32_t t3 = tl<0, 0>(c, a3) ^ tl<1, 1>(c, a0) ^ tl<2, 2>(c, a1) ^ tl<3, 3>(c, a2) ^ key(r, 3);
    a0 = t0; a1 = t1; a2 = t2; a3 = t3;
  }
  // Last round (no MixColumns): S[x] sits in byte 0 of U2, byte 1 of U3, byte 2 of U0, byte 3 of U1.
#define FSS_AES_LAST(w0, w1, w2, w3, kj, sj)                                              \
  ((prmt(prmt(tl<2, 0>(c, w0), tl<3, 1>(c, w1), 0x0050u), prmt(tl<0, 2>(c, w2), tl<1, 3>(c, w3), 0x7200u), 0x7610u)) ^ \
   (kj) ^ (sj))
  blk o;
  o.x = FSS_AES_LAST(a0, a1, a2, a3, key(10, 0), s.x);
  o.y = FSS_AES_LAST(a1, a2, a3, a0, key(10, 1), s.y);
  o.z = FSS_AES_LAST(a2, a3, a0, a1, key(10, 2), s.z);
  o.w = FSS_AES_LAST(a3, a0, a1, a2, key(10, 3), s.w);
#undef FSS_AES_LAST
  return o;
}

#if defined(__CUDACC__)
// Fill the lane-replicated tables.  `a0` = 64 KiB-aligned shared-window address of region A.
// Must be followed by __syncthreads().
__device__ inline void aes_tables_init(uint32_t a0) {
  for (uint32_t e = threadIdx.x; e < 256u * 32u; e += blockDim.x) {
    const uint32_t x = e >> 5, lane = e & 31u;
    const uint32_t u0 = c_u0.v[x];
#pragma unroll
    for (int t = 0; t < 4; ++t)
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(a0 + aes_tbl_offset(t, x, lane)), "r"(aes_tbl_word(u0, t)) : "memory");
  }
}
#endif

}  // namespace fssb200
