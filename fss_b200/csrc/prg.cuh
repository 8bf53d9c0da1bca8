// SPDX-License-Identifier: Apache-2.0
//
// prg.cuh -- the two PRG plugins of the hot path as device policies.
//
//   Prg<kPrgAes>    : fss::prg::Aes128Mmo<mul>  (prg/aes128_mmo.cuh:72-93): block i = AES_{key_i}(s) ^ s.
//                     Every output block has its OWN key, so a point evaluation only computes the
//                     block(s) of the child it descends into (1 AES/level for DPF, 2 for DCF) -- the
//                     reference always computes all `mul` blocks.
//   Prg<kPrgChaCha> : fss::prg::ChaCha<mul,20>  (prg/chacha.cuh:95-127): one ChaCha block yields all
//                     rows, XOR feed-forward.
//
// Interface (all FSS_HD so tests/host_emul can run them on the CPU):
//   ctx_t                       per-thread PRG context (AES: table base | lane offset)
//   gen<MUL>(k, c, s, out[MUL]) all MUL blocks with the scheme's fixed keys 0..MUL-1
//   gen_child<NB>(k, c, s, bit, out[NB])  the NB blocks of child `bit` (0 left / 1 right):
//                               blocks bit*NB .. bit*NB+NB-1 of a 2*NB-block PRG
#pragma once
#include "aes.cuh"

namespace fssb200 {

template <int PRG>
struct Prg;

template <>
struct Prg<kPrgAes> {
  typedef AesCtx ctx_t;
  static constexpr bool kNeedsTables = true;

  template <int MUL>
  static FSS_HD void gen(const PrgKeys &k, const ctx_t &c, const blk s, blk *out) {
#pragma unroll
    for (int i = 0; i < MUL; ++i) out[i] = aes128_mmo(c, KeyFixed{k.rk[i]}, s);
  }
  template <int NB>
  static FSS_HD void gen_child(const PrgKeys &k, const ctx_t &c, const blk s, uint32_t bit, blk *out) {
    const uint32_t m = 0u - bit;
    if (NB == 1) {
      // DPF / Grotto: keys {0,1}: child bit uses key `bit`
      out[0] = aes128_mmo(c, KeySelect{k.rk[0], k.rkd[0], m}, s);
    } else {
      // DCF: keys {0,1,2,3} = {s_l, v_l, s_r, v_r}: child bit uses keys 2*bit, 2*bit+1, i.e. a
      // selection between rk[0]/rk[2] and rk[1]/rk[3]; rkd2[] holds those differences.
      out[0] = aes128_mmo(c, KeySelect{k.rk[0], k.rkd[0], m}, s);
      out[1] = aes128_mmo(c, KeySelect{k.rk[1], k.rkd[1], m}, s);
    }
  }
  // single block with key index 0 (Half-Tree hash, mul = 1)
  static FSS_HD blk gen1(const PrgKeys &k, const ctx_t &c, const blk s) {
    return aes128_mmo(c, KeyFixed{k.rk[0]}, s);
  }
  // left child of a 2-block PRG alone (Grotto walk, last level)
  static FSS_HD blk gen_left(const PrgKeys &k, const ctx_t &c, const blk s) {
    return aes128_mmo(c, KeyFixed{k.rk[0]}, s);
  }
  // output block I alone (every block has its own key): lets a caller consume blocks one at a time
  static constexpr bool kPerBlock = true;
  template <int I>
  static FSS_HD blk block(const PrgKeys &k, const ctx_t &c, const blk s) {
    return aes128_mmo(c, KeyFixed{k.rk[I]}, s);
  }
  // output block i, i warp-uniform at run time: lets a caller LOOP over the blocks of a node instead of inlining one
  // AES copy per block (7 KB of SASS each; the instruction cache holds 32 KB)
  static FSS_HD blk block_rt(const PrgKeys &k, const ctx_t &c, int i, const blk s) {
    return aes128_mmo(c, KeyFixed{k.rk[i]}, s);
  }
};

// ---- ChaCha (prg/chacha.cuh) ---------------------------------------------------------------------------
FSS_HD uint32_t rotl32(uint32_t v, int n) {
#if FSS_DEVICE_CODE
  return __funnelshift_l(v, v, n);
#else
  return (v << n) | (v >> (32 - n));
#endif
}
#define FSS_QR(a, b, c, d)                                                          \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);             \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

template <int MUL>
FSS_HD void chacha_gen(const PrgKeys &k, const blk s, blk *out) {
  // chacha.cuh:71-83: "expand 16-byte k" for mul <= 2, "expand 32-byte k" for mul = 4
  const uint32_t c0 = 0x61707865u, c1 = MUL <= 2 ? 0x3120646eu : 0x3320646eu,
                 c2 = MUL <= 2 ? 0x79622d36u : 0x79622d32u, c3 = 0x6b206574u;
  uint32_t x0 = c0, x1 = c1, x2 = c2, x3 = c3;
  uint32_t x4 = s.x, x5 = s.y, x6 = s.z, x7 = s.w;
  uint32_t x8 = s.x, x9 = s.y, x10 = s.z, x11 = s.w;          // chacha.cuh:63-65 key = seed || seed
  uint32_t x12 = 0, x13 = 0, x14 = k.nonce[0], x15 = k.nonce[1];  // :106-110
#pragma unroll
  for (int r = 0; r < 20; r += 2) {                            // :47-61
    FSS_QR(x0, x4, x8, x12) FSS_QR(x1, x5, x9, x13) FSS_QR(x2, x6, x10, x14) FSS_QR(x3, x7, x11, x15)
    FSS_QR(x0, x5, x10, x15) FSS_QR(x1, x6, x11, x12) FSS_QR(x2, x7, x8, x13) FSS_QR(x3, x4, x9, x14)
  }
  const blk row1 = make_blk(x4 ^ s.x, x5 ^ s.y, x6 ^ s.z, x7 ^ s.w);
  if (MUL == 1) {                                              // :113-115
    out[0] = row1;
    return;
  }
  out[0] = make_blk(x0 ^ c0, x1 ^ c1, x2 ^ c2, x3 ^ c3);       // :116-118
  out[1] = row1;
  if (MUL == 4) {                                              // :119-124
    out[2] = make_blk(x8 ^ s.x, x9 ^ s.y, x10 ^ s.z, x11 ^ s.w);
    out[3] = make_blk(x12, x13, x14 ^ k.nonce[0], x15 ^ k.nonce[1]);
  }
}

struct NoCtx {};

template <>
struct Prg<kPrgChaCha> {
  typedef NoCtx ctx_t;
  static constexpr bool kNeedsTables = false;

  template <int MUL>
  static FSS_HD void gen(const PrgKeys &k, const ctx_t &, const blk s, blk *out) {
    chacha_gen<MUL>(k, s, out);
  }
  template <int NB>
  static FSS_HD void gen_child(const PrgKeys &k, const ctx_t &, const blk s, uint32_t bit, blk *out) {
    blk all[2 * NB];
    chacha_gen<2 * NB>(k, s, all);
    const uint32_t m = 0u - bit;
#pragma unroll
    for (int i = 0; i < NB; ++i) out[i] = xor_masked(all[i], m, all[i] ^ all[NB + i]);
  }
  static FSS_HD blk gen1(const PrgKeys &k, const ctx_t &, const blk s) {
    blk o[1];
    chacha_gen<1>(k, s, o);
    return o[0];
  }
  static FSS_HD blk gen_left(const PrgKeys &k, const ctx_t &, const blk s) {
    blk o[2];
    chacha_gen<2>(k, s, o);
    return o[0];
  }
  static constexpr bool kPerBlock = false;  // one ChaCha block yields all rows
  template <int I>
  static FSS_HD blk block(const PrgKeys &, const ctx_t &, const blk s) { return s; }  // never called
  static FSS_HD blk block_rt(const PrgKeys &, const ctx_t &, int, const blk s) { return s; }  // never called
};

}  // namespace fssb200
