// SPDX-License-Identifier: Apache-2.0
//
// blake3.cuh -- the BLAKE3 keyed compression behind fss::hash::Blake3 (hash/blake3.cuh), the XorHash / Hash
// plugin of VDPF (vdpf.cuh:55-56), as FSS_HD code (device + tests/host_emul).
//
//   hash(msg 64 B)        = first 32 B of compress(iv, msg, counter 0, block_len 64, flags 0x1B)   hash/blake3.cuh:143-147
//   xor_hash((a, b) 32 B) = compress(iv, {a lsb=0, b, 0, 0}, block_len 32)[0..32) ||
//                           compress(iv, {a lsb=1, b, 0, 0}, block_len 32)[0..32)                    hash/blake3.cuh:158-170
//
// The reference permutes the message array between rounds (Permute, :52-61); here the seven rounds are
// written out with the composed schedule as literal indices, so the message words stay in registers and
// the zero words of the 32-byte blocks fold away at compile time.  Rotations by 16 / 8 are byte
// permutes (PRMT), by 12 / 7 funnel shifts (SHF); three-input adds are IADD3.
#pragma once
#include "common.cuh"

namespace fssb200 {

FSS_HD uint32_t b3_rotr(uint32_t v, int n) {
#if FSS_DEVICE_CODE
  return n == 16 ? __byte_perm(v, v, 0x1032) : (n == 8 ? __byte_perm(v, v, 0x0321) : __funnelshift_r(v, v, n));
#else
  return (v >> n) | (v << (32 - n));
#endif
}

#define FSS_B3_G(a, b, c, d, x, y)                                 \
  a = a + b + (x); d = b3_rotr(d ^ a, 16); c = c + d; b = b3_rotr(b ^ c, 12); \
  a = a + b + (y); d = b3_rotr(d ^ a, 8);  c = c + d; b = b3_rotr(b ^ c, 7);

constexpr uint32_t kB3Flags = 1u | 2u | 8u | 16u;  // CHUNK_START | CHUNK_END | ROOT | KEYED_HASH (hash/blake3.cuh:82-85)

// out[0..8) = first 32 bytes of the compression output (the only part either hash interface uses).
FSS_HD void b3_compress8(const uint32_t h[8], const uint32_t m[16], uint32_t block_len, uint32_t out[8]) {
  uint32_t v0 = h[0], v1 = h[1], v2 = h[2], v3 = h[3], v4 = h[4], v5 = h[5], v6 = h[6], v7 = h[7];
  uint32_t v8 = 0x6A09E667u, v9 = 0xBB67AE85u, v10 = 0x3C6EF372u, v11 = 0xA54FF53Au;  // hash/blake3.cuh:75-80
  uint32_t v12 = 0, v13 = 0, v14 = block_len, v15 = kB3Flags;                         // :104
  /* round 0 */
  FSS_B3_G(v0, v4, v8, v12, m[0], m[1]) FSS_B3_G(v1, v5, v9, v13, m[2], m[3])
  FSS_B3_G(v2, v6, v10, v14, m[4], m[5]) FSS_B3_G(v3, v7, v11, v15, m[6], m[7])
  FSS_B3_G(v0, v5, v10, v15, m[8], m[9]) FSS_B3_G(v1, v6, v11, v12, m[10], m[11])
  FSS_B3_G(v2, v7, v8, v13, m[12], m[13]) FSS_B3_G(v3, v4, v9, v14, m[14], m[15])
  /* round 1 */
  FSS_B3_G(v0, v4, v8, v12, m[2], m[6]) FSS_B3_G(v1, v5, v9, v13, m[3], m[10])
  FSS_B3_G(v2, v6, v10, v14, m[7], m[0]) FSS_B3_G(v3, v7, v11, v15, m[4], m[13])
  FSS_B3_G(v0, v5, v10, v15, m[1], m[11]) FSS_B3_G(v1, v6, v11, v12, m[12], m[5])
  FSS_B3_G(v2, v7, v8, v13, m[9], m[14]) FSS_B3_G(v3, v4, v9, v14, m[15], m[8])
  /* round 2 */
  FSS_B3_G(v0, v4, v8, v12, m[3], m[4]) FSS_B3_G(v1, v5, v9, v13, m[10], m[12])
  FSS_B3_G(v2, v6, v10, v14, m[13], m[2]) FSS_B3_G(v3, v7, v11, v15, m[7], m[14])
  FSS_B3_G(v0, v5, v10, v15, m[6], m[5]) FSS_B3_G(v1, v6, v11, v12, m[9], m[0])
  FSS_B3_G(v2, v7, v8, v13, m[11], m[15]) FSS_B3_G(v3, v4, v9, v14, m[8], m[1])
  /* round 3 */
  FSS_B3_G(v0, v4, v8, v12, m[10], m[7]) FSS_B3_G(v1, v5, v9, v13, m[12], m[9])
  FSS_B3_G(v2, v6, v10, v14, m[14], m[3]) FSS_B3_G(v3, v7, v11, v15, m[13], m[15])
  FSS_B3_G(v0, v5, v10, v15, m[4], m[0]) FSS_B3_G(v1, v6, v11, v12, m[11], m[2])
  FSS_B3_G(v2, v7, v8, v13, m[5], m[8]) FSS_B3_G(v3, v4, v9, v14, m[1], m[6])
  /* round 4 */
  FSS_B3_G(v0, v4, v8, v12, m[12], m[13]) FSS_B3_G(v1, v5, v9, v13, m[9], m[11])
  FSS_B3_G(v2, v6, v10, v14, m[15], m[10]) FSS_B3_G(v3, v7, v11, v15, m[14], m[8])
  FSS_B3_G(v0, v5, v10, v15, m[7], m[2]) FSS_B3_G(v1, v6, v11, v12, m[5], m[3])
  FSS_B3_G(v2, v7, v8, v13, m[0], m[1]) FSS_B3_G(v3, v4, v9, v14, m[6], m[4])
  /* round 5 */
  FSS_B3_G(v0, v4, v8, v12, m[9], m[14]) FSS_B3_G(v1, v5, v9, v13, m[11], m[5])
  FSS_B3_G(v2, v6, v10, v14, m[8], m[12]) FSS_B3_G(v3, v7, v11, v15, m[15], m[1])
  FSS_B3_G(v0, v5, v10, v15, m[13], m[3]) FSS_B3_G(v1, v6, v11, v12, m[0], m[10])
  FSS_B3_G(v2, v7, v8, v13, m[2], m[6]) FSS_B3_G(v3, v4, v9, v14, m[4], m[7])
  /* round 6 */
  FSS_B3_G(v0, v4, v8, v12, m[11], m[15]) FSS_B3_G(v1, v5, v9, v13, m[5], m[0])
  FSS_B3_G(v2, v6, v10, v14, m[1], m[9]) FSS_B3_G(v3, v7, v11, v15, m[8], m[6])
  FSS_B3_G(v0, v5, v10, v15, m[14], m[10]) FSS_B3_G(v1, v6, v11, v12, m[2], m[12])
  FSS_B3_G(v2, v7, v8, v13, m[3], m[4]) FSS_B3_G(v3, v4, v9, v14, m[7], m[13])
  out[0] = v0 ^ v8; out[1] = v1 ^ v9; out[2] = v2 ^ v10; out[3] = v3 ^ v11;           // :117-118
  out[4] = v4 ^ v12; out[5] = v5 ^ v13; out[6] = v6 ^ v14; out[7] = v7 ^ v15;
}
#undef FSS_B3_G

struct blk4 {
  blk b[4];
};

// Hashable::Hash, hash/blake3.cuh:143-147
FSS_HD void b3_hash(const uint32_t iv[8], const blk msg[4], blk out[2]) {
  uint32_t m[16], o[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m[4 * i] = msg[i].x; m[4 * i + 1] = msg[i].y; m[4 * i + 2] = msg[i].z; m[4 * i + 3] = msg[i].w; }
  b3_compress8(iv, m, 64u, o);
  out[0] = make_blk(o[0], o[1], o[2], o[3]);
  out[1] = make_blk(o[4], o[5], o[6], o[7]);
}
// XorHashable::Hash, hash/blake3.cuh:158-170
FSS_HD void b3_xor_hash(const uint32_t iv[8], blk a, blk b, blk out[4]) {
  uint32_t m[16], o[8];
  m[0] = a.x; m[1] = a.y; m[2] = a.z; m[3] = a.w & ~1u;
  m[4] = b.x; m[5] = b.y; m[6] = b.z; m[7] = b.w;
#pragma unroll
  for (int i = 8; i < 16; ++i) m[i] = 0;
  b3_compress8(iv, m, 32u, o);
  out[0] = make_blk(o[0], o[1], o[2], o[3]);
  out[1] = make_blk(o[4], o[5], o[6], o[7]);
  m[3] = a.w | 1u;
  b3_compress8(iv, m, 32u, o);
  out[2] = make_blk(o[0], o[1], o[2], o[3]);
  out[3] = make_blk(o[4], o[5], o[6], o[7]);
}

}  // namespace fssb200
