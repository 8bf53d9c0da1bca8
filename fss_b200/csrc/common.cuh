// SPDX-License-Identifier: Apache-2.0
//
// common.cuh -- shared definitions of the sm_100a kernels behind include/fssb200.h.
//
// Everything on the hot path works on 16-byte blocks held as four little-endian 32-bit
// words (the reference's `int4 {x,y,z,w}`, util.cuh:16-38); bit 0 of word 3 is the clamp /
// control bit.
//
// The per-thread bodies are `__host__ __device__` so that tests/host_emul can run the very same
// scheme logic on the CPU against the oracle (no GPU in the build container); everything that
// touches shared memory, PTX or launch geometry is device-only.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fssb200.h"

#if defined(__CUDA_ARCH__)
#define FSS_DEVICE_CODE 1
#else
#define FSS_DEVICE_CODE 0
#endif

#define FSS_HD __host__ __device__ __forceinline__
#define FSS_D __device__ __forceinline__

namespace fssb200 {

typedef unsigned __int128 u128;

struct blk {
  uint32_t x, y, z, w;
};

FSS_HD blk make_blk(uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  blk b;
  b.x = x; b.y = y; b.z = z; b.w = w;
  return b;
}
FSS_HD blk operator^(blk a, blk b) { return make_blk(a.x ^ b.x, a.y ^ b.y, a.z ^ b.z, a.w ^ b.w); }
// a ^ (m & b) with m an all-ones / all-zeros word: one LOP3 per word.
FSS_HD blk xor_masked(blk a, uint32_t m, blk b) {
  return make_blk(a.x ^ (m & b.x), a.y ^ (m & b.y), a.z ^ (m & b.z), a.w ^ (m & b.w));
}
FSS_HD blk clamp(blk a) { a.w &= ~1u; return a; }        // util.cuh:30-34 SetLsb(v, 0)
FSS_HD uint32_t lsb(blk a) { return a.w & 1u; }           // util.cuh:36-38
FSS_HD blk zero_blk() { return make_blk(0, 0, 0, 0); }

// ---- kernel parameter block ---------------------------------------------------------------------
// Passed by value as a __grid_constant__ kernel argument: it lives in the constant bank, so every
// round-key word is a warp-uniform constant operand (no LSU traffic, no registers) -- the
// reference re-expands the keys per thread into local memory (prg/aes128_mmo_soft.cuh:204-207).
struct PrgKeys {
  uint32_t rk[4][44];   // AES-128 round keys of the (up to) 4 user keys, little-endian words
  uint32_t rkd[2][44];  // rk[2p] ^ rk[2p+1]: per-thread key selection is rk[2p] ^ (mask & rkd[p])
  uint32_t nonce[2];    // ChaCha nonce words (prg/chacha.cuh:108-110)
  uint32_t hash_key[4]; // HalfTreeDpf::hash_key (half_tree_dpf.cuh:44)
  uint32_t hash_iv[2][8]; // VDPF: IVs of the XorHash (0) / Hash (1) Blake3 plugins (hash/blake3.cuh:131)
  uint32_t hash_kind[2];  // FSSB200_HASH_* of the two plugins; a SHA-256 plugin's key = hash_iv[i][0..4) (hash/sha256.cuh:35)
};

struct GroupMod {
  uint32_t mod[4];  // modulus words (general-modulus groups), little-endian
};

enum : int { kPrgAes = FSSB200_PRG_AES128_MMO, kPrgChaCha = FSSB200_PRG_CHACHA };

// Group kinds the kernels are instantiated for (value width + modulus class).
enum : int {
  kGrpBytes = 0,    // group::Bytes
  kGrpU8 = 1,       // Uint<uint8_t>,  wraparound
  kGrpU16 = 2,      // Uint<uint16_t>, wraparound
  kGrpU32 = 3,      // Uint<uint32_t>, wraparound
  kGrpU64 = 4,      // Uint<uint64_t>, wraparound
  kGrpU127 = 5,     // Uint<__uint128_t, 2^127>
  kGrpU32Mod = 6,   // Uint<uint8/16/32_t, mod>, general modulus (value mask by width)
  kGrpU64Mod = 7,   // Uint<uint64_t, mod>
  kGrpU128Mod = 8,  // Uint<__uint128_t, mod>, 0 < mod < 2^127
  kNumGrpKinds = 9
};

struct PointArgs {
  const blk *seeds;       // [nkeys]
  const uint8_t *cws;     // key-major, ncw*32 B per key
  const blk *ocws;        // [nkeys] (Half-Tree)
  const uint8_t *xs;      // In[nkeys]
  blk *ys;                // [nkeys]
  // level-major inputs (fssb200_eval_levelmajor)
  const blk *cw_s;        // [n][nkeys]
  const blk *cw_v;        // [n][nkeys]
  const uint32_t *extra;  // [ceil(n/32)][nkeys]
  const blk *out_cw;      // [nkeys]
  uint64_t nkeys;
  int in_bits;
  int in_bytes;
  int party;
  uint32_t vmask;         // value mask for <= 32-bit groups
  const blk *cs;          // VDPF: [nkeys][4] correction seeds (vdpf.cuh:153-157)
  blk *pis;               // VDPF: [nkeys][4] corrected per-point hashes out
  // CUtensorMap of the key-major Cw array as a 2-D byte tensor [nkeys][ncw*32], box 32 keys x 64 B,
  // SWIZZLE_64B (point modes 4 / 5: correction words fetched by the TMA unit); opaque here
  alignas(64) uint8_t tmap[128];
  // point mode 7 (level-major arrays through the TMA unit): tmap = cw_s as a 2-D uint32 tensor [n][nkeys*4], box 4 levels
  // (DCF: 2) x 32 keys; tmap2 = cw_v likewise (DCF only)
  alignas(64) uint8_t tmap2[128];
};

struct GenArgs {
  const blk *s0s;         // [nkeys][2]
  const uint8_t *alphas;  // In[nkeys]
  const blk *betas;       // [nkeys] or nullptr
  uint8_t *cws;           // key-major out
  blk *ocws;              // [nkeys] out (Half-Tree, VDPF)
  blk *cs;                // VDPF: [nkeys][4] out
  int32_t *status;        // VDPF: [nkeys] out, Gen's return value (vdpf.cuh:160)
  uint64_t nkeys;
  int in_bits;
  int in_bytes;
  int pred;
  uint32_t vmask;
  // CUtensorMap of the key-major output array (as PointArgs::tmap): gen kernels with OUT = 1 write through it
  alignas(64) uint8_t tmap[128];
};

struct EvalAllArgs {
  const blk *seeds;
  const uint8_t *cws;
  const blk *ocws;
  void *ys;               // blk[nkeys][leaf_count] (or bytes for Grotto)
  uint64_t nkeys;
  uint64_t leaf_begin;
  uint64_t leaf_count;
  uint64_t ys_stride;     // elements between the outputs of consecutive keys (leaf_count, or 2N-1 inside a parity tree)
  int in_bits;
  int party;
  int unit_bits;          // log2(leaves per CTA work unit)
  int breadth_bits;       // levels expanded breadth-first in shared memory
  int dfs_bits;           // levels each thread expands depth-first
  uint32_t vmask;
};

// Input bit i (MSB first) of an `In` value held as little-endian words (dpf.cuh:196).
struct InVal {
  uint32_t w[4];
};
FSS_HD InVal load_in(const uint8_t *p, int in_bytes) {
  InVal v;
  v.w[0] = v.w[1] = v.w[2] = v.w[3] = 0;
  if (in_bytes == 4) {
    v.w[0] = *reinterpret_cast<const uint32_t *>(p);
  } else if (in_bytes == 8) {
    const uint2 t = *reinterpret_cast<const uint2 *>(p);
    v.w[0] = t.x; v.w[1] = t.y;
  } else if (in_bytes == 16) {
    const uint4 t = *reinterpret_cast<const uint4 *>(p);
    v.w[0] = t.x; v.w[1] = t.y; v.w[2] = t.z; v.w[3] = t.w;
  } else if (in_bytes == 2) {
    v.w[0] = *reinterpret_cast<const uint16_t *>(p);
  } else {
    v.w[0] = *p;
  }
  return v;
}
// bit `pos` (0 = least significant) as 0/1
FSS_HD uint32_t in_bit(const InVal &v, int pos) {
  // pos is warp-uniform; the word select compiles to a few SEL / a uniform branch
  const uint32_t word = pos < 32 ? v.w[0] : (pos < 64 ? v.w[1] : (pos < 96 ? v.w[2] : v.w[3]));
  return (word >> (pos & 31)) & 1u;
}

}  // namespace fssb200
