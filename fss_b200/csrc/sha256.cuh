// SPDX-License-Identifier: Apache-2.0
//
// sha256.cuh -- the keyed SHA-256 behind fss::hash::Sha256 (hash/sha256.cuh), the second XorHash / Hash plugin
// of VDPF (vdpf.cuh:55-58), as FSS_HD code (device + tests/host_emul).  The reference's plugin is host-only
// (EVP_Digest, `__trap()` on the device, hash/sha256.cuh:47-50); this one runs inside the VDPF kernels.
//
//   hash(msg 64 B)        = SHA-256(key 16 B || msg)                 : 80 bytes, two blocks        hash/sha256.cuh:44-58
//   xor_hash((a, b) 32 B) = SHA-256(key || a lsb=0 || b) ||
//                           SHA-256(key || a lsb=1 || b)             : 48 bytes each, one block    hash/sha256.cuh:69-89
//
// Bytes of the int4 words are hashed in memory order (little-endian words -> byte swap into the big-endian message
// schedule) and the digest is stored the same way.  The 64 rounds run as 4 x 16 with the schedule as a rolling
// 16-word window at literal indices (registers, ~6 KB of SASS instead of ~30 KB fully unrolled).
#pragma once
#include "common.cuh"

namespace fssb200 {

// FIPS 180-4 4.2.2.  Device code reads the __constant__ copy (warp-uniform index: one constant-bank operand per round);
// host code (tests/host_emul, and the host pass of nvcc) reads the plain array.
#define FSS_SHA256_K_INIT \
  0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u, \
  0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u, \
  0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau, \
  0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u, \
  0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u, \
  0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u, \
  0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u, \
  0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u
#if defined(__CUDACC__)
static __constant__ uint32_t kSha256KDev[64] = {FSS_SHA256_K_INIT};
#endif
static const uint32_t kSha256KHost[64] = {FSS_SHA256_K_INIT};
#undef FSS_SHA256_K_INIT
#if FSS_DEVICE_CODE
#define FSS_SHA256_K(i) kSha256KDev[i]
#else
#define FSS_SHA256_K(i) kSha256KHost[i]
#endif

FSS_HD uint32_t sha_rotr(uint32_t v, int n) {
#if FSS_DEVICE_CODE
  return __funnelshift_r(v, v, n);
#else
  return (v >> n) | (v << (32 - n));
#endif
}
FSS_HD uint32_t sha_bswap(uint32_t v) {
#if FSS_DEVICE_CODE
  return __byte_perm(v, 0, 0x0123);
#else
  return (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
#endif
}

// One compression: h += F(h, w); w (16 big-endian message words) is consumed.
FSS_HD void sha256_compress(uint32_t h[8], uint32_t w[16]) {
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#define FSS_SHA_ROUND(A, B, C, D, E, F, G, H, k, wi)                                                      \
  {                                                                                                       \
    const uint32_t t1 = H + (sha_rotr(E, 6) ^ sha_rotr(E, 11) ^ sha_rotr(E, 25)) + ((E & F) ^ (~E & G)) + (k) + (wi); \
    const uint32_t t2 = (sha_rotr(A, 2) ^ sha_rotr(A, 13) ^ sha_rotr(A, 22)) + ((A & B) ^ (A & C) ^ (B & C)); \
    D += t1;                                                                                              \
    H = t1 + t2;                                                                                          \
  }
#define FSS_SHA_SCHED(i)                                                                                   \
  w[(i) & 15] += (sha_rotr(w[((i) + 1) & 15], 7) ^ sha_rotr(w[((i) + 1) & 15], 18) ^ (w[((i) + 1) & 15] >> 3)) + \
      w[((i) + 9) & 15] + (sha_rotr(w[((i) + 14) & 15], 17) ^ sha_rotr(w[((i) + 14) & 15], 19) ^ (w[((i) + 14) & 15] >> 10));
#pragma unroll 1
  for (int r = 0; r < 64; r += 16) {
    if (r) {
      FSS_SHA_SCHED(0) FSS_SHA_SCHED(1) FSS_SHA_SCHED(2) FSS_SHA_SCHED(3) FSS_SHA_SCHED(4) FSS_SHA_SCHED(5)
      FSS_SHA_SCHED(6) FSS_SHA_SCHED(7) FSS_SHA_SCHED(8) FSS_SHA_SCHED(9) FSS_SHA_SCHED(10) FSS_SHA_SCHED(11)
      FSS_SHA_SCHED(12) FSS_SHA_SCHED(13) FSS_SHA_SCHED(14) FSS_SHA_SCHED(15)
    }
    FSS_SHA_ROUND(a, b, c, d, e, f, g, hh, FSS_SHA256_K(r + 0), w[0])
    FSS_SHA_ROUND(hh, a, b, c, d, e, f, g, FSS_SHA256_K(r + 1), w[1])
    FSS_SHA_ROUND(g, hh, a, b, c, d, e, f, FSS_SHA256_K(r + 2), w[2])
    FSS_SHA_ROUND(f, g, hh, a, b, c, d, e, FSS_SHA256_K(r + 3), w[3])
    FSS_SHA_ROUND(e, f, g, hh, a, b, c, d, FSS_SHA256_K(r + 4), w[4])
    FSS_SHA_ROUND(d, e, f, g, hh, a, b, c, FSS_SHA256_K(r + 5), w[5])
    FSS_SHA_ROUND(c, d, e, f, g, hh, a, b, FSS_SHA256_K(r + 6), w[6])
    FSS_SHA_ROUND(b, c, d, e, f, g, hh, a, FSS_SHA256_K(r + 7), w[7])
    FSS_SHA_ROUND(a, b, c, d, e, f, g, hh, FSS_SHA256_K(r + 8), w[8])
    FSS_SHA_ROUND(hh, a, b, c, d, e, f, g, FSS_SHA256_K(r + 9), w[9])
    FSS_SHA_ROUND(g, hh, a, b, c, d, e, f, FSS_SHA256_K(r + 10), w[10])
    FSS_SHA_ROUND(f, g, hh, a, b, c, d, e, FSS_SHA256_K(r + 11), w[11])
    FSS_SHA_ROUND(e, f, g, hh, a, b, c, d, FSS_SHA256_K(r + 12), w[12])
    FSS_SHA_ROUND(d, e, f, g, hh, a, b, c, FSS_SHA256_K(r + 13), w[13])
    FSS_SHA_ROUND(c, d, e, f, g, hh, a, b, FSS_SHA256_K(r + 14), w[14])
    FSS_SHA_ROUND(b, c, d, e, f, g, hh, a, FSS_SHA256_K(r + 15), w[15])
  }
#undef FSS_SHA_ROUND
#undef FSS_SHA_SCHED
  h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

FSS_HD void sha256_init(uint32_t h[8]) {
  h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
  h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
}
FSS_HD void sha256_put(uint32_t w[16], int at, blk v) {
  w[at] = sha_bswap(v.x); w[at + 1] = sha_bswap(v.y); w[at + 2] = sha_bswap(v.z); w[at + 3] = sha_bswap(v.w);
}
FSS_HD void sha256_digest(const uint32_t h[8], blk out[2]) {
  out[0] = make_blk(sha_bswap(h[0]), sha_bswap(h[1]), sha_bswap(h[2]), sha_bswap(h[3]));
  out[1] = make_blk(sha_bswap(h[4]), sha_bswap(h[5]), sha_bswap(h[6]), sha_bswap(h[7]));
}

// Hashable::Hash, hash/sha256.cuh:44-58: 16 + 64 = 80 bytes -> two blocks
FSS_HD void sha_hash(const uint32_t key[4], const blk msg[4], blk out[2]) {
  uint32_t h[8], w[16];
  sha256_init(h);
  sha256_put(w, 0, make_blk(key[0], key[1], key[2], key[3]));
  sha256_put(w, 4, msg[0]);
  sha256_put(w, 8, msg[1]);
  sha256_put(w, 12, msg[2]);
  sha256_compress(h, w);
  sha256_put(w, 0, msg[3]);
  w[4] = 0x80000000u;
#pragma unroll
  for (int i = 5; i < 15; ++i) w[i] = 0;
  w[15] = 80u * 8u;
  sha256_compress(h, w);
  sha256_digest(h, out);
}
// XorHashable::Hash, hash/sha256.cuh:69-89: 16 + 32 = 48 bytes -> one block per digest
FSS_HD void sha_xor_hash(const uint32_t key[4], blk a, blk b, blk out[4]) {
#pragma unroll
  for (uint32_t sigma = 0; sigma < 2; ++sigma) {  // (unrolled: a rolled loop indexes `out` dynamically = local memory)
    uint32_t h[8], w[16];
    sha256_init(h);
    sha256_put(w, 0, make_blk(key[0], key[1], key[2], key[3]));
    sha256_put(w, 4, make_blk(a.x, a.y, a.z, (a.w & ~1u) | sigma));
    sha256_put(w, 8, b);
    w[12] = 0x80000000u;
    w[13] = 0;
    w[14] = 0;
    w[15] = 48u * 8u;
    sha256_compress(h, w);
    sha256_digest(h, out + 2 * sigma);
  }
}

}  // namespace fssb200
