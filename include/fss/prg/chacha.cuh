// SPDX-License-Identifier: Apache-2.0
// fss/prg/chacha.cuh -- ChaCha-based PRG (reference prg/chacha.cuh:24-128): one block per call, the seed
// duplicated as the 256-bit key, XOR feed-forward, 20 rounds.  Same class name / constructor; runs on the GPU.
#pragma once
#include <fss/prg/aes128_mmo.cuh>

namespace fss::prg {

namespace b200_detail {
// The ChaCha block of this PRG for DEVICE callers (the reference's members are `__host__ __device__`,
// prg/chacha.cuh:95-127): state = constants | seed | seed | {0, 0, nonce0, nonce1}, 20 rounds, XOR feed-forward.
// Host callers go to the GPU through the C ABI like every other host member of the shim.
FSS_SHIM_HD unsigned Rotl(unsigned v, int n) { return (v << n) | (v >> (32 - n)); }
template <int mul>
FSS_SHIM_HD cuda::std::array<int4, mul> ChaChaBlock(int4 seed, int n0, int n1) {
  // "expand 16-byte k" for mul <= 2, "expand 32-byte k" for mul = 4 (chacha.cuh:71-83)
  const unsigned c[4] = {0x61707865u, mul <= 2 ? 0x3120646eu : 0x3320646eu, mul <= 2 ? 0x79622d36u : 0x79622d32u, 0x6b206574u};
  const unsigned s[4] = {unsigned(seed.x), unsigned(seed.y), unsigned(seed.z), unsigned(seed.w)};
  unsigned x[16] = {c[0], c[1], c[2], c[3], s[0], s[1], s[2], s[3], s[0], s[1], s[2], s[3], 0u, 0u, unsigned(n0), unsigned(n1)};
  const auto qr = [&](int a, int b, int cc, int d) {
    x[a] += x[b]; x[d] = Rotl(x[d] ^ x[a], 16);
    x[cc] += x[d]; x[b] = Rotl(x[b] ^ x[cc], 12);
    x[a] += x[b]; x[d] = Rotl(x[d] ^ x[a], 8);
    x[cc] += x[d]; x[b] = Rotl(x[b] ^ x[cc], 7);
  };
  for (int r = 0; r < 10; ++r) {
    qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15);
    qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14);
  }
  const int4 row0 = {int(x[0] ^ c[0]), int(x[1] ^ c[1]), int(x[2] ^ c[2]), int(x[3] ^ c[3])};
  const int4 row1 = {int(x[4] ^ s[0]), int(x[5] ^ s[1]), int(x[6] ^ s[2]), int(x[7] ^ s[3])};
  cuda::std::array<int4, mul> out{};
  if constexpr (mul == 1) {
    out[0] = row1;
  } else {
    out[0] = row0;
    out[1] = row1;
    if constexpr (mul == 4) {
      out[2] = int4{int(x[8] ^ s[0]), int(x[9] ^ s[1]), int(x[10] ^ s[2]), int(x[11] ^ s[3])};
      out[3] = int4{int(x[12]), int(x[13]), int(x[14] ^ unsigned(n0)), int(x[15] ^ unsigned(n1))};
    }
  }
  return out;
}
}  // namespace b200_detail

template <int mul, int rounds = 20>
  requires(rounds == 20 && (mul == 1 || mul == 2 || mul == 4))
class ChaCha {
  const int *nonce_;  // borrowed, like the reference (chacha.cuh:89-93)

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_CHACHA;
  FSS_SHIM_HD explicit ChaCha(const int *nonce) : nonce_(nonce) {}
  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, nonce_, 8); }
  FSS_SHIM_HD cuda::std::array<int4, mul> Gen(int4 seed) const {
#if defined(__CUDA_ARCH__)
    return b200_detail::ChaChaBlock<mul>(seed, nonce_[0], nonce_[1]);  // inside the user's kernel (nonce_: device-visible memory)
#else
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
#endif
  }
};
static_assert(Prgable<ChaCha<1>, 1> && Prgable<ChaCha<2>, 2> && Prgable<ChaCha<4>, 4>);

}  // namespace fss::prg
