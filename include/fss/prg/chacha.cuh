// SPDX-License-Identifier: Apache-2.0
// fss/prg/chacha.cuh -- ChaCha-based PRG (reference prg/chacha.cuh:24-128): one block per call, the seed
// duplicated as the 256-bit key, XOR feed-forward, 20 rounds.  Same class name / constructor; runs on the GPU.
#pragma once
#include <fss/prg/aes128_mmo.cuh>

namespace fss::prg {

template <int mul, int rounds = 20>
  requires(rounds == 20 && (mul == 1 || mul == 2 || mul == 4))
class ChaCha {
  const int *nonce_;  // borrowed, like the reference (chacha.cuh:89-93)

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_CHACHA;
  explicit ChaCha(const int *nonce) : nonce_(nonce) {}
  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, nonce_, 8); }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
  }
};
static_assert(Prgable<ChaCha<1>, 1> && Prgable<ChaCha<2>, 2> && Prgable<ChaCha<4>, 4>);

}  // namespace fss::prg
