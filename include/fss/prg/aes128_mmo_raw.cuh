// SPDX-License-Identifier: Apache-2.0
// fss/prg/aes128_mmo_raw.cuh -- `Aes128MmoRaw<mul>` (reference prg/aes128_mmo_raw.cuh:38-111: the AES-NI variant,
// constructed from `mul` 16-byte keys).  Same function as Aes128Mmo (out[i] = AES_{key_i}(seed) ^ seed, pinned bit for
// bit by oracle/make_golden.py), so it maps to the same device kernels.
#pragma once
#include <fss/prg/aes128_mmo.cuh>

namespace fss::prg {

template <int mul>
class Aes128MmoRaw {
  uint8_t keys_[mul][16];

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_AES128_MMO;
  explicit Aes128MmoRaw(const uint8_t keys[][16]) { std::memcpy(keys_, keys, 16 * mul); }   // aes128_mmo_raw.cuh:76-81
  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, keys_, 16 * mul); }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
  }
};

}  // namespace fss::prg
