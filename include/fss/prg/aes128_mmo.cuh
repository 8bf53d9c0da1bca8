// SPDX-License-Identifier: Apache-2.0
// fss/prg/aes128_mmo.cuh -- AES-128 Matyas-Meyer-Oseas PRG, out[i] = AES_{key_i}(seed) ^ seed
// (reference prg/aes128_mmo.cuh:27-94, which wraps OpenSSL EVP on the host and traps on the device).
// Same class name, `CreateCtxs` / `FreeCtxs` / constructor / `Gen` surface; the cipher runs on the GPU
// (lane-replicated T-tables, fss_b200/csrc/aes.cuh).  `Aes128MmoRaw` (aes128_mmo_raw.cuh:38-111) and
// `Aes128Soft` (aes128_mmo_soft.cuh:185-218) compute the same function and map to the same kernels.
#pragma once
#include <array>
#include <cstring>
#include <span>
#include <fss/b200/runtime.hpp>
#include <fss/prg.cuh>

namespace fss::prg {

namespace b200_detail {
struct AesKey {  // stands in for the reference's EVP_CIPHER_CTX*: one user key
  uint8_t key[16];
};
template <int mul>
cuda::std::array<int4, mul> GenOnDevice(int prg_tag, const uint8_t key64[64], int4 seed) {
  fssb200_params p;
  std::memset(&p, 0, sizeof(p));
  p.scheme = mul == 4 ? FSSB200_SCHEME_DCF : (mul == 1 ? FSSB200_SCHEME_HALFTREE : FSSB200_SCHEME_DPF);
  p.in_bits = 8;
  p.in_bytes = 1;
  p.prg = prg_tag;
  std::memcpy(p.prg_key, key64, 64);
  cuda::std::array<int4, mul> r{};
  b200::Check(fssb200_prg_gen_host(b200::ContextFor(p), &seed, r.data(), mul, 1), "fssb200_prg_gen_host");
  return r;
}
}  // namespace b200_detail

template <int mul>
class Aes128Mmo {
  uint8_t keys_[mul][16];

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_AES128_MMO;
  using Ctx = b200_detail::AesKey;

  explicit Aes128Mmo(std::span<Ctx *, mul> ctxs) {
    for (int i = 0; i < mul; ++i) std::memcpy(keys_[i], ctxs[i]->key, 16);
  }
  Aes128Mmo(std::array<Ctx *, mul> &ctxs) : Aes128Mmo(std::span<Ctx *, mul>(ctxs)) {}
  // prg/aes128_mmo.cuh:49-64: one context per 16-byte user key
  static std::array<Ctx *, mul> CreateCtxs(const unsigned char *keys[mul]) {
    std::array<Ctx *, mul> c{};
    for (int i = 0; i < mul; ++i) {
      c[i] = new Ctx;
      std::memcpy(c[i]->key, keys[i], 16);
    }
    return c;
  }
  static void FreeCtxs(std::span<Ctx *, mul> ctxs) {
    for (auto *c : ctxs) delete c;
  }
  static void FreeCtxs(std::array<Ctx *, mul> &ctxs) { FreeCtxs(std::span<Ctx *, mul>(ctxs)); }

  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, keys_, 16 * mul); }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
  }
};

// aes128_mmo_raw.cuh:76-81: constructed from `mul` 16-byte keys
template <int mul>
class Aes128MmoRaw {
  uint8_t keys_[mul][16];

public:
  static constexpr int kFssB200Prg = FSSB200_PRG_AES128_MMO;
  explicit Aes128MmoRaw(const uint8_t keys[][16]) { std::memcpy(keys_, keys, 16 * mul); }
  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, keys_, 16 * mul); }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
  }
};

// aes128_mmo_soft.cuh:197-207: the table pointers of the reference constructor are accepted and ignored
// (the tables live in the evaluator's shared memory).
template <int mul>
class Aes128Soft : public Aes128MmoRaw<mul> {
public:
  Aes128Soft(const uint8_t keys[][16], const uint32_t * /*te0*/, const uint8_t * /*sbox*/) : Aes128MmoRaw<mul>(keys) {}
};

static_assert(Prgable<Aes128Mmo<2>, 2> && Prgable<Aes128Mmo<4>, 4> && b200::DevicePrg<Aes128Mmo<1>, 1>);

}  // namespace fss::prg
