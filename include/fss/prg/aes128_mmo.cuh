// SPDX-License-Identifier: Apache-2.0
// fss/prg/aes128_mmo.cuh -- AES-128 Matyas-Meyer-Oseas PRG, out[i] = AES_{key_i}(seed) ^ seed
// (reference prg/aes128_mmo.cuh:27-94, which wraps OpenSSL EVP on the host and traps on the device).
// Same class name, `CreateCtxs` / `FreeCtxs` / constructor / `Gen` surface; the cipher runs on the GPU
// (lane-replicated T-tables, fss_b200/csrc/aes.cuh).  `Aes128MmoRaw` (aes128_mmo_raw.cuh:38-111) and
// `Aes128Soft` (aes128_mmo_soft.cuh:185-218) compute the same function and map to the same kernels.
#pragma once
#include <cstring>
#include <cuda/std/array>
#include <cuda/std/span>
#include <fss/b200/runtime.hpp>
#include <fss/prg.cuh>

// The reference's Aes128Mmo hands out OpenSSL cipher contexts (`cuda::std::array<EVP_CIPHER_CTX *, mul>`,
// prg/aes128_mmo.cuh:49-70) and its users name that type (src/dpf_test.cu:167).  The shim keeps the type NAME -- the same
// opaque declaration OpenSSL makes, so both headers can be included together -- but what CreateCtxs() returns are handles to
// the 16-byte user keys: the cipher runs on the GPU, nothing here links libcrypto.  Only pass Aes128Mmo what its own
// CreateCtxs() made.
typedef struct evp_cipher_ctx_st EVP_CIPHER_CTX;

namespace fss::prg {

namespace b200_detail {
struct AesKey {  // what an `EVP_CIPHER_CTX *` of this shim points at: one user key
  uint8_t key[16];
};
inline EVP_CIPHER_CTX *ToHandle(AesKey *k) { return reinterpret_cast<EVP_CIPHER_CTX *>(k); }
inline AesKey *FromHandle(EVP_CIPHER_CTX *h) { return reinterpret_cast<AesKey *>(h); }
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

  Aes128Mmo(cuda::std::span<EVP_CIPHER_CTX *, mul> ctxs) {                 // prg/aes128_mmo.cuh:40
    for (int i = 0; i < mul; ++i) std::memcpy(keys_[i], b200_detail::FromHandle(ctxs[i])->key, 16);
  }
  // prg/aes128_mmo.cuh:49-64: one context per 16-byte user key
  static cuda::std::array<EVP_CIPHER_CTX *, mul> CreateCtxs(const unsigned char *keys[mul]) {
    cuda::std::array<EVP_CIPHER_CTX *, mul> c{};
    for (int i = 0; i < mul; ++i) {
      auto *k = new b200_detail::AesKey;
      std::memcpy(k->key, keys[i], 16);
      c[i] = b200_detail::ToHandle(k);
    }
    return c;
  }
  static void FreeCtxs(cuda::std::span<EVP_CIPHER_CTX *, mul> ctxs) {     // :66-70
    for (auto *c : ctxs) delete b200_detail::FromHandle(c);
  }

  void FssB200Key(uint8_t key64[64]) const { std::memcpy(key64, keys_, 16 * mul); }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    uint8_t k[64] = {0};
    FssB200Key(k);
    return b200_detail::GenOnDevice<mul>(kFssB200Prg, k, seed);
  }
};

}  // namespace fss::prg

#include <fss/prg/aes128_mmo_soft.cuh>  // Aes128MmoRaw / Aes128Soft: the same function on the same kernels
