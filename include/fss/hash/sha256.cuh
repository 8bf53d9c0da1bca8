// SPDX-License-Identifier: Apache-2.0
// fss/hash/sha256.cuh -- keyed SHA-256 hash plugin (reference hash/sha256.cuh:25-90): same class name, constructor
// and the two `Hash` overloads.  The reference's is host-only (EVP_Digest; `__trap()` on the device, :47-50); here the
// function runs on the GPU (fss_b200/csrc/sha256.cuh) -- inside the VDPF kernels when it is a Vdpf's XorHash / Hash,
// and through `fssb200_hash` for the stand-alone members below.
#pragma once
#include <cstring>
#include <fss/b200/runtime.hpp>
#include <fss/hash.cuh>

namespace fss::hash {

class Sha256 {
  int4 key_;

  fssb200_ctx *Context() const {
    fssb200_params p;
    std::memset(&p, 0, sizeof(p));
    p.scheme = FSSB200_SCHEME_VDPF;
    p.in_bits = 8;
    p.in_bytes = 1;
    p.prg = FSSB200_PRG_CHACHA;
    FssB200Iv(p.hash_iv[0]);
    FssB200Iv(p.hash_iv[1]);
    p.hash = FSSB200_HASH_SHA256 | (FSSB200_HASH_SHA256 << 8);
    return b200::ContextFor(p);
  }
  template <int NIN, int NOUT>
  cuda::std::array<int4, NOUT> Run(int which, const int4 *msg) const {
    b200::DeviceArray<int4> d(NIN + NOUT);
    d.Upload(0, msg, NIN);
    b200::Check(fssb200_hash(Context(), which, d.ptr, d.ptr + NIN, 1, nullptr), "fssb200_hash");
    cuda::std::array<int4, NOUT> out{};
    d.Download(NIN, out.data(), NOUT);
    return out;
  }

public:
  explicit Sha256(int4 key) : key_(key) {}  // hash/sha256.cuh:35

  static constexpr int kFssB200Hash = FSSB200_HASH_SHA256;
  void FssB200Iv(uint8_t iv32[32]) const {  // the key is the first half of the slot
    std::memset(iv32, 0, 32);
    std::memcpy(iv32, &key_, 16);
  }

  // hash/sha256.cuh:44-58: SHA-256(key || 64-byte message)
  cuda::std::array<int4, 2> Hash(cuda::std::span<const int4, 4> msg) const { return Run<4, 2>(1, msg.data()); }
  // hash/sha256.cuh:69-89: SHA-256(key || a lsb=0 || b) || SHA-256(key || a lsb=1 || b)
  cuda::std::array<int4, 4> Hash(cuda::std::tuple<int4, const int4> msg) const {
    const int4 in[2] = {cuda::std::get<0>(msg), cuda::std::get<1>(msg)};
    return Run<2, 4>(0, in);
  }
};
static_assert(Hashable<Sha256> && XorHashable<Sha256> && b200::DeviceHash<Sha256>);

}  // namespace fss::hash
