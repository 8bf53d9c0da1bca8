// SPDX-License-Identifier: Apache-2.0
// fss/hash/blake3.cuh -- keyed single-compression BLAKE3 hash plugin (reference hash/blake3.cuh:24-172): same
// class name, constructor and the two `Hash` overloads; the compression runs on the GPU (fss_b200/csrc/blake3.cuh)
// through `fssb200_hash`.
#pragma once
#include <cstring>
#include <fss/b200/runtime.hpp>
#include <fss/hash.cuh>

namespace fss::hash {

class Blake3 {
  int4 iv_[2];

  fssb200_ctx *Context() const {
    fssb200_params p;
    std::memset(&p, 0, sizeof(p));
    p.scheme = FSSB200_SCHEME_VDPF;
    p.in_bits = 8;
    p.in_bytes = 1;
    p.prg = FSSB200_PRG_CHACHA;
    std::memcpy(p.hash_iv[0], iv_, 32);
    std::memcpy(p.hash_iv[1], iv_, 32);
    return b200::ContextFor(p);
  }
  template <int NIN, int NOUT>
  cuda::std::array<int4, NOUT> Run(int which, const int4 *msg) const {
    int4 *d = nullptr;
    if (cudaMallocAsync(reinterpret_cast<void **>(&d), sizeof(int4) * (NIN + NOUT), nullptr) != cudaSuccess) throw std::bad_alloc();
    cudaMemcpyAsync(d, msg, sizeof(int4) * NIN, cudaMemcpyHostToDevice, nullptr);
    const int rc = fssb200_hash(Context(), which, d, d + NIN, 1, nullptr);
    cuda::std::array<int4, NOUT> out{};
    if (rc == 0) cudaMemcpy(out.data(), d + NIN, sizeof(int4) * NOUT, cudaMemcpyDeviceToHost);
    cudaFreeAsync(d, nullptr);
    b200::Check(rc, "fssb200_hash");
    return out;
  }

public:
  explicit Blake3(cuda::std::span<const int4, 2> iv) : iv_{iv[0], iv[1]} {}  // hash/blake3.cuh:131

  static constexpr int kFssB200Hash = FSSB200_HASH_BLAKE3;
  void FssB200Iv(uint8_t iv32[32]) const { std::memcpy(iv32, iv_, 32); }

  // hash/blake3.cuh:145-149: 64 B -> 32 B
  cuda::std::array<int4, 2> Hash(cuda::std::span<const int4, 4> msg) const { return Run<4, 2>(1, msg.data()); }
  // hash/blake3.cuh:160-171: (a, b) -> 64 B, a's clamp bit separates the two digests
  cuda::std::array<int4, 4> Hash(cuda::std::tuple<int4, const int4> msg) const {
    const int4 in[2] = {cuda::std::get<0>(msg), cuda::std::get<1>(msg)};
    return Run<2, 4>(0, in);
  }
};
static_assert(Hashable<Blake3> && XorHashable<Blake3> && b200::DeviceHash<Blake3>);

}  // namespace fss::hash
