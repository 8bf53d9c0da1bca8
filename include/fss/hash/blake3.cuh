// SPDX-License-Identifier: Apache-2.0
// fss/hash/blake3.cuh -- keyed single-compression BLAKE3 hash plugin (reference hash/blake3.cuh:24-172): same
// class name, constructor and the two `Hash` overloads.  Host callers: the compression runs on the GPU
// (fss_b200/csrc/blake3.cuh) through `fssb200_hash`.  Device callers (the reference's members are `__host__ __device__`;
// its src/bench_gpu.cu constructs the plugin and calls Vdpf::Gen / Eval per thread inside its own kernels): the
// compression below, per thread.
#pragma once
#include <cstring>
#include <fss/b200/runtime.hpp>
#include <fss/hash.cuh>
#include <fss/util.cuh>

namespace fss::hash::b200_detail {

// BLAKE3 compression of ONE block under the key words `h`, counter 0, flags CHUNK_START | CHUNK_END | ROOT | KEYED_HASH
// (hash/blake3.cuh:82-85,104); out = the first 8 output words (both plugin interfaces use only those).  The message
// schedule is applied to an index table, not to the words: round r reads m[idx[i]], then idx <- idx o permutation.
FSS_SHIM_HD unsigned Ror(unsigned v, int n) { return (v >> n) | (v << (32 - n)); }
FSS_SHIM_HD void Compress(const unsigned h[8], const unsigned m[16], unsigned block_len, unsigned out[8]) {
  unsigned v[16] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], 0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                    0u, 0u, block_len, 1u | 2u | 8u | 16u};
  unsigned char idx[16] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15};
  const unsigned char perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
  const auto mix = [&](int a, int b, int c, int d, unsigned x, unsigned y) {
    v[a] += v[b] + x; v[d] = Ror(v[d] ^ v[a], 16); v[c] += v[d]; v[b] = Ror(v[b] ^ v[c], 12);
    v[a] += v[b] + y; v[d] = Ror(v[d] ^ v[a], 8);  v[c] += v[d]; v[b] = Ror(v[b] ^ v[c], 7);
  };
  for (int round = 0; round < 7; ++round) {
    for (int c = 0; c < 4; ++c) mix(c, 4 + c, 8 + c, 12 + c, m[idx[2 * c]], m[idx[2 * c + 1]]);                                // columns
    for (int c = 0; c < 4; ++c) mix(c, 4 + (c + 1) % 4, 8 + (c + 2) % 4, 12 + (c + 3) % 4, m[idx[8 + 2 * c]], m[idx[9 + 2 * c]]);  // diagonals
    unsigned char next[16];
    for (int i = 0; i < 16; ++i) next[i] = idx[perm[i]];
    for (int i = 0; i < 16; ++i) idx[i] = next[i];
  }
  for (int i = 0; i < 8; ++i) out[i] = v[i] ^ v[i + 8];
}
FSS_SHIM_HD void Words(int4 b, unsigned *w) {
  w[0] = unsigned(b.x); w[1] = unsigned(b.y); w[2] = unsigned(b.z); w[3] = unsigned(b.w);
}
FSS_SHIM_HD int4 Block(const unsigned *w) { return int4{int(w[0]), int(w[1]), int(w[2]), int(w[3])}; }
// Hashable: 64 B -> 32 B (hash/blake3.cuh:145-149)
FSS_SHIM_HD cuda::std::array<int4, 2> Hash64(const int4 iv[2], const int4 msg[4]) {
  unsigned h[8], m[16], o[8];
  Words(iv[0], h); Words(iv[1], h + 4);
  for (int i = 0; i < 4; ++i) Words(msg[i], m + 4 * i);
  Compress(h, m, 64u, o);
  return {Block(o), Block(o + 4)};
}
// XorHashable: (a, b) -> 64 B; a's clamp bit separates the two 32-byte digests (hash/blake3.cuh:160-171)
FSS_SHIM_HD cuda::std::array<int4, 4> HashPair(const int4 iv[2], int4 a, int4 b) {
  unsigned h[8], m[16] = {}, o[8];
  Words(iv[0], h); Words(iv[1], h + 4);
  Words(a, m); Words(b, m + 4);
  cuda::std::array<int4, 4> out{};
  for (unsigned bit = 0; bit < 2; ++bit) {
    m[3] = (m[3] & ~1u) | bit;
    Compress(h, m, 32u, o);
    out[2 * bit] = Block(o);
    out[2 * bit + 1] = Block(o + 4);
  }
  return out;
}

}  // namespace fss::hash::b200_detail

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
  FSS_SHIM_HD explicit Blake3(cuda::std::span<const int4, 2> iv) : iv_{iv[0], iv[1]} {}  // hash/blake3.cuh:131

  static constexpr int kFssB200Hash = FSSB200_HASH_BLAKE3;
  void FssB200Iv(uint8_t iv32[32]) const { std::memcpy(iv32, iv_, 32); }

  // hash/blake3.cuh:145-149: 64 B -> 32 B
  FSS_SHIM_HD cuda::std::array<int4, 2> Hash(cuda::std::span<const int4, 4> msg) const {
#if defined(__CUDA_ARCH__)
    return b200_detail::Hash64(iv_, msg.data());
#else
    return Run<4, 2>(1, msg.data());
#endif
  }
  // hash/blake3.cuh:160-171: (a, b) -> 64 B, a's clamp bit separates the two digests
  FSS_SHIM_HD cuda::std::array<int4, 4> Hash(cuda::std::tuple<int4, const int4> msg) const {
#if defined(__CUDA_ARCH__)
    return b200_detail::HashPair(iv_, cuda::std::get<0>(msg), cuda::std::get<1>(msg));
#else
    const int4 in[2] = {cuda::std::get<0>(msg), cuda::std::get<1>(msg)};
    return Run<2, 4>(0, in);
#endif
  }
};
static_assert(Hashable<Blake3> && XorHashable<Blake3> && b200::DeviceHash<Blake3>);

}  // namespace fss::hash
