// SPDX-License-Identifier: Apache-2.0
// VDMPF, deterministic run: prints a digest of both keys, every output share and both proofs of several BatchEval calls.
// Compiled TWICE from this one source: against the reference's unmodified headers (CPU, OpenSSL) by
// oracle/make_golden_vdmpf.py -> tests/golden/vdmpf_v1.txt, and against include/ of this repository (tests/test_vdmpf.py: on
// the CPU over tests/host_emul/fake_backend.cpp, on the B200 over libfssb200.so).  The outputs must be identical line for line.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/hash/blake3.cuh>
#include <fss/hash/sha256.cuh>
#include <fss/prg/aes128_mmo.cuh>
#include <fss/prg/chacha.cuh>
#include <fss/prp/aes128_feistel.cuh>
#include <fss/vdmpf.cuh>

static uint64_t g_state = 0x9e3779b97f4a7c15ULL;
static uint32_t Next() {  // splitmix64, top half
  uint64_t z = (g_state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return uint32_t((z ^ (z >> 31)) >> 32);
}
static int4 Block(bool clamp) { return int4{int(Next()), int(Next()), int(Next()), int(clamp ? Next() & ~1u : Next())}; }

struct Fnv {
  uint64_t h = 0xcbf29ce484222325ULL;
  void Add(const void *p, size_t n) {
    for (size_t i = 0; i < n; ++i) h = (h ^ static_cast<const uint8_t *>(p)[i]) * 0x100000001b3ULL;
  }
};

template <typename V>
static uint64_t KeyDigest(const typename V::Key &k, int bucket_bits) {
  Fnv f;
  f.Add(&k.sigma, 16);
  f.Add(&k.m_rt, 4);
  f.Add(&k.b_size_rt, 4);
  for (int i = 0; i < V::m; ++i) {
    for (int l = 0; l < bucket_bits; ++l) {  // (not the padding bytes of a Cw)
      f.Add(&k.bks[i].cws[l].s, 16);
      const uint8_t tr = k.bks[i].cws[l].tr;
      f.Add(&tr, 1);
    }
    f.Add(k.bks[i].cs.data(), 64);
    f.Add(&k.bks[i].ocw, 16);
    f.Add(&k.bks[i].s0, 16);
  }
  return f.h;
}

template <typename V, typename In, int kInBits, int kBucketBits>
static void Run(const char *name, V &v, int t, int eta) {
  std::vector<In> as;
  while (int(as.size()) < t) {  // distinct points
    const In a = static_cast<In>(Next() & ((uint64_t(1) << kInBits) - 1));
    bool dup = false;
    for (In o : as) dup |= o == a;
    if (!dup) as.push_back(a);
  }
  std::vector<int4> betas;
  for (int i = 0; i < t; ++i) betas.push_back(Block(true));
  auto *k0 = new typename V::Key, *k1 = new typename V::Key;
  std::memset(k0, 0, sizeof(*k0));
  std::memset(k1, 0, sizeof(*k1));
  int tries = 0, ret;
  do {
    const int4 sigma = Block(false);
    cuda::std::array<cuda::std::array<int4, 2>, V::m> s0s;
    for (int i = 0; i < V::m; ++i) s0s[i] = {Block(true), Block(true)};
    ret = v.Gen(*k0, *k1, sigma, cuda::std::span<const cuda::std::array<int4, 2>, V::m>(s0s), std::span<const In>(as),
        std::span<const int4>(betas), t, /*ch_retry=*/tries == 0 ? 3 : 1000);  // (the first try may run out of evictions)
    ++tries;
  } while (ret != 0);
  std::printf("%s t=%d m=%d m_rt=%d b_rt=%d tries=%d k0=%016llx k1=%016llx\n", name, t, V::m, k0->m_rt, k0->b_size_rt, tries,
      (unsigned long long)KeyDigest<V>(*k0, kBucketBits), (unsigned long long)KeyDigest<V>(*k1, kBucketBits));
  // inputs: every point, then random ones (some repeated), in a shuffled order
  std::vector<In> xs(as.begin(), as.end());
  for (int i = 0; i < eta; ++i) xs.push_back(i % 7 == 3 ? xs[size_t(Next()) % xs.size()] : static_cast<In>(Next() & ((uint64_t(1) << kInBits) - 1)));
  for (size_t i = xs.size(); i > 1; --i) std::swap(xs[i - 1], xs[size_t(Next()) % i]);
  for (size_t count : {xs.size(), size_t(1), size_t(0)}) {
    std::vector<int4> y0(count + 1), y1(count + 1);
    cuda::std::array<int4, 4> p0, p1;
    v.BatchEval(false, *k0, std::span<const In>(xs.data(), count), std::span<int4>(y0), p0);
    v.BatchEval(true, *k1, std::span<const In>(xs.data(), count), std::span<int4>(y1), p1);
    Fnv f0, f1, g0, g1;
    f0.Add(y0.data(), 16 * count);
    f1.Add(y1.data(), 16 * count);
    g0.Add(p0.data(), 64);
    g1.Add(p1.data(), 64);
    std::printf("  eta=%zu ys0=%016llx ys1=%016llx pi0=%016llx pi1=%016llx verify=%d\n", count, (unsigned long long)f0.h,
        (unsigned long long)f1.h, (unsigned long long)g0.h, (unsigned long long)g1.h,
        int(V::Verify(cuda::std::span<const int4, 4>(p0), cuda::std::span<const int4, 4>(p1))));
  }
  delete k0;
  delete k1;
}

int main() {
  static int nonce[2] = {0x12345678, int(0x9abcdef0u)};
  int4 iv[2] = {{0x11111111, 0x22222222, 0x33333333, 0x44444444}, {0x55555555, 0x66666666, 0x77777777, int(0x88888888u)}};
  fss::prp::Aes128Feistel prp;
  {
    using G = fss::group::Bytes;
    using P = fss::prg::ChaCha<2>;
    using H = fss::hash::Blake3;
    P prg(nonce);
    H h(cuda::std::span<const int4, 2>(iv, 2));
    fss::Vdmpf<16, 64, 13, G, P, H, H, fss::prp::Aes128Feistel, uint16_t> v{prg, h, h, prp};
    Run<decltype(v), uint16_t, 16, 13>("bytes/chacha/blake3 n=16", v, 64, 300);
    Run<decltype(v), uint16_t, 16, 13>("bytes/chacha/blake3 n=16, a batch of 6000 inputs", v, 50, 6000);
    Run<decltype(v), uint16_t, 16, 13>("bytes/chacha/blake3 n=16 (fewer points than the key is sized for)", v, 31, 40);
  }
  {
    using G = fss::group::Uint<__uint128_t, (static_cast<__uint128_t>(1) << 127)>;
    using P = fss::prg::ChaCha<2>;
    using H = fss::hash::Blake3;
    P prg(nonce);
    H h(cuda::std::span<const int4, 2>(iv, 2));
    fss::Vdmpf<24, 40, 21, G, P, H, H, fss::prp::Aes128Feistel, uint32_t> v{prg, h, h, prp};
    Run<decltype(v), uint32_t, 24, 21>("u127/chacha/blake3 n=24", v, 40, 100);
  }
  {
    using G = fss::group::Uint<uint64_t>;
    using P = fss::prg::Aes128Mmo<2>;
    using H = fss::hash::Sha256;
    unsigned char key0[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}, key1[16] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1};
    const unsigned char *keys[2] = {key0, key1};
    auto ctxs = P::CreateCtxs(keys);
    P prg(ctxs);
    H xh({0x12345678, int(0x9abcdef0u), 0x13572468, int(0x2468ace0u)}), hh({int(0x0fedcba9u), int(0x87654321u), int(0x2468ace0u), 0x13572468});
    fss::Vdmpf<10, 30, 7, G, P, H, H, fss::prp::Aes128Feistel, uint16_t> v{prg, xh, hh, prp};
    Run<decltype(v), uint16_t, 10, 7>("u64/aes128_mmo/sha256 n=10", v, 30, 200);
    P::FreeCtxs(ctxs);
  }
  return 0;
}
