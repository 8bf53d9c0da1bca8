// SPDX-License-Identifier: Apache-2.0
// The PRP and cuckoo-hash plugins of the multi-point scheme, host only: prints digests of fss::prp::Aes128Feistel::Permu over
// many (seed, domain, x), the cuckoo table sizes ChBucket(t, lambda) over a wide range of t, and the tables Compact::Run builds
// (with enough / too few evictions allowed).  Compiled TWICE from this one source -- against the reference's unmodified headers
// (oracle/make_golden_vdmpf.py -> tests/golden/prp_cuckoo_v1.txt) and against include/ of this repository
// (tests/test_vdmpf.py); the outputs must be identical line for line.
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <span>
#include <vector>
#include <fss/cuckoo_hash.cuh>
#include <fss/prp/aes128_feistel.cuh>

static void Hex(__uint128_t v) { std::printf("%016llx%016llx", (unsigned long long)(v >> 64), (unsigned long long)v); }

int main() {
  fss::prp::Aes128Feistel prp;
  // ---- Permu: domains from 2 to 3 * 2^100, odd and even bit counts, 20 seeds x 50 inputs each ----
  const __uint128_t one = 1;
  const __uint128_t domains[] = {2, 3, 5, 100, 196608, one << 33, (one << 33) + 1, (__uint128_t(3) << 64) + 12345, __uint128_t(3) << 100};
  for (__uint128_t domain : domains) {
    uint64_t acc = 0xcbf29ce484222325ULL;
    __uint128_t first = 0;
    for (int s = 0; s < 20; ++s) {
      const int4 seed{s * 7919 + 1, ~s, s << 20, 0x13572468 ^ s};
      for (int i = 0; i < 50; ++i) {
        const __uint128_t x = (__uint128_t(i) * 0x9e3779b97f4a7c15ULL * 0x1234567ULL + (__uint128_t(i) << 70) + s) % domain;
        const __uint128_t y = prp.Permu(seed, x, domain);
        if (s == 0 && i == 1) first = y;
        acc = (acc ^ uint64_t(y) ^ uint64_t(y >> 64)) * 0x100000001b3ULL;
      }
    }
    std::printf("permu domain=");
    Hex(domain);
    std::printf(" first=");
    Hex(first);
    std::printf(" acc=%016llx\n", (unsigned long long)acc);
  }
  {  // a permutation of a small domain
    const int4 seed{1, 2, 3, 4};
    std::vector<int> image;
    for (int x = 0; x < 1000; ++x) image.push_back(int(prp.Permu(seed, x, 1000)));
    std::vector<int> sorted(image);
    std::sort(sorted.begin(), sorted.end());
    bool bijection = true;
    for (int x = 0; x < 1000; ++x) bijection &= sorted[size_t(x)] == x;
    std::printf("permu 1000: bijection=%d image[0..3]=%d %d %d %d\n", int(bijection), image[0], image[1], image[2], image[3]);
  }
  // ---- ChBucket ----
  for (int lambda : {40, 80, 128}) {
    uint64_t acc = 0xcbf29ce484222325ULL;
    for (int t = 30; t < 2000000; t += (t < 5000 ? 1 : 37)) acc = (acc ^ uint64_t(fss::cuckoo_hash::ChBucket(t, lambda))) * 0x100000001b3ULL;
    std::printf("chbucket lambda=%d m(30)=%d m(1000)=%d m(2^20)=%d acc=%016llx\n", lambda, fss::cuckoo_hash::ChBucket(30, lambda),
        fss::cuckoo_hash::ChBucket(1000, lambda), fss::cuckoo_hash::ChBucket(1 << 20, lambda), (unsigned long long)acc);
  }
  // ---- Compact::Run / PrpHash::Locate ----
  for (int trial = 0; trial < 40; ++trial) {
    std::vector<uint32_t> as;
    for (int i = 0; i < 30 + trial * 17; ++i) as.push_back((uint32_t(i) * 2654435761u + uint32_t(trial) * 97u) & 0xfffffu);
    std::sort(as.begin(), as.end());
    as.erase(std::unique(as.begin(), as.end()), as.end());
    while (as.size() < 30) as.push_back(uint32_t(0xfffff - as.size()));
    const int t = int(as.size()), m = fss::cuckoo_hash::ChBucket(t, 80);
    const __uint128_t n = one << 20;
    const int b_size = int((n * 3 + m - 1) / m);
    const int4 sigma{trial, 2 * trial + 1, 77, ~trial};
    std::vector<std::pair<int, int>> table(size_t(m), std::pair<int, int>{-7, -7});
    for (int retry : {1000, 2, 0}) {
      fss::cuckoo_hash::Compact<fss::prp::Aes128Feistel, uint32_t> compact{prp};
      const int rc = compact.Run(std::span<const uint32_t>(as), m, sigma, n, b_size, retry, std::span<std::pair<int, int>>(table));
      uint64_t h = 0xcbf29ce484222325ULL;
      int placed = 0;
      if (rc == 0)
        for (const auto &e : table) {
          h = (h ^ uint64_t(e.first * 4 + e.second + 5)) * 0x100000001b3ULL;
          placed += e.first >= 0;
        }
      std::printf("compact t=%d m=%d retry=%d rc=%d placed=%d table=%016llx\n", t, m, retry, rc, placed, (unsigned long long)h);
    }
    fss::cuckoo_hash::PrpHash<fss::prp::Aes128Feistel, uint32_t> hasher{prp};
    const auto where = hasher.Locate(sigma, as[3], trial % 3, n, b_size);
    std::printf("  locate(as[3], k=%d) = bucket %d position %d\n", trial % 3, where.first, where.second);
  }
  return 0;
}
