// Host-side probe (no GPU): does the packing loop of host_api.cu run at a core's streaming rate, and do AVX2 / AVX-512 variants change that?
//   g++ -O2 -std=c++17 -pthread tools/probes/pack_simd_probe.cpp -o /tmp/pack_simd_probe && /tmp/pack_simd_probe <threads> [ncw]
// Result on the Xeon (Sapphire / Emerald Rapids) guests of this pool: one thread packs 4.8-5.3 M keys/s = 5.1-5.5 GB/s read + 2.7 GB/s
// written, whatever the instruction set, against 9.0 GB/s for a read-only stream: the loop is bound by the bytes a core can move.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#include <immintrin.h>
static void pack_sse(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw) {
  const size_t in_row = size_t(ncw) * 32u, out_row = size_t(ncw) * 16u + 16u;
  for (size_t k = k0; k < k1; ++k) {
    const uint8_t *r = src + k * in_row; uint8_t *o = dst + k * out_row;
    uint64_t f0 = 0, f1 = 0;
    for (int i = 0; i < ncw && i < 64; ++i) f0 |= uint64_t(r[32 * i + 16] != 0) << i;
    for (int i = 64; i < ncw && i < 128; ++i) f1 |= uint64_t(r[32 * i + 16] != 0) << (i - 64);
    for (int i = 0; i < ncw; ++i) _mm_storeu_si128((__m128i *)(o + 16 * i), _mm_loadu_si128((const __m128i *)(r + 32 * i)));
    _mm_storeu_si128((__m128i *)(o + size_t(ncw) * 16u), _mm_set_epi64x((long long)f1, (long long)f0));
  }
}
__attribute__((target("avx512f,avx512bw,avx512vl,bmi2")))
static void pack_avx512(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw) {
  const size_t in_row = size_t(ncw) * 32u, out_row = size_t(ncw) * 16u + 16u;
  const __m512i idx = _mm512_setr_epi64(0, 1, 4, 5, 0, 1, 4, 5);
  for (size_t k = k0; k < k1; ++k) {
    const uint8_t *r = src + k * in_row; uint8_t *o = dst + k * out_row;
    uint64_t f[2] = {0, 0};
    int i = 0;
    for (; i + 1 < ncw; i += 2) {
      const __m512i v = _mm512_loadu_si512((const void *)(r + 32 * i));
      _mm256_storeu_si256((__m256i *)(o + 16 * i), _mm512_castsi512_si256(_mm512_permutexvar_epi64(idx, v)));
      if (i < 128) {
        const uint64_t m = _cvtmask64_u64(_mm512_test_epi8_mask(v, v));
        f[i >> 6] |= _pext_u64(m, 0x0001000000010000ull) << (i & 63);
      }
    }
    if (i < ncw) {
      const __m256i v = _mm256_loadu_si256((const __m256i *)(r + 32 * i));
      _mm_storeu_si128((__m128i *)(o + 16 * i), _mm256_castsi256_si128(v));
      if (i < 128) f[i >> 6] |= uint64_t(r[32 * i + 16] != 0) << (i & 63);
    }
    _mm_storeu_si128((__m128i *)(o + size_t(ncw) * 16u), _mm_set_epi64x((long long)f[1], (long long)f[0]));
  }
}
__attribute__((target("avx2,bmi2")))
static void pack_avx2(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw) {
  const size_t in_row = size_t(ncw) * 32u, out_row = size_t(ncw) * 16u + 16u;
  const __m256i zero = _mm256_setzero_si256();
  for (size_t k = k0; k < k1; ++k) {
    const uint8_t *r = src + k * in_row; uint8_t *o = dst + k * out_row;
    uint64_t f[2] = {0, 0};
    for (int i = 0; i < ncw; ++i) {
      const __m256i v = _mm256_loadu_si256((const __m256i *)(r + 32 * i));
      _mm_storeu_si128((__m128i *)(o + 16 * i), _mm256_castsi256_si128(v));
      if (i < 128) {
        const uint32_t z = uint32_t(_mm256_movemask_epi8(_mm256_cmpeq_epi8(v, zero)));
        f[i >> 6] |= uint64_t((~z >> 16) & 1u) << (i & 63);
      }
    }
    _mm_storeu_si128((__m128i *)(o + size_t(ncw) * 16u), _mm_set_epi64x((long long)f[1], (long long)f[0]));
  }
}
typedef void (*packfn)(const uint8_t *, uint8_t *, size_t, size_t, int);
int main(int argc, char **argv) {
  const int nt = argc > 1 ? atoi(argv[1]) : 1;
  const int ncw = argc > 2 ? atoi(argv[2]) : 33;
  const size_t nk = (size_t(1) << 20) * 33 / ncw;
  std::vector<uint8_t> src(nk * ncw * 32 + 64);
  for (size_t i = 0; i < src.size(); ++i) src[i] = uint8_t(i * 2654435761u >> 24);
  const size_t out_row = ncw * 16 + 16;
  // correctness
  {
    std::vector<uint8_t> a(1000 * out_row), b(1000 * out_row), c(1000 * out_row);
    pack_sse(src.data(), a.data(), 0, 1000, ncw); pack_avx512(src.data(), b.data(), 0, 1000, ncw); pack_avx2(src.data(), c.data(), 0, 1000, ncw);
    printf("avx512 %s avx2 %s\n", a == b ? "ok" : "MISMATCH", a == c ? "ok" : "MISMATCH");
  }
  const size_t ring = size_t(4) << 14;
  std::vector<std::vector<uint8_t>> dst(nt, std::vector<uint8_t>(ring * out_row / nt + 4096));
  const char *names[3] = {"sse", "avx2", "avx512"};
  packfn fns[3] = {pack_sse, pack_avx2, pack_avx512};
  for (int mode = 0; mode < 3; ++mode) {
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back([&, t] {
      const size_t a = nk * t / nt, b = nk * (t + 1) / nt, blk = 256;
      const size_t cap = dst[t].size() / out_row / blk * blk;
      for (int rep = 0; rep < 3; ++rep)
        for (size_t k = a; k < b; k += blk)
          fns[mode](src.data() + k * ncw * 32, dst[t].data() + ((k - a) % cap) * out_row, 0, std::min(blk, b - k), ncw);
    });
    for (auto &x : th) x.join();
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("%s threads=%d ncw=%d: %.2f GB/s read (%.1f M keys/s)\n", names[mode], nt, ncw, 3.0 * nk * ncw * 32 / s / 1e9, 3.0 * nk / s / 1e6);
  }
}
