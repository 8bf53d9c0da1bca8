// SPDX-License-Identifier: Apache-2.0
// fss/cuckoo_hash.cuh -- PRP-based compact cuckoo hashing for the multi-point scheme (reference cuckoo_hash.cuh: `ChBucket`
// :83-91, `PrpHash::Locate` :115-134, `Compact::Run` :152-199; Lemma 5 / Remark 1 of ePrint 2024/677): same names, template
// parameter lists, signatures and -- for the same inputs -- the same table, so that a key generated here hashes its points
// into the buckets the reference's BatchEval looks in and vice versa.
//
// Layout of the hashed domain: element x under hash function k is the value x + k n of [0, kappa n); the PRP scatters it
// over that range and the range is cut into buckets of b_size values: bucket = y / b_size, position inside it = y % b_size.
//
// Written for batches: `Locations()` evaluates the kappa candidate places of every element up front (independent PRP
// calls, one pass over the inputs), and the insertion walk then only reads that table -- the reference recomputes the PRP
// on every eviction.  The walk itself (a random hash function per placement from std::mt19937(42), evict the occupant,
// give up after ch_retry evictions) is the reference's, draw for draw.
#pragma once
#include <cassert>
#include <cstddef>
#include <random>
#include <span>
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include <fss/prp.cuh>

namespace fss::cuckoo_hash {

namespace detail {
// log2 of a positive double by the shift-and-square method: integer part by scaling into [1, 2), then one fraction bit
// per squaring (52 of them).
constexpr double BinaryLog(double x) {
  double ip = 0;
  while (x >= 2) {
    x *= 0.5;
    ip += 1;
  }
  while (x < 1) {
    x *= 2;
    ip -= 1;
  }
  double frac = 0, bit = 0.5;
  for (int i = 0; i < 52; ++i) {
    x *= x;
    if (x >= 2) {
      x *= 0.5;
      frac += bit;
    }
    bit *= 0.5;
  }
  return ip + frac;
}
}  // namespace detail

// Buckets for t elements at failure probability 2^-lambda with kappa = 3: m = ceil(e t), e = (lambda + 130 + log2 t) / 123.5
// (Remark 1 of the paper, stated for t >= 30; the reference asserts that bound in debug builds while its own VDMPF sample
// hashes 8 points in a release build -- the formula is evaluated as it stands for any t >= 1).
constexpr int ChBucket(int t, int lambda) {
  assert(t >= 1);
  const double e = (double(lambda) + 130.0 + detail::BinaryLog(double(t))) / 123.5, et = e * double(t);
  const long long whole = static_cast<long long>(et);
  return static_cast<int>(double(whole) < et ? whole + 1 : whole);
}

template <typename Prp, typename In, int kappa = 3>
  requires Permutable<Prp>
struct PrpHash {
  Prp prp;
  // (bucket, position in the bucket) of element x under hash function k; n = size of the input domain.
  std::pair<int, int> Locate(int4 sigma, In x, int k, __uint128_t n, int b_size) {
    const __uint128_t y = prp.Permu(sigma, static_cast<__uint128_t>(x) + n * static_cast<__uint128_t>(k), n * kappa);
    const __uint128_t bs = static_cast<__uint128_t>(b_size);
    return {static_cast<int>(y / bs), static_cast<int>(y % bs)};
  }
  // The kappa candidate places of every element, element-major: out[i * kappa + k].  The calls are independent; large
  // batches (the inputs of a BatchEval) are spread over the host's cores when the translation unit is built with OpenMP,
  // every thread with its own copy of the PRP (a PRP object may keep per-seed state and need not be thread-safe).
  std::vector<std::pair<int, int>> Locations(int4 sigma, std::span<const In> xs, __uint128_t n, int b_size) {
    std::vector<std::pair<int, int>> out(xs.size() * size_t(kappa));
    const long long count = static_cast<long long>(xs.size());
#if defined(_OPENMP)
#pragma omp parallel if (count >= 2048)
    {
      PrpHash local{prp};
#pragma omp for schedule(static)
      for (long long i = 0; i < count; ++i)
        for (int k = 0; k < kappa; ++k) out[size_t(i) * size_t(kappa) + size_t(k)] = local.Locate(sigma, xs[size_t(i)], k, n, b_size);
    }
#else
    for (long long i = 0; i < count; ++i)
      for (int k = 0; k < kappa; ++k) out[size_t(i) * size_t(kappa) + size_t(k)] = Locate(sigma, xs[size_t(i)], k, n, b_size);
#endif
    return out;
  }
};

template <typename Prp, typename In, int kappa = 3>
  requires Permutable<Prp>
struct Compact {
  Prp prp;
  // Fills table[bucket] = (index into `as`, hash function that put it there), (-1, -1) for empty buckets.
  // Returns 0, or 1 when an insertion needed more than ch_retry evictions (the caller draws a new sigma).
  int Run(std::span<const In> as, int m, int4 sigma, __uint128_t n, int b_size, int ch_retry,
      std::span<std::pair<int, int>> table) {
    PrpHash<Prp, In, kappa> hasher{prp};
    const std::vector<std::pair<int, int>> where = hasher.Locations(sigma, as, n, b_size);
    for (int i = 0; i < m; ++i) table[size_t(i)] = {-1, -1};
    std::mt19937 rng(42);
    for (int first = 0; first < int(as.size()); ++first) {
      std::pair<int, int> homeless{first, int(rng() % kappa)};  // (element, hash function to try)
      for (int evictions = 0;; ) {
        const int bucket = where[size_t(homeless.first) * size_t(kappa) + size_t(homeless.second)].first % m;
        std::swap(table[size_t(bucket)], homeless);  // move in; whoever lived there is homeless now
        if (homeless.first < 0) break;               // ... nobody did
        homeless.second = int(rng() % kappa);
        if (++evictions > ch_retry) return 1;
      }
    }
    return 0;
  }
};

}  // namespace fss::cuckoo_hash
