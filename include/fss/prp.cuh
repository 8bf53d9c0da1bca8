// SPDX-License-Identifier: Apache-2.0
// fss/prp.cuh -- the small-domain pseudorandom-permutation plugin concept (reference prp.cuh:21-25): a keyed permutation
// of [0, domain) under a 16-byte seed.  The interface IS the contract; the cuckoo hashing of the multi-point scheme
// (fss/cuckoo_hash.cuh, fss/vdmpf.cuh) is written against it.
#pragma once
#include <concepts>
#include <cuda_runtime.h>

template <typename Prp>
concept Permutable = requires(Prp prp, int4 seed, __uint128_t x, __uint128_t domain) {
  { prp.Permu(seed, x, domain) } -> std::same_as<__uint128_t>;
};
