// SPDX-License-Identifier: Apache-2.0
// fss/prp.cuh -- the small-domain pseudorandom-permutation plugin concept (reference prp.cuh:21-25): a keyed permutation
// of [0, domain) under a 16-byte seed.  The interface IS the contract; the cuckoo hashing of the multi-point scheme
// (fss/cuckoo_hash.cuh, fss/vdmpf.cuh) is written against it.
#pragma once
#include <concepts>
#include <utility>
#include <cuda_runtime.h>

namespace fss::b200 {
// what `prp.Permu(seed, value, domain_size)` returns for a non-const plugin object (PRPs may cache per-seed state)
template <typename P>
using PermuResult = decltype(std::declval<P &>().Permu(std::declval<int4>(), std::declval<__uint128_t>(), std::declval<__uint128_t>()));
}  // namespace fss::b200

// A plugin is Permutable when that call is well-formed and yields a 128-bit unsigned value.
template <typename Prp>
concept Permutable = requires { typename fss::b200::PermuResult<Prp>; } && std::same_as<fss::b200::PermuResult<Prp>, __uint128_t>;
