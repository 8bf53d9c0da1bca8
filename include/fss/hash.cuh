// SPDX-License-Identifier: Apache-2.0
// fss/hash.cuh -- the hash plugin concepts (reference hash.cuh:19-30).  In this shim a hash additionally exports
// its 32-byte IV (`FssB200Iv`) so that the evaluator context can run it on the device.
#pragma once
#include <concepts>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda/std/array>
#include <cuda/std/span>
#include <cuda/std/tuple>

template <typename Hash>
concept Hashable = requires(Hash hash, cuda::std::span<const int4, 4> msg) {
  { hash.Hash(msg) } -> std::same_as<cuda::std::array<int4, 2>>;
};

template <typename Hash>
concept XorHashable = requires(Hash hash, cuda::std::tuple<int4, const int4> msg) {
  { hash.Hash(msg) } -> std::same_as<cuda::std::array<int4, 4>>;
};

namespace fss::b200 {
// A hash the evaluator runs on the device: it exports its key material (32-byte slot of fssb200_params::hash_iv) and which
// built-in function it is (FSSB200_HASH_BLAKE3 / FSSB200_HASH_SHA256).
template <typename Hash>
concept DeviceHash = requires(const Hash h, uint8_t *iv32) {
  { h.FssB200Iv(iv32) };
  { Hash::kFssB200Hash } -> std::convertible_to<int>;
};
}  // namespace fss::b200
