// SPDX-License-Identifier: Apache-2.0
// fss/prg.cuh -- the PRG plugin concept (reference prg.cuh:20-23).  In this shim a PRG additionally names
// its device implementation (`kFssB200Prg`) and exports its key material (`FssB200Key`).
#pragma once
#include <concepts>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda/std/array>

template <typename Prg, int mul>
concept Prgable = requires(Prg prg, int4 seed) {
  { prg.Gen(seed) } -> std::same_as<cuda::std::array<int4, mul>>;
};

namespace fss::b200 {
template <typename Prg, int mul>
concept DevicePrg = Prgable<Prg, mul> && requires(const Prg prg, uint8_t *key64) {
  { Prg::kFssB200Prg } -> std::convertible_to<int>;
  { prg.FssB200Key(key64) };
};
}  // namespace fss::b200
