// SPDX-License-Identifier: Apache-2.0
// fss/group.cuh -- the output-group plugin concept (reference group.cuh:39-45).  In this shim a group
// additionally names its device implementation through `kFssB200Group` / `kFssB200Mod{Lo,Hi}`.
#pragma once
#include <concepts>
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

template <typename Group>
concept Groupable = std::is_default_constructible_v<Group> && requires(Group lhs, Group rhs, int4 buf) {
  { lhs + rhs } -> std::same_as<Group>;
  { -lhs } -> std::same_as<Group>;
  { Group::From(buf) } -> std::same_as<Group>;
  { lhs.Into() } -> std::same_as<int4>;
};

namespace fss::b200 {
// A group the B200 evaluator can run: it maps to a (tag, modulus) pair of include/fssb200.h.
template <typename Group>
concept DeviceGroup = Groupable<Group> && requires {
  { Group::kFssB200Group } -> std::convertible_to<int>;
  { Group::kFssB200ModLo } -> std::convertible_to<uint64_t>;
  { Group::kFssB200ModHi } -> std::convertible_to<uint64_t>;
};
}  // namespace fss::b200
