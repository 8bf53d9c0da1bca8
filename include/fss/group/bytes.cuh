// SPDX-License-Identifier: Apache-2.0
// fss/group/bytes.cuh -- 16 bytes with XOR (reference group/bytes.cuh:19-43).
#pragma once
#include <cassert>
#include <fss/group.cuh>
#include <fss/util.cuh>
#include "../../fssb200.h"

namespace fss::group {

struct Bytes {
  int4 val{0, 0, 0, 0};
  static constexpr int kFssB200Group = FSSB200_GROUP_BYTES;
  static constexpr uint64_t kFssB200ModLo = 0, kFssB200ModHi = 0;

  Bytes() = default;
  FSS_SHIM_HD Bytes operator+(Bytes rhs) const { return Bytes(util::Xor(val, rhs.val)); }
  FSS_SHIM_HD Bytes operator-() const { return *this; }
  FSS_SHIM_HD static Bytes From(int4 buf) {
    assert((buf.w & 1) == 0);
    return Bytes(buf);
  }
  FSS_SHIM_HD int4 Into() const { return val; }

private:
  FSS_SHIM_HD explicit Bytes(int4 b) : val(b) {}
};
static_assert(Groupable<Bytes> && fss::b200::DeviceGroup<Bytes>);

}  // namespace fss::group
