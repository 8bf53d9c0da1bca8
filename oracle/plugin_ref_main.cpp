// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY.  The reference's own class templates (unmodified headers under /root/reference/include)
// instantiated with the user-defined plugins of tests/cpp/user_plugin.hpp, on the CPU: the golden output the shim's
// plugin path must reproduce bit for bit (tests/golden/plugin_user_v1.txt, written by oracle/make_golden_plugin.py).
#include "../tests/cpp/user_plugin.hpp"

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/half_tree_dpf.cuh>

#include "../tests/cpp/plugin_user_main.inc"

int main() {
  UserKeys keys;
  SectionReferenceSurface(keys);
  return 0;
}
