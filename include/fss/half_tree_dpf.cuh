// SPDX-License-Identifier: Apache-2.0
// fss/half_tree_dpf.cuh -- Half-Tree DPF (reference half_tree_dpf.cuh:39-355): same class template,
// `prg` / `hash_key` members, `Cw` layout (n entries + a separate output CW) and member signatures.
#pragma once
#include <sys/types.h>
#include <fss/b200/runtime.hpp>
#include <fss/group.cuh>
#include <fss/prg.cuh>
#include <fss/util.cuh>

namespace fss {

template <int in_bits, typename Group, typename Prg, typename In = uint, int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 1>)
class HalfTreeDpf {
public:
  Prg prg;
  int4 hash_key;

  // half_tree_dpf.cuh:53-57: last level stores SetLsb(HCW, LCW_0) in s and LCW_1 in extra
  struct alignas(32) Cw {
    int4 s;
    bool extra;
  };
  static_assert(sizeof(Cw) == 32);
  static constexpr int kNumCw = in_bits;

  fssb200_ctx *Context() const {
    return b200::ContextFor(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_HALFTREE, prg, FSSB200_PRED_LT, &hash_key));
  }

  void Gen(Cw cws[], int4 &ocw, const int4 s0s[2], In a, int4 b_buf) const {            // :68
    b200::Check(fssb200_gen_host(Context(), s0s, &a, &b_buf, cws, &ocw, 1), "HalfTreeDpf::Gen");
  }
  int4 Eval(bool b, int4 s0, const Cw cws[], int4 ocw, In x) const {                     // :187
    int4 y;
    b200::Check(fssb200_eval_host(Context(), b, &s0, cws, &ocw, &x, &y, 1), "HalfTreeDpf::Eval");
    return y;
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], int4 ocw, int4 ys[]) const {             // :246
    b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, &ocw, ys, 1, 0, 0), "HalfTreeDpf::EvalAll");
  }

  void GenBatch(const int4 *s0s, const In *alphas, const int4 *betas, Cw *cws, int4 *ocws, size_t nkeys,
                cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_gen(Context(), s0s, alphas, betas, cws, ocws, nkeys, stream), "HalfTreeDpf::GenBatch");
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *ocws, const In *xs, int4 *ys, size_t nkeys,
                 cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_halftree_eval(Context(), b, seeds, cws, ocws, xs, ys, nkeys, stream), "HalfTreeDpf::EvalBatch");
  }
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *ocws, int4 *ys, size_t nkeys,
                    uint64_t leaf_begin = 0, uint64_t leaf_count = 0, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_eval_all(Context(), b, seeds, cws, ocws, ys, nkeys, leaf_begin, leaf_count, stream),
                "HalfTreeDpf::EvalAllBatch");
  }
};

}  // namespace fss
