// SPDX-License-Identifier: Apache-2.0
// fss/dcf.cuh -- 2-party DCF (reference dcf.cuh:58-386): same enum, class template, members and `Cw`.
#pragma once
#include <sys/types.h>
#include <fss/b200/runtime.hpp>
#include <fss/group.cuh>
#include <fss/prg.cuh>
#include <fss/util.cuh>

namespace fss {

enum class DcfPred {
  kLt, /**< y = b when x < a */
  kGt, /**< y = b when x > a */
};

template <int in_bits, typename Group, typename Prg, typename In = uint, DcfPred pred = DcfPred::kLt, int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 4>)
class Dcf {
public:
  Prg prg;

  // dcf.cuh:91-96: tl in the clamp bit of s, tr in the clamp bit of v
  struct alignas(32) Cw {
    int4 s;
    int4 v;
  };
  static_assert(sizeof(Cw) == 32);
  static constexpr int kNumCw = in_bits + 1;

  fssb200_ctx *Context() const {
    return b200::ContextFor(b200::MakeParams<in_bits, Group, Prg, In>(
        FSSB200_SCHEME_DCF, prg, pred == DcfPred::kLt ? FSSB200_PRED_LT : FSSB200_PRED_GT));
  }

  void Gen(Cw cws[], const int4 s0s[2], In a, int4 b_buf) const {          // dcf.cuh:108
    b200::Check(fssb200_gen_host(Context(), s0s, &a, &b_buf, cws, nullptr, 1), "Dcf::Gen");
  }
  int4 Eval(bool b, int4 s0, const Cw cws[], In x) const {                  // dcf.cuh:205
    int4 y;
    b200::Check(fssb200_eval_host(Context(), b, &s0, cws, nullptr, &x, &y, 1), "Dcf::Eval");
    return y;
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], int4 ys[]) const {          // dcf.cuh:294
    b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, nullptr, ys, 1, 0, 0), "Dcf::EvalAll");
  }

  void GenBatch(const int4 *s0s, const In *alphas, const int4 *betas, Cw *cws, size_t nkeys,
                cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_gen(Context(), s0s, alphas, betas, cws, nullptr, nkeys, stream), "Dcf::GenBatch");
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys,
                 cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_dcf_eval(Context(), b, seeds, cws, xs, ys, nkeys, stream), "Dcf::EvalBatch");
  }
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, int4 *ys, size_t nkeys, uint64_t leaf_begin = 0,
                    uint64_t leaf_count = 0, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_eval_all(Context(), b, seeds, cws, nullptr, ys, nkeys, leaf_begin, leaf_count, stream),
                "Dcf::EvalAllBatch");
  }
  void EvalBatchHost(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys) const {
    b200::Check(fssb200_eval_host(Context(), b, seeds, cws, nullptr, xs, ys, nkeys), "Dcf::EvalBatchHost");
  }
  // one process, ndev GPUs: per-device arrays, one stream per device, no collective (see Dpf::EvalBatchMulti)
  fssb200_params Params() const {
    return b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DCF, prg,
                                                     pred == DcfPred::kLt ? FSSB200_PRED_LT : FSSB200_PRED_GT);
  }
  void EvalBatchMulti(bool b, int ndev, const int *devices, const int4 *const *seeds, const Cw *const *cws,
                      const In *const *xs, int4 *const *ys, const size_t *nkeys,
                      const cudaStream_t *streams = nullptr) const {
    b200::MultiCall m(Params(), ndev, devices);
    m.Check(fssb200_eval_multi(m.ctxs.data(), ndev, b, reinterpret_cast<const void *const *>(seeds),
                               reinterpret_cast<const void *const *>(cws), nullptr,
                               reinterpret_cast<const void *const *>(xs), reinterpret_cast<void *const *>(ys), nkeys,
                               reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dcf::EvalBatchMulti");
  }
  void SyncMulti(int ndev, const int *devices, const cudaStream_t *streams = nullptr) const {
    b200::MultiCall m(Params(), ndev, devices);
    m.Check(fssb200_multi_sync(m.ctxs.data(), ndev, reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dcf::SyncMulti");
  }
};

}  // namespace fss
