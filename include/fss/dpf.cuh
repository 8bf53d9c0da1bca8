// SPDX-License-Identifier: Apache-2.0
// fss/dpf.cuh -- 2-party DPF (reference dpf.cuh:61-304): same class template, template parameter list,
// member signatures and `Cw` layout; every member evaluates on the B200 through the C ABI
// (include/fssb200.h).  Added: batched members taking device or host arrays.
//
// Not supported in this shim: calling Gen/Eval from inside a user's own __global__ kernel (the reference's
// members are __host__ __device__, README.md:198-242) -- batch the keys and call EvalBatch instead.
#pragma once
#include <sys/types.h>
#include <fss/b200/runtime.hpp>
#include <fss/group.cuh>
#include <fss/prg.cuh>
#include <fss/util.cuh>

namespace fss {

template <int in_bits, typename Group, typename Prg, typename In = uint, int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 2>)
class Dpf {
public:
  Prg prg;

  // dpf.cuh:76-81: s with tl in its clamp bit, tr as a bool at byte 16
  struct alignas(32) Cw {
    int4 s;
    bool tr;
  };
  static_assert(sizeof(Cw) == 32);
  static constexpr int kNumCw = in_bits + 1;

  fssb200_ctx *Context() const { return b200::ContextFor(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DPF, prg)); }

  // ---- the reference's single-key members (host arrays) ----
  void Gen(Cw cws[], const int4 s0s[2], In a, int4 b_buf) const {          // dpf.cuh:93
    b200::Check(fssb200_gen_host(Context(), s0s, &a, &b_buf, cws, nullptr, 1), "Dpf::Gen");
  }
  int4 Eval(bool b, int4 s0, const Cw cws[], In x) const {                  // dpf.cuh:170
    int4 y;
    b200::Check(fssb200_eval_host(Context(), b, &s0, cws, nullptr, &x, &y, 1), "Dpf::Eval");
    return y;
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], int4 ys[]) const {          // dpf.cuh:232
    b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, nullptr, ys, 1, 0, 0), "Dpf::EvalAll");
  }

  // ---- batched, device pointers, stream ordered ----
  void GenBatch(const int4 *s0s /*[n][2]*/, const In *alphas, const int4 *betas, Cw *cws, size_t nkeys,
                cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_gen(Context(), s0s, alphas, betas, cws, nullptr, nkeys, stream), "Dpf::GenBatch");
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys,
                 cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_dpf_eval(Context(), b, seeds, cws, xs, ys, nkeys, stream), "Dpf::EvalBatch");
  }
  // leaves [leaf_begin, leaf_begin + leaf_count) of every key; ys[k * leaf_count + (x - leaf_begin)]
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, int4 *ys, size_t nkeys, uint64_t leaf_begin = 0,
                    uint64_t leaf_count = 0, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_eval_all(Context(), b, seeds, cws, nullptr, ys, nkeys, leaf_begin, leaf_count, stream),
                "Dpf::EvalAllBatch");
  }
  // ---- batched, host arrays (copies pipelined inside the library) ----
  void EvalBatchHost(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys) const {
    b200::Check(fssb200_eval_host(Context(), b, seeds, cws, nullptr, xs, ys, nkeys), "Dpf::EvalBatchHost");
  }
  // ---- one process, ndev GPUs: per-device arrays, one stream per device, no collective ----
  // devices == nullptr: ordinals 0..ndev-1.  Stream ordered; SyncMulti waits and surfaces each device's error.
  void EvalBatchMulti(bool b, int ndev, const int *devices, const int4 *const *seeds, const Cw *const *cws,
                      const In *const *xs, int4 *const *ys, const size_t *nkeys,
                      const cudaStream_t *streams = nullptr) const {
    b200::MultiCall m(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DPF, prg), ndev, devices);
    m.Check(fssb200_eval_multi(m.ctxs.data(), ndev, b, reinterpret_cast<const void *const *>(seeds),
                               reinterpret_cast<const void *const *>(cws), nullptr,
                               reinterpret_cast<const void *const *>(xs), reinterpret_cast<void *const *>(ys), nkeys,
                               reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dpf::EvalBatchMulti");
  }
  // leaves [leaf_begin[d], +leaf_count[d]) of device d's keys (subtrees sharded: fssb200_leaf_shard)
  void EvalAllBatchMulti(bool b, int ndev, const int *devices, const int4 *const *seeds, const Cw *const *cws,
                         int4 *const *ys, const size_t *nkeys, const uint64_t *leaf_begin, const uint64_t *leaf_count,
                         const cudaStream_t *streams = nullptr) const {
    b200::MultiCall m(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DPF, prg), ndev, devices);
    m.Check(fssb200_eval_all_multi(m.ctxs.data(), ndev, b, reinterpret_cast<const void *const *>(seeds),
                                   reinterpret_cast<const void *const *>(cws), nullptr,
                                   reinterpret_cast<void *const *>(ys), nkeys, leaf_begin, leaf_count,
                                   reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dpf::EvalAllBatchMulti");
  }
  void SyncMulti(int ndev, const int *devices, const cudaStream_t *streams = nullptr) const {
    b200::MultiCall m(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DPF, prg), ndev, devices);
    m.Check(fssb200_multi_sync(m.ctxs.data(), ndev, reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dpf::SyncMulti");
  }
  // host arrays of the whole batch over ndev GPUs (key ranges of fssb200_key_shard)
  void EvalBatchHostMulti(bool b, int ndev, const int *devices, const int4 *seeds, const Cw *cws, const In *xs,
                          int4 *ys, size_t nkeys) const {
    b200::MultiCall m(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DPF, prg), ndev, devices);
    m.Check(fssb200_eval_host_multi(m.ctxs.data(), ndev, b, seeds, cws, nullptr, xs, ys, nkeys, m.rcs.data()),
            "Dpf::EvalBatchHostMulti");
  }
};

}  // namespace fss
