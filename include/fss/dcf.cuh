// SPDX-License-Identifier: Apache-2.0
// fss/dcf.cuh -- 2-party DCF (reference dcf.cuh:58-386): same enum, class template, members and `Cw`.
// Host members with the built-in plugins run on the B200 through the C ABI; device code and user-defined plugins go
// through the plugin-generic templates of fss/b200/generic.cuh (see fss/dpf.cuh for the three cases).
#pragma once
#include <sys/types.h>
#include <fss/b200/generic.cuh>
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
           Groupable<Group> && Prgable<Prg, 4>)
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
  static constexpr bool kPrebuilt = b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 4>;

  fssb200_params Params() const
    requires kPrebuilt
  {
    return b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_DCF, prg,
                                                     pred == DcfPred::kLt ? FSSB200_PRED_LT : FSSB200_PRED_GT);
  }
  fssb200_ctx *Context() const
    requires kPrebuilt
  {
    return b200::ContextFor(Params());
  }

  FSS_SHIM_HD void Gen(Cw cws[], const int4 s0s[2], In a, int4 b_buf) const {          // dcf.cuh:108
    KeepGenericKernels();
#if defined(__CUDA_ARCH__)
    b200::generic::DcfGen<in_bits, Group, In>(const_cast<Prg &>(prg), pred == DcfPred::kLt, cws, s0s, a, b_buf);
#else
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_gen_host(Context(), s0s, &a, &b_buf, cws, nullptr, 1), "Dcf::Gen");
    } else {
      b200::DeviceArray<int4> ds(3);
      b200::DeviceArray<In> da(1);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, s0s, 2);
      ds.Upload(2, &b_buf, 1);
      da.Upload(0, &a, 1);
      GenBatch(ds.ptr, da.ptr, ds.ptr + 2, dc.ptr, 1);
      dc.Download(0, cws, kNumCw);
    }
#endif
  }
  FSS_SHIM_HD int4 Eval(bool b, int4 s0, const Cw cws[], In x) const {                  // dcf.cuh:205
    KeepGenericKernels();
#if defined(__CUDA_ARCH__)
    return b200::generic::DcfEval<in_bits, Group, In>(const_cast<Prg &>(prg), b, s0, cws, x);
#else
    int4 y;
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_host(Context(), b, &s0, cws, nullptr, &x, &y, 1), "Dcf::Eval");
    } else {
      b200::DeviceArray<int4> ds(2);
      b200::DeviceArray<In> dx(1);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, &s0, 1);
      dx.Upload(0, &x, 1);
      dc.Upload(0, cws, kNumCw);
      EvalBatch(b, ds.ptr, dc.ptr, dx.ptr, ds.ptr + 1, 1);
      ds.Download(1, &y, 1);
    }
    return y;
#endif
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], int4 ys[]) const {          // dcf.cuh:294
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, nullptr, ys, 1, 0, 0), "Dcf::EvalAll");
    } else {
      static_assert(in_bits <= 40, "EvalAll: 2^in_bits leaves");
      const size_t n = size_t(1) << in_bits;
      b200::DeviceArray<int4> ds(1), dy(n);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, &s0, 1);
      dc.Upload(0, cws, kNumCw);
      EvalAllBatch(b, ds.ptr, dc.ptr, dy.ptr, 1);
      dy.Download(0, ys, n);
    }
  }

  void GenBatch(const int4 *s0s, const In *alphas, const int4 *betas, Cw *cws, size_t nkeys,
                cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_gen(Context(), s0s, alphas, betas, cws, nullptr, nkeys, stream), "Dcf::GenBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (nkeys == 0) return;
      b200::generic::GenKernel<Dcf, In><<<b200::generic::GridFor(nkeys, 128), 128, 0, stream>>>(*this, s0s, alphas, betas, cws, nkeys);
      b200::generic::CheckLaunch("Dcf::GenBatch");
#endif
    }
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys,
                 cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_dcf_eval(Context(), b, seeds, cws, xs, ys, nkeys, stream), "Dcf::EvalBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (nkeys == 0) return;
      b200::generic::EvalKernel<Dcf, In><<<b200::generic::GridFor(nkeys, 128), 128, 0, stream>>>(*this, b, seeds, cws, xs, ys, nkeys);
      b200::generic::CheckLaunch("Dcf::EvalBatch");
#endif
    }
  }
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, int4 *ys, size_t nkeys, uint64_t leaf_begin = 0,
                    uint64_t leaf_count = 0, cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_all(Context(), b, seeds, cws, nullptr, ys, nkeys, leaf_begin, leaf_count, stream),
                  "Dcf::EvalAllBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (leaf_count == 0) leaf_count = (uint64_t(1) << in_bits) - leaf_begin;
      if (nkeys == 0) return;
      b200::generic::EvalAllKernel<Dcf, In><<<b200::generic::GridFor(nkeys * leaf_count, 128), 128, 0, stream>>>(
          *this, b, seeds, cws, ys, nkeys, leaf_begin, leaf_count);
      b200::generic::CheckLaunch("Dcf::EvalAllBatch");
#endif
    }
  }
  void EvalBatchHost(bool b, const int4 *seeds, const Cw *cws, const In *xs, int4 *ys, size_t nkeys) const
    requires kPrebuilt
  {
    b200::Check(fssb200_eval_host(Context(), b, seeds, cws, nullptr, xs, ys, nkeys), "Dcf::EvalBatchHost");
  }
  // one process, ndev GPUs: per-device arrays, one stream per device, no collective (see Dpf::EvalBatchMulti)
  void EvalBatchMulti(bool b, int ndev, const int *devices, const int4 *const *seeds, const Cw *const *cws,
                      const In *const *xs, int4 *const *ys, const size_t *nkeys,
                      const cudaStream_t *streams = nullptr) const
    requires kPrebuilt
  {
    b200::MultiCall m(Params(), ndev, devices);
    m.Check(fssb200_eval_multi(m.ctxs.data(), ndev, b, reinterpret_cast<const void *const *>(seeds),
                               reinterpret_cast<const void *const *>(cws), nullptr,
                               reinterpret_cast<const void *const *>(xs), reinterpret_cast<void *const *>(ys), nkeys,
                               reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dcf::EvalBatchMulti");
  }
  void SyncMulti(int ndev, const int *devices, const cudaStream_t *streams = nullptr) const
    requires kPrebuilt
  {
    b200::MultiCall m(Params(), ndev, devices);
    m.Check(fssb200_multi_sync(m.ctxs.data(), ndev, reinterpret_cast<void *const *>(streams), m.rcs.data()),
            "Dcf::SyncMulti");
  }

private:
  // The single-key members above reach the batched members (and through them the generic kernels) only in their HOST
  // branch.  nvcc instantiates a __global__ template for the device only if the instantiation is also seen while
  // __CUDA_ARCH__ is defined, so the members name the kernels once outside the branch.
  FSS_SHIM_HD static void KeepGenericKernels() {
#if defined(__CUDACC__)
    if constexpr (!kPrebuilt) {
      [[maybe_unused]] auto g = &b200::generic::GenKernel<Dcf, In>;
      [[maybe_unused]] auto e = &b200::generic::EvalKernel<Dcf, In>;
      [[maybe_unused]] auto a = &b200::generic::EvalAllKernel<Dcf, In>;
    }
#endif
  }
  static void UserPluginNeedsNvcc() {
#if !defined(__CUDACC__)
    static_assert(kPrebuilt, "a user-defined Group / Prg plugin is compiled for the GPU in YOUR translation unit: "
                             "build it with nvcc (fss/b200/generic.cuh); there is no CPU evaluation path");
#endif
  }
};

}  // namespace fss
