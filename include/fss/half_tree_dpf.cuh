// SPDX-License-Identifier: Apache-2.0
// fss/half_tree_dpf.cuh -- Half-Tree DPF (reference half_tree_dpf.cuh:39-355): same class template,
// `prg` / `hash_key` members, `Cw` layout (n entries + a separate output CW) and member signatures.
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

template <int in_bits, typename Group, typename Prg, typename In = uint, int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           Groupable<Group> && Prgable<Prg, 1>)
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
  static constexpr bool kPrebuilt = b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 1>;

  fssb200_ctx *Context() const
    requires kPrebuilt
  {
    return b200::ContextFor(b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_HALFTREE, prg, FSSB200_PRED_LT, &hash_key));
  }

  FSS_SHIM_HD void Gen(Cw cws[], int4 &ocw, const int4 s0s[2], In a, int4 b_buf) const {            // :68
    KeepGenericKernels();
#if defined(__CUDA_ARCH__)
    b200::generic::HalfTreeGen<in_bits, Group, In>(const_cast<Prg &>(prg), hash_key, cws, ocw, s0s, a, b_buf);
#else
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_gen_host(Context(), s0s, &a, &b_buf, cws, &ocw, 1), "HalfTreeDpf::Gen");
    } else {
      b200::DeviceArray<int4> ds(4);
      b200::DeviceArray<In> da(1);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, s0s, 2);
      ds.Upload(2, &b_buf, 1);
      da.Upload(0, &a, 1);
      GenBatch(ds.ptr, da.ptr, ds.ptr + 2, dc.ptr, ds.ptr + 3, 1);
      dc.Download(0, cws, kNumCw);
      ds.Download(3, &ocw, 1);
    }
#endif
  }
  FSS_SHIM_HD int4 Eval(bool b, int4 s0, const Cw cws[], int4 ocw, In x) const {                     // :187
    KeepGenericKernels();
#if defined(__CUDA_ARCH__)
    return b200::generic::HalfTreeEval<in_bits, Group, In>(const_cast<Prg &>(prg), hash_key, b, s0, cws, ocw, x);
#else
    int4 y;
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_host(Context(), b, &s0, cws, &ocw, &x, &y, 1), "HalfTreeDpf::Eval");
    } else {
      b200::DeviceArray<int4> ds(3);
      b200::DeviceArray<In> dx(1);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, &s0, 1);
      ds.Upload(1, &ocw, 1);
      dx.Upload(0, &x, 1);
      dc.Upload(0, cws, kNumCw);
      EvalBatch(b, ds.ptr, dc.ptr, ds.ptr + 1, dx.ptr, ds.ptr + 2, 1);
      ds.Download(2, &y, 1);
    }
    return y;
#endif
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], int4 ocw, int4 ys[]) const {             // :246
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, &ocw, ys, 1, 0, 0), "HalfTreeDpf::EvalAll");
    } else {
      static_assert(in_bits <= 40, "EvalAll: 2^in_bits leaves");
      const size_t n = size_t(1) << in_bits;
      b200::DeviceArray<int4> ds(2), dy(n);
      b200::DeviceArray<Cw> dc(kNumCw);
      ds.Upload(0, &s0, 1);
      ds.Upload(1, &ocw, 1);
      dc.Upload(0, cws, kNumCw);
      EvalAllBatch(b, ds.ptr, dc.ptr, ds.ptr + 1, dy.ptr, 1);
      dy.Download(0, ys, n);
    }
  }

  void GenBatch(const int4 *s0s, const In *alphas, const int4 *betas, Cw *cws, int4 *ocws, size_t nkeys,
                cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_gen(Context(), s0s, alphas, betas, cws, ocws, nkeys, stream), "HalfTreeDpf::GenBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (nkeys == 0) return;
      b200::generic::HtGenKernel<HalfTreeDpf, In><<<b200::generic::GridFor(nkeys, 128), 128, 0, stream>>>(*this, s0s, alphas, betas, cws, ocws, nkeys);
      b200::generic::CheckLaunch("HalfTreeDpf::GenBatch");
#endif
    }
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *ocws, const In *xs, int4 *ys, size_t nkeys,
                 cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_halftree_eval(Context(), b, seeds, cws, ocws, xs, ys, nkeys, stream), "HalfTreeDpf::EvalBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (nkeys == 0) return;
      b200::generic::HtEvalKernel<HalfTreeDpf, In><<<b200::generic::GridFor(nkeys, 128), 128, 0, stream>>>(*this, b, seeds, cws, ocws, xs, ys, nkeys);
      b200::generic::CheckLaunch("HalfTreeDpf::EvalBatch");
#endif
    }
  }
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *ocws, int4 *ys, size_t nkeys,
                    uint64_t leaf_begin = 0, uint64_t leaf_count = 0, cudaStream_t stream = nullptr) const {
    if constexpr (kPrebuilt) {
      b200::Check(fssb200_eval_all(Context(), b, seeds, cws, ocws, ys, nkeys, leaf_begin, leaf_count, stream),
                  "HalfTreeDpf::EvalAllBatch");
    } else {
      UserPluginNeedsNvcc();
#if defined(__CUDACC__)
      if (leaf_count == 0) leaf_count = (uint64_t(1) << in_bits) - leaf_begin;
      if (nkeys == 0) return;
      b200::generic::HtEvalAllKernel<HalfTreeDpf, In><<<b200::generic::GridFor(nkeys * leaf_count, 128), 128, 0, stream>>>(
          *this, b, seeds, cws, ocws, ys, nkeys, leaf_begin, leaf_count);
      b200::generic::CheckLaunch("HalfTreeDpf::EvalAllBatch");
#endif
    }
  }

private:
  FSS_SHIM_HD static void KeepGenericKernels() {  // (see fss/dpf.cuh)
#if defined(__CUDACC__)
    if constexpr (!kPrebuilt) {
      [[maybe_unused]] auto g = &b200::generic::HtGenKernel<HalfTreeDpf, In>;
      [[maybe_unused]] auto e = &b200::generic::HtEvalKernel<HalfTreeDpf, In>;
      [[maybe_unused]] auto a = &b200::generic::HtEvalAllKernel<HalfTreeDpf, In>;
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
