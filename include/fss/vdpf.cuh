// SPDX-License-Identifier: Apache-2.0
// fss/vdpf.cuh -- verifiable DPF (reference vdpf.cuh:63-403): same class template, template parameter list, member
// signatures and `Cw` layout; every HOST call evaluates on the B200 through the C ABI (include/fssb200.h,
// fssb200_vdpf_*).  `Gen` and `Eval` are also callable per thread inside the user's own kernels, like the reference's
// `__host__ __device__` members (its src/bench_gpu.cu does: VdpfGenKernel / VdpfEvalKernel), through the plugin-generic
// templates of fss/b200/generic.cuh -- with plugins whose members are device-callable (prg::ChaCha, hash::Blake3).
// Added: batched members taking device arrays.
#pragma once
#include <sys/types.h>
#include <vector>
#include <fss/b200/generic.cuh>
#include <fss/b200/runtime.hpp>
#include <fss/group.cuh>
#include <fss/hash.cuh>
#include <fss/prg.cuh>
#include <fss/util.cuh>

namespace fss {

template <int in_bits, typename Group, typename Prg, typename XorHash, typename Hash, typename In = uint,
    int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           b200::DeviceGroup<Group> && b200::DevicePrg<Prg, 2> && XorHashable<XorHash> && Hashable<Hash> &&
           b200::DeviceHash<XorHash> && b200::DeviceHash<Hash>)
class Vdpf {
public:
  Prg prg;
  XorHash xor_hash;
  Hash hash;

  // vdpf.cuh:77-80: in_bits entries, no output entry (the output correction word travels separately)
  struct alignas(32) Cw {
    int4 s;
    bool tr;
  };
  static_assert(sizeof(Cw) == 32);

  fssb200_ctx *Context() const {
    fssb200_params p = b200::MakeParams<in_bits, Group, Prg, In>(FSSB200_SCHEME_VDPF, prg);
    xor_hash.FssB200Iv(p.hash_iv[0]);
    hash.FssB200Iv(p.hash_iv[1]);
    p.hash = XorHash::kFssB200Hash | (Hash::kFssB200Hash << 8);
    return b200::ContextFor(p);
  }

  // ---- the reference's single-key members (host arrays) ----
  // vdpf.cuh:101: returns 1 if t0 == t1 at the end (resample the seeds and retry)
  FSS_SHIM_HD int Gen(Cw cws[], cuda::std::array<int4, 4> &cs, int4 &ocw, cuda::std::span<const int4, 2> s0s, In a, int4 b_buf) const {
#if defined(__CUDA_ARCH__)
    return b200::generic::VdpfGen<in_bits, Group, In>(const_cast<Prg &>(prg), const_cast<XorHash &>(xor_hash), cws, cs, ocw, s0s.data(), a, b_buf);
#else
    int32_t status = 0;
    b200::Check(fssb200_vdpf_gen_host(Context(), s0s.data(), &a, &b_buf, cws, cs.data(), &ocw, &status, 1), "Vdpf::Gen");
    return status;
#endif
  }
  // vdpf.cuh:195: y share through `y`, corrected per-point hash returned
  FSS_SHIM_HD cuda::std::array<int4, 4> Eval(bool b, int4 s0, cuda::std::span<const Cw> cws, cuda::std::span<const int4, 4> cs,
      int4 ocw, In x, int4 &y) const {
#if defined(__CUDA_ARCH__)
    return b200::generic::VdpfEval<in_bits, Group, In>(const_cast<Prg &>(prg), const_cast<XorHash &>(xor_hash), b, s0, cws.data(), cs.data(), ocw, x, y);
#else
    cuda::std::array<int4, 4> pi{};
    b200::Check(fssb200_vdpf_eval_host(Context(), b, &s0, cws.data(), cs.data(), &ocw, &x, &y, pi.data(), 1), "Vdpf::Eval");
    return pi;
#endif
  }
  // vdpf.cuh:259
  void Prove(cuda::std::span<const cuda::std::array<int4, 4>> pi_tildes, cuda::std::span<const int4, 4> cs,
      cuda::std::array<int4, 4> &pi) const {
    const size_t m = pi_tildes.size();
    b200::DeviceArray<int4> d(4 * m + 8);
    if (m) d.Upload(0, reinterpret_cast<const int4 *>(pi_tildes.data()), 4 * m);
    d.Upload(4 * m, cs.data(), 4);
    b200::Check(fssb200_vdpf_prove(Context(), d.ptr, d.ptr + 4 * m, m, d.ptr + 4 * m + 4, 1, nullptr), "Vdpf::Prove");
    d.Download(4 * m + 4, pi.data(), 4);
  }
  // vdpf.cuh:276
  static bool Verify(cuda::std::span<const int4, 4> pi0, cuda::std::span<const int4, 4> pi1) {
    return std::memcmp(pi0.data(), pi1.data(), 64) == 0;
  }
  // vdpf.cuh:302: full domain + accumulated proof
  void EvalAll(bool b, int4 s0, cuda::std::span<const Cw> cws, cuda::std::span<const int4, 4> cs, int4 ocw,
      cuda::std::span<int4> ys, cuda::std::array<int4, 4> &pi) const {
    const size_t n = size_t(1) << in_bits;
    b200::DeviceArray<int4> d(1 + 2 * in_bits + 4 + 1 + n + 4);
    int4 *d_s0 = d.ptr, *d_cws = d_s0 + 1, *d_cs = d_cws + 2 * in_bits, *d_ocw = d_cs + 4, *d_ys = d_ocw + 1, *d_pi = d_ys + n;
    d.Upload(0, &s0, 1);
    d.Upload(1, reinterpret_cast<const int4 *>(cws.data()), 2 * size_t(in_bits));
    d.Upload(1 + 2 * in_bits, cs.data(), 4);
    d.Upload(1 + 2 * in_bits + 4, &ocw, 1);
    b200::Check(fssb200_vdpf_eval_all(Context(), b, d_s0, d_cws, d_cs, d_ocw, d_ys, d_pi, 1, nullptr), "Vdpf::EvalAll");
    d.Download(size_t(d_ys - d.ptr), ys.data(), n);
    d.Download(size_t(d_pi - d.ptr), pi.data(), 4);
  }

  // ---- batched, device pointers, stream ordered ----
  void GenBatch(const int4 *s0s /*[n][2]*/, const In *alphas, const int4 *betas, Cw *cws, int4 *cs /*[n][4]*/, int4 *ocws,
      int32_t *status, size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_vdpf_gen(Context(), s0s, alphas, betas, cws, cs, ocws, status, nkeys, stream), "Vdpf::GenBatch");
  }
  void EvalBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *cs, const int4 *ocws, const In *xs, int4 *ys,
      int4 *pis /*[n][4]*/, size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_vdpf_eval(Context(), b, seeds, cws, cs, ocws, xs, ys, pis, nkeys, stream), "Vdpf::EvalBatch");
  }
  void ProveBatch(const int4 *pi_tildes /*[n][m][4]*/, const int4 *cs, size_t m, int4 *pis, size_t nkeys,
      cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_vdpf_prove(Context(), pi_tildes, cs, m, pis, nkeys, stream), "Vdpf::ProveBatch");
  }
  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, const int4 *cs, const int4 *ocws, int4 *ys, int4 *pis,
      size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_vdpf_eval_all(Context(), b, seeds, cws, cs, ocws, ys, pis, nkeys, stream), "Vdpf::EvalAllBatch");
  }
};

}  // namespace fss
