// SPDX-License-Identifier: Apache-2.0
// fss/grotto_dcf.cuh -- Grotto DCF over F2 (reference grotto_dcf.cuh:45-239): same class template, `Cw`
// (= Dpf<..., Bytes, ...>::Cw), ParityTree, Gen / Preprocess / Eval / EvalAll.
#pragma once
#include <vector>
#include <fss/dpf.cuh>
#include <fss/group/bytes.cuh>

namespace fss {

template <int in_bits, typename Prg, typename In = uint, int par_depth = -1>
  requires((std::is_unsigned_v<In> || std::is_same_v<In, __uint128_t>) && in_bits <= sizeof(In) * 8 &&
           b200::DevicePrg<Prg, 2>)
class GrottoDcf {
  using DpfType = Dpf<in_bits, group::Bytes, Prg, In, par_depth>;

public:
  using Cw = typename DpfType::Cw;
  Prg prg;

  // grotto_dcf.cuh:78-81: p[0..2N-2] level order, leaf x at p[x + N - 1]
  struct ParityTree {
    bool *p;
    bool b;
  };

  fssb200_ctx *Context() const {
    return b200::ContextFor(b200::MakeParams<in_bits, group::Bytes, Prg, In>(FSSB200_SCHEME_GROTTO, prg));
  }

  void Gen(Cw cws[], const int4 s0s[2], In a) const {                                    // :63
    b200::Check(fssb200_gen_host(Context(), s0s, &a, nullptr, cws, nullptr, 1), "GrottoDcf::Gen");
  }
  // Host array in, host array out: the tree is built on the device and copied back.
  void Preprocess(ParityTree &pt, int4 s0, const Cw cws[]) const {                        // :94
    static_assert(in_bits <= 31, "parity tree of 2^(n+1)-1 bytes");
    const size_t bytes = (size_t(2) << in_bits) - 1;
    b200::DeviceBlock seed(s0, nullptr);
    uint8_t *d_tree = nullptr;
    Cw *d_cws = nullptr;
    if (cudaMallocAsync(reinterpret_cast<void **>(&d_tree), bytes, nullptr) != cudaSuccess) throw std::bad_alloc();
    if (cudaMallocAsync(reinterpret_cast<void **>(&d_cws), sizeof(Cw) * (in_bits + 1), nullptr) != cudaSuccess) throw std::bad_alloc();
    cudaMemcpyAsync(d_cws, cws, sizeof(Cw) * (in_bits + 1), cudaMemcpyHostToDevice, nullptr);
    b200::Check(fssb200_grotto_preprocess(Context(), pt.b, seed.ptr, d_cws, d_tree, 1, nullptr), "GrottoDcf::Preprocess");
    cudaMemcpy(pt.p, d_tree, bytes, cudaMemcpyDeviceToHost);
    cudaFreeAsync(d_tree, nullptr);
    cudaFreeAsync(d_cws, nullptr);
  }
  // Static prefix-parity lookup (grotto_dcf.cuh:116-135).  `pt.p` is a host array here, as in the
  // reference; the lookup itself runs on the device (fssb200_grotto_eval), so this convenience member
  // uploads the tree first -- keep trees device-resident and use EvalBatch for anything but spot checks.
  static bool Eval(const ParityTree &pt, In x) {
    static_assert(in_bits <= 31, "parity tree of 2^(n+1)-1 bytes");
    const size_t bytes = (size_t(2) << in_bits) - 1;
    fssb200_params p;
    std::memset(&p, 0, sizeof(p));
    p.scheme = FSSB200_SCHEME_GROTTO;
    p.in_bits = in_bits;
    p.in_bytes = sizeof(In);
    p.prg = FSSB200_PRG_CHACHA;  // the lookup does not touch the PRG
    fssb200_ctx *ctx = b200::ContextFor(p);
    uint8_t *d = nullptr;
    if (cudaMalloc(reinterpret_cast<void **>(&d), bytes + 2 * sizeof(In) + 16) != cudaSuccess) throw std::bad_alloc();
    cudaMemcpy(d, pt.p, bytes, cudaMemcpyHostToDevice);
    uint8_t *d_xa = d + (((bytes + sizeof(In) - 1) / sizeof(In)) * sizeof(In));  // In-aligned slot behind the tree
    cudaMemcpy(d_xa, &x, sizeof(In), cudaMemcpyHostToDevice);
    uint8_t *d_y = d_xa + sizeof(In);
    const int rc = fssb200_grotto_eval(ctx, d, d_xa, d_y, 1, nullptr);
    uint8_t y = 0;
    cudaMemcpy(&y, d_y, 1, cudaMemcpyDeviceToHost);
    cudaFree(d);
    b200::Check(rc, "GrottoDcf::Eval");
    return y != 0;
  }
  void EvalAll(bool b, int4 s0, const Cw cws[], bool ys[]) const {                         // :151
    b200::Check(fssb200_eval_all_host(Context(), b, &s0, cws, nullptr, ys, 1, 0, 0), "GrottoDcf::EvalAll");
  }

  void EvalAllBatch(bool b, const int4 *seeds, const Cw *cws, bool *ys, size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_eval_all(Context(), b, seeds, cws, nullptr, ys, nkeys, 0, 0, stream), "GrottoDcf::EvalAllBatch");
  }
  void PreprocessBatch(bool b, const int4 *seeds, const Cw *cws, bool *pts, size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_grotto_preprocess(Context(), b, seeds, cws, pts, nkeys, stream), "GrottoDcf::PreprocessBatch");
  }
  void EvalBatch(const bool *pts, const In *xs, bool *ys, size_t nkeys, cudaStream_t stream = nullptr) const {
    b200::Check(fssb200_grotto_eval(Context(), pts, xs, ys, nkeys, stream), "GrottoDcf::EvalBatch");
  }
};

}  // namespace fss
