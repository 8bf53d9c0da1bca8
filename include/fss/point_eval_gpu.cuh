// SPDX-License-Identifier: Apache-2.0
// fss/point_eval_gpu.cuh -- the reference's batched GPU point-evaluation entry points
// (point_eval_gpu.cuh:324-526), same names and argument lists, forwarding to the C ABI.
//
// The level-major arrays are the reference's (`cw_s[i*nkeys+k]`, `cw_v`, packed `extra`, `out_cw`); for
// in_bits > 32 `extra` holds ceil(in_bits/32) words per key (`extra[w*nkeys+k]`), which the reference
// cannot represent.  The relayout pre-pass is optional here: `Dpf::EvalBatch` reads the key-major `Cw`
// array directly.  `unroll` is accepted for source compatibility and ignored.
#pragma once
#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/half_tree_dpf.cuh>
#include <fss/vdpf.cuh>

namespace fss::gpu {

namespace detail {
// The relayout does not touch the PRG: a key-less context of the right scheme / domain is enough, which
// keeps the reference's argument lists (no scheme object, point_eval_gpu.cuh:324-381).
template <int in_bits, typename In>
fssb200_ctx *RelayoutContext(int scheme) {
  fssb200_params p;
  std::memset(&p, 0, sizeof(p));
  p.scheme = scheme;
  p.in_bits = in_bits;
  p.in_bytes = sizeof(In);
  p.prg = FSSB200_PRG_CHACHA;
  return b200::ContextFor(p);
}
}  // namespace detail

template <int in_bits, typename Group, typename Prg, typename In>
void DpfRelayoutGpu(const typename Dpf<in_bits, Group, Prg, In>::Cw *cws, int nkeys, int4 *cw_s, uint32_t *extra,
                    int4 *out_cw, cudaStream_t stream = nullptr) {
  b200::Check(fssb200_relayout(detail::RelayoutContext<in_bits, In>(FSSB200_SCHEME_DPF), cws, cw_s, nullptr, extra,
                               out_cw, nkeys, stream), "DpfRelayoutGpu");
}
template <int in_bits, typename Group, typename Prg, typename In>
void DcfRelayoutGpu(const typename Dcf<in_bits, Group, Prg, In>::Cw *cws, int nkeys, int4 *cw_s, int4 *cw_v,
                    int4 *out_cw, cudaStream_t stream = nullptr) {
  b200::Check(fssb200_relayout(detail::RelayoutContext<in_bits, In>(FSSB200_SCHEME_DCF), cws, cw_s, cw_v, nullptr,
                               out_cw, nkeys, stream), "DcfRelayoutGpu");
}
template <int in_bits, typename Group, typename Prg, typename In>
void HalfTreeDpfRelayoutGpu(const typename HalfTreeDpf<in_bits, Group, Prg, In>::Cw *cws, int nkeys, int4 *cw_s,
                            uint32_t *extra, cudaStream_t stream = nullptr) {
  b200::Check(fssb200_relayout(detail::RelayoutContext<in_bits, In>(FSSB200_SCHEME_HALFTREE), cws, cw_s, nullptr,
                               extra, nullptr, nkeys, stream), "HalfTreeDpfRelayoutGpu");
}

// point_eval_gpu.cuh:448-460
template <int unroll = 4, int in_bits, typename Group, typename Prg, typename In>
void DpfEvalPointGpu(bool b, const int4 *seeds, const int4 *cw_s, const uint32_t *extra, const int4 *out_cw,
                     const In *xs, int4 *ys, int nkeys, const Dpf<in_bits, Group, Prg, In> &dpf,
                     cudaStream_t stream = nullptr) {
  b200::Check(fssb200_eval_levelmajor(dpf.Context(), b, seeds, cw_s, nullptr, extra, out_cw, nullptr, xs, ys, nkeys, stream),
              "DpfEvalPointGpu");
}
// point_eval_gpu.cuh:480-492
template <int unroll = 4, int in_bits, typename Group, typename Prg, typename In, DcfPred pred>
void DcfEvalPointGpu(bool b, const int4 *seeds, const int4 *cw_s, const int4 *cw_v, const int4 *out_cw, const In *xs,
                     int4 *ys, int nkeys, const Dcf<in_bits, Group, Prg, In, pred> &dcf, cudaStream_t stream = nullptr) {
  b200::Check(fssb200_eval_levelmajor(dcf.Context(), b, seeds, cw_s, cw_v, nullptr, out_cw, nullptr, xs, ys, nkeys, stream),
              "DcfEvalPointGpu");
}
// point_eval_gpu.cuh:416-428
template <int unroll = 4, int in_bits, typename Group, typename Prg, typename In>
void HalfTreeDpfEvalPointGpu(bool b, const int4 *seeds, const int4 *cw_s, const uint32_t *extra, const int4 *ocws,
                             const In *xs, int4 *ys, int nkeys, const HalfTreeDpf<in_bits, Group, Prg, In> &dpf,
                             cudaStream_t stream = nullptr) {
  b200::Check(fssb200_eval_levelmajor(dpf.Context(), b, seeds, cw_s, nullptr, extra, nullptr, ocws, xs, ys, nkeys, stream),
              "HalfTreeDpfEvalPointGpu");
}

// point_eval_gpu.cuh:389-396: cw_s[i * nkeys + k] and the packed control bits; cs and ocws need no relayout
template <int in_bits, typename Group, typename Prg, typename XorHash, typename Hash, typename In>
void VdpfRelayoutGpu(const typename Vdpf<in_bits, Group, Prg, XorHash, Hash, In>::Cw *cws, int nkeys, int4 *cw_s,
                     uint32_t *extra, cudaStream_t stream = nullptr) {
  b200::Check(fssb200_relayout(detail::RelayoutContext<in_bits, In>(FSSB200_SCHEME_VDPF), cws, cw_s, nullptr, extra,
                               nullptr, nkeys, stream), "VdpfRelayoutGpu");
}
// point_eval_gpu.cuh:513-526: y share and corrected per-point hash (four int4 per key) of every key.  The reference
// takes `const uint32_t *xs` whatever In is; so does this, for the In it is usable with (4 bytes).
template <int unroll = 4, int in_bits, typename Group, typename Prg, typename XorHash, typename Hash, typename In>
void VdpfEvalPointGpu(bool b, const int4 *seeds, const int4 *cw_s, const uint32_t *extra,
                      const cuda::std::array<int4, 4> *cs, const int4 *ocws, const uint32_t *xs, int4 *ys, int4 *pi,
                      int nkeys, const Vdpf<in_bits, Group, Prg, XorHash, Hash, In> &vdpf, cudaStream_t stream = nullptr) {
  static_assert(sizeof(In) == 4, "VdpfEvalPointGpu reads 32-bit inputs (point_eval_gpu.cuh:515)");
  b200::Check(fssb200_vdpf_eval_levelmajor(vdpf.Context(), b, seeds, cw_s, extra, cs, ocws, xs, ys, pi, nkeys, stream),
              "VdpfEvalPointGpu");
}

}  // namespace fss::gpu
