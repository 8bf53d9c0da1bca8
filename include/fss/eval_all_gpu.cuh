// SPDX-License-Identifier: Apache-2.0
// fss/eval_all_gpu.cuh -- the reference's GPU full-domain entry points (eval_all_gpu.cuh:451-535), same
// names and argument lists.  The tuning template parameters <z, b1, bs> of the reference's hybrid kernel
// (frontier depth, block-root depth, block size; constraints in_bits - z <= 8 and 2^(z-b1) == bs) are
// accepted and ignored: the B200 kernel plans its own decomposition, has no such constraints and indexes
// leaves with 64 bits (the reference's `int` indexing overflows at n = 28, key >= 8, eval_all_gpu.cuh:287).
#pragma once
#include <fss/dpf.cuh>
#include <fss/half_tree_dpf.cuh>

namespace fss::gpu {

template <int z = -1, int b1 = 8, int bs = 256, int in_bits, typename Group, typename Prg, typename In>
void DpfEvalAllGpu(bool b, int4 s0, const typename Dpf<in_bits, Group, Prg, In>::Cw *cws, int4 *ys,
                   const Dpf<in_bits, Group, Prg, In> &dpf, cudaStream_t stream = nullptr) {
  b200::DeviceBlock seed(s0, stream);
  dpf.EvalAllBatch(b, seed.ptr, cws, ys, 1, 0, 0, stream);
}
template <int z = -1, int b1 = 8, int bs = 256, int in_bits, typename Group, typename Prg, typename In>
void DpfEvalAllGpuBatch(bool b, const int4 *s0s, const typename Dpf<in_bits, Group, Prg, In>::Cw *cws, int nkeys,
                        int4 *ys, const Dpf<in_bits, Group, Prg, In> &dpf, cudaStream_t stream = nullptr) {
  dpf.EvalAllBatch(b, s0s, cws, ys, nkeys, 0, 0, stream);
}
template <int z = -1, int b1 = 8, int bs = 256, int in_bits, typename Group, typename Prg, typename In>
void HalfTreeDpfEvalAllGpu(bool b, int4 s0, const typename HalfTreeDpf<in_bits, Group, Prg, In>::Cw *cws, int4 ocw,
                           int4 *ys, const HalfTreeDpf<in_bits, Group, Prg, In> &dpf, cudaStream_t stream = nullptr) {
  b200::DeviceBlock seed(s0, stream), o(ocw, stream);
  dpf.EvalAllBatch(b, seed.ptr, cws, o.ptr, ys, 1, 0, 0, stream);
}
template <int z = -1, int b1 = 8, int bs = 256, int in_bits, typename Group, typename Prg, typename In>
void HalfTreeDpfEvalAllGpuBatch(bool b, const int4 *s0s, const typename HalfTreeDpf<in_bits, Group, Prg, In>::Cw *cws,
                                const int4 *ocws, int nkeys, int4 *ys, const HalfTreeDpf<in_bits, Group, Prg, In> &dpf,
                                cudaStream_t stream = nullptr) {
  dpf.EvalAllBatch(b, s0s, cws, ocws, ys, nkeys, 0, 0, stream);
}

}  // namespace fss::gpu
