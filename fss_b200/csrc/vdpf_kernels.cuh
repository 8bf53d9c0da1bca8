// SPDX-License-Identifier: Apache-2.0
//
// vdpf_kernels.cuh -- host launchers of the VDPF helper kernels (vdpf_kernels.cu): Blake3 known-answer hook,
// batched Prove, and the leaf conversion + proof chain of EvalAll.
#pragma once
#include "kernels.cuh"

namespace fssb200 {

cudaError_t launch_hash(const KParams &P, int which, const blk *msgs, blk *out, uint64_t n, cudaStream_t stream);
cudaError_t launch_vdpf_prove(const KParams &P, const blk *pts, const blk *cs, uint64_t m, blk *pis, uint64_t nkeys,
    cudaStream_t stream);
// gk: group kind (common.cuh); returns cudaErrorInvalidValue for a kind without instantiation
cudaError_t launch_vdpf_finish(const KParams &P, int gk, int party, int in_bits, const blk *cs, const blk *ocws, blk *ys,
    blk *pis, uint64_t nkeys, cudaStream_t stream);

}  // namespace fssb200
