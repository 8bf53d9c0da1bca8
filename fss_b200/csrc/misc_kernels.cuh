// SPDX-License-Identifier: Apache-2.0
//
// misc_kernels.cuh -- host launchers of the small helper kernels (misc_kernels.cu): level-major
// relayout, Grotto scan / parity tree / lookup, integer-pipe microbenchmarks.
#pragma once
#include "common.cuh"

namespace fssb200 {

// fssb200_relayout: key-major Cw[nkeys][ncw] -> level-major (point_eval_gpu.cuh:39-91).
cudaError_t launch_relayout(int scheme, int in_bits, int ncw, const uint8_t *cws, blk *cw_s, blk *cw_v,
    uint32_t *extra, blk *out_cw, uint64_t nkeys, cudaStream_t stream);
// In-place inclusive prefix XOR over `len` bytes (values 0/1) for each of nkeys rows
// (grotto_dcf.cuh:160-162).
cudaError_t launch_prefix_xor(uint8_t *ys, uint64_t nkeys, uint64_t len, cudaStream_t stream);
// Heap-ordered parity trees p[j] = p[2j+1] ^ p[2j+2] (grotto_dcf.cuh:100-103) of nkeys keys, rows key_stride
// bytes apart: fills the min(bottom, kParityLevels) levels above level `bottom` (whose 2^bottom nodes are present).
constexpr int kParityLevels = 13;
cudaError_t launch_parity_levels(uint8_t *pt, uint64_t key_stride, uint64_t nkeys, int bottom, cudaStream_t stream);
// GrottoDcf::Eval lookup (grotto_dcf.cuh:116-135), one thread per key.
cudaError_t launch_grotto_lookup(const uint8_t *pt, const uint8_t *xs, uint8_t *ys, uint64_t nkeys, int in_bits,
    int in_bytes, cudaStream_t stream);
// Issue-rate microbenchmarks; see fssb200_microbench.
int run_microbench(int kind, double *ops_per_s);

}  // namespace fssb200
