// SPDX-License-Identifier: Apache-2.0
//
// vdpf_kernels.cu -- VDPF helper kernels (vdpf.cuh): the point / gen walks live in kernels.cuh
// (point_kernel / gen_kernel with SCHEME = VDPF); here are the pieces around them.
#include "vdpf_kernels.cuh"

#include <cstdlib>

namespace fssb200 {

// ---- VDPF helpers -------------------------------------------------------------------------------------------------
// Hash known-answer kernel: which = 0 XorHash ((a, b) -> 64 B), 1 Hash (64 B -> 32 B).
__global__ void __launch_bounds__(256) hash_kernel(const __grid_constant__ KParams P, int which, const blk *msgs,
    blk *out, uint64_t n) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (which == 0) {
    blk o[4];
    vdpf_xor_hash(P.keys, ld_blk(msgs + 2 * i), ld_blk(msgs + 2 * i + 1), o);
#pragma unroll
    for (int j = 0; j < 4; ++j) st_blk(out + 4 * i + j, o[j]);
  } else {
    blk m[4], o[2];
#pragma unroll
    for (int j = 0; j < 4; ++j) m[j] = ld_blk(msgs + 4 * i + j);
    vdpf_hash(P.keys, m, o);
    st_blk(out + 2 * i, o[0]);
    st_blk(out + 2 * i + 1, o[1]);
  }
}

// Vdpf::Prove (vdpf.cuh:254-264) for a batch: thread k walks key k's m hashes in order.
__global__ void __launch_bounds__(128) vdpf_prove_kernel(const __grid_constant__ KParams P, const blk *pts,
    const blk *cs, uint64_t m, blk *pis, uint64_t nkeys) {
  const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= nkeys) return;
  blk pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = ld_blk(cs + 4 * k + j);
  for (uint64_t i = 0; i < m; ++i) {
    blk pt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pt[j] = ld_blk(pts + 4 * (k * m + i) + j);
    vdpf_accumulate(P.keys, pi, pt);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) st_blk(pis + 4 * k + j, pi[j]);
}

// Second half of Vdpf::EvalAll (vdpf.cuh:313-341): ys holds the packed (s | t) leaves written by
// evalall_kernel<4>.  LPK lanes per key: LPK leaves at a time the lanes convert their leaf to the output share
// and compute its corrected hash in parallel (two compressions); the accumulation H'(pi ^ pi_tilde) is
// sequential in x by definition, so the group then replays its LPK hashes in order through shuffles, every lane
// of the group carrying the same proof.  A warp instruction of the replay advances 32 / LPK keys: with many keys the
// replay is issue-bound and a small LPK wins, with few keys it is latency-bound and LPK = 32 (most leaf hashes in
// parallel) wins; launch_vdpf_finish picks LPK from the key count.
template <int G, int LPK>
__global__ void __launch_bounds__(128) vdpf_finish_kernel(const __grid_constant__ KParams P, int party, int in_bits,
    const blk *cs, const blk *ocws, blk *ys, blk *pis, uint64_t nkeys) {
  const uint32_t lane = threadIdx.x & 31u, sub = lane % LPK;
  const uint64_t kraw = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) / LPK;
  if (((uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) & ~uint64_t(31)) / LPK >= nkeys) return;  // a whole warp past the batch
  const bool live = kraw < nkeys;                    // dead groups of the last warp run along (full-mask shuffles), store nothing
  const uint64_t k = live ? kraw : nkeys - 1;
  const uint64_t N = uint64_t(1) << in_bits;
  const blk ocw = ld_blk(ocws + k);
  blk pi[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = ld_blk(cs + 4 * k + j);
  for (uint64_t base = 0; base < N; base += LPK) {
    const uint64_t x = base + sub;
    const bool have = x < N;
    blk pt[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pt[j] = zero_blk();
    if (have) {
      InVal xv;
      xv.w[0] = uint32_t(x); xv.w[1] = uint32_t(x >> 32); xv.w[2] = 0; xv.w[3] = 0;
      const blk y = vdpf_leaf<G>(P.keys, P.ga, uint32_t(party), ys[k * N + x], xv, ocw, cs + 4 * k, pt);
      if (live) st_blk(ys + k * N + x, y);
    }
    const uint32_t cnt = N - base < LPK ? uint32_t(N - base) : uint32_t(LPK);
    for (uint32_t src = 0; src < cnt; ++src) {
      blk q[4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
        q[j] = make_blk(__shfl_sync(0xffffffffu, pt[j].x, src, LPK), __shfl_sync(0xffffffffu, pt[j].y, src, LPK),
            __shfl_sync(0xffffffffu, pt[j].z, src, LPK), __shfl_sync(0xffffffffu, pt[j].w, src, LPK));
      vdpf_accumulate(P.keys, pi, q);
    }
  }
  if (sub == 0 && live) {
#pragma unroll
    for (int j = 0; j < 4; ++j) st_blk(pis + 4 * k + j, pi[j]);
  }
}

cudaError_t launch_hash(const KParams &P, int which, const blk *msgs, blk *out, uint64_t n, cudaStream_t stream) {
  if (n == 0) return cudaSuccess;
  hash_kernel<<<unsigned((n + 255) / 256), 256, 0, stream>>>(P, which, msgs, out, n);
  return cudaGetLastError();
}
cudaError_t launch_vdpf_prove(const KParams &P, const blk *pts, const blk *cs, uint64_t m, blk *pis, uint64_t nkeys,
    cudaStream_t stream) {
  if (nkeys == 0) return cudaSuccess;
  vdpf_prove_kernel<<<unsigned((nkeys + 127) / 128), 128, 0, stream>>>(P, pts, cs, m, pis, nkeys);
  return cudaGetLastError();
}
// Lanes per key of the finish kernel.  Measured on a B200 (profiles/r02_vdpf_finish.md): the leaf phase (share conversion +
// corrected hash) costs ~5000 warp instructions and ~6000 cycles of latency, one link of the chain ~1350 and ~1400.  A warp
// does 32 key-leaves per (leaf phase + LPK links), so per key and leaf the step costs (5000 + 1350 LPK) / 32 issue slots on
// 4 * SMs schedulers, or 1400 + 6000 / LPK cycles of latency, whichever is larger.
static int vdpf_finish_lanes(uint64_t nkeys, int sms) {
  if (const char *e = std::getenv("FSSB200_VDPF_FINISH_LANES")) {
    const int v = std::atoi(e);
    if (v == 1 || v == 2 || v == 4 || v == 8 || v == 16 || v == 32) return v;
  }
  int best = 32;
  double best_cost = 1e300;
  for (int g = 32; g >= 1; g >>= 1) {
    const double issue = double(nkeys) * (5000.0 + 1350.0 * g) / (32.0 * 4.0 * sms), lat = 1400.0 + 6000.0 / g;
    const double cost = issue > lat ? issue : lat;
    if (cost < best_cost * 0.98) best_cost = cost, best = g;
  }
  return best;
}

cudaError_t launch_vdpf_finish(const KParams &P, int gk, int party, int in_bits, const blk *cs, const blk *ocws, blk *ys,
    blk *pis, uint64_t nkeys, cudaStream_t stream) {
  if (nkeys == 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int lpk = vdpf_finish_lanes(nkeys, sms);
  const unsigned grid = unsigned((nkeys * uint64_t(lpk) + 127) / 128);
#define Y(GK, L) \
  case L: vdpf_finish_kernel<GK, L><<<grid, 128, 0, stream>>>(P, party, in_bits, cs, ocws, ys, pis, nkeys); break;
#define X(GK) \
  case GK: \
    switch (lpk) { Y(GK, 1) Y(GK, 2) Y(GK, 4) Y(GK, 8) Y(GK, 16) Y(GK, 32) } \
    break;
  switch (gk) {
    X(kGrpBytes) X(kGrpU32) X(kGrpU64) X(kGrpU127) X(kGrpU32Mod) X(kGrpU64Mod) X(kGrpU128Mod)
    default: return cudaErrorInvalidValue;
  }
#undef X
#undef Y
  return cudaGetLastError();
}

}  // namespace fssb200
