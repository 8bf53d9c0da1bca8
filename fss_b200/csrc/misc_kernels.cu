// SPDX-License-Identifier: Apache-2.0
//
// misc_kernels.cu -- helper kernels around the PRG-tree hot path: level-major relayout, Grotto
// prefix-XOR scan / parity tree / lookup, and the integer-pipe / shared-memory issue-rate
// microbenchmarks that give the roofline denominators.
#include "misc_kernels.cuh"

namespace fssb200 {

// ---- relayout (point_eval_gpu.cuh:39-91) ------------------------------------------------------------------
// One warp per key: lanes read the key's 32-byte Cw entries coalesced (lane l -> level l, l+32, ...)
// and scatter them to the level-major arrays; the control bits are gathered with a ballot.
__global__ void relayout_kernel(int scheme, int n, int ncw, const uint8_t *__restrict__ cws, blk *__restrict__ cw_s,
    blk *__restrict__ cw_v, uint32_t *__restrict__ extra, blk *__restrict__ out_cw, uint64_t nkeys) {
  const uint64_t warp = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nkeys) return;
  const uint64_t k = warp;
  const uint4 *kc = reinterpret_cast<const uint4 *>(cws + k * uint64_t(ncw) * 32u);
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    uint32_t flag = 0;
    if (i < n) {
      const uint4 s = __ldg(kc + 2 * i), v = __ldg(kc + 2 * i + 1);
      reinterpret_cast<uint4 *>(cw_s)[uint64_t(i) * nkeys + k] = s;
      if (scheme == FSSB200_SCHEME_DCF) reinterpret_cast<uint4 *>(cw_v)[uint64_t(i) * nkeys + k] = v;
      flag = (v.x & 0xffu) != 0;  // Dpf::Cw::tr / HalfTreeDpf::Cw::extra (bool at byte 16)
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, flag);
    if (lane == 0 && scheme != FSSB200_SCHEME_DCF) extra[uint64_t(base >> 5) * nkeys + k] = bits;
  }
  if (lane == 0 && scheme != FSSB200_SCHEME_HALFTREE && scheme != FSSB200_SCHEME_VDPF && out_cw) {
    const uint4 s = __ldg(kc + 2 * n), v = __ldg(kc + 2 * n + 1);
    reinterpret_cast<uint4 *>(out_cw)[k] = scheme == FSSB200_SCHEME_DCF ? v : s;
  }
}

cudaError_t launch_relayout(int scheme, int in_bits, int ncw, const uint8_t *cws, blk *cw_s, blk *cw_v,
    uint32_t *extra, blk *out_cw, uint64_t nkeys, cudaStream_t stream) {
  const uint64_t threads = nkeys * 32;
  const unsigned blocks = unsigned((threads + 255) / 256);
  relayout_kernel<<<blocks, 256, 0, stream>>>(scheme, in_bits, ncw, cws, cw_s, cw_v, extra, out_cw, nkeys);
  return cudaGetLastError();
}

// ---- Grotto: prefix XOR over leaf bits (grotto_dcf.cuh:160-162) ---------------------------------------------
constexpr int kScanThreads = 256;
constexpr uint64_t kScanTile = uint64_t(kScanThreads) * 16;  // bytes per tile

__device__ __forceinline__ uint32_t word_prefix(uint32_t w) {  // byte j := XOR of bytes 0..j (values 0/1)
  w ^= w << 8;
  w ^= w << 16;
  return w;
}

// Tile-local inclusive scan, in place.  Requires 16-byte aligned rows and len % 16 == 0.
__global__ void scan_tile_kernel(uint8_t *ys, uint64_t len) {
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  uint8_t *row = ys + uint64_t(blockIdx.y) * len;
  const uint64_t off = uint64_t(blockIdx.x) * kScanTile + uint64_t(threadIdx.x) * 16;
  const bool active = off < len;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (active) v = *reinterpret_cast<const uint4 *>(row + off);
  v.x = word_prefix(v.x);
  v.y = word_prefix(v.y) ^ ((v.x >> 24) * 0x01010101u);
  v.z = word_prefix(v.z) ^ ((v.y >> 24) * 0x01010101u);
  v.w = word_prefix(v.w) ^ ((v.z >> 24) * 0x01010101u);
  const uint32_t tot = v.w >> 24;  // parity of this thread's 16 bytes
  // exclusive scan of thread parities across the block
  const uint32_t ball = __ballot_sync(0xffffffffu, tot & 1u);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t carry = __popc(ball & ((1u << lane) - 1u)) & 1u;
  if (lane == 0) warp_tot[wid] = __popc(ball) & 1u;
  __syncthreads();
  for (int w = 0; w < wid; ++w) carry ^= warp_tot[w];
  const uint32_t m = carry * 0x01010101u;
  if (active) {
    v.x ^= m; v.y ^= m; v.z ^= m; v.w ^= m;
    *reinterpret_cast<uint4 *>(row + off) = v;
  }
}

// Carry of each tile = XOR of the totals (= last bytes) of the tiles before it; parked in bit 1 of
// the tile's first byte so that no scratch buffer is needed.
__global__ void scan_carry_kernel(uint8_t *ys, uint64_t len, uint64_t tiles) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t running;
  uint8_t *row = ys + uint64_t(blockIdx.x) * len;
  if (threadIdx.x == 0) running = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint64_t base = 0; base < tiles; base += blockDim.x) {
    const uint64_t t = base + threadIdx.x;
    uint32_t tot = 0;
    if (t < tiles) {
      const uint64_t last = (t + 1) * kScanTile < len ? (t + 1) * kScanTile - 1 : len - 1;
      tot = row[last] & 1u;
    }
    const uint32_t ball = __ballot_sync(0xffffffffu, tot);
    uint32_t carry = __popc(ball & ((1u << lane) - 1u)) & 1u;
    if (lane == 0) warp_tot[wid] = __popc(ball) & 1u;
    __syncthreads();
    for (int w = 0; w < wid; ++w) carry ^= warp_tot[w];
    carry ^= running;
    if (t < tiles && carry) row[t * kScanTile] |= 2u;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) running = carry ^ tot;
    __syncthreads();
  }
}

__global__ void scan_apply_kernel(uint8_t *ys, uint64_t len) {
  __shared__ uint32_t carry_s;
  uint8_t *row = ys + uint64_t(blockIdx.y) * len;
  const uint64_t tile0 = uint64_t(blockIdx.x) * kScanTile;
  if (threadIdx.x == 0) carry_s = (row[tile0] >> 1) & 1u;
  __syncthreads();
  const uint32_t carry = carry_s;
  const uint64_t off = tile0 + uint64_t(threadIdx.x) * 16;
  if (off >= len || !carry) return;
  uint4 v = *reinterpret_cast<const uint4 *>(row + off);
  const uint32_t m = 0x01010101u;
  v.x ^= m; v.y ^= m; v.z ^= m; v.w ^= m;
  if (threadIdx.x == 0) v.x &= ~2u;  // drop the parked carry bit
  *reinterpret_cast<uint4 *>(row + off) = v;
}

// Small / unaligned rows: one thread per row.
__global__ void scan_serial_kernel(uint8_t *ys, uint64_t nkeys, uint64_t len) {
  const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= nkeys) return;
  uint8_t *row = ys + k * len;
  uint8_t acc = 0;
  for (uint64_t x = 0; x < len; ++x) {
    acc ^= row[x] & 1u;
    row[x] = acc;
  }
}

cudaError_t launch_prefix_xor(uint8_t *ys, uint64_t nkeys, uint64_t len, cudaStream_t stream) {
  if ((len & 15u) || (reinterpret_cast<uintptr_t>(ys) & 15u) || len < 256) {
    scan_serial_kernel<<<unsigned((nkeys + 127) / 128), 128, 0, stream>>>(ys, nkeys, len);
    return cudaGetLastError();
  }
  const uint64_t tiles = (len + kScanTile - 1) / kScanTile;
  const dim3 grid = dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(nkeys), 1);
  scan_tile_kernel<<<grid, kScanThreads, 0, stream>>>(ys, len);
  if (tiles > 1) {
    scan_carry_kernel<<<unsigned(nkeys), 1024, 0, stream>>>(ys, len, tiles);
    scan_apply_kernel<<<grid, kScanThreads, 0, stream>>>(ys, len);
  }
  return cudaGetLastError();
}

// ---- Grotto parity tree / lookup ------------------------------------------------------------------------------
__global__ void parity_level_kernel(uint8_t *tree, uint64_t first, uint64_t count) {
  const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const uint64_t j = first + i;
  tree[j] = tree[2 * j + 1] ^ tree[2 * j + 2];
}
cudaError_t launch_parity_level(uint8_t *tree, int level, cudaStream_t stream) {
  const uint64_t count = uint64_t(1) << level, first = count - 1;
  parity_level_kernel<<<unsigned((count + 255) / 256), 256, 0, stream>>>(tree, first, count);
  return cudaGetLastError();
}

__global__ void grotto_lookup_kernel(const uint8_t *pt, const uint8_t *xs, uint8_t *ys, uint64_t nkeys, int n,
    int in_bytes) {
  const uint64_t k = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= nkeys) return;
  const uint64_t N = uint64_t(1) << n;
  const uint8_t *tree = pt + k * (2 * N - 1);
  // e = x + 1 in the arithmetic of `In` (grotto_dcf.cuh:118); n <= 31 so only the low 64 bits matter,
  // but the wrap to 0 must be detected on the full In width
  uint64_t lo = 0, hi = 0;
  const uint8_t *xp = xs + k * uint64_t(in_bytes);
  for (int b = 0; b < in_bytes && b < 8; ++b) lo |= uint64_t(xp[b]) << (8 * b);
  for (int b = 8; b < in_bytes; ++b) hi |= uint64_t(xp[b]) << (8 * (b - 8));
  uint64_t e = lo + 1;
  if (in_bytes < 8) e &= (uint64_t(1) << (8 * in_bytes)) - 1;
  bool wrapped = (e == 0);
  if (in_bytes == 16) {
    const uint64_t ehi = hi + (e == 0 ? 1 : 0);
    wrapped = (e == 0 && ehi == 0);
    if (ehi != 0) {  // e >= 2^64 > N: the reference walks bits of e below n only
      wrapped = false;
    }
  }
  if (wrapped || e == N) {
    ys[k] = tree[0];
    return;
  }
  uint8_t pi = 0;
  uint64_t cur = 0;
  for (int i = 0; i < n; ++i) {
    if ((e >> (n - 1 - i)) & 1) {
      pi ^= tree[2 * cur + 1];
      cur = 2 * cur + 2;
    } else {
      cur = 2 * cur + 1;
    }
  }
  ys[k] = pi;
}
cudaError_t launch_grotto_lookup(const uint8_t *pt, const uint8_t *xs, uint8_t *ys, uint64_t nkeys, int in_bits,
    int in_bytes, cudaStream_t stream) {
  grotto_lookup_kernel<<<unsigned((nkeys + 127) / 128), 128, 0, stream>>>(pt, xs, ys, nkeys, in_bits, in_bytes);
  return cudaGetLastError();
}

// ---- issue-rate microbenchmarks ----------------------------------------------------------------------------------
constexpr int kMbIters = 2048;
constexpr int kMbIlp = 8;

template <int KIND>
__global__ void __launch_bounds__(1024, 1) microbench_kernel(uint32_t *sink, uint32_t b, uint32_t c) {
  __shared__ uint32_t tbl[32 * 64];
  for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) tbl[i] = i * 2654435761u;
  __syncthreads();
  uint32_t a[kMbIlp];
#pragma unroll
  for (int j = 0; j < kMbIlp; ++j) a[j] = threadIdx.x * 7u + j;
  const uint32_t saddr = static_cast<uint32_t>(__cvta_generic_to_shared(tbl)) + (threadIdx.x & 31u) * 4u;
#pragma unroll 1
  for (int it = 0; it < kMbIters; ++it) {
#pragma unroll
    for (int j = 0; j < kMbIlp; ++j) {
      if (KIND == 0) {
        asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
      } else if (KIND == 1) {
        asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
      } else if (KIND == 2) {
        if (j & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
        else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[j]) : "r"(b), "r"(c));
      } else if (KIND == 3) {
        uint32_t v;
        // ld.volatile: ptxas must not merge or hoist the loads (a plain ld.shared of a loop-invariant
        // address is hoisted out of the loop and the kernel then measures the XOR instead).
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr + uint32_t(j) * 128u));
        a[j] ^= v;  // conflict-free: lane l reads bank l; 8 independent loads per iteration
      } else if (KIND == 4) {
        asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(a[j]) : "r"(b));
      } else if (KIND == 5) {
        asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
      } else {
        if (j & 1) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[j]) : "r"(b), "r"(c));
        else asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(a[j]) : "r"(b));
      }
    }
  }
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < kMbIlp; ++j) acc ^= a[j];
  if (acc == 0x12345678u) sink[0] = acc;  // keeps the chains alive
}

int run_microbench(int kind, double *ops_per_s) {
  if (kind < 0 || kind > 6) return FSSB200_EINVAL;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  uint32_t *sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, 4);
  if (e != cudaSuccess) return int(e);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  const dim3 grid = dim3(static_cast<unsigned>(sms) * 4, 1, 1), block = dim3(1024, 1, 1);
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0);
    switch (kind) {
      case 0: microbench_kernel<0><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      case 1: microbench_kernel<1><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      case 2: microbench_kernel<2><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      case 3: microbench_kernel<3><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      case 4: microbench_kernel<4><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      case 5: microbench_kernel<5><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
      default: microbench_kernel<6><<<grid, block>>>(sink, 0x9e3779b9u, 0x7f4a7c15u); break;
    }
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(sink);
  if (e != cudaSuccess) return int(e);
  const double ops = double(grid.x) * block.x * double(kMbIters) * kMbIlp;
  *ops_per_s = ops / (double(best) * 1e-3);
  return 0;
}

}  // namespace fssb200
