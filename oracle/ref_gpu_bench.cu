// SPDX-License-Identifier: Apache-2.0
//
// TEST / MEASUREMENT INFRASTRUCTURE ONLY -- never linked or imported by fss_b200/.
//
// ref_gpu_bench: runs the UNMODIFIED reference GPU kernels (headers under /root/reference/include, compiled
// here for sm_100a from where they lie) on the same inputs as libfssb200.so, so that
//   (1) the reference's own device code is a second, GPU-side parity oracle (its kernels have no test in the
//       reference: "point_eval_gpu.cuh kernels have no test at all"), and
//   (2) "beat this generic kernel on the same B200"  is a measured statement.
//
// Usage: ref_gpu_bench <mode> <nkeys> <dir> [iters]
//   reads  <dir>/seeds.bin  int4[nkeys]           party-0 seeds
//          <dir>/cws.bin    Cw[nkeys][ncw]        key-major, the reference's own 32-byte Cw structs
//          <dir>/ocws.bin   int4[nkeys]           (Half-Tree only)
//          <dir>/xs.bin     In[nkeys]             (point modes)
//   writes <dir>/ys_ref.bin                       (point: int4[nkeys]; evalall: int4[nkeys][2^n])
//   prints one JSON line: {"mode":..., "ms":best, "ms_avg":..., "regs":..., "units":...}
//
// What is timed is what the reference's own benchmark times (src/bench_gpu.cu:48-68,302-307): the eval launch
// only, CUDA events, after one warm-up; the level-major relayout is a one-time pre-pass outside the loop
// (src/bench_gpu.cu:618-619).  Kernels that wrap `Scheme::Eval` per thread follow the shape of the reference's
// bench kernels (src/bench_gpu.cu:85-95,125-138): one thread per key, 256 threads per block.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/eval_all_gpu.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/half_tree_dpf.cuh>
#include <fss/point_eval_gpu.cuh>
#include <fss/prg/aes128_mmo_soft.cuh>
#include <fss/prg/chacha.cuh>

#define CK(x)                                                                                   \
  do {                                                                                          \
    cudaError_t e_ = (x);                                                                       \
    if (e_ != cudaSuccess) {                                                                    \
      fprintf(stderr, "cuda error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                                  \
    }                                                                                           \
  } while (0)

// fixtures shared with fss_b200.context (DEFAULT_AES_KEYS / DEFAULT_CHACHA_NONCE / DEFAULT_HASH_KEY)
__constant__ int c_nonce[2] = {0x12345678, static_cast<int>(0x9abcdef0u)};
__constant__ uint8_t c_aes_keys[4][16] = {
    {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16},
    {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1},
    {1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8},
    {8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1},
};
static const int4 kHashKey = {0x12345678, static_cast<int>(0x9abcdef0u), 0x0fedcba9, static_cast<int>(0x87654321u)};

using Bytes = fss::group::Bytes;
using U64 = fss::group::Uint<uint64_t>;
// The 2^127 literal cannot sit in a __global__ template argument list (nvcc stub generator),
// so every kernel below is a plain function and names the group inside its body.
using U127 = fss::group::Uint<__uint128_t, (static_cast<__uint128_t>(1) << 127)>;

constexpr int kBs = 256;

// ---- per-thread wrappers around the library's own Eval (key-major Cw: the "naive GPU" baseline) ---------------------
__device__ __forceinline__ void soft_tables(uint32_t *te0, uint8_t *sbox) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    te0[i] = fss::prg::aes_detail::ComputeTe0(static_cast<uint8_t>(i));
    sbox[i] = fss::prg::aes_detail::Sbox(static_cast<uint8_t>(i));
  }
  __syncthreads();
}

__global__ void k_dpf32_chacha_naive(int4 *ys, const int4 *seeds, const void *cws_, const uint32_t *xs, int nkeys) {
  using S = fss::Dpf<32, Bytes, fss::prg::ChaCha<2>, uint>;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nkeys) return;
  fss::prg::ChaCha<2> prg(c_nonce);
  S dpf{prg};
  ys[tid] = dpf.Eval(false, seeds[tid], static_cast<const S::Cw *>(cws_) + size_t(tid) * 33, xs[tid]);
}

__global__ void k_dpf32_aes_naive(int4 *ys, const int4 *seeds, const void *cws_, const uint32_t *xs, int nkeys) {
  using S = fss::Dpf<32, Bytes, fss::prg::Aes128Soft<2>, uint>;
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  soft_tables(te0, sbox);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nkeys) return;
  fss::prg::Aes128Soft<2> prg(c_aes_keys, te0, sbox);
  S dpf{prg};
  ys[tid] = dpf.Eval(false, seeds[tid], static_cast<const S::Cw *>(cws_) + size_t(tid) * 33, xs[tid]);
}

__global__ void k_dcf64_u127_aes_naive(int4 *ys, const int4 *seeds, const void *cws_, const uint64_t *xs, int nkeys) {
  using S = fss::Dcf<64, U127, fss::prg::Aes128Soft<4>, uint64_t>;
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  soft_tables(te0, sbox);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nkeys) return;
  fss::prg::Aes128Soft<4> prg(c_aes_keys, te0, sbox);
  S dcf{prg};
  ys[tid] = dcf.Eval(false, seeds[tid], static_cast<const S::Cw *>(cws_) + size_t(tid) * 65, xs[tid]);
}

__global__ void k_dcf64_u127_chacha_naive(int4 *ys, const int4 *seeds, const void *cws_, const uint64_t *xs,
                                          int nkeys) {
  using S = fss::Dcf<64, U127, fss::prg::ChaCha<4>, uint64_t>;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nkeys) return;
  fss::prg::ChaCha<4> prg(c_nonce);
  S dcf{prg};
  ys[tid] = dcf.Eval(false, seeds[tid], static_cast<const S::Cw *>(cws_) + size_t(tid) * 65, xs[tid]);
}

__global__ void k_ht32_aes_naive(int4 *ys, const int4 *seeds, const void *cws_, const int4 *ocws, const uint32_t *xs,
                                 int nkeys, int4 hash_key) {
  using S = fss::HalfTreeDpf<32, Bytes, fss::prg::Aes128Soft<1>, uint>;
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  soft_tables(te0, sbox);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= nkeys) return;
  fss::prg::Aes128Soft<1> prg(c_aes_keys, te0, sbox);
  S dpf{prg, hash_key};
  ys[tid] = dpf.Eval(false, seeds[tid], static_cast<const S::Cw *>(cws_) + size_t(tid) * 32, ocws[tid], xs[tid]);
}

// ---- harness ---------------------------------------------------------------------------------------------------------
static std::vector<char> slurp(const std::string &path, size_t want) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) {
    fprintf(stderr, "cannot open %s\n", path.c_str());
    exit(2);
  }
  std::vector<char> buf(want);
  const size_t got = fread(buf.data(), 1, want, f);
  fclose(f);
  if (got != want) {
    fprintf(stderr, "%s: wanted %zu bytes, got %zu\n", path.c_str(), want, got);
    exit(2);
  }
  return buf;
}
template <class T>
static T *upload(const std::string &path, size_t count) {
  std::vector<char> h = slurp(path, count * sizeof(T));
  T *d;
  CK(cudaMalloc(&d, h.size()));
  CK(cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice));
  return d;
}
static void download(const std::string &path, const void *d, size_t bytes) {
  std::vector<char> h(bytes);
  CK(cudaMemcpy(h.data(), d, bytes, cudaMemcpyDeviceToHost));
  FILE *f = fopen(path.c_str(), "wb");
  if (!f || fwrite(h.data(), 1, bytes, f) != bytes) {
    fprintf(stderr, "cannot write %s\n", path.c_str());
    exit(2);
  }
  fclose(f);
}

template <class Fn>
static void timed(const char *mode, double units, int iters, Fn &&fn) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  fn();  // warm-up
  CK(cudaDeviceSynchronize());
  double best = 1e30, sum = 0;
  for (int i = 0; i < iters; ++i) {
    CK(cudaEventRecord(a));
    fn();
    CK(cudaPeekAtLastError());
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
    sum += ms;
  }
  printf("{\"mode\": \"%s\", \"ms\": %.6f, \"ms_avg\": %.6f, \"iters\": %d, \"units\": %.0f, \"units_per_s\": %.6e}\n",
         mode, best, sum / iters, iters, units, units / (sum / iters * 1e-3));
}

template <int N, int Z, int B1>
static void run_dpf_evalall_chacha(const char *mode, int nkeys, const std::string &dir, int iters, bool dump) {
  using S = fss::Dpf<N, Bytes, fss::prg::ChaCha<2>, uint>;
  int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
  auto *cws = upload<typename S::Cw>(dir + "/cws.bin", size_t(nkeys) * (N + 1));
  int4 *ys;
  CK(cudaMalloc(&ys, sizeof(int4) * (size_t(nkeys) << N)));
  int *nonce;
  CK(cudaGetSymbolAddress(reinterpret_cast<void **>(&nonce), c_nonce));
  fss::prg::ChaCha<2> prg(nonce);
  S dpf{prg};
  timed(mode, double(nkeys) * double(size_t(1) << N), iters,
        [&] { fss::gpu::DpfEvalAllGpuBatch<Z, B1, 256>(false, seeds, cws, nkeys, ys, dpf); });
  if (dump) download(dir + "/ys_ref.bin", ys, sizeof(int4) * (size_t(nkeys) << N));
}

template <int N, int Z, int B1>
static void run_ht_evalall_chacha(const char *mode, int nkeys, const std::string &dir, int iters, bool dump) {
  using S = fss::HalfTreeDpf<N, Bytes, fss::prg::ChaCha<1>, uint>;
  int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
  auto *cws = upload<typename S::Cw>(dir + "/cws.bin", size_t(nkeys) * N);
  int4 *ocws = upload<int4>(dir + "/ocws.bin", nkeys);
  int4 *ys;
  CK(cudaMalloc(&ys, sizeof(int4) * (size_t(nkeys) << N)));
  int *nonce;
  CK(cudaGetSymbolAddress(reinterpret_cast<void **>(&nonce), c_nonce));
  fss::prg::ChaCha<1> prg(nonce);
  S dpf{prg, kHashKey};
  timed(mode, double(nkeys) * double(size_t(1) << N), iters,
        [&] { fss::gpu::HalfTreeDpfEvalAllGpuBatch<Z, B1, 256>(false, seeds, cws, ocws, nkeys, ys, dpf); });
  if (dump) download(dir + "/ys_ref.bin", ys, sizeof(int4) * (size_t(nkeys) << N));
}

int main(int argc, char **argv) {
  if (argc < 4) {
    fprintf(stderr, "usage: %s <mode> <nkeys> <dir> [iters]\n", argv[0]);
    return 2;
  }
  const std::string mode = argv[1], dir = argv[3];
  const int nkeys = atoi(argv[2]);
  const int iters = argc > 4 ? atoi(argv[4]) : 10;
  const int blocks = (nkeys + kBs - 1) / kBs;
  int *nonce;
  CK(cudaGetSymbolAddress(reinterpret_cast<void **>(&nonce), c_nonce));

  if (mode == "dpf32_chacha_naive" || mode == "dpf32_aes_naive" || mode == "dpf32_chacha_point") {
    using S = fss::Dpf<32, Bytes, fss::prg::ChaCha<2>, uint>;  // Cw layout does not depend on the PRG
    int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
    auto *cws = upload<S::Cw>(dir + "/cws.bin", size_t(nkeys) * 33);
    uint32_t *xs = upload<uint32_t>(dir + "/xs.bin", nkeys);
    int4 *ys;
    CK(cudaMalloc(&ys, sizeof(int4) * nkeys));
    if (mode == "dpf32_chacha_naive") {
      timed(mode.c_str(), nkeys, iters, [&] { k_dpf32_chacha_naive<<<blocks, kBs>>>(ys, seeds, cws, xs, nkeys); });
    } else if (mode == "dpf32_aes_naive") {
      timed(mode.c_str(), nkeys, iters, [&] { k_dpf32_aes_naive<<<blocks, kBs>>>(ys, seeds, cws, xs, nkeys); });
    } else {
      int4 *cw_s, *out_cw;
      uint32_t *extra;
      CK(cudaMalloc(&cw_s, sizeof(int4) * 32 * size_t(nkeys)));
      CK(cudaMalloc(&out_cw, sizeof(int4) * nkeys));
      CK(cudaMalloc(&extra, sizeof(uint32_t) * nkeys));
      fss::gpu::DpfRelayoutGpu<32, Bytes, fss::prg::ChaCha<2>, uint>(cws, nkeys, cw_s, extra, out_cw);
      CK(cudaDeviceSynchronize());
      fss::prg::ChaCha<2> prg(nonce);
      S dpf{prg};
      timed(mode.c_str(), nkeys, iters, [&] {
        fss::gpu::DpfEvalPointGpu<4, 32, Bytes, fss::prg::ChaCha<2>, uint>(false, seeds, cw_s, extra, out_cw, xs, ys,
                                                                            nkeys, dpf);
      });
    }
    CK(cudaDeviceSynchronize());
    download(dir + "/ys_ref.bin", ys, sizeof(int4) * nkeys);
  } else if (mode == "dcf64_u127_aes_naive" || mode == "dcf64_u127_chacha_naive") {
    using S = fss::Dcf<64, U127, fss::prg::ChaCha<4>, uint64_t>;
    int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
    auto *cws = upload<S::Cw>(dir + "/cws.bin", size_t(nkeys) * 65);
    uint64_t *xs = upload<uint64_t>(dir + "/xs.bin", nkeys);
    int4 *ys;
    CK(cudaMalloc(&ys, sizeof(int4) * nkeys));
    if (mode == "dcf64_u127_aes_naive")
      timed(mode.c_str(), nkeys, iters, [&] { k_dcf64_u127_aes_naive<<<blocks, kBs>>>(ys, seeds, cws, xs, nkeys); });
    else
      timed(mode.c_str(), nkeys, iters, [&] { k_dcf64_u127_chacha_naive<<<blocks, kBs>>>(ys, seeds, cws, xs, nkeys); });
    CK(cudaDeviceSynchronize());
    download(dir + "/ys_ref.bin", ys, sizeof(int4) * nkeys);
  } else if (mode == "dcf32_u64_chacha_point") {
    using S = fss::Dcf<32, U64, fss::prg::ChaCha<4>, uint>;
    int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
    auto *cws = upload<S::Cw>(dir + "/cws.bin", size_t(nkeys) * 33);
    uint32_t *xs = upload<uint32_t>(dir + "/xs.bin", nkeys);
    int4 *ys, *cw_s, *cw_v, *out_cw;
    CK(cudaMalloc(&ys, sizeof(int4) * nkeys));
    CK(cudaMalloc(&cw_s, sizeof(int4) * 32 * size_t(nkeys)));
    CK(cudaMalloc(&cw_v, sizeof(int4) * 32 * size_t(nkeys)));
    CK(cudaMalloc(&out_cw, sizeof(int4) * nkeys));
    fss::gpu::DcfRelayoutGpu<32, U64, fss::prg::ChaCha<4>, uint>(cws, nkeys, cw_s, cw_v, out_cw);
    CK(cudaDeviceSynchronize());
    fss::prg::ChaCha<4> prg(nonce);
    S dcf{prg};
    timed(mode.c_str(), nkeys, iters, [&] {
      fss::gpu::DcfEvalPointGpu<4, 32, U64, fss::prg::ChaCha<4>, uint>(false, seeds, cw_s, cw_v, out_cw, xs, ys, nkeys,
                                                                        dcf);
    });
    CK(cudaDeviceSynchronize());
    download(dir + "/ys_ref.bin", ys, sizeof(int4) * nkeys);
  } else if (mode == "ht32_chacha_point" || mode == "ht32_aes_naive") {
    using S = fss::HalfTreeDpf<32, Bytes, fss::prg::ChaCha<1>, uint>;
    int4 *seeds = upload<int4>(dir + "/seeds.bin", nkeys);
    auto *cws = upload<S::Cw>(dir + "/cws.bin", size_t(nkeys) * 32);
    int4 *ocws = upload<int4>(dir + "/ocws.bin", nkeys);
    uint32_t *xs = upload<uint32_t>(dir + "/xs.bin", nkeys);
    int4 *ys;
    CK(cudaMalloc(&ys, sizeof(int4) * nkeys));
    if (mode == "ht32_aes_naive") {
      timed(mode.c_str(), nkeys, iters,
            [&] { k_ht32_aes_naive<<<blocks, kBs>>>(ys, seeds, cws, ocws, xs, nkeys, kHashKey); });
    } else {
      int4 *cw_s;
      uint32_t *extra;
      CK(cudaMalloc(&cw_s, sizeof(int4) * 32 * size_t(nkeys)));
      CK(cudaMalloc(&extra, sizeof(uint32_t) * nkeys));
      fss::gpu::HalfTreeDpfRelayoutGpu<32, Bytes, fss::prg::ChaCha<1>, uint>(cws, nkeys, cw_s, extra);
      CK(cudaDeviceSynchronize());
      fss::prg::ChaCha<1> prg(nonce);
      S dpf{prg, kHashKey};
      timed(mode.c_str(), nkeys, iters, [&] {
        fss::gpu::HalfTreeDpfEvalPointGpu<4, 32, Bytes, fss::prg::ChaCha<1>, uint>(false, seeds, cw_s, extra, ocws, xs,
                                                                                    ys, nkeys, dpf);
      });
    }
    CK(cudaDeviceSynchronize());
    download(dir + "/ys_ref.bin", ys, sizeof(int4) * nkeys);
  } else if (mode == "dpf_evalall20_chacha") {  // the reference benchmark's own shape (src/bench_gpu.cu:578)
    run_dpf_evalall_chacha<20, 17, 9>(mode.c_str(), nkeys, dir, iters, true);
  } else if (mode == "dpf_evalall28_chacha") {  // BASELINE configs[3] domain; nkeys <= 7 (int indexing, eval_all_gpu.cuh:287)
    run_dpf_evalall_chacha<28, 20, 12>(mode.c_str(), nkeys, dir, iters, false);
  } else if (mode == "ht_evalall20_chacha") {
    run_ht_evalall_chacha<20, 17, 9>(mode.c_str(), nkeys, dir, iters, true);
  } else if (mode == "ht_evalall28_chacha") {
    run_ht_evalall_chacha<28, 20, 12>(mode.c_str(), nkeys, dir, iters, false);
  } else {
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
