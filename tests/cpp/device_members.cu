// SPDX-License-Identifier: Apache-2.0
// TEST (nvcc, B200): the device-callable members added for the reference's src/bench_gpu.cu -- `prg::Aes128Soft` on tables this
// file's kernels put in shared memory, `hash::Blake3`, `Vdpf::Gen` / `Vdpf::Eval` -- called per thread inside this file's own
// __global__ kernels exactly as that benchmark does (bench_gpu.cu:105-137, 172-205), and compared bit for bit with the
// batched members (the precompiled sm_100a kernels behind the C ABI, which the -m gpu suite pins against the oracle).
#include <cstdio>
#include <cstring>
#include <vector>
#include <fss/dpf.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/hash/blake3.cuh>
#include <fss/prg/aes128_mmo_soft.cuh>
#include <fss/prg/chacha.cuh>
#include <fss/vdpf.cuh>

static int g_fail = 0;
#define EXPECT(cond, what)                                        \
  do {                                                            \
    if (!(cond)) {                                                \
      std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); \
      ++g_fail;                                                   \
    }                                                             \
  } while (0)
#define CUDA_OK(x) EXPECT((x) == cudaSuccess, #x)

constexpr int kKeys = 1000, kBits = 20;
using Group = fss::group::Uint<uint64_t>;
using SoftPrg = fss::prg::Aes128Soft<2>;
using SoftDpf = fss::Dpf<kBits, Group, SoftPrg, uint>;
using VdpfT = fss::Vdpf<kBits, Group, fss::prg::ChaCha<2>, fss::hash::Blake3, fss::hash::Blake3, uint>;

__constant__ uint8_t kKeysDev[2][16] = {{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}, {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1}};
static const uint8_t kKeysHost[2][16] = {{1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16}, {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1}};
__device__ int kNonceDev[2] = {0x12345678, int(0x9abcdef0u)};
static int kNonceHost[2] = {0x12345678, int(0x9abcdef0u)};
__device__ const int4 kIvDev[2] = {{0x11111111, 0x22222222, 0x33333333, 0x44444444}, {0x55555555, 0x66666666, 0x77777777, int(0x88888888u)}};
static const int4 kIvHost[2] = {{0x11111111, 0x22222222, 0x33333333, 0x44444444}, {0x55555555, 0x66666666, 0x77777777, int(0x88888888u)}};

__device__ void FillTables(uint32_t *te0, uint8_t *sbox) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    te0[i] = fss::prg::aes_detail::ComputeTe0(uint8_t(i));
    sbox[i] = fss::prg::aes_detail::Sbox(uint8_t(i));
  }
  __syncthreads();
}
__global__ void SoftGen(SoftDpf::Cw *cws, const int4 *s0s, const uint *alphas, const int4 *betas, int n) {
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  FillTables(te0, sbox);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  SoftPrg prg(kKeysDev, te0, sbox);
  SoftDpf dpf{prg};
  const int4 s[2] = {s0s[2 * tid], s0s[2 * tid + 1]};
  dpf.Gen(cws + size_t(tid) * (kBits + 1), s, alphas[tid], betas[tid]);
}
__global__ void SoftEval(int4 *ys, bool party, const int4 *seeds, const SoftDpf::Cw *cws, const uint *xs, int n) {
  __shared__ uint32_t te0[256];
  __shared__ uint8_t sbox[256];
  FillTables(te0, sbox);
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  SoftPrg prg(kKeysDev, te0, sbox);
  SoftDpf dpf{prg};
  ys[tid] = dpf.Eval(party, seeds[tid], cws + size_t(tid) * (kBits + 1), xs[tid]);
}
__global__ void VGen(VdpfT::Cw *cws, cuda::std::array<int4, 4> *cs, int4 *ocws, int *status, const int4 *s0s, const uint *alphas,
    const int4 *betas, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  fss::prg::ChaCha<2> prg(kNonceDev);
  fss::hash::Blake3 xh{cuda::std::span<const int4, 2>(kIvDev)}, h{cuda::std::span<const int4, 2>(kIvDev)};
  VdpfT vdpf{prg, xh, h};
  const int4 s[2] = {s0s[2 * tid], s0s[2 * tid + 1]};
  status[tid] = vdpf.Gen(cws + size_t(tid) * kBits, cs[tid], ocws[tid], cuda::std::span<const int4, 2>(s), alphas[tid], betas[tid]);
}
__global__ void VEval(int4 *ys, cuda::std::array<int4, 4> *pis, bool party, const int4 *seeds, const VdpfT::Cw *cws,
    const cuda::std::array<int4, 4> *cs, const int4 *ocws, const uint *xs, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  fss::prg::ChaCha<2> prg(kNonceDev);
  fss::hash::Blake3 xh{cuda::std::span<const int4, 2>(kIvDev)}, h{cuda::std::span<const int4, 2>(kIvDev)};
  VdpfT vdpf{prg, xh, h};
  int4 y;
  pis[tid] = vdpf.Eval(party, seeds[tid], cuda::std::span<const VdpfT::Cw>(cws + size_t(tid) * kBits, kBits),
      cuda::std::span<const int4, 4>(cs[tid].data(), 4), ocws[tid], xs[tid], y);
  ys[tid] = y;
}

template <typename T>
static T *Dev(size_t n, const T *h = nullptr) {
  T *d = nullptr;
  CUDA_OK(cudaMalloc(&d, sizeof(T) * n));
  if (h) CUDA_OK(cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice));
  else CUDA_OK(cudaMemset(d, 0, sizeof(T) * n));
  return d;
}
template <typename T>
static std::vector<T> Host(const T *d, size_t n) {
  std::vector<T> h(n);
  CUDA_OK(cudaMemcpy(h.data(), d, sizeof(T) * n, cudaMemcpyDeviceToHost));
  return h;
}
// compares the information-carrying bytes of Cw arrays ({s, tr}; the padding of the two producers may differ)
template <typename Cw>
static bool SameCws(const std::vector<Cw> &a, const std::vector<Cw> &b, bool with_flag) {
  bool same = a.size() == b.size();
  for (size_t i = 0; i < a.size() && same; ++i) same = std::memcmp(&a[i].s, &b[i].s, 16) == 0 && (!with_flag || a[i].tr == b[i].tr);
  return same;
}

int main() {
  const int n = kKeys;
  uint64_t st = 99;
  auto next = [&] {
    st = st * 6364136223846793005ULL + 1442695040888963407ULL;
    return uint32_t(st >> 33);
  };
  std::vector<int4> s0s(2 * n), betas(n), seeds0(n), seeds1(n);
  std::vector<uint> alphas(n), xs(n);
  for (int k = 0; k < n; ++k) {
    for (int j = 0; j < 2; ++j) s0s[2 * k + j] = int4{int(next()), int(next()), int(next()), int(next() & ~1u)};
    betas[k] = int4{int(next()), int(next()), int(next()), int(next() & ~1u)};
    alphas[k] = next() & ((1u << kBits) - 1);
    xs[k] = k % 3 ? next() & ((1u << kBits) - 1) : alphas[k];
    seeds0[k] = s0s[2 * k];
    seeds1[k] = s0s[2 * k + 1];
  }
  int4 *d_s0s = Dev(2 * n, s0s.data()), *d_betas = Dev(n, betas.data()), *d_seed[2] = {Dev(n, seeds0.data()), Dev(n, seeds1.data())};
  uint *d_alphas = Dev(n, alphas.data()), *d_xs = Dev(n, xs.data());
  const int grid = (n + 127) / 128;

  {  // ---- Dpf with Aes128Soft inside kernels vs the batched members (same keys, same function: the precompiled AES kernels)
    uint32_t te0[256];
    uint8_t sbox[256];
    fss::prg::aes_detail::InitTe0(te0);
    fss::prg::aes_detail::InitSbox(sbox);
    SoftDpf dpf{SoftPrg(kKeysHost, te0, sbox)};
    SoftDpf::Cw *d_cws = Dev<SoftDpf::Cw>(size_t(n) * (kBits + 1)), *d_want = Dev<SoftDpf::Cw>(size_t(n) * (kBits + 1));
    SoftGen<<<grid, 128>>>(d_cws, d_s0s, d_alphas, d_betas, n);
    CUDA_OK(cudaDeviceSynchronize());
    dpf.GenBatch(d_s0s, d_alphas, d_betas, d_want, n);
    CUDA_OK(cudaDeviceSynchronize());
    EXPECT(SameCws(Host(d_cws, size_t(n) * (kBits + 1)), Host(d_want, size_t(n) * (kBits + 1)), false), "Dpf<Aes128Soft>::Gen in a kernel == GenBatch");
    int4 *d_y = Dev<int4>(n), *d_yw = Dev<int4>(n);
    for (int party = 0; party < 2; ++party) {
      SoftEval<<<grid, 128>>>(d_y, party != 0, d_seed[party], d_want, d_xs, n);
      CUDA_OK(cudaDeviceSynchronize());
      dpf.EvalBatch(party != 0, d_seed[party], d_want, d_xs, d_yw, n);
      CUDA_OK(cudaDeviceSynchronize());
      const auto a = Host(d_y, n), b = Host(d_yw, n);
      EXPECT(std::memcmp(a.data(), b.data(), sizeof(int4) * n) == 0, "Dpf<Aes128Soft>::Eval in a kernel == EvalBatch");
    }
    std::printf("Dpf<20, Uint64, Aes128Soft<2>> per thread on shared-memory tables: Gen, Eval x 2 parties, %d keys\n", n);
  }
  {  // ---- Vdpf (ChaCha + Blake3) inside kernels vs the batched members
    fss::prg::ChaCha<2> prg(kNonceHost);
    fss::hash::Blake3 xh{cuda::std::span<const int4, 2>(kIvHost)}, h{cuda::std::span<const int4, 2>(kIvHost)};
    VdpfT vdpf{prg, xh, h};
    VdpfT::Cw *d_cws = Dev<VdpfT::Cw>(size_t(n) * kBits), *d_want = Dev<VdpfT::Cw>(size_t(n) * kBits);
    auto *d_cs = Dev<cuda::std::array<int4, 4>>(n), *d_cs_w = Dev<cuda::std::array<int4, 4>>(n);
    int4 *d_ocws = Dev<int4>(n), *d_ocws_w = Dev<int4>(n);
    int *d_status = Dev<int>(n);
    int32_t *d_status_w = Dev<int32_t>(n);
    VGen<<<grid, 128>>>(d_cws, d_cs, d_ocws, d_status, d_s0s, d_alphas, d_betas, n);
    CUDA_OK(cudaDeviceSynchronize());
    vdpf.GenBatch(d_s0s, d_alphas, d_betas, d_want, reinterpret_cast<int4 *>(d_cs_w), d_ocws_w, d_status_w, n);
    CUDA_OK(cudaDeviceSynchronize());
    EXPECT(SameCws(Host(d_cws, size_t(n) * kBits), Host(d_want, size_t(n) * kBits), true), "Vdpf::Gen in a kernel: cws == GenBatch");
    {
      const auto a = Host(d_cs, n), b = Host(d_cs_w, n);
      EXPECT(std::memcmp(a.data(), b.data(), 64 * size_t(n)) == 0, "Vdpf::Gen in a kernel: cs == GenBatch");
      const auto c = Host(d_ocws, n), d = Host(d_ocws_w, n);
      EXPECT(std::memcmp(c.data(), d.data(), 16 * size_t(n)) == 0, "Vdpf::Gen in a kernel: ocw == GenBatch");
      const auto e = Host(d_status, n);
      const auto f = Host(d_status_w, n);
      bool same = true;
      for (int k = 0; k < n; ++k) same &= e[k] == f[k];
      EXPECT(same, "Vdpf::Gen in a kernel: status == GenBatch");
    }
    int4 *d_y = Dev<int4>(n), *d_yw = Dev<int4>(n);
    auto *d_pi = Dev<cuda::std::array<int4, 4>>(n), *d_pi_w = Dev<cuda::std::array<int4, 4>>(n);
    for (int party = 0; party < 2; ++party) {
      VEval<<<grid, 128>>>(d_y, d_pi, party != 0, d_seed[party], d_want, d_cs_w, d_ocws_w, d_xs, n);
      CUDA_OK(cudaDeviceSynchronize());
      vdpf.EvalBatch(party != 0, d_seed[party], d_want, reinterpret_cast<const int4 *>(d_cs_w), d_ocws_w, d_xs, d_yw, reinterpret_cast<int4 *>(d_pi_w), n);
      CUDA_OK(cudaDeviceSynchronize());
      const auto a = Host(d_y, n), b = Host(d_yw, n);
      EXPECT(std::memcmp(a.data(), b.data(), 16 * size_t(n)) == 0, "Vdpf::Eval in a kernel: y == EvalBatch");
      const auto c = Host(d_pi, n), d = Host(d_pi_w, n);
      EXPECT(std::memcmp(c.data(), d.data(), 64 * size_t(n)) == 0, "Vdpf::Eval in a kernel: pi == EvalBatch");
    }
    std::printf("Vdpf<20, Uint64, ChaCha<2>, Blake3, Blake3> per thread: Gen, Eval x 2 parties, %d keys\n", n);
  }
  if (g_fail == 0) std::printf("device members: all checks passed\n");
  return g_fail ? 1 : 0;
}
