// SPDX-License-Identifier: Apache-2.0
// TEST (nvcc, B200): the header shim with (1) USER-DEFINED plugins -- the reference's class templates instantiated
// with a group and a PRG the library has never heard of (user_plugin.hpp), through the reference's member signatures:
// the section printed by plugin_user_main.inc must equal the reference's own output (tests/golden/plugin_user_v1.txt);
// and (2) DEVICE-CALLABLE members: `dpf.Gen` / `dpf.Eval` called per thread inside this file's own __global__
// kernels, the usage of README.md:198-242 / samples/dpf_dcf_gpu.cu:51-82, for the user plugins and for the built-in
// ChaCha PRG, checked against the batched members (for ChaCha: against the precompiled kernels behind the C ABI).
#include "user_plugin.hpp"

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/half_tree_dpf.cuh>
#include <fss/prg/chacha.cuh>

#include "plugin_user_main.inc"

static int g_fail = 0;
#define EXPECT(cond, what)                                          \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__);   \
      ++g_fail;                                                     \
    }                                                               \
  } while (0)

template <typename T>
static T *ToDevice(const T *h, size_t n) {
  T *d = nullptr;
  cudaMalloc(&d, sizeof(T) * n);
  cudaMemcpy(d, h, sizeof(T) * n, cudaMemcpyHostToDevice);
  return d;
}

// ---- user plugins inside the user's own kernel ----
__global__ void UserEvalKernel(UserDpf dpf, bool party, const int4 *seeds, const UserDpf::Cw *cws, const uint16_t *xs,
                               int4 *ys, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  ys[tid] = dpf.Eval(party, seeds[tid], cws + tid * 13, xs[tid]);
}
__global__ void UserGenKernel(UserDcfLt dcf, UserDcfLt::Cw *cws, const int4 *s0s, const uint32_t *alphas, const int4 *betas, int n) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= n) return;
  const int4 s[2] = {s0s[2 * tid], s0s[2 * tid + 1]};
  dcf.Gen(cws + tid * 21, s, alphas[tid], betas[tid]);
}

static void SectionUserDevice(const UserKeys &K) {
  const int n = kUserKeys;
  UserDpf dpf{MixPrg<2>{0xdecafbadu}};
  std::vector<uint16_t> xs(n);
  std::vector<int4> want0(n), want1(n);
  for (int k = 0; k < n; ++k) {
    xs[k] = uint16_t(k % 2 ? K.alpha[k] & 0xfff : (K.alpha[k] + 5) & 0xfff);
    want0[k] = dpf.Eval(false, K.s0[k], &K.dpf_cws[size_t(k) * 13], xs[k]);   // single-key member (one-thread launch)
    want1[k] = dpf.Eval(true, K.s1[k], &K.dpf_cws[size_t(k) * 13], xs[k]);
  }
  int4 *d_s0 = ToDevice(K.s0.data(), n), *d_s1 = ToDevice(K.s1.data(), n), *d_y;
  UserDpf::Cw *d_cws = ToDevice(K.dpf_cws.data(), size_t(n) * 13);
  uint16_t *d_xs = ToDevice(xs.data(), n);
  cudaMalloc(&d_y, 16 * n);
  std::vector<int4> got(n);
  // batched member -> generic kernel instantiated here with the user's types
  dpf.EvalBatch(false, d_s0, d_cws, d_xs, d_y, n);
  cudaMemcpy(got.data(), d_y, 16 * n, cudaMemcpyDeviceToHost);
  EXPECT(std::memcmp(got.data(), want0.data(), 16 * n) == 0, "user plugins: EvalBatch == Eval");
  // the user's own kernel calling the member per thread
  UserEvalKernel<<<1, 32>>>(dpf, true, d_s1, d_cws, d_xs, d_y, n);
  cudaMemcpy(got.data(), d_y, 16 * n, cudaMemcpyDeviceToHost);
  EXPECT(cudaGetLastError() == cudaSuccess, "UserEvalKernel launch");
  EXPECT(std::memcmp(got.data(), want1.data(), 16 * n) == 0, "user plugins: dpf.Eval inside a user kernel == host member");
  // Gen inside a user kernel == the host member's keys (already checked against the reference's golden output)
  UserDcfLt dcf{MixPrg<4>{0x0badcafeu}};
  std::vector<int4> s0s(2 * n);
  std::vector<uint32_t> al(n);
  for (int k = 0; k < n; ++k) {
    s0s[2 * k] = K.s0[k];
    s0s[2 * k + 1] = K.s1[k];
    al[k] = K.alpha[k] & 0xfffff;
  }
  int4 *d_s0s = ToDevice(s0s.data(), 2 * n), *d_beta = ToDevice(K.beta.data(), n);
  uint32_t *d_al = ToDevice(al.data(), n);
  UserDcfLt::Cw *d_dc;
  cudaMalloc(&d_dc, sizeof(UserDcfLt::Cw) * 21 * n);
  UserGenKernel<<<1, 32>>>(dcf, d_dc, d_s0s, d_al, d_beta, n);
  std::vector<UserDcfLt::Cw> dc(size_t(n) * 21);
  cudaMemcpy(dc.data(), d_dc, sizeof(UserDcfLt::Cw) * dc.size(), cudaMemcpyDeviceToHost);
  EXPECT(std::memcmp(dc.data(), K.dcf_cws.data(), sizeof(UserDcfLt::Cw) * dc.size()) == 0,
         "user plugins: dcf.Gen inside a user kernel == host member");
  // full domain through the batched member, against per-point evaluation
  std::vector<int4> all(size_t(1) << 12);
  dpf.EvalAll(true, K.s1[3], &K.dpf_cws[3 * 13], all.data());
  bool same = true;
  for (int x = 0; x < 4096; x += 97) {
    const int4 y = dpf.Eval(true, K.s1[3], &K.dpf_cws[3 * 13], uint16_t(x));
    same = same && std::memcmp(&y, &all[size_t(x)], 16) == 0;
  }
  EXPECT(same, "user plugins: EvalAll == Eval");
  std::printf("shim: user-defined plugins, batched + device-callable members: %s\n", g_fail ? "FAILED" : "ok");
}

// ---- the built-in ChaCha PRG inside the user's own kernels (samples/dpf_dcf_gpu.cu:40-60 restated) ----
constexpr int kBits = 32;
constexpr int kN = 300;
using CcPrg = fss::prg::ChaCha<2>;
using CcGroup = fss::group::Uint<uint64_t>;
using CcDpf = fss::Dpf<kBits, CcGroup, CcPrg, uint32_t>;
__constant__ int kNonce[2] = {0x12345678, int(0x9abcdef0u)};
static const int kNonceHost[2] = {0x12345678, int(0x9abcdef0u)};

__global__ void CcGenKernel(CcDpf::Cw *cws, const int4 *seeds, const uint32_t *alphas, const int4 *betas) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= kN) return;
  CcPrg prg(kNonce);
  CcDpf dpf{prg};
  const int4 s[2] = {seeds[tid * 2], seeds[tid * 2 + 1]};
  dpf.Gen(cws + tid * (kBits + 1), s, alphas[tid], betas[tid]);
}
__global__ void CcEvalKernel(int4 *ys, bool party, const int4 *seeds, const CcDpf::Cw *cws, const uint32_t *xs) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid >= kN) return;
  CcPrg prg(kNonce);
  CcDpf dpf{prg};
  ys[tid] = dpf.Eval(party, seeds[tid], cws + tid * (kBits + 1), xs[tid]);
}

static void SectionBuiltinDevice() {
  const int before = g_fail;
  unsigned st = 77;
  std::vector<int4> s0s(2 * kN), seeds0(kN), seeds1(kN), betas(kN);
  std::vector<uint32_t> alphas(kN), xs(kN);
  for (int i = 0; i < kN; ++i) {
    s0s[2 * i] = seeds0[i] = RandBlock(st);
    s0s[2 * i + 1] = seeds1[i] = RandBlock(st);
    betas[i] = RandBlock(st);
    alphas[i] = Lcg(st);
    xs[i] = i % 3 ? Lcg(st) : alphas[i];
  }
  int4 *d_s0s = ToDevice(s0s.data(), 2 * kN), *d_seeds0 = ToDevice(seeds0.data(), kN), *d_seeds1 = ToDevice(seeds1.data(), kN),
       *d_betas = ToDevice(betas.data(), kN), *d_y, *d_yb;
  uint32_t *d_al = ToDevice(alphas.data(), kN), *d_xs = ToDevice(xs.data(), kN);
  CcDpf::Cw *d_cws, *d_cws_b;
  cudaMalloc(&d_cws, sizeof(CcDpf::Cw) * (kBits + 1) * kN);
  cudaMalloc(&d_cws_b, sizeof(CcDpf::Cw) * (kBits + 1) * kN);
  cudaMalloc(&d_y, 16 * kN);
  cudaMalloc(&d_yb, 16 * kN);
  // keys: per-thread Gen in a user kernel vs the precompiled batched Gen kernel (C ABI)
  CcGenKernel<<<(kN + 127) / 128, 128>>>(d_cws, d_s0s, d_al, d_betas);
  CcPrg prg(kNonceHost);
  CcDpf dpf{prg};
  dpf.GenBatch(d_s0s, d_al, d_betas, d_cws_b, kN);
  std::vector<CcDpf::Cw> c1(size_t(kBits + 1) * kN), c2(c1.size());
  cudaMemcpy(c1.data(), d_cws, sizeof(CcDpf::Cw) * c1.size(), cudaMemcpyDeviceToHost);
  cudaMemcpy(c2.data(), d_cws_b, sizeof(CcDpf::Cw) * c2.size(), cudaMemcpyDeviceToHost);
  EXPECT(cudaGetLastError() == cudaSuccess, "CcGenKernel launch");
  EXPECT(std::memcmp(c1.data(), c2.data(), sizeof(CcDpf::Cw) * c1.size()) == 0,
         "ChaCha: dpf.Gen inside a user kernel == precompiled batched Gen");
  std::vector<int4> y(kN), yb(kN), y1(kN);
  for (int party = 0; party < 2; ++party) {
    CcEvalKernel<<<(kN + 127) / 128, 128>>>(d_y, party != 0, party ? d_seeds1 : d_seeds0, d_cws, d_xs);
    dpf.EvalBatch(party != 0, party ? d_seeds1 : d_seeds0, d_cws, d_xs, d_yb, kN);
    cudaMemcpy(y.data(), d_y, 16 * kN, cudaMemcpyDeviceToHost);
    cudaMemcpy(yb.data(), d_yb, 16 * kN, cudaMemcpyDeviceToHost);
    EXPECT(std::memcmp(y.data(), yb.data(), 16 * kN) == 0, "ChaCha: dpf.Eval inside a user kernel == precompiled batched Eval");
    if (party == 0) y1 = y;
  }
  int bad = 0;
  for (int i = 0; i < kN; ++i) {
    const int4 sum = (CcGroup::From(y1[i]) + CcGroup::From(y[i])).Into();
    const int4 want = xs[i] == alphas[i] ? CcGroup::From(betas[i]).Into() : int4{0, 0, 0, 0};
    bad += std::memcmp(&sum, &want, 16) != 0;
  }
  EXPECT(bad == 0, "ChaCha: reconstruction of the in-kernel evaluations");
  std::printf("shim: built-in ChaCha PRG, device-callable members: %s\n", g_fail == before ? "ok" : "FAILED");
}

int main() {
  try {
    UserKeys keys;
    SectionReferenceSurface(keys);   // the lines of tests/golden/plugin_user_v1.txt
    SectionUserDevice(keys);
    SectionBuiltinDevice();
  } catch (const std::exception &e) {
    std::printf("FAIL exception: %s\n", e.what());
    return 2;
  }
  std::printf(g_fail ? "plugin test: %d failure(s)\n" : "plugin test: all checks passed\n", g_fail);
  return g_fail ? 1 : 0;
}
