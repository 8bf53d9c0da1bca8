// SPDX-License-Identifier: Apache-2.0
// TEST: the header-only C++ surface (include/fss/*.cuh) used the way the reference's samples use it
// (flow of samples/dpf_dcf_cpu.cu, half_tree_dpf_cpu.cu, grotto_dcf_cpu.cu and src/bench_gpu.cu restated,
// not copied), compiled with plain g++ -std=c++20 and linked against libfssb200.so.  Prints "FAIL ..." and
// returns non-zero on any mismatch; known answers are the SURVEY.md section 8c vectors.
#include <cstdio>
#include <cstring>
#include <vector>

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/eval_all_gpu.cuh>
#include <fss/grotto_dcf.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/half_tree_dpf.cuh>
#include <fss/hash/blake3.cuh>
#include <fss/point_eval_gpu.cuh>
#include <fss/prg/aes128_mmo.cuh>
#include <fss/prg/chacha.cuh>
#include <fss/hash/sha256.cuh>
#include <fss/vdpf.cuh>

static int g_fail = 0;
#define EXPECT(cond, what)                     \
  do {                                         \
    if (!(cond)) {                             \
      std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); \
      ++g_fail;                                \
    }                                          \
  } while (0)

static bool Eq(int4 a, int4 b) { return std::memcmp(&a, &b, 16) == 0; }
static int4 Hex(unsigned x, unsigned y, unsigned z, unsigned w) { return int4{int(x), int(y), int(z), int(w)}; }

template <typename Group>
static int4 Add(int4 a, int4 b) { return (Group::From(a) + Group::From(b)).Into(); }

static unsigned char k0[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static unsigned char k1[16] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1};
static unsigned char k2[16] = {1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8};
static unsigned char k3[16] = {8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1};
static const int4 kSeeds[2] = {{0x11111111, 0x22222222, 0x33333333, 0x44444440},
                               {0x55555555, 0x66666666, 0x77777777, int(0x88888880u)}};
static const int4 kBeta = {7, 0, 0, 0};
static const int4 kZero = {0, 0, 0, 0};

static void DpfN8() {
  using Group = fss::group::Bytes;
  using Prg = fss::prg::Aes128Mmo<2>;
  using Dpf = fss::Dpf<8, Group, Prg, uint8_t>;
  const unsigned char *keys[2] = {k0, k1};
  auto ctxs = Prg::CreateCtxs(keys);
  Prg prg(ctxs);
  Dpf dpf{prg};
  Dpf::Cw cws[9];
  dpf.Gen(cws, kSeeds, 42, kBeta);
  EXPECT(Eq(cws[0].s, Hex(0x84863059, 0xb57c0062, 0xc0a8015f, 0x3520c9e8)) && cws[0].tr, "dpf n8 cw[0]");
  EXPECT(Eq(cws[8].s, Hex(0x0b17cf59, 0xcd54e225, 0xef2cd79f, 0x30ad9cf8)), "dpf n8 cw[8]");
  int4 y0 = dpf.Eval(false, kSeeds[0], cws, 42), y1 = dpf.Eval(true, kSeeds[1], cws, 42);
  EXPECT(Eq(y0, Hex(0xa81892ad, 0xad2e583e, 0x272d1483, 0xd2bfa094)), "dpf n8 y0(42)");
  EXPECT(Eq(y1, Hex(0xa81892aa, 0xad2e583e, 0x272d1483, 0xd2bfa094)), "dpf n8 y1(42)");
  EXPECT(Eq(Add<Group>(y0, y1), kBeta), "dpf n8 reconstruct at alpha");
  EXPECT(Eq(dpf.Eval(false, kSeeds[0], cws, 100), Hex(0x7e2be32f, 0x4508eb0c, 0x537f8e89, 0x39b1bd96)), "dpf n8 y0(100)");
  int4 a0[256], a1[256];
  dpf.EvalAll(false, kSeeds[0], cws, a0);
  dpf.EvalAll(true, kSeeds[1], cws, a1);
  int bad = 0;
  for (int i = 0; i < 256; ++i) bad += !Eq(Add<Group>(a0[i], a1[i]), i == 42 ? kBeta : kZero);
  EXPECT(bad == 0, "dpf n8 EvalAll reconstruct");
  EXPECT(Eq(a0[100], dpf.Eval(false, kSeeds[0], cws, 100)), "dpf n8 EvalAll == Eval");
  auto g = prg.Gen(kSeeds[0]);
  EXPECT(Eq(g[0], Hex(0x1423e6d2, 0x60533602, 0x813b3fbc, 0x412b31dc)), "Aes128Mmo<2>.Gen block 0");
  Prg::FreeCtxs(ctxs);
}

static void DcfN64() {
  using Group = fss::group::Uint<__uint128_t, (static_cast<__uint128_t>(1) << 127)>;
  using Prg = fss::prg::Aes128Mmo<4>;
  using Dcf = fss::Dcf<64, Group, Prg, uint64_t>;
  const unsigned char *keys[4] = {k0, k1, k2, k3};
  auto ctxs = Prg::CreateCtxs(keys);
  Prg prg(ctxs);
  Dcf dcf{prg};
  std::vector<Dcf::Cw> cws(65);
  dcf.Gen(cws.data(), kSeeds, 42, kBeta);
  EXPECT(Eq(cws[64].v, Hex(0x6675d222, 0xa63e84c9, 0xd6f446d0, 0x985772b0)), "dcf n64 cw[64].v");
  int4 y0 = dcf.Eval(false, kSeeds[0], cws.data(), 10), y1 = dcf.Eval(true, kSeeds[1], cws.data(), 10);
  EXPECT(Eq(y0, Hex(0x865812cd, 0xe91b6533, 0xfa4486b0, 0x9bc4ab6c)), "dcf n64 y0(10)");
  EXPECT(Eq(y1, Hex(0x79a7ed3a, 0x16e49acc, 0x05bb794f, 0x643b5492)), "dcf n64 y1(10)");
  EXPECT(Eq(Add<Group>(y0, y1), kBeta), "dcf n64 x < alpha");
  EXPECT(Eq(Add<Group>(dcf.Eval(false, kSeeds[0], cws.data(), 42), dcf.Eval(true, kSeeds[1], cws.data(), 42)), kZero),
         "dcf n64 x == alpha");
  EXPECT(Eq(dcf.Eval(false, kSeeds[0], cws.data(), 1ull << 40), Hex(0x25494a62, 0xf745e3eb, 0x2e8feeb2, 0x48222122)),
         "dcf n64 y0(2^40)");
  Prg::FreeCtxs(ctxs);
}

static void HalfTreeAndGrotto() {
  using Group = fss::group::Bytes;
  using Prg1 = fss::prg::Aes128Mmo<1>;
  using Ht = fss::HalfTreeDpf<8, Group, Prg1, uint8_t>;
  const unsigned char *keys1[1] = {k0};
  auto c1 = Prg1::CreateCtxs(keys1);
  Prg1 prg1(c1);
  Ht ht{prg1, {0x12345678, int(0x9abcdef0u), 0x13572468, int(0x2468ace0u)}};
  Ht::Cw cws[8];
  int4 ocw;
  ht.Gen(cws, ocw, kSeeds, 42, kBeta);
  EXPECT(Eq(ocw, Hex(0xb3205fb1, 0xa28dd128, 0x8cc01d96, 0xda93f6e8)), "halftree ocw");
  int4 y0 = ht.Eval(false, kSeeds[0], cws, ocw, 42), y1 = ht.Eval(true, kSeeds[1], cws, ocw, 42);
  EXPECT(Eq(y0, Hex(0xe58e28f0, 0xdfccd22d, 0x550c9b5d, 0x12125964)), "halftree y0(42)");
  EXPECT(Eq(Add<Group>(y0, y1), kBeta), "halftree reconstruct");
  int4 a0[256], a1[256];
  ht.EvalAll(false, kSeeds[0], cws, ocw, a0);
  ht.EvalAll(true, kSeeds[1], cws, ocw, a1);
  int bad = 0;
  for (int i = 0; i < 256; ++i) bad += !Eq(Add<Group>(a0[i], a1[i]), i == 42 ? kBeta : kZero);
  EXPECT(bad == 0, "halftree EvalAll reconstruct");
  Prg1::FreeCtxs(c1);

  using Prg2 = fss::prg::Aes128Mmo<2>;
  using Gr = fss::GrottoDcf<8, Prg2, uint8_t>;
  const unsigned char *keys2[2] = {k0, k1};
  auto c2 = Prg2::CreateCtxs(keys2);
  Prg2 prg2(c2);
  Gr gr{prg2};
  Gr::Cw gcws[9];
  gr.Gen(gcws, kSeeds, 42);
  bool s0[256], s1[256];
  gr.EvalAll(false, kSeeds[0], gcws, s0);
  gr.EvalAll(true, kSeeds[1], gcws, s1);
  bad = 0;
  for (int x = 0; x < 256; ++x) bad += (s0[x] ^ s1[x]) != (42 <= x);
  EXPECT(bad == 0, "grotto EvalAll reconstructs 1[alpha <= x]");
  const char *want = "1011111110111000";
  for (int x = 0; x < 16; ++x) EXPECT(s0[x] == (want[x] == '1'), "grotto party-0 share bits (survey KAT)");
  std::vector<char> tree(511);
  Gr::ParityTree pt{reinterpret_cast<bool *>(tree.data()), false};
  gr.Preprocess(pt, kSeeds[0], gcws);
  for (int x : {0, 41, 42, 200, 255}) EXPECT(Gr::Eval(pt, uint8_t(x)) == s0[x], "grotto Preprocess+Eval == EvalAll");
  Prg2::FreeCtxs(c2);
}

// Batched device path in the style of src/bench_gpu.cu: keys generated on the device, relayout, point eval.
static void BatchedChaCha() {
  using Group = fss::group::Uint<uint64_t>;
  using Prg = fss::prg::ChaCha<2>;
  using Dpf = fss::Dpf<32, Group, Prg, uint32_t>;
  static const int nonce[2] = {0x12345678, int(0x9abcdef0u)};
  Prg prg(nonce);
  Dpf dpf{prg};
  const int n = 5000;
  std::vector<int4> s0s(2 * n), betas(n), seeds0(n), seeds1(n);
  std::vector<uint32_t> alphas(n), xs(n);
  unsigned state = 42;
  auto rnd = [&] { state = state * 1664525u + 1013904223u; return int(state ^ (state >> 13)); };
  for (int i = 0; i < n; ++i) {
    s0s[2 * i] = {rnd(), rnd(), rnd(), rnd() & ~1};
    s0s[2 * i + 1] = {rnd(), rnd(), rnd(), rnd() & ~1};
    seeds0[i] = s0s[2 * i];
    seeds1[i] = s0s[2 * i + 1];
    betas[i] = {rnd(), rnd(), 0, 0};
    alphas[i] = unsigned(rnd());
    xs[i] = (i % 4 == 0) ? alphas[i] : unsigned(rnd());
  }
  int4 *d_s0s, *d_betas, *d_seeds0, *d_seeds1, *d_y0, *d_y1, *d_cw_s, *d_out_cw, *d_y0lm;
  uint32_t *d_alphas, *d_xs, *d_extra;
  Dpf::Cw *d_cws;
  cudaMalloc(&d_s0s, 32 * n); cudaMalloc(&d_betas, 16 * n); cudaMalloc(&d_seeds0, 16 * n); cudaMalloc(&d_seeds1, 16 * n);
  cudaMalloc(&d_y0, 16 * n); cudaMalloc(&d_y1, 16 * n); cudaMalloc(&d_y0lm, 16 * n); cudaMalloc(&d_alphas, 4 * n);
  cudaMalloc(&d_xs, 4 * n); cudaMalloc(&d_cws, sizeof(Dpf::Cw) * 33 * size_t(n)); cudaMalloc(&d_cw_s, 16 * 32 * size_t(n));
  cudaMalloc(&d_extra, 4 * n); cudaMalloc(&d_out_cw, 16 * n);
  cudaMemcpy(d_s0s, s0s.data(), 32 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_betas, betas.data(), 16 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_seeds0, seeds0.data(), 16 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_seeds1, seeds1.data(), 16 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_alphas, alphas.data(), 4 * n, cudaMemcpyHostToDevice);
  cudaMemcpy(d_xs, xs.data(), 4 * n, cudaMemcpyHostToDevice);
  dpf.GenBatch(d_s0s, d_alphas, d_betas, d_cws, n);
  dpf.EvalBatch(false, d_seeds0, d_cws, d_xs, d_y0, n);
  dpf.EvalBatch(true, d_seeds1, d_cws, d_xs, d_y1, n);
  fss::gpu::DpfRelayoutGpu<32, Group, Prg, uint32_t>(d_cws, n, d_cw_s, d_extra, d_out_cw);
  fss::gpu::DpfEvalPointGpu(false, d_seeds0, d_cw_s, d_extra, d_out_cw, d_xs, d_y0lm, n, dpf);
  std::vector<int4> y0(n), y1(n), y0lm(n);
  cudaMemcpy(y0.data(), d_y0, 16 * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(y1.data(), d_y1, 16 * n, cudaMemcpyDeviceToHost);
  cudaMemcpy(y0lm.data(), d_y0lm, 16 * n, cudaMemcpyDeviceToHost);
  int bad = 0, badlm = 0;
  for (int i = 0; i < n; ++i) {
    const int4 want = xs[i] == alphas[i] ? Group::From(betas[i]).Into() : kZero;
    bad += !Eq(Add<Group>(y0[i], y1[i]), want);
    badlm += !Eq(y0[i], y0lm[i]);
  }
  EXPECT(bad == 0, "batched ChaCha DPF reconstruct (5000 keys)");
  EXPECT(badlm == 0, "level-major path == key-major path");
  // single-key host member on a device-generated key agrees with the batched result
  std::vector<Dpf::Cw> h_cws(33);
  cudaMemcpy(h_cws.data(), d_cws + 33 * 7, sizeof(Dpf::Cw) * 33, cudaMemcpyDeviceToHost);
  EXPECT(Eq(dpf.Eval(false, seeds0[7], h_cws.data(), xs[7]), y0[7]), "Dpf::Eval == EvalBatch");
  // multi-device members: two shards of the key batch, here both on the current device so that
  // the sample runs on a one-GPU box -- per-device pointer arrays, one stream per shard, no collective
  {
    int dev = 0;
    cudaGetDevice(&dev);
    const int devices[2] = {dev, dev};
    size_t b0, e0, b1, e1;
    fssb200_key_shard(size_t(n), 0, 2, &b0, &e0);
    fssb200_key_shard(size_t(n), 1, 2, &b1, &e1);
    EXPECT(b0 == 0 && e0 == b1 && e1 == size_t(n), "key_shard covers the batch");
    cudaStream_t st[2];
    cudaStreamCreate(&st[0]);
    cudaStreamCreate(&st[1]);
    int4 *d_ym;
    cudaMalloc(&d_ym, 16 * n);
    const int4 *sd[2] = {d_seeds0 + b0, d_seeds0 + b1};
    const Dpf::Cw *cw[2] = {d_cws + 33 * b0, d_cws + 33 * b1};
    const uint32_t *xx[2] = {d_xs + b0, d_xs + b1};
    int4 *yy[2] = {d_ym + b0, d_ym + b1};
    const size_t nk[2] = {e0 - b0, e1 - b1};
    dpf.EvalBatchMulti(false, 2, devices, sd, cw, xx, yy, nk, st);
    dpf.SyncMulti(2, devices, st);
    std::vector<int4> ym(n), yh(n);
    cudaMemcpy(ym.data(), d_ym, 16 * n, cudaMemcpyDeviceToHost);
    EXPECT(std::memcmp(ym.data(), y0.data(), 16 * size_t(n)) == 0, "EvalBatchMulti (2 shards) == EvalBatch");
    // host arrays of the whole batch over the same two shards
    std::vector<Dpf::Cw> all_cws(33 * size_t(n));
    cudaMemcpy(all_cws.data(), d_cws, sizeof(Dpf::Cw) * all_cws.size(), cudaMemcpyDeviceToHost);
    dpf.EvalBatchHostMulti(false, 2, devices, seeds0.data(), all_cws.data(), xs.data(), yh.data(), size_t(n));
    EXPECT(std::memcmp(yh.data(), y0.data(), 16 * size_t(n)) == 0, "EvalBatchHostMulti (2 shards) == EvalBatch");
    cudaStreamDestroy(st[0]);
    cudaStreamDestroy(st[1]);
    cudaFree(d_ym);
  }
  // full-domain on the device for a small domain
  using Dpf12 = fss::Dpf<12, Group, Prg, uint32_t>;
  Dpf12 d12{prg};
  std::vector<Dpf12::Cw> c12(13);
  d12.Gen(c12.data(), kSeeds, 1234, kBeta);
  Dpf12::Cw *d_c12; int4 *d_all;
  cudaMalloc(&d_c12, sizeof(Dpf12::Cw) * 13); cudaMalloc(&d_all, 16 << 12);
  cudaMemcpy(d_c12, c12.data(), sizeof(Dpf12::Cw) * 13, cudaMemcpyHostToDevice);
  fss::gpu::DpfEvalAllGpu(false, kSeeds[0], d_c12, d_all, d12);
  std::vector<int4> all(1 << 12);
  cudaMemcpy(all.data(), d_all, 16 << 12, cudaMemcpyDeviceToHost);
  EXPECT(Eq(all[1234], d12.Eval(false, kSeeds[0], c12.data(), 1234)), "DpfEvalAllGpu == Eval at alpha");
  EXPECT(Eq(all[77], d12.Eval(false, kSeeds[0], c12.data(), 77)), "DpfEvalAllGpu == Eval elsewhere");
  cudaDeviceSynchronize();
}

// Verifiable DPF (flow of samples/vdpf_cpu.cu restated; Blake3 for both hashes as in src/bench_gpu.cu)
static void VdpfN8() {
  using Group = fss::group::Bytes;
  using Prg = fss::prg::Aes128Mmo<2>;
  using H = fss::hash::Blake3;
  using Vdpf = fss::Vdpf<8, Group, Prg, H, H, uint8_t>;
  const unsigned char *keys[2] = {k0, k1};
  auto ctxs = Prg::CreateCtxs(keys);
  Prg prg(ctxs);
  const int4 iv0[2] = {{0x12345678, int(0x9abcdef0u), 0x13572468, 0x2468ace0}, {1, 2, 3, 4}};
  const int4 iv1[2] = {{int(0x0fedcba9u), int(0x87654321u), 0x2468ace0, 0x13572468}, {5, 6, 7, 8}};
  H xor_hash{cuda::std::span<const int4, 2>(iv0, 2)}, hash{cuda::std::span<const int4, 2>(iv1, 2)};
  Vdpf vdpf{prg, xor_hash, hash};
  Vdpf::Cw cws[8];
  cuda::std::array<int4, 4> cs;
  int4 ocw, seeds[2];
  int ret, r = 0;
  do {  // Gen returns 1 when the seeds hit t0 == t1: the dealer resamples (vdpf.cuh:169)
    seeds[0] = {0x11111111 + r, 0x22222222 + r, 0x33333333 + r, 0x44444440 + r};
    seeds[1] = {0x55555555 + r, 0x66666666 + r, 0x77777777 + r, int(0x88888880u) + r};
    ret = vdpf.Gen(cws, cs, ocw, cuda::std::span<const int4, 2>(seeds, 2), 42, kBeta);
    ++r;
  } while (ret != 0 && r < 64);
  EXPECT(ret == 0, "vdpf Gen succeeds within 64 seed pairs");
  const auto cwspan = cuda::std::span<const Vdpf::Cw>(cws, 8);
  const auto csspan = cuda::std::span<const int4, 4>(cs);
  int4 y0, y1;
  auto pt0 = vdpf.Eval(false, seeds[0], cwspan, csspan, ocw, 42, y0);
  auto pt1 = vdpf.Eval(true, seeds[1], cwspan, csspan, ocw, 42, y1);
  EXPECT(Eq(Add<Group>(y0, y1), kBeta), "vdpf reconstruct at alpha");
  EXPECT(std::memcmp(pt0.data(), pt1.data(), 64) == 0, "vdpf per-point hashes agree between the parties");
  cuda::std::array<int4, 4> pi0, pi1;
  vdpf.Prove(cuda::std::span<const cuda::std::array<int4, 4>>(&pt0, 1), csspan, pi0);
  vdpf.Prove(cuda::std::span<const cuda::std::array<int4, 4>>(&pt1, 1), csspan, pi1);
  EXPECT(Vdpf::Verify(cuda::std::span<const int4, 4>(pi0), cuda::std::span<const int4, 4>(pi1)), "vdpf Verify (one point)");
  // full domain: outputs reconstruct, proofs match, and the proof equals Prove over the per-point hashes in order
  int4 ys0[256], ys1[256];
  cuda::std::array<int4, 4> qa, qb;
  vdpf.EvalAll(false, seeds[0], cwspan, csspan, ocw, cuda::std::span<int4>(ys0), qa);
  vdpf.EvalAll(true, seeds[1], cwspan, csspan, ocw, cuda::std::span<int4>(ys1), qb);
  int bad = 0;
  for (int x = 0; x < 256; ++x) bad += !Eq(Add<Group>(ys0[x], ys1[x]), x == 42 ? kBeta : kZero);
  EXPECT(bad == 0, "vdpf EvalAll reconstruct");
  EXPECT(Vdpf::Verify(cuda::std::span<const int4, 4>(qa), cuda::std::span<const int4, 4>(qb)), "vdpf EvalAll Verify");
  std::vector<cuda::std::array<int4, 4>> pts(256);
  int4 yx;
  bool same = true;
  for (int x = 0; x < 256; ++x) {
    pts[x] = vdpf.Eval(false, seeds[0], cwspan, csspan, ocw, uint8_t(x), yx);
    same = same && Eq(yx, ys0[x]);
  }
  EXPECT(same, "vdpf EvalAll == Eval");
  cuda::std::array<int4, 4> qp;
  vdpf.Prove(cuda::std::span<const cuda::std::array<int4, 4>>(pts.data(), pts.size()), csspan, qp);
  EXPECT(std::memcmp(qp.data(), qa.data(), 64) == 0, "vdpf Prove(all points) == EvalAll proof");
  // a tampered key is caught
  Vdpf::Cw badcws[8];
  std::memcpy(badcws, cws, sizeof(cws));
  badcws[3].s.x ^= 4;
  vdpf.EvalAll(true, seeds[1], cuda::std::span<const Vdpf::Cw>(badcws, 8), csspan, ocw, cuda::std::span<int4>(ys1), qb);
  EXPECT(!Vdpf::Verify(cuda::std::span<const int4, 4>(qa), cuda::std::span<const int4, 4>(qb)), "vdpf Verify rejects a tampered key");
  // the hash plugin on its own: deterministic, IV-dependent, XorHash separates its two halves
  const int4 msg[4] = {{1, 2, 3, 4}, {5, 6, 7, 8}, {9, 10, 11, 12}, {13, 14, 15, 16}};
  auto h1 = hash.Hash(cuda::std::span<const int4, 4>(msg, 4)), h2 = xor_hash.Hash(cuda::std::span<const int4, 4>(msg, 4));
  EXPECT(!Eq(h1[0], h2[0]), "Blake3 output depends on the IV");
  auto x4 = xor_hash.Hash(cuda::std::tuple<int4, const int4>(msg[0], msg[1]));
  EXPECT(!Eq(x4[0], x4[2]), "Blake3 XorHash halves differ");
  Prg::FreeCtxs(ctxs);
}

// The same scheme with the reference's second hash plugin, fss::hash::Sha256 (hash/sha256.cuh; host-only there, on the device
// here), alone and mixed with Blake3, plus known answers of the plugin computed with Python's hashlib.
template <class HX, class HH>
static void VdpfHashes(const char *tag, HX xor_hash, HH hash) {
  using Group = fss::group::Uint<uint64_t>;
  using Prg = fss::prg::ChaCha<2>;
  using Vdpf = fss::Vdpf<10, Group, Prg, HX, HH, uint16_t>;
  static const int nonce[2] = {0x12345678, int(0x9abcdef0u)};
  Prg prg(nonce);
  Vdpf vdpf{prg, xor_hash, hash};
  typename Vdpf::Cw cws[10];
  cuda::std::array<int4, 4> cs;
  int4 ocw, seeds[2];
  int ret, r = 0;
  do {
    seeds[0] = {0x11111111 + r, 0x22222222 + r, 0x33333333 + r, 0x44444440 + r};
    seeds[1] = {0x55555555 + r, 0x66666666 + r, 0x77777777 + r, int(0x88888880u) + r};
    ret = vdpf.Gen(cws, cs, ocw, cuda::std::span<const int4, 2>(seeds, 2), 777, kBeta);
    ++r;
  } while (ret != 0 && r < 64);
  EXPECT(ret == 0, tag);
  const auto cwspan = cuda::std::span<const typename Vdpf::Cw>(cws, 10);
  const auto csspan = cuda::std::span<const int4, 4>(cs);
  std::vector<int4> ys0(1024), ys1(1024);
  cuda::std::array<int4, 4> qa, qb;
  vdpf.EvalAll(false, seeds[0], cwspan, csspan, ocw, cuda::std::span<int4>(ys0), qa);
  vdpf.EvalAll(true, seeds[1], cwspan, csspan, ocw, cuda::std::span<int4>(ys1), qb);
  int bad = 0;
  const int4 want = Group::From(kBeta).Into();
  for (int x = 0; x < 1024; ++x) bad += !Eq(Add<Group>(ys0[x], ys1[x]), x == 777 ? want : kZero);
  EXPECT(bad == 0, tag);
  EXPECT(Vdpf::Verify(cuda::std::span<const int4, 4>(qa), cuda::std::span<const int4, 4>(qb)), tag);
  int4 y;
  auto pt = vdpf.Eval(false, seeds[0], cwspan, csspan, ocw, 777, y);
  EXPECT(Eq(y, ys0[777]), tag);
  cws[4].s.y ^= 8;
  vdpf.EvalAll(true, seeds[1], cwspan, csspan, ocw, cuda::std::span<int4>(ys1), qb);
  EXPECT(!Vdpf::Verify(cuda::std::span<const int4, 4>(qa), cuda::std::span<const int4, 4>(qb)), tag);
  (void)pt;
}
// fss::gpu::VdpfRelayoutGpu + VdpfEvalPointGpu (point_eval_gpu.cuh:389-396, 513-526) against the batched key-major members
static void VdpfGpuFreeFunctions() {
  using Group = fss::group::Bytes;
  using Prg = fss::prg::Aes128Mmo<2>;
  using H = fss::hash::Blake3;
  using Vdpf = fss::Vdpf<20, Group, Prg, H, H, uint32_t>;
  const unsigned char *keys[2] = {k0, k1};
  auto ctxs = Prg::CreateCtxs(keys);
  Prg prg(ctxs);
  const int4 iv0[2] = {{1, 2, 3, 4}, {5, 6, 7, 8}}, iv1[2] = {{9, 10, 11, 12}, {13, 14, 15, 16}};
  Vdpf vdpf{prg, H{cuda::std::span<const int4, 2>(iv0, 2)}, H{cuda::std::span<const int4, 2>(iv1, 2)}};
  const int nk = 333;
  std::vector<int4> s0s(2 * nk), betas(nk), cs(4 * nk), ocws(nk), seeds(nk), y_a(nk), y_b(nk), pi_a(4 * nk), pi_b(4 * nk);
  std::vector<uint32_t> alphas(nk), xs(nk);
  std::vector<int32_t> status(nk);
  std::vector<Vdpf::Cw> cws(size_t(nk) * 20);
  for (int k = 0; k < nk; ++k) {
    s0s[2 * k] = {k * 7 + 1, k, 3, (k * 5) & ~1};
    s0s[2 * k + 1] = {k * 11 + 2, k, 9, (k * 3) & ~1};
    betas[k] = {k, 1, 2, 4};
    alphas[k] = (k * 2654435761u) & 0xfffffu;
    xs[k] = k % 3 ? ((k * 40503u) & 0xfffffu) : alphas[k];
    seeds[k] = s0s[2 * k + 1];
  }
  auto dev = [](const auto &v) {
    using T = typename std::decay_t<decltype(v)>::value_type;
    T *p = nullptr;
    cudaMalloc(&p, sizeof(T) * v.size());
    cudaMemcpy(p, v.data(), sizeof(T) * v.size(), cudaMemcpyHostToDevice);
    return p;
  };
  int4 *d_s0s = dev(s0s), *d_betas = dev(betas), *d_cs = dev(cs), *d_ocws = dev(ocws), *d_seeds = dev(seeds);
  int4 *d_ya = dev(y_a), *d_yb = dev(y_b), *d_pa = dev(pi_a), *d_pb = dev(pi_b);
  uint32_t *d_al = dev(alphas), *d_xs = dev(xs);
  int32_t *d_st = dev(status);
  Vdpf::Cw *d_cws = dev(cws);
  vdpf.GenBatch(d_s0s, d_al, d_betas, d_cws, d_cs, d_ocws, d_st, nk);
  vdpf.EvalBatch(true, d_seeds, d_cws, d_cs, d_ocws, d_xs, d_ya, d_pa, nk);
  int4 *d_cw_s = nullptr;
  uint32_t *d_extra = nullptr;
  cudaMalloc(&d_cw_s, sizeof(int4) * 20 * nk);
  cudaMalloc(&d_extra, sizeof(uint32_t) * nk);
  fss::gpu::VdpfRelayoutGpu<20, Group, Prg, H, H, uint32_t>(d_cws, nk, d_cw_s, d_extra);
  fss::gpu::VdpfEvalPointGpu(true, d_seeds, d_cw_s, d_extra, reinterpret_cast<const cuda::std::array<int4, 4> *>(d_cs), d_ocws,
                             d_xs, d_yb, d_pb, nk, vdpf);
  cudaDeviceSynchronize();
  cudaMemcpy(y_a.data(), d_ya, sizeof(int4) * nk, cudaMemcpyDeviceToHost);
  cudaMemcpy(y_b.data(), d_yb, sizeof(int4) * nk, cudaMemcpyDeviceToHost);
  cudaMemcpy(pi_a.data(), d_pa, sizeof(int4) * 4 * nk, cudaMemcpyDeviceToHost);
  cudaMemcpy(pi_b.data(), d_pb, sizeof(int4) * 4 * nk, cudaMemcpyDeviceToHost);
  EXPECT(std::memcmp(y_a.data(), y_b.data(), sizeof(int4) * nk) == 0, "VdpfEvalPointGpu == EvalBatch (shares)");
  EXPECT(std::memcmp(pi_a.data(), pi_b.data(), sizeof(int4) * 4 * nk) == 0, "VdpfEvalPointGpu == EvalBatch (hashes)");
  int nonzero = 0;
  for (int k = 0; k < nk; ++k) nonzero += !Eq(y_a[k], kZero);
  EXPECT(nonzero > nk / 2, "VdpfEvalPointGpu wrote shares");
  for (void *p : {(void *)d_s0s, (void *)d_betas, (void *)d_cs, (void *)d_ocws, (void *)d_seeds, (void *)d_ya, (void *)d_yb, (void *)d_pa,
                  (void *)d_pb, (void *)d_al, (void *)d_xs, (void *)d_st, (void *)d_cws, (void *)d_cw_s, (void *)d_extra})
    cudaFree(p);
  Prg::FreeCtxs(ctxs);
}

static void VdpfSha256() {
  using S = fss::hash::Sha256;
  using B = fss::hash::Blake3;
  const int4 key = {0x12345678, int(0x9abcdef0u), 0x13572468, 0x2468ace0};
  const int4 iv[2] = {{int(0x0fedcba9u), int(0x87654321u), 0x2468ace0, 0x13572468}, {5, 6, 7, 8}};
  S sha(key);
  // hashlib.sha256(key || msg) and the two digests of (a, b), words little-endian
  const int4 msg[4] = {{1, 2, 3, 4}, {5, 6, 7, 8}, {9, 10, 11, 12}, {13, 14, 15, 16}};
  auto h = sha.Hash(cuda::std::span<const int4, 4>(msg, 4));
  const uint32_t want_h[8] = {0xf045bcdfu, 0x57a787deu, 0xdc80996fu, 0xc989b531u, 0x0036c49cu, 0x2c667c7du, 0x8452b952u, 0x8f54d914u};
  EXPECT(std::memcmp(h.data(), want_h, 32) == 0, "Sha256::Hash(64 B) == hashlib");
  auto x4 = sha.Hash(cuda::std::tuple<int4, const int4>(int4{1, 2, 3, 5}, msg[1]));  // a's lsb is overwritten: 4 and 5
  const uint32_t want_x[16] = {0x4e9d17d5u, 0x3f300836u, 0x880bfb60u, 0x16f8a5d0u, 0x321211efu, 0x9d7bcd2bu, 0xc4a1f849u, 0xc980a395u,
                               0xd1ee6fa3u, 0x16d749f6u, 0x49e83e72u, 0x75db73e7u, 0xcdcc72acu, 0x571697f3u, 0x1ecf46bfu, 0x5557a2ffu};
  EXPECT(std::memcmp(x4.data(), want_x, 64) == 0, "Sha256::Hash(a, b) == hashlib");
  VdpfHashes("vdpf<Sha256, Sha256>", S(key), S(int4{9, 8, 7, 6}));
  VdpfHashes("vdpf<Sha256, Blake3>", S(key), B(cuda::std::span<const int4, 2>(iv, 2)));
  VdpfHashes("vdpf<Blake3, Sha256>", B(cuda::std::span<const int4, 2>(iv, 2)), S(key));
}

int main() {
  try {
    VdpfN8();
    VdpfSha256();
    VdpfGpuFreeFunctions();
    DpfN8();
    DcfN64();
    HalfTreeAndGrotto();
    BatchedChaCha();
  } catch (const std::exception &e) {
    std::printf("FAIL exception: %s\n", e.what());
    return 2;
  }
  std::printf(g_fail ? "shim sample: %d failure(s)\n" : "shim sample: all checks passed\n", g_fail);
  return g_fail ? 1 : 0;
}
