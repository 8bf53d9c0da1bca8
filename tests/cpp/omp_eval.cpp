// SPDX-License-Identifier: Apache-2.0
// TEST: re-entrancy of the header shim's single-key members.  The reference's `Eval` is a const pure function and its
// own CPU benchmark calls it on ONE scheme object from every OpenMP thread (src/bench_cpu.cu:157-161: `#pragma omp
// parallel for` over keys, `dpf.Eval(...)` in the body).  The same user code must work against this shim: every
// thread's call checks its own staging arena out of the library's pool.  Compiled with g++ -std=c++20 -fopenmp.
// Checks: threaded results == single-threaded results, and reconstruction (dpf.cuh:170-214, dcf.cuh:205-276).
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/prg/aes128_mmo.cuh>

static int g_fail = 0;
#define EXPECT(cond, what)                                          \
  do {                                                              \
    if (!(cond)) {                                                  \
      std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__);   \
      ++g_fail;                                                     \
    }                                                               \
  } while (0)

static unsigned char k0[16] = {1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16};
static unsigned char k1[16] = {16, 15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1};
static unsigned char k2[16] = {1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8};
static unsigned char k3[16] = {8, 8, 7, 7, 6, 6, 5, 5, 4, 4, 3, 3, 2, 2, 1, 1};

static int4 RandBlock(std::mt19937_64 &rng) {
  int4 v = {int(rng()), int(rng()), int(rng()), int(rng())};
  v.w &= ~1;
  return v;
}
static bool Eq(int4 a, int4 b) { return std::memcmp(&a, &b, 16) == 0; }

template <class Scheme, class Group, class In>
static void Run(Scheme &sch, const char *name, bool lt_pred) {
  using Cw = typename Scheme::Cw;
  constexpr int kKeys = 256;
  std::mt19937_64 rng(42);
  std::vector<Cw> cws(size_t(kKeys) * Scheme::kNumCw);
  std::vector<int4> s0(kKeys), s1(kKeys), beta(kKeys), y0(kKeys), y1(kKeys), z0(kKeys), z1(kKeys);
  std::vector<In> alpha(kKeys), x(kKeys);
  for (int i = 0; i < kKeys; ++i) {
    s0[i] = RandBlock(rng);
    s1[i] = RandBlock(rng);
    beta[i] = RandBlock(rng);
    alpha[i] = In(rng());
    x[i] = (i % 3 == 0) ? alpha[i] : In(rng());
    const int4 ss[2] = {s0[i], s1[i]};
    sch.Gen(&cws[size_t(i) * Scheme::kNumCw], ss, alpha[i], beta[i]);
  }
  // single-threaded results first
  for (int i = 0; i < kKeys; ++i) {
    z0[i] = sch.Eval(false, s0[i], &cws[size_t(i) * Scheme::kNumCw], x[i]);
    z1[i] = sch.Eval(true, s1[i], &cws[size_t(i) * Scheme::kNumCw], x[i]);
  }
  int threads_seen = 0;
#pragma omp parallel for schedule(dynamic, 4)
  for (int i = 0; i < kKeys; ++i) {  // the loop of src/bench_cpu.cu:157-161
    y0[i] = sch.Eval(false, s0[i], &cws[size_t(i) * Scheme::kNumCw], x[i]);
    y1[i] = sch.Eval(true, s1[i], &cws[size_t(i) * Scheme::kNumCw], x[i]);
    if (i == 0) threads_seen = omp_get_num_threads();
  }
  int bad = 0, bad_rec = 0;
  for (int i = 0; i < kKeys; ++i) {
    if (!Eq(y0[i], z0[i]) || !Eq(y1[i], z1[i])) ++bad;
    const int4 sum = (Group::From(y0[i]) + Group::From(y1[i])).Into();
    const bool hit = lt_pred ? (x[i] < alpha[i]) : (x[i] == alpha[i]);
    const int4 want = hit ? Group::From(beta[i]).Into() : int4{0, 0, 0, 0};  // (Uint<u64>: words 2, 3 of beta do not count)
    if (!Eq(sum, want)) ++bad_rec;
  }
  std::printf("%s: %d keys, %d OpenMP threads, %d mismatches vs single-threaded, %d reconstruction failures\n", name, kKeys,
              threads_seen, bad, bad_rec);
  EXPECT(bad == 0, name);
  EXPECT(bad_rec == 0, name);
  EXPECT(threads_seen >= 2, "OpenMP ran with one thread");
}

int main() {
  {
    using Prg = fss::prg::Aes128Mmo<2>;
    const unsigned char *keys[2] = {k0, k1};
    auto ctxs = Prg::CreateCtxs(keys);
    Prg prg(ctxs);
    fss::Dpf<32, fss::group::Bytes, Prg, uint32_t> dpf{prg};
    Run<decltype(dpf), fss::group::Bytes, uint32_t>(dpf, "Dpf<32,Bytes,Aes128Mmo<2>>::Eval", false);
    Prg::FreeCtxs(ctxs);
  }
  {
    using Prg = fss::prg::Aes128Mmo<4>;
    using Group = fss::group::Uint<uint64_t>;
    const unsigned char *keys[4] = {k0, k1, k2, k3};
    auto ctxs = Prg::CreateCtxs(keys);
    Prg prg(ctxs);
    fss::Dcf<64, Group, Prg, uint64_t> dcf{prg};
    Run<decltype(dcf), Group, uint64_t>(dcf, "Dcf<64,Uint<u64>,Aes128Mmo<4>>::Eval", true);
    Prg::FreeCtxs(ctxs);
  }
  if (g_fail == 0) std::printf("all checks passed\n");
  return g_fail ? 1 : 0;
}
