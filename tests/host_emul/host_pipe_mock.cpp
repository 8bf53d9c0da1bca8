// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY.  fss_b200/csrc/host_api.cu -- the product's host-array entry points, unmodified -- compiled for
// the CPU against the mock CUDA runtime of mock_cuda.h (-DFSSB200_HOST_MOCK), so that the worker crew, the staging ring,
// the piece / chunk hand-offs, the adaptive direct pieces, the arena pool and the error paths run in the GPU-less build
// container, under ThreadSanitizer too (tests/test_host_mock.py).  The "kernels" digest every key's bytes: a result is
// right only if every byte of every key reached the "device" intact and in the format the launch claimed.
#define FSSB200_HOST_MOCK 1
#include <cstdio>
#include <vector>

#include "../../fss_b200/csrc/host_api.cu"

// ---- the slice of api.cu that host_api.cu calls, as stream operations on the mock device ------------------------------
static uint64_t Mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  return h * 0xff51afd7ed558ccdull;
}
static uint64_t Load64(const uint8_t *p) {
  uint64_t v;
  std::memcpy(&v, p, 8);
  return v;
}
// ys[k] = digest(party, seed, x, ocw, every s entry, every flag): the same value from reference-layout rows and packed rows
static void DigestKey(const fssb200_ctx *c, int party, const uint8_t *seed, const uint8_t *row, bool packed, const uint8_t *ocw,
                      const uint8_t *x, uint8_t *y) {
  uint64_t h = Mix(0x1234, uint64_t(party));
  h = Mix(Mix(h, Load64(seed)), Load64(seed + 8));
  uint64_t xv = 0;
  std::memcpy(&xv, x, size_t(c->p.in_bytes) < 8 ? size_t(c->p.in_bytes) : 8);
  h = Mix(h, xv);
  if (ocw) h = Mix(Mix(h, Load64(ocw)), Load64(ocw + 8));
  const bool flagged = c->p.scheme == FSSB200_SCHEME_DPF || c->p.scheme == FSSB200_SCHEME_HALFTREE;
  for (int i = 0; i < c->ncw; ++i) {
    const uint8_t *s = packed ? row + 16 * i : row + 32 * i;
    h = Mix(Mix(h, Load64(s)), Load64(s + 8));
    if (flagged && i < 128) {  // (a packed row carries the flags of entries 0..127: for n = 128 entry 128 is the output
                               //  correction word, whose flag byte no scheme reads -- dpf.cuh:158)
      const bool f = packed ? ((row[size_t(c->ncw) * 16 + (i >> 3)] >> (i & 7)) & 1) != 0 : row[32 * i + 16] != 0;
      h = Mix(h, f ? 0x77 : 0x11);
    } else if (!flagged) {  // DCF: the v half is payload
      h = Mix(Mix(h, Load64(row + 32 * i + 16)), Load64(row + 32 * i + 24));
    }
  }
  const uint64_t out[2] = {h, Mix(h, 0xabcdef)};
  std::memcpy(y, out, 16);
}

static std::atomic<long> g_kernel_launches{0};
static std::atomic<int> g_fail_after{-1};  // error injection: the n-th launch from now returns an error

extern "C" {
size_t fssb200_packed_row_bytes(const fssb200_ctx *c) {
  if (!c || (c->p.scheme != FSSB200_SCHEME_DPF && c->p.scheme != FSSB200_SCHEME_HALFTREE)) return 0;
  return size_t(c->ncw) * 16u + 16u;
}
uint64_t fssb200_eval_all_granule(const fssb200_ctx *) { return 1; }
static int EvalOp(const fssb200_ctx *c, int party, const void *seeds, const void *rows, const void *ocws, const void *xs, void *ys,
                  size_t nkeys, void *stream, bool packed) {
  if (g_fail_after.load() >= 0 && g_fail_after.fetch_sub(1) == 0) return 700;  // an injected "cudaError_t"
  ++g_kernel_launches;
  const size_t rowb = packed ? size_t(c->ncw) * 16 + 16 : size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k)
      DigestKey(c, party, static_cast<const uint8_t *>(seeds) + 16 * k, static_cast<const uint8_t *>(rows) + rowb * k, packed,
                ocws ? static_cast<const uint8_t *>(ocws) + 16 * k : nullptr, static_cast<const uint8_t *>(xs) + ib * k,
                static_cast<uint8_t *>(ys) + 16 * k);
  });
  return 0;
}
int fssb200_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws, const void *xs, void *ys,
                 size_t nkeys, void *stream) {
  return EvalOp(c, party, seeds, cws, ocws, xs, ys, nkeys, stream, false);
}
int fssb200_eval_packed(const fssb200_ctx *c, int party, const void *seeds, const void *rows, const void *ocws, const void *xs,
                        void *ys, size_t nkeys, void *stream) {
  return EvalOp(c, party, seeds, rows, ocws, xs, ys, nkeys, stream, true);
}
// the other device entry points host_api.cu references: not exercised by this test
int fssb200_gen(const fssb200_ctx *, const void *, const void *, const void *, void *, void *, size_t, void *) { return FSSB200_ESCHEME; }
int fssb200_vdpf_gen(const fssb200_ctx *, const void *, const void *, const void *, void *, void *, void *, void *, size_t, void *) { return FSSB200_ESCHEME; }
int fssb200_vdpf_eval(const fssb200_ctx *, int, const void *, const void *, const void *, const void *, const void *, void *, void *, size_t, void *) { return FSSB200_ESCHEME; }
int fssb200_eval_levelmajor(const fssb200_ctx *, int, const void *, const void *, const void *, const void *, const void *, const void *, const void *, void *, size_t, void *) { return FSSB200_ESCHEME; }
int fssb200_eval_all(const fssb200_ctx *, int, const void *, const void *, const void *, void *, size_t, uint64_t, uint64_t, void *) { return FSSB200_ESCHEME; }
int fssb200_prg_gen(const fssb200_ctx *, const void *, void *, int, size_t, void *) { return FSSB200_ESCHEME; }
}

// ---- scenarios ------------------------------------------------------------------------------------------------------------
static int g_bad = 0;
#define CHECK(cond, ...)                        \
  do {                                          \
    if (!(cond)) {                              \
      std::printf("FAIL: " __VA_ARGS__);        \
      std::printf("\n");                        \
      ++g_bad;                                  \
    }                                           \
  } while (0)

struct Batch {
  std::vector<uint8_t> seeds, cws, ocws, xs, want;
  size_t n;
};
static Batch MakeBatch(const fssb200_ctx *c, size_t n, unsigned seed, bool with_ocws) {
  Batch b;
  b.n = n;
  std::minstd_rand rng(seed);
  auto fill = [&](std::vector<uint8_t> &v, size_t bytes) {
    v.resize(bytes);
    for (auto &x : v) x = uint8_t(rng());
  };
  fill(b.seeds, 16 * n);
  fill(b.cws, size_t(c->ncw) * 32 * n);
  fill(b.xs, size_t(c->p.in_bytes) * n);
  if (with_ocws) fill(b.ocws, 16 * n);
  if (c->p.scheme != FSSB200_SCHEME_DCF)  // the padding of {int4 s; bool flag}: garbage on purpose (it must never matter)
    for (size_t i = 0; i < n * size_t(c->ncw); ++i) b.cws[32 * i + 16] = (rng() & 1) ? uint8_t(rng() | 1) : 0;
  b.want.resize(16 * n);
  for (size_t k = 0; k < n; ++k)
    DigestKey(c, 1, &b.seeds[16 * k], &b.cws[size_t(c->ncw) * 32 * k], false, with_ocws ? &b.ocws[16 * k] : nullptr,
              &b.xs[size_t(c->p.in_bytes) * k], &b.want[16 * k]);
  return b;
}
static fssb200_ctx *MakeCtx(int scheme, int in_bits, int in_bytes) {
  fssb200_ctx *c = new fssb200_ctx();
  std::memset(&c->p, 0, sizeof(c->p));
  c->p.scheme = scheme;
  c->p.in_bits = in_bits;
  c->p.in_bytes = in_bytes;
  c->ncw = scheme == FSSB200_SCHEME_HALFTREE ? in_bits : in_bits + 1;
  return c;
}
static void RunOne(fssb200_ctx *c, const Batch &b, int mode, bool pin_in, bool pin_out, const char *what) {
  fssb200_ctx_set_host_mode(c, mode);
  if (pin_in) {
    mockcuda::register_pinned(b.seeds.data(), b.seeds.size());
    mockcuda::register_pinned(b.cws.data(), b.cws.size());
    mockcuda::register_pinned(b.xs.data(), b.xs.size());
    if (!b.ocws.empty()) mockcuda::register_pinned(b.ocws.data(), b.ocws.size());
  }
  std::vector<uint8_t> ys(16 * b.n, 0xEE);
  if (pin_out) mockcuda::register_pinned(ys.data(), ys.size());
  const int rc = fssb200_eval_host(c, 1, b.seeds.data(), b.cws.data(), b.ocws.empty() ? nullptr : b.ocws.data(), b.xs.data(),
                                   ys.data(), b.n);
  CHECK(rc == 0, "%s: rc = %d", what, rc);
  size_t bad = 0;
  for (size_t k = 0; k < b.n; ++k) bad += std::memcmp(&ys[16 * k], &b.want[16 * k], 16) != 0;
  CHECK(bad == 0, "%s: %zu of %zu keys wrong (mode %d, pinned in %d out %d)", what, bad, b.n, mode, int(pin_in), int(pin_out));
  uint64_t pk = 0, dk = 0;
  int th = 0;
  fssb200_ctx_host_stats(c, &pk, &dk, &th);
  CHECK(pk + dk == b.n, "%s: stats %llu + %llu != %zu", what, (unsigned long long)pk, (unsigned long long)dk, b.n);
  if (mode == 1) CHECK(pk == 0, "%s: mode 1 staged %llu keys", what, (unsigned long long)pk);
  if ((mode == 2 || !pin_in) && fssb200_packed_row_bytes(c) && b.n >= 8192 && mode != 1 && th >= 1)  // (th = 0: no worker crew)
    CHECK(dk == 0, "%s: %llu keys crossed as they are", what, (unsigned long long)dk);
  if (pin_in) {
    mockcuda::unregister_pinned(b.seeds.data());
    mockcuda::unregister_pinned(b.cws.data());
    mockcuda::unregister_pinned(b.xs.data());
    if (!b.ocws.empty()) mockcuda::unregister_pinned(b.ocws.data());
  }
  if (pin_out) mockcuda::unregister_pinned(ys.data());
}

int main(int argc, char **argv) {
  const bool quick = argc > 1 && std::string(argv[1]) == "quick";
  // geometry knobs: small pieces and chunks so that rings, device sets and ragged tails wrap many times
  struct Geo { const char *chunk_bits, *piece_bits, *slots, *nt; } geos[] = {
      {"12", "10", "3", "0"}, {"13", "11", "2", "1"}, {"11", "11", "5", "0"}, {"14", "9", "4", "1"}};
  fssb200_ctx *dpf = MakeCtx(FSSB200_SCHEME_DPF, 32, 4), *ht = MakeCtx(FSSB200_SCHEME_HALFTREE, 20, 4),
              *dcf = MakeCtx(FSSB200_SCHEME_DCF, 16, 2), *wide = MakeCtx(FSSB200_SCHEME_DPF, 128, 16);
  const Batch b_dpf = MakeBatch(dpf, quick ? 20011 : 50021, 1, false), b_ht = MakeBatch(ht, 17000, 2, true),
              b_dcf = MakeBatch(dcf, 12345, 3, false), b_wide = MakeBatch(wide, 9001, 4, false),
              b_small = MakeBatch(dpf, 100, 5, false), b_one = MakeBatch(dpf, 1, 6, false);
  for (const Geo &g : geos) {
    setenv("FSSB200_PIPE_CHUNK_BITS", g.chunk_bits, 1);
    setenv("FSSB200_PIPE_PIECE_BITS", g.piece_bits, 1);
    setenv("FSSB200_PIPE_SLOTS", g.slots, 1);
    setenv("FSSB200_PACK_NT", g.nt, 1);
    for (int mode = 0; mode <= 3; ++mode)
      for (int pin = 0; pin < 4; ++pin) {
        RunOne(dpf, b_dpf, mode, pin & 1, pin & 2, "dpf n=32");
        if (quick && (mode == 3 || pin == 1)) continue;
        RunOne(ht, b_ht, mode, pin & 1, pin & 2, "halftree n=20");
        RunOne(dcf, b_dcf, mode, pin & 1, pin & 2, "dcf n=16 (no padding: staged copy / direct)");
        RunOne(wide, b_wide, mode, pin & 1, pin & 2, "dpf n=128 (129 flags: two flag words)");
      }
    RunOne(dpf, b_small, 0, true, true, "100 keys (plain chunked path)");
    RunOne(dpf, b_one, 0, false, false, "1 key");
  }
  // preferred chunk sizes that do not divide by the piece size; a chunk smaller than a piece
  for (size_t ck : {size_t(3000), size_t(5000), size_t(700)}) {
    fssb200_ctx_reserve_host(dpf, ck);
    RunOne(dpf, b_dpf, 0, true, true, "odd chunk size");
    RunOne(dpf, b_dpf, 2, false, false, "odd chunk size, pageable");
  }
  fssb200_ctx_reserve_host(dpf, 0);
  // concurrency: 6 threads, one context each kind of call at the same time (the crew is lent to one large call, the others
  // stage with their own thread)
  {
    std::vector<std::thread> th;
    for (int t = 0; t < 6; ++t)
      th.emplace_back([&, t] {
        if (t % 3 == 0) RunOne(t == 0 ? dpf : ht, t == 0 ? b_dpf : b_ht, 0, true, true, "concurrent large");
        else for (int i = 0; i < 20; ++i) RunOne(MakeCtx(FSSB200_SCHEME_DPF, 32, 4), i % 2 ? b_small : b_one, 0, false, false, "concurrent small");
      });
    for (auto &t : th) t.join();
  }
  // multi-device entry point: 3 "devices" (>= 3: the automatic mode sends plain pieces), then 2 (adaptive, crew shared)
  for (int ndev : {3, 2}) {
    std::vector<fssb200_ctx *> cs;
    for (int d = 0; d < ndev; ++d) cs.push_back(MakeCtx(FSSB200_SCHEME_DPF, 32, 4));
    std::vector<uint8_t> ys(16 * b_dpf.n);
    mockcuda::register_pinned(b_dpf.seeds.data(), b_dpf.seeds.size());
    mockcuda::register_pinned(b_dpf.cws.data(), b_dpf.cws.size());
    mockcuda::register_pinned(b_dpf.xs.data(), b_dpf.xs.size());
    std::vector<int> rcs(size_t(ndev), -1);
    const int rc = fssb200_eval_host_multi(cs.data(), ndev, 1, b_dpf.seeds.data(), b_dpf.cws.data(), nullptr, b_dpf.xs.data(), ys.data(),
                                           b_dpf.n, rcs.data());
    CHECK(rc == 0 && std::memcmp(ys.data(), b_dpf.want.data(), ys.size()) == 0, "eval_host_multi ndev=%d rc=%d", ndev, rc);
    mockcuda::unregister_pinned(b_dpf.seeds.data());
    mockcuda::unregister_pinned(b_dpf.cws.data());
    mockcuda::unregister_pinned(b_dpf.xs.data());
  }
  // the same with blocks small enough for the BALANCED split (devices claim key blocks from one counter, two calls in
  // flight per device), with and without an injected launch error
  setenv("FSSB200_MULTI_MIN_BLOCK_BITS", "10", 1);
  setenv("FSSB200_MULTI_MAX_BLOCK_BITS", "13", 1);
  for (int fail : {-1, 5}) {
    const int ndev = 3;
    std::vector<fssb200_ctx *> cs;
    for (int d = 0; d < ndev; ++d) cs.push_back(MakeCtx(FSSB200_SCHEME_DPF, 32, 4));
    std::vector<uint8_t> ys(16 * b_dpf.n);
    mockcuda::register_pinned(b_dpf.seeds.data(), b_dpf.seeds.size());
    mockcuda::register_pinned(b_dpf.cws.data(), b_dpf.cws.size());
    mockcuda::register_pinned(b_dpf.xs.data(), b_dpf.xs.size());
    std::vector<int> rcs(size_t(ndev), -1);
    const long launches0 = g_kernel_launches.load();
    g_fail_after.store(fail);
    const int rc = fssb200_eval_host_multi(cs.data(), ndev, 1, b_dpf.seeds.data(), b_dpf.cws.data(), nullptr, b_dpf.xs.data(), ys.data(),
                                           b_dpf.n, rcs.data());
    g_fail_after.store(-1);
    if (fail < 0) {
      CHECK(rc == 0 && std::memcmp(ys.data(), b_dpf.want.data(), ys.size()) == 0, "balanced eval_host_multi rc=%d", rc);
      CHECK(g_kernel_launches.load() - launches0 >= 6, "balanced split did not split: %ld launches", g_kernel_launches.load() - launches0);
    } else {
      int bad = 0;
      for (int d = 0; d < ndev; ++d) bad += rcs[size_t(d)] == 700;
      CHECK(rc == 700 && bad >= 1, "balanced: injected launch error not reported: rc = %d", rc);
    }
    mockcuda::unregister_pinned(b_dpf.seeds.data());
    mockcuda::unregister_pinned(b_dpf.cws.data());
    mockcuda::unregister_pinned(b_dpf.xs.data());
  }
  // balanced split with PAGEABLE inputs (every block is staged by the workers), and with two devices in host mode 1
  for (int variant = 0; variant < 2; ++variant) {
    const int ndev = variant == 0 ? 3 : 2;
    std::vector<fssb200_ctx *> cs;
    for (int d = 0; d < ndev; ++d) {
      cs.push_back(MakeCtx(FSSB200_SCHEME_DPF, 32, 4));
      if (variant == 1) fssb200_ctx_set_host_mode(cs.back(), 1);
    }
    if (variant == 1) {
      mockcuda::register_pinned(b_dpf.seeds.data(), b_dpf.seeds.size());
      mockcuda::register_pinned(b_dpf.cws.data(), b_dpf.cws.size());
      mockcuda::register_pinned(b_dpf.xs.data(), b_dpf.xs.size());
    }
    std::vector<uint8_t> ys(16 * b_dpf.n);
    std::vector<int> rcs(size_t(ndev), -1);
    const long launches0 = g_kernel_launches.load();
    const int rc = fssb200_eval_host_multi(cs.data(), ndev, 1, b_dpf.seeds.data(), b_dpf.cws.data(), nullptr, b_dpf.xs.data(), ys.data(),
                                           b_dpf.n, rcs.data());
    CHECK(rc == 0 && std::memcmp(ys.data(), b_dpf.want.data(), ys.size()) == 0, "balanced eval_host_multi variant %d rc=%d", variant, rc);
    CHECK(g_kernel_launches.load() - launches0 >= 4, "balanced split (variant %d) did not split", variant);
    if (variant == 1) {
      mockcuda::unregister_pinned(b_dpf.seeds.data());
      mockcuda::unregister_pinned(b_dpf.cws.data());
      mockcuda::unregister_pinned(b_dpf.xs.data());
    }
  }
  unsetenv("FSSB200_MULTI_MIN_BLOCK_BITS");
  unsetenv("FSSB200_MULTI_MAX_BLOCK_BITS");
  // error path: the 2nd launch of a call fails -> the call returns the error, nothing stays in flight, the next call is fine
  for (int mode : {0, 1, 2}) {
    mockcuda::register_pinned(b_dpf.seeds.data(), b_dpf.seeds.size());
    mockcuda::register_pinned(b_dpf.cws.data(), b_dpf.cws.size());
    mockcuda::register_pinned(b_dpf.xs.data(), b_dpf.xs.size());
    fssb200_ctx_set_host_mode(dpf, mode);
    g_fail_after.store(1);  // the second launch of the call fails
    std::vector<uint8_t> ys(16 * b_dpf.n);
    const int rc = fssb200_eval_host(dpf, 1, b_dpf.seeds.data(), b_dpf.cws.data(), nullptr, b_dpf.xs.data(), ys.data(), b_dpf.n);
    CHECK(rc == 700, "injected launch error not reported (mode %d): rc = %d", mode, rc);
    g_fail_after.store(-1);
    mockcuda::unregister_pinned(b_dpf.seeds.data());
    mockcuda::unregister_pinned(b_dpf.cws.data());
    mockcuda::unregister_pinned(b_dpf.xs.data());
    RunOne(dpf, b_dpf, mode, true, true, "call after a failed call");
  }
  // arena pool: what is cached is bounded; trim frees everything the pool holds
  uint64_t dev_b = 0, pin_b = 0;
  fssb200_host_cached_bytes(&dev_b, &pin_b);
  CHECK(dev_b > 0, "no arena cached");
  fssb200_host_trim();
  fssb200_host_cached_bytes(&dev_b, &pin_b);
  CHECK(dev_b == 0 && pin_b == 0, "trim left %llu + %llu bytes", (unsigned long long)dev_b, (unsigned long long)pin_b);
  CHECK(mockcuda::live_allocs().load() == 0, "%ld mock allocations leaked", mockcuda::live_allocs().load());
  std::printf("kernel launches: %ld\n", g_kernel_launches.load());
  std::printf(g_bad ? "host pipeline mock: %d failure(s)\n" : "host pipeline mock: all checks passed\n", g_bad);
  return g_bad ? 1 : 0;
}
