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
// The other device entry points host_api.cu forwards to: every output byte is a digest of every input byte of its key, so
// the chunked host loops (offsets inside a device set, strided gathers, double buffering, the D2H of each chunk) are
// right only if the mock device saw what the caller passed.
static uint64_t DigestBytes(uint64_t h, const uint8_t *p, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) h = Mix(h, Load64(p + i));
  for (; i < n; ++i) h = Mix(h, p[i]);
  return h;
}
static void FillFrom(uint64_t h, uint8_t *out, size_t n) {
  for (size_t i = 0; i < n; i += 8) {
    h = Mix(h, i);
    std::memcpy(out + i, &h, n - i < 8 ? n - i : 8);
  }
}
// gen: cws[k] (ncw x 32 B) and ocws[k] from (s0s[k], alpha[k], beta[k])
static void GenKey(const fssb200_ctx *c, const uint8_t *s0s, const uint8_t *alpha, const uint8_t *beta, uint8_t *cws, uint8_t *ocw) {
  uint64_t h = DigestBytes(0x6e6, s0s, 32);
  h = DigestBytes(h, alpha, size_t(c->p.in_bytes));
  if (beta) h = DigestBytes(h, beta, 16);
  FillFrom(h, cws, size_t(c->ncw) * 32);
  if (ocw) FillFrom(Mix(h, 0x0c), ocw, 16);
}
int fssb200_gen(const fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *ocws, size_t nkeys,
                void *stream) {
  if (g_fail_after.load() >= 0 && g_fail_after.fetch_sub(1) == 0) return 700;
  ++g_kernel_launches;
  const size_t ib = size_t(c->p.in_bytes), cwb = size_t(c->ncw) * 32;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k)
      GenKey(c, static_cast<const uint8_t *>(s0s) + 32 * k, static_cast<const uint8_t *>(alphas) + ib * k,
             betas ? static_cast<const uint8_t *>(betas) + 16 * k : nullptr, static_cast<uint8_t *>(cws) + cwb * k,
             ocws ? static_cast<uint8_t *>(ocws) + 16 * k : nullptr);
  });
  return 0;
}
// vdpf gen: cws, cs (64 B), ocws, status; vdpf eval: ys, pis (64 B) from everything of the key
static void VdpfGenKey(const fssb200_ctx *c, const uint8_t *s0s, const uint8_t *alpha, const uint8_t *beta, uint8_t *cws, uint8_t *cs,
                       uint8_t *ocw, int32_t *status) {
  uint64_t h = DigestBytes(DigestBytes(DigestBytes(0x7d9f, s0s, 32), alpha, size_t(c->p.in_bytes)), beta, 16);
  FillFrom(h, cws, size_t(c->ncw) * 32);
  FillFrom(Mix(h, 1), cs, 64);
  FillFrom(Mix(h, 2), ocw, 16);
  *status = int32_t(h & 1);
}
int fssb200_vdpf_gen(const fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *cs, void *ocws,
                     void *status, size_t nkeys, void *stream) {
  ++g_kernel_launches;
  const size_t ib = size_t(c->p.in_bytes), cwb = size_t(c->ncw) * 32;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k)
      VdpfGenKey(c, static_cast<const uint8_t *>(s0s) + 32 * k, static_cast<const uint8_t *>(alphas) + ib * k,
                 static_cast<const uint8_t *>(betas) + 16 * k, static_cast<uint8_t *>(cws) + cwb * k, static_cast<uint8_t *>(cs) + 64 * k,
                 static_cast<uint8_t *>(ocws) + 16 * k, static_cast<int32_t *>(status) + k);
  });
  return 0;
}
static void VdpfEvalKey(const fssb200_ctx *c, int party, const uint8_t *seed, const uint8_t *cws, const uint8_t *cs, const uint8_t *ocw,
                        const uint8_t *x, uint8_t *y, uint8_t *pi) {
  uint64_t h = DigestBytes(Mix(0xe7a1, uint64_t(party)), seed, 16);
  h = DigestBytes(DigestBytes(DigestBytes(h, cws, size_t(c->ncw) * 32), cs, 64), ocw, 16);
  h = DigestBytes(h, x, size_t(c->p.in_bytes));
  FillFrom(h, y, 16);
  FillFrom(Mix(h, 3), pi, 64);
}
int fssb200_vdpf_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs, const void *ocws,
                      const void *xs, void *ys, void *pis, size_t nkeys, void *stream) {
  ++g_kernel_launches;
  const size_t ib = size_t(c->p.in_bytes), cwb = size_t(c->ncw) * 32;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k)
      VdpfEvalKey(c, party, static_cast<const uint8_t *>(seeds) + 16 * k, static_cast<const uint8_t *>(cws) + cwb * k,
                  static_cast<const uint8_t *>(cs) + 64 * k, static_cast<const uint8_t *>(ocws) + 16 * k,
                  static_cast<const uint8_t *>(xs) + ib * k, static_cast<uint8_t *>(ys) + 16 * k, static_cast<uint8_t *>(pis) + 64 * k);
  });
  return 0;
}
// level-major eval: ys[k] from seed, x, every cw_s[i][k] (and cw_v), the key's control-bit words, out_cw, ocw
static void LmKey(const fssb200_ctx *c, int party, size_t k, size_t nkeys, const uint8_t *seeds, const uint8_t *cw_s, const uint8_t *cw_v,
                  const uint8_t *extra, const uint8_t *out_cw, const uint8_t *ocws, const uint8_t *xs, uint8_t *y) {
  const size_t n = size_t(c->p.in_bits), nw = (n + 31) / 32;
  uint64_t h = DigestBytes(Mix(0x1e7e1, uint64_t(party)), seeds + 16 * k, 16);
  h = DigestBytes(h, xs + size_t(c->p.in_bytes) * k, size_t(c->p.in_bytes));
  for (size_t i = 0; i < n; ++i) {
    h = DigestBytes(h, cw_s + (i * nkeys + k) * 16, 16);
    if (cw_v) h = DigestBytes(h, cw_v + (i * nkeys + k) * 16, 16);
  }
  if (extra)
    for (size_t w = 0; w < nw; ++w) h = DigestBytes(h, extra + (w * nkeys + k) * 4, 4);
  if (out_cw) h = DigestBytes(h, out_cw + 16 * k, 16);
  if (ocws) h = DigestBytes(h, ocws + 16 * k, 16);
  FillFrom(h, y, 16);
}
int fssb200_eval_levelmajor(const fssb200_ctx *c, int party, const void *seeds, const void *cw_s, const void *cw_v, const void *extra,
                            const void *out_cw, const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream) {
  if (g_fail_after.load() >= 0 && g_fail_after.fetch_sub(1) == 0) return 700;
  ++g_kernel_launches;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k)
      LmKey(c, party, k, nkeys, static_cast<const uint8_t *>(seeds), static_cast<const uint8_t *>(cw_s), static_cast<const uint8_t *>(cw_v),
            static_cast<const uint8_t *>(extra), static_cast<const uint8_t *>(out_cw), static_cast<const uint8_t *>(ocws),
            static_cast<const uint8_t *>(xs), static_cast<uint8_t *>(ys) + 16 * k);
  });
  return 0;
}
// full-domain: leaf x of key k = digest(key) mixed with x (16 B; Grotto: 1 byte)
static uint64_t AllKeyDigest(const fssb200_ctx *c, int party, const uint8_t *seed, const uint8_t *cws, const uint8_t *ocw) {
  uint64_t h = DigestBytes(Mix(0xa11, uint64_t(party)), seed, 16);
  h = DigestBytes(h, cws, size_t(c->ncw) * 32);
  return ocw ? DigestBytes(h, ocw, 16) : h;
}
static void AllLeaf(const fssb200_ctx *c, uint64_t hk, uint64_t x, uint8_t *out) {
  const uint64_t v[2] = {Mix(hk, x), Mix(hk, ~x)};
  std::memcpy(out, v, c->p.scheme == FSSB200_SCHEME_GROTTO ? 1 : 16);
}
int fssb200_eval_all(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws, void *ys, size_t nkeys,
                     uint64_t leaf_begin, uint64_t leaf_count, void *stream) {
  if (g_fail_after.load() >= 0 && g_fail_after.fetch_sub(1) == 0) return 700;
  ++g_kernel_launches;
  const size_t cwb = size_t(c->ncw) * 32, lb = c->p.scheme == FSSB200_SCHEME_GROTTO ? 1 : 16;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t k = 0; k < nkeys; ++k) {
      const uint64_t hk = AllKeyDigest(c, party, static_cast<const uint8_t *>(seeds) + 16 * k, static_cast<const uint8_t *>(cws) + cwb * k,
                                       ocws ? static_cast<const uint8_t *>(ocws) + 16 * k : nullptr);
      for (uint64_t i = 0; i < leaf_count; ++i) AllLeaf(c, hk, leaf_begin + i, static_cast<uint8_t *>(ys) + (k * leaf_count + i) * lb);
    }
  });
  return 0;
}
int fssb200_prg_gen(const fssb200_ctx *, const void *seeds, void *out, int mul, size_t nseeds, void *stream) {
  ++g_kernel_launches;
  mockcuda::S(static_cast<cudaStream_t>(stream))->push([=] {
    for (size_t i = 0; i < nseeds; ++i)
      for (int j = 0; j < mul; ++j)
        FillFrom(Mix(DigestBytes(0x9e6, static_cast<const uint8_t *>(seeds) + 16 * i, 16), uint64_t(j)),
                 static_cast<uint8_t *>(out) + (i * size_t(mul) + size_t(j)) * 16, 16);
  });
  return 0;
}
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
  c->ncw = (scheme == FSSB200_SCHEME_HALFTREE || scheme == FSSB200_SCHEME_VDPF) ? in_bits : in_bits + 1;
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
  // ---- the other host entry points: chunked, double-buffered loops over the same arena pool -------------------------------
  {
    std::minstd_rand rng(99);
    auto fill = [&](std::vector<uint8_t> &v, size_t bytes) {
      v.resize(bytes);
      for (auto &x : v) x = uint8_t(rng());
    };
    // gen: DPF n=32 and Half-Tree n=20 (separate output correction words), chunk sizes that leave a ragged tail
    for (fssb200_ctx *c : {dpf, ht}) {
      const size_t nk = 5003, cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
      const bool half = c->p.scheme == FSSB200_SCHEME_HALFTREE;
      std::vector<uint8_t> s0s, al, be, cws(nk * cwb, 0xEE), ocws(half ? nk * 16 : 0, 0xEE), want_c(nk * cwb), want_o(half ? nk * 16 : 0);
      fill(s0s, nk * 32); fill(al, nk * ib); fill(be, nk * 16);
      for (size_t k = 0; k < nk; ++k)
        GenKey(c, &s0s[32 * k], &al[ib * k], &be[16 * k], &want_c[cwb * k], half ? &want_o[16 * k] : nullptr);
      for (size_t ck : {size_t(0), size_t(1000), size_t(4999), size_t(1)}) {
        if (ck == 1 && c == ht) continue;  // (one key per chunk: once is enough)
        fssb200_ctx_reserve_host(c, ck);
        const size_t use = ck == 1 ? 37 : nk;
        std::fill(cws.begin(), cws.end(), 0xEE);
        const int rc = fssb200_gen_host(c, s0s.data(), al.data(), be.data(), cws.data(), half ? ocws.data() : nullptr, use);
        CHECK(rc == 0 && std::memcmp(cws.data(), want_c.data(), use * cwb) == 0, "gen_host chunk %zu rc=%d", ck, rc);
        if (half) CHECK(std::memcmp(ocws.data(), want_o.data(), use * 16) == 0, "gen_host ocws chunk %zu", ck);
        CHECK(use == nk || cws[use * cwb] == 0xEE, "gen_host wrote past the batch");
      }
      fssb200_ctx_reserve_host(c, 0);
    }
    // prg_gen_host
    for (int mul : {1, 2, 4}) {
      const size_t ns = 3001;
      std::vector<uint8_t> seeds, out(ns * 16 * size_t(mul), 0xEE), want(ns * 16 * size_t(mul));
      fill(seeds, ns * 16);
      for (size_t i = 0; i < ns; ++i)
        for (int j = 0; j < mul; ++j) FillFrom(Mix(DigestBytes(0x9e6, &seeds[16 * i], 16), uint64_t(j)), &want[(i * size_t(mul) + size_t(j)) * 16], 16);
      const int rc = fssb200_prg_gen_host(dpf, seeds.data(), out.data(), mul, ns);
      CHECK(rc == 0 && out == want, "prg_gen_host mul=%d rc=%d", mul, rc);
    }
    // eval_all_host: several keys, leaf ranges (granule 1 in the mock), Half-Tree ocws, Grotto bytes
    fssb200_ctx *ea_dpf = MakeCtx(FSSB200_SCHEME_DPF, 12, 2), *ea_ht = MakeCtx(FSSB200_SCHEME_HALFTREE, 10, 2),
                *ea_gr = MakeCtx(FSSB200_SCHEME_GROTTO, 11, 2);
    // geometries: every key in one launch | groups of 2 + a ragged last group | one key at a time in ranges of 1 MiB of leaves
    // (n = 17: 2 MiB per key, two ranges) -- (keys-per-chunk cap, FSSB200_ALL_SET_MB, in_bits of the DPF / Half-Tree contexts)
    fssb200_ctx *ea_big = MakeCtx(FSSB200_SCHEME_DPF, 17, 4), *ea_big_ht = MakeCtx(FSSB200_SCHEME_HALFTREE, 17, 4);
    struct AllGeo { size_t cap; const char *set_mb; bool big; };
    for (const AllGeo geo : {AllGeo{0, "64", false}, AllGeo{2, "64", false}, AllGeo{0, "1", true}})
    for (fssb200_ctx *c : {geo.big ? ea_big : ea_dpf, geo.big ? ea_big_ht : ea_ht, ea_gr}) {
      setenv("FSSB200_ALL_SET_MB", geo.set_mb, 1);
      fssb200_ctx_reserve_host(c, geo.cap);
      const long launches0 = g_kernel_launches.load();
      const size_t nk = 5, cwb = size_t(c->ncw) * 32, lb = c->p.scheme == FSSB200_SCHEME_GROTTO ? 1 : 16;
      const bool half = c->p.scheme == FSSB200_SCHEME_HALFTREE, grotto = c->p.scheme == FSSB200_SCHEME_GROTTO;
      const uint64_t N = uint64_t(1) << c->p.in_bits;
      std::vector<uint8_t> seeds, cws, ocws;
      fill(seeds, nk * 16); fill(cws, nk * cwb); fill(ocws, nk * 16);
      for (auto range : {std::pair<uint64_t, uint64_t>(0, 0), std::pair<uint64_t, uint64_t>(N / 4, N / 2), std::pair<uint64_t, uint64_t>(N - 7, 7)}) {
        if (grotto && range.first) continue;  // (Grotto: the scan needs the domain from leaf 0)
        const uint64_t cnt = range.second ? range.second : N - range.first;
        std::vector<uint8_t> ys(nk * cnt * lb + 16, 0xEE);
        const int rc = fssb200_eval_all_host(c, 1, seeds.data(), cws.data(), half ? ocws.data() : nullptr, ys.data(), nk, range.first, range.second);
        size_t bad = 0;
        for (size_t k = 0; k < nk && rc == 0; ++k) {
          const uint64_t hk = AllKeyDigest(c, 1, &seeds[16 * k], &cws[cwb * k], half ? &ocws[16 * k] : nullptr);
          for (uint64_t i = 0; i < cnt; ++i) {
            uint8_t w[16];
            AllLeaf(c, hk, range.first + i, w);
            bad += std::memcmp(w, &ys[(k * cnt + i) * lb], lb) != 0;
          }
        }
        CHECK(rc == 0 && bad == 0 && ys[nk * cnt * lb] == 0xEE, "eval_all_host scheme %d range %llu+%llu rc=%d bad=%zu", c->p.scheme,
              (unsigned long long)range.first, (unsigned long long)range.second, rc, bad);
      }
      CHECK(fssb200_eval_all_host(c, 1, seeds.data(), cws.data(), ocws.data(), seeds.data(), 1, N, 0) == FSSB200_ERANGE, "eval_all_host range check");
      // launches of the three (Grotto: one) ranges: 1 per range with every key in one chunk, 3 with groups of 2, 2, 1
      const long launches = g_kernel_launches.load() - launches0, ranges = grotto ? 1 : 3;
      if (!geo.big) CHECK(launches == ranges * (geo.cap ? 3 : 1), "eval_all_host scheme %d cap %zu: %ld launches", c->p.scheme, geo.cap, launches);
      // 1 MiB sets: the whole range = 2 chunks per key, the half range = 1 per key, the 7-leaf range = one launch for all keys
      if (geo.big && !grotto) CHECK(launches == long(nk) * 3 + 1, "eval_all_host scheme %d ranges: %ld launches", c->p.scheme, launches);
      fssb200_ctx_reserve_host(c, 0);
      unsetenv("FSSB200_ALL_SET_MB");
    }
    // eval_levelmajor_host: strided gathers of [level][key] arrays, chunks with a ragged tail; DPF (extra + out_cw), DCF (cw_v + out_cw),
    // Half-Tree (extra + ocws), and a 100-level domain (4 control-bit words per key)
    fssb200_ctx *lm_wide = MakeCtx(FSSB200_SCHEME_DPF, 100, 16);
    for (fssb200_ctx *c : {dpf, dcf, ht, lm_wide}) {
      const size_t nk = 4001, n = size_t(c->p.in_bits), nw = (n + 31) / 32, ib = size_t(c->p.in_bytes);
      const bool isdcf = c->p.scheme == FSSB200_SCHEME_DCF, half = c->p.scheme == FSSB200_SCHEME_HALFTREE;
      std::vector<uint8_t> seeds, cw_s, cw_v, extra, out_cw, ocws, xs, want(nk * 16);
      fill(seeds, nk * 16); fill(cw_s, n * nk * 16); fill(cw_v, n * nk * 16); fill(extra, nw * nk * 4); fill(out_cw, nk * 16);
      fill(ocws, nk * 16); fill(xs, nk * ib);
      for (size_t k = 0; k < nk; ++k)
        LmKey(c, 1, k, nk, seeds.data(), cw_s.data(), isdcf ? cw_v.data() : nullptr, isdcf ? nullptr : extra.data(),
              half ? nullptr : out_cw.data(), half ? ocws.data() : nullptr, xs.data(), &want[16 * k]);
      for (size_t ck : {size_t(0), size_t(1500), size_t(4000)}) {
        fssb200_ctx_reserve_host(c, ck);
        std::vector<uint8_t> ys(nk * 16, 0xEE);
        const int rc = fssb200_eval_levelmajor_host(c, 1, seeds.data(), cw_s.data(), isdcf ? cw_v.data() : nullptr, isdcf ? nullptr : extra.data(),
                                                    half ? nullptr : out_cw.data(), half ? ocws.data() : nullptr, xs.data(), ys.data(), nk);
        CHECK(rc == 0 && ys == want, "eval_levelmajor_host scheme %d n=%zu chunk %zu rc=%d", c->p.scheme, n, ck, rc);
      }
      fssb200_ctx_reserve_host(c, 0);
    }
    // an injected launch error in the chunked loops: reported, nothing left in flight, the next call is fine
    {
      const size_t nk = 3000, cwb = size_t(dpf->ncw) * 32;
      std::vector<uint8_t> s0s, al, be, cws(nk * cwb);
      fill(s0s, nk * 32); fill(al, nk * 4); fill(be, nk * 16);
      fssb200_ctx_reserve_host(dpf, 1000);
      g_fail_after.store(1);
      const int rc = fssb200_gen_host(dpf, s0s.data(), al.data(), be.data(), cws.data(), nullptr, nk);
      g_fail_after.store(-1);
      CHECK(rc == 700, "gen_host: injected launch error not reported: rc = %d", rc);
      CHECK(fssb200_gen_host(dpf, s0s.data(), al.data(), be.data(), cws.data(), nullptr, nk) == 0, "gen_host after a failed call");
      fssb200_ctx_reserve_host(dpf, 0);
    }
    {  // ... and in the key-group loop of eval_all_host (groups of 2 keys: the second launch fails)
      const size_t nk = 5, cwb = size_t(ea_ht->ncw) * 32, N = size_t(1) << ea_ht->p.in_bits;
      std::vector<uint8_t> seeds, cws, ocws, ys(nk * N * 16, 0xEE);
      fill(seeds, nk * 16); fill(cws, nk * cwb); fill(ocws, nk * 16);
      fssb200_ctx_reserve_host(ea_ht, 2);
      g_fail_after.store(1);
      const int rc = fssb200_eval_all_host(ea_ht, 0, seeds.data(), cws.data(), ocws.data(), ys.data(), nk, 0, 0);
      g_fail_after.store(-1);
      CHECK(rc == 700, "eval_all_host: injected launch error not reported: rc = %d", rc);
      CHECK(fssb200_eval_all_host(ea_ht, 0, seeds.data(), cws.data(), ocws.data(), ys.data(), nk, 0, 0) == 0, "eval_all_host after a failed call");
      size_t bad = 0;
      for (size_t k = 0; k < nk; ++k) {
        const uint64_t hk = AllKeyDigest(ea_ht, 0, &seeds[16 * k], &cws[cwb * k], &ocws[16 * k]);
        for (uint64_t i = 0; i < N; ++i) {
          uint8_t w[16];
          AllLeaf(ea_ht, hk, i, w);
          bad += std::memcmp(w, &ys[(k * N + i) * 16], 16) != 0;
        }
      }
      CHECK(bad == 0, "eval_all_host after a failed call: %zu wrong leaves", bad);
      fssb200_ctx_reserve_host(ea_ht, 0);
    }
    // VDPF host calls
    fssb200_ctx *vd = MakeCtx(FSSB200_SCHEME_VDPF, 24, 4);
    {
      const size_t nk = 2503, cwb = size_t(vd->ncw) * 32;
      std::vector<uint8_t> s0s, al, be, xs, cws(nk * cwb, 0xEE), cs(nk * 64), ocws(nk * 16), ys(nk * 16), pis(nk * 64);
      std::vector<int32_t> status(nk, -1);
      fill(s0s, nk * 32); fill(al, nk * 4); fill(be, nk * 16); fill(xs, nk * 4);
      fssb200_ctx_reserve_host(vd, 600);
      int rc = fssb200_vdpf_gen_host(vd, s0s.data(), al.data(), be.data(), cws.data(), cs.data(), ocws.data(), status.data(), nk);
      size_t bad = 0;
      for (size_t k = 0; k < nk && rc == 0; ++k) {
        std::vector<uint8_t> wc(cwb), wcs(64), wo(16);
        int32_t ws = -1;
        VdpfGenKey(vd, &s0s[32 * k], &al[4 * k], &be[16 * k], wc.data(), wcs.data(), wo.data(), &ws);
        bad += std::memcmp(wc.data(), &cws[cwb * k], cwb) != 0 || std::memcmp(wcs.data(), &cs[64 * k], 64) != 0 ||
            std::memcmp(wo.data(), &ocws[16 * k], 16) != 0 || ws != status[k];
      }
      CHECK(rc == 0 && bad == 0, "vdpf_gen_host rc=%d bad=%zu", rc, bad);
      std::vector<uint8_t> seeds(nk * 16);
      for (size_t k = 0; k < nk; ++k) std::memcpy(&seeds[16 * k], &s0s[32 * k + 16], 16);
      rc = fssb200_vdpf_eval_host(vd, 1, seeds.data(), cws.data(), cs.data(), ocws.data(), xs.data(), ys.data(), pis.data(), nk);
      bad = 0;
      for (size_t k = 0; k < nk && rc == 0; ++k) {
        uint8_t wy[16], wp[64];
        VdpfEvalKey(vd, 1, &seeds[16 * k], &cws[cwb * k], &cs[64 * k], &ocws[16 * k], &xs[4 * k], wy, wp);
        bad += std::memcmp(wy, &ys[16 * k], 16) != 0 || std::memcmp(wp, &pis[64 * k], 64) != 0;
      }
      CHECK(rc == 0 && bad == 0, "vdpf_eval_host rc=%d bad=%zu", rc, bad);
      fssb200_ctx_reserve_host(vd, 0);
    }
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
