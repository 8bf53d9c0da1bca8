// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// ref_shim.cpp: a thin extern "C" shim that instantiates the UNMODIFIED reference
// headers (included from where they lie, /root/reference/include -- nothing is
// copied into this repository) for a table of (scheme, in_bits, group, prg)
// parameter sets and exposes their Gen / Eval / EvalAll through runtime-dispatched
// C functions.  It is compiled by oracle/Makefile into oracle/_ref/libfssref_*.so
// and is used
//   * to pin the C restatement (oracle/fss_oracle.c) and to generate the golden
//     fixtures under tests/golden/ (oracle/make_golden.py),
//   * as the CPU baseline of bench.py (`cpu_baseline.kind = "reference"`,
//     `--impl reference`): the reference's own Eval / EvalAll with its OpenSSL
//     AES-NI PRG, one PRG context set per host thread.
//
// The translation unit is compiled several times with -DREF_PART=<k> so that the
// template instantiations build in parallel; every part registers its parameter
// sets into the same table type and exports `ref_part<k>_lookup`.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>
#include <omp.h>

#include <fss/dcf.cuh>
#include <fss/dpf.cuh>
#include <fss/grotto_dcf.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/half_tree_dpf.cuh>
#include <fss/prg/aes128_mmo.cuh>
#include <fss/prg/aes128_mmo_raw.cuh>
#include <fss/prg/aes128_mmo_soft.cuh>
#include <fss/prg/chacha.cuh>

#include "ref_shim.h"

namespace {

using u128 = __uint128_t;

// ---- PRG holders: own the key material a reference PRG object borrows --------
template <class Prg>
struct Holder;

template <int mul>
struct Holder<fss::prg::Aes128Mmo<mul>> {
  cuda::std::array<EVP_CIPHER_CTX *, mul> ctxs;
  explicit Holder(const RefParams &p) {
    const unsigned char *ks[mul];
    for (int i = 0; i < mul; ++i) ks[i] = p.prg_key + 16 * i;
    ctxs = fss::prg::Aes128Mmo<mul>::CreateCtxs(ks);
  }
  ~Holder() { fss::prg::Aes128Mmo<mul>::FreeCtxs(ctxs); }
  fss::prg::Aes128Mmo<mul> make() { return fss::prg::Aes128Mmo<mul>(ctxs); }
};

template <int mul>
struct Holder<fss::prg::Aes128MmoRaw<mul>> {
  uint8_t k[mul][16];
  explicit Holder(const RefParams &p) { memcpy(k, p.prg_key, sizeof(k)); }
  fss::prg::Aes128MmoRaw<mul> make() { return fss::prg::Aes128MmoRaw<mul>(k); }
};

template <int mul>
struct Holder<fss::prg::ChaCha<mul>> {
  int nonce[2];
  explicit Holder(const RefParams &p) { memcpy(nonce, p.prg_key, 8); }
  fss::prg::ChaCha<mul> make() { return fss::prg::ChaCha<mul>(nonce); }
};

template <int tag, int mul>
struct PrgOf;
template <int mul>
struct PrgOf<REF_PRG_AES128_MMO, mul> { using type = fss::prg::Aes128Mmo<mul>; };
template <int mul>
struct PrgOf<REF_PRG_CHACHA, mul> { using type = fss::prg::ChaCha<mul>; };
template <int mul>
struct PrgOf<REF_PRG_AES128_MMO_RAW, mul> { using type = fss::prg::Aes128MmoRaw<mul>; };

template <int N>
using InOf = std::conditional_t<(N <= 32), uint32_t, std::conditional_t<(N <= 64), uint64_t, u128>>;

template <class In>
In LoadIn(const uint8_t *p, int in_bytes) {
  u128 v = 0;
  memcpy(&v, p, in_bytes);
  return static_cast<In>(v);
}

int4 ToInt4(const uint8_t *p) {
  int4 v;
  memcpy(&v, p, 16);
  return v;
}

// ---- scheme adapters ------------------------------------------------------------
template <int N, class G, int prg_tag>
struct DpfAd {
  using Prg = typename PrgOf<prg_tag, 2>::type;
  using S = fss::Dpf<N, G, Prg, InOf<N>>;
  static constexpr int kNcw = N + 1;
  static S Make(Holder<Prg> &h, const RefParams &) { return S{h.make()}; }
  static void Gen(S &s, typename S::Cw *cws, int4 *, const int4 *s0s, InOf<N> a, int4 beta) { s.Gen(cws, s0s, a, beta); }
  static int4 Eval(S &s, bool b, int4 s0, const typename S::Cw *cws, int4, InOf<N> x) { return s.Eval(b, s0, cws, x); }
  static void EvalAll(S &s, bool b, int4 s0, const typename S::Cw *cws, int4, void *ys) {
    s.EvalAll(b, s0, cws, static_cast<int4 *>(ys));
  }
  static constexpr size_t kLeafBytes = 16;
};

template <int N, class G, int prg_tag, fss::DcfPred pred>
struct DcfAd {
  using Prg = typename PrgOf<prg_tag, 4>::type;
  using S = fss::Dcf<N, G, Prg, InOf<N>, pred>;
  static constexpr int kNcw = N + 1;
  static S Make(Holder<Prg> &h, const RefParams &) { return S{h.make()}; }
  static void Gen(S &s, typename S::Cw *cws, int4 *, const int4 *s0s, InOf<N> a, int4 beta) { s.Gen(cws, s0s, a, beta); }
  static int4 Eval(S &s, bool b, int4 s0, const typename S::Cw *cws, int4, InOf<N> x) { return s.Eval(b, s0, cws, x); }
  static void EvalAll(S &s, bool b, int4 s0, const typename S::Cw *cws, int4, void *ys) {
    s.EvalAll(b, s0, cws, static_cast<int4 *>(ys));
  }
  static constexpr size_t kLeafBytes = 16;
};

template <int N, class G, int prg_tag>
struct HtAd {
  using Prg = typename PrgOf<prg_tag, 1>::type;
  using S = fss::HalfTreeDpf<N, G, Prg, InOf<N>>;
  static constexpr int kNcw = N;
  static S Make(Holder<Prg> &h, const RefParams &p) { return S{h.make(), ToInt4(p.hash_key)}; }
  static void Gen(S &s, typename S::Cw *cws, int4 *ocw, const int4 *s0s, InOf<N> a, int4 beta) {
    s.Gen(cws, *ocw, s0s, a, beta);
  }
  static int4 Eval(S &s, bool b, int4 s0, const typename S::Cw *cws, int4 ocw, InOf<N> x) {
    return s.Eval(b, s0, cws, ocw, x);
  }
  static void EvalAll(S &s, bool b, int4 s0, const typename S::Cw *cws, int4 ocw, void *ys) {
    s.EvalAll(b, s0, cws, ocw, static_cast<int4 *>(ys));
  }
  static constexpr size_t kLeafBytes = 16;
};

template <int N, int prg_tag>
struct GrAd {
  using Prg = typename PrgOf<prg_tag, 2>::type;
  using S = fss::GrottoDcf<N, Prg, InOf<N>>;
  static constexpr int kNcw = N + 1;
  static S Make(Holder<Prg> &h, const RefParams &) { return S{h.make()}; }
  static void Gen(S &s, typename S::Cw *cws, int4 *, const int4 *s0s, InOf<N> a, int4) { s.Gen(cws, s0s, a); }
  static void EvalAll(S &s, bool b, int4 s0, const typename S::Cw *cws, int4, void *ys) {
    s.EvalAll(b, s0, cws, static_cast<bool *>(ys));
  }
  static constexpr size_t kLeafBytes = 1;
};

// ---- type-erased batch loops -----------------------------------------------------
template <class Ad, int N>
void GenBatch(const RefParams *pp, size_t nkeys, const void *s0s_, const void *alphas_, const void *betas_,
    void *cws_, void *ocws_, int threads) {
  using S = typename Ad::S;
  const RefParams &p = *pp;
  auto *s0s = static_cast<const int4 *>(s0s_);
  auto *alphas = static_cast<const uint8_t *>(alphas_);
  auto *betas = static_cast<const int4 *>(betas_);
  auto *cws = static_cast<typename S::Cw *>(cws_);
  auto *ocws = static_cast<int4 *>(ocws_);
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    Holder<typename Ad::Prg> h(p);
    S s = Ad::Make(h, p);
#pragma omp for schedule(static)
    for (size_t k = 0; k < nkeys; ++k) {
      int4 beta = betas ? betas[k] : int4{0, 0, 0, 0};
      int4 dummy;
      Ad::Gen(s, cws + k * Ad::kNcw, ocws ? ocws + k : &dummy, s0s + 2 * k,
          LoadIn<InOf<N>>(alphas + k * p.in_bytes, p.in_bytes), beta);
    }
  }
}

template <class Ad, int N>
void EvalBatch(const RefParams *pp, int party, size_t nkeys, const void *seeds_, const void *cws_,
    const void *ocws_, const void *xs_, void *ys_, int threads) {
  using S = typename Ad::S;
  const RefParams &p = *pp;
  auto *seeds = static_cast<const int4 *>(seeds_);
  auto *cws = static_cast<const typename S::Cw *>(cws_);
  auto *ocws = static_cast<const int4 *>(ocws_);
  auto *xs = static_cast<const uint8_t *>(xs_);
  auto *ys = static_cast<int4 *>(ys_);
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    Holder<typename Ad::Prg> h(p);
    S s = Ad::Make(h, p);
#pragma omp for schedule(static)
    for (size_t k = 0; k < nkeys; ++k) {
      ys[k] = Ad::Eval(s, party != 0, seeds[k], cws + k * Ad::kNcw, ocws ? ocws[k] : int4{0, 0, 0, 0},
          LoadIn<InOf<N>>(xs + k * p.in_bytes, p.in_bytes));
    }
  }
}

// Keys are spread over host threads (one PRG context set per thread); inside a key
// the reference's own EvalAll runs (its nested `omp parallel` gets a team of one
// unless the caller passes threads == -1, which runs keys serially and lets the
// reference's par_depth = -1 task recursion use every core, dpf.cuh:242-246).
template <class Ad, int N>
void EvalAllBatch(const RefParams *pp, int party, size_t nkeys, const void *seeds_, const void *cws_,
    const void *ocws_, void *ys_, int threads) {
  using S = typename Ad::S;
  const RefParams &p = *pp;
  auto *seeds = static_cast<const int4 *>(seeds_);
  auto *cws = static_cast<const typename S::Cw *>(cws_);
  auto *ocws = static_cast<const int4 *>(ocws_);
  auto *ys = static_cast<uint8_t *>(ys_);
  const size_t stride = (size_t(1) << N) * Ad::kLeafBytes;
  if (threads == -1) {
    Holder<typename Ad::Prg> h(p);
    S s = Ad::Make(h, p);
    for (size_t k = 0; k < nkeys; ++k)
      Ad::EvalAll(s, party != 0, seeds[k], cws + k * Ad::kNcw, ocws ? ocws[k] : int4{0, 0, 0, 0}, ys + k * stride);
    return;
  }
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    Holder<typename Ad::Prg> h(p);
    S s = Ad::Make(h, p);
#pragma omp for schedule(dynamic, 1)
    for (size_t k = 0; k < nkeys; ++k)
      Ad::EvalAll(s, party != 0, seeds[k], cws + k * Ad::kNcw, ocws ? ocws[k] : int4{0, 0, 0, 0}, ys + k * stride);
  }
}

template <class Ad, int N>
void GrottoPreprocessBatch(const RefParams *pp, int party, size_t nkeys, const void *seeds_, const void *cws_,
    void *pt_, int threads) {
  using S = typename Ad::S;
  const RefParams &p = *pp;
  auto *seeds = static_cast<const int4 *>(seeds_);
  auto *cws = static_cast<const typename S::Cw *>(cws_);
  auto *pt = static_cast<bool *>(pt_);
  const size_t stride = (size_t(2) << N) - 1;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
  {
    Holder<typename Ad::Prg> h(p);
    S s = Ad::Make(h, p);
#pragma omp for schedule(dynamic, 1)
    for (size_t k = 0; k < nkeys; ++k) {
      typename S::ParityTree t{pt + k * stride, party != 0};
      s.Preprocess(t, seeds[k], cws + k * Ad::kNcw);
    }
  }
}

template <class Ad, int N>
void GrottoLookupBatch(const RefParams *pp, size_t nkeys, const void *pt_, const void *xs_, void *ys_) {
  using S = typename Ad::S;
  const RefParams &p = *pp;
  auto *pt = static_cast<const bool *>(pt_);
  auto *xs = static_cast<const uint8_t *>(xs_);
  auto *ys = static_cast<bool *>(ys_);
  const size_t stride = (size_t(2) << N) - 1;
  for (size_t k = 0; k < nkeys; ++k) {
    typename S::ParityTree t{const_cast<bool *>(pt) + k * stride, false};
    ys[k] = S::Eval(t, LoadIn<InOf<N>>(xs + k * p.in_bytes, p.in_bytes));
  }
}

// ---- registration -------------------------------------------------------------------
using Table = std::map<std::tuple<int, int, int, uint64_t, uint64_t, int, int>, RefOps>;

Table &table() {
  static Table t;
  return t;
}

constexpr int kMaxEvalAllBits = 30;

template <class G>
struct GroupTag;
#define GT(TYPE, TAG, LO, HI)                       \
  template <>                                       \
  struct GroupTag<TYPE> {                           \
    static constexpr int tag = TAG;                 \
    static constexpr uint64_t lo = LO, hi = HI;     \
  }
using GBytes = fss::group::Bytes;
using GU8 = fss::group::Uint<uint8_t>;
using GU16 = fss::group::Uint<uint16_t>;
using GU32 = fss::group::Uint<uint32_t>;
using GU64 = fss::group::Uint<uint64_t>;
using GU127 = fss::group::Uint<u128, (u128(1) << 127)>;
using GU8p = fss::group::Uint<uint8_t, 251>;
using GU16p = fss::group::Uint<uint16_t, 65521>;
using GU32p = fss::group::Uint<uint32_t, 4294967291u>;
using GU64p = fss::group::Uint<uint64_t, 18446744073709551557ull>;
using GU128p = fss::group::Uint<u128, ((u128(1) << 127) - 1)>;
using GU128q = fss::group::Uint<u128, ((u128(0x1234567812345678ull) << 64) | 0x9abcdef09abcdef1ull)>;
GT(GBytes, 0, 0, 0);
GT(GU8, 1, 0, 0);
GT(GU16, 2, 0, 0);
GT(GU32, 3, 0, 0);
GT(GU64, 4, 0, 0);
GT(GU127, 5, 0, 0x8000000000000000ull);
GT(GU8p, 1, 251, 0);
GT(GU16p, 2, 65521, 0);
GT(GU32p, 3, 4294967291u, 0);
GT(GU64p, 4, 18446744073709551557ull, 0);
GT(GU128p, 5, 0xffffffffffffffffull, 0x7fffffffffffffffull);
GT(GU128q, 5, 0x9abcdef09abcdef1ull, 0x1234567812345678ull);

template <class Ad, int N, bool has_eval = true>
RefOps MakeOps() {
  RefOps o{};
  o.gen = &GenBatch<Ad, N>;
  if constexpr (has_eval) o.eval = &EvalBatch<Ad, N>;
  if constexpr (N <= kMaxEvalAllBits) o.evalall = &EvalAllBatch<Ad, N>;
  o.ncw = Ad::kNcw;
  return o;
}

template <int N, class G, int prg_tag>
void RegGroupSchemes(bool with_gt) {
  using T = GroupTag<G>;
  table()[{REF_SCHEME_DPF, N, T::tag, T::lo, T::hi, prg_tag, 0}] = MakeOps<DpfAd<N, G, prg_tag>, N>();
  table()[{REF_SCHEME_DCF, N, T::tag, T::lo, T::hi, prg_tag, 0}] =
      MakeOps<DcfAd<N, G, prg_tag, fss::DcfPred::kLt>, N>();
  if (with_gt) {
    if constexpr (N == 8 || N == 32) {
      table()[{REF_SCHEME_DCF, N, T::tag, T::lo, T::hi, prg_tag, 1}] =
          MakeOps<DcfAd<N, G, prg_tag, fss::DcfPred::kGt>, N>();
    }
  }
  table()[{REF_SCHEME_HALFTREE, N, T::tag, T::lo, T::hi, prg_tag, 0}] = MakeOps<HtAd<N, G, prg_tag>, N>();
}

template <int N, int prg_tag>
void RegGrotto() {
  RefOps o = MakeOps<GrAd<N, prg_tag>, N, false>();
  if constexpr (N <= 24) {
    o.grotto_preprocess = &GrottoPreprocessBatch<GrAd<N, prg_tag>, N>;
    o.grotto_lookup = &GrottoLookupBatch<GrAd<N, prg_tag>, N>;
  }
  table()[{REF_SCHEME_GROTTO, N, 0, 0, 0, prg_tag, 0}] = o;
}

template <int N, int prg_tag>
void RegFull() {
  RegGroupSchemes<N, GBytes, prg_tag>(true);
  RegGroupSchemes<N, GU8, prg_tag>(false);
  RegGroupSchemes<N, GU16, prg_tag>(false);
  RegGroupSchemes<N, GU32, prg_tag>(false);
  RegGroupSchemes<N, GU64, prg_tag>(true);
  RegGroupSchemes<N, GU127, prg_tag>(true);
  RegGroupSchemes<N, GU8p, prg_tag>(false);
  RegGroupSchemes<N, GU16p, prg_tag>(false);
  RegGroupSchemes<N, GU32p, prg_tag>(false);
  RegGroupSchemes<N, GU64p, prg_tag>(false);
  RegGroupSchemes<N, GU128p, prg_tag>(false);
  RegGroupSchemes<N, GU128q, prg_tag>(false);
  RegGrotto<N, prg_tag>();
}

template <int N, int prg_tag>
void RegLight() {
  RegGroupSchemes<N, GBytes, prg_tag>(false);
  RegGroupSchemes<N, GU64, prg_tag>(false);
  RegGroupSchemes<N, GU127, prg_tag>(false);
  RegGrotto<N, prg_tag>();
}

template <int N>
void RegFullAll() {
  RegFull<N, REF_PRG_AES128_MMO>();
  RegFull<N, REF_PRG_CHACHA>();
}
template <int N>
void RegLightAll() {
  RegLight<N, REF_PRG_AES128_MMO>();
  RegLight<N, REF_PRG_CHACHA>();
}
// AES-NI intrinsics PRG: only the timing configurations.
template <int N>
void RegRaw() {
  RegGroupSchemes<N, GBytes, REF_PRG_AES128_MMO_RAW>(false);
  RegGroupSchemes<N, GU127, REF_PRG_AES128_MMO_RAW>(false);
}

struct Init {
  Init() {
#if REF_PART == 0
    RegFullAll<8>();
    RegLightAll<1>();
    RegLightAll<2>();
    RegLightAll<3>();
#elif REF_PART == 1
    RegFullAll<16>();
    RegLightAll<5>();
    RegLightAll<12>();
    RegLightAll<20>();
#elif REF_PART == 2
    RegFullAll<32>();
    RegLightAll<24>();
    RegLightAll<28>();
#elif REF_PART == 3
    RegFullAll<64>();
    RegLightAll<33>();
    RegLightAll<48>();
#elif REF_PART == 4
    RegLightAll<10>();
    RegLightAll<40>();
    RegLightAll<100>();
    RegLightAll<128>();
    RegRaw<20>();
    RegRaw<24>();
    RegRaw<28>();
    RegRaw<32>();
    RegRaw<64>();
#else
#error "REF_PART must be 0..4"
#endif
  }
} g_init;

template <int mul, int prg_tag>
void PrgGenImpl(const RefParams &p, size_t n, const int4 *seeds, int4 *out) {
  using Prg = typename PrgOf<prg_tag, mul>::type;
  Holder<Prg> h(p);
  Prg prg = h.make();
  for (size_t i = 0; i < n; ++i) {
    auto o = prg.Gen(seeds[i]);
    for (int j = 0; j < mul; ++j) out[i * mul + j] = o[j];
  }
}

}  // namespace

#define CAT2(a, b) a##b
#define CAT(a, b) CAT2(a, b)

extern "C" {

const RefOps *CAT(CAT(ref_part, REF_PART), _lookup)(int scheme, int in_bits, int group, uint64_t mod_lo,
    uint64_t mod_hi, int prg, int pred) {
  if (scheme != REF_SCHEME_DCF) pred = 0;
  if (scheme == REF_SCHEME_GROTTO) {
    group = 0;
    mod_lo = mod_hi = 0;
  }
  auto it = table().find({scheme, in_bits, group, mod_lo, mod_hi, prg, pred});
  return it == table().end() ? nullptr : &it->second;
}

int CAT(CAT(ref_part, REF_PART), _count)(void) { return static_cast<int>(table().size()); }

#if REF_PART == 0
// out[i*mul + j] = block j of prg.Gen(seeds[i]); Aes128Soft (prg_tag 3) is checked
// here as well so that the three reference AES PRGs are pinned against each other.
int ref_prg_gen(const RefParams *p, int prg, int mul, size_t n, const void *seeds_, void *out_) {
  auto *seeds = static_cast<const int4 *>(seeds_);
  auto *out = static_cast<int4 *>(out_);
#define PG(TAG)                                                   \
  if (prg == TAG) {                                               \
    if (mul == 1) return PrgGenImpl<1, TAG>(*p, n, seeds, out), 0; \
    if (mul == 2) return PrgGenImpl<2, TAG>(*p, n, seeds, out), 0; \
    if (mul == 4) return PrgGenImpl<4, TAG>(*p, n, seeds, out), 0; \
    return -1;                                                    \
  }
  PG(REF_PRG_AES128_MMO)
  PG(REF_PRG_CHACHA)
  PG(REF_PRG_AES128_MMO_RAW)
#undef PG
  if (prg == 3 && (mul == 1 || mul == 2)) {
    uint32_t te0[256];
    uint8_t sbox[256];
    fss::prg::aes_detail::InitTe0(te0);
    fss::prg::aes_detail::InitSbox(sbox);
    uint8_t keys[2][16];
    memcpy(keys, p->prg_key, 32);
    for (size_t i = 0; i < n; ++i) {
      if (mul == 1) {
        fss::prg::Aes128Soft<1> prg1(keys, te0, sbox);
        out[i] = prg1.Gen(seeds[i])[0];
      } else {
        fss::prg::Aes128Soft<2> prg2(keys, te0, sbox);
        auto o = prg2.Gen(seeds[i]);
        out[2 * i] = o[0];
        out[2 * i + 1] = o[1];
      }
    }
    return 0;
  }
  return -1;
}

int ref_host_threads(void) { return omp_get_max_threads(); }
#endif

}  // extern "C"
