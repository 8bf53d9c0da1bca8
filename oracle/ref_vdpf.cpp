// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// ref_vdpf.cpp: extern "C" shim over the UNMODIFIED reference `fss::Vdpf` (vdpf.cuh) with
// XorHash = Hash = fss::hash::Blake3 or fss::hash::Sha256 (RefSel::pad = 0 / 1), included from
// /root/reference/include, for a table of (in_bits, group, prg, hash) instantiations.  Part of oracle/_ref/libfssref.so; pins the VDPF part of the
// C restatement (oracle/fss_oracle.c) and generates the VDPF golden fixtures.
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <type_traits>
#include <vector>
#include <omp.h>

#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/hash/blake3.cuh>
#include <fss/hash/sha256.cuh>
#include <fss/prg/aes128_mmo.cuh>
#include <fss/prg/chacha.cuh>
#include <fss/vdpf.cuh>

#include "ref_shim.h"

namespace {

using u128 = __uint128_t;

template <class Prg>
struct Holder;
template <>
struct Holder<fss::prg::Aes128Mmo<2>> {
  cuda::std::array<EVP_CIPHER_CTX *, 2> ctxs;
  explicit Holder(const RefParams &p) {
    const unsigned char *ks[2] = {p.prg_key, p.prg_key + 16};
    ctxs = fss::prg::Aes128Mmo<2>::CreateCtxs(ks);
  }
  ~Holder() { fss::prg::Aes128Mmo<2>::FreeCtxs(ctxs); }
  fss::prg::Aes128Mmo<2> make() { return fss::prg::Aes128Mmo<2>(ctxs); }
};
template <>
struct Holder<fss::prg::ChaCha<2>> {
  int nonce[2];
  explicit Holder(const RefParams &p) { memcpy(nonce, p.prg_key, 8); }
  fss::prg::ChaCha<2> make() { return fss::prg::ChaCha<2>(nonce); }
};

template <int N>
using InOf = std::conditional_t<(N <= 32), uint32_t, std::conditional_t<(N <= 64), uint64_t, u128>>;

template <class In>
In LoadIn(const uint8_t *p, int in_bytes) {
  u128 v = 0;
  memcpy(&v, p, in_bytes);
  return static_cast<In>(v);
}

template <class H>
H MakeHash(const uint8_t iv[32]);
template <>
fss::hash::Blake3 MakeHash<fss::hash::Blake3>(const uint8_t iv[32]) {
  int4 v[2];
  memcpy(v, iv, 32);
  return fss::hash::Blake3(cuda::std::span<const int4, 2>(v, 2));
}
template <>
fss::hash::Sha256 MakeHash<fss::hash::Sha256>(const uint8_t iv[32]) {  // 16-byte key = the first half of the slot
  int4 k;
  memcpy(&k, iv, 16);
  return fss::hash::Sha256(k);
}

struct VdpfOps {
  void (*gen)(const RefParams *, const uint8_t *ivs, size_t, const void *, const void *, const void *, void *, void *,
      void *, void *, int);
  void (*eval)(const RefParams *, const uint8_t *ivs, int, size_t, const void *, const void *, const void *,
      const void *, const void *, void *, void *, int);
  void (*prove)(const RefParams *, const uint8_t *ivs, size_t, size_t, const void *, const void *, void *);
  void (*evalall)(const RefParams *, const uint8_t *ivs, int, size_t, const void *, const void *, const void *,
      const void *, void *, void *, int);
};

template <int N, class G, class Prg, class H>
struct Ad {
  using In = InOf<N>;
  using S = fss::Vdpf<N, G, Prg, H, H, In>;
  using Cw = typename S::Cw;
  using Arr4 = cuda::std::array<int4, 4>;

  static void Gen(const RefParams *pp, const uint8_t *ivs, size_t nkeys, const void *s0s_, const void *alphas_,
      const void *betas_, void *cws_, void *cs_, void *ocws_, void *status_, int threads) {
    const RefParams &p = *pp;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
      Holder<Prg> h(p);
      S s{h.make(), MakeHash<H>(ivs), MakeHash<H>(ivs + 32)};
#pragma omp for schedule(static)
      for (size_t k = 0; k < nkeys; ++k) {
        const int4 *s0s = static_cast<const int4 *>(s0s_) + 2 * k;
        static_cast<int32_t *>(status_)[k] = s.Gen(static_cast<Cw *>(cws_) + k * N, static_cast<Arr4 *>(cs_)[k],
            static_cast<int4 *>(ocws_)[k], cuda::std::span<const int4, 2>(s0s, 2),
            LoadIn<In>(static_cast<const uint8_t *>(alphas_) + k * p.in_bytes, p.in_bytes),
            static_cast<const int4 *>(betas_)[k]);
      }
    }
  }
  static void Eval(const RefParams *pp, const uint8_t *ivs, int party, size_t nkeys, const void *seeds_,
      const void *cws_, const void *cs_, const void *ocws_, const void *xs_, void *ys_, void *pis_, int threads) {
    const RefParams &p = *pp;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
      Holder<Prg> h(p);
      S s{h.make(), MakeHash<H>(ivs), MakeHash<H>(ivs + 32)};
#pragma omp for schedule(static)
      for (size_t k = 0; k < nkeys; ++k) {
        const Arr4 &cs = static_cast<const Arr4 *>(cs_)[k];
        int4 y;
        Arr4 pi = s.Eval(party != 0, static_cast<const int4 *>(seeds_)[k],
            cuda::std::span<const Cw>(static_cast<const Cw *>(cws_) + k * N, N),
            cuda::std::span<const int4, 4>(cs.data(), 4), static_cast<const int4 *>(ocws_)[k],
            LoadIn<In>(static_cast<const uint8_t *>(xs_) + k * p.in_bytes, p.in_bytes), y);
        static_cast<int4 *>(ys_)[k] = y;
        static_cast<Arr4 *>(pis_)[k] = pi;
      }
    }
  }
  static void Prove(const RefParams *pp, const uint8_t *ivs, size_t nkeys, size_t m, const void *pts_,
      const void *cs_, void *pis_) {
    Holder<Prg> h(*pp);
    S s{h.make(), MakeHash<H>(ivs), MakeHash<H>(ivs + 32)};
    for (size_t k = 0; k < nkeys; ++k) {
      const Arr4 &cs = static_cast<const Arr4 *>(cs_)[k];
      s.Prove(cuda::std::span<const Arr4>(static_cast<const Arr4 *>(pts_) + k * m, m),
          cuda::std::span<const int4, 4>(cs.data(), 4), static_cast<Arr4 *>(pis_)[k]);
    }
  }
  static void EvalAll(const RefParams *pp, const uint8_t *ivs, int party, size_t nkeys, const void *seeds_,
      const void *cws_, const void *cs_, const void *ocws_, void *ys_, void *pis_, int threads) {
    const RefParams &p = *pp;
    const size_t nl = size_t(1) << N;
#pragma omp parallel num_threads(threads > 0 ? threads : 1)
    {
      Holder<Prg> h(p);
      S s{h.make(), MakeHash<H>(ivs), MakeHash<H>(ivs + 32)};
#pragma omp for schedule(dynamic, 1)
      for (size_t k = 0; k < nkeys; ++k) {
        const Arr4 &cs = static_cast<const Arr4 *>(cs_)[k];
        s.EvalAll(party != 0, static_cast<const int4 *>(seeds_)[k],
            cuda::std::span<const Cw>(static_cast<const Cw *>(cws_) + k * N, N),
            cuda::std::span<const int4, 4>(cs.data(), 4), static_cast<const int4 *>(ocws_)[k],
            cuda::std::span<int4>(static_cast<int4 *>(ys_) + k * nl, nl), static_cast<Arr4 *>(pis_)[k]);
      }
    }
  }
  static VdpfOps Ops() {
    VdpfOps o{};
    o.gen = &Gen;
    o.eval = &Eval;
    o.prove = &Prove;
    if constexpr (N <= 24) o.evalall = &EvalAll;
    return o;
  }
};

using Table = std::map<std::tuple<int, int, uint64_t, uint64_t, int, int>, VdpfOps>;
Table &table() {
  static Table t;
  return t;
}

using GBytes = fss::group::Bytes;
using GU32 = fss::group::Uint<uint32_t>;
using GU64 = fss::group::Uint<uint64_t>;
using GU127 = fss::group::Uint<u128, (u128(1) << 127)>;
using GU64p = fss::group::Uint<uint64_t, 18446744073709551557ull>;

template <int N, class Prg, int prg_tag, class H, int hash_tag>
void RegN() {
  table()[{N, 0, 0, 0, prg_tag, hash_tag}] = Ad<N, GBytes, Prg, H>::Ops();
  table()[{N, 3, 0, 0, prg_tag, hash_tag}] = Ad<N, GU32, Prg, H>::Ops();
  table()[{N, 4, 0, 0, prg_tag, hash_tag}] = Ad<N, GU64, Prg, H>::Ops();
  table()[{N, 5, 0, 0x8000000000000000ull, prg_tag, hash_tag}] = Ad<N, GU127, Prg, H>::Ops();
  table()[{N, 4, 18446744073709551557ull, 0, prg_tag, hash_tag}] = Ad<N, GU64p, Prg, H>::Ops();
}
template <int N>
void Reg() {
  RegN<N, fss::prg::Aes128Mmo<2>, REF_PRG_AES128_MMO, fss::hash::Blake3, 0>();
  RegN<N, fss::prg::ChaCha<2>, REF_PRG_CHACHA, fss::hash::Blake3, 0>();
}
template <int N>
void RegSha() {  // fewer domains (compile time): the hash plugin does not interact with in_bits
  RegN<N, fss::prg::Aes128Mmo<2>, REF_PRG_AES128_MMO, fss::hash::Sha256, 1>();
  RegN<N, fss::prg::ChaCha<2>, REF_PRG_CHACHA, fss::hash::Sha256, 1>();
}

struct Init {
  Init() {
    Reg<1>(); Reg<3>(); Reg<8>(); Reg<12>(); Reg<16>(); Reg<20>(); Reg<32>(); Reg<40>(); Reg<64>(); Reg<128>();
    RegSha<1>(); RegSha<8>(); RegSha<12>(); RegSha<32>(); RegSha<64>(); RegSha<128>();
  }
} g_init;

const VdpfOps *Find(const RefSel *s) {
  auto it = table().find({s->in_bits, s->group, s->mod_lo, s->mod_hi, s->prg, s->pad});
  return it == table().end() ? nullptr : &it->second;
}

// which = 0: XorHash (a, b) -> 64 B; 1: Hash 64 B -> 32 B
template <class H>
static int RunHash(const uint8_t *iv, int which, size_t n, const void *msgs, void *out) {
  H h = MakeHash<H>(iv);
  for (size_t i = 0; i < n; ++i) {
    if (which == 0) {
      const int4 *m = static_cast<const int4 *>(msgs) + 2 * i;
      auto o = h.Hash(cuda::std::tuple<int4, const int4>{m[0], m[1]});
      memcpy(static_cast<int4 *>(out) + 4 * i, o.data(), 64);
    } else {
      const int4 *m = static_cast<const int4 *>(msgs) + 4 * i;
      auto o = h.Hash(cuda::std::span<const int4, 4>(m, 4));
      memcpy(static_cast<int4 *>(out) + 2 * i, o.data(), 32);
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int ref_vdpf_supported(const RefSel *sel) { return Find(sel) != nullptr; }

int ref_vdpf_gen(const RefSel *sel, const RefParams *p, const uint8_t *ivs, size_t nkeys, const void *s0s,
    const void *alphas, const void *betas, void *cws, void *cs, void *ocws, void *status, int threads) {
  auto *o = Find(sel);
  if (!o) return -1;
  o->gen(p, ivs, nkeys, s0s, alphas, betas, cws, cs, ocws, status, threads);
  return 0;
}
int ref_vdpf_eval(const RefSel *sel, const RefParams *p, const uint8_t *ivs, int party, size_t nkeys,
    const void *seeds, const void *cws, const void *cs, const void *ocws, const void *xs, void *ys, void *pis,
    int threads) {
  auto *o = Find(sel);
  if (!o) return -1;
  o->eval(p, ivs, party, nkeys, seeds, cws, cs, ocws, xs, ys, pis, threads);
  return 0;
}
int ref_vdpf_prove(const RefSel *sel, const RefParams *p, const uint8_t *ivs, size_t nkeys, size_t m,
    const void *pi_tildes, const void *cs, void *pis) {
  auto *o = Find(sel);
  if (!o) return -1;
  o->prove(p, ivs, nkeys, m, pi_tildes, cs, pis);
  return 0;
}
int ref_vdpf_evalall(const RefSel *sel, const RefParams *p, const uint8_t *ivs, int party, size_t nkeys,
    const void *seeds, const void *cws, const void *cs, const void *ocws, void *ys, void *pis, int threads) {
  auto *o = Find(sel);
  if (!o || !o->evalall) return -1;
  o->evalall(p, ivs, party, nkeys, seeds, cws, cs, ocws, ys, pis, threads);
  return 0;
}
// which = 0: XorHash (a, b) -> 64 B; 1: Hash 64 B -> 32 B   (hash/blake3.cuh:143-171)
int ref_blake3(const uint8_t *iv, int which, size_t n, const void *msgs, void *out) {
  return RunHash<fss::hash::Blake3>(iv, which, n, msgs, out);
}
// the same two interfaces of fss::hash::Sha256 (hash/sha256.cuh:44-89); iv = 16-byte key (+ 16 ignored bytes)
int ref_sha256(const uint8_t *iv, int which, size_t n, const void *msgs, void *out) {
  return RunHash<fss::hash::Sha256>(iv, which, n, msgs, out);
}
}  // extern "C"
