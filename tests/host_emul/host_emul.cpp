// SPDX-License-Identifier: Apache-2.0
//
// TEST INFRASTRUCTURE ONLY (never shipped, never loaded by fss_b200/).
//
// host_emul.cpp: compiles the `__host__ __device__` per-thread bodies of the CUDA kernels
// (fss_b200/csrc/{aes,prg,group,schemes}.cuh) for the CPU, so that the exact scheme / AES-table /
// group code that runs per CUDA thread can be checked against the oracle in the GPU-less build
// container (`pytest -m "not gpu"`).  The shared-memory table image, the PRMT address formation
// and the lane replication are emulated bit for bit (aes.cuh, `#if !FSS_DEVICE_CODE`).
// Launch geometry, shared-memory staging and the EvalAll work decomposition are device-only and
// are covered by the `-m gpu` tests.
#include <cstring>
#include <vector>

#include "../../fss_b200/csrc/schemes.cuh"

namespace fssb200 {

static std::vector<uint8_t> g_tables;

uint8_t *host_aes_tables() {
  if (g_tables.empty()) {
    g_tables.assign(kAesTblBytes, 0);
    constexpr U0Table u0t = make_u0();
    for (uint32_t x = 0; x < 256; ++x)
      for (uint32_t lane = 0; lane < 32; ++lane) {
        for (int t = 0; t < 4; ++t) {
          const uint32_t w = aes_tbl_word(u0t.v[x], t);
          std::memcpy(&g_tables[aes_tbl_offset(t, x, lane)], &w, 4);
        }
      }
  }
  return g_tables.data();
}

struct EmuCtx {
  PrgKeys keys;
  GroupArgs ga;
  int gk, scheme, n, in_bytes, prg, pred, ncw;
};

static int group_kind(const fssb200_params &p, uint32_t *vmask) {
  const bool has_mod = (p.mod_lo | p.mod_hi) != 0;
  *vmask = 0xffffffffu;
  switch (p.group) {
    case FSSB200_GROUP_BYTES: return kGrpBytes;
    case FSSB200_GROUP_U8: *vmask = 0xff; return has_mod ? kGrpU32Mod : kGrpU32;
    case FSSB200_GROUP_U16: *vmask = 0xffff; return has_mod ? kGrpU32Mod : kGrpU32;
    case FSSB200_GROUP_U32: return has_mod ? kGrpU32Mod : kGrpU32;
    case FSSB200_GROUP_U64: return has_mod ? kGrpU64Mod : kGrpU64;
    default: return (p.mod_hi == 0x8000000000000000ull && p.mod_lo == 0) ? kGrpU127 : kGrpU128Mod;
  }
}

static void make_ctx(const fssb200_params &p, EmuCtx &c) {
  std::memset(&c, 0, sizeof(c));
  c.scheme = p.scheme;
  c.n = p.in_bits;
  c.in_bytes = p.in_bytes;
  c.prg = p.prg;
  c.pred = p.pred;
  c.ncw = (p.scheme == FSSB200_SCHEME_HALFTREE || p.scheme == FSSB200_SCHEME_VDPF) ? p.in_bits : p.in_bits + 1;
  fssb200_params q = p;
  if (q.scheme == FSSB200_SCHEME_GROTTO) { q.group = FSSB200_GROUP_BYTES; q.mod_lo = q.mod_hi = 0; }
  c.gk = group_kind(q, &c.ga.vmask);
  const int mul = q.scheme == FSSB200_SCHEME_DCF ? 4 : (q.scheme == FSSB200_SCHEME_HALFTREE ? 1 : 2);
  if (q.prg == FSSB200_PRG_AES128_MMO) {
    for (int i = 0; i < 4; ++i) aes128_expand_le(q.prg_key + 16 * i, c.keys.rk[i]);
    const int nb = mul >= 2 ? mul / 2 : 1;
    for (int pi = 0; pi < 2 && pi < nb; ++pi)
      for (int i = 0; i < 44; ++i) c.keys.rkd[pi][i] = c.keys.rk[pi][i] ^ c.keys.rk[pi + nb][i];
  } else {
    std::memcpy(c.keys.nonce, q.prg_key, 8);
  }
  std::memcpy(c.keys.hash_key, q.hash_key, 16);
  std::memcpy(c.keys.hash_iv, q.hash_iv, 64);
  c.keys.hash_kind[0] = uint32_t(q.hash) & 0xffu;
  c.keys.hash_kind[1] = (uint32_t(q.hash) >> 8) & 0xffu;
  c.ga.mod[0] = uint32_t(q.mod_lo); c.ga.mod[1] = uint32_t(q.mod_lo >> 32);
  c.ga.mod[2] = uint32_t(q.mod_hi); c.ga.mod[3] = uint32_t(q.mod_hi >> 32);
}

template <int PRG>
static typename Prg<PRG>::ctx_t lane_ctx(uint64_t k);
template <>
AesCtx lane_ctx<kPrgAes>(uint64_t k) { return AesCtx{uint32_t(k & 31u) << 2}; }
template <>
NoCtx lane_ctx<kPrgChaCha>(uint64_t) { return NoCtx{}; }

struct Bufs {
  const blk *seeds; const uint8_t *cws; const blk *ocws; const uint8_t *xs; blk *ys;
  const blk *cw_s, *cw_v; const uint32_t *extra; const blk *out_cw;
  const blk *s0s; const blk *betas; uint8_t *cws_out; blk *ocws_out;
  uint8_t *all_out; uint64_t leaf_begin, leaf_count;
  uint64_t nkeys; int party; bool level_major;
  const blk *cs; blk *pis; blk *cs_out; int32_t *status;   // VDPF
};

template <int G, int PRG>
static void eval_t(const EmuCtx &c, const Bufs &b) {
  for (uint64_t k = 0; k < b.nkeys; ++k) {
    const auto pc = lane_ctx<PRG>(k);
    const InVal x = load_in(b.xs + k * c.in_bytes, c.in_bytes);
    const blk s0 = b.seeds[k];
    blk y;
    if (b.level_major) {
      const CwLevelMajor cw{b.cw_s, b.cw_v, b.extra, b.out_cw, b.nkeys, k};
      if (c.scheme == FSSB200_SCHEME_DPF) y = dpf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw);
      else if (c.scheme == FSSB200_SCHEME_DCF) y = dcf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw);
      else y = ht_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw, b.ocws[k]);
    } else {
      const CwKeyMajor cw{b.cws + k * uint64_t(c.ncw) * 32};
      if (c.scheme == FSSB200_SCHEME_DPF) y = dpf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw);
      else if (c.scheme == FSSB200_SCHEME_DCF) y = dcf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw);
      else y = ht_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, s0, x, cw, b.ocws[k]);
    }
    b.ys[k] = y;
  }
}

template <int G, int PRG>
static void gen_t(const EmuCtx &c, const Bufs &b) {
  for (uint64_t k = 0; k < b.nkeys; ++k) {
    const auto pc = lane_ctx<PRG>(k);
    const InVal a = load_in(b.xs + k * c.in_bytes, c.in_bytes);
    const blk beta = b.betas ? b.betas[k] : zero_blk();
    uint8_t *cws = b.cws_out + k * uint64_t(c.ncw) * 32;
    if (c.scheme == FSSB200_SCHEME_DPF || c.scheme == FSSB200_SCHEME_GROTTO)
      dpf_gen_body<G, PRG>(c.keys, c.ga, pc, c.n, b.s0s[2 * k], b.s0s[2 * k + 1], a, beta, CwOutKeyMajor{cws});
    else if (c.scheme == FSSB200_SCHEME_DCF)
      dcf_gen_body<G, PRG>(c.keys, c.ga, pc, c.n, c.pred, b.s0s[2 * k], b.s0s[2 * k + 1], a, beta, CwOutKeyMajor{cws});
    else
      ht_gen_body<G, PRG>(c.keys, c.ga, pc, c.n, b.s0s[2 * k], b.s0s[2 * k + 1], a, beta, CwOutKeyMajor{cws}, &b.ocws_out[k]);
  }
}

template <int G, int PRG>
static void vdpf_eval_t(const EmuCtx &c, const Bufs &b) {
  for (uint64_t k = 0; k < b.nkeys; ++k) {
    const auto pc = lane_ctx<PRG>(k);
    const InVal x = load_in(b.xs + k * c.in_bytes, c.in_bytes);
    if (b.level_major) {
      const CwLevelMajor cw{b.cw_s, nullptr, b.extra, nullptr, b.nkeys, k};
      b.ys[k] = vdpf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, b.seeds[k], x, cw, b.ocws[k], b.cs + 4 * k,
          b.pis + 4 * k);
    } else {
      const CwKeyMajor cw{b.cws + k * uint64_t(c.ncw) * 32};
      b.ys[k] = vdpf_eval_body<G, PRG>(c.keys, c.ga, pc, c.n, b.party, b.seeds[k], x, cw, b.ocws[k], b.cs + 4 * k,
          b.pis + 4 * k);
    }
  }
}
template <int G, int PRG>
static void vdpf_gen_t(const EmuCtx &c, const Bufs &b) {
  for (uint64_t k = 0; k < b.nkeys; ++k) {
    const auto pc = lane_ctx<PRG>(k);
    const InVal a = load_in(b.xs + k * c.in_bytes, c.in_bytes);
    b.status[k] = vdpf_gen_body<G, PRG>(c.keys, c.ga, pc, c.n, b.s0s[2 * k], b.s0s[2 * k + 1], a, b.betas[k],
        CwOutKeyMajor{b.cws_out + k * uint64_t(c.ncw) * 32}, b.cs_out + 4 * k, b.ocws_out + k);
  }
}

// Full-domain recursion over the same node-expansion / leaf functions the evalall kernel uses.
template <int G, int PRG>
static void tree_t(const EmuCtx &c, const Bufs &b, uint64_t k, blk st, int lvl, uint64_t l, uint64_t r) {
  if (r <= b.leaf_begin || l >= b.leaf_begin + b.leaf_count) return;
  const auto pc = lane_ctx<PRG>(l);
  const uint8_t *kc = b.cws + k * uint64_t(c.ncw) * 32;
  const CwKeyMajor cw{kc};
  const bool half = c.scheme == FSSB200_SCHEME_HALFTREE;
  const uint64_t out0 = k * b.leaf_count - b.leaf_begin;
  if (half && lvl == c.n - 1) {
    const blk cwl = cw.s(lvl);
    for (uint32_t sg = 0; sg < 2; ++sg) {
      const uint64_t x = l + sg;
      if (x < b.leaf_begin || x >= b.leaf_begin + b.leaf_count) continue;
      const blk y = ht_last<G, PRG>(c.keys, c.ga, pc, b.party, st, sg, cwl, sg ? cw.flag(lvl) : (cwl.w & 1u), b.ocws[k]);
      std::memcpy(b.all_out + (out0 + x) * 16, &y, 16);
    }
    return;
  }
  if (!half && lvl == c.n) {
    if (c.scheme == FSSB200_SCHEME_GROTTO) {
      b.all_out[out0 + l] = uint8_t(lsb(st));
    } else {
      const blk y = dpf_leaf<G>(c.ga, b.party, st, cw.s(c.n));
      std::memcpy(b.all_out + (out0 + l) * 16, &y, 16);
    }
    return;
  }
  blk left, right;
  if (half) {
    ht_expand<PRG>(c.keys, pc, st, cw.s(lvl), left, right);
  } else {
    const blk cs = cw.s(lvl);
    blk cr = cs;
    cr.w = (cs.w & ~1u) | cw.flag(lvl);
    dpf_expand<PRG>(c.keys, pc, st, cs, cr, left, right);
  }
  const uint64_t mid = l + ((r - l) >> 1);
  tree_t<G, PRG>(c, b, k, left, lvl + 1, l, mid);
  tree_t<G, PRG>(c, b, k, right, lvl + 1, mid, r);
}
template <int G, int PRG>
static void evalall_t(const EmuCtx &c, const Bufs &b) {
  for (uint64_t k = 0; k < b.nkeys; ++k) {
    blk st = clamp(b.seeds[k]);
    st.w |= uint32_t(b.party);
    tree_t<G, PRG>(c, b, k, st, 0, 0, uint64_t(1) << c.n);
  }
}

template <template <int, int> class F>
struct Dispatch;

#define DISPATCH(FN, c, b)                                                     \
  do {                                                                         \
    const bool aes = (c).prg == FSSB200_PRG_AES128_MMO;                        \
    switch ((c).gk) {                                                          \
      case kGrpBytes: aes ? FN<kGrpBytes, kPrgAes>(c, b) : FN<kGrpBytes, kPrgChaCha>(c, b); break;       \
      case kGrpU32: aes ? FN<kGrpU32, kPrgAes>(c, b) : FN<kGrpU32, kPrgChaCha>(c, b); break;             \
      case kGrpU64: aes ? FN<kGrpU64, kPrgAes>(c, b) : FN<kGrpU64, kPrgChaCha>(c, b); break;             \
      case kGrpU127: aes ? FN<kGrpU127, kPrgAes>(c, b) : FN<kGrpU127, kPrgChaCha>(c, b); break;          \
      case kGrpU32Mod: aes ? FN<kGrpU32Mod, kPrgAes>(c, b) : FN<kGrpU32Mod, kPrgChaCha>(c, b); break;    \
      case kGrpU64Mod: aes ? FN<kGrpU64Mod, kPrgAes>(c, b) : FN<kGrpU64Mod, kPrgChaCha>(c, b); break;    \
      default: aes ? FN<kGrpU128Mod, kPrgAes>(c, b) : FN<kGrpU128Mod, kPrgChaCha>(c, b); break;          \
    }                                                                          \
  } while (0)

}  // namespace fssb200

using namespace fssb200;

extern "C" {

int emul_prg_gen(const fssb200_params *p, int mul, size_t n, const void *seeds, void *out) {
  EmuCtx c;
  make_ctx(*p, c);
  const blk *s = static_cast<const blk *>(seeds);
  blk *o = static_cast<blk *>(out);
  for (size_t i = 0; i < n; ++i) {
    if (p->prg == FSSB200_PRG_AES128_MMO) {
      const AesCtx pc = lane_ctx<kPrgAes>(i);
      if (mul == 1) Prg<kPrgAes>::gen<1>(c.keys, pc, s[i], o + i);
      else if (mul == 2) Prg<kPrgAes>::gen<2>(c.keys, pc, s[i], o + 2 * i);
      else Prg<kPrgAes>::gen<4>(c.keys, pc, s[i], o + 4 * i);
    } else {
      if (mul == 1) Prg<kPrgChaCha>::gen<1>(c.keys, NoCtx{}, s[i], o + i);
      else if (mul == 2) Prg<kPrgChaCha>::gen<2>(c.keys, NoCtx{}, s[i], o + 2 * i);
      else Prg<kPrgChaCha>::gen<4>(c.keys, NoCtx{}, s[i], o + 4 * i);
    }
  }
  return 0;
}

int emul_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas, const void *betas,
    void *cws, void *ocws) {
  EmuCtx c;
  make_ctx(*p, c);
  Bufs b{};
  b.nkeys = nkeys;
  b.s0s = static_cast<const blk *>(s0s);
  b.xs = static_cast<const uint8_t *>(alphas);
  b.betas = p->scheme == FSSB200_SCHEME_GROTTO ? nullptr : static_cast<const blk *>(betas);
  b.cws_out = static_cast<uint8_t *>(cws);
  b.ocws_out = static_cast<blk *>(ocws);
  DISPATCH(gen_t, c, b);
  return 0;
}

int emul_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *ocws, const void *xs, void *ys, int level_major, const void *cw_s, const void *cw_v,
    const void *extra, const void *out_cw) {
  EmuCtx c;
  make_ctx(*p, c);
  Bufs b{};
  b.nkeys = nkeys;
  b.party = party;
  b.seeds = static_cast<const blk *>(seeds);
  b.cws = static_cast<const uint8_t *>(cws);
  b.ocws = static_cast<const blk *>(ocws);
  b.xs = static_cast<const uint8_t *>(xs);
  b.ys = static_cast<blk *>(ys);
  b.level_major = level_major != 0;
  b.cw_s = static_cast<const blk *>(cw_s);
  b.cw_v = static_cast<const blk *>(cw_v);
  b.extra = static_cast<const uint32_t *>(extra);
  b.out_cw = static_cast<const blk *>(out_cw);
  DISPATCH(eval_t, c, b);
  return 0;
}

int emul_evalall(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *ocws, void *ys, uint64_t leaf_begin, uint64_t leaf_count) {
  EmuCtx c;
  make_ctx(*p, c);
  if (p->scheme == FSSB200_SCHEME_DCF) return FSSB200_ESCHEME;
  Bufs b{};
  b.nkeys = nkeys;
  b.party = party;
  b.seeds = static_cast<const blk *>(seeds);
  b.cws = static_cast<const uint8_t *>(cws);
  b.ocws = static_cast<const blk *>(ocws);
  b.all_out = static_cast<uint8_t *>(ys);
  b.leaf_begin = leaf_begin;
  b.leaf_count = leaf_count ? leaf_count : ((uint64_t(1) << p->in_bits) - leaf_begin);
  DISPATCH(evalall_t, c, b);
  return 0;
}

// Grotto O(n) point walk (schemes.cuh: grotto_walk_body): ys[k] = one share bit
int emul_grotto_walk(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *xs, void *ys) {
  EmuCtx c;
  make_ctx(*p, c);
  if (p->scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  const blk *sd = static_cast<const blk *>(seeds);
  for (size_t k = 0; k < nkeys; ++k) {
    const InVal x = load_in(static_cast<const uint8_t *>(xs) + k * c.in_bytes, c.in_bytes);
    const CwKeyMajor cw{static_cast<const uint8_t *>(cws) + k * uint64_t(c.ncw) * 32};
    uint32_t bit;
    if (p->prg == FSSB200_PRG_AES128_MMO)
      bit = grotto_walk_body<kPrgAes>(c.keys, lane_ctx<kPrgAes>(k), c.n, c.in_bytes, uint32_t(party), sd[k], x, cw);
    else
      bit = grotto_walk_body<kPrgChaCha>(c.keys, NoCtx{}, c.n, c.in_bytes, uint32_t(party), sd[k], x, cw);
    static_cast<uint8_t *>(ys)[k] = uint8_t(bit);
  }
  return 0;
}

// The tile accessors of kernels.cuh (CwTileT / CwLmTile) fetch chunk c of a key's correction words when the scheme body calls
// begin_level(j) with j == c << lpc_bits, and request the NEXT chunk (or the next tile's chunk 0) at the same time: a body that
// skips a chunk boundary, visits one twice or stops before the last chunk leaves the prefetch pipeline out of step with the next
// tile (the Grotto walk did, before its trailing begin_level(n)).  This records the boundaries a body really visits.
struct CwCounting {
  CwKeyMajor inner;
  int lpc_bits;
  int *chunks;  // out: chunk index of every boundary visited, in call order
  int *count;
  int cap;
  FSS_HD blk s(int i) const { return inner.s(i); }
  FSS_HD blk v(int i) const { return inner.v(i); }
  FSS_HD uint32_t flag(int i) const { return inner.flag(i); }
  FSS_HD blk out_s(int n) const { return inner.out_s(n); }
  FSS_HD blk out_v(int n) const { return inner.out_v(n); }
  FSS_HD void begin_level(int j) const {
    if (j & ((1 << lpc_bits) - 1)) return;
    if (*count < cap) chunks[*count] = j >> lpc_bits;
    ++*count;
  }
  FSS_HD void done_level(int) const {}
};
// Runs the point-evaluation body of p's scheme on one all-zero key and returns how many chunk boundaries it visited
// (chunks[0..) = their indices).  Bytes group, the scheme's PRG.
int emul_chunk_sequence(const fssb200_params *p, int lpc_bits, int *chunks, int cap) {
  EmuCtx c;
  make_ctx(*p, c);
  std::vector<uint8_t> key(size_t(c.ncw + 1) * 32, 0);
  blk cs[4] = {zero_blk(), zero_blk(), zero_blk(), zero_blk()}, pi[4];
  int count = 0;
  const CwCounting cw{CwKeyMajor{key.data()}, lpc_bits, chunks, &count, cap};
  const InVal x = load_in(key.data(), c.in_bytes);
  const blk s0 = zero_blk();
  const auto pc = lane_ctx<kPrgChaCha>(0);
  switch (p->scheme) {
    case FSSB200_SCHEME_DPF: dpf_eval_body<kGrpBytes, kPrgChaCha>(c.keys, c.ga, pc, c.n, 0u, s0, x, cw); break;
    case FSSB200_SCHEME_DCF: dcf_eval_body<kGrpBytes, kPrgChaCha>(c.keys, c.ga, pc, c.n, 0u, s0, x, cw); break;
    case FSSB200_SCHEME_HALFTREE: ht_eval_body<kGrpBytes, kPrgChaCha>(c.keys, c.ga, pc, c.n, 0u, s0, x, cw, zero_blk()); break;
    case FSSB200_SCHEME_VDPF: vdpf_eval_body<kGrpBytes, kPrgChaCha>(c.keys, c.ga, pc, c.n, 0u, s0, x, cw, zero_blk(), cs, pi); break;
    case FSSB200_SCHEME_GROTTO: grotto_walk_body<kPrgChaCha>(c.keys, pc, c.n, c.in_bytes, 0u, s0, x, cw); break;
    default: return -1;
  }
  return count;
}

int emul_hash(const fssb200_params *p, int which, size_t n, const void *msgs, void *out) {
  EmuCtx c;
  make_ctx(*p, c);
  const blk *m = static_cast<const blk *>(msgs);
  blk *o = static_cast<blk *>(out);
  for (size_t i = 0; i < n; ++i) {
    if (which == 0) vdpf_xor_hash(c.keys, m[2 * i], m[2 * i + 1], o + 4 * i);
    else vdpf_hash(c.keys, m + 4 * i, o + 2 * i);
  }
  return 0;
}

int emul_vdpf_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas, const void *betas,
    void *cws, void *cs, void *ocws, void *status) {
  EmuCtx c;
  make_ctx(*p, c);
  Bufs b{};
  b.nkeys = nkeys;
  b.s0s = static_cast<const blk *>(s0s);
  b.xs = static_cast<const uint8_t *>(alphas);
  b.betas = static_cast<const blk *>(betas);
  b.cws_out = static_cast<uint8_t *>(cws);
  b.cs_out = static_cast<blk *>(cs);
  b.ocws_out = static_cast<blk *>(ocws);
  b.status = static_cast<int32_t *>(status);
  DISPATCH(vdpf_gen_t, c, b);
  return 0;
}

int emul_vdpf_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *cs, const void *ocws, const void *xs, void *ys, void *pis, int level_major, const void *cw_s,
    const void *extra) {
  EmuCtx c;
  make_ctx(*p, c);
  Bufs b{};
  b.nkeys = nkeys;
  b.party = party;
  b.seeds = static_cast<const blk *>(seeds);
  b.cws = static_cast<const uint8_t *>(cws);
  b.cs = static_cast<const blk *>(cs);
  b.ocws = static_cast<const blk *>(ocws);
  b.xs = static_cast<const uint8_t *>(xs);
  b.ys = static_cast<blk *>(ys);
  b.pis = static_cast<blk *>(pis);
  b.level_major = level_major != 0;
  b.cw_s = static_cast<const blk *>(cw_s);
  b.extra = static_cast<const uint32_t *>(extra);
  DISPATCH(vdpf_eval_t, c, b);
  return 0;
}

int emul_vdpf_prove(const fssb200_params *p, size_t nkeys, size_t m, const void *pi_tildes, const void *cs, void *pis) {
  EmuCtx c;
  make_ctx(*p, c);
  const blk *pt = static_cast<const blk *>(pi_tildes);
  for (size_t k = 0; k < nkeys; ++k) {
    blk pi[4];
    std::memcpy(pi, static_cast<const blk *>(cs) + 4 * k, 64);
    for (size_t i = 0; i < m; ++i) vdpf_accumulate(c.keys, pi, pt + 4 * (k * m + i));
    std::memcpy(static_cast<blk *>(pis) + 4 * k, pi, 64);
  }
  return 0;
}

}  // extern "C"
