// SPDX-License-Identifier: Apache-2.0
// TEST (CPU).  The code the header shim runs for DEVICE callers -- fss::prg::aes_detail::{ExpandKey, EncryptMmo} (Aes128Soft on
// the caller's Te0 / S-box tables), fss::hash::b200_detail::{Hash64, HashPair} (BLAKE3), fss::prg::b200_detail::ChaChaBlock and
// the plugin-generic fss::b200::generic::{VdpfGen, VdpfEval} -- is `__host__ __device__`; compiled for the host here, the very
// same functions are compared bit for bit with the oracle (parity-pinned against the reference) and with the survey's
// known answers.  The plugin objects below do what the shim's plugin classes do under `#if defined(__CUDA_ARCH__)`.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include <fss/b200/generic.cuh>
#include <fss/group/bytes.cuh>
#include <fss/group/uint.cuh>
#include <fss/hash/blake3.cuh>
#include <fss/prg/aes128_mmo_soft.cuh>
#include <fss/prg/chacha.cuh>
#include "../../oracle/fss_oracle.h"

static int g_bad = 0;
#define CHECK(cond, ...)                 \
  do {                                   \
    if (!(cond)) {                       \
      std::printf("FAIL: " __VA_ARGS__); \
      std::printf("\n");                 \
      ++g_bad;                           \
    }                                    \
  } while (0)

static uint64_t g_state = 12345;
static uint32_t Next() {
  uint64_t z = (g_state += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return uint32_t((z ^ (z >> 31)) >> 32);
}
static int4 Block(bool clamp) { return int4{int(Next()), int(Next()), int(Next()), int(clamp ? Next() & ~1u : Next())}; }

static uint32_t g_te0[256];
static uint8_t g_sbox[256];

// ---- plugins with the device-path bodies ----------------------------------------------------------------------------
template <int mul>
struct SoftAes {  // = Aes128Soft<mul> under __CUDA_ARCH__
  uint32_t rk[mul][44];
  explicit SoftAes(const uint8_t keys[][16]) {
    for (int i = 0; i < mul; ++i) fss::prg::aes_detail::ExpandKey(keys[i], g_sbox, rk[i]);
  }
  cuda::std::array<int4, mul> Gen(int4 seed) const {
    cuda::std::array<int4, mul> out{};
    for (int i = 0; i < mul; ++i) out[i] = fss::prg::aes_detail::EncryptMmo(rk[i], g_te0, g_sbox, seed);
    return out;
  }
};
template <int mul>
struct DevChaCha {  // = ChaCha<mul> under __CUDA_ARCH__
  int n0, n1;
  cuda::std::array<int4, mul> Gen(int4 seed) const { return fss::prg::b200_detail::ChaChaBlock<mul>(seed, n0, n1); }
};
struct DevBlake3 {  // = hash::Blake3 under __CUDA_ARCH__
  int4 iv[2];
  cuda::std::array<int4, 2> Hash(cuda::std::span<const int4, 4> msg) const { return fss::hash::b200_detail::Hash64(iv, msg.data()); }
  cuda::std::array<int4, 4> Hash(cuda::std::tuple<int4, const int4> msg) const {
    return fss::hash::b200_detail::HashPair(iv, cuda::std::get<0>(msg), cuda::std::get<1>(msg));
  }
};

static fssb200_params Params(int scheme, int in_bits, int in_bytes, int group, uint64_t mod_lo, uint64_t mod_hi, int prg, const void *key,
    size_t key_bytes, const int4 iv[2]) {
  fssb200_params p;
  std::memset(&p, 0, sizeof(p));
  p.scheme = scheme;
  p.in_bits = in_bits;
  p.in_bytes = in_bytes;
  p.group = group;
  p.mod_lo = mod_lo;
  p.mod_hi = mod_hi;
  p.prg = prg;
  std::memcpy(p.prg_key, key, key_bytes);
  if (iv) {
    std::memcpy(p.hash_iv[0], iv, 32);
    std::memcpy(p.hash_iv[1], iv, 32);
  }
  return p;
}

// ---- VDPF through the generic templates vs the oracle ----------------------------------------------------------------------
struct VCw {
  int4 s;
  bool tr;
  char pad[15];
};
static_assert(sizeof(VCw) == 32);

template <int in_bits, typename In, typename Group, typename Prg>
static void VdpfCase(const char *name, Prg prg, const fssb200_params &p, int keys) {
  DevBlake3 xh;
  std::memcpy(xh.iv, p.hash_iv[0], 32);
  int retries = 0;
  for (int k = 0; k < keys; ++k) {
    const int4 s0s[2] = {Block(true), Block(true)};
    const In a = static_cast<In>((uint64_t(Next()) << 32 | Next()) & (in_bits >= 64 ? ~uint64_t(0) : ((uint64_t(1) << in_bits) - 1)));
    const int4 beta = Block(true);
    std::vector<VCw> cws(in_bits), ocws_cws(in_bits);
    cuda::std::array<int4, 4> cs{}, o_cs{};
    int4 ocw{0, 0, 0, 0}, o_ocw{0, 0, 0, 0};
    int32_t o_status = 0;
    std::memset(cws.data(), 0xAB, sizeof(VCw) * in_bits);
    const int status = fss::b200::generic::VdpfGen<in_bits, Group, In>(prg, xh, cws.data(), cs, ocw, s0s, a, beta);
    const int rc = orc_vdpf_gen(&p, 1, s0s, &a, &beta, ocws_cws.data(), o_cs.data(), &o_ocw, &o_status, 1);
    CHECK(rc == 0 && status == o_status, "%s key %d: gen status %d vs %d (rc %d)", name, k, status, o_status, rc);
    bool same = std::memcmp(cs.data(), o_cs.data(), 64) == 0;
    for (int i = 0; i < in_bits; ++i) same &= std::memcmp(&cws[i].s, &ocws_cws[i].s, 16) == 0 && cws[i].tr == ocws_cws[i].tr;
    if (status == 0) same &= std::memcmp(&ocw, &o_ocw, 16) == 0;
    CHECK(same, "%s key %d: generated key differs from the oracle's", name, k);
    for (int i = 0; i < in_bits; ++i) {  // the padding of a slot is defined (zero) here
      const uint8_t *raw = reinterpret_cast<const uint8_t *>(&cws[i]);
      for (int j = 17; j < 32; ++j) same &= raw[j] == 0;
    }
    CHECK(same, "%s key %d: Cw padding not zero", name, k);
    if (status != 0) {
      ++retries;
      continue;
    }
    for (int party = 0; party < 2; ++party)
      for (int trial = 0; trial < 6; ++trial) {
        const In x = trial == 0 ? a : static_cast<In>((uint64_t(Next()) << 32 | Next()) & (in_bits >= 64 ? ~uint64_t(0) : ((uint64_t(1) << in_bits) - 1)));
        int4 y, o_y;
        cuda::std::array<int4, 4> o_pi{};
        const auto pi = fss::b200::generic::VdpfEval<in_bits, Group, In>(prg, xh, party != 0, s0s[party], cws.data(), cs.data(), ocw, x, y);
        const int rc2 = orc_vdpf_eval(&p, party, 1, &s0s[party], ocws_cws.data(), o_cs.data(), &o_ocw, &x, &o_y, o_pi.data(), 1);
        CHECK(rc2 == 0 && std::memcmp(&y, &o_y, 16) == 0 && std::memcmp(pi.data(), o_pi.data(), 64) == 0, "%s key %d party %d trial %d: eval differs",
            name, k, party, trial);
      }
  }
  std::printf("%s: %d keys (%d with the retry status), both parties x 6 points each\n", name, keys, retries);
}

int main() {
  fss::prg::aes_detail::InitTe0(g_te0);
  fss::prg::aes_detail::InitSbox(g_sbox);

  // ---- soft AES: the survey's known answer (Aes128Mmo<4>.Gen(seed0)), then random keys / seeds vs the oracle
  uint8_t keys[4][16];
  for (int j = 0; j < 16; ++j) {
    keys[0][j] = uint8_t(j + 1);
    keys[1][j] = uint8_t(16 - j);
    keys[2][j] = uint8_t(j / 2 + 1);
    keys[3][j] = uint8_t(8 - j / 2);
  }
  {
    const int4 seed0{0x11111111, 0x22222222, 0x33333333, 0x44444440};
    const uint32_t want[4][4] = {{0x1423e6d2, 0x60533602, 0x813b3fbc, 0x412b31dc}, {0x2d18cbe7, 0xa5eddcc9, 0xc691e4f6, 0x97831704},
                                 {0x2c046c24, 0x8a033811, 0xcb0335b6, 0x3d799ddb}, {0x2413b69f, 0xffed7ac4, 0x314804c9, 0x595d2580}};
    SoftAes<4> prg(keys);
    const auto out = prg.Gen(seed0);
    CHECK(std::memcmp(out.data(), want, 64) == 0, "soft AES: survey known answer");
  }
  for (int round = 0; round < 50; ++round) {
    uint8_t rkeys[4][16];
    for (auto &k : rkeys)
      for (auto &b : k) b = uint8_t(Next());
    const fssb200_params p = Params(FSSB200_SCHEME_DCF, 8, 1, FSSB200_GROUP_BYTES, 0, 0, FSSB200_PRG_AES128_MMO, rkeys, 64, nullptr);
    SoftAes<4> prg(rkeys);
    std::vector<int4> seeds(64), want(64 * 4);
    for (auto &s : seeds) s = Block(false);
    CHECK(orc_prg_gen(&p, 4, seeds.size(), seeds.data(), want.data()) == 0, "orc_prg_gen");
    bool same = true;
    for (size_t i = 0; i < seeds.size(); ++i) same &= std::memcmp(prg.Gen(seeds[i]).data(), &want[4 * i], 64) == 0;
    CHECK(same, "soft AES round %d differs from the oracle", round);
  }
  std::printf("soft AES (caller's Te0 + S-box): survey known answer + 50 key sets x 64 seeds x 4 blocks == oracle\n");

  // ---- BLAKE3: both plugin interfaces vs the oracle
  for (int round = 0; round < 40; ++round) {
    const int4 iv[2] = {Block(false), Block(false)};
    const int nonce[2] = {0x12345678, int(0x9abcdef0u)};
    const fssb200_params p = Params(FSSB200_SCHEME_VDPF, 8, 1, FSSB200_GROUP_BYTES, 0, 0, FSSB200_PRG_CHACHA, nonce, 8, iv);
    const size_t n = 32;
    std::vector<int4> pairs(2 * n), msgs(4 * n), want4(4 * n), want2(2 * n);
    for (auto &b : pairs) b = Block(false);
    for (auto &b : msgs) b = Block(false);
    CHECK(orc_hash(&p, 0, n, pairs.data(), want4.data()) == 0 && orc_hash(&p, 1, n, msgs.data(), want2.data()) == 0, "orc_hash");
    bool same = true;
    for (size_t i = 0; i < n; ++i) {
      same &= std::memcmp(fss::hash::b200_detail::HashPair(iv, pairs[2 * i], pairs[2 * i + 1]).data(), &want4[4 * i], 64) == 0;
      same &= std::memcmp(fss::hash::b200_detail::Hash64(iv, &msgs[4 * i]).data(), &want2[2 * i], 32) == 0;
    }
    CHECK(same, "BLAKE3 round %d differs from the oracle", round);
  }
  std::printf("BLAKE3 (Hashable 64 B -> 32 B, XorHashable (a, b) -> 64 B): 40 IVs x 32 messages == oracle\n");

  // ---- VDPF Gen / Eval through the plugin-generic templates
  const int4 iv[2] = {{0x11111111, 0x22222222, 0x33333333, 0x44444444}, {0x55555555, 0x66666666, 0x77777777, int(0x88888888u)}};
  const int nonce[2] = {0x12345678, int(0x9abcdef0u)};
  const uint64_t top = uint64_t(1) << 63;
  using U64 = fss::group::Uint<uint64_t>;
  using U127 = fss::group::Uint<__uint128_t, (static_cast<__uint128_t>(1) << 127)>;
  using U32p = fss::group::Uint<uint32_t, 4294967291u>;
  VdpfCase<8, uint8_t, fss::group::Bytes>("vdpf n=8 bytes chacha", DevChaCha<2>{nonce[0], nonce[1]},
      Params(FSSB200_SCHEME_VDPF, 8, 1, FSSB200_GROUP_BYTES, 0, 0, FSSB200_PRG_CHACHA, nonce, 8, iv), 40);
  VdpfCase<20, uint32_t, U64>("vdpf n=20 u64 chacha (the reference bench's parameters)", DevChaCha<2>{nonce[0], nonce[1]},
      Params(FSSB200_SCHEME_VDPF, 20, 4, FSSB200_GROUP_U64, 0, 0, FSSB200_PRG_CHACHA, nonce, 8, iv), 40);
  VdpfCase<33, uint64_t, U127>("vdpf n=33 u127 soft AES", SoftAes<2>(keys),
      Params(FSSB200_SCHEME_VDPF, 33, 8, FSSB200_GROUP_U128, 0, top, FSSB200_PRG_AES128_MMO, keys, 32, iv), 25);
  VdpfCase<64, uint64_t, fss::group::Bytes>("vdpf n=64 bytes soft AES", SoftAes<2>(keys),
      Params(FSSB200_SCHEME_VDPF, 64, 8, FSSB200_GROUP_BYTES, 0, 0, FSSB200_PRG_AES128_MMO, keys, 32, iv), 15);
  VdpfCase<1, uint8_t, U32p>("vdpf n=1 u32 mod 2^32-5 chacha", DevChaCha<2>{nonce[0], nonce[1]},
      Params(FSSB200_SCHEME_VDPF, 1, 1, FSSB200_GROUP_U32, 4294967291u, 0, FSSB200_PRG_CHACHA, nonce, 8, iv), 40);

  // ---- DPF through the generic templates with the soft AES (what the reference bench's DpfGenKernelAes / DpfEvalKernelAes run)
  {
    struct DCw {
      int4 s;
      bool tr;
      char pad[15];
    };
    const fssb200_params p = Params(FSSB200_SCHEME_DPF, 20, 4, FSSB200_GROUP_U64, 0, 0, FSSB200_PRG_AES128_MMO, keys, 32, nullptr);
    SoftAes<2> prg(keys);
    bool same = true;
    for (int k = 0; k < 40; ++k) {
      const int4 s0s[2] = {Block(true), Block(true)};
      const uint32_t a = Next() & 0xfffff;
      const int4 beta = Block(true);
      std::vector<DCw> cws(21), want(21);
      fss::b200::generic::DpfGen<20, U64, uint32_t>(prg, cws.data(), s0s, a, beta);
      CHECK(orc_gen(&p, 1, s0s, &a, &beta, want.data(), nullptr, 1) == 0, "orc_gen");
      for (int i = 0; i < 21; ++i) same &= std::memcmp(&cws[i].s, &want[i].s, 16) == 0 && (i == 20 || cws[i].tr == want[i].tr);
      for (int party = 0; party < 2; ++party)
        for (uint32_t x : {a, uint32_t(Next() & 0xfffff)}) {
          int4 o_y;
          const int4 y = fss::b200::generic::DpfEval<20, U64, uint32_t>(prg, party != 0, s0s[party], cws.data(), x);
          CHECK(orc_eval(&p, party, 1, &s0s[party], want.data(), nullptr, &x, &o_y, 1) == 0, "orc_eval");
          same &= std::memcmp(&y, &o_y, 16) == 0;
        }
    }
    CHECK(same, "dpf n=20 u64 soft AES differs from the oracle");
    std::printf("dpf n=20 u64 soft AES: 40 keys, both parties == oracle\n");
  }

  if (g_bad == 0) std::printf("shim device paths: all checks passed\n");
  return g_bad ? 1 : 0;
}
