// SPDX-License-Identifier: Apache-2.0
//
// schemes.cuh -- per-thread bodies of the DPF / DCF / Half-Tree walks and key generation.
//
// Node representation: a tree node is ONE packed 16-byte block, seed with the control bit t in
// the clamp bit (the reference's `st`, dpf.cuh:233-236, eval_all_gpu.cuh:154-178).  Because the
// stored correction word already carries tl_cw in its clamp bit (dpf.cuh:148), the whole
// "if (t) { s ^= s_cw; t ^= t_cw }" update of dpf.cuh:189-194 is one masked XOR of the packed
// block: child = G(s) ^ (-(t) & cw'), with cw' = s_cw whose clamp bit is tl_cw (left) or tr_cw
// (right).  4 LOP3 per child.
//
// Everything here is FSS_HD: the same code runs per CUDA thread and, in tests/host_emul, per key
// on the CPU.
#pragma once
#include "blake3.cuh"
#include "sha256.cuh"
#include "group.cuh"
#include "prg.cuh"

// A node expansion can LOOP over its AES blocks instead of inlining one copy per block: a copy is ~7 KB of SASS and the
// instruction cache holds 32 KB (B300_MICROARCH "I-cache").  Measured on a B200 (profiles/r01_evalall_coop.md):
//   DCF Gen level, 8 blocks  (58 KB inlined): looped 2 blocks / iteration  0.78 -> 0.97 of the LDS ceiling
//   DCF EvalAll node, 4 blocks (29 KB + u128 arithmetic): looped by side   0.854 -> 0.874 (u127), 0.861 -> 0.835 (Bytes)
//   DPF / Grotto node, 2 blocks (14 KB, fits): looped 1 block / iteration  0.912 -> 0.900 (less ILP) => stays inlined
#ifndef FSS_LOOPED_EXPAND
#define FSS_LOOPED_EXPAND 0   // DPF / Grotto node: 2 blocks
#endif
#ifndef FSS_LOOPED_DCF
#define FSS_LOOPED_DCF 1      // DCF node (4 blocks) and DCF Gen level (8 blocks)
#endif

namespace fssb200 {

FSS_HD blk ld_blk(const void *p) {
#if FSS_DEVICE_CODE
  const uint4 t = __ldg(reinterpret_cast<const uint4 *>(p));
#else
  const uint4 t = *reinterpret_cast<const uint4 *>(p);
#endif
  return make_blk(t.x, t.y, t.z, t.w);
}
FSS_HD void st_blk(void *p, blk b) { *reinterpret_cast<uint4 *>(p) = make_uint4(b.x, b.y, b.z, b.w); }

// ---- correction-word accessors ------------------------------------------------------------------------
// Key-major: the reference's own array-of-structs, 32 B per level (dpf.cuh:76-81, dcf.cuh:91-96).
struct CwKeyMajor {
  const uint8_t *base;  // this key's Cw[ncw]
  FSS_HD blk s(int i) const { return ld_blk(base + 32 * i); }
  FSS_HD blk v(int i) const { return ld_blk(base + 32 * i + 16); }
  // Dpf::Cw::tr / HalfTreeDpf::Cw::extra: a C++ bool at byte 16, tested as != 0
  FSS_HD uint32_t flag(int i) const {
#if FSS_DEVICE_CODE
    return __ldg(base + 32 * i + 16) != 0;
#else
    return base[32 * i + 16] != 0;
#endif
  }
  FSS_HD blk out_s(int n) const { return s(n); }  // cws[n].s  (DPF)
  FSS_HD blk out_v(int n) const { return v(n); }  // cws[n].v  (DCF)
  FSS_HD void begin_level(int) const {}           // (the staged accessor waits for the level here ...
  FSS_HD void done_level(int) const {}            //  ... and starts the next level's copy here)
};
// Level-major (fssb200_relayout; point_eval_gpu.cuh:39-91 with >32-level control words).
struct CwLevelMajor {
  const blk *cw_s;
  const blk *cw_v;
  const uint32_t *extra;
  const blk *out_cw;
  uint64_t nkeys, k;
  FSS_HD blk s(int i) const { return ld_blk(cw_s + uint64_t(i) * nkeys + k); }
  FSS_HD blk v(int i) const { return ld_blk(cw_v + uint64_t(i) * nkeys + k); }
  FSS_HD uint32_t flag(int i) const { return (extra[uint64_t(i >> 5) * nkeys + k] >> (i & 31)) & 1u; }
  FSS_HD blk out_s(int) const { return ld_blk(out_cw + k); }
  FSS_HD blk out_v(int) const { return ld_blk(out_cw + k); }
  FSS_HD void begin_level(int) const {}
  FSS_HD void done_level(int) const {}
};

// Gen output: entry i of this key's Cw array = {s, v} (32 B).  Key-major direct stores here; the gen kernel
// uses CwTileOut (kernels.cuh: tiles written by the TMA unit) with the same interface.
struct CwOutKeyMajor {
  uint8_t *base;
  FSS_HD void put(int i, blk s, blk v) const {
    st_blk(base + 32 * i, s);
    st_blk(base + 32 * i + 16, v);
  }
};

// ---- DPF ------------------------------------------------------------------------------------------------
// Leaf conversion, dpf.cuh:207-213 / :255-263.
template <int G>
FSS_HD blk dpf_leaf(const GroupArgs &ga, uint32_t party, blk st, blk out_cw) {
  typedef Grp<G> GR;
  typename GR::V y = GR::from(ga, clamp(st));
  y = GR::add_masked(ga, y, 0u - lsb(st), GR::from(ga, out_cw));
  y = GR::cneg(ga, y, party);
  return GR::into(ga, y);
}

// Dpf::Eval, dpf.cuh:170-214.  One PRG block per level: only the child on the path.
template <int G, int PRG, class Cw>
FSS_HD blk dpf_eval_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    uint32_t party, blk s0, const InVal &x, const Cw &cw) {
  blk st = clamp(s0);
  st.w |= party;                                   // t = b (dpf.cuh:173)
  cw.begin_level(0);
  blk cs = cw.s(0);
  uint32_t cf = cw.flag(0);
  cw.done_level(0);
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    // fetch the next level's correction word before this level's PRG call (entry n is the output CW)
    cw.begin_level(i + 1);
    const blk cs_next = (i + 1 < n) ? cw.s(i + 1) : cw.out_s(n);
    const uint32_t cf_next = (i + 1 < n) ? cw.flag(i + 1) : 0u;
    cw.done_level(i + 1);
    const uint32_t xb = in_bit(x, n - 1 - i);      // MSB first (dpf.cuh:196)
    const uint32_t tm = 0u - lsb(st);
    blk c[1];
    Prg<PRG>::template gen_child<1>(K, pc, clamp(st), xb, c);
    blk cwp = cs;                                  // clamp bit := (xb ? tr_cw : tl_cw)
    cwp.w = (cs.w & ~1u) | (xb ? cf : (cs.w & 1u));
    st = xor_masked(c[0], tm, cwp);
    cs = cs_next;
    cf = cf_next;
  }
  return dpf_leaf<G>(ga, party, st, cs);           // cs == cws[n].s
}

// Dpf::Gen, dpf.cuh:93-159.  Writes Cw[n+1] (32 B each, padding zeroed).
template <int G, int PRG, class Out>
FSS_HD void dpf_gen_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    blk s0, blk s1, const InVal &a, blk beta, const Out &out) {
  typedef Grp<G> GR;
  s0 = clamp(s0);
  s1 = clamp(s1);
  uint32_t t0 = 0, t1 = 1;
  beta = clamp(beta);
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    blk g0[2], g1[2];
    Prg<PRG>::template gen<2>(K, pc, s0, g0);
    Prg<PRG>::template gen<2>(K, pc, s1, g1);
    const uint32_t ab = in_bit(a, n - 1 - i);
    const uint32_t am = 0u - ab;
    // keep = child on alpha's path, lose = the other one
    const blk lose0 = xor_masked(g0[1], am, g0[0] ^ g0[1]);   // ab ? g0[0] : g0[1]
    const blk lose1 = xor_masked(g1[1], am, g1[0] ^ g1[1]);
    const blk keep0 = xor_masked(g0[0], am, g0[0] ^ g0[1]);   // ab ? g0[1] : g0[0]
    const blk keep1 = xor_masked(g1[0], am, g1[0] ^ g1[1]);
    blk s_cw = clamp(lose0 ^ lose1);                           // dpf.cuh:115-117
    const uint32_t tl_cw = (lsb(g0[0]) ^ lsb(g1[0]) ^ ab ^ 1u) & 1u;  // :119
    const uint32_t tr_cw = (lsb(g0[1]) ^ lsb(g1[1]) ^ ab) & 1u;       // :120
    const uint32_t tk_cw = ab ? tr_cw : tl_cw;
    const blk ns0 = xor_masked(clamp(keep0), 0u - t0, s_cw);
    const blk ns1 = xor_masked(clamp(keep1), 0u - t1, s_cw);
    t0 = lsb(keep0) ^ (t0 & tk_cw);
    t1 = lsb(keep1) ^ (t1 & tk_cw);
    s0 = ns0;
    s1 = ns1;
    s_cw.w |= tl_cw;
    out.put(i, s_cw, make_blk(tr_cw, 0, 0, 0));                // :151-153
  }
  typename GR::V v = GR::add(ga, GR::add(ga, GR::from(ga, beta), GR::neg(ga, GR::from(ga, s0))), GR::from(ga, s1));
  v = GR::cneg(ga, v, t1);                                      // :156-157
  out.put(n, GR::into(ga, v), zero_blk());
}

// ---- VDPF (vdpf.cuh) -------------------------------------------------------------------------------------
FSS_HD blk pack_in(const InVal &x) { return make_blk(x.w[0], x.w[1], x.w[2], x.w[3]); }  // util.cuh:46-63 Pack<In>

// The scheme's two hash plugins (vdpf.cuh:55-58): XorHash H (x, s) -> 64 B and Hash H' 64 B -> 32 B, each Blake3 or SHA-256.
// The kind is a kernel parameter (warp-uniform branch), not a template parameter: the hashes sit outside the level loop.
FSS_HD void vdpf_xor_hash(const PrgKeys &K, blk a, blk b, blk out[4]) {
  if (K.hash_kind[0] == FSSB200_HASH_SHA256) sha_xor_hash(K.hash_iv[0], a, b, out);
  else b3_xor_hash(K.hash_iv[0], a, b, out);
}
FSS_HD void vdpf_hash(const PrgKeys &K, const blk msg[4], blk out[2]) {
  if (K.hash_kind[1] == FSSB200_HASH_SHA256) sha_hash(K.hash_iv[1], msg, out);
  else b3_hash(K.hash_iv[1], msg, out);
}

// Output share and corrected per-point hash of a packed leaf (s | t) at input x: vdpf.cuh:224-242, :318-331.
template <int G>
FSS_HD blk vdpf_leaf(const PrgKeys &K, const GroupArgs &ga, uint32_t party, blk st, const InVal &x, blk ocw,
    const blk *cs, blk pi[4]) {
  typedef Grp<G> GR;
  const uint32_t tm = 0u - lsb(st);
  const blk s = clamp(st);
  typename GR::V y = GR::from(ga, s);
  y = GR::add_masked(ga, y, tm, GR::from(ga, ocw));
  y = GR::cneg(ga, y, party);
  vdpf_xor_hash(K, pack_in(x), s, pi);
#pragma unroll
  for (int j = 0; j < 4; ++j) pi[j] = xor_masked(pi[j], tm, ld_blk(cs + j));
  return GR::into(ga, y);
}

// Vdpf::Eval, vdpf.cuh:191-243: the DPF walk over n correction words (no entry n), then the leaf above.
template <int G, int PRG, class Cw>
FSS_HD blk vdpf_eval_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    uint32_t party, blk s0, const InVal &x, const Cw &cw, blk ocw, const blk *cs, blk pi[4]) {
  blk st = clamp(s0);
  st.w |= party;
  cw.begin_level(0);
  blk cs_cw = cw.s(0);
  uint32_t cf = cw.flag(0);
  cw.done_level(0);
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    blk cs_next = cs_cw;
    uint32_t cf_next = 0;
    if (i + 1 < n) {
      cw.begin_level(i + 1);
      cs_next = cw.s(i + 1);
      cf_next = cw.flag(i + 1);
      cw.done_level(i + 1);
    }
    const uint32_t xb = in_bit(x, n - 1 - i);
    const uint32_t tm = 0u - lsb(st);
    blk c[1];
    Prg<PRG>::template gen_child<1>(K, pc, clamp(st), xb, c);
    blk cwp = cs_cw;
    cwp.w = (cs_cw.w & ~1u) | (xb ? cf : (cs_cw.w & 1u));
    st = xor_masked(c[0], tm, cwp);
    cs_cw = cs_next;
    cf = cf_next;
  }
  return vdpf_leaf<G>(K, ga, party, st, x, ocw, cs, pi);
}

// Vdpf::Gen, vdpf.cuh:97-177.  Returns Gen's status (1: t0 == t1, ocw not written).
template <int G, int PRG, class Out>
FSS_HD int vdpf_gen_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n, blk s0,
    blk s1, const InVal &a, blk beta, const Out &out, blk *cs, blk *ocw, bool write = true) {
  typedef Grp<G> GR;
  s0 = clamp(s0);
  s1 = clamp(s1);
  uint32_t t0 = 0, t1 = 1;
  beta = clamp(beta);
#pragma unroll 1
  for (int i = 0; i < n; ++i) {   // the walk of Dpf::Gen (dpf_gen_body)
    blk g0[2], g1[2];
    Prg<PRG>::template gen<2>(K, pc, s0, g0);
    Prg<PRG>::template gen<2>(K, pc, s1, g1);
    const uint32_t ab = in_bit(a, n - 1 - i);
    const uint32_t am = 0u - ab;
    const blk lose0 = xor_masked(g0[1], am, g0[0] ^ g0[1]);
    const blk lose1 = xor_masked(g1[1], am, g1[0] ^ g1[1]);
    const blk keep0 = xor_masked(g0[0], am, g0[0] ^ g0[1]);
    const blk keep1 = xor_masked(g1[0], am, g1[0] ^ g1[1]);
    blk s_cw = clamp(lose0 ^ lose1);
    const uint32_t tl_cw = (lsb(g0[0]) ^ lsb(g1[0]) ^ ab ^ 1u) & 1u;
    const uint32_t tr_cw = (lsb(g0[1]) ^ lsb(g1[1]) ^ ab) & 1u;
    const uint32_t tk_cw = ab ? tr_cw : tl_cw;
    const blk ns0 = xor_masked(clamp(keep0), 0u - t0, s_cw);
    const blk ns1 = xor_masked(clamp(keep1), 0u - t1, s_cw);
    t0 = lsb(keep0) ^ (t0 & tk_cw);
    t1 = lsb(keep1) ^ (t1 & tk_cw);
    s0 = ns0;
    s1 = ns1;
    s_cw.w |= tl_cw;
    out.put(i, s_cw, make_blk(tr_cw, 0, 0, 0));                // vdpf.cuh:147-149
  }
  blk p0[4], p1[4];                                             // :153-157
  vdpf_xor_hash(K, pack_in(a), s0, p0);
  vdpf_xor_hash(K, pack_in(a), s1, p1);
  if (write) {  // (idle lanes of a ragged tile shadow the last key and write nothing)
#pragma unroll
    for (int j = 0; j < 4; ++j) st_blk(cs + j, p0[j] ^ p1[j]);
  }
  if (t0 == t1) return 1;                                       // :160
  typename GR::V v = GR::add(ga, GR::add(ga, GR::from(ga, beta), GR::neg(ga, GR::from(ga, s0))), GR::from(ga, s1));
  v = GR::cneg(ga, v, t1);
  if (write) st_blk(ocw, GR::into(ga, v));
  return 0;
}

// One step of Vdpf::Prove (vdpf.cuh:257-263) / the EvalAll accumulation (:336-340).
FSS_HD void vdpf_accumulate(const PrgKeys &K, blk pi[4], const blk pt[4]) {
  blk in[4], h[2];
#pragma unroll
  for (int j = 0; j < 4; ++j) in[j] = pi[j] ^ pt[j];
  vdpf_hash(K, in, h);
  pi[0] = pi[0] ^ h[0];
  pi[1] = pi[1] ^ h[1];
}

// ---- DCF ------------------------------------------------------------------------------------------------
// Dcf::Eval, dcf.cuh:205-276.  Two PRG blocks per level (s and v of the chosen side).  The running
// value is accumulated without the party sign; -(a+b) = (-a)+(-b) in every supported group, so the
// sign of dcf.cuh:244-252 is applied once at the end.
template <int G, int PRG, class Cw>
FSS_HD blk dcf_eval_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    uint32_t party, blk s0, const InVal &x, const Cw &cw) {
  typedef Grp<G> GR;
  blk st = clamp(s0);
  st.w |= party;
  typename GR::V acc = GR::zero(ga);
  cw.begin_level(0);
  blk cs = cw.s(0), cv = cw.v(0);
  cw.done_level(0);
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    cw.begin_level(i + 1);
    const blk cs_next = (i + 1 < n) ? cw.s(i + 1) : zero_blk();
    const blk cv_next = (i + 1 < n) ? cw.v(i + 1) : cw.out_v(n);  // entry n = {0, v_cw_{n+1}}
    cw.done_level(i + 1);
    const uint32_t xb = in_bit(x, n - 1 - i);
    const uint32_t t = lsb(st);
    const uint32_t tm = 0u - t;
    blk c[2];
    Prg<PRG>::template gen_child<2>(K, pc, clamp(st), xb, c);
    // value: v += v_side (+ v_cw if t)   (t of the CURRENT node, dcf.cuh:244-252)
    acc = GR::add(ga, acc, GR::from(ga, clamp(c[1])));
    acc = GR::add_masked(ga, acc, tm, GR::from(ga, clamp(cv)));
    // seed: tl_cw = lsb(cw.s), tr_cw = lsb(cw.v)  (dcf.cuh:214-220)
    blk cwp = cs;
    cwp.w = (cs.w & ~1u) | ((xb ? cv.w : cs.w) & 1u);
    st = xor_masked(c[0], tm, cwp);
    cs = cs_next;
    cv = cv_next;
  }
  acc = GR::add(ga, acc, GR::from(ga, clamp(st)));              // dcf.cuh:263-275
  acc = GR::add_masked(ga, acc, 0u - lsb(st), GR::from(ga, cv));
  return GR::into(ga, GR::cneg(ga, acc, party));
}

// Dcf::Gen, dcf.cuh:108-194.  Per level the reference needs, of the eight PRG blocks {s_l, v_l, s_r, v_r} x 2 parties,
// only: the two seeds of the kept side, the XOR of the seeds of the lost side, the control-bit sums, and per side
// D = v1 - v0 (v_cw = -v + D_lose (+ beta), v += -D_keep + sign * v_cw: dcf.cuh:147-160 regrouped; exact in every
// supported group since all values are canonical).  With a per-block PRG (AES) the blocks are consumed pair by
// pair into that reduced state -- the first version held all eight blocks and spilled ~90 registers per level.
template <int G, int PRG, class Out>
FSS_HD void dcf_gen_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    int pred, blk s0, blk s1, const InVal &a, blk beta, const Out &out) {
  typedef Grp<G> GR;
  typedef typename GR::V V;
  typedef Prg<PRG> PG;
  s0 = clamp(s0);
  s1 = clamp(s1);
  uint32_t t0 = 0, t1 = 1;
  V v = GR::zero(ga);
  const V vbeta = GR::from(ga, clamp(beta));
#pragma unroll 1
  for (int i = 0; i < n; ++i) {
    const uint32_t ab = in_bit(a, n - 1 - i);
    const uint32_t am = 0u - ab;                                // ab = 1: keep right, lose left
    blk keep0, keep1, lose_x;                                   // kept seeds (packed with t), s0_lose ^ s1_lose
    uint32_t tl_sum, tr_sum;                                    // lsb(s0l) ^ lsb(s1l), lsb(s0r) ^ lsb(s1r)
    V d_l, d_r;                                                 // v1l - v0l, v1r - v0r
    if (PG::kPerBlock) {
#if FSS_LOOPED_DCF
      // one output block (of both parties) per iteration: 2 AES copies in the loop body instead of 8 (58 KB of SASS
      // against a 32 KB instruction cache)
      keep0 = keep1 = lose_x = zero_blk();
      tl_sum = tr_sum = 0;
      d_l = d_r = GR::zero(ga);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {                               // j = 0..3: s_l, v_l, s_r, v_r
        const blk x0 = PG::block_rt(K, pc, j, s0), x1 = PG::block_rt(K, pc, j, s1);
        if (j & 1) {
          const V dd = GR::add(ga, GR::from(ga, clamp(x1)), GR::neg(ga, GR::from(ga, clamp(x0))));
          if (j == 1) d_l = dd;
          else d_r = dd;
        } else {
          const uint32_t km = j ? am : ~am;                       // this side is kept
          const uint32_t ts = lsb(x0) ^ lsb(x1);
          if (j == 0) tl_sum = ts;
          else tr_sum = ts;
          keep0 = xor_masked(keep0, km, x0);
          keep1 = xor_masked(keep1, km, x1);
          lose_x = xor_masked(lose_x, ~km, x0 ^ x1);
        }
      }
#else
      blk x0 = PG::template block<0>(K, pc, s0), x1 = PG::template block<0>(K, pc, s1);   // s_l
      tl_sum = lsb(x0) ^ lsb(x1);
      keep0 = xor_masked(zero_blk(), ~am, x0);
      keep1 = xor_masked(zero_blk(), ~am, x1);
      lose_x = xor_masked(zero_blk(), am, x0 ^ x1);
      x0 = PG::template block<1>(K, pc, s0); x1 = PG::template block<1>(K, pc, s1);       // v_l
      d_l = GR::add(ga, GR::from(ga, clamp(x1)), GR::neg(ga, GR::from(ga, clamp(x0))));
      x0 = PG::template block<2>(K, pc, s0); x1 = PG::template block<2>(K, pc, s1);       // s_r
      tr_sum = lsb(x0) ^ lsb(x1);
      keep0 = xor_masked(keep0, am, x0);
      keep1 = xor_masked(keep1, am, x1);
      lose_x = xor_masked(lose_x, ~am, x0 ^ x1);
      x0 = PG::template block<3>(K, pc, s0); x1 = PG::template block<3>(K, pc, s1);       // v_r
      d_r = GR::add(ga, GR::from(ga, clamp(x1)), GR::neg(ga, GR::from(ga, clamp(x0))));
#endif
    } else {
      blk g0[4], g1[4];  // {s_l, v_l, s_r, v_r}  (dcf.cuh:122)
      PG::template gen<4>(K, pc, s0, g0);
      PG::template gen<4>(K, pc, s1, g1);
      tl_sum = lsb(g0[0]) ^ lsb(g1[0]);
      tr_sum = lsb(g0[2]) ^ lsb(g1[2]);
      keep0 = ab ? g0[2] : g0[0];
      keep1 = ab ? g1[2] : g1[0];
      lose_x = ab ? (g0[0] ^ g1[0]) : (g0[2] ^ g1[2]);
      d_l = GR::add(ga, GR::from(ga, clamp(g1[1])), GR::neg(ga, GR::from(ga, clamp(g0[1]))));
      d_r = GR::add(ga, GR::from(ga, clamp(g1[3])), GR::neg(ga, GR::from(ga, clamp(g0[3]))));
    }
    blk s_cw = clamp(lose_x);                                   // :143-145
    const V d_lose = ab ? d_l : d_r, d_keep = ab ? d_r : d_l;
    V v_cw = GR::add(ga, GR::neg(ga, v), d_lose);               // :147-153
    if ((ab != 0) == (pred == FSSB200_PRED_LT)) v_cw = GR::add(ga, v_cw, vbeta);
    v_cw = GR::cneg(ga, v_cw, t1);                              // :155
    v = GR::add(ga, GR::add(ga, v, GR::neg(ga, d_keep)), GR::cneg(ga, v_cw, t1));  // :157-160
    const uint32_t tl_cw = (tl_sum ^ ab ^ 1u) & 1u;
    const uint32_t tr_cw = (tr_sum ^ ab) & 1u;
    const uint32_t tk_cw = ab ? tr_cw : tl_cw;
    const blk ns0 = xor_masked(clamp(keep0), 0u - t0, s_cw);
    const blk ns1 = xor_masked(clamp(keep1), 0u - t1, s_cw);
    t0 = lsb(keep0) ^ (t0 & tk_cw);
    t1 = lsb(keep1) ^ (t1 & tk_cw);
    s0 = ns0;
    s1 = ns1;
    s_cw.w |= tl_cw;
    blk v_buf = GR::into(ga, v_cw);
    v_buf.w = (v_buf.w & ~1u) | tr_cw;                          // :187-189
    out.put(i, s_cw, v_buf);
  }
  V vn = GR::add(ga, GR::add(ga, GR::from(ga, s1), GR::neg(ga, GR::from(ga, s0))), GR::neg(ga, v));
  vn = GR::cneg(ga, vn, t1);                                    // :191-193
  out.put(n, zero_blk(), GR::into(ga, vn));
}

// ---- Half-Tree DPF ---------------------------------------------------------------------------------------
FSS_HD blk hash_key_blk(const PrgKeys &K) {
  return make_blk(K.hash_key[0], K.hash_key[1], K.hash_key[2], K.hash_key[3]);
}
// H(node) = prg.Gen(hash_key ^ node)[0]   (half_tree_dpf.cuh:195)
template <int PRG>
FSS_HD blk ht_hash(const PrgKeys &K, const typename Prg<PRG>::ctx_t &pc, blk node) {
  return Prg<PRG>::gen1(K, pc, hash_key_blk(K) ^ node);
}
// Last level + Convert, half_tree_dpf.cuh:208-230 / :325-354.  lcw = LCW_sigma.
template <int G, int PRG>
FSS_HD blk ht_last(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, uint32_t party,
    blk node, uint32_t sigma, blk cw_last, uint32_t lcw, blk ocw) {
  typedef Grp<G> GR;
  const uint32_t tm = 0u - lsb(node);
  blk in = node;
  in.w = (node.w & ~1u) | sigma;
  blk h = ht_hash<PRG>(K, pc, in);
  blk corr = cw_last;                                           // SetLsb(HCW, LCW_sigma)
  corr.w = (cw_last.w & ~1u) | lcw;
  h = xor_masked(h, tm, corr);                                  // high ^= hcw, low ^= lcw
  typename GR::V y = GR::from(ga, clamp(h));
  y = GR::add_masked(ga, y, 0u - lsb(h), GR::from(ga, ocw));
  return GR::into(ga, GR::cneg(ga, y, party));
}
// HalfTreeDpf::Eval, half_tree_dpf.cuh:187-231.  n hashes per evaluation.
template <int G, int PRG, class Cw>
FSS_HD blk ht_eval_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n,
    uint32_t party, blk s0, const InVal &x, const Cw &cw, blk ocw) {
  blk node = clamp(s0);
  node.w |= party;
  cw.begin_level(0);
  blk cs = cw.s(0);
  cw.done_level(0);
#pragma unroll 1
  for (int i = 0; i < n - 1; ++i) {
    cw.begin_level(i + 1);
    const blk cs_next = cw.s(i + 1);
    cw.done_level(i + 1);
    const uint32_t xm = 0u - in_bit(x, n - 1 - i);
    const uint32_t tm = 0u - lsb(node);
    const blk h = ht_hash<PRG>(K, pc, node);
    node = xor_masked(xor_masked(h, xm, node), tm, cs);          // :202-204 (cw lsb included)
    cs = cs_next;
  }
  const uint32_t xn = in_bit(x, 0);
  const uint32_t lcw = xn ? cw.flag(n - 1) : (cs.w & 1u);        // :213-216
  return ht_last<G, PRG>(K, ga, pc, party, node, xn, cs, lcw, ocw);
}
// HalfTreeDpf::Gen, half_tree_dpf.cuh:68-175.
template <int G, int PRG, class Out>
FSS_HD void ht_gen_body(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, int n, blk s0,
    blk s1, const InVal &a, blk beta, const Out &out, blk *ocw) {
  typedef Grp<G> GR;
  beta = clamp(beta);
  blk node0 = clamp(s0), node1 = clamp(s1);
  node1.w |= 1u;
  blk delta = node0 ^ node1;
#pragma unroll 1
  for (int i = 0; i < n - 1; ++i) {
    const blk h0 = ht_hash<PRG>(K, pc, node0), h1 = ht_hash<PRG>(K, pc, node1);
    const uint32_t am = 0u - in_bit(a, n - 1 - i);
    const blk cw = xor_masked(h0 ^ h1, ~am, delta);              // :83-84
    out.put(i, cw, zero_blk());
    const uint32_t t0m = 0u - lsb(node0), t1m = 0u - lsb(node1);
    node0 = xor_masked(xor_masked(h0, am, node0), t0m, cw);
    node1 = xor_masked(xor_masked(h1, am, node1), t1m, cw);
    delta = node0 ^ node1;
  }
  const uint32_t an = in_bit(a, 0);
  const uint32_t t0 = lsb(node0), t1 = lsb(node1);
  blk in;
  in = clamp(node0); const blk h0_0 = ht_hash<PRG>(K, pc, in);
  in.w |= 1u;        const blk h0_1 = ht_hash<PRG>(K, pc, in);
  in = clamp(node1); const blk h1_0 = ht_hash<PRG>(K, pc, in);
  in.w |= 1u;        const blk h1_1 = ht_hash<PRG>(K, pc, in);
  const blk hcw = an ? clamp(h0_0 ^ h1_0) : clamp(h0_1 ^ h1_1);   // :123-125
  const uint32_t lcw0 = (lsb(h0_0) ^ lsb(h1_0) ^ an ^ 1u) & 1u;   // :132
  const uint32_t lcw1 = (lsb(h0_1) ^ lsb(h1_1) ^ an) & 1u;        // :133
  blk cwn = hcw;
  cwn.w |= lcw0;
  out.put(n - 1, cwn, make_blk(lcw1, 0, 0, 0));                   // :139-141
  blk leaf0 = an ? h0_1 : h0_0, leaf1 = an ? h1_1 : h1_0;          // packed high||low
  blk leaf_cw = hcw;
  leaf_cw.w |= an ? lcw1 : lcw0;
  leaf0 = xor_masked(leaf0, 0u - t0, leaf_cw);
  leaf1 = xor_masked(leaf1, 0u - t1, leaf_cw);
  typename GR::V v = GR::add(ga, GR::add(ga, GR::from(ga, beta), GR::neg(ga, GR::from(ga, clamp(leaf0)))),
      GR::from(ga, clamp(leaf1)));
  v = GR::cneg(ga, v, lsb(leaf1));                                 // :171-173
  *ocw = GR::into(ga, v);
}

// ---- full-domain node expansion (dpf.cuh:265-288, eval_all_gpu.cuh:154-178) ------------------------------
// cwl / cwr: the level's s_cw with tl_cw resp. tr_cw in the clamp bit.
template <int PRG>
FSS_HD void dpf_expand(const PrgKeys &K, const typename Prg<PRG>::ctx_t &pc, blk st, blk cwl, blk cwr, blk &left,
    blk &right) {
  const uint32_t tm = 0u - lsb(st);
#if FSS_LOOPED_EXPAND
  if (Prg<PRG>::kPerBlock) {
    const blk in = clamp(st);
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const blk g = Prg<PRG>::block_rt(K, pc, i, in);
      if (i == 0) left = xor_masked(g, tm, cwl);
      else right = xor_masked(g, tm, cwr);
    }
    return;
  }
#endif
  blk g[2];
  Prg<PRG>::template gen<2>(K, pc, clamp(st), g);
  left = xor_masked(g[0], tm, cwl);
  right = xor_masked(g[1], tm, cwr);
}
// half_tree_dpf.cuh:292-302, eval_all_gpu.cuh:46-52
template <int PRG>
FSS_HD void ht_expand(const PrgKeys &K, const typename Prg<PRG>::ctx_t &pc, blk node, blk cw, blk &left,
    blk &right) {
  const uint32_t tm = 0u - lsb(node);
  const blk h = ht_hash<PRG>(K, pc, node);
  left = xor_masked(h, tm, cw);
  right = left ^ node;
}

// Dcf::EvalTree node (dcf.cuh:338-371) with the value share carried WITHOUT the party sign (see
// dcf_eval_body): children values u + v_side (+ v_cw if t).  cwl / cwr as in dpf_expand; vcw = From(clamp(cw.v)).
template <int G, int PRG>
FSS_HD void dcf_expand(const PrgKeys &K, const GroupArgs &ga, const typename Prg<PRG>::ctx_t &pc, blk st,
    typename Grp<G>::V u, blk cwl, blk cwr, typename Grp<G>::V vcw, blk &left, blk &right, typename Grp<G>::V &ul,
    typename Grp<G>::V &ur) {
  typedef Grp<G> GR;
  const uint32_t tm = 0u - lsb(st);
#if FSS_LOOPED_DCF
  // (only for the groups with multi-word / modular arithmetic: for Bytes / u32 / u64 the four inlined copies still fit and
  // their extra instruction-level parallelism wins at 8 warps per SM: Bytes 0.861 inlined vs 0.835 looped, u127 0.854 vs 0.874)
  if (Prg<PRG>::kPerBlock && G >= kGrpU127) {
    // one side per iteration (2 AES copies in the loop body instead of 4: 29 KB + the group arithmetic does not fit
    // the 32 KB instruction cache)
    const blk in = clamp(st);
    const typename GR::V base = GR::add_masked(ga, u, tm, vcw);
#pragma unroll 1
    for (int side = 0; side < 2; ++side) {
      const blk gs = Prg<PRG>::block_rt(K, pc, 2 * side, in), gv = Prg<PRG>::block_rt(K, pc, 2 * side + 1, in);
      const typename GR::V uv = GR::add(ga, base, GR::from(ga, clamp(gv)));
      if (side == 0) {
        left = xor_masked(gs, tm, cwl);
        ul = uv;
      } else {
        right = xor_masked(gs, tm, cwr);
        ur = uv;
      }
    }
    return;
  }
#endif
  blk g[4];  // {s_l, v_l, s_r, v_r}
  Prg<PRG>::template gen<4>(K, pc, clamp(st), g);
  left = xor_masked(g[0], tm, cwl);
  right = xor_masked(g[2], tm, cwr);
  const typename GR::V base = GR::add_masked(ga, u, tm, vcw);
  ul = GR::add(ga, base, GR::from(ga, clamp(g[1])));
  ur = GR::add(ga, base, GR::from(ga, clamp(g[3])));
}
// Dcf leaf (dcf.cuh:319-329): y = sign * (u + From(s) + (t ? v_cw_{n+1} : 0))
template <int G>
FSS_HD blk dcf_leaf(const GroupArgs &ga, uint32_t party, blk st, typename Grp<G>::V u, blk out_v) {
  typedef Grp<G> GR;
  typename GR::V y = GR::add(ga, u, GR::from(ga, clamp(st)));
  y = GR::add_masked(ga, y, 0u - lsb(st), GR::from(ga, out_v));
  return GR::into(ga, GR::cneg(ga, y, party));
}

// ---- Grotto DCF: O(n) point walk (no counterpart in the reference) ---------------------------
// GrottoDcf::Eval (grotto_dcf.cuh:116-135) is a lookup in a per-key parity tree of 2N-1 bools that Preprocess
// (:94-104) builds from ALL N leaves: 8 GiB per key at n = 32, so a batch of 2^20 such keys cannot exist.  The
// share it returns is the parity of the leaf control bits of [0, x+1).  This walk returns a DIFFERENT share of the
// same secret: [0, e), e = x + 1, is the disjoint union of the left-sibling subtrees along e's path (one for every
// 1 bit of e), and the control bit of a subtree root reconstructs to 1 iff alpha lies in that subtree (the DPF
// invariant t0 ^ t1 = [node on alpha's path], dpf.cuh:119-120) -- exactly what the parity of its leaves
// reconstructs to.  So  share = XOR of the control bits of those left siblings  satisfies
// share_0 ^ share_1 = 1[alpha <= x] like the reference, in O(n) PRG calls and no memory; per party the bit
// differs from the reference's (the parity of pseudorandom leaf bits is not the root's bit): "reconstruction-equal,
// share-parity unpinned".  Full range (e wraps to 0, or e == N): the root's control bit = the party index.
// Blocks: both children at levels 0..n-2, only the left block (its control bit) at the last level.
FSS_HD InVal in_succ(const InVal &x, int in_bytes) {
  InVal e;
  uint32_t c = 1;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    e.w[j] = x.w[j] + c;
    c = (e.w[j] < c) ? 1u : 0u;
  }
  if (in_bytes == 1) e.w[0] &= 0xffu;
  if (in_bytes == 2) e.w[0] &= 0xffffu;
  if (in_bytes <= 4) e.w[1] = 0;
  if (in_bytes <= 8) e.w[2] = e.w[3] = 0;
  return e;
}
template <int PRG, class Cw>
FSS_HD uint32_t grotto_walk_body(const PrgKeys &K, const typename Prg<PRG>::ctx_t &pc, int n, int in_bytes,
    uint32_t party, blk s0, const InVal &x, const Cw &cw) {
  const InVal e = in_succ(x, in_bytes);
  // e == 0 (x + 1 overflowed In) or e == N: the whole domain (grotto_dcf.cuh:121)
  bool full = (e.w[0] | e.w[1] | e.w[2] | e.w[3]) == 0u;
  if (n < 8 * in_bytes) {
    bool is_n = true;
#pragma unroll
    for (int j = 0; j < 4; ++j) is_n = is_n && e.w[j] == ((n >> 5) == j ? (1u << (n & 31)) : 0u);
    full = full || is_n;
  }
  blk st = clamp(s0);
  st.w |= party;
  uint32_t acc = 0;
  cw.begin_level(0);
  blk cs = cw.s(0);
  uint32_t cf = cw.flag(0);
  cw.done_level(0);
#pragma unroll 1
  for (int i = 0; i + 1 < n; ++i) {
    cw.begin_level(i + 1);
    const blk cs_next = cw.s(i + 1);
    const uint32_t cf_next = cw.flag(i + 1);
    cw.done_level(i + 1);
    const uint32_t eb = in_bit(e, n - 1 - i);
    blk cwr = cs;
    cwr.w = (cs.w & ~1u) | cf;
    blk l, r;
    dpf_expand<PRG>(K, pc, st, cs, cwr, l, r);
    acc ^= eb & lsb(l);                              // left sibling lies inside [0, e)
    st = xor_masked(l, 0u - eb, l ^ r);              // eb ? r : l
    cs = cs_next;
    cf = cf_next;
  }
  cw.begin_level(n);  // (entry n, the output CW, is not used: Grotto has beta = 0 -- but a staged accessor must be
  cw.done_level(n);   //  walked to the end of the row so that its prefetch pipeline stays in step with the next tile)
  {  // last level: only the left child's control bit can still count (the leaf e itself is outside [0, e))
    const uint32_t tm = 0u - lsb(st);
    const blk l = xor_masked(Prg<PRG>::gen_left(K, pc, clamp(st)), tm, cs);
    acc ^= in_bit(e, 0) & lsb(l);
  }
  return full ? party : acc;
}

}  // namespace fssb200
