/* SPDX-License-Identifier: Apache-2.0
 *
 * TEST INFRASTRUCTURE ONLY -- see fss_oracle.h.  Plain-C CPU restatement of the
 * reference algorithms on the DPF/DCF hot path; every function cites the reference
 * lines (relative to /root/reference/) it restates.  Deliberately written in the
 * most literal way (byte-wise FIPS-197 AES, recursive trees, unsigned __int128
 * group arithmetic): it shares no code and no table layout with the CUDA kernels it
 * checks.  PARITY PINNED against the compiled reference (oracle/_ref) and the
 * golden fixtures by tests/test_oracle.py; the VDPF part (Blake3 and SHA-256 hash
 * plugins) by tests/test_vdpf.py against oracle/ref_vdpf.cpp, the reference-generated
 * fixtures tests/golden/golden_vdpf{,_sha256}_v1.* and Python's hashlib.
 */
#include "fss_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint32_t w[4]; } blk; /* CUDA int4 {x,y,z,w}, little-endian words */

/* ---- util.cuh:16-38 ---------------------------------------------------------- */
static blk bxor(blk a, blk b) {
  blk r;
  for (int i = 0; i < 4; ++i) r.w[i] = a.w[i] ^ b.w[i];
  return r;
}
static blk set_lsb(blk v, int bit) { /* util.cuh:30-34 */
  if (bit) v.w[3] |= 1u; else v.w[3] &= ~1u;
  return v;
}
static int get_lsb(blk v) { return (int)(v.w[3] & 1u); } /* util.cuh:36-38 */
static blk bzero(void) { blk r = {{0, 0, 0, 0}}; return r; }
static blk bsel(int c, blk v) { return c ? v : bzero(); }

/* ---- AES-128 (FIPS-197), the cipher behind prg/aes128_mmo.cuh:57,84 ------------ */
static uint8_t g_sbox[256];
static int g_sbox_ready = 0;

static uint8_t gf_mul(uint8_t a, uint8_t b) {
  uint8_t r = 0;
  while (b) {
    if (b & 1) r ^= a;
    a = (uint8_t)((a << 1) ^ ((a & 0x80) ? 0x1b : 0));
    b >>= 1;
  }
  return r;
}
static void sbox_init(void) {
  if (g_sbox_ready) return;
  for (int x = 0; x < 256; ++x) {
    uint8_t inv = 0;
    if (x) for (int y = 1; y < 256; ++y) if (gf_mul((uint8_t)x, (uint8_t)y) == 1) { inv = (uint8_t)y; break; }
    uint8_t s = inv, r = inv;
    for (int k = 0; k < 4; ++k) { r = (uint8_t)((r << 1) | (r >> 7)); s ^= r; }
    g_sbox[x] = (uint8_t)(s ^ 0x63);
  }
  g_sbox_ready = 1;
}
static void aes_expand(const uint8_t key[16], uint8_t rk[176]) {
  static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
  memcpy(rk, key, 16);
  for (int i = 4; i < 44; ++i) {
    uint8_t t[4];
    memcpy(t, rk + 4 * (i - 1), 4);
    if (i % 4 == 0) {
      uint8_t u = t[0];
      t[0] = (uint8_t)(g_sbox[t[1]] ^ rcon[i / 4 - 1]);
      t[1] = g_sbox[t[2]];
      t[2] = g_sbox[t[3]];
      t[3] = g_sbox[u];
    }
    for (int j = 0; j < 4; ++j) rk[4 * i + j] = (uint8_t)(rk[4 * (i - 4) + j] ^ t[j]);
  }
}
static void aes_encrypt(const uint8_t rk[176], const uint8_t in[16], uint8_t out[16]) {
  uint8_t s[16], t[16];
  for (int i = 0; i < 16; ++i) s[i] = (uint8_t)(in[i] ^ rk[i]);
  for (int r = 1; r <= 10; ++r) {
    for (int i = 0; i < 16; ++i) s[i] = g_sbox[s[i]];           /* SubBytes  */
    for (int c = 0; c < 4; ++c)                                  /* ShiftRows */
      for (int row = 0; row < 4; ++row) t[4 * c + row] = s[4 * ((c + row) & 3) + row];
    if (r < 10) {                                                /* MixColumns */
      for (int c = 0; c < 4; ++c) {
        const uint8_t *a = t + 4 * c;
        s[4 * c + 0] = (uint8_t)(gf_mul(a[0], 2) ^ gf_mul(a[1], 3) ^ a[2] ^ a[3]);
        s[4 * c + 1] = (uint8_t)(a[0] ^ gf_mul(a[1], 2) ^ gf_mul(a[2], 3) ^ a[3]);
        s[4 * c + 2] = (uint8_t)(a[0] ^ a[1] ^ gf_mul(a[2], 2) ^ gf_mul(a[3], 3));
        s[4 * c + 3] = (uint8_t)(gf_mul(a[0], 3) ^ a[1] ^ a[2] ^ gf_mul(a[3], 2));
      }
    } else {
      memcpy(s, t, 16);
    }
    for (int i = 0; i < 16; ++i) s[i] ^= rk[16 * r + i];         /* AddRoundKey */
  }
  memcpy(out, s, 16);
}

/* ---- ChaCha block, prg/chacha.cuh:36-61 ------------------------------------------ */
static uint32_t rotl32(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
#define QR(a, b, c, d)                                                       \
  do {                                                                       \
    a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);    \
    a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);     \
  } while (0)

/* ---- evaluation context ------------------------------------------------------------ */
typedef struct {
  fssb200_params p;
  uint8_t rk[4][176];
  uint32_t nonce[2];
  blk hash_key;
  u128 mod;   /* 0 = power of two of the value width */
  int vbytes; /* value width of the Uint group; 0 for Bytes */
  int mul;
} octx;

static int ctx_init(octx *c, const fssb200_params *p) {
  if (!p) return FSSB200_EINVAL;
  sbox_init();
  memset(c, 0, sizeof(*c));
  c->p = *p;
  if (p->scheme < 0 || p->scheme > 4) return FSSB200_EINVAL;
  if (p->in_bytes != 1 && p->in_bytes != 2 && p->in_bytes != 4 && p->in_bytes != 8 && p->in_bytes != 16)
    return FSSB200_EINVAL;
  if (p->in_bits < 1 || p->in_bits > 8 * p->in_bytes) return FSSB200_EDOMAIN;
  static const int vb[6] = {0, 1, 2, 4, 8, 16};
  if (p->group < 0 || p->group > 5) return FSSB200_EGROUP;
  c->vbytes = vb[p->group];
  c->mod = ((u128)p->mod_hi << 64) | p->mod_lo;
  if (p->scheme == FSSB200_SCHEME_GROTTO) { c->vbytes = 0; c->mod = 0; c->p.group = FSSB200_GROUP_BYTES; }
  if (c->p.group == FSSB200_GROUP_BYTES && c->mod) return FSSB200_EGROUP;
  if (c->p.group == FSSB200_GROUP_U128 && (c->mod == 0 || c->mod > ((u128)1 << 127))) return FSSB200_EGROUP;
  if (c->vbytes && c->vbytes < 16 && c->mod && (c->mod >> (8 * c->vbytes))) return FSSB200_EGROUP;
  c->mul = p->scheme == FSSB200_SCHEME_DCF ? 4 : (p->scheme == FSSB200_SCHEME_HALFTREE ? 1 : 2);
  if (p->prg == FSSB200_PRG_AES128_MMO) {
    for (int i = 0; i < 4; ++i) aes_expand(p->prg_key + 16 * i, c->rk[i]);
  } else if (p->prg == FSSB200_PRG_CHACHA) {
    memcpy(c->nonce, p->prg_key, 8);
  } else {
    return FSSB200_EINVAL;
  }
  memcpy(&c->hash_key, p->hash_key, 16);
  return 0;
}

/* prg.Gen(seed): prg/aes128_mmo.cuh:72-93 (out[i] = AES_{key_i}(seed) ^ seed) and
 * prg/chacha.cuh:95-127. */
static void prg_gen(const octx *c, int mul, blk seed, blk out[4]) {
  if (c->p.prg == FSSB200_PRG_AES128_MMO) {
    for (int i = 0; i < mul; ++i) {
      uint8_t in[16], o[16];
      memcpy(in, &seed, 16);
      aes_encrypt(c->rk[i], in, o);
      memcpy(&out[i], o, 16);
      out[i] = bxor(out[i], seed);
    }
    return;
  }
  /* chacha.cuh:71-83 constants; :99-110 state; 20 rounds :47-61 */
  static const uint32_t k16[4] = {0x61707865, 0x3120646e, 0x79622d36, 0x6b206574};
  static const uint32_t k32[4] = {0x61707865, 0x3320646e, 0x79622d32, 0x6b206574};
  const uint32_t *kc = mul <= 2 ? k16 : k32;
  uint32_t x[16];
  for (int i = 0; i < 4; ++i) { x[i] = kc[i]; x[4 + i] = seed.w[i]; x[8 + i] = seed.w[i]; }
  x[12] = 0; x[13] = 0; x[14] = c->nonce[0]; x[15] = c->nonce[1];
  for (int r = 0; r < 20; r += 2) {
    QR(x[0], x[4], x[8], x[12]);  QR(x[1], x[5], x[9], x[13]);
    QR(x[2], x[6], x[10], x[14]); QR(x[3], x[7], x[11], x[15]);
    QR(x[0], x[5], x[10], x[15]); QR(x[1], x[6], x[11], x[12]);
    QR(x[2], x[7], x[8], x[13]);  QR(x[3], x[4], x[9], x[14]);
  }
  blk row[4];
  for (int r = 0; r < 4; ++r) for (int i = 0; i < 4; ++i) row[r].w[i] = x[4 * r + i];
  blk kb; for (int i = 0; i < 4; ++i) kb.w[i] = kc[i];
  blk nb = {{0, 0, c->nonce[0], c->nonce[1]}};
  if (mul == 1) { out[0] = bxor(row[1], seed); return; }                    /* :113-115 */
  if (mul == 2) { out[0] = bxor(row[0], kb); out[1] = bxor(row[1], seed); return; } /* :116-118 */
  out[0] = bxor(row[0], kb); out[1] = bxor(row[1], seed);                   /* :119-124 */
  out[2] = bxor(row[2], seed); out[3] = bxor(row[3], nb);
}

/* ---- groups: group/bytes.cuh:19-43, group/uint.cuh:27-88 -------------------------------
 * A group element is carried as a u128: for Bytes the packed 16 bytes, for Uint<T,mod>
 * the value. */
static u128 vmask(const octx *c) { return c->vbytes >= 16 ? ~(u128)0 : (((u128)1 << (8 * c->vbytes)) - 1); }
static u128 g_from(const octx *c, blk b) {
  u128 v;
  if (c->vbytes == 0) return (u128)b.w[0] | ((u128)b.w[1] << 32) | ((u128)b.w[2] << 64) | ((u128)b.w[3] << 96);
  if (c->vbytes < 4) v = b.w[0] & (uint32_t)vmask(c);                   /* uint.cuh:53 */
  else if (c->vbytes == 4) v = b.w[0];                                  /* :55 */
  else if (c->vbytes == 8) v = (u128)b.w[0] | ((u128)b.w[1] << 32);     /* :56-57 */
  else v = (u128)b.w[0] | ((u128)b.w[1] << 32) | ((u128)b.w[2] << 64) | ((u128)(b.w[3] >> 1) << 96); /* :58-62 */
  if (c->mod) v %= c->mod;                                              /* :65 */
  return v;
}
static blk g_into(const octx *c, u128 v) {
  blk b = bzero();
  if (c->vbytes == 0) { for (int i = 0; i < 4; ++i) b.w[i] = (uint32_t)(v >> (32 * i)); return b; }
  if (c->vbytes <= 4) b.w[0] = (uint32_t)v;                             /* :72 */
  else if (c->vbytes == 8) { b.w[0] = (uint32_t)v; b.w[1] = (uint32_t)(v >> 32); } /* :73-75 */
  else { b.w[0] = (uint32_t)v; b.w[1] = (uint32_t)(v >> 32); b.w[2] = (uint32_t)(v >> 64);
         b.w[3] = (uint32_t)((v >> 96) << 1); }                         /* :76-81 */
  return b;
}
static u128 g_add(const octx *c, u128 a, u128 b) {
  if (c->vbytes == 0) return a ^ b;                                     /* bytes.cuh:22-24 */
  if (!c->mod) return (a + b) & vmask(c);                               /* uint.cuh:34 */
  u128 s = a + b;                                                        /* a,b < mod <= 2^127 */
  return s >= c->mod ? s - c->mod : s;                                   /* :36-37 */
}
static u128 g_neg(const octx *c, u128 a) {
  if (c->vbytes == 0) return a;                                         /* bytes.cuh:26-28 */
  if (!c->mod) return (0 - a) & vmask(c);                               /* uint.cuh:41 */
  return a ? c->mod - a : 0;                                             /* :43-44 */
}

static u128 load_in(const uint8_t *p, int in_bytes) { u128 v = 0; memcpy(&v, p, (size_t)in_bytes); return v; }
static int in_bit(u128 x, int n, int i) { return (int)((x >> (n - 1 - i)) & 1); }

/* Cw accessors: 32-byte slots {blk s; blk second} */
typedef struct { blk s; blk v; } cw32;

/* ---- DPF: dpf.cuh ---------------------------------------------------------------------- */
static void dpf_gen(const octx *c, cw32 *cws, const blk s0s[2], u128 a, blk b_buf) { /* :93-159 */
  const int n = c->p.in_bits;
  blk s0 = set_lsb(s0s[0], 0), s1 = set_lsb(s0s[1], 0);
  int t0 = 0, t1 = 1;
  b_buf = set_lsb(b_buf, 0);
  for (int i = 0; i < n; ++i) {
    blk g0[4], g1[4];
    prg_gen(c, 2, s0, g0);
    prg_gen(c, 2, s1, g1);
    int t0l = get_lsb(g0[0]), t0r = get_lsb(g0[1]), t1l = get_lsb(g1[0]), t1r = get_lsb(g1[1]);
    blk s0l = set_lsb(g0[0], 0), s0r = set_lsb(g0[1], 0), s1l = set_lsb(g1[0], 0), s1r = set_lsb(g1[1], 0);
    int a_bit = in_bit(a, n, i);
    blk s_cw = a_bit ? bxor(s0l, s1l) : bxor(s0r, s1r);
    int tl_cw = t0l ^ t1l ^ a_bit ^ 1, tr_cw = t0r ^ t1r ^ a_bit;
    if (!a_bit) {
      s0 = bxor(s0l, bsel(t0, s_cw)); s1 = bxor(s1l, bsel(t1, s_cw));
      t0 = t0l ^ (t0 & tl_cw); t1 = t1l ^ (t1 & tl_cw);
    } else {
      s0 = bxor(s0r, bsel(t0, s_cw)); s1 = bxor(s1r, bsel(t1, s_cw));
      t0 = t0r ^ (t0 & tr_cw); t1 = t1r ^ (t1 & tr_cw);
    }
    cws[i].s = set_lsb(s_cw, tl_cw);
    cws[i].v = bzero(); cws[i].v.w[0] = (uint32_t)tr_cw;     /* {tr,0,0,0} :151-153 */
  }
  u128 v = g_add(c, g_add(c, g_from(c, b_buf), g_neg(c, g_from(c, s0))), g_from(c, s1));
  if (t1) v = g_neg(c, v);
  cws[n].s = g_into(c, v); cws[n].v = bzero();               /* :158 (padding unspecified) */
}

static int cw_tr(const cw32 *cw) { return ((const uint8_t *)cw)[16] != 0; } /* bool at byte 16 */

/* One DPF node expansion, both children, packed (seed | t in lsb): dpf.cuh:265-288 */
static void dpf_expand(const octx *c, blk st, const cw32 *cw, blk *left, blk *right) {
  int t = get_lsb(st);
  blk s = set_lsb(st, 0), g[4];
  blk s_cw = set_lsb(cw->s, 0);
  int tl_cw = get_lsb(cw->s), tr_cw = cw_tr(cw);
  prg_gen(c, 2, s, g);
  int tl = get_lsb(g[0]), tr = get_lsb(g[1]);
  blk sl = set_lsb(g[0], 0), sr = set_lsb(g[1], 0);
  if (t) { sl = bxor(sl, s_cw); sr = bxor(sr, s_cw); tl ^= tl_cw; tr ^= tr_cw; }
  *left = set_lsb(sl, tl); *right = set_lsb(sr, tr);
}
static blk dpf_leaf(const octx *c, int b, blk st, const cw32 *cws) { /* :207-213, :255-263 */
  int t = get_lsb(st);
  u128 y = g_from(c, set_lsb(st, 0));
  if (t) y = g_add(c, y, g_from(c, cws[c->p.in_bits].s));
  if (b) y = g_neg(c, y);
  return g_into(c, y);
}
static blk dpf_eval(const octx *c, int b, blk s0, const cw32 *cws, u128 x) { /* :170-214 */
  const int n = c->p.in_bits;
  blk st = set_lsb(s0, b);
  for (int i = 0; i < n; ++i) {
    blk l, r;
    dpf_expand(c, st, &cws[i], &l, &r);
    st = in_bit(x, n, i) ? r : l;
  }
  return dpf_leaf(c, b, st, cws);
}

/* ---- DCF: dcf.cuh ------------------------------------------------------------------------ */
static void dcf_gen(const octx *c, cw32 *cws, const blk s0s[2], u128 a, blk b_buf) { /* :108-194 */
  const int n = c->p.in_bits;
  blk s0 = set_lsb(s0s[0], 0), s1 = set_lsb(s0s[1], 0);
  int t0 = 0, t1 = 1;
  u128 v = 0;
  b_buf = set_lsb(b_buf, 0);
  const u128 beta = g_from(c, b_buf);
  for (int i = 0; i < n; ++i) {
    blk g0[4], g1[4];
    prg_gen(c, 4, s0, g0);
    prg_gen(c, 4, s1, g1);
    int t0l = get_lsb(g0[0]), t0r = get_lsb(g0[2]), t1l = get_lsb(g1[0]), t1r = get_lsb(g1[2]);
    blk s0l = set_lsb(g0[0], 0), s0r = set_lsb(g0[2], 0), s1l = set_lsb(g1[0], 0), s1r = set_lsb(g1[2], 0);
    u128 v0l = g_from(c, set_lsb(g0[1], 0)), v0r = g_from(c, set_lsb(g0[3], 0));
    u128 v1l = g_from(c, set_lsb(g1[1], 0)), v1r = g_from(c, set_lsb(g1[3], 0));
    int a_bit = in_bit(a, n, i);
    blk s_cw = a_bit ? bxor(s0l, s1l) : bxor(s0r, s1r);
    u128 v_cw = g_neg(c, v);                                              /* :139 */
    if (!a_bit) {
      v_cw = g_add(c, g_add(c, v_cw, v1r), g_neg(c, v0r));
      if (c->p.pred == FSSB200_PRED_GT) v_cw = g_add(c, v_cw, beta);
    } else {
      v_cw = g_add(c, g_add(c, v_cw, v1l), g_neg(c, v0l));
      if (c->p.pred == FSSB200_PRED_LT) v_cw = g_add(c, v_cw, beta);
    }
    if (t1) v_cw = g_neg(c, v_cw);
    if (!a_bit) v = g_add(c, g_add(c, v, g_neg(c, v1l)), v0l);
    else v = g_add(c, g_add(c, v, g_neg(c, v1r)), v0r);
    if (t1) v = g_add(c, v, g_neg(c, v_cw)); else v = g_add(c, v, v_cw);
    int tl_cw = t0l ^ t1l ^ a_bit ^ 1, tr_cw = t0r ^ t1r ^ a_bit;
    if (!a_bit) {
      s0 = bxor(s0l, bsel(t0, s_cw)); s1 = bxor(s1l, bsel(t1, s_cw));
      t0 = t0l ^ (t0 & tl_cw); t1 = t1l ^ (t1 & tl_cw);
    } else {
      s0 = bxor(s0r, bsel(t0, s_cw)); s1 = bxor(s1r, bsel(t1, s_cw));
      t0 = t0r ^ (t0 & tr_cw); t1 = t1r ^ (t1 & tr_cw);
    }
    cws[i].s = set_lsb(s_cw, tl_cw);
    cws[i].v = set_lsb(g_into(c, v_cw), tr_cw);                           /* :187-189 */
  }
  u128 vn = g_add(c, g_add(c, g_from(c, s1), g_neg(c, g_from(c, s0))), g_neg(c, v));
  if (t1) vn = g_neg(c, vn);
  cws[n].s = bzero(); cws[n].v = g_into(c, vn);                           /* :193 */
}

/* One DCF node: children (packed) and their running values; dcf.cuh:338-371 */
static void dcf_expand(const octx *c, int b, blk st, u128 v, const cw32 *cw, blk *left, blk *right,
    u128 *vl_out, u128 *vr_out) {
  int t = get_lsb(st);
  blk s = set_lsb(st, 0), g[4];
  blk s_cw = set_lsb(cw->s, 0);
  int tl_cw = get_lsb(cw->s), tr_cw = get_lsb(cw->v);
  u128 v_cw = g_from(c, set_lsb(cw->v, 0));
  prg_gen(c, 4, s, g);
  int tl = get_lsb(g[0]), tr = get_lsb(g[2]);
  blk sl = set_lsb(g[0], 0), sr = set_lsb(g[2], 0);
  u128 vl = g_from(c, set_lsb(g[1], 0)), vr = g_from(c, set_lsb(g[3], 0));
  if (t) {
    sl = bxor(sl, s_cw); sr = bxor(sr, s_cw); tl ^= tl_cw; tr ^= tr_cw;
    vl = g_add(c, vl, v_cw); vr = g_add(c, vr, v_cw);
  }
  if (b) { vl = g_neg(c, vl); vr = g_neg(c, vr); }
  *vl_out = g_add(c, vl, v); *vr_out = g_add(c, vr, v);
  *left = set_lsb(sl, tl); *right = set_lsb(sr, tr);
}
static blk dcf_leaf(const octx *c, int b, blk st, u128 v, const cw32 *cws) { /* :263-275, :319-329 */
  int t = get_lsb(st);
  u128 term = g_from(c, set_lsb(st, 0));
  if (t) term = g_add(c, term, g_from(c, cws[c->p.in_bits].v));
  if (b) term = g_neg(c, term);
  return g_into(c, g_add(c, v, term));
}
static blk dcf_eval(const octx *c, int b, blk s0, const cw32 *cws, u128 x) { /* :205-276 */
  const int n = c->p.in_bits;
  blk st = set_lsb(s0, b);
  u128 v = 0;
  for (int i = 0; i < n; ++i) {
    blk l, r; u128 vl, vr;
    dcf_expand(c, b, st, v, &cws[i], &l, &r, &vl, &vr);
    if (in_bit(x, n, i)) { st = r; v = vr; } else { st = l; v = vl; }
  }
  return dcf_leaf(c, b, st, v, cws);
}

/* ---- Half-Tree DPF: half_tree_dpf.cuh ------------------------------------------------------ */
static blk ht_hash(const octx *c, blk node) { /* prg.Gen(hash_key ^ node)[0] */
  blk g[4];
  prg_gen(c, 1, bxor(c->hash_key, node), g);
  return g[0];
}
static void ht_gen(const octx *c, cw32 *cws, blk *ocw, const blk s0s[2], u128 a, blk b_buf) { /* :68-175 */
  const int n = c->p.in_bits;
  b_buf = set_lsb(b_buf, 0);
  blk node0 = set_lsb(s0s[0], 0), node1 = set_lsb(s0s[1], 1);
  blk delta = bxor(node0, node1);
  for (int i = 0; i < n - 1; ++i) {
    blk h0 = ht_hash(c, node0), h1 = ht_hash(c, node1);
    int a_bit = in_bit(a, n, i);
    blk cw = bxor(h0, h1);
    if (!a_bit) cw = bxor(cw, delta);
    cws[i].s = cw; cws[i].v = bzero();
    int t0 = get_lsb(node0), t1 = get_lsb(node1);
    node0 = bxor(bxor(h0, bsel(a_bit, node0)), bsel(t0, cw));
    node1 = bxor(bxor(h1, bsel(a_bit, node1)), bsel(t1, cw));
    delta = bxor(node0, node1);
  }
  int a_n = (int)(a & 1);
  int t0 = get_lsb(node0), t1 = get_lsb(node1);
  blk h0_0 = ht_hash(c, set_lsb(node0, 0)), h0_1 = ht_hash(c, set_lsb(node0, 1));
  blk h1_0 = ht_hash(c, set_lsb(node1, 0)), h1_1 = ht_hash(c, set_lsb(node1, 1));
  blk high0_0 = set_lsb(h0_0, 0), high0_1 = set_lsb(h0_1, 0), high1_0 = set_lsb(h1_0, 0), high1_1 = set_lsb(h1_1, 0);
  int low0_0 = get_lsb(h0_0), low0_1 = get_lsb(h0_1), low1_0 = get_lsb(h1_0), low1_1 = get_lsb(h1_1);
  blk hcw = a_n ? bxor(high0_0, high1_0) : bxor(high0_1, high1_1);          /* :123-125 */
  int lcw0 = low0_0 ^ low1_0 ^ !a_n, lcw1 = low0_1 ^ low1_1 ^ a_n;          /* :132-133 */
  cws[n - 1].s = set_lsb(hcw, lcw0);
  cws[n - 1].v = bzero(); cws[n - 1].v.w[0] = (uint32_t)lcw1;               /* :139-141 */
  blk leaf0 = a_n ? set_lsb(high0_1, low0_1) : set_lsb(high0_0, low0_0);
  blk leaf1 = a_n ? set_lsb(high1_1, low1_1) : set_lsb(high1_0, low1_0);
  blk leaf_cw = set_lsb(hcw, a_n ? lcw1 : lcw0);
  if (t0) leaf0 = bxor(leaf0, leaf_cw);
  if (t1) leaf1 = bxor(leaf1, leaf_cw);
  u128 v = g_add(c, g_add(c, g_from(c, b_buf), g_neg(c, g_from(c, set_lsb(leaf0, 0)))),
      g_from(c, set_lsb(leaf1, 0)));
  if (get_lsb(leaf1)) v = g_neg(c, v);
  *ocw = g_into(c, v);
}
static blk ht_last(const octx *c, int b, blk node, const cw32 *cws, blk ocw, int sigma) { /* :208-230, :325-354 */
  const int n = c->p.in_bits;
  int t = get_lsb(node);
  blk h = ht_hash(c, set_lsb(node, sigma));
  blk hcw = set_lsb(cws[n - 1].s, 0);
  int lcw = sigma ? cw_tr(&cws[n - 1]) : get_lsb(cws[n - 1].s);
  blk high = set_lsb(h, 0);
  int low = get_lsb(h);
  if (t) { high = bxor(high, hcw); low ^= lcw; }
  u128 y = g_from(c, high);
  if (low) y = g_add(c, y, g_from(c, ocw));
  if (b) y = g_neg(c, y);
  return g_into(c, y);
}
static blk ht_step(const octx *c, blk node, const cw32 *cw, int x_bit) { /* :191-205 */
  int t = get_lsb(node);
  blk h = ht_hash(c, node);
  return bxor(bxor(h, bsel(x_bit, node)), bsel(t, cw->s));
}
static blk ht_eval(const octx *c, int b, blk s0, const cw32 *cws, blk ocw, u128 x) { /* :187-231 */
  const int n = c->p.in_bits;
  blk node = set_lsb(s0, b);
  for (int i = 0; i < n - 1; ++i) node = ht_step(c, node, &cws[i], in_bit(x, n, i));
  return ht_last(c, b, node, cws, ocw, (int)(x & 1));
}

/* ---- full-domain trees (dpf.cuh:250-303, dcf.cuh:314-385, half_tree_dpf.cuh:284-316,
 *      grotto_dcf.cuh:185-238), pruned to the leaf range [lo, hi) ------------------------------ */
typedef struct {
  const octx *c; int b; const cw32 *cws; blk ocw;
  uint64_t lo, hi; uint8_t *out; /* 16 B per leaf, or 1 B per leaf for grotto */
} walk;

static void dpf_tree(const walk *w, blk st, int i, uint64_t l, uint64_t r, int grotto) {
  if (r <= w->lo || l >= w->hi) return;
  const int n = w->c->p.in_bits;
  if (i == n) {
    if (grotto) w->out[l - w->lo] = (uint8_t)get_lsb(st);                /* grotto_dcf.cuh:190-194 */
    else { blk y = dpf_leaf(w->c, w->b, st, w->cws); memcpy(w->out + 16 * (l - w->lo), &y, 16); }
    return;
  }
  blk left, right;
  dpf_expand(w->c, st, &w->cws[i], &left, &right);
  uint64_t mid = l + ((r - l) >> 1);
  dpf_tree(w, left, i + 1, l, mid, grotto);
  dpf_tree(w, right, i + 1, mid, r, grotto);
}
static void dcf_tree(const walk *w, blk st, u128 v, int i, uint64_t l, uint64_t r) {
  if (r <= w->lo || l >= w->hi) return;
  const int n = w->c->p.in_bits;
  if (i == n) { blk y = dcf_leaf(w->c, w->b, st, v, w->cws); memcpy(w->out + 16 * (l - w->lo), &y, 16); return; }
  blk left, right; u128 vl, vr;
  dcf_expand(w->c, w->b, st, v, &w->cws[i], &left, &right, &vl, &vr);
  uint64_t mid = l + ((r - l) >> 1);
  dcf_tree(w, left, vl, i + 1, l, mid);
  dcf_tree(w, right, vr, i + 1, mid, r);
}
static void ht_tree(const walk *w, blk node, int i, uint64_t l, uint64_t r) {
  if (r <= w->lo || l >= w->hi) return;
  const int n = w->c->p.in_bits;
  if (i == n - 1) { /* node covers leaves l, l+1: half_tree_dpf.cuh:278-280 */
    for (int sigma = 0; sigma < 2; ++sigma) {
      uint64_t x = l + (uint64_t)sigma;
      if (x < w->lo || x >= w->hi) continue;
      blk y = ht_last(w->c, w->b, node, w->cws, w->ocw, sigma);
      memcpy(w->out + 16 * (x - w->lo), &y, 16);
    }
    return;
  }
  int t = get_lsb(node);
  blk h = ht_hash(w->c, node);
  blk left = bxor(h, bsel(t, w->cws[i].s));                              /* :299-302 */
  blk right = bxor(left, node);
  uint64_t mid = l + ((r - l) >> 1);
  ht_tree(w, left, i + 1, l, mid);
  ht_tree(w, right, i + 1, mid, r);
}

/* ---- exported batch API -------------------------------------------------------------------------- */
static int ncw_of(const fssb200_params *p) {
  return (p->scheme == FSSB200_SCHEME_HALFTREE || p->scheme == FSSB200_SCHEME_VDPF) ? p->in_bits : p->in_bits + 1;
}
int orc_ncw(const fssb200_params *p) { return p ? ncw_of(p) : FSSB200_EINVAL; }

int orc_prg_gen(const fssb200_params *p, int mul, size_t n, const void *seeds, void *out) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (mul != 1 && mul != 2 && mul != 4) return FSSB200_EINVAL;
  for (size_t i = 0; i < n; ++i) {
    blk g[4];
    prg_gen(&c, mul, ((const blk *)seeds)[i], g);
    memcpy((blk *)out + i * (size_t)mul, g, 16 * (size_t)mul);
  }
  return 0;
}

int orc_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas, const void *betas,
    void *cws, void *ocws, int threads) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  const int ncw = ncw_of(p);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
  for (size_t k = 0; k < nkeys; ++k) {
    const blk *s = (const blk *)s0s + 2 * k;
    u128 a = load_in((const uint8_t *)alphas + k * (size_t)p->in_bytes, p->in_bytes);
    blk beta = betas ? ((const blk *)betas)[k] : bzero();
    cw32 *kc = (cw32 *)cws + k * (size_t)ncw;
    blk ocw = bzero();
    switch (p->scheme) {
      case FSSB200_SCHEME_DPF: dpf_gen(&c, kc, s, a, beta); break;
      case FSSB200_SCHEME_GROTTO: dpf_gen(&c, kc, s, a, bzero()); break;  /* grotto_dcf.cuh:63-67 */
      case FSSB200_SCHEME_DCF: dcf_gen(&c, kc, s, a, beta); break;
      default: ht_gen(&c, kc, &ocw, s, a, beta); if (ocws) ((blk *)ocws)[k] = ocw; break;
    }
  }
  return 0;
}

int orc_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *ocws, const void *xs, void *ys, int threads) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme == FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  const int ncw = ncw_of(p);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
  for (size_t k = 0; k < nkeys; ++k) {
    blk s0 = ((const blk *)seeds)[k];
    const cw32 *kc = (const cw32 *)cws + k * (size_t)ncw;
    u128 x = load_in((const uint8_t *)xs + k * (size_t)p->in_bytes, p->in_bytes);
    blk y;
    if (p->scheme == FSSB200_SCHEME_DPF) y = dpf_eval(&c, party != 0, s0, kc, x);
    else if (p->scheme == FSSB200_SCHEME_DCF) y = dcf_eval(&c, party != 0, s0, kc, x);
    else y = ht_eval(&c, party != 0, s0, kc, ((const blk *)ocws)[k], x);
    ((blk *)ys)[k] = y;
  }
  return 0;
}

static int range_fix(const fssb200_params *p, uint64_t *begin, uint64_t *count) {
  if (p->in_bits > 40) return FSSB200_EDOMAIN;
  const uint64_t N = (uint64_t)1 << p->in_bits;
  if (*begin >= N) return FSSB200_ERANGE;
  if (*count == 0) *count = N - *begin;
  if (*begin + *count > N) return FSSB200_ERANGE;
  return 0;
}

static void expand_key(const octx *c, int party, blk s0, const cw32 *kc, blk ocw, uint8_t *out, uint64_t lo,
    uint64_t hi, int grotto_bits) {
  const uint64_t N = (uint64_t)1 << c->p.in_bits;
  walk w = {c, party != 0, kc, ocw, lo, hi, out};
  blk st = set_lsb(s0, party != 0);
  switch (c->p.scheme) {
    case FSSB200_SCHEME_DPF: dpf_tree(&w, st, 0, 0, N, 0); break;
    case FSSB200_SCHEME_GROTTO: dpf_tree(&w, st, 0, 0, N, grotto_bits); break;
    case FSSB200_SCHEME_DCF: dcf_tree(&w, st, 0, 0, 0, N); break;
    default:
      if (c->p.in_bits == 1) {                                             /* half_tree_dpf.cuh:253-259 */
        for (int sigma = 0; sigma < 2; ++sigma)
          if ((uint64_t)sigma >= lo && (uint64_t)sigma < hi) {
            blk y = ht_last(c, party != 0, st, kc, ocw, sigma);
            memcpy(out + 16 * ((uint64_t)sigma - lo), &y, 16);
          }
      } else {
        ht_tree(&w, st, 0, 0, N);
      }
  }
}

int orc_evalall(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *ocws, void *ys, uint64_t leaf_begin, uint64_t leaf_count, int threads) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  if ((rc = range_fix(p, &leaf_begin, &leaf_count))) return rc;
  const int grotto = p->scheme == FSSB200_SCHEME_GROTTO;
  if (grotto && leaf_begin) return FSSB200_ERANGE;
  const int ncw = ncw_of(p);
  const size_t leaf_bytes = grotto ? 1 : 16;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
  for (size_t k = 0; k < nkeys; ++k) {
    uint8_t *out = (uint8_t *)ys + k * leaf_count * leaf_bytes;
    expand_key(&c, party, ((const blk *)seeds)[k], (const cw32 *)cws + k * (size_t)ncw,
        ocws ? ((const blk *)ocws)[k] : bzero(), out, leaf_begin, leaf_begin + leaf_count, 1);
    if (grotto) for (uint64_t x = 1; x < leaf_count; ++x) out[x] ^= out[x - 1];  /* grotto_dcf.cuh:160-162 */
  }
  return 0;
}

int orc_grotto_expand(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    void *t, uint64_t leaf_begin, uint64_t leaf_count, int threads) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if ((rc = range_fix(p, &leaf_begin, &leaf_count))) return rc;
  const int ncw = ncw_of(p);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
  for (size_t k = 0; k < nkeys; ++k)
    expand_key(&c, party, ((const blk *)seeds)[k], (const cw32 *)cws + k * (size_t)ncw, bzero(),
        (uint8_t *)t + k * leaf_count, leaf_begin, leaf_begin + leaf_count, 1);
  return 0;
}

int orc_grotto_preprocess(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    void *pt, int threads) { /* grotto_dcf.cuh:94-104 */
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (p->in_bits > 30) return FSSB200_EDOMAIN;
  const uint64_t N = (uint64_t)1 << p->in_bits;
  const int ncw = ncw_of(p);
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
  for (size_t k = 0; k < nkeys; ++k) {
    uint8_t *tree = (uint8_t *)pt + k * (2 * N - 1);
    expand_key(&c, party, ((const blk *)seeds)[k], (const cw32 *)cws + k * (size_t)ncw, bzero(), tree + (N - 1), 0,
        N, 1);
    for (uint64_t j = N - 1; j-- > 0;) tree[j] = tree[2 * j + 1] ^ tree[2 * j + 2];
  }
  return 0;
}

int orc_grotto_lookup(const fssb200_params *p, size_t nkeys, const void *pt, const void *xs, void *ys) {
  /* grotto_dcf.cuh:116-135 */
  if (!p || p->scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (p->in_bits > 30) return FSSB200_EDOMAIN;
  const int n = p->in_bits;
  const uint64_t N = (uint64_t)1 << n;
  const u128 in_mask = p->in_bytes >= 16 ? ~(u128)0 : (((u128)1 << (8 * p->in_bytes)) - 1);
  for (size_t k = 0; k < nkeys; ++k) {
    const uint8_t *tree = (const uint8_t *)pt + k * (2 * N - 1);
    u128 e = (load_in((const uint8_t *)xs + k * (size_t)p->in_bytes, p->in_bytes) + 1) & in_mask; /* In arithmetic */
    uint8_t pi = 0;
    if (e == 0 || e == N) { ((uint8_t *)ys)[k] = tree[0]; continue; }
    uint64_t cur = 0;
    for (int i = 0; i < n; ++i) {
      if ((e >> (n - 1 - i)) & 1) { pi ^= tree[2 * cur + 1]; cur = 2 * cur + 2; }
      else cur = 2 * cur + 1;
    }
    ((uint8_t *)ys)[k] = pi;
  }
  return 0;
}

int orc_relayout(const fssb200_params *p, size_t nkeys, const void *cws, void *cw_s, void *cw_v, void *extra,
    void *out_cw) { /* point_eval_gpu.cuh:39-91, control bits widened to ceil(n/32) words */
  if (!p) return FSSB200_EINVAL;
  const int n = p->in_bits, ncw = ncw_of(p);
  const int words = (n + 31) / 32;
  if (extra) memset(extra, 0, sizeof(uint32_t) * (size_t)words * nkeys);
  for (size_t k = 0; k < nkeys; ++k) {
    const cw32 *kc = (const cw32 *)cws + k * (size_t)ncw;
    for (int i = 0; i < n; ++i) {
      ((blk *)cw_s)[(size_t)i * nkeys + k] = kc[i].s;
      if (p->scheme == FSSB200_SCHEME_DCF) ((blk *)cw_v)[(size_t)i * nkeys + k] = kc[i].v;
      else if (cw_tr(&kc[i])) ((uint32_t *)extra)[(size_t)(i / 32) * nkeys + k] |= 1u << (i % 32);
    }
    if (p->scheme == FSSB200_SCHEME_DCF) ((blk *)out_cw)[k] = kc[n].v;
    else if (p->scheme != FSSB200_SCHEME_HALFTREE) ((blk *)out_cw)[k] = kc[n].s;
  }
  return 0;
}

int orc_group_add(const fssb200_params *p, size_t n, const void *a, const void *b, void *out) {
  octx c; int rc = ctx_init(&c, p);
  if (rc) return rc;
  for (size_t i = 0; i < n; ++i) {
    /* From() asserts a clamped input (bytes.cuh:33, uint.cuh:50) */
    blk x = ((const blk *)a)[i], y = ((const blk *)b)[i];
    ((blk *)out)[i] = g_into(&c, g_add(&c, g_from(&c, x), g_from(&c, y)));
  }
  return 0;
}

/* ---- BLAKE3 keyed compression: hash/blake3.cuh ------------------------------------------------- */
static uint32_t rotr32(uint32_t v, int n) { return (v >> n) | (v << (32 - n)); }
static void b3_g(uint32_t *v, int a, int b, int c, int d, uint32_t x, uint32_t y) { /* :33-42 */
  v[a] = v[a] + v[b] + x; v[d] = rotr32(v[d] ^ v[a], 16);
  v[c] = v[c] + v[d];     v[b] = rotr32(v[b] ^ v[c], 12);
  v[a] = v[a] + v[b] + y; v[d] = rotr32(v[d] ^ v[a], 8);
  v[c] = v[c] + v[d];     v[b] = rotr32(v[b] ^ v[c], 7);
}
/* h: 8 words (the plugin's IV), msg: 16 words, counter = 0; out: 16 words.  :99-124 */
static void b3_compress(const uint32_t h[8], const uint32_t msg[16], uint32_t block_len, uint32_t flags,
    uint32_t out[16]) {
  static const uint32_t iv0[4] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au};  /* :75-80 */
  static const int perm[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};   /* :53 */
  uint32_t v[16], m[16], t[16];
  for (int i = 0; i < 8; ++i) v[i] = h[i];
  for (int i = 0; i < 4; ++i) v[8 + i] = iv0[i];
  v[12] = 0; v[13] = 0; v[14] = block_len; v[15] = flags;                              /* :104 */
  memcpy(m, msg, sizeof(m));
  for (int r = 0; r < 7; ++r) {                                                       /* :109-114 */
    b3_g(v, 0, 4, 8, 12, m[0], m[1]);   b3_g(v, 1, 5, 9, 13, m[2], m[3]);               /* :63-67 */
    b3_g(v, 2, 6, 10, 14, m[4], m[5]);  b3_g(v, 3, 7, 11, 15, m[6], m[7]);
    b3_g(v, 0, 5, 10, 15, m[8], m[9]);  b3_g(v, 1, 6, 11, 12, m[10], m[11]);            /* :68-72 */
    b3_g(v, 2, 7, 8, 13, m[12], m[13]); b3_g(v, 3, 4, 9, 14, m[14], m[15]);
    if (r < 6) { for (int i = 0; i < 16; ++i) t[i] = m[perm[i]]; memcpy(m, t, sizeof(m)); }
  }
  for (int i = 0; i < 8; ++i) { out[i] = v[i] ^ v[8 + i]; out[8 + i] = v[8 + i] ^ h[i]; } /* :117-121 */
}
#define B3_FLAGS (1u | 2u | 8u | 16u) /* CHUNK_START | CHUNK_END | ROOT | KEYED_HASH, :82-85,144 */
/* Blake3::Hash(span<int4,4>) :143-147: 64 B -> first 32 B of the compression output */
static void b3_hash(const uint8_t iv[32], const blk msg[4], blk out[2]) {
  uint32_t h[8], o[16];
  memcpy(h, iv, 32);
  b3_compress(h, (const uint32_t *)msg, 64, B3_FLAGS, o);
  memcpy(out, o, 32);
}
/* Blake3::Hash(tuple<int4,int4>) :158-170: two 32-byte-block compressions, a's lsb 0 / 1 */
static void b3_xor_hash(const uint8_t iv[32], blk a, blk b, blk out[4]) {
  uint32_t h[8], o[16];
  blk padded[4];
  memcpy(h, iv, 32);
  padded[0] = set_lsb(a, 0); padded[1] = b; padded[2] = bzero(); padded[3] = bzero();
  b3_compress(h, (const uint32_t *)padded, 32, B3_FLAGS, o);
  memcpy(&out[0], o, 32);
  padded[0] = set_lsb(a, 1);
  b3_compress(h, (const uint32_t *)padded, 32, B3_FLAGS, o);
  memcpy(&out[2], o, 32);
}
/* ---- SHA-256 keyed hash: hash/sha256.cuh (EVP_Digest over key || message) ------------------------ */
/* FIPS 180-4, byte-oriented: `len` message bytes, digest as 32 bytes. */
static void sha256_bytes(const uint8_t *msg, size_t len, uint8_t digest[32]) {
  static const uint32_t K[64] = {
      0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
      0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
      0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
      0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
      0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
      0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
      0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
      0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
  uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  uint8_t buf[192];
  size_t total = ((len + 9 + 63) / 64) * 64;
  memset(buf, 0, sizeof(buf));
  memcpy(buf, msg, len);
  buf[len] = 0x80;
  for (int i = 0; i < 8; ++i) buf[total - 1 - i] = (uint8_t)(((uint64_t)len * 8) >> (8 * i));
  for (size_t off = 0; off < total; off += 64) {
    uint32_t w[64], s[8];
    for (int i = 0; i < 16; ++i)
      w[i] = ((uint32_t)buf[off + 4 * i] << 24) | ((uint32_t)buf[off + 4 * i + 1] << 16) |
          ((uint32_t)buf[off + 4 * i + 2] << 8) | buf[off + 4 * i + 3];
    for (int i = 16; i < 64; ++i) {
      const uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    memcpy(s, h, sizeof(s));
    for (int i = 0; i < 64; ++i) {
      const uint32_t S1 = rotr32(s[4], 6) ^ rotr32(s[4], 11) ^ rotr32(s[4], 25);
      const uint32_t ch = (s[4] & s[5]) ^ (~s[4] & s[6]);
      const uint32_t t1 = s[7] + S1 + ch + K[i] + w[i];
      const uint32_t S0 = rotr32(s[0], 2) ^ rotr32(s[0], 13) ^ rotr32(s[0], 22);
      const uint32_t maj = (s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]);
      const uint32_t t2 = S0 + maj;
      s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1; s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
    }
    for (int i = 0; i < 8; ++i) h[i] += s[i];
  }
  for (int i = 0; i < 8; ++i) {
    digest[4 * i] = (uint8_t)(h[i] >> 24); digest[4 * i + 1] = (uint8_t)(h[i] >> 16);
    digest[4 * i + 2] = (uint8_t)(h[i] >> 8); digest[4 * i + 3] = (uint8_t)h[i];
  }
}
/* Sha256::Hash(span<int4,4>) hash/sha256.cuh:44-58: SHA-256(key || 64-byte message) */
static void sha_hash(const uint8_t key[16], const blk msg[4], blk out[2]) {
  uint8_t buf[80];
  memcpy(buf, key, 16);
  memcpy(buf + 16, msg, 64);
  sha256_bytes(buf, 80, (uint8_t *)out);
}
/* Sha256::Hash(tuple<int4,int4>) hash/sha256.cuh:69-89: SHA-256(key || a lsb=0 || b) || SHA-256(key || a lsb=1 || b) */
static void sha_xor_hash(const uint8_t key[16], blk a, blk b, blk out[4]) {
  uint8_t buf[48];
  blk a0 = set_lsb(a, 0), a1 = set_lsb(a, 1);
  memcpy(buf, key, 16);
  memcpy(buf + 16, &a0, 16);
  memcpy(buf + 32, &b, 16);
  sha256_bytes(buf, 48, (uint8_t *)&out[0]);
  memcpy(buf + 16, &a1, 16);
  sha256_bytes(buf, 48, (uint8_t *)&out[2]);
}
/* the VDPF hash plugins of a parameter set: fssb200_params::hash, byte 0 = XorHash, byte 1 = Hash (0 Blake3, 1 SHA-256;
   a SHA-256 plugin's key is the first 16 bytes of its hash_iv slot) */
static void p_xor_hash(const fssb200_params *p, blk a, blk b, blk out[4]) {
  if ((p->hash & 0xff) == FSSB200_HASH_SHA256) sha_xor_hash(p->hash_iv[0], a, b, out);
  else b3_xor_hash(p->hash_iv[0], a, b, out);
}
static void p_hash(const fssb200_params *p, const blk msg[4], blk out[2]) {
  if (((p->hash >> 8) & 0xff) == FSSB200_HASH_SHA256) sha_hash(p->hash_iv[1], msg, out);
  else b3_hash(p->hash_iv[1], msg, out);
}
int orc_hash(const fssb200_params *p, int which, size_t n, const void *msgs, void *out) {
  if (!p || !msgs || !out) return FSSB200_EINVAL;
  for (size_t i = 0; i < n; ++i) {
    if (which == 0) {
      const blk *m = (const blk *)msgs + 2 * i;
      p_xor_hash(p, m[0], m[1], (blk *)out + 4 * i);
    } else {
      p_hash(p, (const blk *)msgs + 4 * i, (blk *)out + 2 * i);
    }
  }
  return 0;
}

/* ---- VDPF: vdpf.cuh -------------------------------------------------------------------------------- */
static blk pack_in(const octx *c, u128 x) { /* util.cuh:46-63 Pack<In> */
  blk b = bzero();
  const int nb = c->p.in_bytes;
  b.w[0] = (uint32_t)x;
  if (nb > 4) b.w[1] = (uint32_t)(x >> 32);
  if (nb > 8) { b.w[2] = (uint32_t)(x >> 64); b.w[3] = (uint32_t)(x >> 96); }
  return b;
}
/* Vdpf::Gen :97-177.  Returns 1 when t0 == t1 (the caller resamples), ocw is then left untouched. */
static int vdpf_gen(const octx *c, cw32 *cws, blk cs[4], blk *ocw, const blk s0s[2], u128 a, blk b_buf) {
  const int n = c->p.in_bits;
  blk s0 = set_lsb(s0s[0], 0), s1 = set_lsb(s0s[1], 0);
  int t0 = 0, t1 = 1;
  b_buf = set_lsb(b_buf, 0);
  for (int i = 0; i < n; ++i) {                                   /* same walk as Dpf::Gen */
    blk g0[4], g1[4];
    prg_gen(c, 2, s0, g0);
    prg_gen(c, 2, s1, g1);
    int t0l = get_lsb(g0[0]), t0r = get_lsb(g0[1]), t1l = get_lsb(g1[0]), t1r = get_lsb(g1[1]);
    blk s0l = set_lsb(g0[0], 0), s0r = set_lsb(g0[1], 0), s1l = set_lsb(g1[0], 0), s1r = set_lsb(g1[1], 0);
    int a_bit = in_bit(a, n, i);
    blk s_cw = a_bit ? bxor(s0l, s1l) : bxor(s0r, s1r);
    int tl_cw = t0l ^ t1l ^ a_bit ^ 1, tr_cw = t0r ^ t1r ^ a_bit;
    if (!a_bit) {
      s0 = bxor(s0l, bsel(t0, s_cw)); s1 = bxor(s1l, bsel(t1, s_cw));
      t0 = t0l ^ (t0 & tl_cw); t1 = t1l ^ (t1 & tl_cw);
    } else {
      s0 = bxor(s0r, bsel(t0, s_cw)); s1 = bxor(s1r, bsel(t1, s_cw));
      t0 = t0r ^ (t0 & tr_cw); t1 = t1r ^ (t1 & tr_cw);
    }
    cws[i].s = set_lsb(s_cw, tl_cw);
    cws[i].v = bzero(); cws[i].v.w[0] = (uint32_t)tr_cw;          /* :147-149 */
  }
  blk a_buf = pack_in(c, a), p0[4], p1[4];                         /* :153-157 */
  p_xor_hash(&c->p, a_buf, s0, p0);
  p_xor_hash(&c->p, a_buf, s1, p1);
  for (int j = 0; j < 4; ++j) cs[j] = bxor(p0[j], p1[j]);
  if (t0 == t1) return 1;                                          /* :160 */
  u128 v = g_add(c, g_add(c, g_from(c, b_buf), g_neg(c, g_from(c, s0))), g_from(c, s1));
  if (t1) v = g_neg(c, v);
  *ocw = g_into(c, v);
  return 0;
}
/* y share and corrected per-point hash of one packed leaf (s | t) at input x: :224-242, :318-331 */
static void vdpf_leaf(const octx *c, int b, blk st, const blk cs[4], blk ocw, u128 x, blk *y, blk pi_tilde[4]) {
  int t = get_lsb(st);
  blk s = set_lsb(st, 0);
  u128 g = g_from(c, s);
  if (t) g = g_add(c, g, g_from(c, ocw));
  if (b) g = g_neg(c, g);
  *y = g_into(c, g);
  p_xor_hash(&c->p, pack_in(c, x), s, pi_tilde);
  if (t) for (int j = 0; j < 4; ++j) pi_tilde[j] = bxor(pi_tilde[j], cs[j]);
}
static void vdpf_eval(const octx *c, int b, blk s0, const cw32 *cws, const blk cs[4], blk ocw, u128 x, blk *y,
    blk pi_tilde[4]) { /* :191-243 */
  const int n = c->p.in_bits;
  blk st = set_lsb(s0, b);
  for (int i = 0; i < n; ++i) {
    blk l, r;
    dpf_expand(c, st, &cws[i], &l, &r);
    st = in_bit(x, n, i) ? r : l;
  }
  vdpf_leaf(c, b, st, cs, ocw, x, y, pi_tilde);
}
/* one step of Vdpf::Prove :257-263 / EvalAll :336-340 */
static void vdpf_accumulate(const octx *c, blk pi[4], const blk pi_tilde[4]) {
  blk in[4], h[2];
  for (int j = 0; j < 4; ++j) in[j] = bxor(pi[j], pi_tilde[j]);
  p_hash(&c->p, in, h);
  pi[0] = bxor(pi[0], h[0]);
  pi[1] = bxor(pi[1], h[1]);
}
static void vdpf_tree(const octx *c, blk st, const cw32 *cws, blk *leaves, int i, uint64_t l, uint64_t r) { /* :345-401 */
  if (i == c->p.in_bits) { leaves[l] = st; return; }
  blk a, b;
  dpf_expand(c, st, &cws[i], &a, &b);
  const uint64_t mid = (l + r) / 2;
  vdpf_tree(c, a, cws, leaves, i + 1, l, mid);
  vdpf_tree(c, b, cws, leaves, i + 1, mid, r);
}

int orc_vdpf_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas, const void *betas,
    void *cws, void *cs, void *ocws, void *status, int threads) {
  octx c;
  int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  const int n = p->in_bits;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
  for (size_t k = 0; k < nkeys; ++k) {
    const u128 a = load_in((const uint8_t *)alphas + k * (size_t)p->in_bytes, p->in_bytes);
    ((int32_t *)status)[k] = vdpf_gen(&c, (cw32 *)cws + k * (size_t)n, (blk *)cs + 4 * k, (blk *)ocws + k,
        (const blk *)s0s + 2 * k, a, ((const blk *)betas)[k]);
  }
  return 0;
}
int orc_vdpf_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *cs, const void *ocws, const void *xs, void *ys, void *pis, int threads) {
  octx c;
  int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  const int n = p->in_bits;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(static)
  for (size_t k = 0; k < nkeys; ++k) {
    const u128 x = load_in((const uint8_t *)xs + k * (size_t)p->in_bytes, p->in_bytes);
    vdpf_eval(&c, party, ((const blk *)seeds)[k], (const cw32 *)cws + k * (size_t)n, (const blk *)cs + 4 * k,
        ((const blk *)ocws)[k], x, (blk *)ys + k, (blk *)pis + 4 * k);
  }
  return 0;
}
int orc_vdpf_prove(const fssb200_params *p, size_t nkeys, size_t m, const void *pi_tildes, const void *cs,
    void *pis) {
  octx c;
  int rc = ctx_init(&c, p);
  if (rc) return rc;
  for (size_t k = 0; k < nkeys; ++k) {
    blk *pi = (blk *)pis + 4 * k;
    memcpy(pi, (const blk *)cs + 4 * k, 64);                        /* :256 */
    for (size_t i = 0; i < m; ++i) vdpf_accumulate(&c, pi, (const blk *)pi_tildes + 4 * (k * m + i));
  }
  return 0;
}
int orc_vdpf_evalall(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *cs, const void *ocws, void *ys, void *pis, int threads) {
  octx c;
  int rc = ctx_init(&c, p);
  if (rc) return rc;
  if (p->scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  const int n = p->in_bits;
  if (n > 30) return FSSB200_EDOMAIN;
  const uint64_t N = (uint64_t)1 << n;
#pragma omp parallel for num_threads(threads > 0 ? threads : 1) schedule(dynamic, 1)
  for (size_t k = 0; k < nkeys; ++k) {
    blk *out = (blk *)ys + k * N, *pi = (blk *)pis + 4 * k;
    const blk *kcs = (const blk *)cs + 4 * k;
    vdpf_tree(&c, set_lsb(((const blk *)seeds)[k], party), (const cw32 *)cws + k * (size_t)n, out, 0, 0, N); /* :311 */
    memcpy(pi, kcs, 64);                                            /* :314 */
    for (uint64_t j = 0; j < N; ++j) {                              /* :318-341 */
      blk y, pt[4];
      vdpf_leaf(&c, party, out[j], kcs, ((const blk *)ocws)[k], (u128)j, &y, pt);
      out[j] = y;
      vdpf_accumulate(&c, pi, pt);
    }
  }
  return 0;
}
