// SPDX-License-Identifier: Apache-2.0
//
// kernels.cuh -- sm_100a kernels: batched point evaluation, key generation, full-domain
// evaluation.  Instantiated per (scheme, group kind, PRG) in kernels_*.cu.
//
// Shared-memory plan of the AES kernels (one persistent CTA per SM, 227 KB dynamic smem):
//   [a0, a0+128K)    lane-replicated T-tables, a0 = first 64 KiB-aligned shared-window address
//   [base, a0)       "low" scratch  (63 KB when the driver reserves the first 1 KB)
//   [a0+128K, end)   "high" scratch (36 KB)
// ChaCha kernels have no tables; their scratch is one region.
#pragma once
#include "schemes.cuh"

namespace fssb200 {

struct KParams {
  PrgKeys keys;
  GroupArgs ga;
};

constexpr int kPointThreads = 512;   // AES point / gen kernels: 16 warps, 1 CTA per SM
constexpr int kEvalAllThreads = 512;
constexpr int kEvalAllThreadBits = 9;
constexpr int kMaxDfsBits = 8;
constexpr uint32_t kMaxDynSmem = 232448;  // 227 KB opt-in limit on sm_100

// ---- shared-memory helpers ----------------------------------------------------------------------------------
FSS_D uint32_t smem_base_addr() {
  extern __shared__ __align__(16) uint8_t fss_dyn_smem[];
  return static_cast<uint32_t>(__cvta_generic_to_shared(fss_dyn_smem));
}
FSS_D uint32_t dyn_smem_size() {
  uint32_t d;
  asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(d));
  return d;
}
FSS_D blk lds_blk(uint32_t addr) {
  blk b;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "r"(addr) : "memory");
  return b;
}
FSS_D void sts_blk(uint32_t addr, blk b) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
// two adjacent leaves with one 256-bit store (STG.E.256)
FSS_D void stg_blk2(void *p, blk a, blk b) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}

struct SmemPlan {
  uint32_t a0;               // table base (AES) -- 64 KiB aligned
  uint32_t lo, lo_end;       // scratch region below the tables
  uint32_t hi, hi_end;       // scratch region above the tables
  // bump allocation, `align`-byte aligned (power of two >= 16); prefers the region given by `want_hi`
  FSS_D uint32_t alloc(uint32_t bytes, bool want_hi, uint32_t align = 16u) {
    bytes = (bytes + 15u) & ~15u;
    for (int pass = 0; pass < 2; ++pass) {
      const bool use_hi = (pass == 0) ? want_hi : !want_hi;
      if (use_hi) {
        const uint32_t r = (hi + align - 1u) & ~(align - 1u);
        if (r + bytes <= hi_end) { hi = r + bytes; return r; }
      } else {
        const uint32_t r = (lo + align - 1u) & ~(align - 1u);
        if (r + bytes <= lo_end) { lo = r + bytes; return r; }
      }
    }
    __trap();  // host-side geometry (api.cu: plan_evalall) guarantees this cannot happen
    return 0;
  }
};

template <int PRG>
FSS_D SmemPlan smem_plan() {
  SmemPlan p;
  const uint32_t base = smem_base_addr();
  const uint32_t end = base + dyn_smem_size();
  if (Prg<PRG>::kNeedsTables) {
    p.a0 = (base + 0xffffu) & ~0xffffu;
    if (p.a0 + kAesTblBytes > end) __trap();
    p.lo = (base + 15u) & ~15u;
    p.lo_end = p.a0;
    p.hi = p.a0 + kAesTblBytes;
    p.hi_end = end;
  } else {
    p.a0 = 0;
    p.lo = p.lo_end = 0;
    p.hi = (base + 15u) & ~15u;
    p.hi_end = end;
  }
  return p;
}

template <int PRG>
FSS_D typename Prg<PRG>::ctx_t prg_ctx_init(const SmemPlan &sp);
template <>
FSS_D AesCtx prg_ctx_init<kPrgAes>(const SmemPlan &sp) {
  aes_tables_init(sp.a0);
  __syncthreads();
  AesCtx c;
  c.laneoff = sp.a0 | ((threadIdx.x & 31u) << 2);
  return c;
}
template <>
FSS_D NoCtx prg_ctx_init<kPrgChaCha>(const SmemPlan &) {
  return NoCtx{};
}

// ---- correction words staged through shared memory -------------------------------------------------------------
// The reference layout is key-major (stride (n+1)*32 B per key), so a lane-per-key walk touches 32 different
// lines per level, and -- worse -- a per-level global load issued next to the AES code shares a hardware
// scoreboard slot with the table lookups: ncu showed 35 % of all stall samples on one LOP3 waiting for
// the *prefetch* (profiles/r01_dpf_point_v1.md).  Here each warp copies the next L levels of its 32 keys
// into its own shared-memory slab with cp.async (LDGSTS, no registers, no scoreboard): 2L consecutive
// 16-byte pieces per key are read by 2L adjacent lanes (full 32*L-byte segments), the slab slot stride of
// 32L+16 bytes makes the later lane-per-key 128-bit reads conflict-free, and the chunk after the next is
// prefetched into L2 while this one is consumed.
template <int L>
struct CwStagedWarp {
  static constexpr uint32_t kSlot = 32u * L + 16u;
  static constexpr uint32_t kWarpBytes = 32u * kSlot;
  uint32_t buf;          // shared-window address of this warp's slab
  const uint8_t *gsrc;   // first key of the warp's tile: Cw[.][ncw]
  uint32_t key_bytes;    // ncw * 32
  int ncw;
  int nvalid;            // keys of the tile that exist (1..32)
  uint32_t lane;

  FSS_D void load(int i0) const {
    __syncwarp();  // every lane has taken what it needs from the previous chunk
#pragma unroll
    for (int j = 0; j < 2 * L; ++j) {
      const uint32_t pidx = uint32_t(j) * 32u + lane;
      const uint32_t key = pidx / (2u * L), piece = pidx % (2u * L);
      const bool ok = int(key) < nvalid && i0 + int(piece >> 1) < ncw;
      const uint8_t *src = gsrc + (ok ? uint64_t(key) * key_bytes + uint64_t(i0) * 32u + piece * 16u : 0u);
      const uint32_t dst = buf + key * kSlot + piece * 16u;
      const uint32_t nbytes = ok ? 16u : 0u;  // 0: zero-fill, nothing is read
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      if (ok && i0 + int(piece >> 1) + L < ncw)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + 32u * L));
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
  }
  FSS_D void begin_level(int i) const {
    if ((i & (L - 1)) == 0 && i < ncw) load(i);
  }
  FSS_D void done_level(int) const {}
  FSS_D uint32_t at(int i) const { return buf + lane * kSlot + uint32_t(i & (L - 1)) * 32u; }
  FSS_D blk s(int i) const { return lds_blk(at(i)); }
  FSS_D blk v(int i) const { return lds_blk(at(i) + 16u); }
  FSS_D uint32_t flag(int i) const {  // the C++ bool at byte 16 of Dpf::Cw / HalfTreeDpf::Cw
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(at(i) + 16u) : "memory");
    return (w & 0xffu) != 0;
  }
  FSS_D blk out_s(int n) const { return s(n); }
  FSS_D blk out_v(int n) const { return v(n); }
};

// ---- correction words fetched by the TMA unit (point modes 4 / 5) ---------------------------------------------------
// ncu of the cp.async version above (profiles/r01_ncu_full.md): LSU data pipe 91 % busy, 7.9 % of its shared-memory
// wavefronts are the LDGSTS writes, their 2-way conflicts and the slab reads, plus ~30 integer instructions of
// address / predicate glue per level and a warp-wide wait per chunk.  Here the key-major Cw array is described to
// the TMA unit as a 2-D byte tensor [nkeys][ncw*32] and ONE `cp.async.bulk.tensor.2d` per warp and chunk brings the
// next two levels of the warp's 32 keys (box = 32 rows x 64 B) into a double-buffered 2 KB tile, completion counted
// on a per-warp mbarrier.  No LSU instructions, registers or address arithmetic for the copy; the chunk after the
// current one (also across tiles) is always in flight, so the wait does not stall; SWIZZLE_64B (16-byte chunk index
// ^= address bits 7..8) makes the lane-per-key 128-bit reads conflict-free without padding; rows / levels outside
// the tensor are zero-filled by the hardware, so ragged tiles and odd ncw need no predicates.
//
// PACKED = true (point mode 6, fssb200_eval_packed): the rows are the compact key format of fssb200_pack_rows --
// ncw 16-byte `s` entries followed by one 16-byte word of flag bits (bit i = the bool at byte 16 of entry i) -- so a
// 64-byte chunk holds FOUR levels and the flags travel in registers.  Same box, swizzle and barriers.
template <bool PACKED>
struct CwTileT {
  static constexpr uint32_t kBuf = 2048u;          // 32 keys x 64 B (2 levels x 32 B, or 4 packed levels x 16 B)
  static constexpr uint32_t kWarpBytes = 2u * kBuf;
  static constexpr int kLpcBits = PACKED ? 2 : 1;  // log2(levels per chunk)
  uint32_t buf;        // the warp's two tiles (512-byte aligned)
  uint32_t mbar;       // the warp's two mbarriers
  const void *tmap;
  uint32_t lane, rowoff, sw;
  int row0, next_row0;  // first key of this tile / of the warp's next tile (-1: none)
  int nchunks;
  uint32_t *seq;       // chunks this warp has waited for so far (kernel lifetime; buffer = seq & 1, parity = seq >> 1)
  blk fl;              // PACKED: this key's flag bits

  static FSS_D void init_barriers(uint32_t mbar, uint32_t lane) {
    if (lane == 0) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar + 8u) : "memory");
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();
  }
  // lane 0: request levels [2c, 2c+2) of keys [row, row+32) into buffer (q & 1)
  static FSS_D void issue(const void *tmap, uint32_t buf, uint32_t mbar, uint32_t q, int c, int row) {
    const uint32_t b = q & 1u, mb = mbar + 8u * b;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(kBuf) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            buf + b * kBuf),
        "l"(tmap), "r"(c * 64), "r"(row), "r"(mb)
        : "memory");
  }
  FSS_D void begin_level(int j) const {
    if (j & ((1 << kLpcBits) - 1)) return;
    const int c = j >> kLpcBits;
    const uint32_t q = *seq;
    __syncwarp();  // every lane is done with chunk q-1, whose buffer the next request overwrites
    if (lane == 0) {
      if (c + 1 < nchunks) issue(tmap, buf, mbar, q + 1u, c + 1, row0);
      else if (next_row0 >= 0) issue(tmap, buf, mbar, q + 1u, 0, next_row0);
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FSS_TILE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra FSS_TILE_WAIT;\n"
        "}\n" ::"r"(mbar + 8u * (q & 1u)),
        "r"((q >> 1) & 1u)
        : "memory");
    *seq = q + 1u;
  }
  FSS_D void done_level(int) const {}
  // 16-byte piece h (0 = s, 1 = v / flag word) of level j in this lane's row of the current tile
  FSS_D uint32_t at(int j, uint32_t h) const {
    const uint32_t b = (*seq - 1u) & 1u;
    const uint32_t piece = PACKED ? (uint32_t(j) & 3u) : (uint32_t(j) & 1u) * 2u + h;
    return buf + b * kBuf + rowoff + ((piece ^ sw) << 4);
  }
  FSS_D blk s(int j) const { return lds_blk(at(j, 0)); }
  FSS_D blk v(int j) const { return lds_blk(at(j, 1)); }  // (not PACKED: DCF rows have no padding to strip)
  FSS_D uint32_t flag(int j) const {
    if (PACKED) {
      const uint32_t w = j < 64 ? (j < 32 ? fl.x : fl.y) : (j < 96 ? fl.z : fl.w);
      return (w >> (uint32_t(j) & 31u)) & 1u;
    }
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(at(j, 1)) : "memory");
    return (w & 0xffu) != 0;
  }
  FSS_D blk out_s(int n) const { return s(n); }
  FSS_D blk out_v(int n) const { return v(n); }
};
typedef CwTileT<false> CwTile;

// ---- level-major arrays fetched by the TMA unit (point mode 7, fssb200_eval_levelmajor) -----------------------------------
// The level-major layout of fssb200_relayout (the reference's GPU layout, point_eval_gpu.cuh:39-91) keeps level i of all keys
// contiguous: cw_s[i][key] (16 B), DCF also cw_v[i][key], the control bits bit-packed in `extra`.  Mode 2 reads it with
// per-level global loads next to the AES code (0.919 of the LDS ceiling: latency and scoreboard slots, not wavefronts).
// Here the arrays are 2-D uint32 tensors [n][nkeys * 4] and one `cp.async.bulk.tensor.2d` per warp and chunk brings the
// next 4 levels (DCF: 2 levels of s and of v, two requests) of the warp's 32 keys into the same double-buffered 2 KB tile
// as CwTile: rows of 512 contiguous bytes, lane = key reads conflict-free without a swizzle, the control bits travel in
// registers (one word per 32 levels, loaded per tile), 4 shared-memory wavefronts per level and no per-level flag read.
// Levels past the array (the chunk that holds "entry n" of the key-major walk) are zero-filled by the hardware, so the
// chunk sequence is the key-major one and the scheme bodies need no change.
template <bool DCF>
struct CwLmTile {
  static constexpr uint32_t kBuf = 2048u;
  static constexpr uint32_t kWarpBytes = 2u * kBuf;
  static constexpr int kLpcBits = DCF ? 1 : 2;  // log2(levels per chunk)
  uint32_t buf, mbar;
  const void *tmap_s, *tmap_v;
  uint32_t lane;
  int key0, next_key0;  // first key of this tile / of the warp's next tile (-1: none)
  int nchunks;
  uint32_t *seq;
  blk fl;               // this key's control bits (not DCF)
  const blk *out_cw;    // this key's output correction word
  static FSS_D void issue(const void *tmap_s, const void *tmap_v, uint32_t buf, uint32_t mbar, uint32_t q, int c, int key0) {
    const uint32_t b = q & 1u, mb = mbar + 8u * b;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(kBuf) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            buf + b * kBuf),
        "l"(tmap_s), "r"(key0 * 4), "r"(c << kLpcBits), "r"(mb)
        : "memory");
    if (DCF)
      asm volatile(
          "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
              buf + b * kBuf + 1024u),
          "l"(tmap_v), "r"(key0 * 4), "r"(c << kLpcBits), "r"(mb)
          : "memory");
  }
  FSS_D void begin_level(int j) const {
    if (j & ((1 << kLpcBits) - 1)) return;
    const int c = j >> kLpcBits;
    const uint32_t q = *seq;
    __syncwarp();  // every lane is done with chunk q-1, whose buffer the next request overwrites
    if (lane == 0) {
      if (c + 1 < nchunks) issue(tmap_s, tmap_v, buf, mbar, q + 1u, c + 1, key0);
      else if (next_key0 >= 0) issue(tmap_s, tmap_v, buf, mbar, q + 1u, 0, next_key0);
    }
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FSS_LMTILE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra FSS_LMTILE_WAIT;\n"
        "}\n" ::"r"(mbar + 8u * (q & 1u)),
        "r"((q >> 1) & 1u)
        : "memory");
    *seq = q + 1u;
  }
  FSS_D void done_level(int) const {}
  FSS_D uint32_t at(int j, uint32_t h) const {
    const uint32_t b = (*seq - 1u) & 1u;
    return buf + b * kBuf + (DCF ? h * 1024u + (uint32_t(j) & 1u) * 512u : (uint32_t(j) & 3u) * 512u) + lane * 16u;
  }
  FSS_D blk s(int j) const { return lds_blk(at(j, 0)); }
  FSS_D blk v(int j) const { return lds_blk(at(j, 1)); }
  FSS_D uint32_t flag(int j) const {
    const uint32_t w = j < 64 ? (j < 32 ? fl.x : fl.y) : (j < 96 ? fl.z : fl.w);
    return (w >> (uint32_t(j) & 31u)) & 1u;
  }
  FSS_D blk out_s(int) const { return ld_blk(out_cw); }
  FSS_D blk out_v(int) const { return ld_blk(out_cw); }
};

// ---- batched point evaluation --------------------------------------------------------------------------------
// One key per thread; warps own tiles of 32 consecutive keys (grid-stride over tiles).
// SCHEME: FSSB200_SCHEME_{DPF,DCF,HALFTREE,VDPF}, or GROTTO = the O(n) Grotto point walk (schemes.cuh).
// MODE: 0 = key-major, staged, <= 512 threads, L = 4      (default)
//       1 = key-major, staged, 1024 threads (<= 64 regs), L = 2
//       2 = level-major arrays (fssb200_eval_levelmajor), direct coalesced loads
//       3 = key-major, direct per-thread global loads (the round-1 first version; kept for A/B runs)
//       4 = key-major, TMA tiles (CwTile), 512 threads
//       5 = key-major, TMA tiles (CwTile), 768 threads (<= 85 regs)
//       6 = packed rows (fssb200_pack_rows / fssb200_eval_packed), TMA tiles, 768 threads; DPF / Half-Tree only
//       7 = level-major arrays, TMA tiles (CwLmTile), 768 threads
constexpr int kPointModes = 8;
template <int MODE>
struct PointMode {
  static constexpr int kMaxThreads = MODE == 1 ? 1024 : (MODE >= 5 ? 768 : 512);
  static constexpr int kL = MODE == 1 ? 2 : 4;
  static constexpr bool kStaged = MODE <= 1;
  static constexpr bool kTma = MODE >= 4 && MODE <= 6;
  static constexpr bool kLmTma = MODE == 7;
};

template <int SCHEME, int G, int PRG, class Cw>
FSS_D blk point_eval_one(const KParams &P, const typename Prg<PRG>::ctx_t &pc, const PointArgs &A, blk s0,
    const InVal &x, const Cw &cw, uint64_t k, bool valid) {
  const int n = A.in_bits;
  if (SCHEME == FSSB200_SCHEME_VDPF) {
    blk pi[4];
    const blk y = vdpf_eval_body<G, PRG>(P.keys, P.ga, pc, n, uint32_t(A.party), s0, x, cw, ld_blk(A.ocws + k),
        A.cs + 4 * k, pi);
    if (valid) {
#pragma unroll
      for (int j = 0; j < 4; ++j) st_blk(A.pis + 4 * k + j, pi[j]);
    }
    return y;
  }
  if (SCHEME == FSSB200_SCHEME_GROTTO)  // O(n) walk: the share bit travels in .x (fssb200_grotto_eval_walk)
    return make_blk(grotto_walk_body<PRG>(P.keys, pc, n, A.in_bytes, uint32_t(A.party), s0, x, cw), 0u, 0u, 0u);
  if (SCHEME == FSSB200_SCHEME_DPF) return dpf_eval_body<G, PRG>(P.keys, P.ga, pc, n, uint32_t(A.party), s0, x, cw);
  if (SCHEME == FSSB200_SCHEME_DCF) return dcf_eval_body<G, PRG>(P.keys, P.ga, pc, n, uint32_t(A.party), s0, x, cw);
  return ht_eval_body<G, PRG>(P.keys, P.ga, pc, n, uint32_t(A.party), s0, x, cw, ld_blk(A.ocws + k));
}

template <int SCHEME, int G, int PRG, int MODE>
__global__ void __launch_bounds__(PointMode<MODE>::kMaxThreads, 1)
point_kernel(const __grid_constant__ KParams P, const __grid_constant__ PointArgs A) {
  typedef PointMode<MODE> PM;
  SmemPlan sp = smem_plan<PRG>();
  const typename Prg<PRG>::ctx_t pc = prg_ctx_init<PRG>(sp);
  const int n = A.in_bits;
  const int ncw = (SCHEME == FSSB200_SCHEME_HALFTREE || SCHEME == FSSB200_SCHEME_VDPF) ? n : n + 1;
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t slab = 0, mbar = 0;
  if (PM::kStaged) {
    for (uint32_t w = 0; w < nwarps; ++w) {  // same allocation sequence in every thread
      const uint32_t a = sp.alloc(CwStagedWarp<PM::kL>::kWarpBytes, false);
      if (w == wid) slab = a;
    }
  }
  const uint64_t ntiles = (A.nkeys + 31) >> 5;
  const uint64_t tile_stride = uint64_t(gridDim.x) * nwarps;
  // Warp w of CTA b takes tiles w * gridDim + b, + tile_stride, ...: consecutive tiles go to DIFFERENT SMs, so a batch that
  // does not fill every warp (a 2^16-key chunk of the host pipeline = 2048 tiles for 3552 warps) and the last, partial round
  // of any batch load all SMs evenly instead of filling the first CTAs (the kernels are bound by each SM's lookup pipe).
  const uint64_t tile0 = uint64_t(wid) * gridDim.x + blockIdx.x;
  uint32_t tma_seq = 0;
  if (PM::kTma) {
    mbar = sp.alloc(16u * nwarps, false) + 16u * wid;
    for (uint32_t w = 0; w < nwarps; ++w) {
      const uint32_t a = sp.alloc(CwTile::kWarpBytes, false, 512u);
      if (w == wid) slab = a;
    }
    CwTile::init_barriers(mbar, lane);
    if (lane == 0 && tile0 < ntiles) CwTile::issue(A.tmap, slab, mbar, 0u, 0, int(tile0 * 32));
  }
  typedef CwLmTile<SCHEME == FSSB200_SCHEME_DCF> LmTile;
  if (PM::kLmTma) {
    mbar = sp.alloc(16u * nwarps, false) + 16u * wid;
    for (uint32_t w = 0; w < nwarps; ++w) {
      const uint32_t a = sp.alloc(LmTile::kWarpBytes, false, 512u);
      if (w == wid) slab = a;
    }
    CwTile::init_barriers(mbar, lane);
    if (lane == 0 && tile0 < ntiles) LmTile::issue(A.tmap, A.tmap2, slab, mbar, 0u, 0, int(tile0 * 32));
  }
  for (uint64_t tile = tile0; tile < ntiles; tile += tile_stride) {
    const uint64_t k = tile * 32 + lane;
    const bool valid = k < A.nkeys;
    const uint64_t kk = valid ? k : A.nkeys - 1;  // idle lanes of a ragged tile shadow the last key
    const blk s0 = ld_blk(A.seeds + kk);
    const InVal x = load_in(A.xs + kk * uint64_t(A.in_bytes), A.in_bytes);
    blk y;
    if (PM::kStaged) {
      CwStagedWarp<PM::kL> cw;
      cw.buf = slab;
      cw.gsrc = A.cws + tile * 32u * uint64_t(ncw) * 32u;
      cw.key_bytes = uint32_t(ncw) * 32u;
      cw.ncw = ncw;
      const uint64_t left = A.nkeys - tile * 32;
      cw.nvalid = left < 32 ? int(left) : 32;
      cw.lane = lane;
      y = point_eval_one<SCHEME, G, PRG>(P, pc, A, s0, x, cw, kk, valid);
      __syncwarp();
    } else if (PM::kTma) {
      CwTileT<MODE == 6> cw;
      cw.buf = slab;
      cw.mbar = mbar;
      cw.tmap = A.tmap;
      cw.lane = lane;
      cw.rowoff = lane * 64u;
      cw.sw = (lane >> 1) & 3u;
      cw.row0 = int(tile * 32);
      cw.next_row0 = tile + tile_stride < ntiles ? int((tile + tile_stride) * 32) : -1;
      cw.nchunks = MODE == 6 ? (ncw + 3) >> 2 : (ncw + 1) >> 1;
      cw.seq = &tma_seq;
      cw.fl = MODE == 6 ? ld_blk(A.cws + kk * (uint64_t(ncw) * 16u + 16u) + uint64_t(ncw) * 16u) : blk{0u, 0u, 0u, 0u};
      y = point_eval_one<SCHEME, G, PRG>(P, pc, A, s0, x, cw, kk, valid);
    } else if (PM::kLmTma) {
      LmTile cw;
      cw.buf = slab;
      cw.mbar = mbar;
      cw.tmap_s = A.tmap;
      cw.tmap_v = A.tmap2;
      cw.lane = lane;
      cw.key0 = int(tile * 32);
      cw.next_key0 = tile + tile_stride < ntiles ? int((tile + tile_stride) * 32) : -1;
      cw.nchunks = (ncw + (1 << LmTile::kLpcBits) - 1) >> LmTile::kLpcBits;
      cw.seq = &tma_seq;
      cw.fl = blk{0u, 0u, 0u, 0u};
      if (SCHEME != FSSB200_SCHEME_DCF) {  // ceil(n / 32) words of control bits per key
        cw.fl.x = A.extra[kk];
        if (n > 32) cw.fl.y = A.extra[A.nkeys + kk];
        if (n > 64) cw.fl.z = A.extra[2 * A.nkeys + kk];
        if (n > 96) cw.fl.w = A.extra[3 * A.nkeys + kk];
      }
      cw.out_cw = A.out_cw ? A.out_cw + kk : nullptr;
      y = point_eval_one<SCHEME, G, PRG>(P, pc, A, s0, x, cw, kk, valid);
    } else if (MODE == 2) {
      const CwLevelMajor cw{A.cw_s, A.cw_v, A.extra, A.out_cw, A.nkeys, kk};
      y = point_eval_one<SCHEME, G, PRG>(P, pc, A, s0, x, cw, kk, valid);
    } else {
      const CwKeyMajor cw{A.cws + kk * uint64_t(ncw) * 32u};
      y = point_eval_one<SCHEME, G, PRG>(P, pc, A, s0, x, cw, kk, valid);
    }
    if (valid) {
      if (SCHEME == FSSB200_SCHEME_GROTTO) reinterpret_cast<uint8_t *>(A.ys)[k] = uint8_t(y.x);  // bool ys[nkeys]
      else st_blk(A.ys + k, y);
    }
  }
}

// ---- correction words written by the TMA unit (gen kernels, OUT = 1) -------------------------------------------------------
// The first gen kernels stored each 32-byte entry straight into the key-major array: two STG.128 per level whose 32
// lanes hit 32 different rows (stride ncw*32 B), i.e. ~64 LSU wavefronts per warp and level next to the 320 (Half-Tree)
// .. 1280 (DCF) wavefronts of the level's AES lookups -- the same data pipe.  Here a warp parks two levels of its 32
// keys in a 2 KB shared-memory tile (same swizzled layout as CwTile: conflict-free lane-per-key 128-bit stores) and lane
// 0 hands the tile to the TMA unit (`cp.async.bulk.tensor.2d.global.shared::cta`); rows / levels outside the tensor
// are clipped by the hardware.  Double-buffered: a buffer is rewritten only after the store issued from it two
// chunks earlier has read it (`cp.async.bulk.wait_group.read 1`).
struct CwTileOut {
  static constexpr uint32_t kBuf = 2048u;
  static constexpr uint32_t kWarpBytes = 2u * kBuf;
  uint32_t buf;        // the warp's two tiles (512-byte aligned)
  const void *tmap;
  uint32_t lane, rowoff, sw;
  int row0;            // first key of the warp's tile
  int ncw;
  uint32_t *seq;       // chunks this warp has stored so far (kernel lifetime; buffer = seq & 1)

  FSS_D void put(int i, blk s, blk v) const {
    const uint32_t q = *seq, h = uint32_t(i) & 1u;
    const uint32_t tile = buf + (q & 1u) * kBuf;
    if (h == 0) {
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      __syncwarp();
    }
    sts_blk(tile + rowoff + (((h * 2u) ^ sw) << 4), s);
    sts_blk(tile + rowoff + (((h * 2u + 1u) ^ sw) << 4), v);
    if (h == 1 || i == ncw - 1) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the TMA unit
      __syncwarp();
      if (lane == 0) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tmap),
                     "r"((i >> 1) * 64), "r"(row0), "r"(tile)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      *seq = q + 1u;
    }
  }
  static FSS_D void drain(uint32_t lane) {  // before the CTA's shared memory goes away
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }
};

// ---- batched key generation --------------------------------------------------------------------------------------
// One key per thread, both parties in lock-step; warps own tiles of 32 consecutive keys.
// OUT: 0 = direct key-major stores (first version, FSSB200_GEN_MODE=0), 1 = TMA-written tiles (CwTileOut).
template <int SCHEME, int G, int PRG, class Out>
FSS_D void gen_one(const KParams &P, const typename Prg<PRG>::ctx_t &pc, const GenArgs &A, uint64_t k, bool valid,
    const Out &out) {
  const int n = A.in_bits;
  const blk s0 = ld_blk(A.s0s + 2 * k), s1 = ld_blk(A.s0s + 2 * k + 1);
  const InVal a = load_in(A.alphas + k * uint64_t(A.in_bytes), A.in_bytes);
  const blk beta = A.betas ? ld_blk(A.betas + k) : zero_blk();
  if (SCHEME == FSSB200_SCHEME_VDPF) {
    const int st = vdpf_gen_body<G, PRG>(P.keys, P.ga, pc, n, s0, s1, a, beta, out, A.cs + 4 * k, A.ocws + k, valid);
    if (valid) A.status[k] = st;
  } else if (SCHEME == FSSB200_SCHEME_DPF) {
    dpf_gen_body<G, PRG>(P.keys, P.ga, pc, n, s0, s1, a, beta, out);
  } else if (SCHEME == FSSB200_SCHEME_DCF) {
    dcf_gen_body<G, PRG>(P.keys, P.ga, pc, n, A.pred, s0, s1, a, beta, out);
  } else {
    blk ocw;
    ht_gen_body<G, PRG>(P.keys, P.ga, pc, n, s0, s1, a, beta, out, &ocw);
    if (valid) st_blk(A.ocws + k, ocw);
  }
}

template <int SCHEME, int G, int PRG, int OUT>
__global__ void __launch_bounds__(kPointThreads, 1)
gen_kernel(const __grid_constant__ KParams P, const __grid_constant__ GenArgs A) {
  SmemPlan sp = smem_plan<PRG>();
  const typename Prg<PRG>::ctx_t pc = prg_ctx_init<PRG>(sp);
  const int n = A.in_bits;
  const int ncw = (SCHEME == FSSB200_SCHEME_HALFTREE || SCHEME == FSSB200_SCHEME_VDPF) ? n : n + 1;
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  uint32_t slab = 0, seq = 0;
  if (OUT == 1) {
    for (uint32_t w = 0; w < nwarps; ++w) {  // same allocation sequence in every thread
      const uint32_t a = sp.alloc(CwTileOut::kWarpBytes, false, 512u);
      if (w == wid) slab = a;
    }
  }
  const uint64_t ntiles = (A.nkeys + 31) >> 5;
  for (uint64_t tile = uint64_t(wid) * gridDim.x + blockIdx.x; tile < ntiles; tile += uint64_t(gridDim.x) * nwarps) {  // (see point_kernel)
    const uint64_t k = tile * 32 + lane;
    const bool valid = k < A.nkeys;
    if (OUT == 1) {
      // idle lanes of a ragged tile shadow the last key; their tile rows lie outside the tensor and are clipped
      const CwTileOut out{slab, A.tmap, lane, lane * 64u, (lane >> 1) & 3u, int(tile * 32), ncw, &seq};
      gen_one<SCHEME, G, PRG>(P, pc, A, valid ? k : A.nkeys - 1, valid, out);
    } else if (valid) {
      const CwOutKeyMajor out{A.cws + k * uint64_t(ncw) * 32u};
      gen_one<SCHEME, G, PRG>(P, pc, A, k, true, out);
    }
  }
  if (OUT == 1) CwTileOut::drain(lane);
}

// ---- PRG known-answer kernel ------------------------------------------------------------------------------------------
template <int PRG, int MUL>
__global__ void __launch_bounds__(kPointThreads, 1)
prg_kernel(const __grid_constant__ KParams P, const blk *seeds, blk *out, uint64_t n) {
  SmemPlan sp = smem_plan<PRG>();
  const typename Prg<PRG>::ctx_t pc = prg_ctx_init<PRG>(sp);
  const uint64_t stride = uint64_t(gridDim.x) * blockDim.x;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    blk o[MUL];
    Prg<PRG>::template gen<MUL>(P.keys, pc, ld_blk(seeds + i), o);
#pragma unroll
    for (int j = 0; j < MUL; ++j) st_blk(out + i * MUL + j, o[j]);
  }
}

// ---- Grotto leaf bits: bit-packed in shared memory, written as coalesced bytes ------------------------------------------
// A thread's 2^dfs leaf control bits are packed into 2^(dfs-5) words (`[word][thread]`, conflict-free writes) and the
// CTA then writes the unit's 2^unit_bits output bytes together: 16 bytes per lane and store, consecutive lanes on
// consecutive 16-byte pieces (the per-thread byte stores of the first version put 32 two-byte fragments into 32
// different sectors per instruction).  The destination may have any alignment (leaf row of a heap-ordered parity
// tree, grotto_dcf.cuh:98): unaligned head / tail bytes go out as single bytes.
FSS_D uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
FSS_D void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
// bits 0..3 of b -> bytes 0..3 of the result (0 / 1 each)
FSS_D uint32_t spread4(uint32_t b) { return ((b & 15u) * 0x00204081u) & 0x01010101u; }
// word W of the unit's bit string: thread W >> (dfs-5), word W & (2^(dfs-5) - 1)
FSS_D uint32_t bits_word_addr(uint32_t s_bits, uint32_t W, int dfs) {
  const uint32_t wpt_bits = uint32_t(dfs - 5);
  return s_bits + (((W & ((1u << wpt_bits) - 1u)) * kEvalAllThreads + (W >> wpt_bits)) << 2);
}
FSS_D void write_leaf_bits(uint8_t *dst, uint32_t s_bits, int unit_bits, int dfs, uint32_t tid) {
  const uint32_t total = 1u << unit_bits, nwords = total >> 5;
  uint32_t head = uint32_t(-reinterpret_cast<uintptr_t>(dst)) & 15u;
  if (head > total) head = total;
  const uint32_t nvec = (total - head) >> 4;
  for (uint32_t v = tid; v < nvec; v += kEvalAllThreads) {
    const uint32_t p = head + (v << 4), W = p >> 5;
    const uint32_t lo = lds_u32(bits_word_addr(s_bits, W, dfs));
    const uint32_t hi = (W + 1u < nwords) ? lds_u32(bits_word_addr(s_bits, W + 1u, dfs)) : 0u;
    const uint32_t b = __funnelshift_r(lo, hi, p & 31u);
    uint4 o;
    o.x = spread4(b); o.y = spread4(b >> 4); o.z = spread4(b >> 8); o.w = spread4(b >> 12);
    *reinterpret_cast<uint4 *>(dst + p) = o;
  }
  const uint32_t tail0 = head + (nvec << 4);
  const uint32_t nscalar = head + (total - tail0);  // < 32
  if (tid < nscalar) {
    const uint32_t p = tid < head ? tid : tail0 + (tid - head);
    dst[p] = uint8_t((lds_u32(bits_word_addr(s_bits, p >> 5, dfs)) >> (p & 31u)) & 1u);
  }
}

// ---- full-domain evaluation (DPF / Half-Tree / Grotto leaf bits) -----------------------------------------------------------
// Work unit = the subtree of 2^unit_bits leaves below one node at depth du = n - unit_bits.
//   phase 0  warp 0 walks the du levels from the key's root to the unit root (1 node / level)
//   phase 1  breadth-first in shared memory: `breadth_bits` levels, one __syncthreads each
//   phase 2  every thread owns one node and expands its 2^dfs_bits leaves depth-first: an explicit
//            stack of right siblings in shared memory ([depth][thread], conflict-free 128-bit
//            accesses), ONE copy of the node-expansion code; the bottom step turns a node into two
//            adjacent leaves and writes them with a single 256-bit store.
// Every node is expanded exactly once (2 PRG blocks per DPF node, 1 per Half-Tree node).
// MODE: 0 = DPF leaves (16 B), 1 = Half-Tree leaves (16 B), 2 = Grotto leaf control bits (1 B),
//       4 = VDPF: packed (s | t) leaves, n correction words and no output CW (vdpf.cuh:345-401); the leaf
//           conversion and the proof chain run in vdpf_finish_kernel.
template <int MODE, int G, int PRG>
__global__ void __launch_bounds__(kEvalAllThreads, 1)
evalall_kernel(const __grid_constant__ KParams P, const __grid_constant__ EvalAllArgs A) {
  SmemPlan sp = smem_plan<PRG>();
  const typename Prg<PRG>::ctx_t pc = prg_ctx_init<PRG>(sp);
  const int n = A.in_bits;
  const int tid = threadIdx.x;
  const bool half = (MODE == 1);
  const int ncw = (half || MODE == 4) ? n : n + 1;
  const int dfs = A.dfs_bits, bt = A.breadth_bits;
  const int du = n - A.unit_bits;
  // scratch: per-level correction words {cwl, cwr}, two breadth buffers, the DFS stack
  const uint32_t s_cw = sp.alloc(uint32_t(ncw + 1) * 32u, true);
  const uint32_t s_trm = sp.alloc(16u, true);  // tr_cw bits of levels 0..63 (phase 2 reads cwl only, see below)
  const uint32_t s_bfs = sp.alloc(2u * kEvalAllThreads * 16u, true);
  // leaves of 16 B: cooperative bottom stage (see phase 2) when every warp is full and the thread sub-tree has >= 8 leaves
  const bool coop = MODE != 2 && bt >= 5 && dfs >= 3;
  const int nstk = coop ? dfs - 2 : dfs - 1;
  const uint32_t s_stk = sp.alloc(uint32_t(nstk > 1 ? nstk : 1) * kEvalAllThreads * 16u, false);
  // frontier tiles of the bottom stage, 2 KB per warp: warps 0..7 reuse s_bfs (dead in phase 2), warps 8..15 get s_fr1
  const uint32_t s_fr1 = coop ? sp.alloc((kEvalAllThreads / 2) * 64u, true) : 0u;  // 64 B (four nodes) per thread
  // Grotto, dfs >= 5 (n >= 14, every thread active): the unit's leaf bits, packed (write_leaf_bits)
  const bool packed = MODE == 2 && dfs >= 5;
  const uint32_t s_bits = packed ? sp.alloc((kEvalAllThreads * 4u) << (dfs - 5), true) : 0u;

  const uint64_t upk = A.leaf_count >> A.unit_bits;  // units per key
  const uint64_t total = A.nkeys * upk;
  uint64_t cur_key = ~uint64_t(0);
  for (uint64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const uint64_t key = unit / upk;
    const uint64_t leaf0 = A.leaf_begin + ((unit - key * upk) << A.unit_bits);  // first leaf of the unit
    __syncthreads();  // previous unit done with s_cw / s_bfs
    if (key != cur_key) {
      cur_key = key;
      const uint8_t *kc = A.cws + key * uint64_t(ncw) * 32u;
      for (int i = tid; i < ncw; i += kEvalAllThreads) {
        const blk cs = ld_blk(kc + 32 * i);
        blk cr = cs;  // right-child correction word: clamp bit := tr_cw (bool at byte 16)
        if (!half) cr.w = (cs.w & ~1u) | uint32_t(__ldg(kc + 32 * i + 16) != 0);
        sts_blk(s_cw + 32u * i, cs);
        sts_blk(s_cw + 32u * i + 16u, half ? ld_blk(kc + 32 * i + 16) : cr);
      }
      if (!half && tid < 64) {  // n <= 40: warps 0 and 1 cover every level
        const uint32_t f = tid < ncw ? uint32_t(__ldg(kc + 32 * tid + 16) != 0) : 0u;
        const uint32_t m = __ballot_sync(0xffffffffu, f != 0u);
        if ((tid & 31) == 0) sts_u32(s_trm + 4u * (uint32_t(tid) >> 5), m);
      }
      __syncthreads();
    }
    // ---- phase 0: root -> unit root ----
    if (tid < 32) {
      blk st = clamp(ld_blk(A.seeds + key));
      st.w |= uint32_t(A.party);
      const uint64_t path = leaf0 >> A.unit_bits;
      for (int i = 0; i < du; ++i) {
        const uint32_t bit = uint32_t(path >> (du - 1 - i)) & 1u;
        blk l, r;
        if (half) ht_expand<PRG>(P.keys, pc, st, lds_blk(s_cw + 32u * i), l, r);
        else dpf_expand<PRG>(P.keys, pc, st, lds_blk(s_cw + 32u * i), lds_blk(s_cw + 32u * i + 16u), l, r);
        st = bit ? r : l;
      }
      if (tid == 0) sts_blk(s_bfs, st);
    }
    __syncthreads();
    // ---- phase 1: breadth-first, levels du .. du+bt-1 ----
    for (int j = 0; j < bt; ++j) {
      const uint32_t src = s_bfs + uint32_t(j & 1) * (kEvalAllThreads * 16u);
      const uint32_t dst = s_bfs + uint32_t((j + 1) & 1) * (kEvalAllThreads * 16u);
      if (tid < (1 << j)) {
        const blk st = lds_blk(src + 16u * tid);
        const int lvl = du + j;
        blk l, r;
        if (half) ht_expand<PRG>(P.keys, pc, st, lds_blk(s_cw + 32u * lvl), l, r);
        else dpf_expand<PRG>(P.keys, pc, st, lds_blk(s_cw + 32u * lvl), lds_blk(s_cw + 32u * lvl + 16u), l, r);
        sts_blk(dst + 32u * tid, l);
        sts_blk(dst + 32u * tid + 16u, r);
      }
      __syncthreads();
    }
    // ---- phase 2: per-thread depth-first over dfs levels ----
    // Only TWO node-expansion sites per instantiation (internal node, bottom node): three copies of the AES pair
    // (43 KB of SASS) fall out of the 32 KB instruction cache and cost 5 % (measured, profiles/r01_evalall_coop.md).
    blk cur = {0u, 0u, 0u, 0u};
    if (tid < (1 << bt)) cur = lds_blk(s_bfs + uint32_t(bt & 1) * (kEvalAllThreads * 16u) + 16u * tid);
    if (coop) __syncthreads();  // the frontier tiles of warps 0..7 reuse the breadth buffers
    if (MODE != 2 && tid < (1 << bt)) {
      // Leaves (16 B).  coop: the depth-first walk stops one level above the bottom nodes and parks them -- four per
      // thread, i.e. one 8-leaf sub-tree -- in the warp's frontier tile; the warp then turns the 128 parked nodes into
      // leaves with the lanes re-mapped so that 4 adjacent lanes write one full 128-byte line (8 lines per STG.256
      // instead of 32; the scattered form cost 34 LSU wavefronts per store, profiles/r01_ncu_full.md).
      // !coop (tiny domains): every thread turns its own bottom node into two leaves.
      const int lvl0 = du + bt;  // tree level of `cur`
      const uint32_t lane = uint32_t(tid) & 31u;
      const uint64_t out0 = key * A.ys_stride + (leaf0 - A.leaf_begin) + (uint64_t(tid) << dfs);
      const int dint = coop ? dfs - 2 : dfs - 1;  // internal nodes live at depths < dint (+ depth dint itself when coop)
      const uint32_t fr = !coop ? 0u : (tid < kEvalAllThreads / 2 ? s_bfs + (uint32_t(tid) >> 5) * 2048u
                                                                  : s_fr1 + ((uint32_t(tid) >> 5) - kEvalAllThreads / 64) * 2048u);
      const uint32_t total = 1u << dint;  // coop: nodes at depth dfs-2 (two bottom nodes each); else bottom nodes
      // bottom-level correction words and the output correction word stay in registers for the whole unit: as
      // broadcast LDS.128 they cost 4 wavefronts each, 12 per bottom pass next to its 320 table-lookup wavefronts
      const blk bcw0 = lds_blk(s_cw + 32u * uint32_t(n - 1)), bcw1 = lds_blk(s_cw + 32u * uint32_t(n - 1) + 16u);
      const uint64_t trm = half ? 0ull : (uint64_t(lds_u32(s_trm + 4u)) << 32) | lds_u32(s_trm);
      const blk bocw = MODE == 0 ? lds_blk(s_cw + 32u * n) : (MODE == 1 ? ld_blk(A.ocws + key) : blk{0u, 0u, 0u, 0u});
      uint32_t cnt = 0;
      int d = 0;
      while (true) {
        if (d < dint || coop) {
          const int lvl = lvl0 + d;
          blk l, r;
          if (half) {
            ht_expand<PRG>(P.keys, pc, cur, lds_blk(s_cw + 32u * lvl), l, r);
          } else {
            // one broadcast LDS.128 (4 wavefronts) instead of two: cwr = cwl with tr_cw in the clamp bit
            const blk cwl = lds_blk(s_cw + 32u * lvl);
            blk cwr = cwl;
            cwr.w = (cwl.w & ~1u) | (uint32_t(trm >> lvl) & 1u);
            dpf_expand<PRG>(P.keys, pc, cur, cwl, cwr, l, r);
          }
          if (d < dint) {
            sts_blk(s_stk + (uint32_t(d) * kEvalAllThreads + tid) * 16u, r);  // slot of depth d+1
            cur = l;
            ++d;
            continue;
          }
          // coop, d == dfs-2: l, r are bottom nodes 2h, 2h+1 of the thread's current 8-leaf sub-tree -> tile slot
          // (i, lane), XOR-swizzled so that the lane-major writes and the re-mapped reads are both conflict-free
          const uint32_t h = cnt & 1u;
          sts_blk(fr + (((2u * h) << 5) + (lane ^ (4u * h))) * 16u, l);
          sts_blk(fr + (((2u * h + 1u) << 5) + (lane ^ (4u * h + 2u))) * 16u, r);
          ++cnt;
          if (h == 0u) {  // the right sibling at depth dfs-2 is on top of the stack
            cur = lds_blk(s_stk + (uint32_t(d - 1) * kEvalAllThreads + tid) * 16u);
            continue;
          }
          __syncwarp();
        }
        // bottom stage: nodes at tree level n-1 -> two leaves each
        const uint32_t passes = coop ? 4u : 1u;
#pragma unroll 1
        for (uint32_t p = 0; p < passes; ++p) {
          blk node = cur;
          uint64_t o = out0 + 2u * cnt;
          if (coop) {
            const uint32_t idx = (p << 5) + lane, owner = idx >> 2, sub = idx & 3u;
            node = lds_blk(fr + ((sub << 5) + (owner ^ (2u * sub))) * 16u);
            o = out0 + (int64_t(int32_t(owner) - int32_t(lane)) << dfs) + 8u * ((cnt >> 1) - 1u) + 2u * sub;
          }
          if (MODE == 1) {
            const blk y0 = ht_last<G, PRG>(P.keys, P.ga, pc, uint32_t(A.party), node, 0u, bcw0, bcw0.w & 1u, bocw);
            const blk y1 = ht_last<G, PRG>(P.keys, P.ga, pc, uint32_t(A.party), node, 1u, bcw0, uint32_t((bcw1.x & 0xffu) != 0), bocw);
            stg_blk2(static_cast<blk *>(A.ys) + o, y0, y1);
          } else {
            blk l, r;
            dpf_expand<PRG>(P.keys, pc, node, bcw0, bcw1, l, r);
            if (MODE == 0) {
              stg_blk2(static_cast<blk *>(A.ys) + o, dpf_leaf<G>(P.ga, uint32_t(A.party), l, bocw),
                  dpf_leaf<G>(P.ga, uint32_t(A.party), r, bocw));
            } else {
              stg_blk2(static_cast<blk *>(A.ys) + o, l, r);  // MODE 4: packed (s | t) leaves
            }
          }
        }
        if (coop) __syncwarp();  // the tile is rewritten by the next sub-tree
        else ++cnt;
        if (cnt == total) break;
        // depth of the pending right sibling: cnt counts finished nodes at depth dint
        d = dint - (__ffs(int(cnt)) - 1);
        cur = lds_blk(s_stk + (uint32_t(d - 1) * kEvalAllThreads + tid) * 16u);
      }
    }
    if (MODE == 2 && tid < (1 << bt)) {
      // Grotto: leaf control bits.  Two expansion sites on purpose: at the bottom site only the clamp bits of the two
      // children are live, so most of the last AES round is dead code there.
      const int lvl0 = du + bt;
      const uint64_t out0 = key * A.ys_stride + (leaf0 - A.leaf_begin) + (uint64_t(tid) << dfs);
      const uint32_t pairs = 1u << (dfs - 1);
      uint32_t done = 0;  // leaf pairs emitted
      uint32_t acc = 0;   // leaf bits of the current 32-leaf word
      const uint64_t trm = (uint64_t(lds_u32(s_trm + 4u)) << 32) | lds_u32(s_trm);
      // bottom-level correction words in registers (a broadcast LDS.128 is 4 wavefronts)
      const blk bcw0 = lds_blk(s_cw + 32u * uint32_t(n - 1)), bcw1 = lds_blk(s_cw + 32u * uint32_t(n - 1) + 16u);
      int d = 0;
      while (true) {
        const int lvl = lvl0 + d;
        if (d == dfs - 1) {
          blk l, r;
          dpf_expand<PRG>(P.keys, pc, cur, bcw0, bcw1, l, r);
          // one byte per leaf (grotto_dcf.cuh:190-194)
          if (packed) {
            acc |= (lsb(l) | (lsb(r) << 1)) << ((2u * done) & 31u);
            if ((done & 15u) == 15u) {
              sts_u32(s_bits + (((done >> 4) * kEvalAllThreads + uint32_t(tid)) << 2), acc);
              acc = 0;
            }
          } else {
            uint8_t *o = static_cast<uint8_t *>(A.ys) + out0 + 2 * done;  // any alignment (parity trees)
            o[0] = uint8_t(lsb(l));
            o[1] = uint8_t(lsb(r));
          }
          ++done;
          if (done == pairs) break;
          d = dfs - 1 - __ffs(int(done)) + 1;  // depth of the pending right sibling: dfs-1 - ctz(done)
          cur = lds_blk(s_stk + (uint32_t(d - 1) * kEvalAllThreads + tid) * 16u);
        } else {
          blk l, r;
          const blk cwl = lds_blk(s_cw + 32u * lvl);
          blk cwr = cwl;
          cwr.w = (cwl.w & ~1u) | (uint32_t(trm >> lvl) & 1u);
          dpf_expand<PRG>(P.keys, pc, cur, cwl, cwr, l, r);
          sts_blk(s_stk + (uint32_t(d) * kEvalAllThreads + tid) * 16u, r);  // slot of depth d+1
          cur = l;
          ++d;
        }
      }
    }
    if (packed) {
      __syncthreads();
      write_leaf_bits(static_cast<uint8_t *>(A.ys) + key * A.ys_stride + (leaf0 - A.leaf_begin), s_bits, A.unit_bits, dfs,
          uint32_t(tid));
    }
  }
}

// ---- DCF full-domain evaluation (Dcf::EvalAll, dcf.cuh:294-385; the reference has no GPU version) ---------------------
// Same three phases as evalall_kernel; a node carries its running value share (32 bytes per stack / frontier entry),
// four AES blocks per node with fixed keys 0..3.  Two geometries (TB = log2 threads per CTA):
//   TB = 8  256 threads, up to 7 stack levels (56 KB), 8 warps per SM, work unit 2^16 leaves: the shipped default
//           (0.86-0.88 of the LDS ceiling);
//   TB = 9  512 threads, 16 warps per SM (FSSB200_DCF_ALL_THREADS=512).  99 KB of scratch beside the 128 KB of tables hold
//           five 16 KB levels: the two breadth buffers are dead once every thread has taken its start node and become
//           stack levels 0 and 1, three more levels live below the tables => dfs <= 6, work unit 2^15 leaves.  Built to
//           test the round-1 guess that 8 warps cannot hide the 4-block latency: they can -- 16 warps are no faster
//           (profiles/r02_dcf_evalall_ab.md).
constexpr int kDcfAllThreadBitsSmall = 8;
constexpr int kDcfAllThreadBits = 9;
constexpr int kDcfAllMaxDfsBits = 6;   // TB = 9: five stack levels

// group values travel through shared memory in their Into() form
template <int G>
FSS_D typename Grp<G>::V smem_load_val(const GroupArgs &ga, uint32_t addr) {
  return Grp<G>::from(ga, lds_blk(addr));
}
template <int G>
FSS_D void smem_store_val(const GroupArgs &ga, uint32_t addr, typename Grp<G>::V v) {
  sts_blk(addr, Grp<G>::into(ga, v));
}

template <int G, int PRG, int TB>
__global__ void __launch_bounds__(1 << TB, 1)
dcf_evalall_kernel(const __grid_constant__ KParams P, const __grid_constant__ EvalAllArgs A) {
  typedef Grp<G> GR;
  typedef typename GR::V V;
  constexpr uint32_t T = 1u << TB;
  constexpr bool kReuseBfs = TB == 9;
  SmemPlan sp = smem_plan<PRG>();
  const typename Prg<PRG>::ctx_t pc = prg_ctx_init<PRG>(sp);
  const int n = A.in_bits, ncw = n + 1;
  const int tid = threadIdx.x;
  const int dfs = A.dfs_bits, bt = A.breadth_bits;
  const int du = n - A.unit_bits;
  const uint32_t s_cw = sp.alloc(uint32_t(ncw) * 48u, true);            // {cwl, cwr, Into(vcw)} per level
  const uint32_t s_bfs = sp.alloc(2u * T * 32u, true);                    // {node, value} x 2 buffers
  // stack level d (right sibling parked at depth d+1): levels [0, kReuse) are the breadth buffers, the rest one block
  constexpr int kReuse = kReuseBfs ? 2 : 0;
  const int nlev = dfs > 1 ? dfs - 1 : 1;
  const uint32_t s_more = nlev > kReuse ? sp.alloc(uint32_t(nlev - kReuse) * T * 32u, false) : 0u;
  auto level_addr = [&](int d) -> uint32_t {
    return (d < kReuse ? s_bfs + uint32_t(d) * (T * 32u) : s_more + uint32_t(d - kReuse) * (T * 32u)) + uint32_t(tid) * 32u;
  };

  const uint64_t upk = A.leaf_count >> A.unit_bits;
  const uint64_t total = A.nkeys * upk;
  uint64_t cur_key = ~uint64_t(0);
  for (uint64_t unit = blockIdx.x; unit < total; unit += gridDim.x) {
    const uint64_t key = unit / upk;
    const uint64_t leaf0 = A.leaf_begin + ((unit - key * upk) << A.unit_bits);
    __syncthreads();
    if (key != cur_key) {
      cur_key = key;
      const uint8_t *kc = A.cws + key * uint64_t(ncw) * 32u;
      for (int i = tid; i < ncw; i += int(T)) {
        const blk cs = ld_blk(kc + 32 * i), cv = ld_blk(kc + 32 * i + 16);
        blk cr = cs;
        cr.w = (cs.w & ~1u) | (cv.w & 1u);                                  // tr_cw = lsb(cw.v), dcf.cuh:217-219
        sts_blk(s_cw + 48u * i, cs);
        sts_blk(s_cw + 48u * i + 16u, cr);
        sts_blk(s_cw + 48u * i + 32u, GR::into(P.ga, GR::from(P.ga, clamp(cv))));
      }
      __syncthreads();
    }
#define FSS_DCF_EXPAND(st, u, lvl, l, r, ul, ur)                                                          \
  dcf_expand<G, PRG>(P.keys, P.ga, pc, st, u, lds_blk(s_cw + 48u * (lvl)), lds_blk(s_cw + 48u * (lvl) + 16u), \
      smem_load_val<G>(P.ga, s_cw + 48u * (lvl) + 32u), l, r, ul, ur)
    if (tid < 32) {  // phase 0: root -> unit root
      blk st = clamp(ld_blk(A.seeds + key));
      st.w |= uint32_t(A.party);
      V u = GR::zero(P.ga);
      const uint64_t path = leaf0 >> A.unit_bits;
      for (int i = 0; i < du; ++i) {
        blk l, r;
        V ul, ur;
        FSS_DCF_EXPAND(st, u, i, l, r, ul, ur);
        const bool bit = (path >> (du - 1 - i)) & 1;
        st = bit ? r : l;
        u = bit ? ur : ul;
      }
      if (tid == 0) {
        sts_blk(s_bfs, st);
        smem_store_val<G>(P.ga, s_bfs + 16u, u);
      }
    }
    __syncthreads();
    for (int j = 0; j < bt; ++j) {  // phase 1: breadth-first
      const uint32_t src = s_bfs + uint32_t(j & 1) * (T * 32u), dst = s_bfs + uint32_t((j + 1) & 1) * (T * 32u);
      if (tid < (1 << j)) {
        const blk st = lds_blk(src + 32u * tid);
        const V u = smem_load_val<G>(P.ga, src + 32u * tid + 16u);
        blk l, r;
        V ul, ur;
        FSS_DCF_EXPAND(st, u, du + j, l, r, ul, ur);
        sts_blk(dst + 64u * tid, l);
        smem_store_val<G>(P.ga, dst + 64u * tid + 16u, ul);
        sts_blk(dst + 64u * tid + 32u, r);
        smem_store_val<G>(P.ga, dst + 64u * tid + 48u, ur);
      }
      __syncthreads();
    }
    // phase 2: depth-first per thread
    const uint32_t slot = s_bfs + uint32_t(bt & 1) * (T * 32u) + 32u * tid;
    blk cur = {0u, 0u, 0u, 0u};
    V u = GR::zero(P.ga);
    if (tid < (1 << bt)) {
      cur = lds_blk(slot);
      u = smem_load_val<G>(P.ga, slot + 16u);
    }
    if (kReuseBfs) __syncthreads();  // every start node is in registers: the breadth buffers become stack levels 0 and 1
    if (tid < (1 << bt)) {
      const int lvl0 = du + bt;
      const uint64_t out0 = key * A.ys_stride + (leaf0 - A.leaf_begin) + (uint64_t(tid) << dfs);
      const uint32_t pairs = 1u << (dfs - 1);
      const blk out_v = ld_blk(A.cws + key * uint64_t(ncw) * 32u + 32u * n + 16u);  // cws[n].v
      uint32_t done = 0;
      int d = 0;
      while (true) {
        blk l, r;
        V ul, ur;
        FSS_DCF_EXPAND(cur, u, lvl0 + d, l, r, ul, ur);
        if (d == dfs - 1) {
          stg_blk2(static_cast<blk *>(A.ys) + out0 + 2 * done, dcf_leaf<G>(P.ga, uint32_t(A.party), l, ul, out_v),
              dcf_leaf<G>(P.ga, uint32_t(A.party), r, ur, out_v));
          ++done;
          if (done == pairs) break;
          d = dfs - __ffs(int(done));
          const uint32_t e = level_addr(d - 1);
          cur = lds_blk(e);
          u = smem_load_val<G>(P.ga, e + 16u);
        } else {
          const uint32_t e = level_addr(d);
          sts_blk(e, r);
          smem_store_val<G>(P.ga, e + 16u, ur);
          cur = l;
          u = ul;
          ++d;
        }
      }
    }
#undef FSS_DCF_EXPAND
  }
}

}  // namespace fssb200
