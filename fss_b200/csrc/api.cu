// SPDX-License-Identifier: Apache-2.0
//
// api.cu -- the extern "C" boundary of include/fssb200.h, device-pointer entry points: context, argument
// validation, launch geometry.  There is no CPU evaluation path in this library: every entry point either
// launches sm_100a kernels or returns an error code.  The host-buffer entry points live in host_api.cu, the
// multi-device ones in multi_api.cu.
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no libcuda link dependency)
#include <cstdlib>
#include <cstring>
#include <new>

#include "ctx.h"
#include "misc_kernels.cuh"
#include "vdpf_kernels.cuh"

using namespace fssb200;


namespace {

#define CUDA_TRY(expr) FSS_CUDA_TRY(expr)

// Tensor map of the key-major Cw array for the TMA-fed point kernels (CwTile in kernels.cuh).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_rows_tensor_map(uint8_t out[128], const void *rows, size_t nkeys, size_t row_bytes);
// cuTensorMapEncodeTiled, resolved once through the runtime (no link-time dependency on libcuda)
static int encode_tiled_fn(EncodeTiledFn *out) {
  static std::atomic<EncodeTiledFn> fn{nullptr};
  EncodeTiledFn f = fn.load(std::memory_order_acquire);
  if (!f) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return int(e);
    if (q != cudaDriverEntryPointSuccess || !p) return int(cudaErrorNotSupported);
    f = reinterpret_cast<EncodeTiledFn>(p);
    fn.store(f, std::memory_order_release);
  }
  *out = f;
  return 0;
}
int make_cw_tensor_map(uint8_t out[128], const void *cws, size_t nkeys, int ncw) {
  return make_rows_tensor_map(out, cws, nkeys, size_t(ncw) * 32u);
}
// 2-D byte tensor [nkeys][row_bytes], box 32 rows x 64 B, SWIZZLE_64B (CwTileT / CwTileOut in kernels.cuh)
int make_rows_tensor_map(uint8_t out[128], const void *cws, size_t nkeys, size_t row_bytes) {
  EncodeTiledFn fn = nullptr;
  if (int rc = encode_tiled_fn(&fn)) return rc;
  static_assert(sizeof(CUtensorMap) == 128, "PointArgs::tmap size");
  alignas(64) CUtensorMap m;
  const cuuint64_t gdim[2] = {cuuint64_t(row_bytes), cuuint64_t(nkeys)};
  const cuuint64_t gstride[1] = {cuuint64_t(row_bytes)};
  const cuuint32_t box[2] = {64u, 32u};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void *>(cws), gdim, gstride, box, estride,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return int(cudaErrorInvalidValue);
  std::memcpy(out, &m, 128);
  return 0;
}

// Level-major array [nlevels][nkeys] of 16-byte entries as a 2-D uint32 tensor, box = 32 keys x box_levels (CwLmTile)
int make_lm_tensor_map(uint8_t out[128], const void *base, size_t nkeys, int nlevels, int box_levels) {
  EncodeTiledFn fn = nullptr;
  if (int rc = encode_tiled_fn(&fn)) return rc;
  alignas(64) CUtensorMap m;
  const cuuint64_t gdim[2] = {cuuint64_t(nkeys) * 4u, cuuint64_t(nlevels)};
  const cuuint64_t gstride[1] = {cuuint64_t(nkeys) * 16u};
  const cuuint32_t box[2] = {128u, cuuint32_t(box_levels)};
  const cuuint32_t estride[2] = {1u, 1u};
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void *>(base), gdim, gstride, box, estride,
      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return int(cudaErrorInvalidValue);
  std::memcpy(out, &m, 128);
  return 0;
}

int group_kind(const fssb200_params &p, uint32_t *vmask) {
  const bool has_mod = (p.mod_lo | p.mod_hi) != 0;
  *vmask = 0xffffffffu;
  switch (p.group) {
    case FSSB200_GROUP_BYTES: return has_mod ? -1 : kGrpBytes;
    case FSSB200_GROUP_U8:
    case FSSB200_GROUP_U16:
    case FSSB200_GROUP_U32: {
      const int bits = p.group == FSSB200_GROUP_U8 ? 8 : (p.group == FSSB200_GROUP_U16 ? 16 : 32);
      *vmask = bits == 32 ? 0xffffffffu : ((1u << bits) - 1u);
      if (!has_mod) return kGrpU32;
      if (p.mod_hi || (bits < 32 && (p.mod_lo >> bits)) || (p.mod_lo >> 32)) return -1;
      return kGrpU32Mod;
    }
    case FSSB200_GROUP_U64:
      if (!has_mod) return kGrpU64;
      return p.mod_hi ? -1 : kGrpU64Mod;
    case FSSB200_GROUP_U128:
      if (!has_mod) return -1;                                       // uint.cuh:29: mod > 0 required
      if (p.mod_hi == 0x8000000000000000ull && p.mod_lo == 0) return kGrpU127;
      if (p.mod_hi >> 63) return -1;                                 // mod > 2^127
      return kGrpU128Mod;
  }
  return -1;
}

// Launch geometry of the point / gen / prg kernels: AES = one persistent 512-thread CTA per SM with
// the full dynamic shared memory (tables); ChaCha = plain 256-thread CTAs.
// `mode` < 0: gen / prg kernels (no correction-word staging).
LaunchCfg point_cfg(const fssb200_ctx *c, uint64_t n, cudaStream_t s, int mode = -1) {
  LaunchCfg cfg;
  cfg.stream = s;
  if (c->p.prg == FSSB200_PRG_AES128_MMO) {
    const unsigned threads = mode == 1 ? 1024u : (mode >= 5 ? 768u : unsigned(kPointThreads));
    const uint64_t want = (n + 31) / 32;  // one CTA per SM as soon as there is a tile of 32 keys for each (tiles interleave over CTAs)
    cfg.grid = dim3(unsigned(want < uint64_t(c->sm_count) ? (want ? want : 1) : c->sm_count));
    cfg.block = dim3(threads);
    cfg.smem = kMaxDynSmem;
  } else {
    const uint64_t want = (n + 31) / 32;  // (tiles interleave over the CTAs: spread a small batch over every SM)
    const uint64_t cap = uint64_t(c->sm_count) * 6;
    cfg.grid = dim3(unsigned(want < cap ? (want ? want : 1) : cap));
    cfg.block = dim3(256);
    // staged correction words: one slab per warp (L = 4, or L = 2 in mode 1)
    cfg.smem = (mode == 0 || mode == 1) ? 8 * (mode == 1 ? CwStagedWarp<2>::kWarpBytes : CwStagedWarp<4>::kWarpBytes) + 32
        : (mode >= 4) ? 8 * CwTile::kWarpBytes + 8 * 16 + 512 + 32 : 0;
  }
  return cfg;
}

// Gen kernels: geometry of point_cfg; out_mode 1 adds one CwTileOut pair of tiles per warp to the ChaCha CTAs
// (the AES CTAs always own the full dynamic shared memory).
LaunchCfg gen_cfg(const fssb200_ctx *c, uint64_t n, cudaStream_t s, int out_mode) {
  LaunchCfg cfg = point_cfg(c, n, s);
  if (c->p.prg != FSSB200_PRG_AES128_MMO && out_mode) cfg.smem = 8 * CwTileOut::kWarpBytes + 512 + 32;
  return cfg;
}
// Fills a.tmap and returns the output mode the launch will use (TMA coordinates are 32-bit).
int gen_out_mode(const fssb200_ctx *c, GenArgs &a, int *rc) {
  *rc = 0;
  if (!c->gen_mode || (a.nkeys >> 31)) return 0;
  *rc = make_cw_tensor_map(a.tmap, a.cws, a.nkeys, c->ncw);
  return 1;
}

struct EvalAllPlan {
  int unit_bits, breadth_bits, dfs_bits;
};
// thread_bits: log2(threads per CTA) of the kernel (9; 8 for the DCF kernel whose nodes carry a value)
EvalAllPlan plan_evalall(int n, int thread_bits = kEvalAllThreadBits, int max_dfs = kMaxDfsBits) {
  EvalAllPlan pl;
  int dfs = n - thread_bits;
  if (dfs < 1) dfs = 1;
  if (dfs > max_dfs) dfs = max_dfs;
  int bt = n - dfs;
  if (bt > thread_bits) bt = thread_bits;
  pl.dfs_bits = dfs;
  pl.breadth_bits = bt;
  pl.unit_bits = bt + dfs;
  return pl;
}

int check_common(const fssb200_ctx *c) { return c ? 0 : FSSB200_EINVAL; }

}  // namespace

extern "C" {

int fssb200_version(void) { return FSSB200_VERSION; }

const char *fssb200_strerror(int code) {
  switch (code) {
    case FSSB200_OK: return "ok";
    case FSSB200_EINVAL: return "invalid argument";
    case FSSB200_EDOMAIN: return "in_bits outside the supported domain";
    case FSSB200_EGROUP: return "unsupported group / modulus";
    case FSSB200_ESCHEME: return "entry point does not apply to this scheme";
    case FSSB200_EALIGN: return "pointer is not 16-byte aligned";
    case FSSB200_ENODEVICE: return "no such CUDA device";
    case FSSB200_ERANGE: return "leaf range is not unit-aligned or outside the domain";
    case FSSB200_ENOARENA: return "host staging arena not reserved";
  }
  if (code > 0) return cudaGetErrorString(static_cast<cudaError_t>(code));
  return "unknown error";
}

int fssb200_ctx_create(const fssb200_params *p, fssb200_ctx **out) {
  if (!p || !out) return FSSB200_EINVAL;
  *out = nullptr;
  if (p->scheme < FSSB200_SCHEME_DPF || p->scheme > FSSB200_SCHEME_VDPF) return FSSB200_EINVAL;
  if (p->prg != FSSB200_PRG_AES128_MMO && p->prg != FSSB200_PRG_CHACHA) return FSSB200_EINVAL;
  if (p->pred != FSSB200_PRED_LT && p->pred != FSSB200_PRED_GT) return FSSB200_EINVAL;
  if ((p->hash & ~0x0101) != 0) return FSSB200_EINVAL;  // byte 0 / byte 1: FSSB200_HASH_BLAKE3 or FSSB200_HASH_SHA256
  if (p->in_bytes != 1 && p->in_bytes != 2 && p->in_bytes != 4 && p->in_bytes != 8 && p->in_bytes != 16)
    return FSSB200_EINVAL;
  if (p->in_bits < 1 || p->in_bits > 8 * p->in_bytes) return FSSB200_EDOMAIN;
  fssb200_params q = *p;
  if (q.scheme == FSSB200_SCHEME_GROTTO) {  // GrottoDcf is defined over group::Bytes (grotto_dcf.cuh:46)
    q.group = FSSB200_GROUP_BYTES;
    q.mod_lo = q.mod_hi = 0;
  }
  uint32_t vmask;
  const int gk = group_kind(q, &vmask);
  if (gk < 0 || !grp_kind_instantiated(gk)) return FSSB200_EGROUP;

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return FSSB200_ENODEVICE;
  if (q.device < 0 || q.device >= ndev) return FSSB200_ENODEVICE;
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, q.device));
  if (prop.major < 10) return FSSB200_ENODEVICE;  // sm_100a only
  if (size_t(prop.sharedMemPerBlockOptin) < kMaxDynSmem) return FSSB200_ENODEVICE;

  fssb200_ctx *c = new (std::nothrow) fssb200_ctx();
  if (!c) return FSSB200_EINVAL;
  c->p = q;
  c->gk = gk;
  c->vmask = vmask;
  c->mul = q.scheme == FSSB200_SCHEME_DCF ? 4 : (q.scheme == FSSB200_SCHEME_HALFTREE ? 1 : 2);
  c->ncw = (q.scheme == FSSB200_SCHEME_HALFTREE || q.scheme == FSSB200_SCHEME_VDPF) ? q.in_bits : q.in_bits + 1;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = int(prop.sharedMemPerBlockOptin);
  // measured on B200 (profiles/r01_point_modes.md): correction words fetched by the TMA unit (CwTile) beat the
  // cp.async slabs for every scheme; DPF / DCF gain another 1-3 % from 24 warps per SM -- and so does Half-Tree since
  // the tiles interleave over the CTAs (2^20 keys 0.923 -> 0.926, 2^22 keys 0.939 -> 0.954: profiles/r02_levelmajor_tma.md;
  // before, the 768-thread geometry lost a whole round of tiles at 2^20 keys).
  // (the Grotto walk expands both children per level like the gen kernels: 512 threads, no register cap of 85)
  c->point_mode = q.scheme == FSSB200_SCHEME_GROTTO ? 4 : 5;
  if (const char *e = std::getenv("FSSB200_POINT_MODE")) {  // A/B measurement knob
    const int m = std::atoi(e);
    const bool grotto = q.scheme == FSSB200_SCHEME_GROTTO;  // walk: instantiated for modes 3 / 4 / 5 only
    if ((!grotto && (m == 0 || m == 1)) || m == 3 || m == 4 || m == 5) c->point_mode = m;
  }
  c->gen_mode = 1;
  if (const char *e = std::getenv("FSSB200_GEN_MODE")) c->gen_mode = std::atoi(e) ? 1 : 0;  // A/B measurement knob
  std::memset(&c->kp, 0, sizeof(c->kp));
  if (q.prg == FSSB200_PRG_AES128_MMO) {
    for (int i = 0; i < 4; ++i) aes128_expand_le(q.prg_key + 16 * i, c->kp.keys.rk[i]);
    // per-thread child selection: child bit picks key p (bit 0) or p + mul/2 (bit 1)
    const int nb = c->mul >= 2 ? c->mul / 2 : 1;
    for (int pidx = 0; pidx < 2 && pidx < nb; ++pidx)
      for (int i = 0; i < 44; ++i) c->kp.keys.rkd[pidx][i] = c->kp.keys.rk[pidx][i] ^ c->kp.keys.rk[pidx + nb][i];
  } else {
    std::memcpy(c->kp.keys.nonce, q.prg_key, 8);
  }
  std::memcpy(c->kp.keys.hash_key, q.hash_key, 16);
  std::memcpy(c->kp.keys.hash_iv, q.hash_iv, 64);
  c->kp.keys.hash_kind[0] = uint32_t(q.hash) & 0xffu;
  c->kp.keys.hash_kind[1] = (uint32_t(q.hash) >> 8) & 0xffu;
  c->kp.ga.vmask = vmask;
  c->kp.ga.mod[0] = uint32_t(q.mod_lo);
  c->kp.ga.mod[1] = uint32_t(q.mod_lo >> 32);
  c->kp.ga.mod[2] = uint32_t(q.mod_hi);
  c->kp.ga.mod[3] = uint32_t(q.mod_hi >> 32);
  *out = c;
  return 0;
}

void fssb200_ctx_destroy(fssb200_ctx *c) { delete c; }

int fssb200_ctx_params(const fssb200_ctx *c, fssb200_params *out) {
  if (!c || !out) return FSSB200_EINVAL;
  *out = c->p;
  return 0;
}

int fssb200_ctx_ncw(const fssb200_ctx *c) { return c ? c->ncw : FSSB200_EINVAL; }

uint64_t fssb200_ctx_launch_count(const fssb200_ctx *c) { return c ? c->launches.load() : 0; }

// ---- gen ------------------------------------------------------------------------------------------------------
int fssb200_gen(const fssb200_ctx *cc, const void *s0s, const void *alphas, const void *betas, void *cws,
    void *ocws, size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (!s0s || !alphas || !cws) return FSSB200_EINVAL;
  const int scheme = c->p.scheme;
  if (scheme == FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;  // fssb200_vdpf_gen
  if (scheme == FSSB200_SCHEME_HALFTREE && !ocws) return FSSB200_EINVAL;
  if (scheme != FSSB200_SCHEME_GROTTO && !betas) return FSSB200_EINVAL;
  if (!aligned16(s0s) || !aligned16(cws) || !aligned16(betas) || !aligned16(ocws)) return FSSB200_EALIGN;
  if (reinterpret_cast<uintptr_t>(alphas) % c->p.in_bytes) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  // Grotto keys are DPF keys over Bytes with beta = 0 (grotto_dcf.cuh:63-67)
  const int kscheme = scheme == FSSB200_SCHEME_GROTTO ? FSSB200_SCHEME_DPF : scheme;
  if (!grp_kind_instantiated(c->gk)) return FSSB200_EGROUP;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  GenArgs a;
  std::memset(&a, 0, sizeof(a));
  a.s0s = static_cast<const blk *>(s0s);
  a.alphas = static_cast<const uint8_t *>(alphas);
  a.betas = scheme == FSSB200_SCHEME_GROTTO ? nullptr : static_cast<const blk *>(betas);
  a.cws = static_cast<uint8_t *>(cws);
  a.ocws = static_cast<blk *>(ocws);
  a.nkeys = nkeys;
  a.in_bits = c->p.in_bits;
  a.in_bytes = c->p.in_bytes;
  a.pred = c->p.pred;
  a.vmask = c->vmask;
  int rc = 0;
  const int out_mode = gen_out_mode(c, a, &rc);
  if (rc) return rc;
  gen_launch_fn fn = get_gen_launcher(kscheme, c->gk, c->p.prg, out_mode);
  if (!fn) return FSSB200_EGROUP;
  const LaunchCfg cfg = gen_cfg(c, nkeys, static_cast<cudaStream_t>(stream), out_mode);
  c->launches++;
  return int(fn(c->kp, a, cfg));
}

// ---- point eval ---------------------------------------------------------------------------------------------------
static int eval_impl(const fssb200_ctx *cc, int want_scheme, int party, const void *seeds, const void *cws,
    const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream, bool level_major, const void *cw_s,
    const void *cw_v, const void *extra, const void *out_cw, const void *vdpf_cs = nullptr, void *vdpf_pis = nullptr,
    bool packed = false) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  const int scheme = c->p.scheme;
  // Grotto: EvalAll / Preprocess+Eval, or the O(n) walk through fssb200_grotto_eval_walk only
  if (scheme == FSSB200_SCHEME_GROTTO && want_scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (scheme == FSSB200_SCHEME_GROTTO && level_major) return FSSB200_ESCHEME;
  if (packed && scheme != FSSB200_SCHEME_DPF && scheme != FSSB200_SCHEME_HALFTREE) return FSSB200_ESCHEME;
  if (want_scheme >= 0 && want_scheme != scheme) return FSSB200_ESCHEME;
  if (scheme == FSSB200_SCHEME_VDPF) {  // only through fssb200_vdpf_eval
    if (want_scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
    if (!vdpf_cs || !vdpf_pis || !ocws) return FSSB200_EINVAL;
    if (!aligned16(vdpf_cs) || !aligned16(vdpf_pis)) return FSSB200_EALIGN;
  }
  if (party != 0 && party != 1) return FSSB200_EINVAL;
  if (!seeds || !xs || !ys) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_HALFTREE && !ocws) return FSSB200_EINVAL;
  if (level_major) {
    if (!cw_s) return FSSB200_EINVAL;
    if (scheme == FSSB200_SCHEME_DCF && (!cw_v || !out_cw)) return FSSB200_EINVAL;
    if (scheme == FSSB200_SCHEME_DPF && (!extra || !out_cw)) return FSSB200_EINVAL;
    if ((scheme == FSSB200_SCHEME_HALFTREE || scheme == FSSB200_SCHEME_VDPF) && !extra) return FSSB200_EINVAL;
    if (!aligned16(cw_s) || !aligned16(cw_v) || !aligned16(out_cw)) return FSSB200_EALIGN;
  } else {
    if (!cws) return FSSB200_EINVAL;
    if (!aligned16(cws)) return FSSB200_EALIGN;
  }
  if (!aligned16(seeds) || !aligned16(ocws)) return FSSB200_EALIGN;
  if (scheme != FSSB200_SCHEME_GROTTO && !aligned16(ys)) return FSSB200_EALIGN;  // Grotto: bool ys[nkeys]
  if (reinterpret_cast<uintptr_t>(xs) % c->p.in_bytes) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  // level-major: the TMA tiles of mode 7 (profiles/r02_levelmajor_tma.md) unless the batch does not fit 32-bit tensor
  // coordinates (key * 4) or FSSB200_LM_MODE=2 asks for the direct loads of mode 2 (A/B runs)
  static const int lm_mode = [] {
    const char *e = std::getenv("FSSB200_LM_MODE");
    return (e && std::atoi(e) == 2) ? 2 : 7;
  }();
  const int mode = level_major ? ((nkeys >> 28) ? 2 : lm_mode) : (packed ? 6 : c->point_mode);
  point_launch_fn fn = get_point_launcher(scheme, c->gk, c->p.prg, mode);
  if (!fn) return FSSB200_EGROUP;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  PointArgs a;
  std::memset(&a, 0, sizeof(a));
  a.seeds = static_cast<const blk *>(seeds);
  a.cws = static_cast<const uint8_t *>(cws);
  a.ocws = static_cast<const blk *>(ocws);
  a.xs = static_cast<const uint8_t *>(xs);
  a.ys = static_cast<blk *>(ys);
  a.cw_s = static_cast<const blk *>(cw_s);
  a.cw_v = static_cast<const blk *>(cw_v);
  a.extra = static_cast<const uint32_t *>(extra);
  a.out_cw = static_cast<const blk *>(out_cw);
  a.cs = static_cast<const blk *>(vdpf_cs);
  a.pis = static_cast<blk *>(vdpf_pis);
  a.nkeys = nkeys;
  a.in_bits = c->p.in_bits;
  a.in_bytes = c->p.in_bytes;
  a.party = party;
  a.vmask = c->vmask;
  if (mode == 7) {
    const bool dcf = scheme == FSSB200_SCHEME_DCF;
    if (int rc = make_lm_tensor_map(a.tmap, cw_s, nkeys, c->p.in_bits, dcf ? 2 : 4)) return rc;
    if (dcf)
      if (int rc = make_lm_tensor_map(a.tmap2, cw_v, nkeys, c->p.in_bits, 2)) return rc;
  } else if (mode >= 4) {
    if (nkeys >> 31) return FSSB200_EINVAL;  // TMA coordinates are 32-bit
    if (int rc = make_rows_tensor_map(a.tmap, cws, nkeys, mode == 6 ? size_t(c->ncw) * 16u + 16u : size_t(c->ncw) * 32u))
      return rc;
  }
  const LaunchCfg cfg = point_cfg(c, nkeys, static_cast<cudaStream_t>(stream), mode);
  c->launches++;
  return int(fn(c->kp, a, cfg));
}

int fssb200_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    const void *xs, void *ys, size_t nkeys, void *stream) {
  return eval_impl(c, -1, party, seeds, cws, ocws, xs, ys, nkeys, stream, false, nullptr, nullptr, nullptr, nullptr);
}
int fssb200_dpf_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *xs, void *ys,
    size_t nkeys, void *stream) {
  return eval_impl(c, FSSB200_SCHEME_DPF, party, seeds, cws, nullptr, xs, ys, nkeys, stream, false, nullptr, nullptr,
      nullptr, nullptr);
}
int fssb200_dcf_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *xs, void *ys,
    size_t nkeys, void *stream) {
  return eval_impl(c, FSSB200_SCHEME_DCF, party, seeds, cws, nullptr, xs, ys, nkeys, stream, false, nullptr, nullptr,
      nullptr, nullptr);
}
int fssb200_halftree_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    const void *xs, void *ys, size_t nkeys, void *stream) {
  return eval_impl(c, FSSB200_SCHEME_HALFTREE, party, seeds, cws, ocws, xs, ys, nkeys, stream, false, nullptr,
      nullptr, nullptr, nullptr);
}
int fssb200_eval_levelmajor(const fssb200_ctx *c, int party, const void *seeds, const void *cw_s, const void *cw_v,
    const void *extra, const void *out_cw, const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream) {
  return eval_impl(c, -1, party, seeds, nullptr, ocws, xs, ys, nkeys, stream, true, cw_s, cw_v, extra, out_cw);
}

// ---- packed rows (compact key format of the schemes whose Cw is {int4 s; bool flag} + 15 bytes of padding) ---------
size_t fssb200_packed_row_bytes(const fssb200_ctx *c) {
  if (!c || (c->p.scheme != FSSB200_SCHEME_DPF && c->p.scheme != FSSB200_SCHEME_HALFTREE)) return 0;
  return size_t(c->ncw) * 16u + 16u;
}


int fssb200_eval_packed(const fssb200_ctx *c, int party, const void *seeds, const void *rows, const void *ocws,
    const void *xs, void *ys, size_t nkeys, void *stream) {
  return eval_impl(c, -1, party, seeds, rows, ocws, xs, ys, nkeys, stream, false, nullptr, nullptr, nullptr, nullptr,
      nullptr, nullptr, true);
}

int fssb200_relayout(const fssb200_ctx *cc, const void *cws, void *cw_s, void *cw_v, void *extra, void *out_cw,
    size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  const int scheme = c->p.scheme;
  if (!cws || !cw_s) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_DCF ? (!cw_v || !out_cw) : !extra) return FSSB200_EINVAL;
  if ((scheme == FSSB200_SCHEME_DPF || scheme == FSSB200_SCHEME_GROTTO) && !out_cw) return FSSB200_EINVAL;
  if (!aligned16(cws) || !aligned16(cw_s) || !aligned16(cw_v) || !aligned16(out_cw)) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  c->launches++;
  return int(launch_relayout(scheme, c->p.in_bits, c->ncw, static_cast<const uint8_t *>(cws),
      static_cast<blk *>(cw_s), static_cast<blk *>(cw_v), static_cast<uint32_t *>(extra), static_cast<blk *>(out_cw),
      nkeys, static_cast<cudaStream_t>(stream)));
}

// ---- full-domain evaluation ---------------------------------------------------------------------------------------
// DCF full-domain kernel: 256 threads / dfs <= 8 (default), or 512 threads / dfs <= 6 with FSSB200_DCF_ALL_THREADS=512.
// Measured on a B200 (profiles/r02_dcf_evalall_ab.md): 16 warps per SM do NOT help -- u127 0.857 vs 0.875, Bytes 0.860 vs
// 0.861 of the LDS ceiling -- so the kernel is not bound by the latency 8 warps can hide; the larger work unit stays.
static int dcf_all_thread_bits() {
  static const int tb = [] {
    const char *e = std::getenv("FSSB200_DCF_ALL_THREADS");
    return (e && std::atoi(e) == 512) ? kDcfAllThreadBits : kDcfAllThreadBitsSmall;
  }();
  return tb;
}
static EvalAllPlan plan_for(const fssb200_ctx *c, bool dcf) {
  if (!dcf) return plan_evalall(c->p.in_bits);
  const int tb = dcf_all_thread_bits();
  return plan_evalall(c->p.in_bits, tb, tb == kDcfAllThreadBits ? kDcfAllMaxDfsBits : kMaxDfsBits);
}
uint64_t fssb200_eval_all_granule(const fssb200_ctx *c) {
  if (!c) return 0;
  return uint64_t(1) << plan_for(c, c->p.scheme == FSSB200_SCHEME_DCF).unit_bits;
}

static int evalall_impl(fssb200_ctx *c, int mode, int party, const void *seeds, const void *cws, const void *ocws,
    void *ys, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count, void *stream, uint64_t ys_stride = 0) {
  if (party != 0 && party != 1) return FSSB200_EINVAL;
  if (!seeds || !cws || !ys) return FSSB200_EINVAL;
  if (mode == 1 && !ocws) return FSSB200_EINVAL;
  if (!aligned16(seeds) || !aligned16(cws) || !aligned16(ocws)) return FSSB200_EALIGN;
  if (mode != 2 && !aligned16(ys)) return FSSB200_EALIGN;  // Grotto bytes: any alignment
  const int n = c->p.in_bits;
  if (n > 40) return FSSB200_EDOMAIN;  // 2^40 leaves = 16 TiB per key
  const uint64_t N = uint64_t(1) << n;
  if (leaf_begin >= N) return FSSB200_ERANGE;
  if (leaf_count == 0) leaf_count = N - leaf_begin;
  if (leaf_count > N - leaf_begin) return FSSB200_ERANGE;  // (not begin + count > N: that sum can wrap)
  const int tbits = mode == 3 ? dcf_all_thread_bits() : kEvalAllThreadBits;
  const EvalAllPlan pl = plan_for(c, mode == 3);
  const uint64_t granule = uint64_t(1) << pl.unit_bits;
  if ((leaf_begin | leaf_count) & (granule - 1)) return FSSB200_ERANGE;
  if (nkeys == 0) return 0;
  evalall_launch_fn fn = get_evalall_launcher(mode, c->gk, c->p.prg);
  if (!fn) return FSSB200_EGROUP;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  EvalAllArgs a;
  a.seeds = static_cast<const blk *>(seeds);
  a.cws = static_cast<const uint8_t *>(cws);
  a.ocws = static_cast<const blk *>(ocws);
  a.ys = ys;
  a.nkeys = nkeys;
  a.leaf_begin = leaf_begin;
  a.leaf_count = leaf_count;
  a.ys_stride = ys_stride ? ys_stride : leaf_count;
  a.in_bits = n;
  a.party = party;
  a.unit_bits = pl.unit_bits;
  a.breadth_bits = pl.breadth_bits;
  a.dfs_bits = pl.dfs_bits;
  a.vmask = c->vmask;
  const uint64_t units = nkeys * (leaf_count >> pl.unit_bits);
  LaunchCfg cfg;
  cfg.stream = static_cast<cudaStream_t>(stream);
  const unsigned threads = 1u << tbits;
  const size_t node_bytes = mode == 3 ? 32 : 16;  // DCF nodes carry their value share
  cfg.block = dim3(threads);
  if (c->p.prg == FSSB200_PRG_AES128_MMO) {
    cfg.grid = dim3(unsigned(units < uint64_t(c->sm_count) ? units : uint64_t(c->sm_count)));
    cfg.smem = kMaxDynSmem;
  } else {
    const uint64_t cap = uint64_t(c->sm_count) * 2;
    cfg.grid = dim3(unsigned(units < cap ? units : cap));
    // cw copies + two breadth buffers + DFS stack (+ slack for alignment)
    // (16-byte leaves: the cooperative bottom stage trades one stack level for frontier tiles: 64 B per thread, half of them in the breadth buffers)
    const bool coop = mode != 2 && mode != 3 && pl.breadth_bits >= 5 && pl.dfs_bits >= 3;
    const int nstk = pl.dfs_bits - (coop ? 2 : 1);
    cfg.smem = size_t(c->ncw + 1) * 48 + 2 * threads * node_bytes + size_t(nstk > 1 ? nstk : 1) * threads * node_bytes + 64 +
        (coop ? size_t(threads) * 32 + 16 : 0) +
        (mode == 2 && pl.dfs_bits >= 5 ? (size_t(threads) * 4) << (pl.dfs_bits - 5) : 0);  // Grotto: packed leaf bits
  }
  c->launches++;
  return int(fn(c->kp, a, cfg));
}

int fssb200_eval_all(const fssb200_ctx *cc, int party, const void *seeds, const void *cws, const void *ocws,
    void *ys, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  switch (c->p.scheme) {
    case FSSB200_SCHEME_DPF:
      return evalall_impl(c, 0, party, seeds, cws, nullptr, ys, nkeys, leaf_begin, leaf_count, stream);
    case FSSB200_SCHEME_HALFTREE:
      return evalall_impl(c, 1, party, seeds, cws, ocws, ys, nkeys, leaf_begin, leaf_count, stream);
    case FSSB200_SCHEME_DCF:
      return evalall_impl(c, 3, party, seeds, cws, nullptr, ys, nkeys, leaf_begin, leaf_count, stream);
    case FSSB200_SCHEME_GROTTO: {
      if (leaf_begin != 0) return FSSB200_ERANGE;
      int rc = evalall_impl(c, 2, party, seeds, cws, nullptr, ys, nkeys, 0, leaf_count, stream);
      if (rc) return rc;
      if (nkeys == 0) return 0;
      const uint64_t cnt = leaf_count ? leaf_count : (uint64_t(1) << c->p.in_bits);
      DeviceGuard g(c->p.device);
      c->launches++;
      return int(launch_prefix_xor(static_cast<uint8_t *>(ys), nkeys, cnt, static_cast<cudaStream_t>(stream)));
    }
    default:
      return FSSB200_ESCHEME;
  }
}

// ---- VDPF ---------------------------------------------------------------------------------------------------------
int fssb200_vdpf_gen(const fssb200_ctx *cc, const void *s0s, const void *alphas, const void *betas, void *cws,
    void *cs, void *ocws, void *status, size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!s0s || !alphas || !betas || !cws || !cs || !ocws || !status) return FSSB200_EINVAL;
  if (!aligned16(s0s) || !aligned16(cws) || !aligned16(betas) || !aligned16(ocws) || !aligned16(cs)) return FSSB200_EALIGN;
  if (reinterpret_cast<uintptr_t>(alphas) % c->p.in_bytes || reinterpret_cast<uintptr_t>(status) % 4) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  if (!grp_kind_instantiated(c->gk)) return FSSB200_EGROUP;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  GenArgs a;
  std::memset(&a, 0, sizeof(a));
  a.s0s = static_cast<const blk *>(s0s);
  a.alphas = static_cast<const uint8_t *>(alphas);
  a.betas = static_cast<const blk *>(betas);
  a.cws = static_cast<uint8_t *>(cws);
  a.ocws = static_cast<blk *>(ocws);
  a.cs = static_cast<blk *>(cs);
  a.status = static_cast<int32_t *>(status);
  a.nkeys = nkeys;
  a.in_bits = c->p.in_bits;
  a.in_bytes = c->p.in_bytes;
  a.pred = c->p.pred;
  a.vmask = c->vmask;
  int rc = 0;
  const int out_mode = gen_out_mode(c, a, &rc);
  if (rc) return rc;
  gen_launch_fn fn = get_gen_launcher(FSSB200_SCHEME_VDPF, c->gk, c->p.prg, out_mode);
  if (!fn) return FSSB200_EGROUP;
  const LaunchCfg cfg = gen_cfg(c, nkeys, static_cast<cudaStream_t>(stream), out_mode);
  c->launches++;
  return int(fn(c->kp, a, cfg));
}

int fssb200_vdpf_eval(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs,
    const void *ocws, const void *xs, void *ys, void *pis, size_t nkeys, void *stream) {
  return eval_impl(c, FSSB200_SCHEME_VDPF, party, seeds, cws, ocws, xs, ys, nkeys, stream, false, nullptr, nullptr,
      nullptr, nullptr, cs, pis);
}

int fssb200_vdpf_eval_levelmajor(const fssb200_ctx *c, int party, const void *seeds, const void *cw_s,
    const void *extra, const void *cs, const void *ocws, const void *xs, void *ys, void *pis, size_t nkeys,
    void *stream) {
  return eval_impl(c, FSSB200_SCHEME_VDPF, party, seeds, nullptr, ocws, xs, ys, nkeys, stream, true, cw_s, nullptr,
      extra, nullptr, cs, pis);
}

int fssb200_vdpf_prove(const fssb200_ctx *cc, const void *pi_tildes, const void *cs, size_t m, void *pis,
    size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!cs || !pis || (m && !pi_tildes)) return FSSB200_EINVAL;
  if (!aligned16(pi_tildes) || !aligned16(cs) || !aligned16(pis)) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  c->launches++;
  return int(launch_vdpf_prove(c->kp, static_cast<const blk *>(pi_tildes), static_cast<const blk *>(cs), m,
      static_cast<blk *>(pis), nkeys, static_cast<cudaStream_t>(stream)));
}

int fssb200_vdpf_eval_all(const fssb200_ctx *cc, int party, const void *seeds, const void *cws, const void *cs,
    const void *ocws, void *ys, void *pis, size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!cs || !ocws || !pis) return FSSB200_EINVAL;
  if (!aligned16(cs) || !aligned16(ocws) || !aligned16(pis)) return FSSB200_EALIGN;
  if (c->p.in_bits > 32) return FSSB200_EDOMAIN;
  // tree: packed (s | t) leaves into ys; then leaf conversion + sequential proof chain, one warp per key
  if (int rc = evalall_impl(c, 4, party, seeds, cws, nullptr, ys, nkeys, 0, 0, stream)) return rc;
  if (nkeys == 0) return 0;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  c->launches++;
  return int(launch_vdpf_finish(c->kp, c->gk, party, c->p.in_bits, static_cast<const blk *>(cs),
      static_cast<const blk *>(ocws), static_cast<blk *>(ys), static_cast<blk *>(pis), nkeys,
      static_cast<cudaStream_t>(stream)));
}

int fssb200_hash(const fssb200_ctx *cc, int which, const void *msgs, void *out, size_t n, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if ((which != 0 && which != 1) || !msgs || !out) return FSSB200_EINVAL;
  if (!aligned16(msgs) || !aligned16(out)) return FSSB200_EALIGN;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  c->launches++;
  return int(launch_hash(c->kp, which, static_cast<const blk *>(msgs), static_cast<blk *>(out), n,
      static_cast<cudaStream_t>(stream)));
}

int fssb200_grotto_expand(const fssb200_ctx *cc, int party, const void *seeds, const void *cws, void *t,
    size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  return evalall_impl(c, 2, party, seeds, cws, nullptr, t, nkeys, leaf_begin, leaf_count, stream);
}

int fssb200_grotto_preprocess(const fssb200_ctx *cc, int party, const void *seeds, const void *cws, void *pt,
    size_t nkeys, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (!pt) return FSSB200_EINVAL;
  const int n = c->p.in_bits;
  if (n > 31) return FSSB200_EDOMAIN;
  const uint64_t N = uint64_t(1) << n;
  // leaf bits of key k go to pt[k*(2N-1) + N-1 ...] (grotto_dcf.cuh:98): one expansion launch for all keys,
  // rows 2N-1 bytes apart; then the internal nodes bottom-up (grotto_dcf.cuh:100-103), kParityLevels tree
  // levels per launch
  if (int rc = evalall_impl(c, 2, party, seeds, cws, nullptr, static_cast<uint8_t *>(pt) + (N - 1), nkeys, 0, N, stream,
          2 * N - 1))
    return rc;
  if (nkeys == 0) return 0;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  for (int bottom = n; bottom > 0; bottom -= kParityLevels) {
    c->launches++;
    cudaError_t e = launch_parity_levels(static_cast<uint8_t *>(pt), 2 * N - 1, nkeys, bottom,
        static_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return int(e);
  }
  return 0;
}

int fssb200_grotto_eval(const fssb200_ctx *cc, const void *pt, const void *xs, void *ys, size_t nkeys,
    void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (c->p.scheme != FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (!pt || !xs || !ys) return FSSB200_EINVAL;
  if (c->p.in_bits > 31) return FSSB200_EDOMAIN;
  if (nkeys == 0) return 0;
  DeviceGuard g(c->p.device);
  c->launches++;
  return int(launch_grotto_lookup(static_cast<const uint8_t *>(pt), static_cast<const uint8_t *>(xs),
      static_cast<uint8_t *>(ys), nkeys, c->p.in_bits, c->p.in_bytes, static_cast<cudaStream_t>(stream)));
}

int fssb200_grotto_eval_walk(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *xs,
    void *ys, size_t nkeys, void *stream) {
  return eval_impl(c, FSSB200_SCHEME_GROTTO, party, seeds, cws, nullptr, xs, ys, nkeys, stream, false, nullptr,
      nullptr, nullptr, nullptr);
}

// ---- PRG known-answer hook ------------------------------------------------------------------------------------------
int fssb200_prg_gen(const fssb200_ctx *cc, const void *seeds, void *out, int mul, size_t nseeds, void *stream) {
  fssb200_ctx *c = const_cast<fssb200_ctx *>(cc);
  if (int rc = check_common(c)) return rc;
  if (!seeds || !out) return FSSB200_EINVAL;
  if (!aligned16(seeds) || !aligned16(out)) return FSSB200_EALIGN;
  prg_launch_fn fn = get_prg_launcher(c->p.prg, mul);
  if (!fn) return FSSB200_EINVAL;
  if (nseeds == 0) return 0;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  // the KAT hook uses keys 0..mul-1 as stored (prg_key), independent of the scheme's mul
  const LaunchCfg cfg = point_cfg(c, nseeds, static_cast<cudaStream_t>(stream));
  c->launches++;
  return int(fn(c->kp, static_cast<const blk *>(seeds), static_cast<blk *>(out), nseeds, cfg));
}

int fssb200_microbench(int device, int kind, double *ops_per_s) {
  if (!ops_per_s) return FSSB200_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return FSSB200_ENODEVICE;
  DeviceGuard g(device);
  if (g.err != cudaSuccess) return int(g.err);
  return run_microbench(kind, ops_per_s);
}

}  // extern "C"

namespace fssb200 {

point_launch_fn get_point_launcher(int scheme, int gk, int prg, int mode) {
  if (prg == kPrgAes) {
    if (scheme == FSSB200_SCHEME_DPF) return point_launcher_aes_dpf(gk, mode);
    if (scheme == FSSB200_SCHEME_DCF) return point_launcher_aes_dcf(gk, mode);
    if (scheme == FSSB200_SCHEME_HALFTREE) return point_launcher_aes_ht(gk, mode);
    if (scheme == FSSB200_SCHEME_VDPF) return point_launcher_aes_vdpf(gk, mode);
    if (scheme == FSSB200_SCHEME_GROTTO) return point_launcher_aes_grotto(gk, mode);
  } else {
    if (scheme == FSSB200_SCHEME_DPF) return point_launcher_chacha_dpf(gk, mode);
    if (scheme == FSSB200_SCHEME_DCF) return point_launcher_chacha_dcf(gk, mode);
    if (scheme == FSSB200_SCHEME_HALFTREE) return point_launcher_chacha_ht(gk, mode);
    if (scheme == FSSB200_SCHEME_VDPF) return point_launcher_chacha_vdpf(gk, mode);
    if (scheme == FSSB200_SCHEME_GROTTO) return point_launcher_chacha_grotto(gk, mode);
  }
  return nullptr;
}
gen_launch_fn get_gen_launcher(int scheme, int gk, int prg, int out_mode) {
  if (prg == kPrgAes) {
    if (scheme == FSSB200_SCHEME_DPF) return gen_launcher_aes_dpf(gk, out_mode);
    if (scheme == FSSB200_SCHEME_DCF) return gen_launcher_aes_dcf(gk, out_mode);
    if (scheme == FSSB200_SCHEME_HALFTREE) return gen_launcher_aes_ht(gk, out_mode);
    if (scheme == FSSB200_SCHEME_VDPF) return gen_launcher_aes_vdpf(gk, out_mode);
  } else {
    if (scheme == FSSB200_SCHEME_DPF) return gen_launcher_chacha_dpf(gk, out_mode);
    if (scheme == FSSB200_SCHEME_DCF) return gen_launcher_chacha_dcf(gk, out_mode);
    if (scheme == FSSB200_SCHEME_HALFTREE) return gen_launcher_chacha_ht(gk, out_mode);
    if (scheme == FSSB200_SCHEME_VDPF) return gen_launcher_chacha_vdpf(gk, out_mode);
  }
  return nullptr;
}
evalall_launch_fn get_evalall_launcher(int mode, int gk, int prg) {
  if (prg == kPrgAes) {
    if (mode == 0) return evalall_launcher_aes_dpf(gk);
    if (mode == 1) return evalall_launcher_aes_ht(gk);
    if (mode == 2) return evalall_launcher_aes_grotto(gk);
    if (mode == 3) return evalall_launcher_aes_dcf(gk);
    if (mode == 4) return evalall_launcher_aes_vdpf(gk);
  } else {
    if (mode == 0) return evalall_launcher_chacha_dpf(gk);
    if (mode == 1) return evalall_launcher_chacha_ht(gk);
    if (mode == 2) return evalall_launcher_chacha_grotto(gk);
    if (mode == 3) return evalall_launcher_chacha_dcf(gk);
    if (mode == 4) return evalall_launcher_chacha_vdpf(gk);
  }
  return nullptr;
}
prg_launch_fn get_prg_launcher(int prg, int mul) {
  return prg == kPrgAes ? prg_launcher_aes(mul) : prg_launcher_chacha(mul);
}

}  // namespace fssb200
