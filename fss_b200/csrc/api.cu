// SPDX-License-Identifier: Apache-2.0
//
// api.cu -- the extern "C" boundary of include/fssb200.h: context, argument validation, launch
// geometry, host-buffer staging.  There is no CPU evaluation path in this library: every entry
// point either launches sm_100a kernels or returns an error code.
#include <atomic>
#include <condition_variable>
#include <cuda.h>  // CUtensorMap + enums only; cuTensorMapEncodeTiled is resolved at run time (no libcuda link dependency)
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>
#if defined(__linux__)
#include <sched.h>
#endif
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "dispatch.h"
#include "misc_kernels.cuh"
#include "vdpf_kernels.cuh"

using namespace fssb200;

// Worker threads of the host entry points (row packing).  parallel_for blocks until [0, total) is done; the caller
// takes blocks too.
class PackPool {
 public:
  explicit PackPool(int nworkers) {
    for (int i = 0; i < nworkers; ++i) th_.emplace_back([this] { worker(); });
  }
  ~PackPool() {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_start_.notify_all();
    for (auto &t : th_) t.join();
  }
  int workers() const { return int(th_.size()); }
  void parallel_for(size_t total, size_t grain, const std::function<void(size_t, size_t)> &fn) {
    {
      std::lock_guard<std::mutex> l(mu_);
      fn_ = &fn;
      total_ = total;
      grain_ = grain ? grain : 1;
      next_.store(0);
      active_ = int(th_.size());
      ++gen_;
    }
    cv_start_.notify_all();
    drain(fn);
    std::unique_lock<std::mutex> l(mu_);
    cv_done_.wait(l, [this] { return active_ == 0; });
    fn_ = nullptr;
  }

 private:
  void drain(const std::function<void(size_t, size_t)> &fn) {
    for (;;) {
      const size_t b = next_.fetch_add(grain_);
      if (b >= total_) break;
      fn(b, b + grain_ < total_ ? b + grain_ : total_);
    }
  }
  void worker() {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(size_t, size_t)> *fn;
      {
        std::unique_lock<std::mutex> l(mu_);
        cv_start_.wait(l, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
        fn = fn_;
      }
      drain(*fn);
      {
        std::lock_guard<std::mutex> l(mu_);
        if (--active_ == 0) cv_done_.notify_all();
      }
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_start_, cv_done_;
  const std::function<void(size_t, size_t)> *fn_ = nullptr;
  size_t total_ = 0, grain_ = 1;
  std::atomic<size_t> next_{0};
  uint64_t gen_ = 0;
  int active_ = 0;
  bool stop_ = false;
};

constexpr int kStageSlots = 3;
struct HostArena {
  size_t chunk_keys = 0;
  size_t bytes_per_set = 0;
  uint8_t *dev[2] = {nullptr, nullptr};
  cudaStream_t stream[2] = {nullptr, nullptr};
  // packed host path of fssb200_eval_host (DPF / Half-Tree): pinned staging for the packed rows of a chunk
  uint8_t *stage[kStageSlots] = {nullptr, nullptr, nullptr};
  cudaEvent_t stage_ev[kStageSlots] = {nullptr, nullptr, nullptr};
  bool stage_busy[kStageSlots] = {false, false, false};
  PackPool *pool = nullptr;
};

struct fssb200_ctx {
  fssb200_params p;
  KParams kp;
  int gk;           // group kind (common.cuh)
  int ncw;
  int mul;
  int sm_count;
  int max_smem_optin;
  uint32_t vmask;
  int point_mode;   // PointMode of the key-major point kernels (kernels.cuh); FSSB200_POINT_MODE overrides
  int gen_mode;     // gen kernels: 1 = Cw tiles written by the TMA unit (CwTileOut), 0 = direct stores; FSSB200_GEN_MODE
  std::atomic<uint64_t> launches{0};
  HostArena arena;
};

namespace {

#define CUDA_TRY(expr)                         \
  do {                                         \
    cudaError_t e__ = (expr);                  \
    if (e__ != cudaSuccess) return int(e__);   \
  } while (0)

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Tensor map of the key-major Cw array for the TMA-fed point kernels (CwTile in kernels.cuh).
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int make_rows_tensor_map(uint8_t out[128], const void *rows, size_t nkeys, size_t row_bytes);
int make_cw_tensor_map(uint8_t out[128], const void *cws, size_t nkeys, int ncw) {
  return make_rows_tensor_map(out, cws, nkeys, size_t(ncw) * 32u);
}
// 2-D byte tensor [nkeys][row_bytes], box 32 rows x 64 B, SWIZZLE_64B (CwTileT / CwTileOut in kernels.cuh)
int make_rows_tensor_map(uint8_t out[128], const void *cws, size_t nkeys, size_t row_bytes) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess) return int(e);
    if (q != cudaDriverEntryPointSuccess || !p) return int(cudaErrorNotSupported);
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
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

struct DeviceGuard {
  int prev = -1;
  cudaError_t err;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
  }
  ~DeviceGuard() {
    int cur = -1;
    if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
  }
};

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
    const uint64_t want = (n + threads - 1) / threads;
    cfg.grid = dim3(unsigned(want < uint64_t(c->sm_count) ? (want ? want : 1) : c->sm_count));
    cfg.block = dim3(threads);
    cfg.smem = kMaxDynSmem;
  } else {
    const uint64_t want = (n + 255) / 256;
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
EvalAllPlan plan_evalall(int n, int thread_bits = kEvalAllThreadBits) {
  EvalAllPlan pl;
  int dfs = n - thread_bits;
  if (dfs < 1) dfs = 1;
  if (dfs > kMaxDfsBits) dfs = kMaxDfsBits;
  int bt = n - dfs;
  if (bt > thread_bits) bt = thread_bits;
  pl.dfs_bits = dfs;
  pl.breadth_bits = bt;
  pl.unit_bits = bt + dfs;
  return pl;
}

int check_common(const fssb200_ctx *c) { return c ? 0 : FSSB200_EINVAL; }

}  // namespace

namespace {
// rows [k0, k1): ncw x {16 B s} + 16 B of flag bits (bit i = byte 16 of entry i != 0, i < 128).
// STREAM: non-temporal stores for the whole row, flag word included -- a regular store into a line that is still
// being assembled in a write-combining buffer forces a flush + read-for-ownership and cost 3.5x (measured).
template <bool STREAM>
void pack_rows_range_t(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw) {
  const size_t in_row = size_t(ncw) * 32u, out_row = size_t(ncw) * 16u + 16u;
  const int nflag = ncw < 128 ? ncw : 128;
  for (size_t k = k0; k < k1; ++k) {
    const uint8_t *r = src + k * in_row;
    uint8_t *o = dst + k * out_row;
    uint64_t f0 = 0, f1 = 0;
    for (int i = 0; i < nflag && i < 64; ++i) f0 |= uint64_t(r[32 * i + 16] != 0) << i;
    for (int i = 64; i < nflag; ++i) f1 |= uint64_t(r[32 * i + 16] != 0) << (i - 64);
#if defined(__SSE2__)
    for (int i = 0; i < ncw; ++i) {
      const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(r + 32 * i));
      if (STREAM) _mm_stream_si128(reinterpret_cast<__m128i *>(o + 16 * i), v);
      else _mm_storeu_si128(reinterpret_cast<__m128i *>(o + 16 * i), v);
    }
    const __m128i fv = _mm_set_epi64x(static_cast<long long>(f1), static_cast<long long>(f0));
    if (STREAM) _mm_stream_si128(reinterpret_cast<__m128i *>(o + size_t(ncw) * 16u), fv);
    else _mm_storeu_si128(reinterpret_cast<__m128i *>(o + size_t(ncw) * 16u), fv);
#else
    for (int i = 0; i < ncw; ++i) std::memcpy(o + 16 * i, r + 32 * i, 16);
    const uint64_t f[2] = {f0, f1};
    std::memcpy(o + size_t(ncw) * 16u, f, 16);
#endif
  }
#if defined(__SSE2__)
  if (STREAM) _mm_sfence();
#endif
}
void pack_rows_range(const uint8_t *src, uint8_t *dst, size_t k0, size_t k1, int ncw, bool stream_stores) {
  if (stream_stores) pack_rows_range_t<true>(src, dst, k0, k1, ncw);
  else pack_rows_range_t<false>(src, dst, k0, k1, ncw);
}
int usable_cpus() {
#if defined(__linux__)
  cpu_set_t set;
  if (sched_getaffinity(0, sizeof(set), &set) == 0) return CPU_COUNT(&set);
#endif
  const unsigned h = std::thread::hardware_concurrency();
  return h ? int(h) : 1;
}
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
  // cp.async slabs for every scheme; DPF / DCF gain another 1-3 % from 24 warps per SM, Half-Tree does not
  c->point_mode = q.scheme == FSSB200_SCHEME_HALFTREE ? 4 : 5;
  if (const char *e = std::getenv("FSSB200_POINT_MODE")) {  // A/B measurement knob
    const int m = std::atoi(e);
    if (m == 0 || m == 1 || m == 3 || m == 4 || m == 5) c->point_mode = m;
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
  c->kp.ga.vmask = vmask;
  c->kp.ga.mod[0] = uint32_t(q.mod_lo);
  c->kp.ga.mod[1] = uint32_t(q.mod_lo >> 32);
  c->kp.ga.mod[2] = uint32_t(q.mod_hi);
  c->kp.ga.mod[3] = uint32_t(q.mod_hi >> 32);
  *out = c;
  return 0;
}

void fssb200_ctx_destroy(fssb200_ctx *c) {
  if (!c) return;
  {
    DeviceGuard g(c->p.device);
    for (int i = 0; i < 2; ++i) {
      if (c->arena.dev[i]) cudaFree(c->arena.dev[i]);
      if (c->arena.stream[i]) cudaStreamDestroy(c->arena.stream[i]);
    }
    for (int i = 0; i < kStageSlots; ++i) {
      if (c->arena.stage[i]) cudaFreeHost(c->arena.stage[i]);
      if (c->arena.stage_ev[i]) cudaEventDestroy(c->arena.stage_ev[i]);
    }
  }
  delete c->arena.pool;
  delete c;
}

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
  if (scheme == FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;  // Grotto: EvalAll / Preprocess+Eval only
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
  if (!aligned16(seeds) || !aligned16(ys) || !aligned16(ocws)) return FSSB200_EALIGN;
  if (reinterpret_cast<uintptr_t>(xs) % c->p.in_bytes) return FSSB200_EALIGN;
  if (nkeys == 0) return 0;
  const int mode = level_major ? 2 : (packed ? 6 : c->point_mode);
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
  if (mode >= 4) {
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


int fssb200_ctx_host_pack_threads(const fssb200_ctx *c) {
  return (c && c->arena.pool && c->arena.stage[0]) ? c->arena.pool->workers() + 1 : 0;
}

int fssb200_pack_rows(const fssb200_ctx *c, const void *cws, void *rows, size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!fssb200_packed_row_bytes(c)) return FSSB200_ESCHEME;
  if (!cws || !rows) return FSSB200_EINVAL;
  const bool al = aligned16(rows);
  PackPool *pool = c->arena.pool;
  const uint8_t *src = static_cast<const uint8_t *>(cws);
  uint8_t *dst = static_cast<uint8_t *>(rows);
  const int ncw = c->ncw;
  if (pool && nkeys >= 4096) {
    pool->parallel_for(nkeys, 1024, [=](size_t b, size_t e) { pack_rows_range(src, dst, b, e, ncw, al); });
  } else {
    pack_rows_range(src, dst, 0, nkeys, ncw, al);
  }
  return 0;
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
static int evalall_thread_bits(const fssb200_ctx *c) {
  return c->p.scheme == FSSB200_SCHEME_DCF ? kDcfAllThreadBits : kEvalAllThreadBits;
}
uint64_t fssb200_eval_all_granule(const fssb200_ctx *c) {
  if (!c) return 0;
  return uint64_t(1) << plan_evalall(c->p.in_bits, evalall_thread_bits(c)).unit_bits;
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
  if (leaf_begin + leaf_count > N) return FSSB200_ERANGE;
  const int tbits = mode == 3 ? kDcfAllThreadBits : kEvalAllThreadBits;
  const EvalAllPlan pl = plan_evalall(n, tbits);
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

// ---- host-buffer entry points -------------------------------------------------------------------------------------------
int fssb200_ctx_reserve_host(fssb200_ctx *c, size_t max_keys_per_chunk) {
  if (int rc = check_common(c)) return rc;
  if (max_keys_per_chunk == 0) max_keys_per_chunk = size_t(1) << 18;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  // per key: seeds(2 for gen) + cws + ocw + x/alpha(16) + beta + y; evalall output is staged in the
  // same buffers (>= 64 MiB per set)
  size_t per_key = 32 + size_t(c->ncw) * 32 + 16 + 16 + 16 + 16 + 64 + 64 + 16;  // + VDPF cs, pis, status
  size_t bytes = per_key * max_keys_per_chunk;
  if (bytes < (size_t(64) << 20)) bytes = size_t(64) << 20;
  bytes = (bytes + 255) & ~size_t(255);
  for (int i = 0; i < 2; ++i) {
    if (a.dev[i]) { cudaFree(a.dev[i]); a.dev[i] = nullptr; }
    CUDA_TRY(cudaMalloc(&a.dev[i], bytes));
    if (!a.stream[i]) CUDA_TRY(cudaStreamCreateWithFlags(&a.stream[i], cudaStreamNonBlocking));
  }
  a.chunk_keys = max_keys_per_chunk;
  a.bytes_per_set = bytes;
  // Packed host path (DPF / Half-Tree): 15 of the 32 bytes of a Cw are padding, and the host-buffer calls are bound
  // by the PCIe link (profiles/r01_h2d_probe_n1.json).  If this process has enough cores to strip the padding faster
  // than the link moves it (measured: 11 GB/s per core, 70-100 GB/s on 16 cores), worker threads pack each chunk into
  // pinned staging while the previous chunk is in flight and half the bytes cross the link: 82 -> 57 ms for 2^22 keys.
  // Packing trades link bytes for host-memory traffic (9 GB instead of 4.5 GB per 2^22 keys), so it is only used when
  // this is the only rank on the host (torchrun's LOCAL_WORLD_SIZE): with several GPUs every link already draws its
  // share of the host memory bandwidth (8 ranks: 173 GB/s aggregate, profiles/bench_r01_session6_n8.json).
  // FSSB200_PACK_THREADS overrides the thread count (0 = off).
  for (int i = 0; i < kStageSlots; ++i) {
    if (a.stage[i]) { cudaFreeHost(a.stage[i]); a.stage[i] = nullptr; }
    a.stage_busy[i] = false;
  }
  delete a.pool;
  a.pool = nullptr;
  int threads = usable_cpus();
  if (const char *e = std::getenv("LOCAL_WORLD_SIZE")) {
    if (std::atoi(e) > 1) threads = 0;
  }
  if (threads > 32) threads = 32;
  if (const char *e = std::getenv("FSSB200_PACK_THREADS")) threads = std::atoi(e);
  const size_t row = fssb200_packed_row_bytes(c);
  if (row && threads >= 6) {
    // packing is an optimisation: if the pinned staging cannot be had, the call copies the reference layout as it is
    bool ok = true;
    for (int i = 0; i < kStageSlots && ok; ++i) {
      ok = cudaHostAlloc(reinterpret_cast<void **>(&a.stage[i]), row * max_keys_per_chunk, cudaHostAllocDefault) == cudaSuccess;
      if (ok && !a.stage_ev[i]) ok = cudaEventCreateWithFlags(&a.stage_ev[i], cudaEventDisableTiming) == cudaSuccess;
    }
    if (ok) a.pool = new (std::nothrow) PackPool(threads - 1);
    if (!ok || !a.pool) {
      (void)cudaGetLastError();
      for (int i = 0; i < kStageSlots; ++i) {
        if (a.stage[i]) { cudaFreeHost(a.stage[i]); a.stage[i] = nullptr; }
      }
    }
  }
  return 0;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int fssb200_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    const void *xs, void *ys, size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  if (c->p.scheme == FSSB200_SCHEME_GROTTO || c->p.scheme == FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!seeds || !cws || !xs || !ys) return FSSB200_EINVAL;
  if (c->p.scheme == FSSB200_SCHEME_HALFTREE && !ocws) return FSSB200_EINVAL;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t ck = a.chunk_keys, cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  const size_t rowb = fssb200_packed_row_bytes(c);
  const bool can_pack = a.pool && a.stage[0] && rowb && nkeys >= 8192;  // see fssb200_ctx_reserve_host
  // A/B knob: every `direct_every`-th chunk crosses the link unpacked (0 = pack every chunk, the default).  Measured
  // (profiles/r01_host_pack.md): mixing does not help -- the packing cores and the DMA reads compete for the same host
  // memory bandwidth (~160 GB/s on the box), which, not the cores, is what bounds the packed path.
  static const int direct_every = [] {
    const char *e = std::getenv("FSSB200_PACK_DIRECT_EVERY");
    const int v = e ? std::atoi(e) : 0;
    return v < 0 ? 0 : v;
  }();
  size_t pslot = 0;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = nkeys - k0 < ck ? nkeys - k0 : ck;
    const int b = int(chunk & 1);
    cudaStream_t s = a.stream[b];
    uint8_t *d_seeds = a.dev[b];
    uint8_t *d_cws = d_seeds + align_up(k * 16, 256);
    uint8_t *d_ocws = d_cws + align_up(k * cwb, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * 16, 256);
    CUDA_TRY(cudaMemcpyAsync(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    const bool packed = can_pack && !(direct_every && chunk % size_t(direct_every) == size_t(direct_every) - 1);
    if (packed) {
      // pack this chunk while the copies of the previous ones are in flight
      const int slot = int(pslot++ % kStageSlots);
      if (a.stage_busy[slot]) CUDA_TRY(cudaEventSynchronize(a.stage_ev[slot]));
      const uint8_t *src = static_cast<const uint8_t *>(cws) + k0 * cwb;
      uint8_t *dst = a.stage[slot];
      const int ncw = c->ncw;
      a.pool->parallel_for(k, 512, [=](size_t b, size_t e) { pack_rows_range(src, dst, b, e, ncw, true); });
      CUDA_TRY(cudaMemcpyAsync(d_cws, dst, k * rowb, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaEventRecord(a.stage_ev[slot], s));
      a.stage_busy[slot] = true;
    } else {
      CUDA_TRY(cudaMemcpyAsync(d_cws, static_cast<const uint8_t *>(cws) + k0 * cwb, k * cwb, cudaMemcpyHostToDevice, s));
    }
    if (ocws)
      CUDA_TRY(cudaMemcpyAsync(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, cudaMemcpyHostToDevice, s));
    rc = packed ? fssb200_eval_packed(c, party, d_seeds, d_cws, ocws ? d_ocws : nullptr, d_xs, d_ys, k, s)
                : fssb200_eval(c, party, d_seeds, d_cws, ocws ? d_ocws : nullptr, d_xs, d_ys, k, s);
    if (rc) break;
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  for (int i = 0; i < kStageSlots; ++i) a.stage_busy[i] = false;
  return rc;
}

int fssb200_vdpf_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *cs,
    void *ocws, void *status, size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!s0s || !alphas || !betas || !cws || !cs || !ocws || !status) return FSSB200_EINVAL;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t ck = a.chunk_keys, cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = nkeys - k0 < ck ? nkeys - k0 : ck;
    const int b = int(chunk & 1);
    cudaStream_t s = a.stream[b];
    // the arena holds >= 96 + ncw*32 bytes per key (fssb200_ctx_reserve_host); VDPF needs 32 + 16 + 16 + cwb + 64 + 16 + 4
    uint8_t *d_s0s = a.dev[b];
    uint8_t *d_al = d_s0s + align_up(k * 32, 256);
    uint8_t *d_be = d_al + align_up(k * 16, 256);
    uint8_t *d_cws = d_be + align_up(k * 16, 256);
    uint8_t *d_cs = d_cws + align_up(k * cwb, 256);
    uint8_t *d_ocws = d_cs + align_up(k * 64, 256);
    uint8_t *d_st = d_ocws + align_up(k * 16, 256);
    if (size_t(d_st + k * 4 - a.dev[b]) > a.bytes_per_set) return FSSB200_ENOARENA;
    CUDA_TRY(cudaMemcpyAsync(d_s0s, static_cast<const uint8_t *>(s0s) + k0 * 32, k * 32, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_al, static_cast<const uint8_t *>(alphas) + k0 * ib, k * ib, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_be, static_cast<const uint8_t *>(betas) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemsetAsync(d_ocws, 0, k * 16, s));  // Gen leaves ocw untouched when it returns 1
    rc = fssb200_vdpf_gen(c, d_s0s, d_al, d_be, d_cws, d_cs, d_ocws, d_st, k, s);
    if (rc) break;
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(cws) + k0 * cwb, d_cws, k * cwb, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(cs) + k0 * 64, d_cs, k * 64, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ocws) + k0 * 16, d_ocws, k * 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(status) + k0 * 4, d_st, k * 4, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  return rc;
}

int fssb200_vdpf_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs,
    const void *ocws, const void *xs, void *ys, void *pis, size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  if (c->p.scheme != FSSB200_SCHEME_VDPF) return FSSB200_ESCHEME;
  if (!seeds || !cws || !cs || !ocws || !xs || !ys || !pis) return FSSB200_EINVAL;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t ck = a.chunk_keys, cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = nkeys - k0 < ck ? nkeys - k0 : ck;
    const int b = int(chunk & 1);
    cudaStream_t s = a.stream[b];
    uint8_t *d_seeds = a.dev[b];
    uint8_t *d_cws = d_seeds + align_up(k * 16, 256);
    uint8_t *d_cs = d_cws + align_up(k * cwb, 256);
    uint8_t *d_ocws = d_cs + align_up(k * 64, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * 16, 256);
    uint8_t *d_pis = d_ys + align_up(k * 16, 256);
    if (size_t(d_pis + k * 64 - a.dev[b]) > a.bytes_per_set) return FSSB200_ENOARENA;
    CUDA_TRY(cudaMemcpyAsync(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_cws, static_cast<const uint8_t *>(cws) + k0 * cwb, k * cwb, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_cs, static_cast<const uint8_t *>(cs) + k0 * 64, k * 64, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, cudaMemcpyHostToDevice, s));
    rc = fssb200_vdpf_eval(c, party, d_seeds, d_cws, d_cs, d_ocws, d_xs, d_ys, d_pis, k, s);
    if (rc) break;
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(pis) + k0 * 64, d_pis, k * 64, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  return rc;
}

int fssb200_eval_levelmajor_host(fssb200_ctx *c, int party, const void *seeds, const void *cw_s, const void *cw_v,
    const void *extra, const void *out_cw, const void *ocws, const void *xs, void *ys, size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  const int scheme = c->p.scheme;
  if (scheme == FSSB200_SCHEME_GROTTO) return FSSB200_ESCHEME;
  if (!seeds || !cw_s || !xs || !ys) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_DCF && (!cw_v || !out_cw)) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_DPF && (!extra || !out_cw)) return FSSB200_EINVAL;
  if (scheme == FSSB200_SCHEME_HALFTREE && (!extra || !ocws)) return FSSB200_EINVAL;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t ck = a.chunk_keys, ib = size_t(c->p.in_bytes), n = size_t(c->p.in_bits), nw = (n + 31) / 32;
  const uint8_t *h_s = static_cast<const uint8_t *>(cw_s), *h_v = static_cast<const uint8_t *>(cw_v),
                *h_e = static_cast<const uint8_t *>(extra);
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = nkeys - k0 < ck ? nkeys - k0 : ck;
    const int b = int(chunk & 1);
    cudaStream_t st = a.stream[b];
    // one set: seeds | cw_s[n][k] | cw_v[n][k] | extra[nw][k] | out_cw | ocws | xs | ys   (<= bytes_per_set by construction)
    uint8_t *d_seeds = a.dev[b];
    uint8_t *d_s = d_seeds + align_up(k * 16, 256);
    uint8_t *d_v = d_s + align_up(n * k * 16, 256);
    uint8_t *d_e = d_v + (cw_v ? align_up(n * k * 16, 256) : 0);
    uint8_t *d_oc = d_e + (extra ? align_up(nw * k * 4, 256) : 0);
    uint8_t *d_ocws = d_oc + align_up(k * 16, 256);
    uint8_t *d_xs = d_ocws + align_up(k * 16, 256);
    uint8_t *d_ys = d_xs + align_up(k * 16, 256);
    if (size_t(d_ys + k * 16 - a.dev[b]) > a.bytes_per_set) return FSSB200_ENOARENA;
    CUDA_TRY(cudaMemcpyAsync(d_seeds, static_cast<const uint8_t *>(seeds) + k0 * 16, k * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpy2DAsync(d_s, k * 16, h_s + k0 * 16, nkeys * 16, k * 16, n, cudaMemcpyHostToDevice, st));
    if (cw_v) CUDA_TRY(cudaMemcpy2DAsync(d_v, k * 16, h_v + k0 * 16, nkeys * 16, k * 16, n, cudaMemcpyHostToDevice, st));
    if (extra) CUDA_TRY(cudaMemcpy2DAsync(d_e, k * 4, h_e + k0 * 4, nkeys * 4, k * 4, nw, cudaMemcpyHostToDevice, st));
    if (out_cw)
      CUDA_TRY(cudaMemcpyAsync(d_oc, static_cast<const uint8_t *>(out_cw) + k0 * 16, k * 16, cudaMemcpyHostToDevice, st));
    if (ocws)
      CUDA_TRY(cudaMemcpyAsync(d_ocws, static_cast<const uint8_t *>(ocws) + k0 * 16, k * 16, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(d_xs, static_cast<const uint8_t *>(xs) + k0 * ib, k * ib, cudaMemcpyHostToDevice, st));
    rc = fssb200_eval_levelmajor(c, party, d_seeds, d_s, cw_v ? d_v : nullptr, extra ? d_e : nullptr,
        out_cw ? d_oc : nullptr, ocws ? d_ocws : nullptr, d_xs, d_ys, k, st);
    if (rc) break;
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ys) + k0 * 16, d_ys, k * 16, cudaMemcpyDeviceToHost, st));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  return rc;
}

int fssb200_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *ocws,
    size_t nkeys) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  if (!s0s || !alphas || !cws) return FSSB200_EINVAL;
  const bool grotto = c->p.scheme == FSSB200_SCHEME_GROTTO, half = c->p.scheme == FSSB200_SCHEME_HALFTREE;
  if ((!grotto && !betas) || (half && !ocws)) return FSSB200_EINVAL;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t ck = a.chunk_keys, cwb = size_t(c->ncw) * 32, ib = size_t(c->p.in_bytes);
  int rc = 0;
  size_t chunk = 0;
  for (size_t k0 = 0; k0 < nkeys && !rc; k0 += ck, ++chunk) {
    const size_t k = nkeys - k0 < ck ? nkeys - k0 : ck;
    const int b = int(chunk & 1);
    cudaStream_t s = a.stream[b];
    uint8_t *d_s0s = a.dev[b];
    uint8_t *d_cws = d_s0s + align_up(k * 32, 256);
    uint8_t *d_ocws = d_cws + align_up(k * cwb, 256);
    uint8_t *d_al = d_ocws + align_up(k * 16, 256);
    uint8_t *d_be = d_al + align_up(k * 16, 256);
    CUDA_TRY(cudaMemcpyAsync(d_s0s, static_cast<const uint8_t *>(s0s) + k0 * 32, k * 32, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_al, static_cast<const uint8_t *>(alphas) + k0 * ib, k * ib, cudaMemcpyHostToDevice, s));
    if (!grotto)
      CUDA_TRY(cudaMemcpyAsync(d_be, static_cast<const uint8_t *>(betas) + k0 * 16, k * 16, cudaMemcpyHostToDevice, s));
    rc = fssb200_gen(c, d_s0s, d_al, grotto ? nullptr : d_be, d_cws, half ? d_ocws : nullptr, k, s);
    if (rc) break;
    CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(cws) + k0 * cwb, d_cws, k * cwb, cudaMemcpyDeviceToHost, s));
    if (half)
      CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ocws) + k0 * 16, d_ocws, k * 16, cudaMemcpyDeviceToHost, s));
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  return rc;
}

int fssb200_eval_all_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws,
    void *ys, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count) {
  if (int rc = check_common(c)) return rc;
  if (!c->arena.chunk_keys) return FSSB200_ENOARENA;
  if (!seeds || !cws || !ys) return FSSB200_EINVAL;
  const bool half = c->p.scheme == FSSB200_SCHEME_HALFTREE, grotto = c->p.scheme == FSSB200_SCHEME_GROTTO;
  if (half && !ocws) return FSSB200_EINVAL;
  const int n = c->p.in_bits;
  if (n > 40) return FSSB200_EDOMAIN;
  const uint64_t N = uint64_t(1) << n;
  if (leaf_begin >= N) return FSSB200_ERANGE;
  if (leaf_count == 0) leaf_count = N - leaf_begin;
  if (leaf_begin + leaf_count > N) return FSSB200_ERANGE;
  const uint64_t granule = fssb200_eval_all_granule(c);
  if ((leaf_begin | leaf_count) & (granule - 1)) return FSSB200_ERANGE;
  DeviceGuard g(c->p.device);
  if (g.err != cudaSuccess) return int(g.err);
  HostArena &a = c->arena;
  const size_t cwb = size_t(c->ncw) * 32, leaf_bytes = grotto ? 1 : 16;
  // key material of one key (seed + cws + ocw) at the front of each set, leaves behind it
  const size_t hdr = align_up(16 + cwb + 16, 256);
  uint64_t leaves_per_chunk = ((a.bytes_per_set - hdr) / leaf_bytes) / granule * granule;
  if (leaves_per_chunk == 0) return FSSB200_ENOARENA;
  if (grotto) leaves_per_chunk = leaf_count <= leaves_per_chunk ? leaf_count : 0;  // the scan needs whole keys
  if (leaves_per_chunk == 0) return FSSB200_ENOARENA;
  int rc = 0;
  size_t chunk = 0;
  for (size_t k = 0; k < nkeys && !rc; ++k) {
    for (uint64_t l0 = 0; l0 < leaf_count && !rc; l0 += leaves_per_chunk, ++chunk) {
      const uint64_t cnt = leaf_count - l0 < leaves_per_chunk ? leaf_count - l0 : leaves_per_chunk;
      const int b = int(chunk & 1);
      cudaStream_t s = a.stream[b];
      uint8_t *d_seed = a.dev[b], *d_cws = d_seed + 16, *d_ocw = d_cws + cwb, *d_ys = a.dev[b] + hdr;
      CUDA_TRY(cudaMemcpyAsync(d_seed, static_cast<const uint8_t *>(seeds) + k * 16, 16, cudaMemcpyHostToDevice, s));
      CUDA_TRY(cudaMemcpyAsync(d_cws, static_cast<const uint8_t *>(cws) + k * cwb, cwb, cudaMemcpyHostToDevice, s));
      if (half)
        CUDA_TRY(cudaMemcpyAsync(d_ocw, static_cast<const uint8_t *>(ocws) + k * 16, 16, cudaMemcpyHostToDevice, s));
      rc = fssb200_eval_all(c, party, d_seed, d_cws, half ? d_ocw : nullptr, d_ys, 1, leaf_begin + l0, cnt, s);
      if (rc) break;
      CUDA_TRY(cudaMemcpyAsync(static_cast<uint8_t *>(ys) + (k * leaf_count + l0) * leaf_bytes, d_ys, cnt * leaf_bytes,
          cudaMemcpyDeviceToHost, s));
    }
  }
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaStreamSynchronize(a.stream[i]);
    if (!rc && e != cudaSuccess) rc = int(e);
  }
  return rc;
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
  } else {
    if (scheme == FSSB200_SCHEME_DPF) return point_launcher_chacha_dpf(gk, mode);
    if (scheme == FSSB200_SCHEME_DCF) return point_launcher_chacha_dcf(gk, mode);
    if (scheme == FSSB200_SCHEME_HALFTREE) return point_launcher_chacha_ht(gk, mode);
    if (scheme == FSSB200_SCHEME_VDPF) return point_launcher_chacha_vdpf(gk, mode);
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
