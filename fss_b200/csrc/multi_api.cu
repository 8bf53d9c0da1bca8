// SPDX-License-Identifier: Apache-2.0
//
// multi_api.cu -- multi-device entry points of include/fssb200.h: ONE process drives `ndev` GPUs (SURVEY.md
// section 8b / 8e: "multi-GPU variants taking ndev + per-device pointer arrays", "one process, one stream + events
// per device").  The path shards with no exchange step -- keys are independent (dpf.cuh:170-214) and an EvalAll
// subtree depends only on its root (dpf.cuh:291-301) -- so a multi-device call is `ndev` independent stream-ordered
// launches, one per device, each on that device's arrays; nothing is copied between devices and no collective runs.
// Per-device errors are surfaced through `rcs` (negative = argument class, positive = cudaError_t of that device).
#include <cstring>

#include "ctx.h"

using namespace fssb200;

namespace {

int check_ctxs(const fssb200_ctx *const *ctxs, int ndev) {
  if (!ctxs || ndev < 1 || ndev > 64) return FSSB200_EINVAL;
  for (int d = 0; d < ndev; ++d) {
    if (!ctxs[d]) return FSSB200_EINVAL;
    const fssb200_params &a = ctxs[0]->p, &b = ctxs[d]->p;
    // one key batch = one parameter set (the key material may not differ either: same scheme object on every GPU)
    if (a.scheme != b.scheme || a.in_bits != b.in_bits || a.in_bytes != b.in_bytes || a.group != b.group ||
        a.mod_lo != b.mod_lo || a.mod_hi != b.mod_hi || a.prg != b.prg || std::memcmp(a.prg_key, b.prg_key, 64) ||
        std::memcmp(a.hash_key, b.hash_key, 16))
      return FSSB200_EINVAL;
  }
  return 0;
}

int finish(const int *rc, int ndev, int *rcs) {
  int first = 0;
  for (int d = 0; d < ndev; ++d) {
    if (rcs) rcs[d] = rc[d];
    if (!first && rc[d]) first = rc[d];
  }
  return first;
}

}  // namespace

extern "C" {

int fssb200_eval_multi(const fssb200_ctx *const *ctxs, int ndev, int party, const void *const *seeds,
    const void *const *cws, const void *const *ocws, const void *const *xs, void *const *ys, const size_t *nkeys,
    void *const *streams, int *rcs) {
  if (int rc = check_ctxs(ctxs, ndev)) return rc;
  if (!seeds || !cws || !xs || !ys || !nkeys) return FSSB200_EINVAL;
  int rc[64];
  for (int d = 0; d < ndev; ++d)
    rc[d] = fssb200_eval(ctxs[d], party, seeds[d], cws[d], ocws ? ocws[d] : nullptr, xs[d], ys[d], nkeys[d],
        streams ? streams[d] : nullptr);
  return finish(rc, ndev, rcs);
}

int fssb200_eval_all_multi(const fssb200_ctx *const *ctxs, int ndev, int party, const void *const *seeds,
    const void *const *cws, const void *const *ocws, void *const *ys, const size_t *nkeys,
    const uint64_t *leaf_begin, const uint64_t *leaf_count, void *const *streams, int *rcs) {
  if (int rc = check_ctxs(ctxs, ndev)) return rc;
  if (!seeds || !cws || !ys || !nkeys) return FSSB200_EINVAL;
  int rc[64];
  for (int d = 0; d < ndev; ++d)
    rc[d] = fssb200_eval_all(ctxs[d], party, seeds[d], cws[d], ocws ? ocws[d] : nullptr, ys[d], nkeys[d],
        leaf_begin ? leaf_begin[d] : 0, leaf_count ? leaf_count[d] : 0, streams ? streams[d] : nullptr);
  return finish(rc, ndev, rcs);
}

int fssb200_gen_multi(const fssb200_ctx *const *ctxs, int ndev, const void *const *s0s, const void *const *alphas,
    const void *const *betas, void *const *cws, void *const *ocws, const size_t *nkeys, void *const *streams,
    int *rcs) {
  if (int rc = check_ctxs(ctxs, ndev)) return rc;
  if (!s0s || !alphas || !cws || !nkeys) return FSSB200_EINVAL;
  int rc[64];
  for (int d = 0; d < ndev; ++d)
    rc[d] = fssb200_gen(ctxs[d], s0s[d], alphas[d], betas ? betas[d] : nullptr, cws[d], ocws ? ocws[d] : nullptr,
        nkeys[d], streams ? streams[d] : nullptr);
  return finish(rc, ndev, rcs);
}

// Waits for the work queued on every device's stream; per-device cudaError_t in rcs.
int fssb200_multi_sync(const fssb200_ctx *const *ctxs, int ndev, void *const *streams, int *rcs) {
  if (!ctxs || ndev < 1 || ndev > 64) return FSSB200_EINVAL;
  int rc[64];
  for (int d = 0; d < ndev; ++d) {
    if (!ctxs[d]) return FSSB200_EINVAL;
    DeviceGuard g(ctxs[d]->p.device);
    rc[d] = g.err != cudaSuccess ? int(g.err)
                                 : int(cudaStreamSynchronize(static_cast<cudaStream_t>(streams ? streams[d] : nullptr)));
  }
  return finish(rc, ndev, rcs);
}

// Key range [begin, end) of shard `d` of `n` (contiguous, sizes differ by at most one): the split every multi-device
// call and the Python launcher (fss_b200/sharding.py: key_shard) use.
int fssb200_key_shard(size_t nkeys, int d, int n, size_t *begin, size_t *end) {
  if (n < 1 || d < 0 || d >= n || !begin || !end) return FSSB200_EINVAL;
  const size_t base = nkeys / size_t(n), rem = nkeys % size_t(n);
  *begin = size_t(d) * base + (size_t(d) < rem ? size_t(d) : rem);
  *end = *begin + base + (size_t(d) < rem ? 1 : 0);
  return 0;
}

// Leaf range of shard `d` of `n` for full-domain evaluation: whole work units (fssb200_eval_all_granule); trailing
// shards are empty when the domain has fewer units than shards (*count = 0 means EMPTY here, not "to the end").
int fssb200_leaf_shard(const fssb200_ctx *ctx, int d, int n, uint64_t *begin, uint64_t *count) {
  if (!ctx || !begin || !count) return FSSB200_EINVAL;
  const uint64_t g = fssb200_eval_all_granule(ctx);
  if (ctx->p.in_bits > 40) return FSSB200_EDOMAIN;
  const uint64_t units = (uint64_t(1) << ctx->p.in_bits) / g;
  size_t b = 0, e = 0;
  if (int rc = fssb200_key_shard(size_t(units), d, n, &b, &e)) return rc;
  *begin = uint64_t(b) * g;
  *count = uint64_t(e - b) * g;
  return 0;
}

}  // extern "C"
