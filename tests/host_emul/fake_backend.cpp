// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY.  A CPU stand-in for libfssb200.so + the few CUDA runtime calls the header shim (include/fss/*.cuh)
// makes from its single-key members: every batch is answered by the oracle (oracle/fss_oracle.c, parity-pinned), device
// memory is heap memory.  It exists so that the HOST logic of the shim -- parameter marshalling of every scheme / group / PRG
// combination, and for the multi-point scheme the cuckoo table, bucket grouping, the gathers into one batch and the proof
// chains grouped by visit count -- is tested without a GPU (tests/test_vdmpf.py, tests/test_ref_gtests.py), by running the
// reference's own src/*_test.cu and samples, unmodified, on top of it.  The same sources linked with the real library run in
// the -m gpu tests.  Nothing in the product links this file.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../../include/fssb200.h"
#include "../../oracle/fss_oracle.h"

struct fssb200_ctx {
  fssb200_params p;
};

static long g_calls[3];  // gen / eval / prove batches (reported at exit: the shim must batch, not loop)
static void Report() {
  std::fprintf(stderr, "[fake backend] vdpf_gen_host batches: %ld, vdpf_eval_host batches: %ld, vdpf_prove batches: %ld\n", g_calls[0],
      g_calls[1], g_calls[2]);
}

extern "C" {
int fssb200_ctx_create(const fssb200_params *p, fssb200_ctx **out) {
  static bool once = (std::atexit(Report), true);
  (void)once;
  if (!p || !out) return FSSB200_EINVAL;
  *out = new fssb200_ctx{*p};
  return 0;
}
const char *fssb200_strerror(int rc) {
  static thread_local char buf[48];
  std::snprintf(buf, sizeof(buf), "fake backend: rc %d", rc);
  return buf;
}
int fssb200_vdpf_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *cs, void *ocws,
    void *status, size_t nkeys) {
  ++g_calls[0];
  return orc_vdpf_gen(&c->p, nkeys, s0s, alphas, betas, cws, cs, ocws, status, 4);
}
int fssb200_vdpf_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs, const void *ocws,
    const void *xs, void *ys, void *pis, size_t nkeys) {
  ++g_calls[1];
  return orc_vdpf_eval(&c->p, party, nkeys, seeds, cws, cs, ocws, xs, ys, pis, 4);
}
int fssb200_vdpf_prove(const fssb200_ctx *c, const void *pi_tildes, const void *cs, size_t m, void *pis, size_t nkeys, void *) {
  ++g_calls[2];
  return orc_vdpf_prove(&c->p, nkeys, m, pi_tildes, cs, pis);
}

int fssb200_vdpf_eval_all(const fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *cs, const void *ocws,
    void *ys, void *pis, size_t nkeys, void *) {
  return orc_vdpf_evalall(&c->p, party, nkeys, seeds, cws, cs, ocws, ys, pis, 4);
}
int fssb200_gen_host(fssb200_ctx *c, const void *s0s, const void *alphas, const void *betas, void *cws, void *ocws, size_t nkeys) {
  return orc_gen(&c->p, nkeys, s0s, alphas, betas, cws, ocws, 4);
}
int fssb200_eval_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws, const void *xs, void *ys,
    size_t nkeys) {
  return orc_eval(&c->p, party, nkeys, seeds, cws, ocws, xs, ys, 4);
}
int fssb200_eval_all_host(fssb200_ctx *c, int party, const void *seeds, const void *cws, const void *ocws, void *ys, size_t nkeys,
    uint64_t leaf_begin, uint64_t leaf_count) {
  return orc_evalall(&c->p, party, nkeys, seeds, cws, ocws, ys, leaf_begin, leaf_count, 4);
}
int fssb200_grotto_preprocess(const fssb200_ctx *c, int party, const void *seeds, const void *cws, void *pt, size_t nkeys, void *) {
  return orc_grotto_preprocess(&c->p, party, nkeys, seeds, cws, pt, 4);
}
int fssb200_grotto_eval(const fssb200_ctx *c, const void *pt, const void *xs, void *ys, size_t nkeys, void *) {
  return orc_grotto_lookup(&c->p, nkeys, pt, xs, ys);
}

// "device" memory
cudaError_t cudaGetDevice(int *d) {
  *d = 0;
  return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) {
  *p = std::malloc(n ? n : 1);
  return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) {
  std::free(p);
  return cudaSuccess;
}
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) {
  std::memcpy(dst, src, n);
  return cudaSuccess;
}
cudaError_t cudaMallocAsync(void **p, size_t n, cudaStream_t) { return cudaMalloc(p, n); }
cudaError_t cudaFreeAsync(void *p, cudaStream_t) { return cudaFree(p); }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(dst, src, n, k); }
}
