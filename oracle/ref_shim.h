/* SPDX-License-Identifier: Apache-2.0
 * TEST INFRASTRUCTURE ONLY.  Shared declarations of the reference shim
 * (oracle/ref_shim.cpp) -- see that file's header. */
#ifndef REF_SHIM_H_
#define REF_SHIM_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { REF_SCHEME_DPF = 0, REF_SCHEME_DCF = 1, REF_SCHEME_HALFTREE = 2, REF_SCHEME_GROTTO = 3 };
/* 0/1 match the product ABI (include/fssb200.h); 2 = Aes128MmoRaw (AES-NI intrinsics),
 * 3 = Aes128Soft (ref_prg_gen only). */
enum { REF_PRG_AES128_MMO = 0, REF_PRG_CHACHA = 1, REF_PRG_AES128_MMO_RAW = 2 };

typedef struct RefParams {
  int32_t in_bytes;     /* width of one alpha / x element in the caller's arrays */
  uint8_t prg_key[64];  /* AES: mul x 16 B keys; ChaCha: 2 x int32 nonce          */
  uint8_t hash_key[16]; /* Half-Tree                                              */
} RefParams;

typedef struct RefOps {
  void (*gen)(const RefParams *, size_t nkeys, const void *s0s, const void *alphas,
              const void *betas, void *cws, void *ocws, int threads);
  void (*eval)(const RefParams *, int party, size_t nkeys, const void *seeds, const void *cws,
               const void *ocws, const void *xs, void *ys, int threads);
  void (*evalall)(const RefParams *, int party, size_t nkeys, const void *seeds, const void *cws,
                  const void *ocws, void *ys, int threads);
  void (*grotto_preprocess)(const RefParams *, int party, size_t nkeys, const void *seeds,
                            const void *cws, void *pt, int threads);
  void (*grotto_lookup)(const RefParams *, size_t nkeys, const void *pt, const void *xs, void *ys);
  int ncw;
} RefOps;

/* Which reference instantiation to run (compile-time template arguments in the
 * reference, dpf.cuh:61; selected at run time here). */
typedef struct RefSel {
  int32_t scheme, in_bits, group, prg, pred, pad;
  uint64_t mod_lo, mod_hi;
} RefSel;

/* Flat API of oracle/_ref/libfssref.so (oracle/ref_dispatch.cpp).  All return 0 on
 * success, -1 if the parameter set is not in the instantiation table.
 *   threads > 0 : that many OpenMP threads over keys, one PRG context set each
 *   threads = -1 (evalall only): keys serial, the reference's own par_depth = -1
 *                 OpenMP task recursion inside each key (dpf.cuh:242-246). */
int ref_supported(const RefSel *sel);
int ref_ncw(const RefSel *sel);
int ref_gen(const RefSel *sel, const RefParams *p, size_t nkeys, const void *s0s, const void *alphas,
            const void *betas, void *cws, void *ocws, int threads);
int ref_eval(const RefSel *sel, const RefParams *p, int party, size_t nkeys, const void *seeds,
             const void *cws, const void *ocws, const void *xs, void *ys, int threads);
int ref_evalall(const RefSel *sel, const RefParams *p, int party, size_t nkeys, const void *seeds,
                const void *cws, const void *ocws, void *ys, int threads);
int ref_grotto_preprocess(const RefSel *sel, const RefParams *p, int party, size_t nkeys,
                          const void *seeds, const void *cws, void *pt, int threads);
int ref_grotto_lookup(const RefSel *sel, const RefParams *p, size_t nkeys, const void *pt,
                      const void *xs, void *ys);
int ref_prg_gen(const RefParams *p, int prg, int mul, size_t n, const void *seeds, void *out);
int ref_host_threads(void);
int ref_table_size(void);

#ifdef __cplusplus
}
#endif
#endif
