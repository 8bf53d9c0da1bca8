/* SPDX-License-Identifier: Apache-2.0
 *
 * TEST INFRASTRUCTURE ONLY -- never linked into, imported by or called from the
 * product path (fss_b200/, include/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * fss_oracle: a plain-C CPU restatement of the reference's DPF / DCF / Half-Tree DPF
 * / Grotto DCF Gen, Eval and EvalAll with run-time parameters (the reference fixes
 * them as template arguments).  PARITY PINNED: tests/test_oracle.py checks every
 * function here bit-for-bit against oracle/_ref/libfssref.so (the unmodified
 * reference headers compiled by oracle/Makefile) and against the committed golden
 * fixtures in tests/golden/ that were generated from the reference
 * (oracle/make_golden.py), including the survey KATs of SURVEY.md section 8c.
 *
 * Parameters reuse `fssb200_params` from include/fssb200.h (`device` is ignored).
 */
#ifndef FSS_ORACLE_H_
#define FSS_ORACLE_H_
#include <stddef.h>
#include <stdint.h>
#include "../include/fssb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* All return 0 or a negative FSSB200_E* code.  `threads` = OpenMP threads over keys. */
int orc_prg_gen(const fssb200_params *p, int mul, size_t n, const void *seeds, void *out);
int orc_ncw(const fssb200_params *p);
int orc_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas,
            const void *betas, void *cws, void *ocws, int threads);
int orc_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
             const void *ocws, const void *xs, void *ys, int threads);
/* leaves [leaf_begin, leaf_begin + leaf_count) per key (leaf_count = 0: to 2^n).
 * Grotto writes 1 byte per leaf (prefix parity, leaf_begin must be 0). */
int orc_evalall(const fssb200_params *p, int party, size_t nkeys, const void *seeds,
                const void *cws, const void *ocws, void *ys, uint64_t leaf_begin,
                uint64_t leaf_count, int threads);
int orc_grotto_expand(const fssb200_params *p, int party, size_t nkeys, const void *seeds,
                      const void *cws, void *t, uint64_t leaf_begin, uint64_t leaf_count,
                      int threads);
int orc_grotto_preprocess(const fssb200_params *p, int party, size_t nkeys, const void *seeds,
                          const void *cws, void *pt, int threads);
int orc_grotto_lookup(const fssb200_params *p, size_t nkeys, const void *pt, const void *xs,
                      void *ys);
/* Level-major relayout restatement (point_eval_gpu.cuh:39-91), see fssb200_relayout. */
int orc_relayout(const fssb200_params *p, size_t nkeys, const void *cws, void *cw_s, void *cw_v,
                 void *extra, void *out_cw);
/* Group helper for reconstruction checks: out[i] = From(a[i]) + From(b[i]) -> Into. */
int orc_group_add(const fssb200_params *p, size_t n, const void *a, const void *b, void *out);

/* VDPF (vdpf.cuh) with XorHash = Hash = fss::hash::Blake3 (hash/blake3.cuh); see the fssb200_vdpf_* entry
 * points in include/fssb200.h for the array shapes.  which: 0 = XorHash (32 B -> 64 B), 1 = Hash (64 B -> 32 B). */
int orc_hash(const fssb200_params *p, int which, size_t n, const void *msgs, void *out);
int orc_vdpf_gen(const fssb200_params *p, size_t nkeys, const void *s0s, const void *alphas,
                 const void *betas, void *cws, void *cs, void *ocws, void *status, int threads);
int orc_vdpf_eval(const fssb200_params *p, int party, size_t nkeys, const void *seeds, const void *cws,
                  const void *cs, const void *ocws, const void *xs, void *ys, void *pis, int threads);
int orc_vdpf_prove(const fssb200_params *p, size_t nkeys, size_t m, const void *pi_tildes,
                   const void *cs, void *pis);
int orc_vdpf_evalall(const fssb200_params *p, int party, size_t nkeys, const void *seeds,
                     const void *cws, const void *cs, const void *ocws, void *ys, void *pis, int threads);

#ifdef __cplusplus
}
#endif
#endif
