// SPDX-License-Identifier: Apache-2.0
// TEST INFRASTRUCTURE ONLY.  Flat C entry points of oracle/_ref/libfssref.so:
// looks the parameter set up in the instantiation tables of the ref_shim.cpp parts.
#include "ref_shim.h"

#define DECL(k)                                                                                   \
  extern "C" const RefOps *ref_part##k##_lookup(int, int, int, uint64_t, uint64_t, int, int);     \
  extern "C" int ref_part##k##_count(void);
DECL(0) DECL(1) DECL(2) DECL(3) DECL(4)

static const RefOps *Find(const RefSel *s) {
  const RefOps *o;
#define TRY(k)                                                                                          \
  if ((o = ref_part##k##_lookup(s->scheme, s->in_bits, s->group, s->mod_lo, s->mod_hi, s->prg, s->pred))) \
    return o;
  TRY(0) TRY(1) TRY(2) TRY(3) TRY(4)
  return nullptr;
}

extern "C" {

int ref_supported(const RefSel *sel) { return Find(sel) != nullptr; }
int ref_ncw(const RefSel *sel) {
  auto *o = Find(sel);
  return o ? o->ncw : -1;
}
int ref_table_size(void) {
  return ref_part0_count() + ref_part1_count() + ref_part2_count() + ref_part3_count() + ref_part4_count();
}

int ref_gen(const RefSel *sel, const RefParams *p, size_t nkeys, const void *s0s, const void *alphas,
    const void *betas, void *cws, void *ocws, int threads) {
  auto *o = Find(sel);
  if (!o || !o->gen) return -1;
  o->gen(p, nkeys, s0s, alphas, betas, cws, ocws, threads);
  return 0;
}

int ref_eval(const RefSel *sel, const RefParams *p, int party, size_t nkeys, const void *seeds, const void *cws,
    const void *ocws, const void *xs, void *ys, int threads) {
  auto *o = Find(sel);
  if (!o || !o->eval) return -1;
  o->eval(p, party, nkeys, seeds, cws, ocws, xs, ys, threads);
  return 0;
}

int ref_evalall(const RefSel *sel, const RefParams *p, int party, size_t nkeys, const void *seeds,
    const void *cws, const void *ocws, void *ys, int threads) {
  auto *o = Find(sel);
  if (!o || !o->evalall) return -1;
  o->evalall(p, party, nkeys, seeds, cws, ocws, ys, threads);
  return 0;
}

int ref_grotto_preprocess(const RefSel *sel, const RefParams *p, int party, size_t nkeys, const void *seeds,
    const void *cws, void *pt, int threads) {
  auto *o = Find(sel);
  if (!o || !o->grotto_preprocess) return -1;
  o->grotto_preprocess(p, party, nkeys, seeds, cws, pt, threads);
  return 0;
}

int ref_grotto_lookup(const RefSel *sel, const RefParams *p, size_t nkeys, const void *pt, const void *xs,
    void *ys) {
  auto *o = Find(sel);
  if (!o || !o->grotto_lookup) return -1;
  o->grotto_lookup(p, nkeys, pt, xs, ys);
  return 0;
}

}  // extern "C"
