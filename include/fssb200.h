/* SPDX-License-Identifier: Apache-2.0
 *
 * fssb200.h -- C ABI of the B200-native batched evaluator for the DPF / DCF /
 * Half-Tree DPF / Grotto DCF PRG-tree hot path of myl7/fss.
 *
 * This header is the drop-in boundary: plain pointers and sizes, no C++ or torch
 * types.  Every entry point names the reference interface it replaces (paths are
 * relative to the reference checkout, `include/fss/...`).
 *
 * Memory layouts are the reference's, byte for byte:
 *   seed / output / group element : 16 B, CUDA `int4 {x,y,z,w}` little-endian words;
 *                                   bit 0 of `.w` is the clamp / control bit
 *                                   (util.cuh:30-38, group.cuh:28-34).
 *   Dpf::Cw        (dpf.cuh:76-81)          32 B {int4 s; bool tr; pad}      n+1 per key
 *   Dcf::Cw        (dcf.cuh:91-96)          32 B {int4 s; int4 v}            n+1 per key
 *   HalfTreeDpf::Cw(half_tree_dpf.cuh:53-57) 32 B {int4 s; bool extra; pad}  n   per key
 *   GrottoDcf::Cw  (grotto_dcf.cuh:50)      = Dpf::Cw                        n+1 per key
 *   keys are key-major: key k's correction words start at cws + k*ncw*32.
 *   inputs x / alpha : `In[nkeys]`, little-endian unsigned of `in_bytes` bytes
 *                      (1,2,4,8,16) -- the reference's `In` template parameter.
 *   EvalAll output : `int4 ys[nkeys][2^n]` natural order of x (dpf.cuh:291-301);
 *                    Grotto: `bool ys[nkeys][2^n]` one byte per leaf
 *                    (grotto_dcf.cuh:151-163).
 *
 * Ownership (dpf.cuh:86,225; eval_all_gpu.cuh:498): the caller allocates every
 * buffer; the library never allocates or frees device memory inside an eval /
 * gen call.  The `_host` variants stage through arenas of a process-wide pool
 * (see "host-buffer entry points").
 *
 * Errors: every function returns 0 on success, a negative FSSB200_E* code for an
 * invalid argument, or a positive `cudaError_t`.  Nothing throws or aborts.
 *
 * Streams: device-pointer entry points are stream ordered and never synchronise
 * the device; a context is immutable after creation and can be shared by threads
 * (eval_all_gpu.cuh:459-460 semantics).  `stream` is a `cudaStream_t` passed as
 * `void*` (NULL = default stream).
 */
#ifndef FSSB200_H_
#define FSSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FSSB200_VERSION 200 /* 0.2.0: re-entrant host entry points (arena pool), multi-device calls, Grotto walk */

/* ---- enums --------------------------------------------------------------- */

/* Scheme: which reference class template the context stands for. */
enum {
  FSSB200_SCHEME_DPF = 0,      /* fss::Dpf          dpf.cuh:61-304            */
  FSSB200_SCHEME_DCF = 1,      /* fss::Dcf          dcf.cuh:74-386            */
  FSSB200_SCHEME_HALFTREE = 2, /* fss::HalfTreeDpf  half_tree_dpf.cuh:39-355  */
  FSSB200_SCHEME_GROTTO = 3,   /* fss::GrottoDcf    grotto_dcf.cuh:45-239     */
  FSSB200_SCHEME_VDPF = 4      /* fss::Vdpf         vdpf.cuh:63-403 (XorHash / Hash = fss::hash::Blake3 or ::Sha256);
                                  only the fssb200_vdpf_* entry points apply          */
};

/* Output group (the `Group` template parameter, group.cuh:39-45). */
enum {
  FSSB200_GROUP_BYTES = 0, /* fss::group::Bytes             group/bytes.cuh:19-43 */
  FSSB200_GROUP_U8 = 1,    /* fss::group::Uint<uint8_t ,mod> group/uint.cuh:27-88 */
  FSSB200_GROUP_U16 = 2,   /* fss::group::Uint<uint16_t,mod>                      */
  FSSB200_GROUP_U32 = 3,   /* fss::group::Uint<uint32_t,mod>                      */
  FSSB200_GROUP_U64 = 4,   /* fss::group::Uint<uint64_t,mod>                      */
  FSSB200_GROUP_U128 = 5   /* fss::group::Uint<__uint128_t,mod>, 0 < mod <= 2^127 */
};

/* PRG (the `Prg` template parameter, prg.cuh:20-23). */
enum {
  FSSB200_PRG_AES128_MMO = 0, /* fss::prg::Aes128Mmo<mul> prg/aes128_mmo.cuh:27-94
                                 (== Aes128MmoRaw == Aes128Soft bit for bit)      */
  FSSB200_PRG_CHACHA = 1      /* fss::prg::ChaCha<mul,20> prg/chacha.cuh:24-128    */
};

/* DCF predicate (dcf.cuh:58-61); only Gen depends on it. */
enum { FSSB200_PRED_LT = 0, FSSB200_PRED_GT = 1 };
/* VDPF hash plugins (vdpf.cuh:55-58): fss::hash::Blake3 hash/blake3.cuh:24-172, fss::hash::Sha256 hash/sha256.cuh:25-90
 * (host-only in the reference; here both run on the device) */
enum { FSSB200_HASH_BLAKE3 = 0, FSSB200_HASH_SHA256 = 1 };

/* Error codes (negative).  Positive return values are cudaError_t. */
enum {
  FSSB200_OK = 0,
  FSSB200_EINVAL = -1,     /* NULL pointer / bad enum / bad size                  */
  FSSB200_EDOMAIN = -2,    /* in_bits outside [1, 8*in_bytes] or unsupported n    */
  FSSB200_EGROUP = -3,     /* group / modulus combination not representable       */
  FSSB200_ESCHEME = -4,    /* entry point does not apply to the context's scheme  */
  FSSB200_EALIGN = -5,     /* pointer not 16-byte aligned                         */
  FSSB200_ENODEVICE = -6,  /* no CUDA device / device index out of range          */
  FSSB200_ERANGE = -7,     /* leaf range not aligned / out of the domain          */
  FSSB200_ENOARENA = -8    /* (0.1 only: _host call without a reserved arena)     */
};

/* ---- context --------------------------------------------------------------- */

/* Everything the reference passes as template arguments or scheme members.
 *   in_bits   : `in_bits`                     (dpf.cuh:61)
 *   in_bytes  : sizeof(In)                    (dpf.cuh:61; fss_crypto/_jit.py:57-62)
 *   group,mod : `Group`; mod_hi:mod_lo is the 128-bit modulus, 0 = 2^(8*sizeof(T))
 *               (group/uint.cuh:27-31).  U128 requires 0 < mod <= 2^127.
 *   prg       : `Prg`; prg_key = mul 16-byte AES user keys, key i encrypts PRG
 *               output block i (prg/aes128_mmo.cuh:49-64,79-89) -- mul = 2 (DPF,
 *               Grotto), 4 (DCF), 1 (Half-Tree); or, for ChaCha, two little-endian
 *               int32 nonce words in prg_key[0..7] (prg/chacha.cuh:89-93,108-110).
 *   hash_key  : HalfTreeDpf::hash_key         (half_tree_dpf.cuh:44)
 *   pred      : DcfPred                       (dcf.cuh:58-61)
 *   device    : CUDA device ordinal the context's kernels run on.
 *   hash_iv   : VDPF only: key material of the two hash plugins, [0] = XorHash `H` (vdpf.cuh:55),
 *               [1] = Hash `H'` (:56): the 32-byte IV of a fss::hash::Blake3 (hash/blake3.cuh:131) or the
 *               16-byte key of a fss::hash::Sha256 (hash/sha256.cuh:35) in the first half of the slot.
 *   hash      : VDPF only: which function each plugin is, FSSB200_HASH_* in byte 0 (XorHash) and byte 1 (Hash).
 */
typedef struct fssb200_params {
  int32_t scheme;
  int32_t in_bits;
  int32_t in_bytes;
  int32_t group;
  uint64_t mod_lo;
  uint64_t mod_hi;
  int32_t prg;
  int32_t pred;
  uint8_t prg_key[64];
  uint8_t hash_key[16];
  int32_t device;
  int32_t hash;            /* VDPF hash plugins: byte 0 = XorHash (H), byte 1 = Hash (H'); FSSB200_HASH_* (0 = Blake3) */
  uint8_t hash_iv[2][32];  /* [0] XorHash, [1] Hash: Blake3 IV (32 B, hash/blake3.cuh:131) or SHA-256 key (first 16 B,
                              hash/sha256.cuh:35) */
} fssb200_params;

typedef struct fssb200_ctx fssb200_ctx;

int fssb200_version(void);
const char *fssb200_strerror(int code);

/* Replaces scheme-object construction `Dpf dpf{prg}` / `HalfTreeDpf{prg,hash_key}`
 * (dpf.cuh:64-66, README.md:125-126) and `Aes128Mmo::CreateCtxs`
 * (prg/aes128_mmo.cuh:49-64): expands the AES round keys once, validates the
 * parameter set.  Key material is copied; the caller may free `p` afterwards. */
int fssb200_ctx_create(const fssb200_params *p, fssb200_ctx **out);
/* Replaces `Aes128Mmo::FreeCtxs` (prg/aes128_mmo.cuh:66-70). */
void fssb200_ctx_destroy(fssb200_ctx *ctx);
int fssb200_ctx_params(const fssb200_ctx *ctx, fssb200_params *out);
/* Number of 32-byte Cw entries per key: n+1 (DPF, DCF, Grotto), n (Half-Tree, VDPF). */
int fssb200_ctx_ncw(const fssb200_ctx *ctx);

/* ---- batched key generation (device pointers) ------------------------------
 * ys-independent dealer step; replaces `Dpf::Gen` dpf.cuh:93-159, `Dcf::Gen`
 * dcf.cuh:108-194, `HalfTreeDpf::Gen` half_tree_dpf.cuh:68-175, `GrottoDcf::Gen`
 * grotto_dcf.cuh:63-67 (and the bench kernels src/bench_gpu.cu:72-83,142-153,
 * 212-223), one key per thread.
 *   s0s   : int4[nkeys][2]    party-0 and party-1 seeds
 *   alphas: In[nkeys]
 *   betas : int4[nkeys]       (ignored, may be NULL, for Grotto: beta = 0)
 *   cws   : Cw[nkeys][ncw]    out, key-major
 *   ocws  : int4[nkeys]       out, Half-Tree output correction word; else NULL
 */
int fssb200_gen(const fssb200_ctx *ctx, const void *s0s, const void *alphas, const void *betas,
                void *cws, void *ocws, size_t nkeys, void *stream);

/* ---- batched point evaluation (device pointers) ----------------------------
 * ys[k] = Eval(party, seeds[k], cws[k], xs[k]).
 *   fssb200_dpf_eval      replaces `Dpf::Eval` dpf.cuh:170-214 and
 *                         `fss::gpu::DpfEvalPointGpu` point_eval_gpu.cuh:448-460
 *                         (without needing DpfRelayoutGpu :346-353)
 *   fssb200_dcf_eval      replaces `Dcf::Eval` dcf.cuh:205-276 and
 *                         `fss::gpu::DcfEvalPointGpu` point_eval_gpu.cuh:480-492
 *   fssb200_halftree_eval replaces `HalfTreeDpf::Eval` half_tree_dpf.cuh:187-231 and
 *                         `fss::gpu::HalfTreeDpfEvalPointGpu` point_eval_gpu.cuh:416-428
 *   seeds : int4[nkeys]   the party's seed per key
 *   cws   : Cw[nkeys][ncw] key-major (the layout Gen writes)
 *   ocws  : int4[nkeys]   (Half-Tree only)
 *   xs    : In[nkeys]
 *   ys    : int4[nkeys]   out
 * fssb200_eval dispatches on the context's scheme (Grotto is EvalAll-only).
 */
int fssb200_eval(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                 const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream);
int fssb200_dpf_eval(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                     const void *xs, void *ys, size_t nkeys, void *stream);
int fssb200_dcf_eval(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                     const void *xs, void *ys, size_t nkeys, void *stream);
int fssb200_halftree_eval(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                          const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream);

/* ---- full-domain evaluation (device pointers) -------------------------------
 * Writes leaves [leaf_begin, leaf_begin+leaf_count) of every key:
 *   ys[k*leaf_count + (x-leaf_begin)] = Eval(party, seeds[k], cws[k], x).
 * leaf_begin/leaf_count select a subtree range so disjoint ranges can be sharded
 * across GPUs (BASELINE config 4); both must be multiples of the unit returned by
 * fssb200_eval_all_granule() (a power of two), leaf_count = 0 means "to 2^n".
 *   DPF      replaces `Dpf::EvalAll` dpf.cuh:232-303 and
 *            `fss::gpu::DpfEvalAllGpu[Batch]` eval_all_gpu.cuh:483-520
 *   DCF      replaces `Dcf::EvalAll` dcf.cuh:294-385 (no reference GPU version)
 *   HALFTREE replaces `HalfTreeDpf::EvalAll` half_tree_dpf.cuh:246-354 and
 *            `fss::gpu::HalfTreeDpfEvalAllGpu[Batch]` eval_all_gpu.cuh:451-535
 *   GROTTO   replaces `GrottoDcf::EvalAll` grotto_dcf.cuh:151-163: ys is
 *            `bool[nkeys][2^n]` (prefix parity of the leaf control bits; a
 *            sub-range is only allowed with leaf_begin = 0).
 */
int fssb200_eval_all(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                     const void *ocws, void *ys, size_t nkeys, uint64_t leaf_begin,
                     uint64_t leaf_count, void *stream);
uint64_t fssb200_eval_all_granule(const fssb200_ctx *ctx);

/* Grotto leaf control bits without the scan: t[k][x] (one byte per leaf); replaces
 * the private `GrottoDcf::ExpandTree` grotto_dcf.cuh:174-238. */
int fssb200_grotto_expand(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                          void *t, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count,
                          void *stream);
/* Replaces `GrottoDcf::Preprocess` grotto_dcf.cuh:94-104: pt[k] is the heap-ordered
 * parity tree of 2N-1 bytes (root p[0], leaf x at p[N-1+x]). */
int fssb200_grotto_preprocess(const fssb200_ctx *ctx, int party, const void *seeds,
                              const void *cws, void *pt, size_t nkeys, void *stream);
/* Replaces the static `GrottoDcf::Eval` lookup grotto_dcf.cuh:116-135:
 * ys[k] (1 byte) = prefix parity of pt[k] at xs[k]. */
int fssb200_grotto_eval(const fssb200_ctx *ctx, const void *pt, const void *xs, void *ys,
                        size_t nkeys, void *stream);

/* O(n) Grotto point evaluation for domains where the parity tree cannot exist (BASELINE configs[4]:
 * n = 32, 2^20 keys; the tree of grotto_dcf.cuh:78-81 would be 8 GiB per key -- SURVEY.md H6):
 * ys[k] (1 byte) = XOR of the control bits of the left-sibling subtree roots along the path of
 * e = xs[k] + 1, i.e. of the disjoint subtrees that tile [0, e).  ys0[k] ^ ys1[k] = 1[alpha_k <= xs[k]]
 * exactly like `GrottoDcf::Eval` (grotto_dcf.cuh:116-135), with the same edge rule (e == 0 or
 * e == N: the whole domain, share = party), but the per-party bit is NOT the reference's: that one
 * is the parity of N pseudorandom leaf bits and cannot be had without expanding the tree.
 * "Reconstruction-equal, share-parity unpinned": use fssb200_grotto_preprocess + fssb200_grotto_eval
 * where per-share bit-exactness with the reference matters (n <= 31, 2N-1 bytes per key).
 *   seeds : int4[nkeys]   cws : Cw[nkeys][n+1] (Dpf::Cw)   xs : In[nkeys]   ys : uint8[nkeys]  */
int fssb200_grotto_eval_walk(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                             const void *xs, void *ys, size_t nkeys, void *stream);

/* ---- VDPF (verifiable DPF) -----------------------------
 * Context scheme FSSB200_SCHEME_VDPF.  Key of party i = cws (n entries of Dpf-style
 * 32-byte Cw, vdpf.cuh:77-80) + cs (4 x int4 correction seed) + ocw + s0s[i].
 *   fssb200_vdpf_gen      replaces `Vdpf::Gen` vdpf.cuh:97-177 (and VdpfGenKernel,
 *                         src/bench_gpu.cu:173-186).  status[k] = Gen's return value:
 *                         1 = t0 == t1 at the end, the caller resamples the seeds (:169);
 *                         cws / cs of such a key are still written, ocw is not.
 *   fssb200_vdpf_eval     replaces `Vdpf::Eval` vdpf.cuh:191-243 and
 *                         `fss::gpu::VdpfEvalPointGpu` point_eval_gpu.cuh:514-527:
 *                         ys[k] = y share, pis[k] = corrected per-point hash (4 x int4).
 *   fssb200_vdpf_prove    replaces `Vdpf::Prove` vdpf.cuh:254-264 for a batch of keys:
 *                         key k accumulates its m hashes pi_tildes[k][0..m) in order.
 *   fssb200_vdpf_eval_all replaces `Vdpf::EvalAll` vdpf.cuh:294-342 (the reference has no
 *                         GPU version): ys[k][x] for the whole domain and the accumulated
 *                         proof pis[k].  The proof chain is sequential in x by definition
 *                         (:336-340); 1 ... 32 lanes walk it per key (chosen from nkeys: many
 *                         keys share a warp, few keys get a warp each).
 * Verify (vdpf.cuh:271-276) is a 64-byte comparison of two proofs; no entry point.
 *   cws : Cw[nkeys][n]      cs : int4[nkeys][4]      ocws : int4[nkeys]
 *   pis / pi_tildes : int4[nkeys][4] / int4[nkeys][m][4]      status : int32[nkeys]
 */
int fssb200_vdpf_gen(const fssb200_ctx *ctx, const void *s0s, const void *alphas, const void *betas,
                     void *cws, void *cs, void *ocws, void *status, size_t nkeys, void *stream);
int fssb200_vdpf_eval(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                      const void *cs, const void *ocws, const void *xs, void *ys, void *pis,
                      size_t nkeys, void *stream);
/* fssb200_vdpf_eval on the level-major arrays of fssb200_relayout (cw_s[n][nkeys], extra); replaces
 * `fss::gpu::VdpfRelayoutGpu` + `VdpfEvalPointGpu` point_eval_gpu.cuh:390-397,514-527. */
int fssb200_vdpf_eval_levelmajor(const fssb200_ctx *ctx, int party, const void *seeds, const void *cw_s,
                                 const void *extra, const void *cs, const void *ocws, const void *xs,
                                 void *ys, void *pis, size_t nkeys, void *stream);
int fssb200_vdpf_prove(const fssb200_ctx *ctx, const void *pi_tildes, const void *cs, size_t m,
                       void *pis, size_t nkeys, void *stream);
int fssb200_vdpf_eval_all(const fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                          const void *cs, const void *ocws, void *ys, void *pis, size_t nkeys,
                          void *stream);
/* Hash known-answer hook (hash/blake3.cuh:143-171, hash/sha256.cuh:44-89; the context's plugins): which = 0: XorHash, msgs = (a, b) pairs of
 * int4 (32 B) -> 64 B each; which = 1: Hash, msgs = 64 B -> 32 B each.  Device pointers. */
int fssb200_hash(const fssb200_ctx *ctx, int which, const void *msgs, void *out, size_t n,
                 void *stream);

/* ---- level-major layout (optional pre-pass) ----------------------------------
 * Replaces `fss::gpu::{Dpf,Dcf,HalfTreeDpf}RelayoutGpu` point_eval_gpu.cuh:324-381:
 * key-major Cw[nkeys][ncw] -> the compact level-major layout the reference's point
 * kernels read (cw_s[i*nkeys+k], DCF also cw_v[i*nkeys+k], packed tr bits, out_cw).
 * Unlike the reference (`uint32_t extra`, n <= 32) the packed control bits are
 * ceil(n/32) words per key: extra[w*nkeys+k] bit j = level 32*w+j.
 *   cw_s  : int4[n][nkeys]          out
 *   cw_v  : int4[n][nkeys]          out (DCF only, else NULL)
 *   extra : uint32[ceil(n/32)][nkeys] out (bit i = the bool at byte 16 of Cw[i]: DPF tr bits;
 *                                     Half-Tree: only bit n-1, the last level's `extra`; DCF: NULL)
 *   out_cw: int4[nkeys]             out (DPF: cws[n].s, DCF: cws[n].v; Half-Tree: unused)
 */
int fssb200_relayout(const fssb200_ctx *ctx, const void *cws, void *cw_s, void *cw_v, void *extra,
                     void *out_cw, size_t nkeys, void *stream);
/* Replaces `fss::gpu::{Dpf,Dcf,HalfTreeDpf}EvalPointGpu` point_eval_gpu.cuh:416-492
 * on the layout above (ocws: Half-Tree only).  The arrays reach shared memory through the TMA
 * unit (2-D tensors [n][nkeys*4] of uint32, 4 levels x 32 keys per request); cw_s / cw_v must
 * be 16-byte aligned. */
int fssb200_eval_levelmajor(const fssb200_ctx *ctx, int party, const void *seeds, const void *cw_s,
                            const void *cw_v, const void *extra, const void *out_cw,
                            const void *ocws, const void *xs, void *ys, size_t nkeys,
                            void *stream);

/* ---- host-buffer entry points (what a CPU caller of the reference binds) -----
 * Same semantics with HOST pointers (pageable or pinned): inputs are staged to the
 * device in chunks, evaluated, and results copied back, with copies and kernels
 * overlapped on internal streams.  The call returns when the outputs are complete.
 *
 * Re-entrant and thread-safe like the reference's members (dpf.cuh:170 is a const
 * pure function; src/bench_cpu.cu:157-161 calls it under `#pragma omp parallel for`):
 * a context owns no staging memory.  Each call checks an arena (device staging, pinned
 * staging, streams, events) out of a process-wide per-device pool, sized from ITS batch
 * (a 1-key call takes 1 MiB), and hands it back when it returns -- also on every error
 * path, after its streams are drained.  The pool keeps up to 4 idle arenas per device for
 * reuse; fssb200_host_trim() frees them.  fssb200_ctx_reserve_host() is kept from 0.1:
 * it only records the preferred keys per pipeline chunk (0 = library default) and never
 * allocates; no call returns FSSB200_ENOARENA any more.
 *
 * fssb200_eval_host, batches of >= 8192 keys on a host with >= 2 usable cores per rank
 * (cores / LOCAL_WORLD_SIZE; FSSB200_PACK_THREADS overrides): adaptive pack / direct
 * pipeline.  For DPF / Half-Tree keys worker threads strip the 15 padding bytes of every
 * 32-byte Cw into pinned staging ("packed rows" below), chunk after chunk from the front
 * of the batch, and the calling thread submits each finished chunk as one H2D copy +
 * fssb200_eval_packed + D2H; whenever the link is about to run dry and no packed chunk is
 * ready, a chunk from the back of the batch crosses in the reference layout straight from
 * the caller's (pinned) buffer.  Pageable inputs are always staged by the workers (packed,
 * or copied for schemes without padding).  fssb200_ctx_set_host_mode(): 0 = automatic (default:
 * the adaptive pipeline with one or two ranks per host; from three ranks on the links together
 * ask for more than the host memory serves and packing only takes bandwidth from the copy
 * engines, so the rows cross as they are), 1 = reference layout only, 2 = staged chunks only,
 * 3 = adaptive pipeline whatever the rank count.  A call that finds no idle worker thread stages
 * with the calling thread alone. */
int fssb200_ctx_reserve_host(fssb200_ctx *ctx, size_t max_keys_per_chunk);
int fssb200_ctx_set_host_mode(fssb200_ctx *ctx, int mode);
/* Keys of this context's last fssb200_eval_host call that crossed the link staged by the host
 * threads (packed rows; plain copies for schemes without padding) / in the reference layout
 * straight from the caller's buffers, and the host threads that call used (0: plain chunked path). */
int fssb200_ctx_host_stats(const fssb200_ctx *ctx, uint64_t *packed_keys, uint64_t *direct_keys,
                           int *threads);
/* Frees the idle arenas of the pool; bytes they hold right now. */
void fssb200_host_trim(void);
int fssb200_host_cached_bytes(uint64_t *device_bytes, uint64_t *pinned_bytes);
int fssb200_eval_host(fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                      const void *ocws, const void *xs, void *ys, size_t nkeys);
/* ys = [nkeys][leaf_count] in host memory.  Sets of 64 MiB of leaves (FSSB200_ALL_SET_MB): a small
 * domain's keys go as many whole keys per launch and copy as fit one set (at most the keys per chunk
 * of fssb200_ctx_reserve_host), a large domain's one key at a time in leaf ranges. */
int fssb200_eval_all_host(fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                          const void *ocws, void *ys, size_t nkeys, uint64_t leaf_begin,
                          uint64_t leaf_count);
int fssb200_gen_host(fssb200_ctx *ctx, const void *s0s, const void *alphas, const void *betas,
                     void *cws, void *ocws, size_t nkeys);
/* fssb200_eval_levelmajor with HOST arrays in the level-major layout of fssb200_relayout (the compact
 * key format: 16 B + 1 bit per level instead of the 32-byte Cw): chunks of
 * keys are gathered from the [level][key] arrays with strided copies.  Same NULL rules as
 * fssb200_eval_levelmajor. */
/* VDPF with host arrays (same chunked pipeline). */
int fssb200_vdpf_gen_host(fssb200_ctx *ctx, const void *s0s, const void *alphas, const void *betas,
                          void *cws, void *cs, void *ocws, void *status, size_t nkeys);
int fssb200_vdpf_eval_host(fssb200_ctx *ctx, int party, const void *seeds, const void *cws,
                           const void *cs, const void *ocws, const void *xs, void *ys, void *pis,
                           size_t nkeys);
int fssb200_eval_levelmajor_host(fssb200_ctx *ctx, int party, const void *seeds, const void *cw_s,
                                 const void *cw_v, const void *extra, const void *out_cw,
                                 const void *ocws, const void *xs, void *ys, size_t nkeys);

/* ---- multi-device entry points (one process, ndev GPUs) ----------
 * The path shards with no exchange step: keys are independent (dpf.cuh:170-214) and an EvalAll
 * subtree depends only on its root (dpf.cuh:291-301).  A multi-device call is ndev independent,
 * stream-ordered launches; nothing moves between devices, there is no collective (the reference
 * has no multi-GPU code at all).  ctxs[d] is a context created for device d' = its params.device
 * with the SAME parameter set and key material; every array argument is an array of ndev
 * per-device pointers (device memory of ctxs[d]'s device); streams[d] is a stream of that device
 * (streams == NULL: default streams).  Returns the first non-zero per-device code; rcs (may be
 * NULL) receives all ndev codes.  The calls do not synchronise: fssb200_multi_sync() waits for
 * every device's stream and surfaces each device's cudaError_t.
 *   fssb200_eval_multi      device d: ys[d][k] = Eval(party, seeds[d][k], cws[d][k], xs[d][k]), k < nkeys[d]
 *   fssb200_eval_all_multi  device d: leaves [leaf_begin[d], +leaf_count[d]) of its nkeys[d] keys
 *                           (BASELINE configs[3] "subtrees sharded": same keys on every device,
 *                           ranges from fssb200_leaf_shard; or a key range per device)
 *   fssb200_gen_multi       device d: Gen of its nkeys[d] keys
 * fssb200_key_shard / fssb200_leaf_shard: the contiguous split (sizes differ by at most one key /
 * work unit) the multi-device calls, bench.py and fss_b200/sharding.py use. */
int fssb200_eval_multi(const fssb200_ctx *const *ctxs, int ndev, int party, const void *const *seeds,
                       const void *const *cws, const void *const *ocws, const void *const *xs,
                       void *const *ys, const size_t *nkeys, void *const *streams, int *rcs);
int fssb200_eval_all_multi(const fssb200_ctx *const *ctxs, int ndev, int party,
                           const void *const *seeds, const void *const *cws, const void *const *ocws,
                           void *const *ys, const size_t *nkeys, const uint64_t *leaf_begin,
                           const uint64_t *leaf_count, void *const *streams, int *rcs);
int fssb200_gen_multi(const fssb200_ctx *const *ctxs, int ndev, const void *const *s0s,
                      const void *const *alphas, const void *const *betas, void *const *cws,
                      void *const *ocws, const size_t *nkeys, void *const *streams, int *rcs);
int fssb200_multi_sync(const fssb200_ctx *const *ctxs, int ndev, void *const *streams, int *rcs);
int fssb200_key_shard(size_t nkeys, int d, int n, size_t *begin, size_t *end);
int fssb200_leaf_shard(const fssb200_ctx *ctx, int d, int n, uint64_t *begin, uint64_t *count);
/* Host arrays of the WHOLE batch, evaluated on ndev GPUs.  Returns when ys is complete.
 *  - 1-2 devices (host modes 0 / 2 / 3): device d takes key range fssb200_key_shard(nkeys, d, ndev);
 *    one host thread per device runs the fssb200_eval_host pipeline on its range with its share of
 *    the worker threads.
 *  - from 3 devices on, or in host mode 1 (keys cross in the reference layout): the links of one
 *    host are not equally fast (22.8 ... 34.4 GB/s at 8 GPUs), so the devices CLAIM key blocks
 *    (2^15 ... 2^17 keys, shrinking towards the end) from one counter, two calls in flight per
 *    device; which device evaluates a key is not fixed.  8 GPUs: 167.5 ms per 8 x 2^22 keys against
 *    205.0 ms for the equal split (FSSB200_MULTI_BALANCE=0 keeps the equal split).
 * rcs[d] (may be NULL) = the first error device d saw; the return value = the first non-zero one. */
int fssb200_eval_host_multi(fssb200_ctx *const *ctxs, int ndev, int party, const void *seeds,
                            const void *cws, const void *ocws, const void *xs, void *ys, size_t nkeys,
                            int *rcs);

/* ---- packed rows: compact key format for DPF / Half-Tree keys ---------------------------
 * Dpf::Cw (dpf.cuh:76-81) and HalfTreeDpf::Cw (half_tree_dpf.cuh:53-57) are {int4 s; bool flag}
 * padded to 32 bytes: 15 of every 32 bytes carry nothing.  A packed row holds the ncw 16-byte
 * `s` entries of a key followed by one 16-byte word of flag bits (bit i = the bool at byte 16
 * of entry i, i < 128): fssb200_packed_row_bytes() = ncw*16 + 16 (0 for schemes whose Cw has
 * no padding).  fssb200_pack_rows() converts HOST arrays (reference layout in, packed rows
 * out; multi-threaded on the library's host worker threads when they are idle) --
 * a format conversion, no evaluation happens on the CPU.  fssb200_eval_packed() is
 * fssb200_eval() on packed rows in device memory (rows fetched by the TMA unit, four levels per
 * 64-byte chunk).  fssb200_eval_host() uses both internally (adaptive pipeline above). */
size_t fssb200_packed_row_bytes(const fssb200_ctx *ctx);
/* Threads a large fssb200_eval_host() call of this process stages rows with, the calling
 * thread included (0: scheme without padding, or fewer than 2 cores per rank). */
int fssb200_ctx_host_pack_threads(const fssb200_ctx *ctx);
int fssb200_pack_rows(const fssb200_ctx *ctx, const void *cws, void *rows, size_t nkeys);
int fssb200_eval_packed(const fssb200_ctx *ctx, int party, const void *seeds, const void *rows,
                        const void *ocws, const void *xs, void *ys, size_t nkeys, void *stream);

/* ---- introspection / measurement helpers -------------------------------------- */

/* PRG known-answer hook: out[i] = block i of prg.Gen(seed), i < mul, for nseeds
 * seeds (device pointers).  Replaces a direct `prg.Gen(seed)` call
 * (prg/aes128_mmo.cuh:72-93, prg/chacha.cuh:95-127). */
int fssb200_prg_gen(const fssb200_ctx *ctx, const void *seeds, void *out, int mul, size_t nseeds,
                    void *stream);
/* The same with HOST arrays (arena pool; re-entrant): the `prg.Gen(seed)` member of the C++ shim. */
int fssb200_prg_gen_host(fssb200_ctx *ctx, const void *seeds, void *out, int mul, size_t nseeds);
/* Number of kernel launches this context has issued (bench.py's gpu_launches). */
uint64_t fssb200_ctx_launch_count(const fssb200_ctx *ctx);
/* Integer-pipe / shared-memory issue-rate microbenchmarks used for the roofline
 * denominator.  kind: 0 = LOP3 chain, 1 = IMAD chain, 2 = LOP3+IMAD
 * mixed, 3 = conflict-free LDS.32, 4 = PRMT, 5 = IDP.4A, 6 = PRMT+IDP.4A mixed; lookup-path probes (per-lane
 * data-dependent index into a 256-entry table): 7 = texture fetches, 8 / 9 = 8 LDS + 4 / 2 texture fetches, 10 = cached
 * read-only global loads, 11 = 8 LDS + 2 global loads, 12 = 8 LDS.  Returns ops (or lookups) per second through
 * *ops_per_s. */
int fssb200_microbench(int device, int kind, double *ops_per_s);

#ifdef __cplusplus
}
#endif
#endif /* FSSB200_H_ */
