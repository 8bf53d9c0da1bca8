// SPDX-License-Identifier: Apache-2.0
// fss/b200/generic.cuh -- the plugin-generic form of the hot path: DPF / DCF / Half-Tree / VDPF key generation and point
// evaluation written ONLY against the reference's plugin concepts
//     Groupable  (group.cuh:39-45):  default ctor = zero, a + b, -a, Group::From(int4), a.Into()
//     Prgable    (prg.cuh:20-23):    prg.Gen(int4) -> cuda::std::array<int4, mul>
// so that ANY user-defined Group / Prg that satisfies them works, both
//   * per thread inside the user's own __global__ kernels -- the reference's members are `__host__ __device__` and
//     documented for that use (README.md:198-242, samples/dpf_dcf_gpu.cu:51-82), and
//   * batched from the host: the kernels at the bottom of this header are instantiated in the user's translation unit
//     (nvcc) with the user's types and launched by the scheme classes' members.
// The precompiled sm_100a kernels of libfssb200.so (lane-replicated T-table AES, TMA-staged correction words, ...)
// remain the fast path for the built-in plugins; this header is the extensibility path and trades speed for
// generality: one key per thread, every PRG output block computed, correction words read from global memory.
//
// A tree node is one packed block: the seed with its control bit t in the clamp bit.  Because a stored correction
// word carries tl_cw in its clamp bit (dpf.cuh:148), "if (t) { s ^= s_cw; t ^= t_cw }" (dpf.cuh:189-194) is one masked
// XOR of the packed block with cw' = s_cw whose clamp bit is tl_cw (left child) or tr_cw (right child).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <cuda/std/array>
#include <cuda/std/tuple>
#include <fss/group.cuh>
#include <fss/prg.cuh>
#include <fss/util.cuh>

namespace fss::b200::generic {

FSS_SHIM_HD int4 Masked(int4 a, bool on, int4 b) {  // a ^ (on ? b : 0)
  const int m = on ? -1 : 0;
  return int4{a.x ^ (m & b.x), a.y ^ (m & b.y), a.z ^ (m & b.z), a.w ^ (m & b.w)};
}
FSS_SHIM_HD int4 Clamp(int4 v) { return util::SetLsb(v, false); }
template <typename In>
FSS_SHIM_HD bool BitMsbFirst(In v, int in_bits, int level) {  // dpf.cuh:196
  return ((v >> (in_bits - 1 - level)) & 1) != 0;
}

// ---- DPF ------------------------------------------------------------------------------------------------------------
// Dpf::Eval, dpf.cuh:170-214.
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD int4 DpfEval(Prg &prg, bool b, int4 s0, const Cw cws[], In x) {
  int4 node = util::SetLsb(s0, b);  // seed | t, t = b
  for (int i = 0; i < in_bits; ++i) {
    const bool t = util::GetLsb(node);
    const bool right = BitMsbFirst(x, in_bits, i);
    auto g = prg.Gen(Clamp(node));
    const int4 cw = util::SetLsb(cws[i].s, right ? bool(cws[i].tr) : util::GetLsb(cws[i].s));
    node = Masked(right ? g[1] : g[0], t, cw);
  }
  Group y = Group::From(Clamp(node));
  if (util::GetLsb(node)) y = y + Group::From(cws[in_bits].s);
  if (b) y = -y;
  return y.Into();
}

// Dpf::Gen, dpf.cuh:93-159.  Bytes 16..31 of every Cw are written ({tr, 0, ...}, :151-153; entry n: zero).
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD void DpfGen(Prg &prg, Cw cws[], const int4 s0s[2], In a, int4 b_buf) {
  static_assert(sizeof(Cw) == 32, "Dpf::Cw is {int4 s; bool tr} padded to 32 bytes (dpf.cuh:76-81)");
  int4 n0 = Clamp(s0s[0]);                       // party 0: t = 0
  int4 n1 = util::SetLsb(s0s[1], true);          // party 1: t = 1
  b_buf = Clamp(b_buf);
  for (int i = 0; i < in_bits; ++i) {
    auto g0 = prg.Gen(Clamp(n0));
    auto g1 = prg.Gen(Clamp(n1));
    const bool right = BitMsbFirst(a, in_bits, i);  // alpha's path keeps this child, the other one is "lost"
    const int4 lose0 = right ? g0[0] : g0[1], lose1 = right ? g1[0] : g1[1];
    const int4 keep0 = right ? g0[1] : g0[0], keep1 = right ? g1[1] : g1[0];
    const bool tl_cw = util::GetLsb(g0[0]) ^ util::GetLsb(g1[0]) ^ right ^ true;  // :119
    const bool tr_cw = util::GetLsb(g0[1]) ^ util::GetLsb(g1[1]) ^ right;          // :120
    const int4 s_cw = Clamp(util::Xor(lose0, lose1));                              // :115-117
    const int4 cw_keep = util::SetLsb(s_cw, right ? tr_cw : tl_cw);
    n0 = Masked(keep0, util::GetLsb(n0), cw_keep);
    n1 = Masked(keep1, util::GetLsb(n1), cw_keep);
    int4 *raw = reinterpret_cast<int4 *>(&cws[i]);
    raw[0] = util::SetLsb(s_cw, tl_cw);
    raw[1] = int4{tr_cw ? 1 : 0, 0, 0, 0};
  }
  Group v = Group::From(b_buf) + (-Group::From(Clamp(n0))) + Group::From(Clamp(n1));  // :155
  if (util::GetLsb(n1)) v = -v;                                                      // :156-157
  int4 *raw = reinterpret_cast<int4 *>(&cws[in_bits]);
  raw[0] = v.Into();
  raw[1] = int4{0, 0, 0, 0};
}

// ---- DCF ------------------------------------------------------------------------------------------------------------
// Dcf::Eval, dcf.cuh:205-276.  PRG blocks: {s_l, v_l, s_r, v_r} (:223); tl_cw = lsb(cw.s), tr_cw = lsb(cw.v).
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD int4 DcfEval(Prg &prg, bool b, int4 s0, const Cw cws[], In x) {
  int4 node = util::SetLsb(s0, b);
  Group v;  // zero
  for (int i = 0; i < in_bits; ++i) {
    const bool t = util::GetLsb(node);
    const bool right = BitMsbFirst(x, in_bits, i);
    auto g = prg.Gen(Clamp(node));
    const Group v_side = Group::From(Clamp(right ? g[3] : g[1]));
    const Group v_cw = Group::From(Clamp(cws[i].v));
    // v += sign * (v_side + (t ? v_cw : 0)), t of the CURRENT node (dcf.cuh:244-252)
    if (b) {
      v = v + (-v_side);
      if (t) v = v + (-v_cw);
    } else {
      v = v + v_side;
      if (t) v = v + v_cw;
    }
    const int4 cw = util::SetLsb(cws[i].s, right ? util::GetLsb(cws[i].v) : util::GetLsb(cws[i].s));
    node = Masked(right ? g[2] : g[0], t, cw);
  }
  const Group last = Group::From(Clamp(node)), v_np1 = Group::From(cws[in_bits].v);   // :263-275
  if (b) {
    v = v + (-last);
    if (util::GetLsb(node)) v = v + (-v_np1);
  } else {
    v = v + last;
    if (util::GetLsb(node)) v = v + v_np1;
  }
  return v.Into();
}

// Dcf::Gen, dcf.cuh:108-194.  `lt`: DcfPred::kLt (beta is added on the levels where alpha's bit is 1), else kGt.
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD void DcfGen(Prg &prg, bool lt, Cw cws[], const int4 s0s[2], In a, int4 b_buf) {
  int4 n0 = Clamp(s0s[0]);
  int4 n1 = util::SetLsb(s0s[1], true);
  Group v;  // zero
  const Group beta = Group::From(Clamp(b_buf));
  for (int i = 0; i < in_bits; ++i) {
    auto g0 = prg.Gen(Clamp(n0));
    auto g1 = prg.Gen(Clamp(n1));
    const bool right = BitMsbFirst(a, in_bits, i);
    const bool t1 = util::GetLsb(n1);
    const int ks = right ? 2 : 0, ls = right ? 0 : 2;   // block index of the kept / lost side's seed; value = +1
    const Group v0_lose = Group::From(Clamp(g0[ls + 1])), v1_lose = Group::From(Clamp(g1[ls + 1]));
    const Group v0_keep = Group::From(Clamp(g0[ks + 1])), v1_keep = Group::From(Clamp(g1[ks + 1]));
    Group v_cw = (-v) + v1_lose + (-v0_lose);           // :147-153
    if (right == lt) v_cw = v_cw + beta;
    if (t1) v_cw = -v_cw;                               // :155
    v = v + (-v1_keep) + v0_keep;                       // :157-160
    v = t1 ? v + (-v_cw) : v + v_cw;
    const bool tl_cw = util::GetLsb(g0[0]) ^ util::GetLsb(g1[0]) ^ right ^ true;
    const bool tr_cw = util::GetLsb(g0[2]) ^ util::GetLsb(g1[2]) ^ right;
    const int4 s_cw = Clamp(util::Xor(g0[ls], g1[ls]));
    const int4 cw_keep = util::SetLsb(s_cw, right ? tr_cw : tl_cw);
    n0 = Masked(g0[ks], util::GetLsb(n0), cw_keep);
    n1 = Masked(g1[ks], t1, cw_keep);
    cws[i].s = util::SetLsb(s_cw, tl_cw);
    cws[i].v = util::SetLsb(v_cw.Into(), tr_cw);        // :187-189
  }
  Group v_np1 = Group::From(Clamp(n1)) + (-Group::From(Clamp(n0))) + (-v);  // :191
  if (util::GetLsb(n1)) v_np1 = -v_np1;
  cws[in_bits].s = int4{0, 0, 0, 0};
  cws[in_bits].v = v_np1.Into();
}

// ---- Half-Tree DPF ------------------------------------------------------------------------------------------------------
// H(node) = prg.Gen(hash_key ^ node)[0] with the control bit t = lsb(node) INSIDE the hashed value (half_tree_dpf.cuh:195).
template <typename Prg>
FSS_SHIM_HD int4 HtHash(Prg &prg, int4 hash_key, int4 node) { return prg.Gen(util::Xor(hash_key, node))[0]; }

// HalfTreeDpf::Eval, half_tree_dpf.cuh:187-231: n hashes per evaluation.
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD int4 HalfTreeEval(Prg &prg, int4 hash_key, bool b, int4 s0, const Cw cws[], int4 ocw, In x) {
  int4 node = util::SetLsb(s0, b);
  for (int i = 0; i < in_bits - 1; ++i) {
    const bool t = util::GetLsb(node), right = BitMsbFirst(x, in_bits, i);
    const int4 h = HtHash(prg, hash_key, node);
    node = Masked(Masked(h, right, node), t, cws[i].s);                  // :202-204 (the cw's lsb included)
  }
  const bool sigma = (x & 1) != 0, t = util::GetLsb(node);
  const Cw &last = cws[in_bits - 1];
  int4 h = HtHash(prg, hash_key, util::SetLsb(node, sigma));              // :208-216
  h = Masked(h, t, util::SetLsb(last.s, sigma ? bool(last.extra) : util::GetLsb(last.s)));   // high ^= HCW, low ^= LCW_sigma
  Group y = Group::From(Clamp(h));
  if (util::GetLsb(h)) y = y + Group::From(ocw);
  if (b) y = -y;
  return y.Into();
}

// HalfTreeDpf::Gen, half_tree_dpf.cuh:68-175.  Levels 0..n-2: only `.s` and `.extra = false` are defined by the
// reference (aggregate assignment, :91); both 16-byte halves are written here, padding zero.
template <int in_bits, typename Group, typename In, typename Prg, typename Cw>
FSS_SHIM_HD void HalfTreeGen(Prg &prg, int4 hash_key, Cw cws[], int4 &ocw, const int4 s0s[2], In a, int4 b_buf) {
  static_assert(sizeof(Cw) == 32, "HalfTreeDpf::Cw is {int4 s; bool extra} padded to 32 bytes (half_tree_dpf.cuh:53-57)");
  int4 n0 = Clamp(s0s[0]), n1 = util::SetLsb(s0s[1], true);
  for (int i = 0; i < in_bits - 1; ++i) {
    const int4 h0 = HtHash(prg, hash_key, n0), h1 = HtHash(prg, hash_key, n1);
    const bool right = BitMsbFirst(a, in_bits, i);
    const int4 cw = Masked(util::Xor(h0, h1), !right, util::Xor(n0, n1));   // :83-89
    int4 *raw = reinterpret_cast<int4 *>(&cws[i]);
    raw[0] = cw;
    raw[1] = int4{0, 0, 0, 0};
    const bool t0 = util::GetLsb(n0), t1 = util::GetLsb(n1);
    n0 = Masked(Masked(h0, right, n0), t0, cw);
    n1 = Masked(Masked(h1, right, n1), t1, cw);
  }
  const bool an = (a & 1) != 0, t0 = util::GetLsb(n0), t1 = util::GetLsb(n1);
  const int4 h00 = HtHash(prg, hash_key, Clamp(n0)), h01 = HtHash(prg, hash_key, util::SetLsb(n0, true));
  const int4 h10 = HtHash(prg, hash_key, Clamp(n1)), h11 = HtHash(prg, hash_key, util::SetLsb(n1, true));
  const int4 hcw = an ? Clamp(util::Xor(h00, h10)) : Clamp(util::Xor(h01, h11));            // :123-125
  const bool lcw0 = util::GetLsb(h00) ^ util::GetLsb(h10) ^ !an;                             // :132
  const bool lcw1 = util::GetLsb(h01) ^ util::GetLsb(h11) ^ an;                              // :133
  int4 *raw = reinterpret_cast<int4 *>(&cws[in_bits - 1]);
  raw[0] = util::SetLsb(hcw, lcw0);
  raw[1] = int4{lcw1 ? 1 : 0, 0, 0, 0};
  const int4 leaf_cw = util::SetLsb(hcw, an ? lcw1 : lcw0);
  const int4 leaf0 = Masked(an ? h01 : h00, t0, leaf_cw), leaf1 = Masked(an ? h11 : h10, t1, leaf_cw);
  Group v = Group::From(Clamp(b_buf)) + (-Group::From(Clamp(leaf0))) + Group::From(Clamp(leaf1));   // :168-170
  if (util::GetLsb(leaf1)) v = -v;
  ocw = v.Into();
}

// ---- VDPF (verifiable DPF) ----------------------------------------------------------------------------------------------
// The DPF walk without an output entry in `cws` (the output correction word travels separately) plus one XorHash of
// (point, final seed) per party / evaluation (vdpf.cuh:101-180, 195-246).  XorHash: `xh.Hash(tuple<int4, const int4>)`
// -> 4 blocks (hash.cuh).
template <typename XorHash>
FSS_SHIM_HD cuda::std::array<int4, 4> PointHash(XorHash &xh, int4 point, int4 seed) {
  return xh.Hash(cuda::std::tuple<int4, const int4>{point, seed});
}

// Vdpf::Gen: 0, or 1 when both parties end on the same control bit (the caller draws new seeds).  `cs` is written in
// both cases, `ocw` only on success -- as the reference does (vdpf.cuh:167-179).
template <int in_bits, typename Group, typename In, typename Prg, typename XorHash, typename Cw>
FSS_SHIM_HD int VdpfGen(Prg &prg, XorHash &xh, Cw cws[], cuda::std::array<int4, 4> &cs, int4 &ocw, const int4 s0s[2], In a, int4 b_buf) {
  static_assert(sizeof(Cw) == 32, "Vdpf::Cw is {int4 s; bool tr} padded to 32 bytes (vdpf.cuh:77-80)");
  int4 n[2] = {Clamp(s0s[0]), util::SetLsb(s0s[1], true)};  // packed nodes: seed | t
  for (int i = 0; i < in_bits; ++i) {
    const bool right = BitMsbFirst(a, in_bits, i);
    auto g0 = prg.Gen(Clamp(n[0]));
    auto g1 = prg.Gen(Clamp(n[1]));
    const int lose = right ? 0 : 1, keep = right ? 1 : 0;
    const int4 s_cw = Clamp(util::Xor(g0[lose], g1[lose]));
    const bool t_cw[2] = {bool(util::GetLsb(g0[0]) ^ util::GetLsb(g1[0]) ^ right ^ true), bool(util::GetLsb(g0[1]) ^ util::GetLsb(g1[1]) ^ right)};
    const int4 cw_keep = util::SetLsb(s_cw, t_cw[keep]);
    n[0] = Masked(g0[keep], util::GetLsb(n[0]), cw_keep);
    n[1] = Masked(g1[keep], util::GetLsb(n[1]), cw_keep);
    int4 *raw = reinterpret_cast<int4 *>(&cws[i]);  // both halves of the 32-byte slot (the padding is defined: zero)
    raw[0] = util::SetLsb(s_cw, t_cw[0]);
    raw[1] = int4{t_cw[1] ? 1 : 0, 0, 0, 0};
  }
  const int4 point = util::Pack(a);
  const auto h0 = PointHash(xh, point, Clamp(n[0])), h1 = PointHash(xh, point, Clamp(n[1]));
  for (int j = 0; j < 4; ++j) cs[j] = util::Xor(h0[j], h1[j]);
  if (util::GetLsb(n[0]) == util::GetLsb(n[1])) return 1;
  Group v = Group::From(Clamp(b_buf)) + (-Group::From(Clamp(n[0]))) + Group::From(Clamp(n[1]));
  if (util::GetLsb(n[1])) v = -v;
  ocw = v.Into();
  return 0;
}

// Vdpf::Eval: the share through `y`, the corrected per-point hash returned.
template <int in_bits, typename Group, typename In, typename Prg, typename XorHash, typename Cw>
FSS_SHIM_HD cuda::std::array<int4, 4> VdpfEval(Prg &prg, XorHash &xh, bool b, int4 s0, const Cw cws[], const int4 cs[4], int4 ocw, In x,
    int4 &y) {
  int4 node = util::SetLsb(s0, b);
  for (int i = 0; i < in_bits; ++i) {
    const bool right = BitMsbFirst(x, in_bits, i);
    auto g = prg.Gen(Clamp(node));
    const int4 cw = util::SetLsb(cws[i].s, right ? bool(cws[i].tr) : util::GetLsb(cws[i].s));
    node = Masked(right ? g[1] : g[0], util::GetLsb(node), cw);
  }
  const bool t = util::GetLsb(node);
  Group share = Group::From(Clamp(node));
  if (t) share = share + Group::From(ocw);
  if (b) share = -share;
  y = share.Into();
  auto pi = PointHash(xh, util::Pack(x), Clamp(node));
  for (int j = 0; j < 4; ++j) pi[j] = Masked(pi[j], t, cs[j]);
  return pi;
}

// ---- batched kernels for user-defined plugins (instantiated in the user's translation unit) --------------------------
#if defined(__CUDACC__)
// One key per thread; `Scheme` is a scheme object (fss::Dpf / fss::Dcf of this shim) passed by value, like the
// reference passes its scheme objects to its kernels (eval_all_gpu.cuh:486-487).
template <typename Scheme, typename In>
__global__ void EvalKernel(Scheme sch, bool b, const int4 *seeds, const typename Scheme::Cw *cws, const In *xs,
                           int4 *ys, size_t nkeys) {
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nkeys; k += size_t(gridDim.x) * blockDim.x)
    ys[k] = sch.Eval(b, seeds[k], cws + k * Scheme::kNumCw, xs[k]);
}
template <typename Scheme, typename In>
__global__ void GenKernel(Scheme sch, const int4 *s0s, const In *alphas, const int4 *betas,
                          typename Scheme::Cw *cws, size_t nkeys) {
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nkeys; k += size_t(gridDim.x) * blockDim.x) {
    const int4 s[2] = {s0s[2 * k], s0s[2 * k + 1]};
    sch.Gen(cws + k * Scheme::kNumCw, s, alphas[k], betas[k]);
  }
}
// Full domain of one key per launch slice: leaf x by its own point evaluation (in_bits PRG calls per leaf; the built-in
// plugins use the tree kernels of libfssb200.so instead, which expand every node once).
template <typename Scheme, typename In>
__global__ void EvalAllKernel(Scheme sch, bool b, const int4 *seeds, const typename Scheme::Cw *cws, int4 *ys,
                              size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count) {
  const uint64_t total = uint64_t(nkeys) * leaf_count;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k = i / leaf_count, x = leaf_begin + (i - k * leaf_count);
    ys[i] = sch.Eval(b, seeds[k], cws + k * Scheme::kNumCw, In(x));
  }
}

// Runs `fn(stream)` (kernel launches) and turns a launch / execution error into the shim's exception.
inline void CheckLaunch(const char *what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
// Half-Tree: the output correction word is a second per-key array
template <typename Scheme, typename In>
__global__ void HtEvalKernel(Scheme sch, bool b, const int4 *seeds, const typename Scheme::Cw *cws, const int4 *ocws,
                             const In *xs, int4 *ys, size_t nkeys) {
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nkeys; k += size_t(gridDim.x) * blockDim.x)
    ys[k] = sch.Eval(b, seeds[k], cws + k * Scheme::kNumCw, ocws[k], xs[k]);
}
template <typename Scheme, typename In>
__global__ void HtGenKernel(Scheme sch, const int4 *s0s, const In *alphas, const int4 *betas, typename Scheme::Cw *cws,
                            int4 *ocws, size_t nkeys) {
  for (size_t k = size_t(blockIdx.x) * blockDim.x + threadIdx.x; k < nkeys; k += size_t(gridDim.x) * blockDim.x) {
    const int4 s[2] = {s0s[2 * k], s0s[2 * k + 1]};
    int4 ocw;
    sch.Gen(cws + k * Scheme::kNumCw, ocw, s, alphas[k], betas[k]);
    ocws[k] = ocw;
  }
}
template <typename Scheme, typename In>
__global__ void HtEvalAllKernel(Scheme sch, bool b, const int4 *seeds, const typename Scheme::Cw *cws, const int4 *ocws,
                                int4 *ys, size_t nkeys, uint64_t leaf_begin, uint64_t leaf_count) {
  const uint64_t total = uint64_t(nkeys) * leaf_count;
  for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint64_t k = i / leaf_count, x = leaf_begin + (i - k * leaf_count);
    ys[i] = sch.Eval(b, seeds[k], cws + k * Scheme::kNumCw, ocws[k], In(x));
  }
}

inline unsigned GridFor(uint64_t work, unsigned block) {
  const uint64_t want = (work + block - 1) / block;
  return unsigned(want < 1 ? 1 : (want > 148u * 16u ? 148u * 16u : want));  // a few waves of the 148 SMs, grid-stride beyond
}
#endif

}  // namespace fss::b200::generic
