// SPDX-License-Identifier: Apache-2.0
//
// kernels_inst.cu -- explicit kernel instantiations + launchers.  Compiled once per
//   -DFSS_INST_KIND={1 point, 2 gen, 3 evalall, 4 prg}
//   -DFSS_INST_PRG={0 aes, 1 chacha}
//   -DFSS_INST_SCHEME={0 dpf, 1 dcf, 2 halftree, 3 grotto, 4 vdpf}   (ignored for kind 4)
// (see fss_b200/csrc/Makefile) so the ~160 instantiations build in parallel.
#include "dispatch.h"

namespace fssb200 {

template <class Kern, class... Args>
static cudaError_t launch_kernel(Kern kern, const LaunchCfg &c, const Args &...args) {
  if (c.smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(c.smem));
    if (e != cudaSuccess) return e;
  }
  kern<<<c.grid, c.block, c.smem, c.stream>>>(args...);
  return cudaGetLastError();
}

#if FSS_INST_PRG == 0
#define PRGNAME aes
constexpr int kInstPrg = kPrgAes;
#else
#define PRGNAME chacha
constexpr int kInstPrg = kPrgChaCha;
#endif

#if FSS_INST_SCHEME == 0
#define SCHNAME dpf
#elif FSS_INST_SCHEME == 1
#define SCHNAME dcf
#elif FSS_INST_SCHEME == 2
#define SCHNAME ht
#elif FSS_INST_SCHEME == 3
#define SCHNAME grotto
#else
#define SCHNAME vdpf
#endif

#define CAT3_(a, b, c) a##b##_##c
#define CAT3(a, b, c) CAT3_(a, b, c)

#define FOR_EACH_GK(X) X(kGrpBytes) X(kGrpU32) X(kGrpU64) X(kGrpU127) X(kGrpU32Mod) X(kGrpU64Mod) X(kGrpU128Mod)

#if FSS_INST_KIND == 1 && FSS_INST_SCHEME == 3
// Grotto point walk: defined over group::Bytes only; key-major Cw through the TMA unit (modes 4 / 5), or direct loads (3)
template <int MODE>
static cudaError_t point_launch_grotto(const KParams &P, const PointArgs &A, const LaunchCfg &c) {
  return launch_kernel(point_kernel<FSSB200_SCHEME_GROTTO, kGrpBytes, kInstPrg, MODE>, c, P, A);
}
point_launch_fn CAT3(point_launcher_, PRGNAME, SCHNAME)(int, int mode) {
  return mode == 3 ? &point_launch_grotto<3> : (mode == 4 ? &point_launch_grotto<4> : (mode == 5 ? &point_launch_grotto<5> : nullptr));
}
#elif FSS_INST_KIND == 1
template <int G, int MODE>
static cudaError_t point_launch(const KParams &P, const PointArgs &A, const LaunchCfg &c) {
  return launch_kernel(point_kernel<FSS_INST_SCHEME, G, kInstPrg, MODE>, c, P, A);
}
point_launch_fn CAT3(point_launcher_, PRGNAME, SCHNAME)(int gk, int mode) {
  switch (gk) {
// mode 6 (packed rows): only for the schemes whose Cw carries padding (DPF, Half-Tree, VDPF); DCF rows are all payload
#if FSS_INST_SCHEME != 1
#define FSS_CASE_PACKED(GK) case 6: return &point_launch<GK, 6>;
#else
#define FSS_CASE_PACKED(GK)
#endif
#define X(GK)                                                   \
  case GK:                                                      \
    switch (mode) {                                             \
      case 0: return &point_launch<GK, 0>;                      \
      case 1: return &point_launch<GK, 1>;                      \
      case 2: return &point_launch<GK, 2>;                      \
      case 3: return &point_launch<GK, 3>;                      \
      case 4: return &point_launch<GK, 4>;                      \
      case 5: return &point_launch<GK, 5>;                      \
      case 7: return &point_launch<GK, 7>;                      \
      FSS_CASE_PACKED(GK)                                       \
    }                                                           \
    return nullptr;
    FOR_EACH_GK(X)
#undef X
  }
  return nullptr;
}
#elif FSS_INST_KIND == 2
template <int G, int OUT>
static cudaError_t gen_launch(const KParams &P, const GenArgs &A, const LaunchCfg &c) {
  return launch_kernel(gen_kernel<FSS_INST_SCHEME, G, kInstPrg, OUT>, c, P, A);
}
gen_launch_fn CAT3(gen_launcher_, PRGNAME, SCHNAME)(int gk, int out_mode) {
  switch (gk) {
#define X(GK) \
  case GK: return out_mode ? &gen_launch<GK, 1> : &gen_launch<GK, 0>;
    FOR_EACH_GK(X)
#undef X
  }
  return nullptr;
}
#elif FSS_INST_KIND == 3
#if FSS_INST_SCHEME == 3
static cudaError_t evalall_launch_grotto(const KParams &P, const EvalAllArgs &A, const LaunchCfg &c) {
  return launch_kernel(evalall_kernel<2, kGrpBytes, kInstPrg>, c, P, A);
}
evalall_launch_fn CAT3(evalall_launcher_, PRGNAME, SCHNAME)(int) { return &evalall_launch_grotto; }
#elif FSS_INST_SCHEME == 4
static cudaError_t evalall_launch_vdpf(const KParams &P, const EvalAllArgs &A, const LaunchCfg &c) {
  return launch_kernel(evalall_kernel<4, kGrpBytes, kInstPrg>, c, P, A);
}
evalall_launch_fn CAT3(evalall_launcher_, PRGNAME, SCHNAME)(int) { return &evalall_launch_vdpf; }
#elif FSS_INST_SCHEME == 1
template <int G>
static cudaError_t evalall_launch_dcf(const KParams &P, const EvalAllArgs &A, const LaunchCfg &c) {
  // geometry by the CTA size api.cu chose: 512 threads (default) or 256 (FSSB200_DCF_ALL_THREADS=256, A/B runs)
  if (c.block.x == 512u) return launch_kernel(dcf_evalall_kernel<G, kInstPrg, 9>, c, P, A);
  return launch_kernel(dcf_evalall_kernel<G, kInstPrg, 8>, c, P, A);
}
evalall_launch_fn CAT3(evalall_launcher_, PRGNAME, SCHNAME)(int gk) {
  switch (gk) {
#define X(GK) \
  case GK: return &evalall_launch_dcf<GK>;
    FOR_EACH_GK(X)
#undef X
  }
  return nullptr;
}
#else
constexpr int kMode = FSS_INST_SCHEME == 2 ? 1 : 0;
template <int G>
static cudaError_t evalall_launch(const KParams &P, const EvalAllArgs &A, const LaunchCfg &c) {
  return launch_kernel(evalall_kernel<kMode, G, kInstPrg>, c, P, A);
}
evalall_launch_fn CAT3(evalall_launcher_, PRGNAME, SCHNAME)(int gk) {
  switch (gk) {
#define X(GK) \
  case GK: return &evalall_launch<GK>;
    FOR_EACH_GK(X)
#undef X
  }
  return nullptr;
}
#endif
#elif FSS_INST_KIND == 4
template <int MUL>
static cudaError_t prg_launch(const KParams &P, const blk *seeds, blk *out, uint64_t n, const LaunchCfg &c) {
  return launch_kernel(prg_kernel<kInstPrg, MUL>, c, P, seeds, out, n);
}
#define CAT2_(a, b) a##b
#define CAT2(a, b) CAT2_(a, b)
prg_launch_fn CAT2(prg_launcher_, PRGNAME)(int mul) {
  return mul == 1 ? &prg_launch<1> : (mul == 2 ? &prg_launch<2> : (mul == 4 ? &prg_launch<4> : nullptr));
}
#else
#error "FSS_INST_KIND must be 1..4"
#endif

}  // namespace fssb200
