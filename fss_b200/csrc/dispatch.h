// SPDX-License-Identifier: Apache-2.0
//
// dispatch.h -- run-time dispatch from (scheme, group kind, PRG) to the kernel instantiations.
// kernels_inst.cu is compiled once per (FSS_INST_KIND, FSS_INST_PRG, FSS_INST_SCHEME) so that the
// template instantiations build in parallel; each object file defines one `*_launcher_*` getter.
#pragma once
#include "kernels.cuh"

namespace fssb200 {

struct LaunchCfg {
  dim3 grid, block;
  size_t smem;
  cudaStream_t stream;
};

typedef cudaError_t (*point_launch_fn)(const KParams &, const PointArgs &, const LaunchCfg &);
typedef cudaError_t (*gen_launch_fn)(const KParams &, const GenArgs &, const LaunchCfg &);
typedef cudaError_t (*evalall_launch_fn)(const KParams &, const EvalAllArgs &, const LaunchCfg &);
typedef cudaError_t (*prg_launch_fn)(const KParams &, const blk *, blk *, uint64_t, const LaunchCfg &);

// Group kinds that have instantiations (kGrpU8/kGrpU16 run as kGrpU32 with a value mask).
inline bool grp_kind_instantiated(int gk) {
  return gk == kGrpBytes || gk == kGrpU32 || gk == kGrpU64 || gk == kGrpU127 || gk == kGrpU32Mod ||
      gk == kGrpU64Mod || gk == kGrpU128Mod;
}

// scheme in {DPF, DCF, HALFTREE, VDPF}; returns nullptr when not instantiated.
// mode: see PointMode in kernels.cuh (0/1 staged key-major, 2 level-major, 3 direct key-major)
point_launch_fn get_point_launcher(int scheme, int gk, int prg, int mode);
// out_mode: 0 = direct key-major stores, 1 = tiles written by the TMA unit (CwTileOut)
gen_launch_fn get_gen_launcher(int scheme, int gk, int prg, int out_mode);
// mode 0 = DPF leaves, 1 = Half-Tree leaves, 2 = Grotto leaf bits (gk ignored), 3 = DCF leaves,
// 4 = VDPF packed leaves (gk ignored)
evalall_launch_fn get_evalall_launcher(int mode, int gk, int prg);
prg_launch_fn get_prg_launcher(int prg, int mul);

#define FSS_DECL_POINT(PRGNAME, SCHNAME) \
  point_launch_fn point_launcher_##PRGNAME##_##SCHNAME(int gk, int mode);
#define FSS_DECL_GEN(PRGNAME, SCHNAME) gen_launch_fn gen_launcher_##PRGNAME##_##SCHNAME(int gk, int out_mode);
#define FSS_DECL_EVALALL(PRGNAME, MODENAME) evalall_launch_fn evalall_launcher_##PRGNAME##_##MODENAME(int gk);
FSS_DECL_POINT(aes, dpf) FSS_DECL_POINT(aes, dcf) FSS_DECL_POINT(aes, ht) FSS_DECL_POINT(aes, vdpf) FSS_DECL_POINT(aes, grotto)
FSS_DECL_POINT(chacha, dpf) FSS_DECL_POINT(chacha, dcf) FSS_DECL_POINT(chacha, ht) FSS_DECL_POINT(chacha, vdpf)
FSS_DECL_POINT(chacha, grotto)
FSS_DECL_GEN(aes, dpf) FSS_DECL_GEN(aes, dcf) FSS_DECL_GEN(aes, ht) FSS_DECL_GEN(aes, vdpf)
FSS_DECL_GEN(chacha, dpf) FSS_DECL_GEN(chacha, dcf) FSS_DECL_GEN(chacha, ht) FSS_DECL_GEN(chacha, vdpf)
FSS_DECL_EVALALL(aes, dpf) FSS_DECL_EVALALL(aes, ht) FSS_DECL_EVALALL(aes, grotto) FSS_DECL_EVALALL(aes, dcf)
FSS_DECL_EVALALL(aes, vdpf)
FSS_DECL_EVALALL(chacha, dpf) FSS_DECL_EVALALL(chacha, ht) FSS_DECL_EVALALL(chacha, grotto) FSS_DECL_EVALALL(chacha, dcf)
FSS_DECL_EVALALL(chacha, vdpf)
prg_launch_fn prg_launcher_aes(int mul);
prg_launch_fn prg_launcher_chacha(int mul);

}  // namespace fssb200
