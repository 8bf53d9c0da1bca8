#!/usr/bin/env python
"""Measurement of the SURVEY.md section 8(f) rows (the callers / data formats either side of the hot path) on one B200,
each against the roofline that bounds it -- the same bar bench.py applies to the hot path itself:

  f-1  batched Gen          DPF n=32 / DCF n=64 u127 / Half-Tree n=32    keys/s, AES blocks/s vs the LDS lookup ceiling
  f-2  relayout + level-major eval                                        GB/s vs HBM; evals/s vs the LDS ceiling
  f-3  DCF EvalAll, Half-Tree EvalAll, Grotto expand / EvalAll / Preprocess / lookup      leaves/s vs the LDS ceiling
  f-4  VDPF Eval / EvalAll (+ proof chain)                                evals/s, leaves/s

Every timing: CUDA events on the launching (torch current) stream, warm-up launches first, mean of the timed launches;
buffers are far larger than the 126 MB L2 where the row moves data.  AES blocks per unit are the minimal counts of
the algorithm, 160 shared-memory lookups per block; the ceiling is the conflict-free LDS.32 rate
measured in the same run (fssb200_microbench kind 3).  One JSON object on stdout (and --out).

Measurement tool only: nothing here is on the product path, and nothing under oracle/ is touched.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LOOKUPS = 160


def main():
    import torch

    import fss_b200

    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--keys", type=int, default=1 << 22)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "bench_next.json"))
    ap.add_argument("--sections", default="f1,f2,f3,f4", help="comma-separated subset of f1,f2,f3,f4 (A/B runs)")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    gen = torch.Generator(device=dev).manual_seed(777)

    def rand_i32(shape):
        return torch.randint(-2 ** 31, 2 ** 31, shape, dtype=torch.int64, device=dev, generator=gen).to(torch.int32)

    def rand_i64(shape):
        return torch.randint(-2 ** 63, 2 ** 63 - 1, shape, dtype=torch.int64, device=dev, generator=gen)

    def timed(fn, iters=None, warmup=None):
        iters, warmup = iters or args.iters, warmup or args.warmup
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(iters):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return sum(ms) / len(ms)

    lds_peak = fss_b200.microbench(3, 0)
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        hbm_src = "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        hbm_peak, hbm_src = 6650.0, "fallback 6.65 TB/s"
    res = {"lds32_lookups_per_s": lds_peak, "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src, "rows": {}}
    rows = res["rows"]

    def lsu(units, blocks_per_unit, ms):
        return units * blocks_per_unit * LOOKUPS / (ms * 1e-3) / lds_peak

    def seeds_for(n, wide=False):
        s0s = rand_i32((n, 2, 4))
        betas = rand_i32((n, 4))
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        alphas = rand_i64((n,)) if wide else rand_i32((n,))
        xs = rand_i64((n,)) if wide else rand_i32((n,))
        xs[::16] = alphas[::16]
        return s0s, alphas, betas, xs

    K = args.keys
    sections = set(args.sections.split(","))
    # ---- f-1: Gen ---------------------------------------------------------------------------------------------------
    if "f1" in sections:
        for name, scheme, n, group, wide, blocks, k in (
                ("gen_dpf_n32_bytes", "dpf", 32, "bytes", False, 4 * 32, K),
                ("gen_dcf_n64_u127", "dcf", 64, "u128", True, 8 * 64, K // 2),
                ("gen_halftree_n32_bytes", "halftree", 32, "bytes", False, 2 * 32 + 2, K)):
            ctx = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
            s0s, alphas, betas, xs = seeds_for(k, wide)
            al = ctx.in_tensor(alphas, dev)
            ms = timed(lambda: ctx.gen(s0s, al, betas))
            out_bytes = k * ctx.ncw * 32
            rows[name] = {"keys": k, "ms": ms, "keys_per_s": k / (ms * 1e-3), "aes_blocks_per_key": blocks,
                          "aes_blocks_per_s": k * blocks / (ms * 1e-3), "lsu_roofline_frac": lsu(k, blocks, ms),
                          "hbm_write_gbs": out_bytes / (ms * 1e-3) / 1e9, "hbm_frac": out_bytes / (ms * 1e-3) / 1e9 / hbm_peak}
            del ctx, s0s, betas
            torch.cuda.empty_cache()

    # ---- f-2: relayout + level-major evaluation ---------------------------------------------------------------------------
    if "f2" in sections:
        ctx = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs = seeds_for(K)
        cws = ctx.gen(s0s, alphas, betas)
        seeds0 = s0s[:, 0].contiguous()
        ms = timed(lambda: ctx.relayout(cws))
        lay = ctx.relayout(cws)
        rd = cws.numel() * 4
        wr = sum(t.numel() * 4 for t in lay if t is not None)
        rows["relayout_dpf_n32"] = {"keys": K, "ms": ms, "read_bytes": rd, "write_bytes": wr,
                                    "note": "includes the torch.empty / torch.zeros of the output arrays (allocator hit)",
                                    "hbm_gbs": (rd + wr) / (ms * 1e-3) / 1e9, "hbm_frac": (rd + wr) / (ms * 1e-3) / 1e9 / hbm_peak}
        ys = torch.empty((K, 4), dtype=torch.int32, device=dev)
        ms = timed(lambda: ctx.eval_levelmajor(0, seeds0, lay, xs, out=ys))
        rows["eval_levelmajor_dpf_n32"] = {"keys": K, "ms": ms, "evals_per_s": K / (ms * 1e-3),
                                           "lsu_roofline_frac": lsu(K, 32, ms)}
        ms = timed(lambda: ctx.eval(0, seeds0, cws, xs, out=ys))
        rows["eval_keymajor_dpf_n32"] = {"keys": K, "ms": ms, "evals_per_s": K / (ms * 1e-3),
                                         "lsu_roofline_frac": lsu(K, 32, ms)}
        # packed rows (fssb200_pack_rows / fssb200_eval_packed): 16 B + 1 bit per level, key-major
        kp = min(K, 1 << 21)
        prow = ctx.pack_rows(cws[:kp].cpu()).to(dev)
        ms = timed(lambda: ctx.eval_packed(0, seeds0[:kp], prow, xs[:kp], out=ys[:kp]))
        rows_ok = torch.equal(ctx.eval_packed(0, seeds0[:kp], prow, xs[:kp]), ctx.eval(0, seeds0[:kp], cws[:kp], xs[:kp]))
        rows["eval_packed_dpf_n32"] = {"keys": kp, "ms": ms, "evals_per_s": kp / (ms * 1e-3), "matches_keymajor": bool(rows_ok),
                                       "lsu_roofline_frac": lsu(kp, 32, ms)}
        del cws, lay, ys, s0s, betas, seeds0, prow
        torch.cuda.empty_cache()

    # ---- f-3: DCF EvalAll, Half-Tree EvalAll, Grotto --------------------------------------------------------------------------
    if "f3" in sections:
        for name, scheme, n, group, blocks, k in (
                ("evalall_dcf_n24_u127", "dcf", 24, "u128", 4.0, 16),
                ("evalall_dcf_n24_bytes", "dcf", 24, "bytes", 4.0, 16),
                ("evalall_halftree_n28_bytes", "halftree", 28, "bytes", 1.5, 4),
                ("evalall_dpf_n28_bytes", "dpf", 28, "bytes", 2.0, 4)):
            ctx = fss_b200.Context(scheme, n, group, prg="aes128_mmo")
            s0s, alphas, betas, xs = seeds_for(k)
            alphas = alphas & ((1 << n) - 1)
            r = ctx.gen(s0s, alphas, betas)
            cws, ocws = r if scheme == "halftree" else (r, None)
            seeds0 = s0s[:, 0].contiguous()
            out = torch.empty((k, 1 << n, 4), dtype=torch.int32, device=dev)
            ms = timed(lambda: ctx.eval_all(0, seeds0, cws, ocws, out=out), max(3, args.iters // 2), 2)
            leaves = k * (1 << n)
            rows[name] = {"keys": k, "in_bits": n, "ms": ms, "leaves_per_s": leaves / (ms * 1e-3), "aes_blocks_per_leaf": blocks,
                          "lsu_roofline_frac": lsu(leaves, blocks, ms), "hbm_write_gbs": leaves * 16 / (ms * 1e-3) / 1e9}
            del out, cws, ctx
            torch.cuda.empty_cache()
        n, k = 26, 16
        ctx = fss_b200.Context("grotto", n, prg="aes128_mmo")
        s0s, alphas, betas, xs = seeds_for(k)
        alphas = alphas & ((1 << n) - 1)
        cws = ctx.gen(s0s, alphas)
        seeds0 = s0s[:, 0].contiguous()
        leaves = k * (1 << n)
        ms = timed(lambda: ctx.grotto_expand(0, seeds0, cws), 5, 2)
        rows["grotto_expand_n26"] = {"keys": k, "in_bits": n, "ms": ms, "leaves_per_s": leaves / (ms * 1e-3),
                                     "aes_blocks_per_leaf": 2.0, "lsu_roofline_frac": lsu(leaves, 2.0, ms)}
        ms = timed(lambda: ctx.eval_all(0, seeds0, cws), 5, 2)
        rows["grotto_evalall_n26"] = {"keys": k, "in_bits": n, "ms": ms, "leaves_per_s": leaves / (ms * 1e-3),
                                      "note": "leaf-bit expansion + in-place prefix-XOR scan (grotto_dcf.cuh:151-163)",
                                      "lsu_roofline_frac": lsu(leaves, 2.0, ms)}
        ms = timed(lambda: ctx.grotto_preprocess(0, seeds0, cws), 5, 2)
        rows["grotto_preprocess_n26"] = {"keys": k, "in_bits": n, "ms": ms, "leaves_per_s": leaves / (ms * 1e-3),
                                         "note": "expansion + heap-ordered parity tree of 2N-1 bytes (grotto_dcf.cuh:94-104)",
                                         "lsu_roofline_frac": lsu(leaves, 2.0, ms)}
        pt = ctx.grotto_preprocess(0, seeds0, cws)
        q = 1 << 22
        qx = (rand_i32((k, q // k)) & ((1 << n) - 1))
        # one lookup call evaluates one x per key row; time a batch of rows by repeating the parity tree rows
        ms = timed(lambda: ctx.grotto_lookup(pt, qx[:, 0].contiguous()), 5, 2)
        rows["grotto_lookup_n26"] = {"keys": k, "ms": ms, "note": "GrottoDcf::Eval (tree lookup, grotto_dcf.cuh:116-135), one x per key; "
                                     "launch-latency bound at this batch size"}
        del pt, cws, ctx
        torch.cuda.empty_cache()

    # ---- f-4: VDPF ----------------------------------------------------------------------------------------------------------
    if "f4" in sections:
        k = min(K, 1 << 21)
        ctx = fss_b200.Context("vdpf", 32, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs = seeds_for(k)
        vcws, vcs, vocws, _st = ctx.vdpf_gen(s0s, alphas, betas)
        seeds0 = s0s[:, 0].contiguous()
        ms = timed(lambda: ctx.vdpf_eval(0, seeds0, vcws, vcs, vocws, xs))
        rows["vdpf_eval_n32"] = {"keys": k, "ms": ms, "evals_per_s": k / (ms * 1e-3),
                                 "note": "n AES blocks for the walk + 2 Blake3 compressions for the proof share per evaluation; "
                                         "the fraction counts the AES lookups only",
                                 "lsu_roofline_frac_aes_only": lsu(k, 32, ms)}
        ms = timed(lambda: ctx.vdpf_gen(s0s, alphas, betas))
        rows["vdpf_gen_n32"] = {"keys": k, "ms": ms, "keys_per_s": k / (ms * 1e-3)}
        del vcws, vcs, vocws, ctx
        torch.cuda.empty_cache()
        n, k = 24, 8
        ctx = fss_b200.Context("vdpf", n, "bytes", prg="aes128_mmo")
        s0s, alphas, betas, xs = seeds_for(k)
        alphas = alphas & ((1 << n) - 1)
        vcws, vcs, vocws, _st = ctx.vdpf_gen(s0s, alphas, betas)
        seeds0 = s0s[:, 0].contiguous()
        ms = timed(lambda: ctx.vdpf_eval_all(0, seeds0, vcws, vcs, vocws), 3, 2)
        leaves = k * (1 << n)
        rows["vdpf_evalall_n24"] = {"keys": k, "in_bits": n, "ms": ms, "leaves_per_s": leaves / (ms * 1e-3),
                                    "note": "tree expansion (2 AES blocks per leaf) + per-leaf Blake3 + the sequential proof chain "
                                            "(vdpf.cuh:294-342); includes the torch.empty of the outputs"}

    txt = json.dumps(res)
    print(txt, flush=True)
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        f.write(txt + "\n")


if __name__ == "__main__":
    main()
