#!/usr/bin/env python
"""A/B sweep of the host-buffer entry point (fssb200_eval_host, C2: 2^22 DPF keys, n = 32, pinned host buffers in the
reference layout): pipeline mode x chunk size x ring slots x store kind, one subprocess per worker-thread count (the
crew is sized once per process).  Measurement tool only; writes one JSON object per configuration to --out."""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(args):
    import torch

    import fss_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1)
    k = args.keys
    r = lambda s: torch.randint(-2 ** 31, 2 ** 31, s, dtype=torch.int64, device=dev, generator=g).to(torch.int32)  # noqa: E731
    ctx = fss_b200.Context("dpf", 32, "bytes", prg="aes128_mmo")
    s0s, betas, alphas, xs = r((k, 2, 4)), r((k, 4)), r((k,)), r((k,))
    s0s[:, :, 3] &= ~1
    betas[:, 3] &= ~1
    cws = ctx.gen(s0s, alphas, betas)
    seeds0 = s0s[:, 0].contiguous()
    want = ctx.eval(0, seeds0, cws, xs).cpu()
    h_seeds, h_cws, h_xs = seeds0.cpu().pin_memory(), cws.cpu().pin_memory(), xs.cpu().pin_memory()
    h_ys = torch.empty((k, 4), dtype=torch.int32).pin_memory()
    del cws, s0s
    rows = []
    for cfg in json.loads(args.configs):
        for key in ("FSSB200_PIPE_CHUNK_BITS", "FSSB200_PIPE_SLOTS", "FSSB200_PACK_NT", "FSSB200_PIPE_PIECE_BITS", "FSSB200_PIPE_LOW_WATER_MB"):
            os.environ.pop(key, None)
        for key, v in cfg.get("env", {}).items():
            os.environ[key] = str(v)
        ctx.set_host_mode(cfg.get("mode", 0))
        for _ in range(2):
            ctx.eval(0, h_seeds, h_cws, h_xs, out=h_ys)
        times = []
        for _ in range(args.steps):
            t0 = time.perf_counter()
            ctx.eval(0, h_seeds, h_cws, h_xs, out=h_ys)
            times.append(time.perf_counter() - t0)
        ok = bool(torch.equal(h_ys, want))
        st = ctx.host_stats()
        times.sort()
        rows.append({"threads_env": os.environ.get("FSSB200_PACK_THREADS"), **cfg, "ms_median": times[len(times) // 2] * 1e3,
                     "ms_min": times[0] * 1e3, "mevals_per_s": k / times[len(times) // 2] / 1e6, "ok": ok, **st})
        print(json.dumps(rows[-1]), flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--configs", default="")
    ap.add_argument("--keys", type=int, default=1 << 22)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--threads", default="0,4,8,16", help="FSSB200_PACK_THREADS values (0 = library default for the box)")
    ap.add_argument("--tune", default="single")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_sweep.jsonl"))
    args = ap.parse_args()
    if args.child:
        child(args)
        return
    base = [{"mode": 0}, {"mode": 1}, {"mode": 2}]

    def cfg(piece, slots, nt, chunk=16, mode=0, low=None):
        env = {"FSSB200_PACK_NT": nt, "FSSB200_PIPE_PIECE_BITS": piece, "FSSB200_PIPE_SLOTS": slots,
               "FSSB200_PIPE_CHUNK_BITS": chunk}
        if low is not None:
            env["FSSB200_PIPE_LOW_WATER_MB"] = low
        return {"mode": mode, "env": env}
    tune = [cfg(14, 4, 0), cfg(14, 5, 0), cfg(14, 6, 0), cfg(13, 6, 0), cfg(13, 8, 0), cfg(13, 10, 0), cfg(12, 8, 0), cfg(12, 12, 0),
            cfg(12, 16, 0), cfg(14, 4, 0, 17), cfg(14, 4, 0, 15), cfg(14, 4, 0, low=8), cfg(14, 4, 0, low=32), cfg(14, 4, 0, mode=2),
            cfg(14, 8, 1), cfg(15, 6, 1), cfg(16, 6, 1)]
    if args.tune == "multi":   # several ranks share the host: small cache-resident rings vs streaming rings
        tune = [cfg(13, 4, 0), cfg(12, 4, 0), cfg(12, 8, 0), cfg(11, 8, 0), cfg(14, 4, 0), cfg(14, 6, 1), cfg(16, 6, 1), cfg(12, 16, 1)]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        for i, t in enumerate(args.threads.split(",")):
            env = dict(os.environ)
            if t != "0":
                env["FSSB200_PACK_THREADS"] = t
            cfgs = base + (tune if i == 0 and args.tune != "none" else [])
            r = subprocess.run([sys.executable, __file__, "--child", "--configs", json.dumps(cfgs), "--keys", str(args.keys),
                                "--steps", str(args.steps)], env=env, capture_output=True, text=True)
            f.write(r.stdout)
            f.flush()
            sys.stdout.write(r.stdout)
            if r.returncode != 0:
                sys.stdout.write(r.stderr[-3000:])


if __name__ == "__main__":
    main()
