#!/usr/bin/env python
"""A/B of the DCF full-domain kernel geometry: 512 threads x dfs <= 6 (default) vs 256 threads x dfs <= 8
(FSSB200_DCF_ALL_THREADS=256, round 1).  One subprocess per geometry; n = 24, 16 keys, u127 and Bytes; CUDA events,
fraction of the LDS lookup ceiling measured in the same process; both outputs hashed (they must be identical)."""
import hashlib
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child():
    import torch

    import fss_b200
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(3)
    r = lambda s: torch.randint(-2 ** 31, 2 ** 31, s, dtype=torch.int64, device=dev, generator=g).to(torch.int32)  # noqa: E731
    lds = fss_b200.microbench(3, 0)
    out_rows = {}
    for group in ("u128", "bytes"):
        n, k = 24, 16
        ctx = fss_b200.Context("dcf", n, group, prg="aes128_mmo")
        s0s, betas = r((k, 2, 4)), r((k, 4))
        s0s[:, :, 3] &= ~1
        betas[:, 3] &= ~1
        alphas = r((k,)) & ((1 << n) - 1)
        cws = ctx.gen(s0s, alphas, betas)
        seeds0 = s0s[:, 0].contiguous()
        out = torch.empty((k, 1 << n, 4), dtype=torch.int32, device=dev)
        for _ in range(2):
            ctx.eval_all(0, seeds0, cws, out=out)
        torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ctx.eval_all(0, seeds0, cws, out=out)
            b.record()
            torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        t = sum(ms) / len(ms)
        leaves = k * (1 << n)
        out_rows[group] = {"ms": t, "gleaves_per_s": leaves / t / 1e6, "lsu_frac": leaves * 4 * 160 / (t * 1e-3) / lds,
                           "granule": ctx.granule(), "sha256_first_key": hashlib.sha256(out[0].cpu().numpy().tobytes()).hexdigest()}
    print(json.dumps({"threads": os.environ.get("FSSB200_DCF_ALL_THREADS", "512"), **out_rows}), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child()
    else:
        for t in ("512", "256"):
            env = dict(os.environ, FSSB200_DCF_ALL_THREADS=t)
            r = subprocess.run([sys.executable, __file__, "--child"], env=env, capture_output=True, text=True)
            sys.stdout.write(r.stdout)
            if r.returncode:
                sys.stdout.write(r.stderr[-2000:])
